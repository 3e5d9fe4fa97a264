// fp32 SIMT shared-MLP kernels: the "exact" precision mode of the point-wise MLPs and the FC heads.
//
// Replaces nn.Conv1d(k=1) / nn.Linear + folded eval BatchNorm + ReLU of the reference
// (tools/static_model.py:279-294,330-338; tools/dynamic_model.py:241-248,278-285,307-311) and the
// max over points (tools/static_model.py:284,334).  Accumulation is plain fp32 FMA.
#include "common.cuh"
#include "../../include/al3d.h"

namespace al3d {

// ---------------------------------------------------------------------------------------------
// First layer: K = C <= 8 input channels read straight from the strided (bs,C,n) tensor.
// One thread per (point, 4 output channels); weights + bias staged in shared memory.
// ---------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256)
pointwise_first_kernel(const float *__restrict__ x, int64_t sb, int64_t sc, int64_t sp, int64_t M, int n,
                       const float *__restrict__ w, const float *__restrict__ bias, int cout, int act,
                       float *__restrict__ y)
{
    extern __shared__ float sw[];            // cout*C weights, then cout biases
    float *sbias = sw + cout * C;
    for (int i = threadIdx.x; i < cout * C; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < cout; i += blockDim.x) sbias[i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const int quads = cout >> 2;             // cout is a multiple of 4 (checked by the launcher)
    const int64_t total = M * quads;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = t / quads;
        const int q = (int)(t - m * quads);
        const int64_t b = m / n, p = m - b * n;
        const float *px = x + b * sb + p * sp;
        float v[C];
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = __ldg(px + c * sc);
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float *wr = sw + (q * 4 + j) * C;
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) acc = fmaf(v[c], wr[c], acc);
            acc += sbias[q * 4 + j];
            o[j] = (act & AL3D_ACT_RELU) ? fmaxf(acc, 0.f) : acc;
        }
        *reinterpret_cast<float4 *>(y + m * cout + q * 4) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// ---------------------------------------------------------------------------------------------
// Generic NT GEMM with fused bias / per-group bias / ReLU and an optional max-pool epilogue.
// 128x64 tile, BK = 16, 256 threads, 8x4 outputs per thread.
// ---------------------------------------------------------------------------------------------
constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4;

__global__ void __launch_bounds__(256)
linear_f32_kernel(const float *__restrict__ a, int64_t lda, int64_t M, int K, const float *__restrict__ w, int64_t ldw,
                  const float *__restrict__ bias, const float *__restrict__ rowbias, int64_t rows_per_group,
                  int cout, int act, float *__restrict__ y, int64_t ldy, float *__restrict__ y_max)
{
    __shared__ float As[BK][BM + 4];
    __shared__ float Ws[BK][BN + 4];
    __shared__ int red[BN];

    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int tr = tid / 16;      // 0..15 -> rows tr*8 .. +7
    const int tc = tid % 16;      // 0..15 -> cols tc*4 .. +3

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const bool vec_ok = ((lda & 3) == 0) && ((ldw & 3) == 0) && ((K & 3) == 0) &&
                        ((reinterpret_cast<uintptr_t>(a) & 15) == 0) && ((reinterpret_cast<uintptr_t>(w) & 15) == 0);

    for (int k0 = 0; k0 < K; k0 += BK) {
        // A tile: 128 rows x 16 k = 512 float4 -> 2 per thread ; W tile: 64 x 16 = 256 float4 -> 1 per thread
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int f = tid + it * 256;
            const int r = f >> 2, kq = (f & 3) * 4;
            const int64_t m = m0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < M) {
                const float *src = a + m * lda + k0 + kq;
                if (vec_ok && k0 + kq + 3 < K) v = *reinterpret_cast<const float4 *>(src);
                else {
                    if (k0 + kq + 0 < K) v.x = src[0];
                    if (k0 + kq + 1 < K) v.y = src[1];
                    if (k0 + kq + 2 < K) v.z = src[2];
                    if (k0 + kq + 3 < K) v.w = src[3];
                }
            }
            As[kq + 0][r] = v.x; As[kq + 1][r] = v.y; As[kq + 2][r] = v.z; As[kq + 3][r] = v.w;
        }
        {
            const int r = tid >> 2, kq = (tid & 3) * 4;
            const int o = n0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (o < cout) {
                const float *src = w + (int64_t)o * ldw + k0 + kq;
                if (vec_ok && k0 + kq + 3 < K) v = *reinterpret_cast<const float4 *>(src);
                else {
                    if (k0 + kq + 0 < K) v.x = src[0];
                    if (k0 + kq + 1 < K) v.y = src[1];
                    if (k0 + kq + 2 < K) v.z = src[2];
                    if (k0 + kq + 3 < K) v.w = src[3];
                }
            }
            Ws[kq + 0][r] = v.x; Ws[kq + 1][r] = v.y; Ws[kq + 2][r] = v.z; Ws[kq + 3][r] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float av[TM], wv[TN];
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[k][tr * 8]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[k][tr * 8 + 4]);
            av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w;
            av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
            const float4 w0 = *reinterpret_cast<const float4 *>(&Ws[k][tc * 4]);
            wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w;
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue ----
    const int64_t g_first = m0 / rows_per_group;
    const int64_t m_last = (m0 + BM - 1 < M ? m0 + BM - 1 : M - 1);
    const bool one_group = (m_last / rows_per_group) == g_first;
    if (y_max != nullptr && one_group) {
        if (tid < BN) red[tid] = 0;        // results are >= 0 after ReLU: int order == float order
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        const int o = n0 + tc * 4 + j;
        if (o >= cout) continue;
        const float bj = bias ? bias[o] : 0.f;
        float colmax = 0.f;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int64_t m = m0 + tr * 8 + i;
            if (m >= M) continue;
            float v = acc[i][j] + bj;
            const int64_t g = m / rows_per_group;
            if (rowbias) v += rowbias[g * cout + o];
            if (act & AL3D_ACT_RELU) v = fmaxf(v, 0.f);
            if (y_max == nullptr) y[m * ldy + o] = (act & AL3D_ACT_ACCUMULATE) ? y[m * ldy + o] + v : v;
            else if (one_group) colmax = fmaxf(colmax, v);
            else atomicMax(reinterpret_cast<int *>(y_max + g * cout + o), __float_as_int(v));
        }
        if (y_max != nullptr && one_group) atomicMax(&red[tc * 4 + j], __float_as_int(colmax));
    }
    if (y_max != nullptr && one_group) {
        __syncthreads();
        if (tid < BN && n0 + tid < cout)
            atomicMax(reinterpret_cast<int *>(y_max + g_first * cout + n0 + tid), red[tid]);
    }
}

}  // namespace al3d

using namespace al3d;

extern "C" int al3d_pointwise_first_f32(const float *x, int64_t sb, int64_t sc, int64_t sp, int bs, int C, int n,
                                        const float *w, const float *bias, int cout, int act, float *y, void *stream)
{
    AL3D_CHECK_ARG(x && w && y, "al3d_pointwise_first_f32: null pointer");
    AL3D_CHECK_ARG(C >= 1 && C <= 8, "al3d_pointwise_first_f32: C=%d not in [1,8]", C);
    AL3D_CHECK_ARG(cout > 0 && (cout & 3) == 0, "al3d_pointwise_first_f32: cout=%d must be a positive multiple of 4", cout);
    AL3D_CHECK_ARG(bs >= 0 && n >= 0, "al3d_pointwise_first_f32: negative size");
    const int64_t M = (int64_t)bs * n;
    if (M == 0) return 0;
    const int64_t total = M * (cout / 4);
    const int grid = (int)std::min<int64_t>(ceil_div(total, 256), 148 * 16);
    const size_t smem = (size_t)(cout * C + cout) * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
#define AL3D_PF(CC) case CC: pointwise_first_kernel<CC><<<grid, 256, smem, st>>>(x, sb, sc, sp, M, n, w, bias, cout, act, y); break;
    switch (C) { AL3D_PF(1) AL3D_PF(2) AL3D_PF(3) AL3D_PF(4) AL3D_PF(5) AL3D_PF(6) AL3D_PF(7) AL3D_PF(8) }
#undef AL3D_PF
    AL3D_CHECK_LAUNCH("pointwise_first_kernel");
    return 0;
}

extern "C" int al3d_linear_f32(const float *a, int64_t lda, int64_t M, int K, const float *w, int64_t ldw,
                               const float *bias, const float *rowbias, int64_t rows_per_group, int cout, int act,
                               float *y, int64_t ldy, float *y_max, void *stream)
{
    AL3D_CHECK_ARG(a && w, "al3d_linear_f32: null input");
    AL3D_CHECK_ARG(y || y_max, "al3d_linear_f32: no output");
    AL3D_CHECK_ARG(K > 0 && cout > 0 && M >= 0, "al3d_linear_f32: bad shape M=%lld K=%d cout=%d", (long long)M, K, cout);
    AL3D_CHECK_ARG(lda >= K && ldw >= K, "al3d_linear_f32: leading dimension smaller than K");
    AL3D_CHECK_ARG(!y_max || act == AL3D_ACT_RELU, "al3d_linear_f32: max-pool epilogue requires ReLU (and no accumulate)");
    if (rows_per_group <= 0) rows_per_group = M > 0 ? M : 1;
    if (M == 0) return 0;
    dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(cout, BN));
    AL3D_CHECK_ARG(ceil_div(M, BM) < (1ll << 31), "al3d_linear_f32: M too large");
    linear_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, lda, M, K, w, ldw, bias, rowbias, rows_per_group, cout, act,
                                                               y, ldy, y_max);
    AL3D_CHECK_LAUNCH("linear_f32_kernel");
    return 0;
}
