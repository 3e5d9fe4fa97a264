// Training-step kernels of the auto-label models (BASELINE.json configs[4]; reference loop tools/static_train.py:65-90,
// tools/dynamic_train.py:37-133): train-mode BatchNorm1d (batch statistics + running-stat update), ReLU / Dropout,
// max-pool with arg-max routing, weight gradients, the fused loss backward (tools/static_model.py:341-425), seg accuracy
// (tools/static_train.py:128-129) and a fused Adam step over one flat parameter bucket (torch.optim.Adam semantics,
// tools/static_train.py:220).
//
// In training the shared MLPs cannot be fused across layers: every BatchNorm needs the statistics of the whole batch
// (bs*n rows) of its input before anything downstream can run.  Each layer is therefore  GEMM (al3d_linear_f32, rows =
// points) -> column statistics -> normalise + ReLU, with every activation kept in HBM (row-major (M, C), M = bs*n) for
// the backward pass, and the backward mirrors it: BN/ReLU backward (two column reductions + one elementwise pass),
// dgrad = dY.W (the same NT GEMM on W^T), wgrad = dY^T.X (split over row slabs, deterministic two-stage reduction).
// All arithmetic is fp32 with fp64 finalisation of the long column sums; every reduction has a fixed order, so a step
// is bit-reproducible.
#include <algorithm>
#include "common.cuh"
#include "../../include/al3d.h"

namespace al3d {
namespace train {

constexpr int kColTile = 32;           // columns per CTA of the column-reduction kernels (one per lane)
constexpr int kColWarps = 8;

// ---------------------------------------------------------------------------------------------------------------
// Column reductions over a row-major (M, C) matrix.  grid = (ceil(C/32), n_slabs); slab s covers rows
// [s*rows_per_slab, (s+1)*rows_per_slab).  MODE 0: sum(x), sum(x^2)  (BatchNorm statistics)
//                                            MODE 1: sum(g), sum(g * xhat) with g = dz * relu'(.) * drop  (BN backward)
//                                            MODE 2: sum(x)  (bias gradients, per-group sums)
// out: (n_slabs, 2, C) for modes 0 / 1, (n_slabs, C) for mode 2.
// ---------------------------------------------------------------------------------------------------------------
struct BnCtx {
    const float *mean, *rstd, *gamma, *beta;     // per channel
    const float *drop; int64_t drop_sg, drop_sc, drop_sr; int64_t rows_per_group;   // optional dropout multiplier, element strides
    int relu;
};

__device__ __forceinline__ float bn_drop(const BnCtx &c, int64_t m, int col)
{
    if (!c.drop) return 1.f;
    const int64_t g = m / c.rows_per_group, r = m - g * c.rows_per_group;
    return __ldg(c.drop + g * c.drop_sg + (int64_t)col * c.drop_sc + r * c.drop_sr);
}

template <int MODE>
__global__ void __launch_bounds__(kColTile * kColWarps)
col_reduce_kernel(const float *__restrict__ x, int64_t ldx, const float *__restrict__ y, int64_t ldy, int64_t M, int C,
                  int64_t rows_per_slab, BnCtx ctx, float *__restrict__ out)
{
    __shared__ float red[2][kColWarps][kColTile];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int col = blockIdx.x * kColTile + lane;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_slab, r1 = min(r0 + rows_per_slab, M);
    float a = 0.f, b = 0.f;
    if (col < C) {
        float mean = 0.f, rstd = 0.f, gamma = 0.f, beta = 0.f;
        if (MODE == 1) { mean = ctx.mean[col]; rstd = ctx.rstd[col]; gamma = ctx.gamma[col]; beta = ctx.beta[col]; }
        // MODE 0 sums (x - s) and (x - s)^2 with the shift s = row 0 of the matrix: a sample lies within a few standard
        // deviations of the mean, so E[(x-s)^2] - E[x-s]^2 does not cancel the way E[x^2] - E[x]^2 does when |mean| >> std
        const float shift = MODE == 0 ? __ldg(x + col) : 0.f;
        for (int64_t m = r0 + w; m < r1; m += kColWarps) {
            const float v = __ldg(x + m * ldx + col) - shift;
            if (MODE == 0) { a += v; b = fmaf(v, v, b); }
            else if (MODE == 2) a += v;
            else {
                // x = dz (gradient w.r.t. the layer output), y = pre-BN activations
                const float xhat = (__ldg(y + m * ldy + col) - mean) * rstd;
                float g = v * bn_drop(ctx, m, col);
                if (ctx.relu && !(fmaf(gamma, xhat, beta) > 0.f)) g = 0.f;
                a += g; b = fmaf(g, xhat, b);
            }
        }
    }
    red[0][w][lane] = a; red[1][w][lane] = b;
    __syncthreads();
    if (w == 0 && col < C) {
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int i = 0; i < kColWarps; ++i) { sa += red[0][i][lane]; sb += red[1][i][lane]; }      // fixed order
        if (MODE == 2) out[(int64_t)blockIdx.y * C + col] = sa;
        else { out[((int64_t)blockIdx.y * 2 + 0) * C + col] = sa; out[((int64_t)blockIdx.y * 2 + 1) * C + col] = sb; }
    }
}

// Sum of one channel's slab partials by a warp: lane l adds slabs l, l + 32, ... in order, then a fixed shuffle tree --
// the same association for a given n_slabs on every run (fp64 accumulation).
__device__ __forceinline__ double warp_slab_sum(const float *__restrict__ part, int n_slabs, int64_t stride, int lane)
{
    double s = 0.0;
    for (int i = lane; i < n_slabs; i += 32) s += (double)part[(int64_t)i * stride];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

// BatchNorm statistics from the slab partials: mean, rstd = 1/sqrt(biased var + eps), running-stat update
// (running = (1 - momentum) * running + momentum * batch, with the UNBIASED variance, nn.BatchNorm1d semantics).
// One warp per channel.
__global__ void bn_finalize_kernel(const float *__restrict__ part, int n_slabs, int C, int64_t M, const float *__restrict__ shift_row,
                                   float eps, float momentum, float *__restrict__ mean, float *__restrict__ rstd,
                                   float *__restrict__ running_mean, float *__restrict__ running_var)
{
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= C) return;
    const double s = warp_slab_sum(part + c, n_slabs, 2 * (int64_t)C, lane);            // sums of (x - shift) and (x - shift)^2
    const double ss = warp_slab_sum(part + C + c, n_slabs, 2 * (int64_t)C, lane);
    if (lane != 0) return;
    const double d = s / (double)M;
    const double mu = (double)shift_row[c] + d;
    double var = ss / (double)M - d * d;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)mu;
    rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
    if (running_var) {
        const double unbiased = M > 1 ? var * ((double)M / (double)(M - 1)) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

// sums of slab partials -> out[c] (fp64 accumulation, fixed order); n_vec = 1 or 2 interleaved vectors per slab.
// One warp per channel.
__global__ void slab_sum_kernel(const float *__restrict__ part, int n_slabs, int n_vec, int C, float *__restrict__ out0,
                                float *__restrict__ out1)
{
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= C) return;
    const double s0 = warp_slab_sum(part + c, n_slabs, (int64_t)n_vec * C, lane);
    const double s1 = n_vec == 2 ? warp_slab_sum(part + C + c, n_slabs, (int64_t)n_vec * C, lane) : 0.0;
    if (lane != 0) return;
    if (out0) out0[c] = (float)s0;
    if (n_vec == 2 && out1) out1[c] = (float)s1;
}

// z = relu(gamma * (y - mean) * rstd + beta) [* dropout multiplier]
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float *__restrict__ y, int64_t M, int C, BnCtx ctx, float *__restrict__ z)
{
    const int64_t total = M * (int64_t)C;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int64_t m = i / C;
        const int c = (int)(i - m * C);
        float v = fmaf(ctx.gamma[c], (y[i] - ctx.mean[c]) * ctx.rstd[c], ctx.beta[c]);
        if (ctx.relu) v = fmaxf(v, 0.f);
        z[i] = v * bn_drop(ctx, m, c);
    }
}

// dy = gamma * rstd * (g - sum_g / M - xhat * sum_gx / M), g as in col_reduce MODE 1.  dz and dy may alias.
__global__ void __launch_bounds__(256)
bn_backward_apply_kernel(const float *dz, const float *__restrict__ y, int64_t M, int C, BnCtx ctx,
                         const float *__restrict__ sum_g, const float *__restrict__ sum_gx, float *dy)
{
    const int64_t total = M * (int64_t)C;
    const float inv_m = 1.f / (float)M;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int64_t m = i / C;
        const int c = (int)(i - m * C);
        const float xhat = (y[i] - ctx.mean[c]) * ctx.rstd[c];
        float g = dz[i] * bn_drop(ctx, m, c);
        if (ctx.relu && !(fmaf(ctx.gamma[c], xhat, ctx.beta[c]) > 0.f)) g = 0.f;
        dy[i] = ctx.gamma[c] * ctx.rstd[c] * (g - sum_g[c] * inv_m - xhat * sum_gx[c] * inv_m);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// float4 forms of the three passes above for contiguous (M, C) matrices with C a multiple of 64 (every BatchNorm of the
// nets at training batch sizes): 16-byte loads, four rows in flight per thread, per-thread channel constants in
// registers (no per-element index division).  Same outputs, same fixed reduction order within a launch geometry.
// ---------------------------------------------------------------------------------------------------------------
// four consecutive per-channel constants; the vectors may be views of a flat parameter / gradient buffer (4-byte aligned)
__device__ __forceinline__ float4 ld4(const float *p) { return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3)); }

__device__ __forceinline__ float bn_g_of(float dz, float y, float mean, float rstd, float gamma, float beta, float drop, int relu, float &xhat)
{
    xhat = (y - mean) * rstd;
    float g = dz * drop;
    if (relu && !(fmaf(gamma, xhat, beta) > 0.f)) g = 0.f;
    return g;
}

// grid = (C / (4 * tile_cv), n_slabs), 256 threads; tile_cv = 32 (one row per warp step) or 16 (two rows per warp step)
template <int MODE>
__global__ void __launch_bounds__(256)
col_reduce_v4_kernel(const float *__restrict__ x, const float *__restrict__ y, int64_t M, int C, int tile_cv, int64_t rows_per_slab,
                     BnCtx ctx, float *__restrict__ out)
{
    __shared__ float4 red[2][kColWarps][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int rows_w = 32 / tile_cv;                                  // rows per warp step
    const int cvl = lane % tile_cv, rsub = lane / tile_cv;
    const int col = (blockIdx.x * tile_cv + cvl) * 4;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_slab, r1 = min(r0 + rows_per_slab, M);
    const int64_t step = (int64_t)kColWarps * rows_w;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    float4 mean = a, rstd = a, gamma = a, beta = a;
    if (MODE == 1) { mean = ld4(ctx.mean + col); rstd = ld4(ctx.rstd + col); gamma = ld4(ctx.gamma + col); beta = ld4(ctx.beta + col); }
    const float4 shift = MODE == 0 ? ld4(x + col) : make_float4(0.f, 0.f, 0.f, 0.f);     // row 0, see col_reduce_kernel
    for (int64_t m0 = r0 + (int64_t)w * rows_w + rsub; m0 < r1; m0 += 4 * step) {
        float4 v[4], u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t m = m0 + k * step;
            v[k] = make_float4(0.f, 0.f, 0.f, 0.f); u[k] = v[k];
            if (m < r1) {
                v[k] = __ldg(reinterpret_cast<const float4 *>(x + m * C + col));
                if (MODE == 1) u[k] = __ldg(reinterpret_cast<const float4 *>(y + m * C + col));
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t m = m0 + k * step;
            if (m >= r1) break;
            if (MODE == 0) {
                const float4 t = make_float4(v[k].x - shift.x, v[k].y - shift.y, v[k].z - shift.z, v[k].w - shift.w);
                a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
                b.x = fmaf(t.x, t.x, b.x); b.y = fmaf(t.y, t.y, b.y); b.z = fmaf(t.z, t.z, b.z); b.w = fmaf(t.w, t.w, b.w);
            } else if (MODE == 2) {
                a.x += v[k].x; a.y += v[k].y; a.z += v[k].z; a.w += v[k].w;
            } else {
                float d0 = 1.f, d1 = 1.f, d2 = 1.f, d3 = 1.f;
                if (ctx.drop) { d0 = bn_drop(ctx, m, col); d1 = bn_drop(ctx, m, col + 1); d2 = bn_drop(ctx, m, col + 2); d3 = bn_drop(ctx, m, col + 3); }
                float xh, g;
                g = bn_g_of(v[k].x, u[k].x, mean.x, rstd.x, gamma.x, beta.x, d0, ctx.relu, xh); a.x += g; b.x = fmaf(g, xh, b.x);
                g = bn_g_of(v[k].y, u[k].y, mean.y, rstd.y, gamma.y, beta.y, d1, ctx.relu, xh); a.y += g; b.y = fmaf(g, xh, b.y);
                g = bn_g_of(v[k].z, u[k].z, mean.z, rstd.z, gamma.z, beta.z, d2, ctx.relu, xh); a.z += g; b.z = fmaf(g, xh, b.z);
                g = bn_g_of(v[k].w, u[k].w, mean.w, rstd.w, gamma.w, beta.w, d3, ctx.relu, xh); a.w += g; b.w = fmaf(g, xh, b.w);
            }
        }
    }
    if (rows_w == 2) {                                                // the two row phases of a warp: fixed order (low half + high half)
        a.x += __shfl_xor_sync(0xffffffffu, a.x, 16); a.y += __shfl_xor_sync(0xffffffffu, a.y, 16);
        a.z += __shfl_xor_sync(0xffffffffu, a.z, 16); a.w += __shfl_xor_sync(0xffffffffu, a.w, 16);
        b.x += __shfl_xor_sync(0xffffffffu, b.x, 16); b.y += __shfl_xor_sync(0xffffffffu, b.y, 16);
        b.z += __shfl_xor_sync(0xffffffffu, b.z, 16); b.w += __shfl_xor_sync(0xffffffffu, b.w, 16);
    }
    red[0][w][lane] = a; red[1][w][lane] = b;
    __syncthreads();
    if (w == 0 && lane < tile_cv) {
        float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa;
#pragma unroll
        for (int i = 0; i < kColWarps; ++i) {                         // fixed order
            const float4 p = red[0][i][lane], q = red[1][i][lane];
            sa.x += p.x; sa.y += p.y; sa.z += p.z; sa.w += p.w; sb.x += q.x; sb.y += q.y; sb.z += q.z; sb.w += q.w;
        }
        if (MODE == 2) *reinterpret_cast<float4 *>(out + (int64_t)blockIdx.y * C + col) = sa;
        else {
            *reinterpret_cast<float4 *>(out + ((int64_t)blockIdx.y * 2 + 0) * C + col) = sa;
            *reinterpret_cast<float4 *>(out + ((int64_t)blockIdx.y * 2 + 1) * C + col) = sb;
        }
    }
}

// every thread keeps ONE group of four channels (its constants live in registers) and strides over the rows;
// requires (gridDim.x * 256) % (C / 4) == 0.  BACKWARD: dy = gamma * rstd * (g - sum_g / M - xhat * sum_gx / M).
template <bool BACKWARD>
__global__ void __launch_bounds__(256)
bn_apply_v4_kernel(const float *dz, const float *__restrict__ y, int64_t M, int C, BnCtx ctx, const float *__restrict__ sum_g,
                   const float *__restrict__ sum_gx, float *out)
{
    const int CV = C >> 2;
    const int64_t gtid = (int64_t)blockIdx.x * 256 + threadIdx.x, nthr = (int64_t)gridDim.x * 256;
    const int col = (int)(gtid % CV) * 4;
    const int64_t row_step = nthr / CV;
    const float4 mean = ld4(ctx.mean + col), rstd = ld4(ctx.rstd + col), gamma = ld4(ctx.gamma + col), beta = ld4(ctx.beta + col);
    float4 sg = make_float4(0.f, 0.f, 0.f, 0.f), sgx = sg;
    const float inv_m = 1.f / (float)M;
    if (BACKWARD) {
        sg = ld4(sum_g + col); sgx = ld4(sum_gx + col);
        sg.x *= inv_m; sg.y *= inv_m; sg.z *= inv_m; sg.w *= inv_m; sgx.x *= inv_m; sgx.y *= inv_m; sgx.z *= inv_m; sgx.w *= inv_m;
    }
    for (int64_t m0 = gtid / CV; m0 < M; m0 += 4 * row_step) {
        float4 v[4], u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t m = m0 + k * row_step;
            if (m < M) {
                v[k] = __ldg(reinterpret_cast<const float4 *>(y + m * C + col));
                if (BACKWARD) u[k] = *reinterpret_cast<const float4 *>(dz + m * C + col);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t m = m0 + k * row_step;
            if (m >= M) break;
            float d0 = 1.f, d1 = 1.f, d2 = 1.f, d3 = 1.f;
            if (ctx.drop) { d0 = bn_drop(ctx, m, col); d1 = bn_drop(ctx, m, col + 1); d2 = bn_drop(ctx, m, col + 2); d3 = bn_drop(ctx, m, col + 3); }
            float4 o;
            if (!BACKWARD) {
                o.x = fmaf(gamma.x, (v[k].x - mean.x) * rstd.x, beta.x); o.y = fmaf(gamma.y, (v[k].y - mean.y) * rstd.y, beta.y);
                o.z = fmaf(gamma.z, (v[k].z - mean.z) * rstd.z, beta.z); o.w = fmaf(gamma.w, (v[k].w - mean.w) * rstd.w, beta.w);
                if (ctx.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                o.x *= d0; o.y *= d1; o.z *= d2; o.w *= d3;
            } else {
                float xh, g;
                g = bn_g_of(u[k].x, v[k].x, mean.x, rstd.x, gamma.x, beta.x, d0, ctx.relu, xh); o.x = gamma.x * rstd.x * (g - sg.x - xh * sgx.x);
                g = bn_g_of(u[k].y, v[k].y, mean.y, rstd.y, gamma.y, beta.y, d1, ctx.relu, xh); o.y = gamma.y * rstd.y * (g - sg.y - xh * sgx.y);
                g = bn_g_of(u[k].z, v[k].z, mean.z, rstd.z, gamma.z, beta.z, d2, ctx.relu, xh); o.z = gamma.z * rstd.z * (g - sg.z - xh * sgx.z);
                g = bn_g_of(u[k].w, v[k].w, mean.w, rstd.w, gamma.w, beta.w, d3, ctx.relu, xh); o.w = gamma.w * rstd.w * (g - sg.w - xh * sgx.w);
            }
            *reinterpret_cast<float4 *>(out + m * C + col) = o;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// max over the n rows of each group (object) with the arg-max row (first maximum), and its backward scatter
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kColTile * kColWarps)
group_max_kernel(const float *__restrict__ z, int C, int64_t n, float *__restrict__ g, int32_t *__restrict__ arg)
{
    __shared__ float rv[kColWarps][kColTile];
    __shared__ int ri[kColWarps][kColTile];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int col = blockIdx.x * kColTile + lane;
    const int64_t grp = blockIdx.y;
    float best = -INFINITY;
    int bi = 0;
    if (col < C)
        for (int64_t r = w; r < n; r += kColWarps) {
            const float v = __ldg(z + (grp * n + r) * C + col);
            if (v > best) { best = v; bi = (int)r; }           // rows visited in ascending order: keeps the first maximum
        }
    rv[w][lane] = best; ri[w][lane] = bi;
    __syncthreads();
    if (w == 0 && col < C) {
#pragma unroll
        for (int i = 1; i < kColWarps; ++i) {
            const float v = rv[i][lane];
            const int idx = ri[i][lane];
            if (v > best || (v == best && idx < bi)) { best = v; bi = idx; }
        }
        g[grp * C + col] = best;
        arg[grp * C + col] = bi;
    }
}

__global__ void group_max_backward_kernel(const float *__restrict__ dg, const int32_t *__restrict__ arg, int64_t G, int C, int64_t n,
                                          float *__restrict__ dz)
{
    const int64_t total = G * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t grp = i / C;
        const int c = (int)(i - grp * C);
        dz[(grp * n + arg[i]) * C + c] = dg[i];               // dz is zero-filled by the caller
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Weight gradient dW (N, K) = dY^T (N, M) . X (M, K): 64 x 64 output tile per CTA over one slab of rows
// ---------------------------------------------------------------------------------------------------------------
constexpr int WG_T = 64, WG_R = 32;
__global__ void __launch_bounds__(256)
wgrad_partial_kernel(const float *__restrict__ dy, int64_t ldy, const float *__restrict__ x, int64_t ldx, int64_t M, int N, int K,
                     int64_t rows_per_slab, float *__restrict__ part)
{
    __shared__ float sd[WG_R][WG_T + 4];
    __shared__ float sx[WG_R][WG_T + 4];
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * WG_T, k0 = blockIdx.y * WG_T;
    const int64_t r0 = (int64_t)blockIdx.z * rows_per_slab, r1 = min(r0 + rows_per_slab, M);
    const int ti = tid / 16, tj = tid % 16;                   // outputs (n0 + ti*4 + i, k0 + tj*4 + j)
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int64_t rb = r0; rb < r1; rb += WG_R) {
        // 32 rows x 64 columns of each operand: 2048 floats -> 8 per thread (coalesced along the columns)
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int f = tid + it * 256;
            const int r = f >> 6, c = f & 63;
            const int64_t m = rb + r;
            sd[r][c] = (m < r1 && n0 + c < N) ? __ldg(dy + m * ldy + n0 + c) : 0.f;
            sx[r][c] = (m < r1 && k0 + c < K) ? __ldg(x + m * ldx + k0 + c) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < WG_R; ++r) {
            const float4 a = *reinterpret_cast<const float4 *>(&sd[r][ti * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&sx[r][tj * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *dst = part + (int64_t)blockIdx.z * N * K;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = n0 + ti * 4 + i, k = k0 + tj * 4 + j;
            if (o < N && k < K) dst[(int64_t)o * K + k] = acc[i][j];
        }
}

__global__ void wgrad_reduce_kernel(const float *__restrict__ part, int n_slabs, int64_t NK, float *__restrict__ dw, int64_t lddw, int K,
                                    int accumulate)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < NK; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int sl = 0; sl < n_slabs; ++sl) s += part[(int64_t)sl * NK + i];        // fixed order
        const int64_t o = i / K, k = i - o * K;
        float *d = dw + o * lddw + k;
        *d = accumulate ? *d + s : s;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Loss backward (gradient of  sum_t w[t] * term_t  w.r.t. the network outputs; terms as in csrc/loss.cu:
// 0 mask NLL, 1 centre Huber(2), 2 heading CE, 3 size CE, 4 heading-residual Huber(1), 5 size-residual Huber(1))
// ---------------------------------------------------------------------------------------------------------------
__constant__ float c_mean_size[9] = {4.8f, 1.8f, 1.5f, 10.0f, 2.6f, 3.2f, 2.0f, 1.0f, 1.6f};

__global__ void __launch_bounds__(256)
loss_mask_backward_kernel(const float *__restrict__ logits, const float *__restrict__ label, int64_t M, const float *__restrict__ w6,
                          float *__restrict__ dlogits)
{
    const float scale = w6[0] / (float)M;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < M; i += (int64_t)gridDim.x * 256) {
        const float2 l = __ldg(reinterpret_cast<const float2 *>(logits) + i);
        const float m = fmaxf(l.x, l.y);
        const float e0 = expf(l.x - m), e1 = expf(l.y - m), inv = 1.f / (e0 + e1);
        const int lab = (int)(long long)__ldg(label + i);
        reinterpret_cast<float2 *>(dlogits)[i] = make_float2((e0 * inv - (lab == 0 ? 1.f : 0.f)) * scale,
                                                             (e1 * inv - (lab == 0 ? 0.f : 1.f)) * scale);
    }
}

__device__ __forceinline__ float huber_grad(float err, float delta)          // d huber / d err
{
    return fabsf(err) <= delta ? err : (err > 0.f ? delta : -delta);
}

__global__ void loss_heads_backward_kernel(const float *__restrict__ center, const float *__restrict__ center_label,
                                           const float *__restrict__ hs, const int64_t *__restrict__ hcls, const float *__restrict__ hrn,
                                           const float *__restrict__ hres_label, const float *__restrict__ ss,
                                           const int64_t *__restrict__ scls, const float *__restrict__ srn,
                                           const float *__restrict__ sres_label, int bs, const float *__restrict__ w6,
                                           float *__restrict__ dbox)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= bs) return;
    // dbox row: [centre 0..2 | heading scores 3..14 | heading residuals (normalised) 15..26 | size scores 27..29 | size residuals 30..38]
    float *dcenter = dbox + b * 39 - b * 3, *dhs = dbox + b * 39 + 3 - b * 12, *dhrn = dbox + b * 39 + 15 - b * 12,
          *dss = dbox + b * 39 + 27 - b * 3, *dsrn = dbox + b * 39 + 30 - b * 9;
    const float inv_bs = 1.f / (float)bs;
    {
        const float dx = center[b * 3] - center_label[b * 3], dy = center[b * 3 + 1] - center_label[b * 3 + 1],
                    dz = center[b * 3 + 2] - center_label[b * 3 + 2];
        const float d = sqrtf(dx * dx + dy * dy + dz * dz);
        const float k = d > 0.f ? huber_grad(d, 2.0f) / d * w6[1] * inv_bs : 0.f;
        dcenter[b * 3] = k * dx; dcenter[b * 3 + 1] = k * dy; dcenter[b * 3 + 2] = k * dz;
    }
    {
        const float *v = hs + b * 12;
        float m = v[0];
        for (int i = 1; i < 12; ++i) m = fmaxf(m, v[i]);
        float se = 0.f;
        for (int i = 0; i < 12; ++i) se += expf(v[i] - m);
        const int c = (int)hcls[b];
        const float lab = hres_label[b] / 0.2617993877991494f;
        const float ge = huber_grad(hrn[b * 12 + c] - lab, 1.0f) * w6[4] * inv_bs;
        for (int i = 0; i < 12; ++i) {
            dhs[b * 12 + i] = (expf(v[i] - m) / se - (i == c ? 1.f : 0.f)) * w6[2] * inv_bs;
            dhrn[b * 12 + i] = (i == c) ? ge : 0.f;
        }
    }
    {
        const float *v = ss + b * 3;
        const float m = fmaxf(v[0], fmaxf(v[1], v[2]));
        const float e[3] = {expf(v[0] - m), expf(v[1] - m), expf(v[2] - m)};
        const float se = e[0] + e[1] + e[2];
        const int c = (int)scls[b];
        float er[3], d2 = 0.f;
        for (int k = 0; k < 3; ++k) { er[k] = sres_label[b * 3 + k] / c_mean_size[c * 3 + k] - srn[b * 9 + c * 3 + k]; d2 += er[k] * er[k]; }
        const float d = sqrtf(d2);
        const float kk = d > 0.f ? huber_grad(d, 1.0f) / d * w6[5] * inv_bs : 0.f;
        for (int i = 0; i < 3; ++i) {
            dss[b * 3 + i] = (e[i] / se - (i == c ? 1.f : 0.f)) * w6[3] * inv_bs;
            for (int k = 0; k < 3; ++k) dsrn[b * 9 + i * 3 + k] = (i == c) ? -kk * er[k] : 0.f;
        }
    }
}

// number of points whose arg-max class equals the label (torch.argmax: first maximum -> class 1 only if l1 > l0)
__global__ void __launch_bounds__(256)
seg_correct_kernel(const float *__restrict__ logits, const float *__restrict__ label, int64_t M, unsigned long long *__restrict__ count)
{
    unsigned int local = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < M; i += (int64_t)gridDim.x * 256) {
        const float2 l = __ldg(reinterpret_cast<const float2 *>(logits) + i);
        const int pred = l.y > l.x ? 1 : 0;
        local += (pred == (int)(long long)__ldg(label + i)) ? 1u : 0u;
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, (unsigned long long)local);      // integer adds commute: exact
}

// torch.optim.Adam (amsgrad = False, maximize = False) over one flat bucket:
//   g += wd * p;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__global__ void __launch_bounds__(256)
adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, int64_t n, float lr,
            float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt, float grad_scale)
{
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        float gi = g[i] * grad_scale;
        const float pi = p[i];
        gi = fmaf(wd, pi, gi);
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - (lr / bc1) * (mi / denom);
    }
}

static int slabs_for(int64_t M, int64_t tiles, int64_t min_rows, int64_t *rows_per_slab)
{
    int64_t slabs = std::max<int64_t>(1, std::min<int64_t>(ceil_div(M, min_rows), std::max<int64_t>(1, 592 / std::max<int64_t>(tiles, 1))));
    *rows_per_slab = ceil_div(M, slabs);
    return (int)ceil_div(M, *rows_per_slab);
}

}  // namespace train
}  // namespace al3d

using namespace al3d;
using namespace al3d::train;

// geometry of the float4 kernels; 0 if (M, C) does not qualify
static int v4_mask()
{
    const char *e = getenv("AL3D_BN_V4");        // debugging aid: bit mask of the float4 kernels to use (default: all)
    return e ? atoi(e) : 63;
}
static int v4_tile_cv(int C, const void *a, const void *b, const void *c)
{
    if (C % 64 != 0 || C / 4 > 256 || 256 % (C / 4) != 0) return 0;       // the apply kernel needs (C / 4) | 256
    if (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) return 0;
    return (C / 4) % 32 == 0 ? 32 : 16;
}
static int v4_slabs(int64_t M, int C, int tile_cv, int64_t *rps) { return slabs_for(M, (C / 4) / tile_cv, 512, rps); }
static int v4_apply_grid(int64_t M, int C)
{
    const int64_t want = ceil_div(M * (int64_t)(C / 4), 256 * 4);        // ~4 float4 per thread
    return (int)std::max<int64_t>(1, std::min<int64_t>(want, 148 * 8));
}

static BnCtx make_ctx(const float *mean, const float *rstd, const float *gamma, const float *beta, const float *drop, int64_t sg,
                      int64_t sc, int64_t sr, int64_t rows_per_group, int relu, int64_t M)
{
    BnCtx c;
    c.mean = mean; c.rstd = rstd; c.gamma = gamma; c.beta = beta;
    c.drop = drop; c.drop_sg = sg; c.drop_sc = sc; c.drop_sr = sr; c.rows_per_group = rows_per_group > 0 ? rows_per_group : (M > 0 ? M : 1);
    c.relu = relu;
    return c;
}

extern "C" int al3d_train_ws_floats(int64_t M, int C)
{
    // largest scratch any column reduction over (M, C) needs, in floats
    int64_t rps;
    int slabs = slabs_for(M, ceil_div(C, kColTile), 256, &rps);
    if (C % 64 == 0 && C / 4 <= 256) slabs = std::max(slabs, v4_slabs(M, C, (C / 4) % 32 == 0 ? 32 : 16, &rps));
    return (int)std::min<int64_t>((int64_t)slabs * 2 * C, 0x7fffffff);
}

extern "C" int al3d_bn_train_forward(const float *y, int64_t M, int C, const float *gamma, const float *beta, float eps, float momentum,
                                     float *running_mean, float *running_var, const float *drop, int64_t drop_sg, int64_t drop_sc,
                                     int64_t drop_sr, int64_t rows_per_group, int relu, float *ws, float *mean, float *rstd, float *z,
                                     void *stream)
{
    AL3D_CHECK_ARG(y && gamma && beta && ws && mean && rstd && z, "al3d_bn_train_forward: null pointer");
    AL3D_CHECK_ARG(M >= 2 && C >= 1, "al3d_bn_train_forward: M=%lld C=%d (batch statistics need >= 2 rows)", (long long)M, C);
    cudaStream_t st = (cudaStream_t)stream;
    int64_t rps;
    BnCtx none = make_ctx(nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0, 0, M);
    const int tcv = v4_tile_cv(C, y, z, ws);
    int slabs;
    if (tcv && (v4_mask() & 1)) {
        slabs = v4_slabs(M, C, tcv, &rps);
        col_reduce_v4_kernel<0><<<dim3((unsigned)((C / 4) / tcv), slabs), 256, 0, st>>>(y, nullptr, M, C, tcv, rps, none, ws);
        AL3D_CHECK_LAUNCH("col_reduce_v4_kernel<0>");
    } else {
        slabs = slabs_for(M, ceil_div(C, kColTile), 256, &rps);
        col_reduce_kernel<0><<<dim3((unsigned)ceil_div(C, kColTile), slabs), kColTile * kColWarps, 0, st>>>(y, C, nullptr, 0, M, C, rps, none, ws);
        AL3D_CHECK_LAUNCH("col_reduce_kernel<0>");
    }
    bn_finalize_kernel<<<(unsigned)ceil_div(C, 4), 128, 0, st>>>(ws, slabs, C, M, y, eps, momentum, mean, rstd, running_mean, running_var);
    AL3D_CHECK_LAUNCH("bn_finalize_kernel");
    BnCtx ctx = make_ctx(mean, rstd, gamma, beta, drop, drop_sg, drop_sc, drop_sr, rows_per_group, relu, M);
    if (tcv && (v4_mask() & 2)) {
        bn_apply_v4_kernel<false><<<v4_apply_grid(M, C), 256, 0, st>>>(nullptr, y, M, C, ctx, nullptr, nullptr, z);
        AL3D_CHECK_LAUNCH("bn_apply_v4_kernel");
        return 0;
    }
    const int grid = (int)std::min<int64_t>(ceil_div(M * C, 256), 148 * 16);
    bn_apply_kernel<<<grid, 256, 0, st>>>(y, M, C, ctx, z);
    AL3D_CHECK_LAUNCH("bn_apply_kernel");
    return 0;
}

extern "C" int al3d_bn_train_backward(const float *dz, const float *y, int64_t M, int C, const float *gamma, const float *beta,
                                      const float *mean, const float *rstd, const float *drop, int64_t drop_sg, int64_t drop_sc,
                                      int64_t drop_sr, int64_t rows_per_group, int relu, float *ws, float *dgamma, float *dbeta,
                                      float *dy, void *stream)
{
    AL3D_CHECK_ARG(dz && y && gamma && beta && mean && rstd && ws && dgamma && dbeta && dy, "al3d_bn_train_backward: null pointer");
    AL3D_CHECK_ARG(M >= 2 && C >= 1, "al3d_bn_train_backward: M=%lld C=%d", (long long)M, C);
    cudaStream_t st = (cudaStream_t)stream;
    int64_t rps;
    BnCtx ctx = make_ctx(mean, rstd, gamma, beta, drop, drop_sg, drop_sc, drop_sr, rows_per_group, relu, M);
    const int tcv = (((uintptr_t)ws & 15) == 0) ? v4_tile_cv(C, dz, y, dy) : 0;
    int slabs;
    if (tcv && (v4_mask() & 4)) {
        slabs = v4_slabs(M, C, tcv, &rps);
        col_reduce_v4_kernel<1><<<dim3((unsigned)((C / 4) / tcv), slabs), 256, 0, st>>>(dz, y, M, C, tcv, rps, ctx, ws);
        AL3D_CHECK_LAUNCH("col_reduce_v4_kernel<1>");
    } else {
        slabs = slabs_for(M, ceil_div(C, kColTile), 256, &rps);
        col_reduce_kernel<1><<<dim3((unsigned)ceil_div(C, kColTile), slabs), kColTile * kColWarps, 0, st>>>(dz, C, y, C, M, C, rps, ctx, ws);
        AL3D_CHECK_LAUNCH("col_reduce_kernel<1>");
    }
    // dbeta = sum g, dgamma = sum g * xhat
    slab_sum_kernel<<<(unsigned)ceil_div(C, 4), 128, 0, st>>>(ws, slabs, 2, C, dbeta, dgamma);
    AL3D_CHECK_LAUNCH("slab_sum_kernel");
    if (tcv && (v4_mask() & 8)) {
        bn_apply_v4_kernel<true><<<v4_apply_grid(M, C), 256, 0, st>>>(dz, y, M, C, ctx, dbeta, dgamma, dy);
        AL3D_CHECK_LAUNCH("bn_apply_v4_kernel");
        return 0;
    }
    const int grid = (int)std::min<int64_t>(ceil_div(M * C, 256), 148 * 16);
    bn_backward_apply_kernel<<<grid, 256, 0, st>>>(dz, y, M, C, ctx, dbeta, dgamma, dy);
    AL3D_CHECK_LAUNCH("bn_backward_apply_kernel");
    return 0;
}

extern "C" int al3d_group_colsum(const float *x, int64_t M, int C, int64_t rows_per_group, float *ws, float *out, void *stream)
{
    AL3D_CHECK_ARG(x && out, "al3d_group_colsum: null pointer");
    AL3D_CHECK_ARG(M >= 1 && C >= 1, "al3d_group_colsum: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    BnCtx none = make_ctx(nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0, 0, M);
    if (rows_per_group > 0 && rows_per_group < M) {
        // one slab per group: out (G, C) directly
        AL3D_CHECK_ARG(M % rows_per_group == 0, "al3d_group_colsum: M not a multiple of rows_per_group");
        const int64_t G = M / rows_per_group;
        AL3D_CHECK_ARG(G <= 65535, "al3d_group_colsum: too many groups");
        if (const int tcv = (v4_mask() & 32) ? v4_tile_cv(C, x, out, nullptr) : 0) {
            col_reduce_v4_kernel<2><<<dim3((unsigned)((C / 4) / tcv), (unsigned)G), 256, 0, st>>>(x, nullptr, M, C, tcv, rows_per_group, none, out);
            AL3D_CHECK_LAUNCH("col_reduce_v4_kernel<2>");
            return 0;
        }
        col_reduce_kernel<2><<<dim3((unsigned)ceil_div(C, kColTile), (unsigned)G), kColTile * kColWarps, 0, st>>>(x, C, nullptr, 0, M, C,
                                                                                                                 rows_per_group, none, out);
        AL3D_CHECK_LAUNCH("col_reduce_kernel<2>");
        return 0;
    }
    AL3D_CHECK_ARG(ws, "al3d_group_colsum: workspace needed for the whole-matrix sum");
    int64_t rps;
    int slabs;
    if (const int tcv = (v4_mask() & 16) ? v4_tile_cv(C, x, ws, nullptr) : 0) {
        slabs = v4_slabs(M, C, tcv, &rps);
        col_reduce_v4_kernel<2><<<dim3((unsigned)((C / 4) / tcv), slabs), 256, 0, st>>>(x, nullptr, M, C, tcv, rps, none, ws);
        AL3D_CHECK_LAUNCH("col_reduce_v4_kernel<2>");
    } else {
        slabs = slabs_for(M, ceil_div(C, kColTile), 256, &rps);
        col_reduce_kernel<2><<<dim3((unsigned)ceil_div(C, kColTile), slabs), kColTile * kColWarps, 0, st>>>(x, C, nullptr, 0, M, C, rps, none, ws);
        AL3D_CHECK_LAUNCH("col_reduce_kernel<2>");
    }
    slab_sum_kernel<<<(unsigned)ceil_div(C, 4), 128, 0, st>>>(ws, slabs, 1, C, out, nullptr);
    AL3D_CHECK_LAUNCH("slab_sum_kernel");
    return 0;
}

extern "C" int al3d_group_max_forward(const float *z, int64_t G, int64_t n, int C, float *g, int32_t *arg, void *stream)
{
    AL3D_CHECK_ARG(z && g && arg, "al3d_group_max_forward: null pointer");
    AL3D_CHECK_ARG(G >= 1 && G <= 65535 && n >= 1 && C >= 1, "al3d_group_max_forward: bad shape G=%lld n=%lld C=%d", (long long)G, (long long)n, C);
    group_max_kernel<<<dim3((unsigned)ceil_div(C, kColTile), (unsigned)G), kColTile * kColWarps, 0, (cudaStream_t)stream>>>(z, C, n, g, arg);
    AL3D_CHECK_LAUNCH("group_max_kernel");
    return 0;
}

extern "C" int al3d_group_max_backward(const float *dg, const int32_t *arg, int64_t G, int64_t n, int C, float *dz_zeroed, void *stream)
{
    AL3D_CHECK_ARG(dg && arg && dz_zeroed, "al3d_group_max_backward: null pointer");
    const int grid = (int)std::min<int64_t>(ceil_div(G * C, 256), 148 * 8);
    group_max_backward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dg, arg, G, C, n, dz_zeroed);
    AL3D_CHECK_LAUNCH("group_max_backward_kernel");
    return 0;
}

extern "C" int al3d_wgrad_ws_floats(int64_t M, int N, int K)
{
    int64_t rps;
    const int slabs = slabs_for(M, ceil_div(N, WG_T) * ceil_div(K, WG_T), 512, &rps);
    return (int)std::min<int64_t>((int64_t)slabs * N * K, 0x7fffffff);
}

extern "C" int al3d_wgrad_f32(const float *dy, int64_t ldy, const float *x, int64_t ldx, int64_t M, int N, int K, float *ws,
                              float *dw, int64_t lddw, int accumulate, void *stream)
{
    AL3D_CHECK_ARG(dy && x && ws && dw, "al3d_wgrad_f32: null pointer");
    AL3D_CHECK_ARG(M >= 1 && N >= 1 && K >= 1 && ldy >= N && ldx >= K && lddw >= K, "al3d_wgrad_f32: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t rps;
    const int slabs = slabs_for(M, ceil_div(N, WG_T) * ceil_div(K, WG_T), 512, &rps);
    AL3D_CHECK_ARG(slabs <= 65535, "al3d_wgrad_f32: too many slabs");
    wgrad_partial_kernel<<<dim3((unsigned)ceil_div(N, WG_T), (unsigned)ceil_div(K, WG_T), slabs), 256, 0, st>>>(dy, ldy, x, ldx, M, N, K, rps, ws);
    AL3D_CHECK_LAUNCH("wgrad_partial_kernel");
    const int64_t NK = (int64_t)N * K;
    wgrad_reduce_kernel<<<(unsigned)std::min<int64_t>(ceil_div(NK, 256), 148 * 8), 256, 0, st>>>(ws, slabs, NK, dw, lddw, K, accumulate);
    AL3D_CHECK_LAUNCH("wgrad_reduce_kernel");
    return 0;
}

extern "C" int al3d_loss_backward(const float *logits, const float *mask_label, int64_t M, const float *center, const float *center_label,
                                  const float *heading_scores, const int64_t *heading_cls_label, const float *heading_res_norm,
                                  const float *heading_res_label, const float *size_scores, const int64_t *size_cls_label,
                                  const float *size_res_norm, const float *size_res_label, int bs, const float *w6, float *dlogits,
                                  float *dbox, void *stream)
{
    AL3D_CHECK_ARG(center && center_label && heading_scores && heading_cls_label && heading_res_norm && heading_res_label && size_scores &&
                   size_cls_label && size_res_norm && size_res_label && w6 && dbox, "al3d_loss_backward: null pointer");
    AL3D_CHECK_ARG(bs > 0, "al3d_loss_backward: bs must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    if (logits) {
        AL3D_CHECK_ARG(mask_label && dlogits && M > 0, "al3d_loss_backward: mask term needs labels and an output");
        loss_mask_backward_kernel<<<(unsigned)std::min<int64_t>(ceil_div(M, 256), 148 * 8), 256, 0, st>>>(logits, mask_label, M, w6, dlogits);
        AL3D_CHECK_LAUNCH("loss_mask_backward_kernel");
    }
    loss_heads_backward_kernel<<<(unsigned)ceil_div(bs, 128), 128, 0, st>>>(center, center_label, heading_scores, heading_cls_label,
                                                                            heading_res_norm, heading_res_label, size_scores, size_cls_label,
                                                                            size_res_norm, size_res_label, bs, w6, dbox);
    AL3D_CHECK_LAUNCH("loss_heads_backward_kernel");
    return 0;
}

extern "C" int al3d_seg_correct(const float *logits, const float *mask_label, int64_t M, unsigned long long *count_zeroed, void *stream)
{
    AL3D_CHECK_ARG(logits && mask_label && count_zeroed && M >= 0, "al3d_seg_correct: bad arguments");
    if (M == 0) return 0;
    seg_correct_kernel<<<(unsigned)std::min<int64_t>(ceil_div(M, 256), 148 * 8), 256, 0, (cudaStream_t)stream>>>(logits, mask_label, M, count_zeroed);
    AL3D_CHECK_LAUNCH("seg_correct_kernel");
    return 0;
}

extern "C" int al3d_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr, float beta1,
                              float beta2, float eps, float weight_decay, int step, float grad_scale, void *stream)
{
    AL3D_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "al3d_adam_step: bad arguments");
    if (n == 0) return 0;
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
    adam_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n, 256), 148 * 8), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr,
                                                                                                         beta1, beta2, eps, weight_decay, bc1,
                                                                                                         bc2_sqrt, grad_scale);
    AL3D_CHECK_LAUNCH("adam_kernel");
    return 0;
}
