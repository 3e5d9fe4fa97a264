// Tensor-core (tcgen05 / TMEM) kernels for the shared point-wise MLPs, bf16 operands, fp32 accumulate.
//
//   chain_max_kernel   first layer (tiny K, CUDA cores) -> 2-3 chained MMA layers whose activations
//                      never leave shared memory -> last layer computed transposed (channels on TMEM
//                      lanes, points on columns) so the max over points is a per-thread reduction.
//                      Serves ins_seg conv1-5 + max (tools/static_model.py:279-284), the static box
//                      head trunk (:330-334), PointEmbedding and BoxEmbedding trunks
//                      (tools/dynamic_model.py:241-245, 278-282).
//   seg_pass2_kernel   conv1-2 recomputed, dconv1 (+ per-object global-feature bias) pipelined in
//                      128-channel chunks into dconv2's accumulation, dconv3, dconv4, and the 128->2
//                      logits + mask in the last epilogue (tools/static_model.py:286-295, :59).
//
// Both are persistent, warp-specialised: warp 0 streams packed weight blocks with cp.async.bulk into a
// ring, warp 1 (one thread) issues tcgen05.mma, warps 2-5 own the 128 TMEM lanes and run every
// epilogue.  BatchNorm is folded into the weights by the host; ReLU and bias are applied in the epilogue.
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "umma.cuh"
#include "../../include/al3d.h"

namespace al3d {
using namespace umma;

// Delay-injection hooks of the protocol stress tests.  They are compiled into a second library
// (libal3d_stress.so, -DAL3D_TC_STRESS_HOOKS) only: the hooks sit inside unrolled loops of kernels whose instruction
// footprint is already large, and the product library must not pay for them.
#ifdef AL3D_TC_STRESS_HOOKS
#define AL3D_STRESS(st, salt) stress_delay(st, salt)
#define AL3D_STRESS_WARP(st, salt) stress_delay_warp(st, salt)
#else
#define AL3D_STRESS(st, salt) do { } while (0)
#define AL3D_STRESS_WARP(st, salt) do { } while (0)
#endif

constexpr int kTile = 128;                 // points per tile == TMEM lanes
constexpr int kStageBytes = 16384;         // one weight block: 128 rows x 64 K bf16
constexpr int kPlane = kTile * 16;         // bytes of one activation K-plane (128 rows x 16 B)
constexpr int kEpiThreads = 256;            // 8 epilogue warps: (lane quarter, column half)
constexpr int kThreads = 64 + kEpiThreads; // + producer warp + MMA warp

// Epilogue thread geometry: warps 2..9.  A warp may only touch TMEM lanes 32*(warp%4)..+31, so the
// two warps that share a lane quarter split the accumulator columns between them.
__device__ __forceinline__ int epi_row() { return ((threadIdx.x >> 5) & 3) * 32 + (threadIdx.x & 31); }
__device__ __forceinline__ int epi_half() { return ((threadIdx.x >> 5) - 2) >> 2; }

// packed fp32 pairs in one 64-bit register (sm_100 add / fma .f32x2): half the instructions of the scalar form
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t f2_relu(uint64_t a)
{
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
    return f2_pack(fmaxf(lo, 0.f), fmaxf(hi, 0.f));
}
__device__ __forceinline__ float f2_hsum(uint64_t a)
{
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
    return lo + hi;
}

__device__ __forceinline__ uint32_t relu_pack_bf16x2(float lo, float hi)
{
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ float fmax3(float a, float b, float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// bias + ReLU + bf16 pack of 32 accumulator columns -> four 16-byte plane rows
__device__ __forceinline__ void store_act32(uint8_t *buf, int plane0, int row, const uint32_t (&v)[32], const float *bias)
{
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 b0 = *reinterpret_cast<const float4 *>(bias + j * 8);
        const float4 b1 = *reinterpret_cast<const float4 *>(bias + j * 8 + 4);
        uint4 o;
        o.x = relu_pack_bf16x2(__uint_as_float(v[j * 8 + 0]) + b0.x, __uint_as_float(v[j * 8 + 1]) + b0.y);
        o.y = relu_pack_bf16x2(__uint_as_float(v[j * 8 + 2]) + b0.z, __uint_as_float(v[j * 8 + 3]) + b0.w);
        o.z = relu_pack_bf16x2(__uint_as_float(v[j * 8 + 4]) + b1.x, __uint_as_float(v[j * 8 + 5]) + b1.y);
        o.w = relu_pack_bf16x2(__uint_as_float(v[j * 8 + 6]) + b1.z, __uint_as_float(v[j * 8 + 7]) + b1.w);
        *reinterpret_cast<uint4 *>(buf + (plane0 + j) * kPlane + row * 16) = o;
    }
}

// Accumulator columns [c0, c0 + NC) of this thread's TMEM lane -> bias + ReLU -> bf16 planes.
// Two 32-column loads are kept in flight per wait.
template <int NC>
__device__ __forceinline__ void epilogue_cols(uint32_t taddr, uint8_t *buf, int c0, int row, const float *bias)
{
    static_assert(NC == 32 || NC % 64 == 0, "NC");
    if constexpr (NC == 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        tmem_ld_wait();
        store_act32(buf, c0 >> 3, row, v, bias + c0);
    } else {
#pragma unroll
        for (int c = 0; c < NC; c += 64) {
            uint32_t v0[32], v1[32];
            tmem_ld32(taddr + c0 + c, v0);
            tmem_ld32(taddr + c0 + c + 32, v1);
            tmem_ld_wait();
            store_act32(buf, (c0 + c) >> 3, row, v0, bias + c0 + c);
            store_act32(buf, (c0 + c + 32) >> 3, row, v1, bias + c0 + c + 32);
        }
    }
}
__device__ __forceinline__ void epilogue_cols_n(int ncols, uint32_t taddr, uint8_t *buf, int c0, int row, const float *bias)
{
    if (ncols == 32) epilogue_cols<32>(taddr, buf, c0, row, bias);
    else if (ncols == 64) epilogue_cols<64>(taddr, buf, c0, row, bias);
    else epilogue_cols<128>(taddr, buf, c0, row, bias);
}

// First layer on CUDA cores: x[c] (c < c_in) -> output channels [ch0, ch0 + nch), bf16, into the KP
// buffer.  Weights are stored transposed in shared memory: sw[c * w0 + ch].
__device__ __forceinline__ void first_layer(uint8_t *buf, int row, const float *xv, int c_in, int w0, int ch0, int nch,
                                            const float *sw, const float *sb)
{
    for (int ch = ch0; ch < ch0 + nch; ch += 8) {
        float4 a0 = *reinterpret_cast<const float4 *>(sb + ch), a1 = *reinterpret_cast<const float4 *>(sb + ch + 4);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (c >= c_in) break;
            const float4 w0v = *reinterpret_cast<const float4 *>(sw + c * w0 + ch);
            const float4 w1v = *reinterpret_cast<const float4 *>(sw + c * w0 + ch + 4);
            const float x = xv[c];
            a0.x = fmaf(x, w0v.x, a0.x); a0.y = fmaf(x, w0v.y, a0.y); a0.z = fmaf(x, w0v.z, a0.z); a0.w = fmaf(x, w0v.w, a0.w);
            a1.x = fmaf(x, w1v.x, a1.x); a1.y = fmaf(x, w1v.y, a1.y); a1.z = fmaf(x, w1v.z, a1.z); a1.w = fmaf(x, w1v.w, a1.w);
        }
        uint4 o;
        o.x = relu_pack_bf16x2(a0.x, a0.y); o.y = relu_pack_bf16x2(a0.z, a0.w);
        o.z = relu_pack_bf16x2(a1.x, a1.y); o.w = relu_pack_bf16x2(a1.z, a1.w);
        *reinterpret_cast<uint4 *>(buf + (ch >> 3) * kPlane + row * 16) = o;
    }
}

// ================================================================================================
// UMMA self-test: D(128 x N) = A(128 x K) * B(N x K)^T from KP-packed bf16 operands.
// ================================================================================================
__global__ void __launch_bounds__(128)
umma_selftest_kernel(const uint8_t *__restrict__ a_kp, const uint8_t *__restrict__ b_kp, int N, int K, float *__restrict__ d, int swap, const TcStatus wd)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *sa = smem;
    uint8_t *sb = smem + (size_t)128 * K * 2;
    const int a_bytes = 128 * K * 2, b_bytes = N * K * 2;
    for (int i = threadIdx.x * 16; i < a_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4 *>(sa + i) = *reinterpret_cast<const uint4 *>(a_kp + i);
    for (int i = threadIdx.x * 16; i < b_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4 *>(sb + i) = *reinterpret_cast<const uint4 *>(b_kp + i);
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc<256>(&tmem_base_s);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N);
        for (int k0 = 0; k0 < K; k0 += 16) {
            const uint64_t da = make_desc(smem_u32(sa) + (k0 / 8) * 128 * 16, 128, swap != 0);
            const uint64_t db = make_desc(smem_u32(sb) + (k0 / 8) * N * 16, N, swap != 0);
            mma_bf16(tmem, da, db, idesc, k0 > 0 ? 1u : 0u);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0, 0xE001, wd);
    tc_fence_after();
    const int row = threadIdx.x;       // warps 0..3 <-> lane quarters 0..3
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(row & ~31) << 16) + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) if (c0 + i < N) d[row * N + c0 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<256>(tmem);
}

// Same product with the A operand staged in TMEM: every thread packs its row of A (fp32 in global memory)
// to bf16 pairs and tcgen05.st's them to columns [256, 256 + K/2); checks the A-from-TMEM conventions.
__global__ void __launch_bounds__(128)
umma_selftest_ts_kernel(const float *__restrict__ a, const uint8_t *__restrict__ b_kp, int N, int K, float *__restrict__ d, const TcStatus wd)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *sb = smem;
    const int b_bytes = N * K * 2;
    for (int i = threadIdx.x * 16; i < b_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4 *>(sb + i) = *reinterpret_cast<const uint4 *>(b_kp + i);
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&tmem_base_s);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int row = threadIdx.x;
    const uint32_t lane_addr = (uint32_t)(row & ~31) << 16;
    for (int k0 = 0; k0 < K; k0 += 32) {
        uint32_t v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = pack_bf16x2(a[row * K + k0 + 2 * i], a[row * K + k0 + 2 * i + 1]);
        tmem_st16(tmem + lane_addr + 256 + k0 / 2, v);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N);
        for (int k0 = 0; k0 < K; k0 += 16) {
            const uint64_t db = make_desc(smem_u32(sb) + (k0 / 8) * N * 16, N);
            mma_bf16_ts(tmem, tmem + 256 + k0 / 2, db, idesc, k0 > 0 ? 1u : 0u);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0, 0xE002, wd);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_addr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) if (c0 + i < N) d[row * N + c0 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

// CTA-pair variant (cluster of 2, cta_group::2): D(256 x N) = A(256 x K) * B(N x K)^T.  CTA r stages rows
// [128r, 128r+128) of A in its TMEM and rows [r*N/2, (r+1)*N/2) of B (KP-packed, N/2 rows) in its shared memory;
// the leader issues the M = 256 MMAs; each CTA reads back its 128 rows of D.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
umma_selftest_pair_kernel(const float *__restrict__ a, const uint8_t *__restrict__ b_kp_halves, int N, int K, float *__restrict__ d, const TcStatus wd)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t rank = cluster_ctarank();
    const int half_bytes = (N / 2) * K * 2;
    const uint8_t *src = b_kp_halves + (size_t)rank * half_bytes;
    for (int i = threadIdx.x * 16; i < half_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4 *>(smem + i) = *reinterpret_cast<const uint4 *>(src + i);
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc_pair<512>(&tmem_base_s);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int row = threadIdx.x;
    const uint32_t lane_addr = (uint32_t)(row & ~31) << 16;
    const float *arow = a + (size_t)(rank * 128 + row) * K;
    for (int k0 = 0; k0 < K; k0 += 32) {
        uint32_t v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = pack_bf16x2(arow[k0 + 2 * i], arow[k0 + 2 * i + 1]);
        tmem_st16(tmem + lane_addr + 256 + k0 / 2, v);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                         // both CTAs' operands and barriers are ready
    tc_fence_after();
    if (rank == 0 && threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(256, N);
        for (int k0 = 0; k0 < K; k0 += 16) {
            const uint64_t db = make_desc(smem_u32(smem) + (k0 / 8) * (N / 2) * 16, N / 2);
            mma_bf16_ts_pair(tmem, tmem + 256 + k0 / 2, db, idesc, k0 > 0 ? 1u : 0u);
        }
        mma_commit_pair(&bar, 0x3);
    }
    mbar_wait(&bar, 0, 0xE003, wd);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_addr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) if (c0 + i < N) d[(size_t)(rank * 128 + row) * N + c0 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (threadIdx.x < 32) tmem_dealloc_pair<512>(tmem);
}

// CTA-pair variant with BOTH operands in shared memory (the transposed conv5 of seg_pass1_kernel): CTA r stages rows
// [128r, 128r+128) of A as a 128-row KP tile, and its N/2 = 64 rows of B as rows [64,128) of a 128-row KP tile (a
// sub-tile descriptor: plane stride 2048 B), exactly how the kernel addresses half of a tile's conv4 output.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
umma_selftest_pair_ss_kernel(const uint8_t *__restrict__ a_kp_halves, const uint8_t *__restrict__ b_kp_halves, int K, float *__restrict__ d, const TcStatus wd)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t rank = cluster_ctarank();
    const int N = 128;
    uint8_t *sa = smem, *sbt = smem + 128 * K * 2;               // A tile | B container tile (128 rows)
    const int a_bytes = 128 * K * 2;
    for (int i = threadIdx.x * 16; i < a_bytes; i += blockDim.x * 16)
        *reinterpret_cast<uint4 *>(sa + i) = *reinterpret_cast<const uint4 *>(a_kp_halves + (size_t)rank * a_bytes + i);
    // B half: 64 rows x K, KP-packed with 64-row planes in global memory -> rows 64..127 of the 128-row container
    const uint8_t *bsrc = b_kp_halves + (size_t)rank * 64 * K * 2;
    for (int i = threadIdx.x; i < (K / 8) * 64; i += blockDim.x) {
        const int plane = i / 64, r = i % 64;
        *reinterpret_cast<uint4 *>(sbt + plane * 2048 + (64 + r) * 16) = *reinterpret_cast<const uint4 *>(bsrc + plane * 1024 + r * 16);
    }
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc_pair<512>(&tmem_base_s);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (rank == 0 && threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(256, N);
        for (int k0 = 0; k0 < K; k0 += 16) {
            const uint64_t da = make_desc(smem_u32(sa) + (k0 / 8) * 2048, 128);
            const uint64_t db = make_desc_lbo(smem_u32(sbt) + 64 * 16 + (k0 / 8) * 2048, 2048);
            mma_bf16_ss_pair(tmem, da, db, idesc, k0 > 0 ? 1u : 0u);
        }
        mma_commit_pair(&bar, 0x3);
    }
    mbar_wait(&bar, 0, 0xE004, wd);
    tc_fence_after();
    const int row = threadIdx.x;
    const uint32_t lane_addr = (uint32_t)(row & ~31) << 16;
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_addr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) d[(size_t)(rank * 128 + row) * N + c0 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (threadIdx.x < 32) tmem_dealloc_pair<512>(tmem);
}

// ================================================================================================
// chain_max_kernel
// ================================================================================================
struct ChainParams {
    const float *x; int64_t sb, sc, sp; int bs, n;       // input (bs, c_in, n), strides in elements
    int c_in, w0, n_mid, mid[3], last;
    const float *w0_w, *w0_b, *mid_b, *last_b;           // fp32: (8, w0) transposed + zero-padded, (w0), concat(mid), (last)
    const uint8_t *wstream;                              // packed bf16 blocks, 16 KB slots, consumption order
    float *out;                                          // (bs, last) fp32, zero-initialised; max-pooled with atomicMax
    int splits;                                          // work items per object
    int n_items;
    int bufA_bytes, bufB_bytes, n_stages;                // shared-memory carve-up chosen by the launcher
    TcStatus wd;                                         // watchdog status word of this launch
};

constexpr int kChainMaxStages = 12;
// Dynamic shared memory: [bufA | bufB | weight ring (n_stages x 16 KB) | ChainSmemTail]
struct ChainSmemTail {
    float w0_w[128 * 8];
    float w0_b[128];
    float mid_b[512];
    uint64_t w_full[kChainMaxStages], w_empty[kChainMaxStages];
    uint64_t act_ready, acc_ready;
    uint64_t last_full[2], last_empty[2];
    uint32_t tmem_base;
};
struct ChainSmemView {
    uint8_t *bufA, *bufB, *wring;
    ChainSmemTail *t;
};

__device__ __forceinline__ int chain_in_width(const ChainParams &p, int l) { return l == 0 ? p.w0 : p.mid[l - 1]; }

__global__ void __launch_bounds__(kThreads, 1)
chain_max_kernel(const ChainParams p)
{
    const TcStatus &wd = p.wd;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *const s_bufA = smem_raw;
    uint8_t *const s_bufB = smem_raw + p.bufA_bytes;
    uint8_t *const s_wring = s_bufB + p.bufB_bytes;
    ChainSmemTail &s = *reinterpret_cast<ChainSmemTail *>(s_wring + (size_t)p.n_stages * kStageBytes);
    const int kChainStages = p.n_stages;
    const int warp = threadIdx.x >> 5;

    for (int i = threadIdx.x; i < p.w0 * 8; i += kThreads) s.w0_w[i] = p.w0_w[i];
    for (int i = threadIdx.x; i < p.w0; i += kThreads) s.w0_b[i] = p.w0_b[i];
    {
        int tot = 0;
        for (int l = 0; l < p.n_mid; ++l) tot += p.mid[l];
        for (int i = threadIdx.x; i < tot; i += kThreads) s.mid_b[i] = p.mid_b[i];
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < kChainStages; ++i) { mbar_init(&s.w_full[i], 1); mbar_init(&s.w_empty[i], 1); }
        mbar_init(&s.act_ready, kEpiThreads);
        mbar_init(&s.acc_ready, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&s.last_full[i], 1); mbar_init(&s.last_empty[i], kEpiThreads); }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<512>(&s.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;
    const int tiles_per_obj = (p.n + kTile - 1) / kTile;
    const int n_last_chunks = p.last / 128;
    const int k_last = p.mid[p.n_mid - 1];

    if (warp == 0) {
        // ------------------------------------------------------------ weight producer (one thread)
        if (elect_one_sync()) {
            int stage = 0; uint32_t phase = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const int sp_i = item % p.splits;
                const int t0 = (int)((int64_t)tiles_per_obj * sp_i / p.splits), t1 = (int)((int64_t)tiles_per_obj * (sp_i + 1) / p.splits);
                for (int t = t0; t < t1; ++t) {
                    int blk = 0;
                    for (int l = 0; l <= p.n_mid; ++l) {
                        const int K = (l < p.n_mid) ? chain_in_width(p, l) : k_last;
                        const int N = (l < p.n_mid) ? p.mid[l] : p.last;
                        const int rows = N < 128 ? N : 128;
                        const int nblk = (N / rows) * (K / 64);
                        for (int i = 0; i < nblk; ++i, ++blk) {
                            AL3D_STRESS(wd, 0xC1);
                            if (!mbar_wait(&s.w_empty[stage], phase ^ 1, 0xC100 + stage, wd)) goto done;
                            const uint32_t bytes = rows * 64 * 2;
                            mbar_arrive_expect_tx(&s.w_full[stage], bytes);
                            bulk_g2s((s_wring + (size_t)stage * kStageBytes), p.wstream + (size_t)blk * kStageBytes, bytes, &s.w_full[stage]);
                            if (++stage == kChainStages) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (one thread)
        if (elect_one_sync()) {
            int stage = 0; uint32_t wphase = 0, act_phase = 0, le_phase[2] = {0, 0};
            const uint32_t aA = smem_u32(s_bufA), aB = smem_u32(s_bufB);
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const int sp_i = item % p.splits;
                const int t0 = (int)((int64_t)tiles_per_obj * sp_i / p.splits), t1 = (int)((int64_t)tiles_per_obj * (sp_i + 1) / p.splits);
                for (int t = t0; t < t1; ++t) {
                    // mid layers: D[points x channels]; input buffer alternates A, B, A, ...
                    for (int l = 0; l < p.n_mid; ++l) {
                        const int K = chain_in_width(p, l), N = p.mid[l];
                        const int rows = N < 128 ? N : 128;
                        const uint32_t in_addr = (l & 1) ? aB : aA;
                        if (!mbar_wait(&s.act_ready, act_phase, 0xC200 + l, wd)) goto done;
                        act_phase ^= 1;
                        tc_fence_after();
                        const uint32_t idesc = make_idesc_bf16(128, rows);
                        for (int nc = 0; nc < N / rows; ++nc)
                            for (int kb = 0; kb < K / 64; ++kb) {
                                if (!mbar_wait(&s.w_full[stage], wphase, 0xC300 + stage, wd)) goto done;
                                tc_fence_after();
                                mma_block_k64(tmem + nc * 128, in_addr + kb * 8 * kPlane, 128, smem_u32((s_wring + (size_t)stage * kStageBytes)), rows, idesc, kb > 0);
                                mma_commit(&s.w_empty[stage]);
                                if (++stage == kChainStages) { stage = 0; wphase ^= 1; }
                            }
                        mma_commit(&s.acc_ready);
                    }
                    // last layer, transposed: D^T[channels x points], double-buffered in TMEM cols 256..511
                    {
                        const uint32_t in_addr = (p.n_mid & 1) ? aB : aA;
                        if (!mbar_wait(&s.act_ready, act_phase, 0xC2F0, wd)) goto done;
                        act_phase ^= 1;
                        tc_fence_after();
                        const uint32_t idesc = make_idesc_bf16(128, 128);
                        for (int cc = 0; cc < n_last_chunks; ++cc) {
                            const int b = cc & 1;
                            AL3D_STRESS(wd, 0xC4);
                            if (!mbar_wait(&s.last_empty[b], le_phase[b] ^ 1, 0xC400 + b, wd)) goto done;
                            le_phase[b] ^= 1;
                            tc_fence_after();
                            for (int kb = 0; kb < k_last / 64; ++kb) {
                                if (!mbar_wait(&s.w_full[stage], wphase, 0xC500 + stage, wd)) goto done;
                                tc_fence_after();
                                mma_block_k64(tmem + 256 + b * 128, smem_u32((s_wring + (size_t)stage * kStageBytes)), 128, in_addr + kb * 8 * kPlane, 128, idesc, kb > 0);
                                mma_commit(&s.w_empty[stage]);
                                if (++stage == kChainStages) { stage = 0; wphase ^= 1; }
                            }
                            mma_commit(&s.last_full[b]);
                        }
                    }
                }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue warps (256 threads)
        const int row = epi_row(), half = epi_half();
        const uint32_t lane_addr = (uint32_t)(row & ~31) << 16;
        uint32_t acc_phase = 0, lf_phase[2] = {0, 0};
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            const int b = item / p.splits, sp_i = item % p.splits;
            const int t0 = (int)((int64_t)tiles_per_obj * sp_i / p.splits), t1 = (int)((int64_t)tiles_per_obj * (sp_i + 1) / p.splits);
            float rmax[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) rmax[i] = -INFINITY;
            for (int t = t0; t < t1; ++t) {
                // ---- first layer (rows past the end of the object replicate its last point: the
                //      max-pool is idempotent under duplicates); each half computes w0/2 channels
                {
                    int pidx = t * kTile + row;
                    if (pidx > p.n - 1) pidx = p.n - 1;
                    const float *px = p.x + (int64_t)b * p.sb + (int64_t)pidx * p.sp;
                    float xv[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) xv[c] = (c < p.c_in) ? __ldg(px + c * p.sc) : 0.f;
                    first_layer(s_bufA, row, xv, p.c_in, p.w0, half * (p.w0 >> 1), p.w0 >> 1, s.w0_w, s.w0_b);
                    fence_proxy_async_smem();
                    mbar_arrive(&s.act_ready);
                }
                // ---- mid layers: this thread converts columns [half*N/2, (half+1)*N/2) of its row
                int boff = 0;
                for (int l = 0; l < p.n_mid; ++l) {
                    const int N = p.mid[l];
                    uint8_t *outb = (l & 1) ? s_bufA : s_bufB;
                    AL3D_STRESS_WARP(wd, 0xD1);
                    if (!mbar_wait(&s.acc_ready, acc_phase, 0xD100 + l, wd)) goto done;
                    acc_phase ^= 1;
                    tc_fence_after();
                    epilogue_cols_n(N >> 1, tmem + lane_addr, outb, half * (N >> 1), row, s.mid_b + boff);
                    boff += N;
                    tc_fence_before();
                    fence_proxy_async_smem();
                    mbar_arrive(&s.act_ready);
                }
                // ---- last layer: this thread owns channel (cc*128 + row) and 64 of the tile's 128 points
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    if (cc < n_last_chunks) {
                        const int bsel = cc & 1;
                        AL3D_STRESS_WARP(wd, 0xD2);
                        if (!mbar_wait(&s.last_full[bsel], lf_phase[bsel], 0xD200 + cc, wd)) goto done;
                        lf_phase[bsel] ^= 1;
                        tc_fence_after();
                        uint32_t v0[32], v1[32];
                        const uint32_t ta = tmem + lane_addr + 256 + bsel * 128 + half * 64;
                        tmem_ld32(ta, v0);
                        tmem_ld32(ta + 32, v1);
                        tmem_ld_wait();
                        tc_fence_before();
                        mbar_arrive(&s.last_empty[bsel]);          // the values are in registers: free the buffer early
                        float m0 = rmax[cc], m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
                        for (int i = 0; i < 32; i += 8) {
                            m0 = fmax3(m0, __uint_as_float(v0[i]), __uint_as_float(v0[i + 1]));
                            m1 = fmax3(m1, __uint_as_float(v0[i + 2]), __uint_as_float(v0[i + 3]));
                            m2 = fmax3(m2, __uint_as_float(v0[i + 4]), __uint_as_float(v0[i + 5]));
                            m3 = fmax3(m3, __uint_as_float(v0[i + 6]), __uint_as_float(v0[i + 7]));
                            m0 = fmax3(m0, __uint_as_float(v1[i]), __uint_as_float(v1[i + 1]));
                            m1 = fmax3(m1, __uint_as_float(v1[i + 2]), __uint_as_float(v1[i + 3]));
                            m2 = fmax3(m2, __uint_as_float(v1[i + 4]), __uint_as_float(v1[i + 5]));
                            m3 = fmax3(m3, __uint_as_float(v1[i + 6]), __uint_as_float(v1[i + 7]));
                        }
                        rmax[cc] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                    }
                }
            }
            // ---- publish: relu(max + bias) >= 0, so integer atomicMax on the bit pattern is exact
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
                if (cc < n_last_chunks && t1 > t0) {
                    const int ch = cc * 128 + row;
                    const float v = fmaxf(rmax[cc] + __ldg(p.last_b + ch), 0.f);
                    atomicMax(reinterpret_cast<int *>(p.out + (int64_t)b * p.last + ch), __float_as_int(v));
                }
            }
        }
    }
done:
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// ================================================================================================
// seg_pass2_kernel -- activations stay in TMEM
//
// Per 128-point tile every layer is D[points x channels] = A[points x K] * W^T with the A operand read from
// TENSOR MEMORY (tcgen05.mma, A-from-TMEM): the epilogue warps turn an fp32 accumulator into the next layer's
// bf16 operand in place (tcgen05.ld -> bias + ReLU -> cvt.bf16x2 -> tcgen05.st), so shared memory carries
// nothing but the streamed weight blocks (13-stage ring) and the MMA only fetches B from it.
//
// TMEM columns (fp32 accumulators / packed bf16 operands, 2 channels per column):
//   [0,256)    D2 = dconv2 accumulator; later A3 (its bf16 image): channels 0..127 at [0,64), 128..255 at [128,192)
//   [256,448)  three 64-column dconv1 chunk buffers D1b[0..2]; each is rewritten in place as the bf16 operand of
//              dconv2's partial sum (channels 0..31 at +0, 32..63 at +32).  D1b[2] first holds conv2's accumulator.
//              Later D3 = dconv3 accumulator at [256,384) and A4 (its bf16 image) at [256,288) + [320,352)
//   [448,480)  A2 = conv2 output (64 ch);  [480,512)  A1 = conv1 output (64 ch)
//   [0,128)    finally D4 = dconv4 accumulator (over the dead A3), which leaves [384,512) free for the NEXT tile's
//              conv1 / conv2 (A1, A2, conv2 accumulator) while this tile's dconv3 / dconv4 are still running
// ================================================================================================
struct Pass2Params {
    const float *x; int64_t sb, sc, sp; int bs, n; int c_in;
    const float *w1_w, *w1_b;          // conv1 fp32 (8, 64) transposed + padded, (64)
    const float *b2;                   // conv2 bias (64)
    const float *gbias;                // (bs, 512) per-object dconv1 bias (global-feature half + folded BN bias)
    const float *bd2, *bd3, *bd4;      // dconv2-4 biases (256),(128),(128)
    const float *w5, *b5;              // dconv5 fp32 (2,128), (2)
    const uint8_t *wstream;            // two per-CTA halves (kP2HalfBytes each) of the 23 packed blocks, tightly packed
    float *logits;                     // (bs, n, 2)
    uint8_t *mask;                     // (bs, n)
    int tiles_per_obj; int n_items;    // items = bs * tiles_per_obj
    long long *dbg;                    // optional clock64 timeline of CTA 0 (al3d_set_debug_buffer), else NULL
    int dbg_skip;                      // first recorded tile of the timeline (environment AL3D_DEBUG_SKIP)
    TcStatus wd;                       // watchdog status word of this launch
};

// timeline stamps: role 0 = MMA thread, 1 = epilogue thread 0 (CTA 0), 2 = epilogue thread 0 of CTA 1 (pass 2: the
// peer of the pair; its clock is another SM's, compare durations only); 64 stamps x 4 items each
#define AL3D_TS(role)                                                                         \
    do { const int ts_it = it_local - p.dbg_skip;                                             \
         if (p.dbg && blockIdx.x == ((role) == 2 ? 1 : 0) && ts_it >= 0 && ts_it < 4 && ts_i < 64) \
             p.dbg[((role) * 4 + ts_it) * 64 + ts_i++] = clock64(); } while (0)

constexpr int kP2Blocks = 23;
constexpr uint32_t kColD2 = 0, kColD1 = 256, kColA2 = 448, kColA1 = 480, kColD3 = 256, kColD4 = 0;

// The kernel runs as CTA PAIRS (cluster of 2, tcgen05 cta_group::2): one M = 256 MMA covers the 128-point tiles of
// both CTAs, and each CTA holds only HALF of every weight block (N/2 rows of B).  That halves the weight bytes
// per SM -- the per-SM L2 -> smem ingress (~41 B/cycle measured) is what bounded the single-CTA version -- and makes
// the whole 424 KB weight set fit: 212 KB per CTA stay resident in shared memory for the life of the kernel, so
// there is no weight streaming at all.
// Per-tile block list (order of use): 0 conv2 | 1-3 dconv1 chunks 0-2 | 4-13 five groups {dconv2 partial, dconv1 chunk}
// | 14-16 dconv2 partials | 17-20 dconv3 | 21-22 dconv4.  64-row blocks are 8 KB, 128-row blocks 16 KB, the 256-row
// dconv2 partials 32 KB; a CTA keeps half of each (one KP tile of R/2 rows).
__host__ __device__ constexpr uint32_t p2_block_bytes(int blk)
{
    return blk <= 3 ? 8192u : (blk < 14 ? ((blk - 4) % 2 == 0 ? 32768u : 8192u) : (blk < 17 ? 32768u : 16384u));
}
__host__ __device__ constexpr uint32_t p2_half_off(int blk)
{
    uint32_t off = 0;
    for (int i = 0; i < blk; ++i) off += p2_block_bytes(i) / 2;
    return off;
}
constexpr uint32_t kP2HalfBytes = p2_half_off(kP2Blocks);      // 217088 bytes of weights per CTA
struct Pass2Smem {
    uint8_t wres[kP2HalfBytes];        // this CTA's half of every weight block, resident
    float w1_w[64 * 8], w1_b[64], b2[64], gb[2][512], bd2[256], bd3[128], bd4[128], w5[256], b5[2];
    float lpart[2][2 * kTile];         // logits partial sums of the upper column half (double-buffered by tile parity)
    uint64_t res_full, res_peer;       // weights landed in this CTA / in the peer (leader only)
    uint64_t act_f, act_f2, acc_f, act_t, acc_t;   // front stream (conv1 -> act_f, conv2 epilogue -> act_f2 of the next tile), tail stream
    uint64_t d1_full[3], d1_act[3], d1_free[3];    // dconv1 chunk buffers: accumulator ready / bf16 image ready / consumed
    uint32_t tmem_base;
};
static_assert(sizeof(Pass2Smem) + 128 <= 232448, "Pass2Smem exceeds the 227 KB opt-in limit");

// 32 fp32 accumulator columns -> bias + ReLU -> 16 packed bf16x2 words
__device__ __forceinline__ void pack_act32(const uint32_t (&v)[32], const float *bias, uint32_t (&o)[16])
{
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 b = *reinterpret_cast<const float4 *>(bias + j * 4);
        o[2 * j] = relu_pack_bf16x2(__uint_as_float(v[4 * j]) + b.x, __uint_as_float(v[4 * j + 1]) + b.y);
        o[2 * j + 1] = relu_pack_bf16x2(__uint_as_float(v[4 * j + 2]) + b.z, __uint_as_float(v[4 * j + 3]) + b.w);
    }
}

// conv1 on CUDA cores for 32 output channels starting at ch0 -> 16 packed bf16x2 words.
// CIN > 0: compile-time channel count (all weight loads are hoisted; the runtime-bounded loop is latency-bound with
// two warps per scheduler); CIN == 0: runtime c_in <= 8.  WS: channel count of the layer (row stride of w_t).
template <int CIN, int WS>
__device__ __forceinline__ void conv1_pack32_t(const float (&xv)[8], int c_in, const float *w_t, const float *bias, int ch0,
                                               uint32_t (&o)[16])
{
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        float4 a0 = *reinterpret_cast<const float4 *>(bias + ch0 + g * 8);
        float4 a1 = *reinterpret_cast<const float4 *>(bias + ch0 + g * 8 + 4);
#pragma unroll
        for (int c = 0; c < (CIN > 0 ? CIN : 8); ++c) {
            if (CIN == 0 && c >= c_in) break;
            const float4 w0v = *reinterpret_cast<const float4 *>(w_t + c * WS + ch0 + g * 8);
            const float4 w1v = *reinterpret_cast<const float4 *>(w_t + c * WS + ch0 + g * 8 + 4);
            const float xx = xv[c];
            a0.x = fmaf(xx, w0v.x, a0.x); a0.y = fmaf(xx, w0v.y, a0.y); a0.z = fmaf(xx, w0v.z, a0.z); a0.w = fmaf(xx, w0v.w, a0.w);
            a1.x = fmaf(xx, w1v.x, a1.x); a1.y = fmaf(xx, w1v.y, a1.y); a1.z = fmaf(xx, w1v.z, a1.z); a1.w = fmaf(xx, w1v.w, a1.w);
        }
        o[g * 4 + 0] = relu_pack_bf16x2(a0.x, a0.y); o[g * 4 + 1] = relu_pack_bf16x2(a0.z, a0.w);
        o[g * 4 + 2] = relu_pack_bf16x2(a1.x, a1.y); o[g * 4 + 3] = relu_pack_bf16x2(a1.z, a1.w);
    }
}
template <int WS = 64>
__device__ __forceinline__ void conv1_pack32(const float (&xv)[8], int c_in, const float *w_t, const float *bias, int ch0,
                                             uint32_t (&o)[16])
{
    if (c_in == 3)      conv1_pack32_t<3, WS>(xv, c_in, w_t, bias, ch0, o);     // static model (x, y, z)
    else if (c_in == 4) conv1_pack32_t<4, WS>(xv, c_in, w_t, bias, ch0, o);     // dynamic model (x, y, z, t)
    else                conv1_pack32_t<0, WS>(xv, c_in, w_t, bias, ch0, o);
}

// Constant-bank variants: bias / conv1 weights that live in the kernel parameter block are read as instruction operands
// (c[0][..]) when every index is a compile-time constant -- no load instruction, no shared-memory bandwidth.  (An
// LDS.128 broadcast still returns 512 B per warp through the 128 B/clk shared-memory data path: the per-point conv1
// weights and the epilogue biases, read that way by every epilogue warp, cost more than the arithmetic.)
template <int OFF, int N>
__device__ __forceinline__ void pack_act32_cb(const uint32_t (&v)[32], const float (&bias)[N], uint32_t (&o)[16])
{
#pragma unroll
    for (int j = 0; j < 16; ++j)
        o[j] = relu_pack_bf16x2(__uint_as_float(v[2 * j]) + bias[OFF + 2 * j], __uint_as_float(v[2 * j + 1]) + bias[OFF + 2 * j + 1]);
}
// conv1 for 32 output channels CH0..CH0+31 of a 64-channel layer: w (8 x 64, input-channel-major) and b from the
// parameter block; same arithmetic order as conv1_pack32 (bias, then input channels in order)
template <int CIN, int CH0>
__device__ __forceinline__ void conv1_pack32_cb(const float (&xv)[8], const float (&w)[512], const float (&b)[64], uint32_t (&o)[16])
{
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        float a0 = b[CH0 + 2 * j], a1 = b[CH0 + 2 * j + 1];
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
            a0 = fmaf(xv[c], w[c * 64 + CH0 + 2 * j], a0);
            a1 = fmaf(xv[c], w[c * 64 + CH0 + 2 * j + 1], a1);
        }
        o[j] = relu_pack_bf16x2(a0, a1);
    }
}

// 4 MMAs (K = 64) with A in TMEM: a_col[s] is the TMEM column of K-slice s (8 columns each)
__device__ __forceinline__ void mma_ts_k64(uint32_t tmem_d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                           uint32_t b_addr, uint32_t rows_b, uint32_t idesc, bool accumulate_first)
{
    const uint32_t a[4] = {a0, a1, a2, a3};
#pragma unroll
    for (int k = 0; k < 4; ++k)
        mma_bf16_ts(tmem_d, a[k], make_desc(b_addr + k * 2 * rows_b * 16, rows_b), idesc, (accumulate_first || k > 0) ? 1u : 0u);
}

// CTA-pair version: M = 256 (128 rows in each CTA's TMEM), B = this block's rows split over the two CTAs
__device__ __forceinline__ void mma_pair_k64(uint32_t tmem_d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                             uint32_t b_addr, uint32_t rows_half, uint32_t idesc, bool accumulate_first)
{
    const uint32_t a[4] = {a0, a1, a2, a3};
#pragma unroll
    for (int k = 0; k < 4; ++k)
        mma_bf16_ts_pair(tmem_d, a[k], make_desc(b_addr + k * 2 * rows_half * 16, rows_half), idesc, (accumulate_first || k > 0) ? 1u : 0u);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads + 32, 1)
seg_pass2_kernel(const Pass2Params p)
{
    const TcStatus &wd = p.wd;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    Pass2Smem &s = *reinterpret_cast<Pass2Smem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();            // 0 = leader (issues the pair's MMAs), 1 = peer

    for (int i = threadIdx.x; i < 64 * 8; i += kThreads) s.w1_w[i] = p.w1_w[i];
    for (int i = threadIdx.x; i < 64; i += kThreads) { s.w1_b[i] = p.w1_b[i]; s.b2[i] = p.b2[i]; }
    for (int i = threadIdx.x; i < 256; i += kThreads) { s.bd2[i] = p.bd2[i]; s.w5[i] = p.w5[i]; }
    for (int i = threadIdx.x; i < 128; i += kThreads) { s.bd3[i] = p.bd3[i]; s.bd4[i] = p.bd4[i]; }
    if (threadIdx.x < 2) s.b5[threadIdx.x] = p.b5[threadIdx.x];
    if (threadIdx.x == 0) {
        mbar_init(&s.res_full, 1); mbar_init(&s.res_peer, 1);
        // epilogue -> MMA barriers live in the leader and collect one arrival per epilogue WARP of BOTH CTAs (an elected
        // lane arrives after the warp has synchronised: 8 remote transactions per hand-off instead of 256)
        mbar_init(&s.act_f, 2 * kEpiThreads / 32); mbar_init(&s.act_f2, 2 * kEpiThreads / 32); mbar_init(&s.act_t, 2 * kEpiThreads / 32);
        for (int i = 0; i < 3; ++i) mbar_init(&s.d1_free[i], 1);
        mbar_init(&s.acc_f, 1); mbar_init(&s.acc_t, 1);
        for (int i = 0; i < 3; ++i) { mbar_init(&s.d1_full[i], 1); mbar_init(&s.d1_act[i], 2 * kEpiThreads / 32); }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc_pair<512>(&s.tmem_base);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                  // both CTAs' barriers exist before anything signals them
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;
    // both CTAs of a pair run the same number of rounds; a CTA whose item index is past the end recomputes the last
    // item without writing anything ("ghost"), so the pair's shared MMA stream never changes shape
    const int n_rounds = (p.n_items + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0) {
        // ------------------------------------------------------------ weight loader: this CTA's half of every block, once
        if (elect_one_sync()) {
            mbar_arrive_expect_tx(&s.res_full, kP2HalfBytes);
            const uint8_t *src = p.wstream + (size_t)crank * kP2HalfBytes;
            for (uint32_t off = 0; off < kP2HalfBytes; off += 16384) {
                const uint32_t bytes = (kP2HalfBytes - off) < 16384u ? (kP2HalfBytes - off) : 16384u;
                bulk_g2s(s.wres + off, src + off, bytes, &s.res_full);
            }
            if (crank != 0) {
                // tell the leader that the peer's weights are in place
                if (!mbar_wait(&s.res_full, 0, 0xA1F0, wd)) goto done;
                mbar_arrive_remote_release(&s.res_peer, 0);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA only)
        // Two interleaved streams: the "tail" of round r (dconv3, dconv4) and the "front" of round r+1 (conv2), each
        // with its own pair of barriers, so the next tiles' conv1/conv2 are computed while this round's serial tail is
        // in flight.  Every MMA is a cta_group::2 instruction covering both CTAs' tiles; every commit is multicast to
        // both CTAs' barriers.
        if (crank == 0 && elect_one_sync()) {
            uint32_t actf_phase = 0, actt_phase = 0, d1a_phase[3] = {0, 0, 0};
            const uint32_t id64 = make_idesc_bf16(256, 64), id128 = make_idesc_bf16(256, 128), id256 = make_idesc_bf16(256, 256);
            const uint32_t wres = smem_u32(s.wres);
#define P2_W(blk) (wres + p2_half_off(blk))
#define P2_WAIT(bar, ph, code)                                                           \
            AL3D_STRESS(wd, 0xA2);                                                      \
            if (!mbar_wait_cluster(&(bar), ph, code, wd)) goto done;                         \
            AL3D_TS(0);                                                                  \
            ph ^= 1; tc_fence_after();
#define P2_ISSUE_CONV2()                                                                 \
            {                                                                            \
                P2_WAIT(s.act_f, actf_phase, 0xA200)                                     \
                mma_pair_k64(tmem + kColD1 + 128, tmem + kColA1, tmem + kColA1 + 8, tmem + kColA1 + 16, tmem + kColA1 + 24, P2_W(0), 32, id64, false); \
                mma_commit_pair(&s.acc_f, 0x3);                                          \
            }
            if (!mbar_wait(&s.res_full, 0, 0xA2FE, wd)) goto done;
            if (!mbar_wait_cluster(&s.res_peer, 0, 0xA2FF, wd)) goto done;
            tc_fence_after();
            int it_local = 0;
            int ts_i = 0;
            if (n_rounds > 0) P2_ISSUE_CONV2()                               // front of the first round
            for (int r = 0; r < n_rounds; ++r, ++it_local) {
                ts_i = 0;
                AL3D_TS(0);
                const bool has_next = r + 1 < n_rounds;
#pragma unroll
                for (int kc = 0; kc < 8; ++kc) {
                    const int j = kc % 3;
                    const int blk0 = kc < 5 ? 4 + 2 * kc : 14 + (kc - 5);           // dconv2 block of this group
                    AL3D_STRESS(wd, 0xA4);
                    if (!mbar_wait_cluster(&s.d1_act[j], d1a_phase[j], 0xA400 + kc, wd)) goto done;
                    AL3D_TS(0);
                    d1a_phase[j] ^= 1; tc_fence_after();
                    const uint32_t a = tmem + kColD1 + j * 64;        // bf16 image of chunk kc (in place)
                    // one N = 256 instruction per K slice: the issuing thread, not the tensor pipe, bounded the N = 128 form
                    mma_pair_k64(tmem + kColD2, a, a + 8, a + 32, a + 40, P2_W(blk0), 128, id256, kc > 0);
                    if (kc + 3 < 8) mma_commit_pair(&s.d1_free[j], 0x1);   // chunk kc + 3 may overwrite D1b[j] once this has run
                }
                mma_commit_pair(&s.acc_t, 0x3);                            // dconv2 accumulator complete
                // dconv3: A3 x Wd3 -> D3                                            (blocks 17..20)
                P2_WAIT(s.act_t, actt_phase, 0xA202)
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) {
                    const uint32_t a = tmem + kColD2 + (kb >> 1) * 128 + (kb & 1) * 32;
                    mma_pair_k64(tmem + kColD3, a, a + 8, a + 16, a + 24, P2_W(17 + kb), 64, id128, kb > 0);
                }
                mma_commit_pair(&s.acc_t, 0x3);
                AL3D_TS(0);
                // front of the next round: conv2 on its conv1 output (A1 at [480,512) -> D1b[2]; both are free now)
                if (has_next) P2_ISSUE_CONV2()
                // dconv4: A4 x Wd4 -> D4 (over the dead A3 at columns 0..127)       (blocks 21, 22)
                P2_WAIT(s.act_t, actt_phase, 0xA203)
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    const uint32_t a = tmem + kColD3 + kb * 64;
                    mma_pair_k64(tmem + kColD4, a, a + 8, a + 16, a + 24, P2_W(21 + kb), 64, id128, kb > 0);
                }
                mma_commit_pair(&s.acc_t, 0x3);
                // D3 / A4 lived over D1b[0], D1b[1]: the next round's first two dconv1 chunks may go once dconv4 has run
                mma_commit_pair(&s.d1_free[0], 0x1);
                mma_commit_pair(&s.d1_free[1], 0x1);
                AL3D_TS(0);
            }
#undef P2_W
#undef P2_WAIT
#undef P2_ISSUE_CONV2
        }
    } else if (warp == kThreads / 32) {
        // ------------------------------------------------------------ second MMA issuer (leader CTA only): dconv1 chunks.
        // One issuing thread spends ~780 cycles per chunk group (a barrier wait, eight instructions, a commit) against
        // ~640 cycles of tensor-pipe work, so the dconv1 chunks are issued from their own thread.  Its MMAs are not
        // ordered against the first issuer's: d1_free[j] (committed behind the partial sum that read D1b[j], and behind
        // dconv4 for the two buffers D3 / A4 were laid over) says when a chunk buffer may be overwritten.
        if (crank == 0 && elect_one_sync()) {
            uint32_t actf2_phase = 0, d1free_phase[3] = {0, 0, 0};
            const uint32_t id64 = make_idesc_bf16(256, 64);
            const uint32_t wres = smem_u32(s.wres);
            if (!mbar_wait(&s.res_full, 0, 0xA3FE, wd)) goto done;
            if (!mbar_wait_cluster(&s.res_peer, 0, 0xA3FF, wd)) goto done;
            for (int r = 0; r < n_rounds; ++r) {
                if (!mbar_wait_cluster(&s.act_f2, actf2_phase, 0xA300, wd)) goto done;      // A2 of this round's tiles is in place
                actf2_phase ^= 1; tc_fence_after();
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int j = c % 3;
                    if ((c >= 3 || (c < 2 && r > 0))) {
                        AL3D_STRESS(wd, 0xA3);
                        if (!mbar_wait(&s.d1_free[j], d1free_phase[j], 0xA310 + c, wd)) goto done;
                        d1free_phase[j] ^= 1; tc_fence_after();
                    }
                    const int blk = c < 3 ? 1 + c : 5 + 2 * (c - 3);
                    mma_pair_k64(tmem + kColD1 + j * 64, tmem + kColA2, tmem + kColA2 + 8, tmem + kColA2 + 16, tmem + kColA2 + 24,
                                 wres + p2_half_off(blk), 32, id64, false);
                    mma_commit_pair(&s.d1_full[j], 0x3);
                }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue warps (256 threads)
        const int row = epi_row(), half = epi_half();
        const uint32_t lane_addr = (uint32_t)(row & ~31) << 16;
        const uint32_t tl = tmem + lane_addr;
        const int etid = threadIdx.x - 64;
        uint32_t accf_phase = 0, acct_phase = 0, d1f_phase[3] = {0, 0, 0};
        int it_local = 0;
        int ts_i = 0;
        const bool ts_on = (threadIdx.x == 64);
#define AL3D_TSE() do { if (ts_on) { if (blockIdx.x == 0) AL3D_TS(1); else AL3D_TS(2); } } while (0)
        // the epilogue -> MMA barriers are the leader's: the peer CTA arrives on them through the cluster address space
#define P2_PUBLISH(bar) do { tmem_st_wait(); tc_fence_before(); __syncwarp();                                          \
                             if (lane == 0) { if (crank == 0) mbar_arrive(bar); else mbar_arrive_remote(bar, 0); } } while (0)
        // software prefetch of the tile's input point (hides the global-memory latency behind the previous tile)
        float xv[8];
        auto load_x = [&](int item, float (&dst)[8]) {
            if (item > p.n_items - 1) item = p.n_items - 1;        // ghost round: recompute the last item
            {
                const int b = item / p.tiles_per_obj, t = item % p.tiles_per_obj;
                int pidx = t * kTile + row;
                if (pidx > p.n - 1) pidx = p.n - 1;
                const float *px = p.x + (int64_t)b * p.sb + (int64_t)pidx * p.sp;
#pragma unroll
                for (int c = 0; c < 8; ++c) dst[c] = (c < p.c_in) ? __ldg(px + c * p.sc) : 0.f;
            }
        };
        // front, part 1: conv1 on CUDA cores (this thread: 32 channels) -> A1, then prefetch the following tile's point
        auto front_conv1 = [&](int next_item) {
            uint32_t o[16];
            conv1_pack32(xv, p.c_in, s.w1_w, s.w1_b, half * 32, o);
            tmem_st16(tl + kColA1 + half * 16, o);
            AL3D_TSE();
            P2_PUBLISH(&s.act_f);
            AL3D_TSE();
            load_x(next_item, xv);
        };
        // front, part 2: conv2 accumulator (D1b[2]) -> A2
        auto front_conv2 = [&]() -> bool {
            AL3D_STRESS_WARP(wd, 0xB1);
            if (!mbar_wait(&s.acc_f, accf_phase, 0xB100, wd)) return false;
            accf_phase ^= 1; tc_fence_after();
            uint32_t v[32], o[16];
            tmem_ld32(tl + kColD1 + 128 + half * 32, v);
            tmem_ld_wait();
            pack_act32(v, s.b2 + half * 32, o);
            tmem_st16(tl + kColA2 + half * 16, o);
            P2_PUBLISH(&s.act_f2);
            return true;
        };
        if (n_rounds > 0) {
            // first tile of this CTA: its per-object dconv1 bias and its front are not hidden behind a previous tile
            const int b0 = min((int)blockIdx.x, p.n_items - 1) / p.tiles_per_obj;
            for (int i = etid; i < 512; i += kEpiThreads) s.gb[0][i] = __ldg(p.gbias + (int64_t)b0 * 512 + i);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            load_x(blockIdx.x, xv);
            front_conv1(blockIdx.x + gridDim.x);
            if (!front_conv2()) goto done;
        }
        for (int r = 0; r < n_rounds; ++r, ++it_local) {
            ts_i = 0;
            AL3D_TSE();
            const int item_raw = (int)blockIdx.x + r * (int)gridDim.x;
            const bool ghost = item_raw >= p.n_items;
            const int item = ghost ? p.n_items - 1 : item_raw;
            const int b = item / p.tiles_per_obj, t = item % p.tiles_per_obj;
            const int pidx_raw = t * kTile + row;
            const bool valid = !ghost && pidx_raw < p.n;
            const int pidx = pidx_raw < p.n ? pidx_raw : p.n - 1;
            const int par = it_local & 1;
            const int next_item = item_raw + (int)gridDim.x;
            const bool has_next = r + 1 < n_rounds;                // uniform over the CTA pair
            if (has_next) {
                // the next tile's per-object bias goes into the other buffer: nobody reads gb[par ^ 1] during this tile,
                // and the bar.sync of the logits epilogue below orders these writes before the next tile's reads
                const int nb = min(next_item, p.n_items - 1) / p.tiles_per_obj;
                for (int i = etid; i < 512; i += kEpiThreads) s.gb[par ^ 1][i] = __ldg(p.gbias + (int64_t)nb * 512 + i);
            }
            // ---- dconv1 chunk epilogues, in place.  (Moving the next tile's conv1 into these slots was tried: the loop
            //      is bound by the buffer round trip epilogue -> partial sum -> dconv1 -> epilogue, so every cycle added
            //      here lengthens it; net loss.)
#pragma unroll
            for (int kc = 0; kc < 8; ++kc) {
                const int j = kc % 3;
                AL3D_STRESS_WARP(wd, 0xB2);
                if (!mbar_wait(&s.d1_full[j], d1f_phase[j], 0xB200 + kc, wd)) goto done;
                AL3D_TSE();
                d1f_phase[j] ^= 1; tc_fence_after();
                uint32_t v[32], o[16];
                const uint32_t ta = tl + kColD1 + j * 64 + half * 32;
                tmem_ld32(ta, v);
                tmem_ld_wait();
                pack_act32(v, s.gb[par] + kc * 64 + half * 32, o);
                tmem_st16(ta, o);
                P2_PUBLISH(&s.d1_act[j]);
                AL3D_TSE();
            }
            // ---- dconv2 epilogue: D2 -> A3 in place (this thread: 128 columns in two batches)
            AL3D_STRESS_WARP(wd, 0xB101);
            if (!mbar_wait(&s.acc_t, acct_phase, 0xB101, wd)) goto done;
            AL3D_TSE();
            acct_phase ^= 1; tc_fence_after();
#pragma unroll
            for (int bt = 0; bt < 2; ++bt) {
                uint32_t v0[32], v1[32], o0[16], o1[16];
                const uint32_t src = tl + kColD2 + half * 128 + bt * 64;
                tmem_ld32(src, v0);
                tmem_ld32(src + 32, v1);
                tmem_ld_wait();
                pack_act32(v0, s.bd2 + half * 128 + bt * 64, o0);
                pack_act32(v1, s.bd2 + half * 128 + bt * 64 + 32, o1);
                const uint32_t dst = tl + kColD2 + half * 128 + bt * 32;
                tmem_st16(dst, o0);
                tmem_st16(dst + 16, o1);
                AL3D_TSE();
            }
            P2_PUBLISH(&s.act_t);
            AL3D_TSE();
            // ---- while dconv3 runs: conv1 of the NEXT tile (A1 lives at [480,512), free since the chunk MMAs finished)
            if (has_next) front_conv1(next_item + (int)gridDim.x);
            AL3D_TSE();
            // ---- dconv3 epilogue: D3 -> A4 in place (64 columns)
            AL3D_STRESS_WARP(wd, 0xB102);
            if (!mbar_wait(&s.acc_t, acct_phase, 0xB102, wd)) goto done;
            AL3D_TSE();
            acct_phase ^= 1; tc_fence_after();
            {
                uint32_t v0[32], v1[32], o0[16], o1[16];
                const uint32_t src = tl + kColD3 + half * 64;
                tmem_ld32(src, v0);
                tmem_ld32(src + 32, v1);
                tmem_ld_wait();
                pack_act32(v0, s.bd3 + half * 64, o0);
                pack_act32(v1, s.bd3 + half * 64 + 32, o1);
                tmem_st16(src, o0);
                tmem_st16(src + 16, o1);
                P2_PUBLISH(&s.act_t);
                AL3D_TSE();
            }
            // ---- while dconv4 runs: conv2 epilogue of the NEXT tile (D1b[2] -> A2 at [448,480))
            if (has_next) { if (!front_conv2()) goto done; }
            AL3D_TSE();
            // ---- dconv4 epilogue: bias + ReLU in fp32, then the 128 -> 2 layer, logits and mask.
            //      Each half reduces 64 channels; the upper half hands its partial sums over in smem and the
            //      lower half adds them in a fixed order (deterministic).
            AL3D_STRESS_WARP(wd, 0xB103);
            if (!mbar_wait(&s.acc_t, acct_phase, 0xB103, wd)) goto done;
            AL3D_TSE();
            acct_phase ^= 1; tc_fence_after();
            {
                uint32_t v0[32], v1[32];
                const int c0 = half * 64;
                tmem_ld32(tl + kColD4 + c0, v0);
                tmem_ld32(tl + kColD4 + c0 + 32, v1);
                tmem_ld_wait();
                tc_fence_before();
                AL3D_TSE();
                // packed fp32 pairs (add / fma .f32x2): even channels accumulate in the low lane, odd ones in the high lane
                uint64_t a0e = 0, a1e = 0, a0o = 0, a1o = 0;      // logit 0 / 1 sums over the v0 (e) and v1 (o) column batches
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 bb0 = *reinterpret_cast<const float4 *>(s.bd4 + c0 + i);
                    const float4 bb1 = *reinterpret_cast<const float4 *>(s.bd4 + c0 + 32 + i);
                    const float4 wa0 = *reinterpret_cast<const float4 *>(s.w5 + c0 + i);
                    const float4 wb0 = *reinterpret_cast<const float4 *>(s.w5 + 128 + c0 + i);
                    const float4 wa1 = *reinterpret_cast<const float4 *>(s.w5 + c0 + 32 + i);
                    const float4 wb1 = *reinterpret_cast<const float4 *>(s.w5 + 128 + c0 + 32 + i);
                    uint64_t t;
                    t = f2_relu(f2_add(f2_pack(__uint_as_float(v0[i + 0]), __uint_as_float(v0[i + 1])), f2_pack(bb0.x, bb0.y)));
                    a0e = f2_fma(t, f2_pack(wa0.x, wa0.y), a0e); a1e = f2_fma(t, f2_pack(wb0.x, wb0.y), a1e);
                    t = f2_relu(f2_add(f2_pack(__uint_as_float(v0[i + 2]), __uint_as_float(v0[i + 3])), f2_pack(bb0.z, bb0.w)));
                    a0e = f2_fma(t, f2_pack(wa0.z, wa0.w), a0e); a1e = f2_fma(t, f2_pack(wb0.z, wb0.w), a1e);
                    t = f2_relu(f2_add(f2_pack(__uint_as_float(v1[i + 0]), __uint_as_float(v1[i + 1])), f2_pack(bb1.x, bb1.y)));
                    a0o = f2_fma(t, f2_pack(wa1.x, wa1.y), a0o); a1o = f2_fma(t, f2_pack(wb1.x, wb1.y), a1o);
                    t = f2_relu(f2_add(f2_pack(__uint_as_float(v1[i + 2]), __uint_as_float(v1[i + 3])), f2_pack(bb1.z, bb1.w)));
                    a0o = f2_fma(t, f2_pack(wa1.z, wa1.w), a0o); a1o = f2_fma(t, f2_pack(wb1.z, wb1.w), a1o);
                }
                const float l0a = f2_hsum(a0e), l1a = f2_hsum(a1e), l0b = f2_hsum(a0o), l1b = f2_hsum(a1o);
                const float l0 = l0a + l0b, l1 = l1a + l1b;
                if (half == 1) { s.lpart[par][row] = l0; s.lpart[par][kTile + row] = l1; }
                AL3D_TSE();
                asm volatile("bar.sync 2, 256;" ::: "memory");
                AL3D_TSE();
                if (half == 0 && valid) {
                    const float f0 = (s.b5[0] + l0) + s.lpart[par][row];
                    const float f1 = (s.b5[1] + l1) + s.lpart[par][kTile + row];
                    const int64_t o = (int64_t)b * p.n + pidx;
                    *reinterpret_cast<float2 *>(p.logits + o * 2) = make_float2(f0, f1);
                    p.mask[o] = (f0 < f1) ? 1 : 0;
                }
                AL3D_TSE();
            }
        }
#undef P2_PUBLISH
    }
done:
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                  // the pair's MMAs touch both CTAs: leave together
    if (warp == 0) tmem_dealloc_pair<512>(tmem);
}


// ================================================================================================
// seg_pass1_kernel -- ins_seg conv1..conv5 + max over points, tile PAIRS, warp-specialised pipeline
//
// A STEP is a pair of 128-point tiles (X, Y) of one (object, split) item.  conv5 is computed TRANSPOSED per tile
// (M = 128 channels, N = 128 points, K = 128; eight channel chunks), so the max over points is a per-thread reduction
// over TMEM columns and each streamed 16 KB conv5 block serves both tiles (256 points).  While the tensor pipe runs
// conv5 of step i, the front layers (conv1 on CUDA cores, conv2-4 with their A operand in TMEM and their 32 KB of
// weights resident in shared memory) of step i+1 run beside it into the other half of a double-buffered shared-memory
// operand -- nothing of the front sits on conv5's critical path.  Roles (384 threads):
//   warp 0      producer: resident front weights once, then the conv5 weight ring (cp.async.bulk)
//   warp 1, 10  conv5 issuers, tile X / tile Y (own accumulator, own barriers; they alternate on the pipe)
//   warp 11     front issuer: conv2, conv3, conv4 (two N = 64 halves) of the next step's two tiles, interleaved
//   warps 2-5   reducers: drain the conv5 accumulators (tcgen05.ld), running max in registers, atomicMax per object
//   warps 6-9   front epilogues: conv1, accumulator -> bf16 operand conversions, conv4 output -> shared memory
// Every epilogue thread owns one whole point row (TMEM lane) of what it touches, so there are no cross-warp hazards
// inside a TMEM region.  TMEM: [0,128) conv5 accumulator of X | [128,256) of Y | [256,384) front region of X' |
// [384,512) front region of Y'.  Front region (128 columns, R = its base):
//   R+96 A1 (conv1, bf16) -> conv2 -> R+0 D2 (64) -> R+0 A2 (32, in place) -> conv3 -> R+32 D3 (64) -> R+96 A3 (over
//   the dead A1) -> conv4 channels 0-63 -> R+0 D4a -> shared memory -> conv4 channels 64-127 -> R+0 D4b -> shared memory.
// ================================================================================================
struct Pass1Params {
    const float *x; int64_t sb, sc, sp; int bs, n; int c_in;
    const float *w1_w, *w1_b;          // conv1 fp32 (8, 64) transposed + padded, (64)
    const float *b2, *b3, *b4, *b5;    // conv2-5 biases (64),(64),(128),(1024)
    const uint8_t *wfront;             // conv2, conv3 (64 rows x 64 K, 8 KB each in 16 KB slots), conv4 (128 x 64)
    const uint8_t *w5stream;           // 16 blocks of 128 channels x 64 K, order (chunk, k-block)
    float *out;                        // (bs, 1024) zero-initialised
    int splits, n_items;
    long long *dbg;                    // optional clock64 timeline (al3d_set_debug_buffer), else NULL
    int dbg_skip;                      // first recorded step of the timeline (environment AL3D_DEBUG_SKIP)
    TcStatus wd;                       // watchdog status word of this launch
    // copies in the parameter block (constant bank): operands of the front warps' arithmetic, no loads
    float w1c[512], b1c[64], b2c[64], b3c[64], b4c[128];
};

constexpr int kP1Stages = 4;
constexpr uint32_t kP1FrontCol = 256;  // front regions: X' at 256, Y' at 384
struct Pass1Smem {
    uint8_t out4[2][16 * 4096];        // conv4 output of a tile pair: KP tile of 256 rows x 128 channels, double-buffered
    uint8_t wfront[32768];             // resident conv2 (8 KB) | conv3 (8 KB) | conv4 (16 KB) weights
    uint8_t wring[kP1Stages][kStageBytes];
    float w1_w[64 * 8], w1_b[64];     // conv2-4 biases are read through the read-only cache (no room here)
    uint64_t w_full[kP1Stages], w_empty[kP1Stages];
    uint64_t res_full;
    // front warps -> conv5 issuers (operand written) and back (operand consumed): ONE BARRIER PAIR PER out4 BUFFER.  A
    // parity wait must never be lapped (two completions before the waiter looks).  With a single out4_free the issuers
    // could complete the phases of steps s-2 AND s-1 before a slow front reached its wait for s-2: the parity aliased
    // and the pipeline deadlocked (round-1 watchdog 0x9100); a single out4_ready has the mirror-image hazard for a slow
    // issuer.  Per buffer, the next completion of either barrier needs the other side to have passed its wait first.
    uint64_t out4_ready[2], out4_free[2];
    uint64_t act[2], acc[2];           // front hand-overs per tile: operand in TMEM / accumulator complete
    uint64_t c5_full[2], c5_empty[2];  // conv5 accumulator of tile X / Y: complete / drained
    uint32_t tmem_base;
};
static_assert(sizeof(Pass1Smem) + 128 <= 232448, "Pass1Smem exceeds the 227 KB opt-in limit");

// Steps of a CTA: items blockIdx.x, blockIdx.x + gridDim.x, ...; within an item tiles t, t+2, ... (two tiles per step)
struct P1Step { int item, b, t, t1; bool valid; };
__device__ __forceinline__ void p1_item_range(const Pass1Params &p, int tiles_per_obj, int item, int &b, int &t0, int &t1)
{
    const int sp_i = item % p.splits;
    b = item / p.splits;
    t0 = (int)((int64_t)tiles_per_obj * sp_i / p.splits);
    t1 = (int)((int64_t)tiles_per_obj * (sp_i + 1) / p.splits);
}
__device__ __forceinline__ P1Step p1_seek(const Pass1Params &p, int tiles_per_obj, int stride, P1Step s)
{
    s.valid = false;
    while (s.item < p.n_items) {
        int t0;
        p1_item_range(p, tiles_per_obj, s.item, s.b, t0, s.t1);
        if (t0 < s.t1) { s.t = t0; s.valid = true; break; }
        s.item += stride;
    }
    return s;
}
__device__ __forceinline__ P1Step p1_first(const Pass1Params &p, int tiles_per_obj, int first, int stride)
{
    P1Step s; s.item = first; s.b = 0; s.t = 0; s.t1 = 0; s.valid = false;
    return p1_seek(p, tiles_per_obj, stride, s);
}
__device__ __forceinline__ P1Step p1_next(const Pass1Params &p, int tiles_per_obj, int stride, P1Step s)
{
    if (!s.valid) return s;
    s.t += 2;
    if (s.t < s.t1) return s;
    s.item += stride;
    return p1_seek(p, tiles_per_obj, stride, s);
}

constexpr int kP1Threads = kThreads + 64;
__global__ void __launch_bounds__(kP1Threads, 1)
seg_pass1_kernel(const Pass1Params p)
{
#define P1_TS(role)                                                                                   \
    do { const int ts_it = it_local - p.dbg_skip;                                                     \
         if (p.dbg && blockIdx.x == 0 && ts_it >= 0 && ts_it < 4 && ts_i < 64)                         \
             p.dbg[((role) * 4 + ts_it) * 64 + ts_i++] = clock64(); } while (0)
    const TcStatus &wd = p.wd;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    Pass1Smem &s = *reinterpret_cast<Pass1Smem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < 64 * 8; i += kP1Threads) s.w1_w[i] = p.w1_w[i];
    for (int i = threadIdx.x; i < 64; i += kP1Threads) s.w1_b[i] = p.w1_b[i];
    if (threadIdx.x == 0) {
        for (int i = 0; i < kP1Stages; ++i) { mbar_init(&s.w_full[i], 1); mbar_init(&s.w_empty[i], 2); }
        mbar_init(&s.res_full, 1);
        mbar_init(&s.out4_ready[0], 2 * 4); mbar_init(&s.out4_ready[1], 2 * 4);   // front warps x tiles, per out4 buffer
        mbar_init(&s.out4_free[0], 2); mbar_init(&s.out4_free[1], 2);   // the two conv5 issuers, per out4 buffer
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s.act[i], 4); mbar_init(&s.acc[i], 1);
            mbar_init(&s.c5_full[i], 1); mbar_init(&s.c5_empty[i], 4);
        }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<512>(&s.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;
    const int tiles_per_obj = (p.n + kTile - 1) / kTile;
    const int stride = (int)gridDim.x;
#define P1_ARRIVE(bar) do { __syncwarp(); if (lane == 0) mbar_arrive(bar); } while (0)

    if (warp == 0) {
        // ------------------------------------------------------------ producer: resident front weights, then the conv5 ring
        if (elect_one_sync()) {
            mbar_arrive_expect_tx(&s.res_full, 8192 + 8192 + 16384);
            bulk_g2s(s.wfront, p.wfront, 8192, &s.res_full);
            bulk_g2s(s.wfront + 8192, p.wfront + kStageBytes, 8192, &s.res_full);
            bulk_g2s(s.wfront + 16384, p.wfront + 2 * kStageBytes, 16384, &s.res_full);
            int stage = 0; uint32_t phase = 0;
            for (P1Step st = p1_first(p, tiles_per_obj, blockIdx.x, stride); st.valid; st = p1_next(p, tiles_per_obj, stride, st)) {
                for (int blk = 0; blk < 16; ++blk) {
                    AL3D_STRESS(wd, 0x91);
                    if (!mbar_wait(&s.w_empty[stage], phase ^ 1, 0x9100 + stage, wd)) goto done;
                    mbar_arrive_expect_tx(&s.w_full[stage], kStageBytes);
                    bulk_g2s(s.wring[stage], p.w5stream + (size_t)blk * kStageBytes, kStageBytes, &s.w_full[stage]);
                    if (++stage == kP1Stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1 || warp == kThreads / 32) {
        // ------------------------------------------------------------ conv5 issuers: warp 1 -> tile X, warp 10 -> tile Y
        if (elect_one_sync()) {
            const int q = (warp == 1) ? 0 : 1;
            uint32_t o4_phase[2] = {0, 0}, ce_phase = 0;
            const uint32_t id128 = make_idesc_bf16(128, 128);
            const uint32_t o4[2] = {smem_u32(s.out4[0]), smem_u32(s.out4[1])};
            const uint32_t d = tmem + q * 128;
            int ob = 0, g = 0, it_local = 0, ts_i = 0;
            for (P1Step st = p1_first(p, tiles_per_obj, blockIdx.x, stride); st.valid; st = p1_next(p, tiles_per_obj, stride, st), ++it_local) {
                ts_i = 0;
                if (q == 0) P1_TS(0);
                AL3D_STRESS(wd, 0x94 + q);
                if (!mbar_wait(&s.out4_ready[ob], o4_phase[ob], 0x9400 + q, wd)) goto done;
                o4_phase[ob] ^= 1; tc_fence_after();
                if (q == 0) P1_TS(0);
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {
                    AL3D_STRESS(wd, 0x95 + q);
                    if (!mbar_wait(&s.c5_empty[q], ce_phase ^ 1, 0x9500 + q, wd)) goto done;
                    ce_phase ^= 1; tc_fence_after();
                    if (q == 0) P1_TS(0);
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb, ++g) {
                        const int stage = g % kP1Stages;
                        if (!mbar_wait(&s.w_full[stage], (uint32_t)(g / kP1Stages) & 1u, 0x9600 + stage, wd)) goto done;
                        tc_fence_after();
                        const uint32_t wa = smem_u32(s.wring[stage]);
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)
                            mma_bf16(d, make_desc(wa + k4 * 4096, 128),
                                     make_desc_lbo(o4[ob] + q * 2048 + (kb * 8 + k4 * 2) * 4096, 4096), id128, (kb > 0 || k4 > 0) ? 1u : 0u);
                        mma_commit(&s.w_empty[stage]);           // the stage is free once BOTH tiles' issuers have consumed it
                    }
                    mma_commit(&s.c5_full[q]);
                    if (q == 0) P1_TS(0);
                }
                mma_commit(&s.out4_free[(wd.trap & 2) ? 0 : ob]); // this tile's reads of out4[ob] are done (bit 1 of
                                                                  // wd.trap: test hook, the round-1 single-barrier protocol)
                ob ^= 1;
            }
        }
    } else if (warp == kThreads / 32 + 1) {
        // ------------------------------------------------------------ front issuer: conv2, conv3, conv4a, conv4b of both
        // tiles of every step, interleaved (X, Y, X, Y ...) as the front warps hand the operands over
        if (elect_one_sync()) {
            uint32_t act_phase[2] = {0, 0};
            const uint32_t id64 = make_idesc_bf16(128, 64);
            const uint32_t wf = smem_u32(s.wfront);
            if (!mbar_wait(&s.res_full, 0, 0x9200, wd)) goto done;
            tc_fence_after();
            for (P1Step st = p1_first(p, tiles_per_obj, blockIdx.x, stride); st.valid; st = p1_next(p, tiles_per_obj, stride, st)) {
#pragma unroll 1
                for (int l = 0; l < 4; ++l) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const uint32_t R = tmem + kP1FrontCol + q * 128;
                        AL3D_STRESS(wd, 0x93);
                        if (!mbar_wait(&s.act[q], act_phase[q], 0x9300 + l * 2 + q, wd)) goto done;
                        act_phase[q] ^= 1; tc_fence_after();
                        if (l == 0)      mma_ts_k64(R, R + 96, R + 104, R + 112, R + 120, wf, 64, id64, false);              // conv2
                        else if (l == 1) mma_ts_k64(R + 32, R, R + 8, R + 16, R + 24, wf + 8192, 64, id64, false);          // conv3
                        else {
                            // conv4, output channels (l-2)*64 .. +64: rows of the 128-row KP tile (plane stride 2048 B)
                            const uint32_t wb = wf + 16384 + (l - 2) * 64 * 16;
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                mma_bf16_ts(R, R + 96 + k4 * 8, make_desc_lbo(wb + k4 * 4096, 2048), id64, k4 > 0 ? 1u : 0u);
                        }
                        mma_commit(&s.acc[q]);
                    }
                }
            }
        }
    } else if (warp < 6) {
        // ------------------------------------------------------------ reducers (warps 2-5): thread = channel row of a chunk.
        // (Tried and slower: separate reducer warps per tile -- the two conv5 streams then finish and drain together, or, when
        // forced to alternate, hold their weight stages too long for the 4-stage ring; all four tcgen05.ld of a tile in flight
        // at once -- the drain is bound by TMEM bandwidth under the running MMAs, not by load latency.)
        const int row = epi_row();
        const uint32_t tl = tmem + ((uint32_t)(row & ~31) << 16);
        uint32_t cf_phase[2] = {0, 0};
        int it_local = 0, ts_i = 0;
        const bool ts_on = (threadIdx.x == 64);
        float rmax[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rmax[i] = -INFINITY;
        P1Step cur = p1_first(p, tiles_per_obj, blockIdx.x, stride);
        while (cur.valid) {
            const P1Step nxt = p1_next(p, tiles_per_obj, stride, cur);
            ts_i = 0;
            if (ts_on) P1_TS(1);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    AL3D_STRESS_WARP(wd, 0x83);
                    if (!mbar_wait(&s.c5_full[q], cf_phase[q], 0x8300 + c * 2 + q, wd)) goto done;
                    cf_phase[q] ^= 1; tc_fence_after();
                    if (ts_on) P1_TS(1);
                    float m0 = rmax[c], m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        uint32_t v0[32], v1[32];
                        const uint32_t ta = tl + q * 128 + hh * 64;
                        tmem_ld32(ta, v0);
                        tmem_ld32(ta + 32, v1);
                        tmem_ld_wait();
                        if (hh == 1) { tc_fence_before(); P1_ARRIVE(&s.c5_empty[q]); }     // all 128 values are in registers
#pragma unroll
                        for (int i = 0; i < 32; i += 8) {
                            m0 = fmax3(m0, __uint_as_float(v0[i]), __uint_as_float(v0[i + 1]));
                            m1 = fmax3(m1, __uint_as_float(v0[i + 2]), __uint_as_float(v0[i + 3]));
                            m2 = fmax3(m2, __uint_as_float(v0[i + 4]), __uint_as_float(v0[i + 5]));
                            m3 = fmax3(m3, __uint_as_float(v0[i + 6]), __uint_as_float(v0[i + 7]));
                            m0 = fmax3(m0, __uint_as_float(v1[i]), __uint_as_float(v1[i + 1]));
                            m1 = fmax3(m1, __uint_as_float(v1[i + 2]), __uint_as_float(v1[i + 3]));
                            m2 = fmax3(m2, __uint_as_float(v1[i + 4]), __uint_as_float(v1[i + 5]));
                            m3 = fmax3(m3, __uint_as_float(v1[i + 6]), __uint_as_float(v1[i + 7]));
                        }
                    }
                    rmax[c] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                    if (ts_on) P1_TS(1);
                }
            }
            if (!nxt.valid || nxt.item != cur.item) {
                // ---- item finished: relu(max + bias) >= 0, so integer atomicMax on the bit pattern is exact
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int ch = c * 128 + row;
                    const float v = fmaxf(rmax[c] + __ldg(p.b5 + ch), 0.f);
                    atomicMax(reinterpret_cast<int *>(p.out + (int64_t)cur.b * 1024 + ch), __float_as_int(v));
                    rmax[c] = -INFINITY;
                }
            }
            cur = nxt;
            ++it_local;
        }
    } else {
        // ------------------------------------------------------------ front warps (6-9): thread = point row of both tiles
        const int row = epi_row();
        const uint32_t tl = tmem + ((uint32_t)(row & ~31) << 16) + kP1FrontCol;
        uint32_t acc_phase[2] = {0, 0}, of_phase[2] = {0, 0};
        int it_local = 0, ts_i = 0;
        const bool ts_on = (threadIdx.x == 6 * 32);
        float xq[2][8];
        auto load_pair = [&](const P1Step &st) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int tq = (st.t + q < st.t1) ? st.t + q : st.t1 - 1;          // odd tail: repeat the last tile (max is idempotent)
                int pidx = tq * kTile + row;
                if (pidx > p.n - 1) pidx = p.n - 1;
                const float *px = p.x + (int64_t)st.b * p.sb + (int64_t)pidx * p.sp;
#pragma unroll
                for (int c = 0; c < 8; ++c) xq[q][c] = (c < p.c_in) ? __ldg(px + c * p.sc) : 0.f;
            }
        };
        P1Step cur = p1_first(p, tiles_per_obj, blockIdx.x, stride);
        if (cur.valid) load_pair(cur);
        int step = 0;
        while (cur.valid) {
            const P1Step nxt = p1_next(p, tiles_per_obj, stride, cur);
            const int buf = step & 1;
            ts_i = 0;
            if (ts_on) P1_TS(2);
            // ---- conv1 (CUDA cores, 64 channels of this point) -> A1
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                uint32_t o0[16], o1[16];
                if (p.c_in == 3)      { conv1_pack32_cb<3, 0>(xq[q], p.w1c, p.b1c, o0); conv1_pack32_cb<3, 32>(xq[q], p.w1c, p.b1c, o1); }
                else if (p.c_in == 4) { conv1_pack32_cb<4, 0>(xq[q], p.w1c, p.b1c, o0); conv1_pack32_cb<4, 32>(xq[q], p.w1c, p.b1c, o1); }
                else                  { conv1_pack32(xq[q], p.c_in, s.w1_w, s.w1_b, 0, o0); conv1_pack32(xq[q], p.c_in, s.w1_w, s.w1_b, 32, o1); }
                tmem_st16(tl + q * 128 + 96, o0);
                tmem_st16(tl + q * 128 + 112, o1);
                tmem_st_wait(); tc_fence_before();
                P1_ARRIVE(&s.act[q]);
            }
            if (ts_on) P1_TS(2);
            if (nxt.valid) load_pair(nxt);                                 // prefetch the next step's points
            // ---- conv2 / conv3 epilogues: 64 accumulator columns -> bias + ReLU -> bf16 operand of the next layer
#pragma unroll
            for (int l = 0; l < 2; ++l) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    AL3D_STRESS_WARP(wd, 0x81);
                    if (!mbar_wait(&s.acc[q], acc_phase[q], 0x8100 + l * 2 + q, wd)) goto done;
                    acc_phase[q] ^= 1; tc_fence_after();
                    uint32_t v0[32], v1[32], o0[16], o1[16];
                    const uint32_t R = tl + q * 128;
                    tmem_ld32(R + (l == 0 ? 0 : 32), v0);
                    tmem_ld32(R + (l == 0 ? 0 : 32) + 32, v1);
                    tmem_ld_wait();
                    if (l == 0) { pack_act32_cb<0>(v0, p.b2c, o0); pack_act32_cb<32>(v1, p.b2c, o1); }
                    else        { pack_act32_cb<0>(v0, p.b3c, o0); pack_act32_cb<32>(v1, p.b3c, o1); }
                    tmem_st16(R + (l == 0 ? 0 : 96), o0);
                    tmem_st16(R + (l == 0 ? 0 : 96) + 16, o1);
                    tmem_st_wait(); tc_fence_before();
                    P1_ARRIVE(&s.act[q]);
                    if (ts_on) P1_TS(2);
                }
            }
            // ---- conv4 epilogues (two 64-channel halves): -> shared-memory operand of conv5, rows of tile q.
            //      out4[buf] was last read by conv5 two steps ago: wait until both conv5 issuers have released it.
            AL3D_STRESS_WARP(wd, 0x84);
            if (step >= 2) {
                const int fb = (wd.trap & 2) ? 0 : buf;
                if (!mbar_wait(&s.out4_free[fb], of_phase[fb], 0x8400 + fb, wd)) goto done;
                of_phase[fb] ^= 1;
            }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    AL3D_STRESS_WARP(wd, 0x82);
                    if (!mbar_wait(&s.acc[q], acc_phase[q], 0x8200 + hh * 2 + q, wd)) goto done;
                    acc_phase[q] ^= 1; tc_fence_after();
                    uint32_t v0[32], v1[32], o0[16], o1[16];
                    const uint32_t R = tl + q * 128;
                    tmem_ld32(R, v0);
                    tmem_ld32(R + 32, v1);
                    tmem_ld_wait();
                    tc_fence_before();
                    if (hh == 0) P1_ARRIVE(&s.act[q]);                     // D4a is in registers: conv4b may overwrite it
                    if (hh == 0) { pack_act32_cb<0>(v0, p.b4c, o0); pack_act32_cb<32>(v1, p.b4c, o1); }
                    else         { pack_act32_cb<64>(v0, p.b4c, o0); pack_act32_cb<96>(v1, p.b4c, o1); }
                    uint8_t *dst = s.out4[buf] + (size_t)(hh * 8) * 4096 + (size_t)(q * kTile + row) * 16;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        *reinterpret_cast<uint4 *>(dst + (size_t)j * 4096) = make_uint4(o0[4 * j], o0[4 * j + 1], o0[4 * j + 2], o0[4 * j + 3]);
                        *reinterpret_cast<uint4 *>(dst + (size_t)(4 + j) * 4096) = make_uint4(o1[4 * j], o1[4 * j + 1], o1[4 * j + 2], o1[4 * j + 3]);
                    }
                    if (hh == 1) { fence_proxy_async_smem(); P1_ARRIVE(&s.out4_ready[buf]); }
                    if (ts_on) P1_TS(2);
                }
            }
            cur = nxt;
            ++step; ++it_local;
        }
    }
#undef P1_ARRIVE
#undef P1_TS
done:
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// ================================================================================================
// trunk_pair_kernel -- the box-head / embedding trunks: conv1 (CUDA cores) -> conv2 -> conv3 -> conv4 + max over points
// (tools/static_model.py:330-334, tools/dynamic_model.py:241-245, 278-282), same ideas as seg_pass1_kernel:
// tiles are processed in PAIRS; conv2 / conv3 read their A operand from TMEM (conv2's output is converted in place);
// conv3's output of both tiles goes to one 256-row shared-memory operand and conv4 is computed transposed
// (M = 128 channels, N = 256 points), so every streamed conv4 block serves 256 points and the max over points is a
// per-thread reduction.  All weights stream through one ring (conv2 / conv3 blocks twice per pair, conv4 once).
// TMEM (front, one tile at a time): [0,W0/2) A1 | [64,64+M1) D(conv2), its bf16 image in place | [256,256+M2) D(conv3);
// conv4 then double-buffers 2 x 256 columns.
// ================================================================================================
struct TrunkParams {
    const float *x; int64_t sb, sc, sp; int bs, n; int c_in;
    const float *w0_w, *w0_b, *mid_b, *last_b;
    const uint8_t *wstream;            // 16 KB slots: conv2 (k-blocks) | conv3 (row-chunk, k-block) | conv4 (chunk, k-block)
    float *out;                        // (bs, 512) zero-initialised
    int splits, n_items;
    TcStatus wd;                       // watchdog status word of this launch
};

template <int W0, int M1, int M2>
struct TrunkCfg {
    static constexpr int kNkb2 = W0 / 64, kRows2 = M1, kNc3 = M2 / 128, kNkb3 = M1 / 64, kNkb4 = M2 / 64;
    static constexpr int kFrontBlocks = kNkb2 + kNc3 * kNkb3, kLastBlocks = 4 * kNkb4;
    static constexpr int kOut3Bytes = (M2 / 8) * 4096;                 // KP tile of 256 rows x M2 channels
    static constexpr int kStages = (232448 - 1024 - kOut3Bytes - 4096 - (W0 * 9 + M1 + M2) * 4) / kStageBytes;
    static constexpr uint32_t kColA1 = 0, kColD2 = 64, kColD3 = 256;
};

template <int W0, int M1, int M2>
__global__ void __launch_bounds__(kThreads, 1)
trunk_pair_kernel(const TrunkParams p)
{
    const TcStatus &wd = p.wd;
    using C = TrunkCfg<W0, M1, M2>;
    constexpr int S = C::kStages;
    static_assert(S >= 3 && S <= 12, "weight ring depth");
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *const s_out3 = smem_raw;
    uint8_t *const s_ring = smem_raw + C::kOut3Bytes;
    float *const s_w0w = reinterpret_cast<float *>(s_ring + (size_t)S * kStageBytes);
    float *const s_w0b = s_w0w + W0 * 8;
    float *const s_b2 = s_w0b + W0;
    float *const s_b3 = s_b2 + M1;
    uint64_t *const bars = reinterpret_cast<uint64_t *>(s_b3 + M2);
    uint64_t *const w_full = bars, *const w_empty = bars + 12;
    uint64_t *const act = bars + 24, *const acc = bars + 25, *const out3_ready = bars + 26;
    uint64_t *const last_full = bars + 27, *const last_empty = bars + 29;
    uint32_t *const tmem_slot = reinterpret_cast<uint32_t *>(bars + 31);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < W0 * 8; i += kThreads) s_w0w[i] = p.w0_w[i];
    for (int i = threadIdx.x; i < W0; i += kThreads) s_w0b[i] = p.w0_b[i];
    for (int i = threadIdx.x; i < M1 + M2; i += kThreads) s_b2[i] = p.mid_b[i];
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        mbar_init(act, kEpiThreads / 32); mbar_init(acc, 1);
        mbar_init(out3_ready, 2 * kEpiThreads / 32);
        for (int i = 0; i < 2; ++i) { mbar_init(&last_full[i], 1); mbar_init(&last_empty[i], kEpiThreads / 32); }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int tiles_per_obj = (p.n + kTile - 1) / kTile;

    if (warp == 0) {
        // ------------------------------------------------------------ weight producer
        if (elect_one_sync()) {
            int stage = 0; uint32_t phase = 0;
            auto push = [&](int blk, uint32_t bytes) -> bool {
                AL3D_STRESS(wd, 0x71);
                if (!mbar_wait(&w_empty[stage], phase ^ 1, 0x7100 + stage, wd)) return false;
                mbar_arrive_expect_tx(&w_full[stage], bytes);
                bulk_g2s(s_ring + (size_t)stage * kStageBytes, p.wstream + (size_t)blk * kStageBytes, bytes, &w_full[stage]);
                if (++stage == S) { stage = 0; phase ^= 1; }
                return true;
            };
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const int sp_i = item % p.splits;
                const int t0 = (int)((int64_t)tiles_per_obj * sp_i / p.splits), t1 = (int)((int64_t)tiles_per_obj * (sp_i + 1) / p.splits);
                for (int t = t0; t < t1; t += 2) {
                    for (int q = 0; q < 2; ++q) {
                        for (int b = 0; b < C::kNkb2; ++b) if (!push(b, C::kRows2 * 128)) goto done;
                        for (int b = 0; b < C::kNc3 * C::kNkb3; ++b) if (!push(C::kNkb2 + b, kStageBytes)) goto done;
                    }
                    for (int b = 0; b < C::kLastBlocks; ++b) if (!push(C::kFrontBlocks + b, kStageBytes)) goto done;
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (elect_one_sync()) {
            int stage = 0; uint32_t wphase = 0, act_phase = 0, o3_phase = 0, le_phase[2] = {0, 0};
            const uint32_t id2 = make_idesc_bf16(128, M1), id128 = make_idesc_bf16(128, 128), id256 = make_idesc_bf16(128, 256);
            const uint32_t a_out3 = smem_u32(s_out3), ring = smem_u32(s_ring);
#define TK_NEXT_W(code) if (!mbar_wait(&w_full[stage], wphase, code, wd)) goto done; tc_fence_after(); const uint32_t wb_ = ring + stage * kStageBytes;
#define TK_REL_W() mma_commit(&w_empty[stage]); if (++stage == S) { stage = 0; wphase ^= 1; }
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const int sp_i = item % p.splits;
                const int t0 = (int)((int64_t)tiles_per_obj * sp_i / p.splits), t1 = (int)((int64_t)tiles_per_obj * (sp_i + 1) / p.splits);
                for (int t = t0; t < t1; t += 2) {
                    for (int q = 0; q < 2; ++q) {
                        // conv2: A1 (TMEM) x W2 -> D2
                        AL3D_STRESS(wd, 0x72);
                        if (!mbar_wait(act, act_phase, 0x7200, wd)) goto done;
                        act_phase ^= 1; tc_fence_after();
#pragma unroll
                        for (int kb = 0; kb < C::kNkb2; ++kb) {
                            TK_NEXT_W(0x7300)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                mma_bf16_ts(tmem + C::kColD2, tmem + C::kColA1 + (kb * 4 + k) * 8, make_desc(wb_ + k * 2 * C::kRows2 * 16, C::kRows2),
                                            id2, (kb > 0 || k > 0) ? 1u : 0u);
                            TK_REL_W()
                        }
                        mma_commit(acc);
                        // conv3: A2 (bf16 image of D2, in place: channels < M1/2 at +0, the rest at +M1/2) x W3 -> D3
                        if (!mbar_wait(act, act_phase, 0x7201, wd)) goto done;
                        act_phase ^= 1; tc_fence_after();
#pragma unroll
                        for (int nc = 0; nc < C::kNc3; ++nc)
#pragma unroll
                            for (int kb = 0; kb < C::kNkb3; ++kb) {
                                TK_NEXT_W(0x7310)
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const int k0 = (kb * 4 + k) * 16;                      // first channel of this K step
                                    const int hf = k0 / (M1 / 2);
                                    const uint32_t acol = tmem + C::kColD2 + hf * (M1 / 2) + (k0 - hf * (M1 / 2)) / 2;
                                    mma_bf16_ts(tmem + C::kColD3 + nc * 128, acol, make_desc(wb_ + k * 2 * 128 * 16, 128), id128, (kb > 0 || k > 0) ? 1u : 0u);
                                }
                                TK_REL_W()
                            }
                        mma_commit(acc);
                    }
                    // conv4, transposed, N = 256 points (both tiles)
                    if (!mbar_wait(out3_ready, o3_phase, 0x7400, wd)) goto done;
                    o3_phase ^= 1; tc_fence_after();
#pragma unroll 1
                    for (int cc = 0; cc < 4; ++cc) {
                        const int b = cc & 1;
                        AL3D_STRESS(wd, 0x75);
                        if (!mbar_wait(&last_empty[b], le_phase[b] ^ 1, 0x7500 + b, wd)) goto done;
                        le_phase[b] ^= 1; tc_fence_after();
                        for (int kb = 0; kb < C::kNkb4; ++kb) {
                            TK_NEXT_W(0x7600)
                            mma_block_k64(tmem + b * 256, wb_, 128, a_out3 + kb * 8 * 4096, 256, id256, kb > 0);
                            TK_REL_W()
                        }
                        mma_commit(&last_full[b]);
                    }
                }
            }
#undef TK_NEXT_W
#undef TK_REL_W
        }
    } else {
        // ------------------------------------------------------------ epilogue warps (256 threads)
        const int row = epi_row(), half = epi_half();
        const uint32_t tl = tmem + ((uint32_t)(row & ~31) << 16);
        uint32_t acc_phase = 0, lf_phase[2] = {0, 0};
#define TK_ARRIVE(bar) do { __syncwarp(); if (lane == 0) mbar_arrive(bar); } while (0)
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            const int b = item / p.splits, sp_i = item % p.splits;
            const int t0 = (int)((int64_t)tiles_per_obj * sp_i / p.splits), t1 = (int)((int64_t)tiles_per_obj * (sp_i + 1) / p.splits);
            float rmax[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) rmax[i] = -INFINITY;
            for (int t = t0; t < t1; t += 2) {
#pragma unroll 1
                for (int q = 0; q < 2; ++q) {
                    // ---- conv1 on CUDA cores: this thread's W0/2 channels -> A1 (an odd tail pair repeats its tile)
                    {
                        const int tq = (t + q < t1) ? t + q : t1 - 1;
                        int pidx = tq * kTile + row;
                        if (pidx > p.n - 1) pidx = p.n - 1;
                        const float *px = p.x + (int64_t)b * p.sb + (int64_t)pidx * p.sp;
                        float xv[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) xv[c] = (c < p.c_in) ? __ldg(px + c * p.sc) : 0.f;
#pragma unroll
                        for (int g = 0; g < W0 / 64; ++g) {
                            uint32_t o[16];
                            const int ch0 = half * (W0 / 2) + g * 32;
                            conv1_pack32<W0>(xv, p.c_in, s_w0w, s_w0b, ch0, o);
                            tmem_st16(tl + C::kColA1 + ch0 / 2, o);
                        }
                        tmem_st_wait(); tc_fence_before();
                        TK_ARRIVE(act);
                    }
                    // ---- conv2 epilogue: D2 -> its bf16 image, in place (this thread: M1/2 columns)
                    AL3D_STRESS_WARP(wd, 0x61);
                    if (!mbar_wait(acc, acc_phase, 0x6100, wd)) goto done;
                    acc_phase ^= 1; tc_fence_after();
#pragma unroll
                    for (int g = 0; g < M1 / 64; ++g) {
                        uint32_t v[32], o[16];
                        const uint32_t src = tl + C::kColD2 + half * (M1 / 2) + g * 32;
                        tmem_ld32(src, v);
                        tmem_ld_wait();
                        pack_act32(v, s_b2 + half * (M1 / 2) + g * 32, o);
                        tmem_st16(tl + C::kColD2 + half * (M1 / 2) + g * 16, o);
                    }
                    tmem_st_wait(); tc_fence_before();
                    TK_ARRIVE(act);
                    // ---- conv3 epilogue: D3 -> shared-memory operand of conv4 (this thread: M2/2 channels of its row)
                    AL3D_STRESS_WARP(wd, 0x62);
                    if (!mbar_wait(acc, acc_phase, 0x6101, wd)) goto done;
                    acc_phase ^= 1; tc_fence_after();
#pragma unroll
                    for (int g = 0; g < M2 / 64; ++g) {
                        uint32_t v[32], o[16];
                        const int c0 = half * (M2 / 2) + g * 32;
                        tmem_ld32(tl + C::kColD3 + c0, v);
                        tmem_ld_wait();
                        pack_act32(v, s_b3 + c0, o);
                        uint8_t *dst = s_out3 + (size_t)(c0 / 8) * 4096 + (size_t)(q * kTile + row) * 16;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            *reinterpret_cast<uint4 *>(dst + (size_t)j * 4096) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                    }
                    tc_fence_before();
                    fence_proxy_async_smem();
                    TK_ARRIVE(out3_ready);
                }
                // ---- conv4: this thread owns channel (cc*128 + row) and 128 of the pair's 256 points
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const int bsel = cc & 1;
                    AL3D_STRESS_WARP(wd, 0x63);
                    if (!mbar_wait(&last_full[bsel], lf_phase[bsel], 0x6300 + cc, wd)) goto done;
                    lf_phase[bsel] ^= 1; tc_fence_after();
                    float m0 = rmax[cc], m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
                    for (int c0 = 0; c0 < 128; c0 += 64) {
                        uint32_t v0[32], v1[32];
                        const uint32_t ta = tl + bsel * 256 + half * 128 + c0;
                        tmem_ld32(ta, v0);
                        tmem_ld32(ta + 32, v1);
                        tmem_ld_wait();
                        if (c0 == 64) { tc_fence_before(); TK_ARRIVE(&last_empty[bsel]); }   // all values are in registers
#pragma unroll
                        for (int i = 0; i < 32; i += 8) {
                            m0 = fmax3(m0, __uint_as_float(v0[i]), __uint_as_float(v0[i + 1]));
                            m1 = fmax3(m1, __uint_as_float(v0[i + 2]), __uint_as_float(v0[i + 3]));
                            m2 = fmax3(m2, __uint_as_float(v0[i + 4]), __uint_as_float(v0[i + 5]));
                            m3 = fmax3(m3, __uint_as_float(v0[i + 6]), __uint_as_float(v0[i + 7]));
                            m0 = fmax3(m0, __uint_as_float(v1[i]), __uint_as_float(v1[i + 1]));
                            m1 = fmax3(m1, __uint_as_float(v1[i + 2]), __uint_as_float(v1[i + 3]));
                            m2 = fmax3(m2, __uint_as_float(v1[i + 4]), __uint_as_float(v1[i + 5]));
                            m3 = fmax3(m3, __uint_as_float(v1[i + 6]), __uint_as_float(v1[i + 7]));
                        }
                    }
                    rmax[cc] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                }
                // The two warps of a lane quarter split the accumulator COLUMNS, and the next pair's conv1 writes A1 into
                // columns [0, W0/2) of the same lanes: a warp that ran ahead into the next pair would overwrite conv4
                // columns its sibling has not read yet (found by the delay-injection stress test).  All epilogue warps
                // leave the pair together.
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            if (t1 > t0) {
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const int ch = cc * 128 + row;
                    const float v = fmaxf(rmax[cc] + __ldg(p.last_b + ch), 0.f);
                    atomicMax(reinterpret_cast<int *>(p.out + (int64_t)b * 512 + ch), __float_as_int(v));
                }
            }
        }
#undef TK_ARRIVE
    }
done:
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int W0, int M1, int M2>
static int launch_trunk_pair(const TrunkParams &p0, int bs, int n, cudaStream_t stream, int sms)
{
    using C = TrunkCfg<W0, M1, M2>;
    TrunkParams p = p0;
    const int tiles = (n + kTile - 1) / kTile;
    int splits = 1;
    if (bs < 2 * sms) splits = (int)std::min<int64_t>((tiles + 1) / 2, ceil_div(2 * sms, bs));
    p.splits = std::max(splits, 1);
    p.n_items = bs * p.splits;
    const int grid = std::min(p.n_items, sms);
    const size_t smem = (size_t)C::kOut3Bytes + (size_t)C::kStages * kStageBytes + (size_t)(W0 * 9 + M1 + M2) * 4 + 32 * 8 + 128;
    AL3D_CHECK_CUDA(cudaFuncSetAttribute(trunk_pair_kernel<W0, M1, M2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    trunk_pair_kernel<W0, M1, M2><<<grid, kThreads, smem, stream>>>(p);
    AL3D_CHECK_LAUNCH("trunk_pair_kernel");
    return 0;
}

}  // namespace al3d

using namespace al3d;

static long long *g_debug_buffer = nullptr;
extern "C" int al3d_set_debug_buffer(void *dev_ptr) { g_debug_buffer = (long long *)dev_ptr; return 0; }
static int debug_skip()
{
    const char *e = g_debug_buffer ? std::getenv("AL3D_DEBUG_SKIP") : nullptr;
    return e ? std::atoi(e) : 0;
}

extern "C" int al3d_umma_selftest(const void *a_kp, const void *b_kp, int N, int K, float *d_out, int swap_lbo_sbo, void *stream)
{
    AL3D_CHECK_ARG(a_kp && b_kp && d_out, "al3d_umma_selftest: null pointer");
    AL3D_CHECK_ARG(N >= 16 && N <= 256 && N % 16 == 0 && K >= 16 && K % 16 == 0, "al3d_umma_selftest: bad N=%d K=%d", N, K);
    const size_t smem = (size_t)(128 + N) * K * 2;
    AL3D_CHECK_ARG(smem <= 200 * 1024, "al3d_umma_selftest: tile too large");
    AL3D_CHECK_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TcStatus wd;
    if (tc_launch_status(&wd)) return 1;
    umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const uint8_t *)a_kp, (const uint8_t *)b_kp, N, K, d_out, swap_lbo_sbo, wd);
    AL3D_CHECK_LAUNCH("umma_selftest_kernel");
    return 0;
}

extern "C" int al3d_umma_selftest_ts(const float *a, const void *b_kp, int N, int K, float *d_out, void *stream)
{
    AL3D_CHECK_ARG(a && b_kp && d_out, "al3d_umma_selftest_ts: null pointer");
    AL3D_CHECK_ARG(N >= 16 && N <= 256 && N % 16 == 0 && K >= 32 && K % 32 == 0 && K <= 512, "al3d_umma_selftest_ts: bad N=%d K=%d", N, K);
    const size_t smem = (size_t)N * K * 2;
    AL3D_CHECK_ARG(smem <= 200 * 1024, "al3d_umma_selftest_ts: tile too large");
    AL3D_CHECK_CUDA(cudaFuncSetAttribute(umma_selftest_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TcStatus wd;
    if (tc_launch_status(&wd)) return 1;
    umma_selftest_ts_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a, (const uint8_t *)b_kp, N, K, d_out, wd);
    AL3D_CHECK_LAUNCH("umma_selftest_ts_kernel");
    return 0;
}

extern "C" int al3d_umma_selftest_pair(const float *a, const void *b_kp_halves, int N, int K, float *d_out, void *stream)
{
    AL3D_CHECK_ARG(a && b_kp_halves && d_out, "al3d_umma_selftest_pair: null pointer");
    AL3D_CHECK_ARG(N >= 32 && N <= 256 && N % 32 == 0 && K >= 32 && K % 32 == 0 && K <= 512, "al3d_umma_selftest_pair: bad N=%d K=%d", N, K);
    const size_t smem = (size_t)(N / 2) * K * 2;
    AL3D_CHECK_CUDA(cudaFuncSetAttribute(umma_selftest_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TcStatus wd;
    if (tc_launch_status(&wd)) return 1;
    umma_selftest_pair_kernel<<<2, 128, smem, (cudaStream_t)stream>>>(a, (const uint8_t *)b_kp_halves, N, K, d_out, wd);
    AL3D_CHECK_LAUNCH("umma_selftest_pair_kernel");
    return 0;
}

extern "C" int al3d_umma_selftest_pair_ss(const void *a_kp_halves, const void *b_kp_halves, int K, float *d_out, void *stream)
{
    AL3D_CHECK_ARG(a_kp_halves && b_kp_halves && d_out, "al3d_umma_selftest_pair_ss: null pointer");
    AL3D_CHECK_ARG(K >= 16 && K % 16 == 0 && K <= 256, "al3d_umma_selftest_pair_ss: bad K=%d", K);
    const size_t smem = (size_t)2 * 128 * K * 2;
    AL3D_CHECK_CUDA(cudaFuncSetAttribute(umma_selftest_pair_ss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TcStatus wd;
    if (tc_launch_status(&wd)) return 1;
    umma_selftest_pair_ss_kernel<<<2, 128, smem, (cudaStream_t)stream>>>((const uint8_t *)a_kp_halves, (const uint8_t *)b_kp_halves, K, d_out, wd);
    AL3D_CHECK_LAUNCH("umma_selftest_pair_ss_kernel");
    return 0;
}

static int num_sms() { return tc_num_sms(); }      // of the CURRENT device (cached per device ordinal)

extern "C" int al3d_chain_maxpool_bf16(const al3d_chain_weights *w, const float *x, int64_t sb, int64_t sc, int64_t sp,
                                       int bs, int n, float *out, void *stream)
{
    AL3D_CHECK_ARG(w && x && out, "al3d_chain_maxpool_bf16: null pointer");
    AL3D_CHECK_ARG(w->c_in >= 1 && w->c_in <= 8, "al3d_chain_maxpool_bf16: c_in=%d", w->c_in);
    AL3D_CHECK_ARG(w->w0 == 64 || w->w0 == 128, "al3d_chain_maxpool_bf16: w0=%d must be 64 or 128", w->w0);
    AL3D_CHECK_ARG(w->n_mid == 2 || w->n_mid == 3, "al3d_chain_maxpool_bf16: n_mid=%d", w->n_mid);
    int prev = w->w0, tot = 0;
    for (int l = 0; l < w->n_mid; ++l) {
        const int N = w->mid[l];
        AL3D_CHECK_ARG(N == 64 || N == 128 || N == 256, "al3d_chain_maxpool_bf16: mid width %d", N);
        AL3D_CHECK_ARG(prev % 64 == 0, "al3d_chain_maxpool_bf16: K=%d", prev);
        // output buffer capacities: layers 0,2 write bufB (32 KB = 128 ch), layer 1 writes bufA (64 KB = 256 ch)
        AL3D_CHECK_ARG((l & 1) ? N <= 256 : N <= 128, "al3d_chain_maxpool_bf16: layer %d width %d exceeds its buffer", l, N);
        prev = N; tot += N;
    }
    AL3D_CHECK_ARG(w->w0 <= 128 && tot <= 512, "al3d_chain_maxpool_bf16: widths too large");
    AL3D_CHECK_ARG(w->last % 256 == 0 && w->last >= 256 && w->last <= 1024, "al3d_chain_maxpool_bf16: last=%d", w->last);
    AL3D_CHECK_ARG(bs >= 0 && n >= 1, "al3d_chain_maxpool_bf16: bad shape");
    if (bs == 0) return 0;
    if (w->n_mid == 2 && w->last == 512) {
        // the three trunks of the models: specialised tile-pair kernels (the generic chain kernel below keeps any other shape)
        TrunkParams tp;
        tp.x = x; tp.sb = sb; tp.sc = sc; tp.sp = sp; tp.bs = bs; tp.n = n; tp.c_in = w->c_in;
        tp.w0_w = w->w0_w; tp.w0_b = w->w0_b; tp.mid_b = w->mid_b; tp.last_b = w->last_b;
        tp.wstream = (const uint8_t *)w->wstream; tp.out = out; tp.splits = 1; tp.n_items = 0;
        if (tc_launch_status(&tp.wd)) return 1;
        const int sms_ = num_sms();
        if (w->w0 == 128 && w->mid[0] == 128 && w->mid[1] == 256) return launch_trunk_pair<128, 128, 256>(tp, bs, n, (cudaStream_t)stream, sms_);
        if (w->w0 == 64 && w->mid[0] == 128 && w->mid[1] == 256) return launch_trunk_pair<64, 128, 256>(tp, bs, n, (cudaStream_t)stream, sms_);
        if (w->w0 == 64 && w->mid[0] == 64 && w->mid[1] == 128) return launch_trunk_pair<64, 64, 128>(tp, bs, n, (cudaStream_t)stream, sms_);
    }
    ChainParams p;
    p.x = x; p.sb = sb; p.sc = sc; p.sp = sp; p.bs = bs; p.n = n;
    p.c_in = w->c_in; p.w0 = w->w0; p.n_mid = w->n_mid;
    for (int l = 0; l < 3; ++l) p.mid[l] = w->mid[l];
    p.last = w->last;
    p.w0_w = w->w0_w; p.w0_b = w->w0_b; p.mid_b = w->mid_b; p.last_b = w->last_b;
    p.wstream = (const uint8_t *)w->wstream; p.out = out;
    if (tc_launch_status(&p.wd)) return 1;
    const int tiles = (n + kTile - 1) / kTile;
    const int sms = num_sms();
    int splits = 1;
    if (bs < 2 * sms) splits = (int)std::min<int64_t>(tiles, ceil_div(2 * sms, bs));
    p.splits = splits;
    p.n_items = bs * splits;
    const int grid = std::min(p.n_items, sms);
    // activation buffers sized for this chain (layer 0 and odd mid layers write A, even mid layers write B);
    // whatever shared memory is left becomes weight-ring stages: the deeper the ring, the better the
    // L2 -> smem latency of the weight stream is hidden.
    int wA = w->w0, wB = 0;
    for (int l = 0; l < w->n_mid; ++l) { if (l & 1) wA = std::max(wA, w->mid[l]); else wB = std::max(wB, w->mid[l]); }
    p.bufA_bytes = wA * kTile * 2;
    p.bufB_bytes = wB * kTile * 2;
    const int budget = 232448 - 1024;            // 227 KB opt-in limit minus static smem / alignment slack
    int stages = (budget - p.bufA_bytes - p.bufB_bytes - (int)sizeof(ChainSmemTail)) / kStageBytes;
    stages = std::min(stages, kChainMaxStages);
    AL3D_CHECK_ARG(stages >= 2, "al3d_chain_maxpool_bf16: no room for the weight ring");
    p.n_stages = stages;
    const size_t smem = (size_t)p.bufA_bytes + p.bufB_bytes + (size_t)stages * kStageBytes + sizeof(ChainSmemTail);
    AL3D_CHECK_CUDA(cudaFuncSetAttribute(chain_max_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    chain_max_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(p);
    AL3D_CHECK_LAUNCH("chain_max_kernel");
    return 0;
}

extern "C" int al3d_seg_pass1_bf16(const al3d_pass1_weights *w, const float *x, int64_t sb, int64_t sc, int64_t sp,
                                   int bs, int n, float *out, void *stream)
{
    AL3D_CHECK_ARG(w && x && out, "al3d_seg_pass1_bf16: null pointer");
    AL3D_CHECK_ARG(w->c_in >= 1 && w->c_in <= 8, "al3d_seg_pass1_bf16: c_in=%d", w->c_in);
    AL3D_CHECK_ARG(bs >= 0 && n >= 1, "al3d_seg_pass1_bf16: bad shape");
    if (bs == 0) return 0;
    Pass1Params p;
    p.x = x; p.sb = sb; p.sc = sc; p.sp = sp; p.bs = bs; p.n = n; p.c_in = w->c_in;
    p.w1_w = w->w1_w; p.w1_b = w->w1_b; p.b2 = w->b2; p.b3 = w->b3; p.b4 = w->b4; p.b5 = w->b5;
    p.wfront = (const uint8_t *)w->wfront; p.w5stream = (const uint8_t *)w->w5stream; p.out = out;
    AL3D_CHECK_ARG(w->consts_host, "al3d_seg_pass1_bf16: consts_host is null");
    {
        const float *h = w->consts_host;
        std::memcpy(p.w1c, h, sizeof(p.w1c)); std::memcpy(p.b1c, h + 512, sizeof(p.b1c)); std::memcpy(p.b2c, h + 576, sizeof(p.b2c));
        std::memcpy(p.b3c, h + 640, sizeof(p.b3c)); std::memcpy(p.b4c, h + 704, sizeof(p.b4c));
    }
    if (tc_launch_status(&p.wd)) return 1;
    p.dbg = g_debug_buffer ? g_debug_buffer + 3 * 4 * 64 : nullptr;      // second half of the debug buffer
    p.dbg_skip = debug_skip();
    const int tiles = (n + kTile - 1) / kTile;
    const int sms = num_sms();
    int splits = 1;
    if (bs < 2 * sms) splits = (int)std::min<int64_t>((tiles + 1) / 2, ceil_div(2 * sms, bs));
    p.splits = std::max(splits, 1);
    p.n_items = bs * p.splits;
    const int grid = std::min(p.n_items, sms);
    const size_t smem = sizeof(Pass1Smem) + 128;
    static_assert(sizeof(Pass1Smem) + 128 <= 232448, "Pass1Smem exceeds the 227 KB opt-in limit");
    AL3D_CHECK_CUDA(cudaFuncSetAttribute(seg_pass1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    seg_pass1_kernel<<<grid, kP1Threads, smem, (cudaStream_t)stream>>>(p);
    AL3D_CHECK_LAUNCH("seg_pass1_kernel");
    return 0;
}

extern "C" int al3d_seg_pass2_bf16(const al3d_pass2_weights *w, const float *x, int64_t sb, int64_t sc, int64_t sp,
                                   int bs, int n, const float *gbias, float *logits, uint8_t *mask, void *stream)
{
    AL3D_CHECK_ARG(w && x && gbias && logits && mask, "al3d_seg_pass2_bf16: null pointer");
    AL3D_CHECK_ARG(w->c_in >= 1 && w->c_in <= 8, "al3d_seg_pass2_bf16: c_in=%d", w->c_in);
    AL3D_CHECK_ARG(bs >= 0 && n >= 1, "al3d_seg_pass2_bf16: bad shape");
    if (bs == 0) return 0;
    Pass2Params p;
    p.x = x; p.sb = sb; p.sc = sc; p.sp = sp; p.bs = bs; p.n = n; p.c_in = w->c_in;
    p.w1_w = w->w1_w; p.w1_b = w->w1_b; p.b2 = w->b2; p.gbias = gbias;
    p.bd2 = w->bd2; p.bd3 = w->bd3; p.bd4 = w->bd4; p.w5 = w->w5; p.b5 = w->b5;
    p.wstream = (const uint8_t *)w->wstream; p.logits = logits; p.mask = mask; p.dbg = g_debug_buffer; p.dbg_skip = debug_skip();
    if (tc_launch_status(&p.wd)) return 1;
    p.tiles_per_obj = (n + kTile - 1) / kTile;
    const int64_t items = (int64_t)bs * p.tiles_per_obj;
    AL3D_CHECK_ARG(items < (1ll << 31), "al3d_seg_pass2_bf16: too many tiles");
    p.n_items = (int)items;
    int grid = std::min(p.n_items, num_sms());
    grid = ((grid + 1) / 2) * 2;                         // whole CTA pairs; a ghost CTA recomputes the last tile without output
    const size_t smem = sizeof(Pass2Smem) + 128;
    AL3D_CHECK_CUDA(cudaFuncSetAttribute(seg_pass2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    seg_pass2_kernel<<<grid, kThreads + 32, smem, (cudaStream_t)stream>>>(p);     // __cluster_dims__(2,1,1)
    AL3D_CHECK_LAUNCH("seg_pass2_kernel");
    return 0;
}
