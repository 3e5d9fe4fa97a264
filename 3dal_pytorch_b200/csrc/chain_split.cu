// Split-precision ("bf16x3") tensor-core kernels: the parity-grade mode of the shared point-wise MLPs.
//
// A bf16 operand keeps 8 significant bits; ten chained layers of bf16 x bf16 products leave the segmentation logits
// ~2e-2 (relative to the largest logit) away from the fp32 reference and flip ~1 % of the mask bits (measured,
// profiles/r2_parity_per_tensor.jsonl).  fp16 / tf32 operands (11 bits) still miss the 1e-3 bar (~3e-3).  Here every
// fp32 value x is carried as TWO bf16 numbers, hi = bf16(x) and lo = bf16(x - hi) (16 significant bits together), for
// activations and weights alike, and every product is evaluated as
//        a * w  ~=  a_hi * w_hi + a_lo * w_hi + a_hi * w_lo          (a_lo * w_lo ~ 2^-16 relative, dropped)
// i.e. three tcgen05.mma (kind::f16, fp32 accumulation in TMEM) per K = 16 step instead of one.  The result matches the
// fp32 reference to ~5e-5 of max|ref| on the logits; a mask bit can only differ where |l1 - l0| is within that error.
//
// "mixed" mode (ChainParams::last_f16, TailParams::d2_mode; the default of the Python side): the two widest layers of the
// segmentation net -- conv5 128 -> 1024 and dconv2 512 -> 256, 73 % of its MACs and the two the logits are LEAST
// sensitive to (profiles/r2_precision_study.txt) -- multiply IEEE fp16 operands (11 significant bits) instead: conv5
// as f16(a) * f16(w), ONE MMA per product; dconv2 as (f16 hi + f16 lo of a) * f16(w), two.  Their weight blocks are
// single fp16 slots, the epilogue in front of them writes fp16 planes (saturating at 65504), the instruction descriptor
// names fp16 operands; nothing else changes.  63 % of the MMAs of bf16x3; logits within 4e-5 .. 4.4e-4 of the fp32
// reference (profiles/r2_precision_study_mixed.txt, profiles/r2_parity_per_tensor.jsonl).
//
//   split_chain_kernel   first layer (tiny K, CUDA cores) -> 2-3 chained MMA layers, activations (hi | lo planes) in
//                        shared memory, converted IN PLACE by the epilogue warps -> last layer computed transposed
//                        (channels on TMEM lanes, points on columns) so the max over points is a per-thread reduction.
//                        Single tiles: the static box-head trunk (tools/static_model.py:330-334), PointEmbedding and
//                        BoxEmbedding trunks (tools/dynamic_model.py:241-245, 278-282).
//   split_chain_pair_kernel  the same chain on tile PAIRS for ins_seg conv1-5 + max (tools/static_model.py:279-284):
//                        N = 256 points per streamed conv5 block, the two tiles' fronts interleaved, the next unit's
//                        first layer computed under the streamed last layer.
//   split_tail_kernel    conv1-2 recomputed, dconv1 (+ per-object global-feature bias) in four 128-channel chunks
//                        pipelined into dconv2's accumulation, dconv3, dconv4, and the 128 -> 2 logits + mask in fp32
//                        (tools/static_model.py:286-295, :59); software-pipelined across tiles.
//   split_tail_pair_kernel  cta_group::2 experiment of the tail (off by default).
//
// All are persistent (one CTA per SM, 320 threads): warp 0 streams the packed weight blocks (hi block, lo block per
// 128 x 64 tile of W) with cp.async.bulk into a ring, one thread of warp 1 issues every MMA, warps 2-9 own the 128
// TMEM lanes (two warps per lane quarter, splitting the columns) and run the epilogues.  Every mbarrier that guards
// a buffer is private to that buffer and strictly ping-pongs with its counterpart, so no parity wait can be lapped.
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#include "umma.cuh"
#include "split.cuh"
#include "../../include/al3d.h"

namespace al3d {
namespace split {
using namespace umma;

// Timeline builds (scripts/split_timeline.py; -DAL3D_SPLIT_TIMELINE: split_tail_kernel, -DAL3D_CHAIN_TIMELINE:
// split_chain_kernel in pair mode): CTA 0's MMA issuer and first epilogue thread stamp (id, %clock) pairs into the
// scratch area behind the status word.
#ifdef AL3D_SPLIT_TIMELINE
#define TL_STAMP(id) do { if (blockIdx.x == 0 && tl_n < 1500) { unsigned int c_; asm volatile("mov.u32 %0, %%clock;" : "=r"(c_)); \
                          tl_buf[tl_n * 2] = (unsigned int)(id); tl_buf[tl_n * 2 + 1] = c_; ++tl_n; } } while (0)
#define TL_STAMP_DECL(off) unsigned int *tl_buf = p.wd.word + (off); int tl_n = 0;
#define TL_STAMP_E(id) do { if (etid == 0) TL_STAMP(id); } while (0)
#else
#define TL_STAMP(id) do { } while (0)
#define TL_STAMP_E(id) do { } while (0)
#define TL_STAMP_DECL(off)
#endif
#ifdef AL3D_CHAIN_TIMELINE
#define CH_STAMP(id) do { if (p.pair && blockIdx.x == 0 && tl_n < 1500) { unsigned int c_; asm volatile("mov.u32 %0, %%clock;" : "=r"(c_)); \
                          tl_buf[tl_n * 2] = (unsigned int)(id); tl_buf[tl_n * 2 + 1] = c_; ++tl_n; } } while (0)
#define CH_STAMP_DECL(off) unsigned int *tl_buf = p.wd.word + (off); int tl_n = 0;
#define CH_STAMP_E(id) do { if (etid == 0) CH_STAMP(id); } while (0)
#else
#define CH_STAMP(id) do { } while (0)
#define CH_STAMP_E(id) do { } while (0)
#define CH_STAMP_DECL(off)
#endif

// ================================================================================================
// split_chain_kernel
// ================================================================================================
struct ChainParams {
    const float *x; int64_t sb, sc, sp; int bs, n;       // input (bs, c_in, n), strides in elements
    int c_in, w0, n_mid, mid[3], last;
    const float *w0_w, *w0_b, *mid_b, *last_b;           // fp32: (8, w0) transposed + zero-padded, (w0), concat(mid), (last)
    const uint8_t *wstream;                              // 16 KB slots, (hi, lo) per block, consumption order
    float *out;                                          // (bs, last) fp32, zero-initialised; max-pooled with atomicMax
    int splits, n_items;
    int pair;                                            // 1: tile pairs, split_chain_pair_kernel (last layer with N = 256 points)
    int act_bytes, pair_bytes, n_stages;                 // shared-memory carve-up chosen by the launcher
    int front_blocks, last_blocks;                       // 16 KB slots per tile (mid layers) / per unit (last layer)
    int last_f16;                                        // the streamed last layer as ONE fp16 MMA per product
    int y_front_lo;                                      // pair kernel, last_f16 only: tile Y's front lives in the (then idle) lo half of the pair buffer
    TcStatus wd;
};

// Dynamic shared memory: [act | pair buffer | weight ring | ChainTail]
struct ChainTail {
    float w0_w[128 * 8];
    float w0_b[128];
    float mid_b[512];
    uint64_t w_full[kMaxStages], w_empty[kMaxStages];
    uint64_t act_ready, acc_ready;
    uint64_t last_full[2], last_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ int chain_in_width(const ChainParams &p, int l) { return l == 0 ? p.w0 : p.mid[l - 1]; }

__global__ void __launch_bounds__(kThreads, 1)
split_chain_kernel(const ChainParams p)
{
    const TcStatus wd = p.wd;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *const s_act = smem_raw;
    uint8_t *const s_ring = smem_raw + p.act_bytes;
    ChainTail &s = *reinterpret_cast<ChainTail *>(s_ring + (size_t)p.n_stages * kStage);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t act_lo = (uint32_t)p.act_bytes / 2;

    for (int i = threadIdx.x; i < p.w0 * 8; i += kThreads) s.w0_w[i] = p.w0_w[i];
    for (int i = threadIdx.x; i < p.w0; i += kThreads) s.w0_b[i] = p.w0_b[i];
    {
        int tot = 0;
        for (int l = 0; l < p.n_mid; ++l) tot += p.mid[l];
        for (int i = threadIdx.x; i < tot; i += kThreads) s.mid_b[i] = p.mid_b[i];
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.n_stages; ++i) { mbar_init(&s.w_full[i], 1); mbar_init(&s.w_empty[i], 1); }
        mbar_init(&s.act_ready, kEpiThreads / 32);
        mbar_init(&s.acc_ready, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&s.last_full[i], 1); mbar_init(&s.last_empty[i], kEpiThreads / 32); }
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
#define CH_ARRIVE(bar) do { __syncwarp(); if (lane == 0) mbar_arrive(bar); } while (0)

    if (warp == 0) {
        // ------------------------------------------------------------ weight producer (one thread)
        if (elect_one_sync()) {
            int stage = 0; uint32_t phase = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const int sp_i = item % p.splits;
                const int t0 = (int)((int64_t)tiles_per_obj * sp_i / p.splits), t1 = (int)((int64_t)tiles_per_obj * (sp_i + 1) / p.splits);
                for (int t = t0; t < t1; ++t) {
                    for (int blk = 0; blk < p.front_blocks + p.last_blocks; ++blk) {
                        SPLIT_STRESS(wd, 0x51);
                        if (!mbar_wait(&s.w_empty[stage], phase ^ 1, 0x5100 + stage, wd)) goto done;
                        mbar_arrive_expect_tx(&s.w_full[stage], kStage);
                        bulk_g2s(s_ring + (size_t)stage * kStage, p.wstream + (size_t)blk * kStage, kStage, &s.w_full[stage]);
                        if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (one thread)
        if (elect_one_sync()) {
            RingView ring{smem_u32(s_ring), s.w_full, s.w_empty, p.n_stages, 0, 0u};
            uint32_t act_phase = 0, le_phase[2] = {0, 0};
            const uint32_t a_act = smem_u32(s_act);
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const int sp_i = item % p.splits;
                const int t0 = (int)((int64_t)tiles_per_obj * sp_i / p.splits), t1 = (int)((int64_t)tiles_per_obj * (sp_i + 1) / p.splits);
                for (int t = t0; t < t1; ++t) {
                    // mid layers: D[points x channels], input = the activation buffer (converted in place)
                    for (int l = 0; l < p.n_mid; ++l) {
                        const int K = chain_in_width(p, l), N = p.mid[l];
                        const int rows = N < 128 ? N : 128;
                        SPLIT_STRESS(wd, 0x52);
                        if (!mbar_wait(&s.act_ready, act_phase, 0x5200 + l, wd)) goto done;
                        act_phase ^= 1;
                        tc_fence_after();
                        const uint32_t idesc = make_idesc_bf16(128, rows);
                        for (int nc = 0; nc < N / rows; ++nc)
                            for (int kb = 0; kb < K / 64; ++kb)
                                SPLIT_MMA_BLOCK(ring, tmem + nc * 128, a_act + kb * 8 * kPlane, act_lo, kPlane, 128, rows, idesc, kb == 0, 0x5300)
                        mma_commit(&s.acc_ready);
                    }
                    // last layer, transposed: D^T[channels x points], double-buffered in TMEM
                    {
                        SPLIT_STRESS(wd, 0x52);
                        if (!mbar_wait(&s.act_ready, act_phase, 0x52F0, wd)) goto done;
                        act_phase ^= 1;
                        tc_fence_after();
                        const uint32_t idesc = p.last_f16 ? make_idesc_f16(128, 128) : make_idesc_bf16(128, 128);
                        for (int cc = 0; cc < n_last_chunks; ++cc) {
                            const int b = cc & 1;
                            SPLIT_STRESS(wd, 0x54);
                            if (!mbar_wait(&s.last_empty[b], le_phase[b] ^ 1, 0x5400 + b, wd)) goto done;
                            le_phase[b] ^= 1;
                            tc_fence_after();
                            const uint32_t d = tmem + 256 + b * 128;
                            if (p.last_f16) {
                                for (int kb = 0; kb < k_last / 64; ++kb)
                                    SPLIT_MMA_BLOCK_T_F16(ring, d, a_act + kb * 8 * kPlane, kPlane, 128, idesc, kb == 0, 0x5500)
                            } else {
                                for (int kb = 0; kb < k_last / 64; ++kb)
                                    SPLIT_MMA_BLOCK_T(ring, d, a_act + kb * 8 * kPlane, act_lo, kPlane, 128, idesc, kb == 0, 0x5500)
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
                // ---- first layer (rows past the end of the object replicate its last point: the max-pool is
                //      idempotent under duplicates)
                {
                    int pidx = t * kTile + row;
                    if (pidx > p.n - 1) pidx = p.n - 1;
                    const float *px = p.x + (int64_t)b * p.sb + (int64_t)pidx * p.sp;
                    float xv[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) xv[c] = (c < p.c_in) ? __ldg(px + c * p.sc) : 0.f;
                    first_layer_split(s_act, act_lo, row, xv, p.c_in, p.w0, half * (p.w0 >> 1), p.w0 >> 1, s.w0_w, s.w0_b);
                    fence_proxy_async_smem();
                    CH_ARRIVE(&s.act_ready);
                }
                // ---- mid layers: this thread converts columns [half*N/2, (half+1)*N/2) of its row, in place (the
                //      layer's MMAs are complete when acc_ready fires)
                int boff = 0;
                for (int l = 0; l < p.n_mid; ++l) {
                    const int N = p.mid[l];
                    SPLIT_STRESS_WARP(wd, 0x41);
                    if (!mbar_wait(&s.acc_ready, acc_phase, 0x4100 + l, wd)) goto done;
                    acc_phase ^= 1;
                    tc_fence_after();
                    if (p.last_f16 && l == p.n_mid - 1) epilogue_split<kFmtF16>(tmem + lane_addr, half * (N >> 1), N >> 1, s_act, act_lo, kPlane, row, s.mid_b + boff);
                    else epilogue_split(tmem + lane_addr, half * (N >> 1), N >> 1, s_act, act_lo, kPlane, row, s.mid_b + boff);
                    boff += N;
                    tc_fence_before();
                    fence_proxy_async_smem();
                    CH_ARRIVE(&s.act_ready);
                }
                // ---- last layer: this thread owns channel (cc*128 + row) and half of the unit's points
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    if (cc < n_last_chunks) {
                        const int bsel = cc & 1;
                        SPLIT_STRESS_WARP(wd, 0x42);
                        if (!mbar_wait(&s.last_full[bsel], lf_phase[bsel], 0x4200 + cc, wd)) goto done;
                        lf_phase[bsel] ^= 1;
                        tc_fence_after();
                        float m0 = rmax[cc], m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
                        const uint32_t ta = tmem + lane_addr + 256 + bsel * 128 + half * 64;
                        {
                            uint32_t v0[32], v1[32];
                            tmem_ld32(ta, v0);
                            tmem_ld32(ta + 32, v1);
                            tmem_ld_wait();
                            tc_fence_before(); CH_ARRIVE(&s.last_empty[bsel]);       // all values are in registers
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
#undef CH_ARRIVE
done:
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// ================================================================================================
// split_chain_pair_kernel -- the chain on TILE PAIRS (ins_seg conv1-5 + max, tools/static_model.py:279-284)
//
// A unit is two 128-point tiles X, Y of one object; the last layer runs transposed on both at once (N = 256 points per
// streamed 128-channel block, which halves its weight traffic per point).  Measured with scripts/split_timeline.py
// (profiles/r2_split_chain_timeline_before.txt) the first cut of this path spent 44 % of a unit in the strictly serial
// front of the two tiles, so here
//   * the two tiles' fronts are INTERLEAVED: X lives in the activation buffer, Y in the Y rows of the pair buffer
//     (which is where its last mid layer lands anyway; in-place conversion as everywhere), with accumulators in TMEM
//     columns [0,128) / [128,256) and barriers of their own, so one tile's epilogue runs under the other's MMAs;
//   * the next unit's input points are loaded, and X's first layer is computed into the (idle) activation buffer,
//     UNDER the current unit's last layer, and X's first MMA layer is queued right behind the last streamed block --
//     it only needs TMEM buffer 0, which the second-to-last block has left by then.
// Weight stream image: the mid layers' blocks, then the last layer's; consumption order per unit: every mid layer's
// blocks twice (X, Y), then the last layer's.
// ================================================================================================
struct PairTail {
    float w0_w[128 * 8];
    float w0_b[128];
    float mid_b[512];
    uint64_t w_full[kMaxStages], w_empty[kMaxStages];
    uint64_t act_ready[2], acc_ready[2];           // per tile of the pair: operand written / accumulator complete
    uint64_t pair_ready;                           // both tiles' last mid layer written into the pair buffer
    uint64_t last_full[2], last_empty[2];
    uint32_t tmem_base;
};

struct UnitIter { int item, t, t1; };
__device__ __forceinline__ void unit_span(const ChainParams &p, int tiles_per_obj, int item, int &t0, int &t1)
{
    const int sp_i = item % p.splits;
    t0 = (int)((int64_t)tiles_per_obj * sp_i / p.splits);
    t1 = (int)((int64_t)tiles_per_obj * (sp_i + 1) / p.splits);
}
__device__ __forceinline__ bool unit_first(const ChainParams &p, int tiles_per_obj, UnitIter &u)
{
    u.item = blockIdx.x;
    if (u.item >= p.n_items) return false;
    unit_span(p, tiles_per_obj, u.item, u.t, u.t1);
    return true;
}
__device__ __forceinline__ bool unit_next(const ChainParams &p, int tiles_per_obj, const UnitIter &c, UnitIter &n)
{
    if (c.t + 2 < c.t1) { n = c; n.t += 2; return true; }
    n.item = c.item + gridDim.x;
    if (n.item >= p.n_items) return false;
    unit_span(p, tiles_per_obj, n.item, n.t, n.t1);
    return true;
}

__global__ void __launch_bounds__(kThreads, 1)
split_chain_pair_kernel(const ChainParams p)
{
    const TcStatus wd = p.wd;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *const s_act = smem_raw;
    uint8_t *const s_pair = smem_raw + p.act_bytes;
    uint8_t *const s_ring = s_pair + p.pair_bytes;
    PairTail &s = *reinterpret_cast<PairTail *>(s_ring + (size_t)p.n_stages * kStage);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t act_lo = (uint32_t)p.act_bytes / 2, pair_lo = (uint32_t)p.pair_bytes / 2;

    for (int i = threadIdx.x; i < p.w0 * 8; i += kThreads) s.w0_w[i] = p.w0_w[i];
    for (int i = threadIdx.x; i < p.w0; i += kThreads) s.w0_b[i] = p.w0_b[i];
    {
        int tot = 0;
        for (int l = 0; l < p.n_mid; ++l) tot += p.mid[l];
        for (int i = threadIdx.x; i < tot; i += kThreads) s.mid_b[i] = p.mid_b[i];
    }
    if (threadIdx.x == 0) {
        constexpr int kW = kEpiThreads / 32;
        for (int i = 0; i < p.n_stages; ++i) { mbar_init(&s.w_full[i], 1); mbar_init(&s.w_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s.act_ready[i], kW); mbar_init(&s.acc_ready[i], 1);
            mbar_init(&s.last_full[i], 1); mbar_init(&s.last_empty[i], kW);
        }
        mbar_init(&s.pair_ready, 2 * kW);
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
#define CH_ARRIVE(bar) do { __syncwarp(); if (lane == 0) mbar_arrive(bar); } while (0)

    if (warp == 0) {
        // ------------------------------------------------------------ weight producer (one thread)
        if (elect_one_sync()) {
            int stage = 0; uint32_t phase = 0;
#define CP_PUSH(blk_, bytes_)                                                                               \
            {                                                                                               \
                SPLIT_STRESS(wd, 0x51);                                                                     \
                if (!mbar_wait(&s.w_empty[stage], phase ^ 1, 0x5100 + stage, wd)) goto done;                \
                mbar_arrive_expect_tx(&s.w_full[stage], (bytes_));                                          \
                bulk_g2s(s_ring + (size_t)stage * kStage, p.wstream + (size_t)(blk_) * kStage, (bytes_), &s.w_full[stage]); \
                if (++stage == p.n_stages) { stage = 0; phase ^= 1; }                                       \
            }
            UnitIter u, nx;
            for (bool ok = unit_first(p, tiles_per_obj, u); ok; ok = unit_next(p, tiles_per_obj, u, nx), u = nx) {
                int off = 0;
                for (int l = 0; l < p.n_mid; ++l) {
                    const int nb = 2 * (chain_in_width(p, l) / 64);
                    const uint32_t bytes = (uint32_t)p.mid[l] * 128u;          // rows x 64 bf16
                    for (int q = 0; q < 2; ++q)
                        for (int i = 0; i < nb; ++i) CP_PUSH(off + i, bytes)
                    off += nb;
                }
                for (int i = 0; i < p.last_blocks; ++i) CP_PUSH(p.front_blocks + i, (uint32_t)kStage)
            }
#undef CP_PUSH
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (one thread)
        if (elect_one_sync()) {
            RingView ring{smem_u32(s_ring), s.w_full, s.w_empty, p.n_stages, 0, 0u};
            uint32_t act_phase[2] = {0, 0}, pair_phase = 0, le_phase[2] = {0, 0};
            const uint32_t a_act = smem_u32(s_act), a_pair = smem_u32(s_pair);
            CH_STAMP_DECL(1024)
            UnitIter u, nx;
            for (bool ok = unit_first(p, tiles_per_obj, u); ok; ok = unit_next(p, tiles_per_obj, u, nx), u = nx) {
                CH_STAMP(0x100);
                // TMEM buffer 0 (columns [0,256)) holds the front's accumulators: the previous unit's second-to-last
                // streamed block must have been drained from it (buffer 1 may still be in use)
                SPLIT_STRESS(wd, 0x54);
                if (!mbar_wait(&s.last_empty[0], le_phase[0] ^ 1, 0x5410, wd)) goto done;
                tc_fence_after();
                for (int l = 0; l < p.n_mid; ++l) {
                    const int K = chain_in_width(p, l), N = p.mid[l];
                    const uint32_t idesc = make_idesc_bf16(128, N);
                    for (int q = 0; q < 2; ++q) {
                        SPLIT_STRESS(wd, 0x52);
                        if (!mbar_wait(&s.act_ready[q], act_phase[q], 0x5200 + q * 16 + l, wd)) goto done;
                        act_phase[q] ^= 1;
                        tc_fence_after();
                        CH_STAMP(0x110 + l * 2 + q);
                        // tile Y's front: the Y rows of the pair buffer, or (y_front_lo) a buffer of X's shape in its lo half
                        const bool yrows = q && !p.y_front_lo;
                        const uint32_t a_hi = yrows ? a_pair + kTile * 16 : (q ? a_pair + pair_lo : a_act), a_lo = yrows ? pair_lo : act_lo;
                        const uint32_t a_pl = yrows ? 2 * kPlane : kPlane, a_rows = yrows ? 2 * kTile : kTile;
                        for (int kb = 0; kb < K / 64; ++kb)
                            SPLIT_MMA_BLOCK(ring, tmem + q * 128, a_hi + kb * 8 * a_pl, a_lo, a_pl, a_rows, N, idesc, kb == 0, 0x5300)
                        mma_commit(&s.acc_ready[q]);
                    }
                }
                // last layer, transposed: D^T[channels x 256 points], double-buffered in TMEM
                SPLIT_STRESS(wd, 0x52);
                if (!mbar_wait(&s.pair_ready, pair_phase, 0x52F0, wd)) goto done;
                pair_phase ^= 1;
                tc_fence_after();
                CH_STAMP(0x120);
                const uint32_t idesc = p.last_f16 ? make_idesc_f16(128, 256) : make_idesc_bf16(128, 256);
                for (int cc = 0; cc < n_last_chunks; ++cc) {
                    const int b = cc & 1;
                    SPLIT_STRESS(wd, 0x54);
                    if (!mbar_wait(&s.last_empty[b], le_phase[b] ^ 1, 0x5400 + b, wd)) goto done;
                    le_phase[b] ^= 1;
                    tc_fence_after();
                    if (p.last_f16) {
                        for (int kb = 0; kb < k_last / 64; ++kb)
                            SPLIT_MMA_BLOCK_T_F16(ring, tmem + b * 256, a_pair + kb * 8 * 2 * kPlane, 2 * kPlane, 256, idesc, kb == 0, 0x5500)
                    } else {
                        for (int kb = 0; kb < k_last / 64; ++kb)
                            SPLIT_MMA_BLOCK_T(ring, tmem + b * 256, a_pair + kb * 8 * 2 * kPlane, pair_lo, 2 * kPlane, 256, idesc, kb == 0, 0x5500)
                    }
                    mma_commit(&s.last_full[b]);
                    CH_STAMP(0x130 + cc);
                }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue warps (256 threads)
        const int row = epi_row(), half = epi_half();
        const uint32_t lane_addr = (uint32_t)(row & ~31) << 16;
        const int etid = threadIdx.x - 64;
        (void)etid;
        uint32_t acc_phase[2] = {0, 0}, lf_phase[2] = {0, 0};
        CH_STAMP_DECL(8192)
        float xv[2][8];
        // input points of a unit -> registers (rows past the end of the object replicate its last point, an odd tail
        // pair repeats its tile: the max-pool is idempotent under duplicates)
#define CP_LOAD(u_)                                                                                           \
        {                                                                                                     \
            const int b_ = (u_).item / p.splits;                                                              \
            _Pragma("unroll")                                                                                 \
            for (int q_ = 0; q_ < 2; ++q_) {                                                                  \
                const int tq_ = ((u_).t + q_ < (u_).t1) ? (u_).t + q_ : (u_).t1 - 1;                          \
                int pidx_ = tq_ * kTile + row;                                                                \
                if (pidx_ > p.n - 1) pidx_ = p.n - 1;                                                         \
                const float *px_ = p.x + (int64_t)b_ * p.sb + (int64_t)pidx_ * p.sp;                          \
                _Pragma("unroll")                                                                             \
                for (int c_ = 0; c_ < 8; ++c_) xv[q_][c_] = (c_ < p.c_in) ? __ldg(px_ + c_ * p.sc) : 0.f;     \
            }                                                                                                 \
        }
        // first layer of tile X -> activation buffer
#define CP_FIRST_X()                                                                                          \
        {                                                                                                     \
            first_layer_split(s_act, act_lo, row, xv[0], p.c_in, p.w0, half * (p.w0 >> 1), p.w0 >> 1, s.w0_w, s.w0_b); \
            fence_proxy_async_smem();                                                                         \
            CH_ARRIVE(&s.act_ready[0]);                                                                       \
        }
        // first layer of tile Y -> its front buffer in the lo half of the pair buffer (y_front_lo: nothing reads that half
        // while the fp16 last layer streams, so this runs under it like X's)
#define CP_FIRST_Y_LO()                                                                                       \
        {                                                                                                     \
            first_layer_split(s_pair + pair_lo, act_lo, row, xv[1], p.c_in, p.w0, half * (p.w0 >> 1), p.w0 >> 1, s.w0_w, s.w0_b); \
            fence_proxy_async_smem();                                                                         \
            CH_ARRIVE(&s.act_ready[1]);                                                                       \
        }
        UnitIter u, nx;
        bool ok = unit_first(p, tiles_per_obj, u);
        if (ok) {
            CP_LOAD(u)
            CP_FIRST_X()
            if (p.y_front_lo) CP_FIRST_Y_LO()
        }
        float rmax[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rmax[i] = -INFINITY;
        while (ok) {
            const bool has_next = unit_next(p, tiles_per_obj, u, nx);
            CH_STAMP_E(0x200);
            // ---- first layer of tile Y -> the Y rows of the pair buffer (the previous unit's last layer has read it:
            //      this thread has seen its final accumulator)
            if (!p.y_front_lo) {
                first_layer_split(s_pair, pair_lo, kTile + row, xv[1], p.c_in, p.w0, half * (p.w0 >> 1), p.w0 >> 1, s.w0_w, s.w0_b, 2 * kPlane);
                fence_proxy_async_smem();
                CH_ARRIVE(&s.act_ready[1]);
            }
            CH_STAMP_E(0x201);
            // ---- mid layers, X and Y alternating: this thread converts columns [half*N/2, (half+1)*N/2) of its row,
            //      in place (the layer's MMAs are complete when acc_ready fires); the last mid layer of both tiles
            //      lands in the pair buffer
            int boff = 0;
            for (int l = 0; l < p.n_mid; ++l) {
                const int N = p.mid[l];
                const bool last_mid = l == p.n_mid - 1;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    SPLIT_STRESS_WARP(wd, 0x41);
                    if (!mbar_wait(&s.acc_ready[q], acc_phase[q], 0x4100 + q * 16 + l, wd)) goto done;
                    acc_phase[q] ^= 1;
                    tc_fence_after();
                    CH_STAMP_E(0x210 + l * 4 + q * 2);
                    if (last_mid && p.last_f16) epilogue_split<kFmtF16>(tmem + lane_addr + q * 128, half * (N >> 1), N >> 1, s_pair, pair_lo, 2 * kPlane, q * kTile + row, s.mid_b + boff);
                    else if (q == 1 && p.y_front_lo) epilogue_split(tmem + lane_addr + 128, half * (N >> 1), N >> 1, s_pair + pair_lo, act_lo, kPlane, row, s.mid_b + boff);
                    else if (q == 1 || last_mid) epilogue_split(tmem + lane_addr + q * 128, half * (N >> 1), N >> 1, s_pair, pair_lo, 2 * kPlane, q * kTile + row, s.mid_b + boff);
                    else                    epilogue_split(tmem + lane_addr, half * (N >> 1), N >> 1, s_act, act_lo, kPlane, row, s.mid_b + boff);
                    tc_fence_before();
                    fence_proxy_async_smem();
                    CH_ARRIVE(last_mid ? &s.pair_ready : &s.act_ready[q]);
                    CH_STAMP_E(0x211 + l * 4 + q * 2);
                }
                boff += N;
            }
            // ---- last layer: this thread owns channel (cc*128 + row) and half of the unit's points.  Under it: the
            //      next unit's input loads (after block 0) and tile X's first layer (after block 2; the activation
            //      buffer is idle -- X's last mid layer has been read out of it)
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
                if (cc < n_last_chunks) {
                    const int bsel = cc & 1;
                    SPLIT_STRESS_WARP(wd, 0x42);
                    if (!mbar_wait(&s.last_full[bsel], lf_phase[bsel], 0x4200 + cc, wd)) goto done;
                    lf_phase[bsel] ^= 1;
                    tc_fence_after();
                    CH_STAMP_E(0x230 + cc * 2);
                    float m0 = rmax[cc], m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
                    const uint32_t ta = tmem + lane_addr + bsel * 256 + half * 128;
#pragma unroll
                    for (int c0 = 0; c0 < 128; c0 += 64) {
                        uint32_t v0[32], v1[32];
                        tmem_ld32(ta + c0, v0);
                        tmem_ld32(ta + c0 + 32, v1);
                        tmem_ld_wait();
                        if (c0 + 64 >= 128) { tc_fence_before(); CH_ARRIVE(&s.last_empty[bsel]); }   // all values are in registers
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
                    CH_STAMP_E(0x231 + cc * 2);
                    if (has_next && cc == 0) CP_LOAD(nx)
                    if (has_next && cc == (n_last_chunks > 2 ? 2 : n_last_chunks - 1)) CP_FIRST_X()
                    if (has_next && p.y_front_lo && cc == (n_last_chunks > 4 ? 4 : n_last_chunks - 1)) CP_FIRST_Y_LO()
                }
            }
            // ---- end of the object's share: publish relu(max + bias) >= 0, so integer atomicMax on the bit pattern is exact
            if (!has_next || nx.item != u.item) {
                const int b = u.item / p.splits;
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    if (cc < n_last_chunks) {
                        const int ch = cc * 128 + row;
                        const float v = fmaxf(rmax[cc] + __ldg(p.last_b + ch), 0.f);
                        atomicMax(reinterpret_cast<int *>(p.out + (int64_t)b * p.last + ch), __float_as_int(v));
                        rmax[cc] = -INFINITY;
                    }
                }
            }
            u = nx; ok = has_next;
        }
#undef CP_FIRST_Y_LO
#undef CP_FIRST_X
#undef CP_LOAD
    }
#undef CH_ARRIVE
done:
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}


// ================================================================================================
// split_tail_kernel -- second half of the segmentation net in split precision
//
// Shared memory: A2 (32 KB: conv1 output A1, converted in place into the conv2 output A2, kept for the four dconv1
// chunks) | BIG (128 KB): two chunk buffers CB[0..1] of 128 channels (hi | lo: 64 KB each); later all of BIG holds A3
// (dconv2 output, 256 ch) and then A4 (dconv3 output, 128 ch) | weight ring.
// TMEM: D2 = dconv2 accumulator [0,256); once it is drained, dconv4's accumulator sits in [0,128) and the NEXT tile's
// conv2 accumulator in [128,192) | DA [256,384): dconv1 chunks 0 / 2, dconv3 accumulator | DB [384,512): chunks 1 / 3.
//
// Software pipeline across tiles (measured with scripts/split_timeline.py; profiles/r2_split_tail_timeline.txt): the
// front of tile k+1 -- input loads, conv1 on CUDA cores, conv2, dconv1 chunks 0 and 1 -- runs UNDER tile k's dconv3 /
// dconv4, whose shared-memory and TMEM regions are disjoint from it by then; the dconv2 and dconv3 epilogues hand their
// output over per 64-channel K slab so the next layer's MMAs start after the first slab; and tile k's logits are
// computed from registers after chunk 0 of tile k+1 has been handed to the issuer.
// Weight stream image (16 KB slots, hi then lo per block): conv2 | d1(0) | d1(1) | d1(2) | p(0) | d1(3) | p(1) | p(2) |
// p(3) | dconv3 | dconv4, where d1(c) = dconv1 output channels c*128.., p(c) = dconv2 partial sum over input channels
// c*128.. as (row half, k block) x 4.  Steady-state consumption order per tile: d1(2) .. dconv3 | conv2' | dconv4 |
// d1(0)' | d1(1)' (primes: next tile).
// Every mbarrier is private to one buffer / accumulator and completes exactly once (a1/a2/c2/d2/a3/d3/a4/d4) or twice
// (d1_full/d1_act per chunk buffer, strictly alternating with its consumer) per tile, so no parity wait can be lapped.
// ================================================================================================
struct TailParams {
    const float *x; int64_t sb, sc, sp; int bs, n; int c_in;
    const float *w1_w, *w1_b, *b2;     // conv1 fp32 (8, 64) transposed + padded, (64); conv2 bias (64)
    const float *gbias;                // (bs, 512) per-object dconv1 bias
    const float *bd2, *bd3, *bd4;      // dconv2-4 biases (256), (128), (128)
    const float *w5, *b5;              // dconv5 fp32 (2, 128), (2)
    const uint8_t *wstream;
    float *logits; uint8_t *mask;
    int tiles_per_obj, n_items;
    int d2_mode;                       // dconv2 operands: 0 bf16 hi+lo x hi+lo (3 MMAs), 1 fp16 x fp16 (1), 2 fp16 hi+lo x fp16 (2)
    TcStatus wd;
};
constexpr int kTailStages = 3;
constexpr int kTailBlocks = 2 + 4 * 2 + 4 * 8 + 8 + 4;       // 54 slots per tile
constexpr int kBlkConv2 = 0, kBlkD1a = 2, kBlkSteady = 6, kBlkD4 = 50;   // conv2 (2) | d1(0), d1(1) (4) | d1(2) .. dconv3 (44) | dconv4 (4)
constexpr int kBlkD4F16 = 34;      // d2_mode != 0: the sixteen dconv2 blocks are single fp16 slots, the steady section has 28
struct TailSmem {
    uint8_t a2[32768];
    uint8_t big[131072];
    uint8_t ring[kTailStages][kStage];
    float w1_w[64 * 8], w1_b[64], b2[64], gb[512], bd2[256], bd3[128], bd4[128], w5[256], b5[2];
    float lpart[2 * kTile];
    uint64_t w_full[kTailStages], w_empty[kTailStages];
    uint64_t a1_ready, c2_full, a2_ready;          // front of a tile: conv1 output written / conv2 accumulator / conv2 output written
    uint64_t d1_full[2], d1_act[2], cb_free[2];    // per chunk buffer: accumulator ready / operand written / operand consumed
    uint64_t d2_full, a3_ready[4], d3_full, a4_ready[2], d4_full;
    uint32_t tmem_base;
};
static_assert(sizeof(TailSmem) + 128 <= 232448, "TailSmem exceeds the 227 KB opt-in limit");
constexpr uint32_t kTD2 = 0, kTDA = 256, kTDB = 384, kTD4 = 0, kTC2 = 128;

// dconv1 chunk (64 accumulator columns of this thread's row) -> dconv2's operand in chunk buffer `cb`, in the format
// dconv2 multiplies in.
__device__ __forceinline__ void chunk_epilogue(int d2_mode, uint32_t taddr, int c0, uint8_t *cb, int row, const float *bias)
{
    if (d2_mode == 0)      epilogue_split<kFmtBf16x2>(taddr, c0, 64, cb, 32768u, kPlane, row, bias);
    else if (d2_mode == 1) epilogue_split<kFmtF16>(taddr, c0, 64, cb, 32768u, kPlane, row, bias);
    else                   epilogue_split<kFmtF16x2>(taddr, c0, 64, cb, 32768u, kPlane, row, bias);
}

__global__ void __launch_bounds__(kThreads, 1)
split_tail_kernel(const TailParams p)
{
    const TcStatus wd = p.wd;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    TailSmem &s = *reinterpret_cast<TailSmem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < 64 * 8; i += kThreads) s.w1_w[i] = p.w1_w[i];
    for (int i = threadIdx.x; i < 64; i += kThreads) { s.w1_b[i] = p.w1_b[i]; s.b2[i] = p.b2[i]; }
    for (int i = threadIdx.x; i < 256; i += kThreads) { s.bd2[i] = p.bd2[i]; s.w5[i] = p.w5[i]; }
    for (int i = threadIdx.x; i < 128; i += kThreads) { s.bd3[i] = p.bd3[i]; s.bd4[i] = p.bd4[i]; }
    if (threadIdx.x < 2) s.b5[threadIdx.x] = p.b5[threadIdx.x];
    if (threadIdx.x == 0) {
        constexpr int kW = kEpiThreads / 32;
        for (int i = 0; i < kTailStages; ++i) { mbar_init(&s.w_full[i], 1); mbar_init(&s.w_empty[i], 1); }
        mbar_init(&s.a1_ready, kW); mbar_init(&s.c2_full, 1); mbar_init(&s.a2_ready, kW);
        for (int i = 0; i < 2; ++i) { mbar_init(&s.d1_full[i], 1); mbar_init(&s.d1_act[i], kW); mbar_init(&s.cb_free[i], 1); }
        mbar_init(&s.d2_full, 1); mbar_init(&s.d3_full, 1); mbar_init(&s.d4_full, 1);
        for (int i = 0; i < 4; ++i) mbar_init(&s.a3_ready[i], kW);
        for (int i = 0; i < 2; ++i) mbar_init(&s.a4_ready[i], kW);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<512>(&s.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;
    const int item0 = blockIdx.x, stride = gridDim.x;
#define TL_ARRIVE(bar) do { __syncwarp(); if (lane == 0) mbar_arrive(bar); } while (0)

    if (item0 >= p.n_items) {
        // nothing to do for this CTA
    } else if (warp == 0) {
        // ------------------------------------------------------------ weight producer
        if (elect_one_sync()) {
            int stage = 0; uint32_t phase = 0;
#define TL_PUSH(first, count)                                                                         \
            for (int blk = (first); blk < (first) + (count); ++blk) {                                 \
                SPLIT_STRESS(wd, 0x31);                                                               \
                if (!mbar_wait(&s.w_empty[stage], phase ^ 1, 0x3100 + stage, wd)) goto done;          \
                const uint32_t bytes_ = blk < kBlkD1a ? kStage / 2 : kStage;   /* conv2: 64 rows */   \
                mbar_arrive_expect_tx(&s.w_full[stage], bytes_);                                      \
                bulk_g2s(s.ring[stage], p.wstream + (size_t)blk * kStage, bytes_, &s.w_full[stage]);  \
                if (++stage == kTailStages) { stage = 0; phase ^= 1; }                                \
            }
            const int blk_d4 = p.d2_mode ? kBlkD4F16 : kBlkD4;
            TL_PUSH(kBlkConv2, 2)
            TL_PUSH(kBlkD1a, 4)
            for (int item = item0; item < p.n_items; item += stride) {
                const bool has_next = item + stride < p.n_items;
                TL_PUSH(kBlkSteady, blk_d4 - kBlkSteady)
                if (has_next) { TL_PUSH(kBlkConv2, 2) }
                TL_PUSH(blk_d4, 4)
                if (has_next) { TL_PUSH(kBlkD1a, 4) }
            }
#undef TL_PUSH
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (elect_one_sync()) {
            RingView ring{smem_u32(s.ring[0]), s.w_full, s.w_empty, kTailStages, 0, 0u};
            uint32_t d1a_phase[2] = {0, 0};
            const uint32_t id64 = make_idesc_bf16(128, 64), id128 = make_idesc_bf16(128, 128), id128h = make_idesc_f16(128, 128);
            const uint32_t a2 = smem_u32(s.a2), big = smem_u32(s.big);
            TL_STAMP_DECL(1024)
#define TL_WAIT(bar, par, code)                                                  \
            SPLIT_STRESS(wd, 0x32);                                              \
            if (!mbar_wait(bar, par, code, wd)) goto done;                       \
            tc_fence_after();
            // dconv1 chunk c: A2 x Wd1[c*128.., 0:64] -> DA / DB
#define TL_ISSUE_D1(c)                                                                                        \
            { SPLIT_MMA_BLOCK(ring, tmem + (((c) & 1) ? kTDB : kTDA), a2, 16384u, kPlane, 128, 128, id128, true, 0x3310) \
              mma_commit(&s.d1_full[(c) & 1]); }
            // front of a tile: conv2 on A1 (in A2's place, lo half at +16 KB) -> [kTC2, +64); then chunks 0 and 1
#define TL_ISSUE_CONV2(par)                                                                                   \
            { TL_WAIT(&s.a1_ready, par, 0x3200)                                                               \
              SPLIT_MMA_BLOCK(ring, tmem + kTC2, a2, 16384u, kPlane, 128, 64, id64, true, 0x3300)             \
              mma_commit(&s.c2_full); }
#define TL_ISSUE_D1_FIRST(par)                                                                                \
            { TL_WAIT(&s.a2_ready, par, 0x3201)                                                               \
              TL_ISSUE_D1(0)                                                                                  \
              TL_ISSUE_D1(1) }
            TL_ISSUE_CONV2(0u)
            TL_ISSUE_D1_FIRST(0u)
            uint32_t k = 0;                                            // local tile counter; once-per-tile barriers: parity k & 1
            for (int item = item0; item < p.n_items; item += stride, ++k) {
                const bool has_next = item + stride < p.n_items;
                const uint32_t par = k & 1;
                TL_STAMP(0x100);
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const int j = c & 1;
                    SPLIT_STRESS(wd, 0x34);
                    // CB[j] written, DA / DB drained (c == 0: also the previous tile's dconv4 accumulator in [0,128))
                    if (!mbar_wait(&s.d1_act[j], d1a_phase[j], 0x3400 + c, wd)) goto done;
                    d1a_phase[j] ^= 1; tc_fence_after();
                    TL_STAMP(0x110 + c);
                    if (c + 2 < 4) TL_ISSUE_D1(c + 2)
                    // dconv2 partial sum over input channels c*128..+128: D2[:, nc*128..] += CB[j] x Wd2[nc*128.., c*128..]
                    const uint32_t cb = big + j * 65536;
                    if (p.d2_mode) {
                        const bool two = p.d2_mode == 2;
                        for (int nc = 0; nc < 2; ++nc)
                            for (int kb = 0; kb < 2; ++kb)
                                SPLIT_MMA_BLOCK_F16(ring, tmem + kTD2 + nc * 128, cb + kb * 8 * kPlane, 32768u, kPlane, 128, 128, id128h, c == 0 && kb == 0, two, 0x3320)
                    } else {
                        for (int nc = 0; nc < 2; ++nc)
                            for (int kb = 0; kb < 2; ++kb)
                                SPLIT_MMA_BLOCK(ring, tmem + kTD2 + nc * 128, cb + kb * 8 * kPlane, 32768u, kPlane, 128, 128, id128, c == 0 && kb == 0, 0x3320)
                    }
                    if (c + 2 < 4) mma_commit(&s.cb_free[j]);          // chunk c + 2 may overwrite CB[j] once these have run
                    TL_STAMP(0x120 + c);
                }
                mma_commit(&s.d2_full);
                // dconv3: A3 (256 ch over all of BIG, lo half at +64 KB) -> DA, slab by slab
                for (int kb = 0; kb < 4; ++kb) {
                    TL_WAIT(&s.a3_ready[kb], par, 0x3500 + kb)
                    TL_STAMP(0x130 + kb);
                    SPLIT_MMA_BLOCK(ring, tmem + kTDA, big + kb * 8 * kPlane, 65536u, kPlane, 128, 128, id128, kb == 0, 0x3330)
                }
                mma_commit(&s.d3_full);
                TL_STAMP(0x138);
                if (has_next) TL_ISSUE_CONV2(par ^ 1)                  // D2 is drained: every A3 slab has been written
                TL_STAMP(0x139);
                // dconv4: A4 (128 ch, lo half at +32 KB) -> [0,128)
                for (int kb = 0; kb < 2; ++kb) {
                    TL_WAIT(&s.a4_ready[kb], par, 0x3600 + kb)
                    TL_STAMP(0x140 + kb);
                    SPLIT_MMA_BLOCK(ring, tmem + kTD4, big + kb * 8 * kPlane, 32768u, kPlane, 128, 128, id128, kb == 0, 0x3340)
                }
                mma_commit(&s.d4_full);
                TL_STAMP(0x148);
                if (has_next) TL_ISSUE_D1_FIRST(par ^ 1)               // DA: dconv3's accumulator drained (A4 written); DB: chunk 3 drained
                TL_STAMP(0x149);
            }
#undef TL_ISSUE_D1_FIRST
#undef TL_ISSUE_CONV2
#undef TL_ISSUE_D1
#undef TL_WAIT
        }
    } else {
        // ------------------------------------------------------------ epilogue warps (256 threads)
        const int row = epi_row(), half = epi_half();
        const uint32_t tl = tmem + ((uint32_t)(row & ~31) << 16);
        const int etid = threadIdx.x - 64;
        uint32_t d1f_phase[2] = {0, 0}, cbf_phase[2] = {0, 0};
        TL_STAMP_DECL(8192)
#define TL_WAIT(bar, par, code)                                                  \
        SPLIT_STRESS_WARP(wd, 0x21);                                             \
        if (!mbar_wait(bar, par, code, wd)) goto done;                           \
        tc_fence_after();
#define TL_PUBLISH(bar) do { tc_fence_before(); fence_proxy_async_smem(); TL_ARRIVE(bar); } while (0)
        // input point + per-object bias of a tile -> registers (the loads are issued early, consumed by TL_FRONT)
        float xv[8], gbv[2];
#define TL_LOAD(item_, b_, pidx_, valid_)                                                                     \
        {                                                                                                     \
            b_ = (item_) / p.tiles_per_obj;                                                                   \
            const int pr_ = ((item_) % p.tiles_per_obj) * kTile + row;                                        \
            valid_ = pr_ < p.n; pidx_ = valid_ ? pr_ : p.n - 1;                                               \
            const float *px_ = p.x + (int64_t)(b_) * p.sb + (int64_t)(pidx_) * p.sp;                          \
            _Pragma("unroll")                                                                                 \
            for (int c_ = 0; c_ < 8; ++c_) xv[c_] = (c_ < p.c_in) ? __ldg(px_ + c_ * p.sc) : 0.f;             \
            gbv[0] = __ldg(p.gbias + (int64_t)(b_) * 512 + etid);                                             \
            gbv[1] = __ldg(p.gbias + (int64_t)(b_) * 512 + 256 + etid);                                       \
        }
        // conv1 on CUDA cores -> A1 in A2's place (every dconv1 chunk of the previous tile has read A2); per-object
        // bias -> s.gb (every chunk epilogue of the previous tile has read it)
#define TL_FRONT()                                                                                            \
        {                                                                                                     \
            s.gb[etid] = gbv[0]; s.gb[256 + etid] = gbv[1];                                                   \
            first_layer_split(s.a2, 16384u, row, xv, p.c_in, 64, half * 32, 32, s.w1_w, s.w1_b);              \
            TL_PUBLISH(&s.a1_ready);                                                                          \
            asm volatile("bar.sync 1, 256;" ::: "memory");            /* s.gb complete; s.lpart consumed */   \
        }
        // conv2 epilogue: [kTC2, +64) -> A2, in place of A1
#define TL_CONV2_EPI(par)                                                                                     \
        {                                                                                                     \
            TL_WAIT(&s.c2_full, par, 0x2100)                                                                  \
            epilogue_split(tl + kTC2, half * 32, 32, s.a2, 16384u, kPlane, row, s.b2);                        \
            TL_PUBLISH(&s.a2_ready);                                                                          \
        }
        // dconv1 chunk epilogue: DA / DB (128 columns) + per-object bias -> CB[c & 1]
#define TL_CHUNK_EPI(c)                                                                                       \
        {                                                                                                     \
            const int j_ = (c) & 1;                                                                           \
            SPLIT_STRESS_WARP(wd, 0x22);                                                                      \
            if (!mbar_wait(&s.d1_full[j_], d1f_phase[j_], 0x2200 + (c), wd)) goto done;                       \
            d1f_phase[j_] ^= 1; tc_fence_after();                                                             \
            if ((c) >= 2) {                                                                                   \
                if (!mbar_wait(&s.cb_free[j_], cbf_phase[j_], 0x2210 + (c), wd)) goto done;   /* dconv2 partial c - 2 has read CB[j] */ \
                cbf_phase[j_] ^= 1;                                                                           \
            }                                                                                                 \
            chunk_epilogue(p.d2_mode, tl + (j_ ? kTDB : kTDA), half * 64, s.big + j_ * 65536, row, s.gb + (c) * 128); \
            TL_PUBLISH(&s.d1_act[j_]);                                                                        \
        }
        int b, pidx; bool valid;
        TL_LOAD(item0, b, pidx, valid)
        TL_FRONT()
        TL_CONV2_EPI(0u)
        TL_CHUNK_EPI(0)
        uint32_t k = 0;
        for (int item = item0; item < p.n_items; item += stride, ++k) {
            const bool has_next = item + stride < p.n_items;
            const uint32_t par = k & 1;
            int nb = 0, npidx = 0; bool nvalid = false;
            TL_STAMP_E(0x200);
#pragma unroll 1
            for (int c = 1; c < 4; ++c) TL_CHUNK_EPI(c)
            TL_STAMP_E(0x201);
            if (has_next) TL_LOAD(item + stride, nb, npidx, nvalid)
            // ---- dconv2 epilogue: D2 (256 columns) -> A3 over all of BIG (every partial sum has been consumed), per K slab
            TL_WAIT(&s.d2_full, par, 0x2101)
            TL_STAMP_E(0x202);
#pragma unroll 1
            for (int kb = 0; kb < 4; ++kb) {
                epilogue_split(tl + kTD2, kb * 64 + half * 32, 32, s.big, 65536u, kPlane, row, s.bd2);
                TL_PUBLISH(&s.a3_ready[kb]);
            }
            TL_STAMP_E(0x203);
            if (has_next) TL_FRONT()
            TL_STAMP_E(0x204);
            // ---- dconv3 epilogue: DA (128 columns) -> A4, per K slab
            TL_WAIT(&s.d3_full, par, 0x2102)
            TL_STAMP_E(0x205);
#pragma unroll 1
            for (int kb = 0; kb < 2; ++kb) {
                epilogue_split(tl + kTDA, kb * 64 + half * 32, 32, s.big, 32768u, kPlane, row, s.bd3);
                TL_PUBLISH(&s.a4_ready[kb]);
            }
            TL_STAMP_E(0x206);
            if (has_next) TL_CONV2_EPI(par ^ 1)
            TL_STAMP_E(0x207);
            // ---- dconv4 epilogue: accumulator -> registers; chunk 0 of the next tile goes first (the issuer is waiting
            //      for it), then bias + ReLU in fp32, the 128 -> 2 layer, logits and mask from the registers.  Each half
            //      reduces 64 channels; the upper half hands its partial sums over in smem and the lower half adds them in
            //      a fixed order (deterministic).
            TL_WAIT(&s.d4_full, par, 0x2103)
            TL_STAMP_E(0x208);
            {
                uint32_t v0[32], v1[32];
                const int c0 = half * 64;
                tmem_ld32(tl + kTD4 + c0, v0);
                tmem_ld32(tl + kTD4 + c0 + 32, v1);
                tmem_ld_wait();
                tc_fence_before();
                if (has_next) TL_CHUNK_EPI(0)
                TL_STAMP_E(0x209);
                float l0 = 0.f, l1 = 0.f, m0 = 0.f, m1 = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float ya = fmaxf(__uint_as_float(v0[i]) + s.bd4[c0 + i], 0.f);
                    const float yb = fmaxf(__uint_as_float(v1[i]) + s.bd4[c0 + 32 + i], 0.f);
                    l0 = fmaf(ya, s.w5[c0 + i], l0);       l1 = fmaf(ya, s.w5[128 + c0 + i], l1);
                    m0 = fmaf(yb, s.w5[c0 + 32 + i], m0);  m1 = fmaf(yb, s.w5[128 + c0 + 32 + i], m1);
                }
                l0 += m0; l1 += m1;
                if (half == 1) { s.lpart[row] = l0; s.lpart[kTile + row] = l1; }
                asm volatile("bar.sync 2, 256;" ::: "memory");
                if (half == 0 && valid) {
                    const float f0 = (s.b5[0] + l0) + s.lpart[row];
                    const float f1 = (s.b5[1] + l1) + s.lpart[kTile + row];
                    const int64_t o = (int64_t)b * p.n + pidx;
                    *reinterpret_cast<float2 *>(p.logits + o * 2) = make_float2(f0, f1);
                    p.mask[o] = (f0 < f1) ? 1 : 0;
                }
                // s.lpart is rewritten one tile later, after the bar.sync in that tile's TL_FRONT
            }
            TL_STAMP_E(0x20a);
            b = nb; pidx = npidx; valid = nvalid;
        }
#undef TL_CHUNK_EPI
#undef TL_CONV2_EPI
#undef TL_FRONT
#undef TL_LOAD
#undef TL_PUBLISH
#undef TL_WAIT
    }
#undef TL_ARRIVE
done:
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}


// ================================================================================================
// split_tail_pair_kernel -- the same computation on CTA PAIRS (cluster of 2, tcgen05 cta_group::2)
//
// EXPERIMENT, off by default (engine_split.USE_PAIR_KERNEL): split_tail_kernel streams 864 KB of weight slots per
// 128-point tile (~21 k cycles at the measured ~41 B/cycle per-SM L2 -> smem ingress, against ~20.7 k cycles of tensor
// work).  Here one M = 256 MMA covers the 128-point tiles of both CTAs of a pair and each CTA holds only HALF of every
// weight slot (N/2 rows of B): the bytes per SM halve and the same 48 KB of ring hold six stages.  Measured: 39.6 ms
// against 37.3 ms -- the weight stream was not the limiter (the unhidden epilogues are), and the pair adds a cluster round
// trip to every hand-over.
// Every activation buffer, TMEM column and the order of the weight stream are those of split_tail_kernel.
//   leader (cluster rank 0), warp 1: issues every MMA (cta_group::2) and every commit (multicast to both CTAs)
//   both CTAs, warp 0: stream their half slots;  peer, warp 1: forwards "my half of slot s has landed" to the leader
//   both CTAs, warps 2-9: epilogues of their own tile; hand-overs to the issuer arrive on the LEADER's barriers
// ================================================================================================
constexpr int kPairStages = 6;
constexpr int kHalfStage = kStage / 2;
struct TailPairSmem {
    uint8_t a2[32768];
    uint8_t big[131072];
    uint8_t ring[kPairStages][kHalfStage];
    float w1_w[64 * 8], w1_b[64], b2[64], gb[512], bd2[256], bd3[128], bd4[128], w5[256], b5[2];
    float lpart[2 * kTile];
    uint64_t w_full[kPairStages], w_empty[kPairStages], w_peer[kPairStages];   // w_peer: leader only (peer's half landed)
    uint64_t act, acc;                             // act: leader only, one arrival per epilogue warp of BOTH CTAs
    uint64_t d1_full[2], d1_act[2], cb_free[2];    // d1_act: leader only
    uint32_t tmem_base;
};
static_assert(sizeof(TailPairSmem) + 128 <= 232448, "TailPairSmem exceeds the 227 KB opt-in limit");

// One K = 64 block on the pair: D[256 x rows] (+)= A[256 x 64] * W[rows x 64]^T, rows/2 of W in each CTA.
#define PAIR_RING_NEXT(code)                                                                        \
    if (!mbar_wait(&s.w_full[stage], wphase, (code) + stage, wd)) goto done;                         \
    if (!mbar_wait_cluster(&s.w_peer[stage], wphase, (code) + 8 + stage, wd)) goto done;             \
    tc_fence_after();                                                                                \
    const uint32_t wst_ = ring0 + (uint32_t)stage * kHalfStage;
#define PAIR_RING_RELEASE()                                                                         \
    mma_commit_pair(&s.w_empty[stage], 0x3);                                                         \
    if (++stage == kPairStages) { stage = 0; wphase ^= 1; }
#define PAIR_MMA_BLOCK(d_tmem, a_hi, a_lo_off, w_rows, idesc, first, code)                                           \
    {                                                                                                                \
        { PAIR_RING_NEXT(code)                                                                                       \
          _Pragma("unroll")                                                                                          \
          for (int k_ = 0; k_ < 4; ++k_) {                                                                           \
              const uint64_t db_ = make_desc(wst_ + k_ * 2 * ((w_rows) / 2) * 16, (w_rows) / 2);                     \
              mma_bf16_ss_pair((d_tmem), make_desc((a_hi) + k_ * 2 * kPlane, 128), db_, (idesc), ((first) && k_ == 0) ? 0u : 1u); \
              mma_bf16_ss_pair((d_tmem), make_desc((a_hi) + (a_lo_off) + k_ * 2 * kPlane, 128), db_, (idesc), 1u);   \
          }                                                                                                          \
          PAIR_RING_RELEASE() }                                                                                      \
        { PAIR_RING_NEXT(code)                                                                                       \
          _Pragma("unroll")                                                                                          \
          for (int k_ = 0; k_ < 4; ++k_)                                                                             \
              mma_bf16_ss_pair((d_tmem), make_desc((a_hi) + k_ * 2 * kPlane, 128),                                   \
                               make_desc(wst_ + k_ * 2 * ((w_rows) / 2) * 16, (w_rows) / 2), (idesc), 1u);           \
          PAIR_RING_RELEASE() }                                                                                      \
    }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
split_tail_pair_kernel(const TailParams p)
{
    const TcStatus wd = p.wd;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    TailPairSmem &s = *reinterpret_cast<TailPairSmem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();

    for (int i = threadIdx.x; i < 64 * 8; i += kThreads) s.w1_w[i] = p.w1_w[i];
    for (int i = threadIdx.x; i < 64; i += kThreads) { s.w1_b[i] = p.w1_b[i]; s.b2[i] = p.b2[i]; }
    for (int i = threadIdx.x; i < 256; i += kThreads) { s.bd2[i] = p.bd2[i]; s.w5[i] = p.w5[i]; }
    for (int i = threadIdx.x; i < 128; i += kThreads) { s.bd3[i] = p.bd3[i]; s.bd4[i] = p.bd4[i]; }
    if (threadIdx.x < 2) s.b5[threadIdx.x] = p.b5[threadIdx.x];
    if (threadIdx.x == 0) {
        for (int i = 0; i < kPairStages; ++i) { mbar_init(&s.w_full[i], 1); mbar_init(&s.w_empty[i], 1); mbar_init(&s.w_peer[i], 1); }
        mbar_init(&s.act, 2 * kEpiThreads / 32); mbar_init(&s.acc, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&s.d1_full[i], 1); mbar_init(&s.d1_act[i], 2 * kEpiThreads / 32); mbar_init(&s.cb_free[i], 1); }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc_pair<512>(&s.tmem_base);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                  // both CTAs' barriers exist before anything signals them
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;
    // both CTAs of a pair run the same number of rounds; a CTA whose item is past the end recomputes the last item
    // without writing ("ghost"), so the pair's shared MMA stream never changes shape
    const int n_rounds = (p.n_items + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0) {
        // ------------------------------------------------------------ weight producer: this CTA's half of every slot
        if (elect_one_sync()) {
            int stage = 0; uint32_t phase = 0;
            const uint8_t *src = p.wstream + (size_t)crank * kTailBlocks * kHalfStage;
            for (int r = 0; r < n_rounds; ++r)
                for (int blk = 0; blk < kTailBlocks; ++blk) {
                    SPLIT_STRESS(wd, 0x11);
                    if (!mbar_wait(&s.w_empty[stage], phase ^ 1, 0x1100 + stage, wd)) goto done;
                    mbar_arrive_expect_tx(&s.w_full[stage], kHalfStage);
                    bulk_g2s(s.ring[stage], src + (size_t)blk * kHalfStage, kHalfStage, &s.w_full[stage]);
                    if (++stage == kPairStages) { stage = 0; phase ^= 1; }
                }
        }
    } else if (warp == 1) {
        if (crank != 0) {
            // ------------------------------------------------------------ peer: forward "my half has landed" to the leader
            if (elect_one_sync()) {
                int stage = 0; uint32_t phase = 0;
                for (int r = 0; r < n_rounds; ++r)
                    for (int blk = 0; blk < kTailBlocks; ++blk) {
                        if (!mbar_wait(&s.w_full[stage], phase, 0x1200 + stage, wd)) goto done;
                        // default semantics: the half slot was written by this CTA's TMA (complete before the barrier
                        // flipped) and is read by this CTA's tensor core; nothing generic crosses the CTA boundary.  The
                        // .release.cluster form stalls ~250 cycles per arrive and made this forwarding loop the bottleneck
                        // of the whole kernel (58.7 ms instead of 37.3 ms of the single-CTA kernel).
                        mbar_arrive_remote(&s.w_peer[stage], 0);
                        if (++stage == kPairStages) { stage = 0; phase ^= 1; }
                    }
            }
        } else if (elect_one_sync()) {
            // ------------------------------------------------------------ leader: MMA issuer of the pair
            int stage = 0; uint32_t wphase = 0;
            uint32_t act_phase = 0, d1a_phase[2] = {0, 0};
            const uint32_t id64 = make_idesc_bf16(256, 64), id128 = make_idesc_bf16(256, 128);
            const uint32_t a2 = smem_u32(s.a2), big = smem_u32(s.big), ring0 = smem_u32(s.ring[0]);
#define TP_WAIT_ACT(code)                                                        \
            SPLIT_STRESS(wd, 0x12);                                              \
            if (!mbar_wait_cluster(&s.act, act_phase, code, wd)) goto done;      \
            act_phase ^= 1; tc_fence_after();
#define TP_ISSUE_D1(c)                                                                                        \
            { PAIR_MMA_BLOCK(tmem + (((c) & 1) ? kTDB : kTDA), a2, 16384u, 128, id128, true, 0x1310)         \
              mma_commit_pair(&s.d1_full[(c) & 1], 0x3); }
            for (int r = 0; r < n_rounds; ++r) {
                TP_WAIT_ACT(0x1300)
                PAIR_MMA_BLOCK(tmem + kTDA, big, 16384u, 64, id64, true, 0x1320)              // conv2
                mma_commit_pair(&s.acc, 0x3);
                TP_WAIT_ACT(0x1301)
                TP_ISSUE_D1(0)
                TP_ISSUE_D1(1)
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const int j = c & 1;
                    SPLIT_STRESS(wd, 0x14);
                    if (!mbar_wait_cluster(&s.d1_act[j], d1a_phase[j], 0x1400 + c, wd)) goto done;
                    d1a_phase[j] ^= 1; tc_fence_after();
                    if (c + 2 < 4) TP_ISSUE_D1(c + 2)
                    const uint32_t cb = big + j * 65536;
                    for (int nc = 0; nc < 2; ++nc)
                        for (int kb = 0; kb < 2; ++kb)
                            PAIR_MMA_BLOCK(tmem + kTD2 + nc * 128, cb + kb * 8 * kPlane, 32768u, 128, id128, c == 0 && kb == 0, 0x1330)
                    if (c + 2 < 4) mma_commit_pair(&s.cb_free[j], 0x3);
                }
                mma_commit_pair(&s.acc, 0x3);
                TP_WAIT_ACT(0x1302)
                for (int kb = 0; kb < 4; ++kb)
                    PAIR_MMA_BLOCK(tmem + kTDA, big + kb * 8 * kPlane, 65536u, 128, id128, kb == 0, 0x1340)      // dconv3
                mma_commit_pair(&s.acc, 0x3);
                TP_WAIT_ACT(0x1303)
                for (int kb = 0; kb < 2; ++kb)
                    PAIR_MMA_BLOCK(tmem + kTDB, big + kb * 8 * kPlane, 32768u, 128, id128, kb == 0, 0x1350)      // dconv4
                mma_commit_pair(&s.acc, 0x3);
            }
#undef TP_ISSUE_D1
#undef TP_WAIT_ACT
        }
    } else {
        // ------------------------------------------------------------ epilogue warps (256 threads), own tile
        const int row = epi_row(), half = epi_half();
        const uint32_t tl = tmem + ((uint32_t)(row & ~31) << 16);
        const int etid = threadIdx.x - 64;
        uint32_t acc_phase = 0, d1f_phase[2] = {0, 0}, cbf_phase[2] = {0, 0};
#define TP_WAIT_ACC(code)                                                        \
        SPLIT_STRESS_WARP(wd, 0x01);                                             \
        if (!mbar_wait(&s.acc, acc_phase, code, wd)) goto done;                  \
        acc_phase ^= 1; tc_fence_after();
        // the operand was written to this CTA's shared memory through the generic proxy and is consumed by MMAs the OTHER
        // CTA's thread issues: proxy fence, then a cluster-scope release on the leader's barrier
#define TP_PUBLISH(bar) do { tc_fence_before(); fence_proxy_async_smem(); __syncwarp();                                  \
                             if (lane == 0) { if (crank == 0) mbar_arrive(bar); else mbar_arrive_remote_release(bar, 0); } } while (0)
        for (int r = 0; r < n_rounds; ++r) {
            const int item_raw = (int)blockIdx.x + r * (int)gridDim.x;
            const bool ghost = item_raw >= p.n_items;
            const int item = ghost ? p.n_items - 1 : item_raw;
            const int b = item / p.tiles_per_obj, t = item % p.tiles_per_obj;
            const int pidx_raw = t * kTile + row;
            const bool valid = !ghost && pidx_raw < p.n;
            const int pidx = pidx_raw < p.n ? pidx_raw : p.n - 1;
            for (int i = etid; i < 512; i += kEpiThreads) s.gb[i] = __ldg(p.gbias + (int64_t)b * 512 + i);
            {
                const float *px = p.x + (int64_t)b * p.sb + (int64_t)pidx * p.sp;
                float xv[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) xv[c] = (c < p.c_in) ? __ldg(px + c * p.sc) : 0.f;
                first_layer_split(s.big, 16384u, row, xv, p.c_in, 64, half * 32, 32, s.w1_w, s.w1_b);
                TP_PUBLISH(&s.act);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");            // s.gb complete
            TP_WAIT_ACC(0x0100)
            epilogue_split(tl + kTDA, half * 32, 32, s.a2, 16384u, kPlane, row, s.b2);
            TP_PUBLISH(&s.act);
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int j = c & 1;
                SPLIT_STRESS_WARP(wd, 0x02);
                if (!mbar_wait(&s.d1_full[j], d1f_phase[j], 0x0200 + c, wd)) goto done;
                d1f_phase[j] ^= 1; tc_fence_after();
                if (c >= 2) {
                    if (!mbar_wait(&s.cb_free[j], cbf_phase[j], 0x0210 + c, wd)) goto done;
                    cbf_phase[j] ^= 1;
                }
                epilogue_split(tl + (j ? kTDB : kTDA), half * 64, 64, s.big + j * 65536, 32768u, kPlane, row, s.gb + c * 128);
                TP_PUBLISH(&s.d1_act[j]);
            }
            TP_WAIT_ACC(0x0101)
            epilogue_split(tl + kTD2, half * 128, 128, s.big, 65536u, kPlane, row, s.bd2);
            TP_PUBLISH(&s.act);
            TP_WAIT_ACC(0x0102)
            epilogue_split(tl + kTDA, half * 64, 64, s.big, 32768u, kPlane, row, s.bd3);
            TP_PUBLISH(&s.act);
            TP_WAIT_ACC(0x0103)
            {
                uint32_t v0[32], v1[32];
                const int c0 = half * 64;
                tmem_ld32(tl + kTDB + c0, v0);
                tmem_ld32(tl + kTDB + c0 + 32, v1);
                tmem_ld_wait();
                tc_fence_before();
                float l0 = 0.f, l1 = 0.f, m0 = 0.f, m1 = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float ya = fmaxf(__uint_as_float(v0[i]) + s.bd4[c0 + i], 0.f);
                    const float yb = fmaxf(__uint_as_float(v1[i]) + s.bd4[c0 + 32 + i], 0.f);
                    l0 = fmaf(ya, s.w5[c0 + i], l0);       l1 = fmaf(ya, s.w5[128 + c0 + i], l1);
                    m0 = fmaf(yb, s.w5[c0 + 32 + i], m0);  m1 = fmaf(yb, s.w5[128 + c0 + 32 + i], m1);
                }
                l0 += m0; l1 += m1;
                if (half == 1) { s.lpart[row] = l0; s.lpart[kTile + row] = l1; }
                asm volatile("bar.sync 2, 256;" ::: "memory");
                if (half == 0 && valid) {
                    const float f0 = (s.b5[0] + l0) + s.lpart[row];
                    const float f1 = (s.b5[1] + l1) + s.lpart[kTile + row];
                    const int64_t o = (int64_t)b * p.n + pidx;
                    *reinterpret_cast<float2 *>(p.logits + o * 2) = make_float2(f0, f1);
                    p.mask[o] = (f0 < f1) ? 1 : 0;
                }
                asm volatile("bar.sync 2, 256;" ::: "memory");
            }
        }
#undef TP_PUBLISH
#undef TP_WAIT_ACC
    }
done:
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                  // the pair's MMAs touch both CTAs: leave together
    if (warp == 0) tmem_dealloc_pair<512>(tmem);
}

}  // namespace split
}  // namespace al3d

using namespace al3d;

extern "C" int al3d_chain_maxpool_bf16x3(const al3d_split_chain_weights *w, const float *x, int64_t sb, int64_t sc, int64_t sp,
                                         int bs, int n, float *out, void *stream)
{
    using namespace al3d::split;
    AL3D_CHECK_ARG(w && x && out, "al3d_chain_maxpool_bf16x3: null pointer");
    AL3D_CHECK_ARG(w->c_in >= 1 && w->c_in <= 8, "al3d_chain_maxpool_bf16x3: c_in=%d", w->c_in);
    AL3D_CHECK_ARG(w->w0 == 64 || w->w0 == 128, "al3d_chain_maxpool_bf16x3: w0=%d must be 64 or 128", w->w0);
    AL3D_CHECK_ARG(w->n_mid == 2 || w->n_mid == 3, "al3d_chain_maxpool_bf16x3: n_mid=%d", w->n_mid);
    int tot = 0, front_blocks = 0, act_w = w->w0, prev = w->w0;
    for (int l = 0; l < w->n_mid; ++l) {
        const int N = w->mid[l];
        AL3D_CHECK_ARG(N == 64 || N == 128 || N == 256, "al3d_chain_maxpool_bf16x3: mid width %d", N);
        const int rows = std::min(N, 128);
        front_blocks += 2 * (N / rows) * (prev / 64);
        if (!(w->pair && l == w->n_mid - 1)) act_w = std::max(act_w, N);
        prev = N; tot += N;
    }
    AL3D_CHECK_ARG(tot <= 512, "al3d_chain_maxpool_bf16x3: widths too large");
    AL3D_CHECK_ARG(w->last % 128 == 0 && w->last >= 128 && w->last <= 1024, "al3d_chain_maxpool_bf16x3: last=%d", w->last);
    AL3D_CHECK_ARG(bs >= 0 && n >= 1, "al3d_chain_maxpool_bf16x3: bad shape");
    if (bs == 0) return 0;
    ChainParams p;
    p.x = x; p.sb = sb; p.sc = sc; p.sp = sp; p.bs = bs; p.n = n;
    p.c_in = w->c_in; p.w0 = w->w0; p.n_mid = w->n_mid;
    for (int l = 0; l < 3; ++l) p.mid[l] = w->mid[l];
    p.last = w->last;
    p.w0_w = w->w0_w; p.w0_b = w->w0_b; p.mid_b = w->mid_b; p.last_b = w->last_b;
    p.wstream = (const uint8_t *)w->wstream; p.out = out;
    p.pair = w->pair ? 1 : 0;
    p.last_f16 = w->last_f16 ? 1 : 0;
    p.y_front_lo = 0;
    p.front_blocks = front_blocks;
    p.last_blocks = (p.last_f16 ? 1 : 2) * (w->last / 128) * (prev / 64);
    AL3D_CHECK_ARG(w->n_blocks == p.front_blocks + p.last_blocks, "al3d_chain_maxpool_bf16x3: n_blocks=%d, expected %d", w->n_blocks,
                   p.front_blocks + p.last_blocks);
    if (tc_launch_status(&p.wd)) return 1;
    const int tiles = (n + kTile - 1) / kTile;
    const int units = p.pair ? (tiles + 1) / 2 : tiles;
    const int sms = tc_num_sms();
    int splits = 1;
    if (bs < 2 * sms) splits = (int)std::min<int64_t>(units, ceil_div(2 * sms, bs));
    p.splits = std::max(splits, 1);
    p.n_items = bs * p.splits;
    const int grid = std::min(p.n_items, sms);
    p.act_bytes = act_w * kTile * 2 * 2;                       // hi | lo
    p.pair_bytes = p.pair ? prev * 2 * kTile * 2 * 2 : 0;      // 256 rows, hi | lo
    const int budget = 232448 - 1024;
    int stages = (budget - p.act_bytes - p.pair_bytes - (int)sizeof(ChainTail)) / kStage;
    stages = std::min(stages, kMaxStages);
    AL3D_CHECK_ARG(stages >= 2, "al3d_chain_maxpool_bf16x3: no room for the weight ring (act %d B, pair %d B)", p.act_bytes, p.pair_bytes);
    p.n_stages = stages;
    const size_t smem = (size_t)p.act_bytes + p.pair_bytes + (size_t)stages * kStage + sizeof(ChainTail);
    if (p.pair) {
        // tile Y's front lives in the pair buffer, both fronts' accumulators in TMEM buffer 0
        static_assert(sizeof(PairTail) <= sizeof(ChainTail) + 64, "PairTail must fit the ChainTail budget");
        AL3D_CHECK_ARG(act_w <= prev && prev <= 128, "al3d_chain_maxpool_bf16x3: pair mode needs mid widths <= 128 (got %d, %d)", act_w, prev);
        for (int l = 0; l < w->n_mid; ++l) AL3D_CHECK_ARG(w->mid[l] <= 128, "al3d_chain_maxpool_bf16x3: pair mode, mid width %d", w->mid[l]);
        // fp16 last layer: the lo half of the pair buffer is idle while it streams; tile Y's front (a buffer of X's shape)
        // moves there so that its first layer is computed under the previous unit's last layer (12.2 -> 11.8 ms per
        // 8192 x 4096 launch).  AL3D_PAIR_Y_FRONT_LO=0 keeps it in the Y rows.
        {
            static const int ylo_env = [] { const char *e = getenv("AL3D_PAIR_Y_FRONT_LO"); return e ? atoi(e) : 1; }();
            p.y_front_lo = (p.last_f16 && ylo_env && p.act_bytes <= p.pair_bytes / 2) ? 1 : 0;
        }
        const size_t smem_pair = smem - sizeof(ChainTail) + sizeof(PairTail);
        AL3D_CHECK_CUDA(cudaFuncSetAttribute(split_chain_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pair));
        split_chain_pair_kernel<<<grid, kThreads, smem_pair, (cudaStream_t)stream>>>(p);
        AL3D_CHECK_LAUNCH("split_chain_pair_kernel");
        return 0;
    }
    AL3D_CHECK_CUDA(cudaFuncSetAttribute(split_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    split_chain_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(p);
    AL3D_CHECK_LAUNCH("split_chain_kernel");
    return 0;
}

extern "C" int al3d_seg_pass2_bf16x3(const al3d_split_tail_weights *w, const float *x, int64_t sb, int64_t sc, int64_t sp,
                                     int bs, int n, const float *gbias, float *logits, uint8_t *mask, void *stream)
{
    using namespace al3d::split;
    AL3D_CHECK_ARG(w && x && gbias && logits && mask, "al3d_seg_pass2_bf16x3: null pointer");
    AL3D_CHECK_ARG(w->c_in >= 1 && w->c_in <= 8, "al3d_seg_pass2_bf16x3: c_in=%d", w->c_in);
    AL3D_CHECK_ARG(bs >= 0 && n >= 1, "al3d_seg_pass2_bf16x3: bad shape");
    if (bs == 0) return 0;
    TailParams p;
    p.x = x; p.sb = sb; p.sc = sc; p.sp = sp; p.bs = bs; p.n = n; p.c_in = w->c_in;
    p.w1_w = w->w1_w; p.w1_b = w->w1_b; p.b2 = w->b2; p.gbias = gbias;
    p.bd2 = w->bd2; p.bd3 = w->bd3; p.bd4 = w->bd4; p.w5 = w->w5; p.b5 = w->b5;
    p.wstream = (const uint8_t *)w->wstream; p.logits = logits; p.mask = mask;
    AL3D_CHECK_ARG(w->d2_mode >= 0 && w->d2_mode <= 2, "al3d_seg_pass2_bf16x3: d2_mode=%d", w->d2_mode);
    AL3D_CHECK_ARG(w->d2_mode == 0 || w->wstream_pair == nullptr, "al3d_seg_pass2_bf16x3: the CTA-pair kernel has no fp16 dconv2");
    p.d2_mode = w->d2_mode;
    if (tc_launch_status(&p.wd)) return 1;
    p.tiles_per_obj = (n + kTile - 1) / kTile;
    const int64_t items = (int64_t)bs * p.tiles_per_obj;
    AL3D_CHECK_ARG(items < (1ll << 31), "al3d_seg_pass2_bf16x3: too many tiles");
    p.n_items = (int)items;
    if (w->wstream_pair != nullptr && p.n_items >= 2) {
        // CTA pairs: each CTA streams its half of every weight slot (two per-CTA images in wstream_pair)
        p.wstream = (const uint8_t *)w->wstream_pair;
        int grid = std::min(p.n_items, tc_num_sms());
        grid = ((grid + 1) / 2) * 2;                     // whole pairs; a ghost CTA recomputes the last tile without output
        if (grid > tc_num_sms()) grid -= 2;
        const size_t smem = sizeof(TailPairSmem) + 128;
        AL3D_CHECK_CUDA(cudaFuncSetAttribute(split_tail_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        split_tail_pair_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(p);     // __cluster_dims__(2,1,1)
        AL3D_CHECK_LAUNCH("split_tail_pair_kernel");
        return 0;
    }
    const int grid = std::min(p.n_items, tc_num_sms());
    const size_t smem = sizeof(TailSmem) + 128;
    AL3D_CHECK_CUDA(cudaFuncSetAttribute(split_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    split_tail_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(p);
    AL3D_CHECK_LAUNCH("split_tail_kernel");
    return 0;
}
