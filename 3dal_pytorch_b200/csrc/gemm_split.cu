// Split-precision ("bf16x3") tensor-core GEMMs on fp32 operands in global memory: the layer GEMMs of the training step.
//
//   al3d_gemm_split_nt    C[M x N] (+)= A[M x K] . B[N x K]^T (+ bias[N] | per-group row bias)
//                         forward  (A = activations, B = conv weight;      tools/static_model.py:279-295 in .train())
//                         dgrad    (A = dY,          B = weight transposed; autograd of the same Conv1d / Linear)
//   al3d_gemm_split_tn    C[N x K] (+)= A[M x N]^T . B[M x K]       (reduction over the M rows = points)
//                         wgrad    (A = dY, B = layer input)
//
// Arithmetic, PARTS = 2 ("bf16x3", as in chain_split.cu): every fp32 value is carried as hi = bf16(x), lo = bf16(x - hi)
// and every product is three tcgen05.mma (kind::f16, fp32 accumulation in TMEM): ~1e-5 relative to an fp32 GEMM.
// PARTS = 3 ("bf16x6"): x = hi + mid + lo carries all 24 mantissa bits and a product is the six MMAs of combined order
// <= 2 (hh, hm, mh, hl, lh, mm): ~1e-7 relative, i.e. fp32-grade -- the mode the training step uses by default.
//
// NT kernel.  Persistent, one CTA per SM, 320 threads.  An item is (128-row tile of A, pass of <= 256 output columns);
// accumulators ping-pong between TMEM columns [0,256) and [256,512).  Warp 0 streams the packed weight blocks (hi, lo per
// 128 x 64 tile of B; a small kernel packs them from the fp32 weights before the launch) with cp.async.bulk into a
// ring; one thread of warp 1 issues the MMAs; warps 2-9 STAGE the A operand -- each thread loads 32 consecutive fp32
// of its row, splits them and writes hi / lo K-planes of a 64-wide K slab (ring of 4 slabs) -- and run the epilogue of
// the PREVIOUS item (TMEM -> registers -> + bias -> fp32 rows in global memory) while the MMAs of the current one run.
//
// TN kernel.  The reduction dimension is the row index of both global operands, so both are staged MN-major: a thread
// takes 8 consecutive channels of one row (one 16-byte chunk of the canonical no-swizzle MN-major core matrix
// ((8 mn x 2 B) x 8 k x 16 B)) -- no transposition anywhere.  A CTA owns one (128 x <=256) tile of C and a contiguous
// range of rows; partial tiles go to a workspace and are reduced in a fixed order (deterministic).
//
// Every mbarrier guards one buffer and strictly alternates with its consumer (slab ring, weight ring, two accumulators).
#include <algorithm>
#include "common.cuh"
#include "umma.cuh"
#include "split.cuh"
#include "../../include/al3d.h"

namespace al3d {
namespace split {
using namespace umma;

constexpr int kPartBytes = 16384;          // one part (hi / mid / lo) of an A slab: 128 rows x 64 k bf16
template <int PARTS> struct NtCfg;
template <> struct NtCfg<2> { static constexpr int kSlabs = 4, kStages = 5; };
template <> struct NtCfg<3> { static constexpr int kSlabs = 3, kStages = 4; };

// ------------------------------------------------------------------------------------------------ weight packing
// B fp32 (element (n, k) at b[n * ldb + k], or b[k * ldb + n] when `trans`) -> slots in consumption order:
// pass p (np columns) > K slab s > column block nc (rows) > PARTS slots (hi, [mid,] lo); a slot is a KP tile [rows x 64].
template <int PARTS>
__global__ void split_pack_kernel(const float *__restrict__ b, int64_t ldb, int trans, int N, int K, int np, int rows, uint8_t *__restrict__ img)
{
    const int n_nc = np / rows, S = K / 64;
    const int64_t total = (int64_t)N * (K / 8);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i % N), k8 = (int)(i / N);           // consecutive threads: consecutive n (coalesced when trans)
        const int p = n / np, nc = (n % np) / rows, r = n % rows, s = k8 / 8, pl = k8 % 8;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = trans ? b[(int64_t)(k8 * 8 + j) * ldb + n] : b[(int64_t)n * ldb + k8 * 8 + j];
        uint8_t *slot = img + ((((int64_t)p * S + s) * n_nc + nc) * PARTS) * kStage + (size_t)pl * rows * 16 + (size_t)r * 16;
        uint4 h, m, l;
        if (PARTS == 2) {
            split2(v[0], v[1], h.x, l.x); split2(v[2], v[3], h.y, l.y); split2(v[4], v[5], h.z, l.z); split2(v[6], v[7], h.w, l.w);
            *reinterpret_cast<uint4 *>(slot) = h;
            *reinterpret_cast<uint4 *>(slot + kStage) = l;
        } else {
            split3(v[0], v[1], h.x, m.x, l.x); split3(v[2], v[3], h.y, m.y, l.y); split3(v[4], v[5], h.z, m.z, l.z); split3(v[6], v[7], h.w, m.w, l.w);
            *reinterpret_cast<uint4 *>(slot) = h;
            *reinterpret_cast<uint4 *>(slot + kStage) = m;
            *reinterpret_cast<uint4 *>(slot + 2 * kStage) = l;
        }
    }
}

// One K = 64 block of D (+)= A . W^T with operands in PARTS bf16 parts (planes of A part q at a_hi + q * 16 KB):
// weight slot j (part j of W) multiplies the A parts 0 .. PARTS-1-j, so every product of combined order < PARTS is taken.
#define GEMM_MMA_BLOCK(PARTS_, r, d_tmem, a_hi, w_rows, idesc, first, code)                                          \
    {                                                                                                                \
        _Pragma("unroll")                                                                                            \
        for (int j_ = 0; j_ < (PARTS_); ++j_) {                                                                      \
            SPLIT_RING_NEXT(r, code)                                                                                 \
            _Pragma("unroll")                                                                                        \
            for (int k_ = 0; k_ < 4; ++k_) {                                                                         \
                const uint64_t db_ = make_desc(wst_ + k_ * 2 * (w_rows) * 16, (w_rows));                             \
                _Pragma("unroll")                                                                                    \
                for (int q_ = 0; q_ < (PARTS_) - j_; ++q_)                                                           \
                    mma_bf16((d_tmem), make_desc((a_hi) + q_ * kPartBytes + k_ * 2 * kPlane, 128), db_, (idesc),     \
                             ((first) && j_ == 0 && k_ == 0 && q_ == 0) ? 0u : 1u);                                  \
            }                                                                                                        \
            SPLIT_RING_RELEASE(r)                                                                                    \
        }                                                                                                            \
    }

// ------------------------------------------------------------------------------------------------ NT kernel
struct GemmNtParams {
    const float *a; int64_t lda; int M, K;
    const uint8_t *wimg;
    const float *bias, *rowbias; int rows_per_group;
    float *c; int64_t ldc; int N, accumulate;
    int np, rows, n_pass, n_mtiles;
    TcStatus wd;
};

template <int PARTS>
struct GemmNtSmem {
    static constexpr int kSlabs = NtCfg<PARTS>::kSlabs, kStages = NtCfg<PARTS>::kStages, kSlabBytes = PARTS * kPartBytes;
    uint8_t slab[kSlabs][kSlabBytes];
    uint8_t ring[kStages][kStage];
    uint64_t w_full[kStages], w_empty[kStages];
    uint64_t a_full[kSlabs], a_empty[kSlabs];
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};
static_assert(sizeof(GemmNtSmem<2>) + 128 <= 232448 && sizeof(GemmNtSmem<3>) + 128 <= 232448, "GemmNtSmem exceeds the 227 KB opt-in limit");

template <int PARTS>
__global__ void __launch_bounds__(kThreads, 1)
split_gemm_nt_kernel(const GemmNtParams p)
{
    constexpr int kGemmSlabs = GemmNtSmem<PARTS>::kSlabs, kGemmStages = GemmNtSmem<PARTS>::kStages, kSlabBytes = GemmNtSmem<PARTS>::kSlabBytes;
    const TcStatus wd = p.wd;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    GemmNtSmem<PARTS> &s = *reinterpret_cast<GemmNtSmem<PARTS> *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        constexpr int kW = kEpiThreads / 32;
        for (int i = 0; i < kGemmStages; ++i) { mbar_init(&s.w_full[i], 1); mbar_init(&s.w_empty[i], 1); }
        for (int i = 0; i < kGemmSlabs; ++i) { mbar_init(&s.a_full[i], kW); mbar_init(&s.a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&s.acc_full[i], 1); mbar_init(&s.acc_empty[i], kW); }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<512>(&s.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;
    const int S = p.K / 64, n_nc = p.np / p.rows;
    const int n_items = p.n_mtiles * p.n_pass;            // item = mtile * n_pass + pass: a tile's passes are consecutive (A stays in L2)
#define GM_ARRIVE(bar) do { __syncwarp(); if (lane == 0) mbar_arrive(bar); } while (0)

    if (warp == 0) {
        // ------------------------------------------------------------ weight producer (one thread)
        if (elect_one_sync()) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t bytes = (uint32_t)p.rows * 128u;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int pass = item % p.n_pass;
                const uint8_t *src = p.wimg + (size_t)pass * S * n_nc * PARTS * kStage;
                for (int blk = 0; blk < S * n_nc * PARTS; ++blk) {
                    SPLIT_STRESS(wd, 0x71);
                    if (!mbar_wait(&s.w_empty[stage], phase ^ 1, 0x7100 + stage, wd)) goto done;
                    mbar_arrive_expect_tx(&s.w_full[stage], bytes);
                    bulk_g2s(s.ring[stage], src + (size_t)blk * kStage, bytes, &s.w_full[stage]);
                    if (++stage == kGemmStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (one thread)
        if (elect_one_sync()) {
            RingView ring{smem_u32(s.ring[0]), s.w_full, s.w_empty, kGemmStages, 0, 0u};
            const uint32_t idesc = make_idesc_bf16(128, p.rows);
            const uint32_t slab0 = smem_u32(s.slab[0]);
            int ab = 0; uint32_t a_phase = 0, ae_phase[2] = {0, 0};
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int buf = it & 1;
                SPLIT_STRESS(wd, 0x72);
                if (!mbar_wait(&s.acc_empty[buf], ae_phase[buf] ^ 1, 0x7200 + buf, wd)) goto done;     // epilogue of item it - 2 has drained it
                ae_phase[buf] ^= 1;
                tc_fence_after();
                for (int sl = 0; sl < S; ++sl) {
                    SPLIT_STRESS(wd, 0x73);
                    if (!mbar_wait(&s.a_full[ab], a_phase, 0x7300 + ab, wd)) goto done;
                    tc_fence_after();
                    const uint32_t a_hi = slab0 + (uint32_t)ab * kSlabBytes;
                    for (int nc = 0; nc < n_nc; ++nc)
                        GEMM_MMA_BLOCK(PARTS, ring, tmem + buf * 256 + nc * 128, a_hi, p.rows, idesc, sl == 0, 0x7400)
                    mma_commit(&s.a_empty[ab]);
                    if (++ab == kGemmSlabs) { ab = 0; a_phase ^= 1; }
                }
                mma_commit(&s.acc_full[buf]);
            }
        }
    } else {
        // ------------------------------------------------------------ stagers + epilogue (256 threads)
        const int row = epi_row(), half = epi_half();
        const uint32_t lane_addr = (uint32_t)(row & ~31) << 16;
        int ab = 0; uint32_t ae_phase = 0, af_phase[2] = {0, 0};
        // epilogue of one item: this thread owns columns [half * np/2, +np/2) of its row
#define GM_EPILOGUE(item_, buf_)                                                                               \
        {                                                                                                     \
            const int mt_ = (item_) / p.n_pass, pass_ = (item_) % p.n_pass;                                   \
            SPLIT_STRESS_WARP(wd, 0x61);                                                                      \
            if (!mbar_wait(&s.acc_full[buf_], af_phase[buf_], 0x6100 + (buf_), wd)) goto done;                \
            af_phase[buf_] ^= 1;                                                                              \
            tc_fence_after();                                                                                 \
            const int r_ = mt_ * kTile + row;                                                                 \
            const int hc_ = p.np >> 1;                                                                        \
            const int col0_ = pass_ * p.np + half * hc_;                                                      \
            const float *bias_ = p.rowbias ? p.rowbias + (int64_t)((r_ < p.M ? r_ : p.M - 1) / p.rows_per_group) * p.N : p.bias; \
            for (int c_ = 0; c_ < hc_; c_ += 32) {                                                            \
                uint32_t v_[32];                                                                              \
                tmem_ld32(tmem + lane_addr + (buf_) * 256 + half * hc_ + c_, v_);                             \
                tmem_ld_wait();                                                                               \
                if (c_ + 32 >= hc_) { tc_fence_before(); GM_ARRIVE(&s.acc_empty[buf_]); }                     \
                if (r_ < p.M) {                                                                               \
                    float *dst_ = p.c + (int64_t)r_ * p.ldc + col0_ + c_;                                     \
                    _Pragma("unroll")                                                                         \
                    for (int j_ = 0; j_ < 32; j_ += 4) {                                                      \
                        float4 o_ = make_float4(__uint_as_float(v_[j_]), __uint_as_float(v_[j_ + 1]),         \
                                                __uint_as_float(v_[j_ + 2]), __uint_as_float(v_[j_ + 3]));    \
                        if (bias_) {                                                                          \
                            const float *bq_ = bias_ + col0_ + c_ + j_;      /* views of a flat parameter buffer: 4-byte aligned only */ \
                            o_.x += __ldg(bq_); o_.y += __ldg(bq_ + 1); o_.z += __ldg(bq_ + 2); o_.w += __ldg(bq_ + 3); \
                        }                                                                                     \
                        if (p.accumulate) {                                                                   \
                            const float4 q_ = *reinterpret_cast<const float4 *>(dst_ + j_);                   \
                            o_.x += q_.x; o_.y += q_.y; o_.z += q_.z; o_.w += q_.w;                           \
                        }                                                                                     \
                        *reinterpret_cast<float4 *>(dst_ + j_) = o_;                                          \
                    }                                                                                         \
                }                                                                                             \
            }                                                                                                 \
        }
        int it = 0, prev_item = -1;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int mt = item / p.n_pass;
            const int r = mt * kTile + row;
            const float *src = p.a + (int64_t)(r < p.M ? r : p.M - 1) * p.lda + half * 32;
            for (int sl = 0; sl < S; ++sl) {
                float4 x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = __ldg(reinterpret_cast<const float4 *>(src + sl * 64) + j);
                SPLIT_STRESS_WARP(wd, 0x62);
                if (!mbar_wait(&s.a_empty[ab], ae_phase ^ 1, 0x6200 + ab, wd)) goto done;
                uint8_t *dst = s.slab[ab] + (size_t)(half * 4) * kPlane + (size_t)row * 16;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 h, m, l;
                    if (PARTS == 2) {
                        split2(x[2 * j].x, x[2 * j].y, h.x, l.x); split2(x[2 * j].z, x[2 * j].w, h.y, l.y);
                        split2(x[2 * j + 1].x, x[2 * j + 1].y, h.z, l.z); split2(x[2 * j + 1].z, x[2 * j + 1].w, h.w, l.w);
                        *reinterpret_cast<uint4 *>(dst + (size_t)j * kPlane) = h;
                        *reinterpret_cast<uint4 *>(dst + kPartBytes + (size_t)j * kPlane) = l;
                    } else {
                        split3(x[2 * j].x, x[2 * j].y, h.x, m.x, l.x); split3(x[2 * j].z, x[2 * j].w, h.y, m.y, l.y);
                        split3(x[2 * j + 1].x, x[2 * j + 1].y, h.z, m.z, l.z); split3(x[2 * j + 1].z, x[2 * j + 1].w, h.w, m.w, l.w);
                        *reinterpret_cast<uint4 *>(dst + (size_t)j * kPlane) = h;
                        *reinterpret_cast<uint4 *>(dst + kPartBytes + (size_t)j * kPlane) = m;
                        *reinterpret_cast<uint4 *>(dst + 2 * kPartBytes + (size_t)j * kPlane) = l;
                    }
                }
                fence_proxy_async_smem();
                GM_ARRIVE(&s.a_full[ab]);
                if (++ab == kGemmSlabs) { ab = 0; ae_phase ^= 1; }
            }
            if (prev_item >= 0) GM_EPILOGUE(prev_item, (it - 1) & 1)
            prev_item = item;
        }
        if (prev_item >= 0) GM_EPILOGUE(prev_item, (it - 1) & 1)
#undef GM_EPILOGUE
    }
#undef GM_ARRIVE
done:
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------ TN kernel (wgrad)
// C tile [128 rows (n) x kc columns (k)] = sum over this CTA's rows m of A[m, n] * B[m, k].  Both operands MN-major:
// smem slab of 64 m: [m / 8][channel / 8][m % 8][channel % 8] bf16, i.e. 128-byte core matrices (8 channels contiguous,
// 8 m at 16-byte steps), channel groups 128 B apart (SBO), m groups (channels / 8) * 128 B apart (LBO).
constexpr int kTnStages = 2;
constexpr int kTnA = 16384;                                // one part of an A slab (64 m x 128 n bf16)
constexpr int kTnStageBytes = 98304;                       // PARTS A parts | PARTS B parts: 2 x (16 + 32) KB or 3 x (16 + 16) KB
template <int PARTS> struct TnCfg;
template <> struct TnCfg<2> { static constexpr int kMaxKc = 256, kB = 32768; };
template <> struct TnCfg<3> { static constexpr int kMaxKc = 128, kB = 16384; };

struct GemmTnParams {
    const float *a; int64_t lda;          // (M, N)
    const float *b; int64_t ldb;          // (M, K)
    int64_t M; int N, K;
    int kc, n_ntiles, n_ktiles, splits;
    int64_t rows_per_split;               // multiple of 64
    float *part;                          // [splits][N][K]
    TcStatus wd;
};

struct GemmTnSmem {
    uint8_t stage[kTnStages][kTnStageBytes];
    uint64_t full[kTnStages], empty[kTnStages];
    uint64_t acc_full;
    uint32_t tmem_base;
};
static_assert(sizeof(GemmTnSmem) + 128 <= 232448, "GemmTnSmem exceeds the 227 KB opt-in limit");

__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int M, int N) { return make_idesc_bf16(M, N) | (1u << 15) | (1u << 16); }

// 64 rows x (8 * groups) channels of fp32 -> MN-major slab parts.  A warp task covers 8 rows x 4 channel groups: a
// quarter-warp writes the 8 rows of one core matrix (128 contiguous bytes, conflict-free) and reads 32-byte sectors of 8
// consecutive rows.  A thread owns tasks i = 0 .. groups/4 - 1 (warp task w = warp_e + 8 i) and keeps their fp32 values
// in registers: tn_load issues the 32-byte loads of one task, tn_store converts and writes it -- the kernel reloads a
// task for the NEXT slab right after storing it, so the global loads are in flight under the conversion of the other
// tasks, the mbarrier wait and the MMAs (the first cut loaded and converted task by task: ~11 k cycles per slab, all
// of it exposed load latency, against 1.5 k cycles of MMAs).  Measured with the MMAs / the staging switched off in turn
// (256 x 512, M = 262 144): MMAs alone 0.13 ms (parts 2) / 0.22 ms (parts 3), staging alone 0.75 / 0.88 ms -- the kernel is
// bound by re-reading and re-converting the operands (every dY chunk once per K tile, every X chunk once per N tile).
struct TnTask { float4 x0, x1; };

__device__ __forceinline__ void tn_load(TnTask &t, int i, const float *src, int64_t ld, int64_t m0, int64_t m_end, int c0, int c_end,
                                        int warp_e, int lane)
{
    const int w = warp_e + i * (kEpiThreads / 32);
    const int64_t m = m0 + (w & 7) * 8 + (lane & 7);
    const int c = c0 + ((w >> 3) * 4 + (lane >> 3)) * 8;
    t.x0 = make_float4(0.f, 0.f, 0.f, 0.f); t.x1 = t.x0;
    if (m < m_end && c < c_end) {
        const float4 *q = reinterpret_cast<const float4 *>(src + m * ld + c);
        t.x0 = __ldg(q); t.x1 = __ldg(q + 1);
    }
}

template <int PARTS>
__device__ __forceinline__ void tn_store(const TnTask &t, int i, uint8_t *part0, uint32_t part_stride, int groups, int warp_e, int lane)
{
    const int w = warp_e + i * (kEpiThreads / 32);
    const uint32_t off = (uint32_t)(w & 7) * ((uint32_t)groups * 128u) + (uint32_t)((w >> 3) * 4 + (lane >> 3)) * 128u + (uint32_t)(lane & 7) * 16u;
    uint4 h, md, l;
    if (PARTS == 2) {
        split2(t.x0.x, t.x0.y, h.x, l.x); split2(t.x0.z, t.x0.w, h.y, l.y); split2(t.x1.x, t.x1.y, h.z, l.z); split2(t.x1.z, t.x1.w, h.w, l.w);
        *reinterpret_cast<uint4 *>(part0 + off) = h;
        *reinterpret_cast<uint4 *>(part0 + part_stride + off) = l;
    } else {
        split3(t.x0.x, t.x0.y, h.x, md.x, l.x); split3(t.x0.z, t.x0.w, h.y, md.y, l.y);
        split3(t.x1.x, t.x1.y, h.z, md.z, l.z); split3(t.x1.z, t.x1.w, h.w, md.w, l.w);
        *reinterpret_cast<uint4 *>(part0 + off) = h;
        *reinterpret_cast<uint4 *>(part0 + part_stride + off) = md;
        *reinterpret_cast<uint4 *>(part0 + 2 * part_stride + off) = l;
    }
}

template <int PARTS>
__global__ void __launch_bounds__(kThreads, 1)
split_gemm_tn_kernel(const GemmTnParams p)
{
    constexpr uint32_t kTnB = TnCfg<PARTS>::kB;
    const TcStatus wd = p.wd;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    GemmTnSmem &s = *reinterpret_cast<GemmTnSmem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kTnStages; ++i) { mbar_init(&s.full[i], kEpiThreads / 32); mbar_init(&s.empty[i], 1); }
        mbar_init(&s.acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<256>(&s.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;
    const int n_tiles = p.n_ntiles * p.n_ktiles;
    const int tile = blockIdx.x % n_tiles, split = blockIdx.x / n_tiles;
    const int n0 = (tile / p.n_ktiles) * 128, k0 = (tile % p.n_ktiles) * p.kc;
    const int64_t r0 = (int64_t)split * p.rows_per_split;
    const int64_t r1 = r0 + p.rows_per_split < p.M ? r0 + p.rows_per_split : p.M;
    const int n_slabs = (int)((r1 - r0 + 63) / 64);
    const int gb = p.kc / 8;                                  // channel groups of the B slab

    if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (one thread)
        if (elect_one_sync()) {
            const uint32_t idesc = make_idesc_bf16_mn(128, p.kc);
            const uint32_t lbo_a = 16u * 128u, lbo_b = (uint32_t)gb * 128u;
            int st = 0; uint32_t phase = 0;
            for (int sl = 0; sl < n_slabs; ++sl) {
                SPLIT_STRESS(wd, 0x75);
                if (!mbar_wait(&s.full[st], phase, 0x7500 + st, wd)) goto done;
                tc_fence_after();
                const uint32_t a0 = smem_u32(s.stage[st]), b0 = a0 + PARTS * kTnA;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    // B part j multiplies the A parts 0 .. PARTS-1-j: every product of combined order < PARTS
#pragma unroll
                    for (int j = 0; j < PARTS; ++j) {
                        const uint64_t db = make_desc_lbo(b0 + j * kTnB + ks * 2 * lbo_b, lbo_b);
#pragma unroll
                        for (int q = 0; q < PARTS - j; ++q)
                            mma_bf16(tmem, make_desc_lbo(a0 + q * kTnA + ks * 2 * lbo_a, lbo_a), db, idesc, (sl == 0 && ks == 0 && j == 0 && q == 0) ? 0u : 1u);
                    }
                }
                mma_commit(&s.empty[st]);
                if (++st == kTnStages) { st = 0; phase ^= 1; }
            }
            mma_commit(&s.acc_full);
        }
    } else if (warp >= 2) {
        // ------------------------------------------------------------ stagers + epilogue (256 threads)
        const int warp_e = warp - 2;
        constexpr int kNiB = TnCfg<PARTS>::kMaxKc / 32;           // tasks per thread: A 4 (16 channel groups), B kc / 32
        const int ni_b = gb / 4;
        TnTask ta[4], tb[kNiB];
        if (n_slabs > 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) tn_load(ta[i], i, p.a, p.lda, r0, r1, n0, p.N, warp_e, lane);
#pragma unroll
            for (int i = 0; i < kNiB; ++i) if (i < ni_b) tn_load(tb[i], i, p.b, p.ldb, r0, r1, k0, p.K, warp_e, lane);
        }
        int st = 0; uint32_t phase = 0;
        for (int sl = 0; sl < n_slabs; ++sl) {
            SPLIT_STRESS_WARP(wd, 0x65);
            if (!mbar_wait(&s.empty[st], phase ^ 1, 0x6500 + st, wd)) goto done;
            uint8_t *base = s.stage[st];
            const int64_t m_next = r0 + (int64_t)(sl + 1) * 64;
            const bool more = sl + 1 < n_slabs;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                tn_store<PARTS>(ta[i], i, base, kTnA, 16, warp_e, lane);
                if (more) tn_load(ta[i], i, p.a, p.lda, m_next, r1, n0, p.N, warp_e, lane);
            }
#pragma unroll
            for (int i = 0; i < kNiB; ++i) {
                if (i < ni_b) {
                    tn_store<PARTS>(tb[i], i, base + PARTS * kTnA, kTnB, gb, warp_e, lane);
                    if (more) tn_load(tb[i], i, p.b, p.ldb, m_next, r1, k0, p.K, warp_e, lane);
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.full[st]);
            if (++st == kTnStages) { st = 0; phase ^= 1; }
        }
        // epilogue: this thread owns row n0 + row, columns [half * kc/2, +kc/2) of the tile
        const int row = epi_row(), half = epi_half();
        if (n_slabs > 0) {
            if (!mbar_wait(&s.acc_full, 0, 0x6600, wd)) goto done;
            tc_fence_after();
        }
        const int hc = p.kc >> 1;
        float *dst = p.part + ((int64_t)split * p.N + n0 + row) * p.K + k0 + half * hc;
        for (int c = 0; c < hc; c += 32) {
            uint32_t v[32];
            if (n_slabs > 0) {
                tmem_ld32(tmem + ((uint32_t)(row & ~31) << 16) + half * hc + c, v);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
            if (n0 + row < p.N) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4 *>(dst + c + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                           __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            }
        }
        tc_fence_before();
    }
done:
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem);
}

// out[n, k] (+)= sum over splits of part[s][n][k], in a fixed order
__global__ void split_gemm_tn_reduce_kernel(const float *__restrict__ part, int splits, int64_t NK, int K, float *__restrict__ c, int64_t ldc, int accumulate)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < NK; i += (int64_t)gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int sp = 0; sp < splits; ++sp) acc += part[(int64_t)sp * NK + i];
        const int64_t n = i / K, k = i - n * K;
        float *d = c + n * ldc + k;
        *d = accumulate ? *d + acc : acc;
    }
}

}  // namespace split
}  // namespace al3d

using namespace al3d;

extern "C" int64_t al3d_gemm_split_ws_bytes(int N, int K, int parts)
{
    // packed weight image: N/rows x K/64 blocks x `parts` slots of 16 KB
    const int rows = N < 128 ? N : 128;
    return (int64_t)(N / rows) * (K / 64) * parts * al3d::split::kStage;
}

template <int PARTS>
static int launch_nt(const al3d::split::GemmNtParams &p, const float *b, int64_t ldb, int b_transposed, void *ws, cudaStream_t st)
{
    using namespace al3d::split;
    {
        const int64_t total = (int64_t)p.N * (p.K / 8);
        const int blocks = (int)std::min<int64_t>((total + 255) / 256, 4096);
        split_pack_kernel<PARTS><<<blocks, 256, 0, st>>>(b, ldb, b_transposed ? 1 : 0, p.N, p.K, p.np, p.rows, (uint8_t *)ws);
        AL3D_CHECK_LAUNCH("split_pack_kernel");
    }
    const int grid = std::min(p.n_mtiles * p.n_pass, tc_num_sms());
    const size_t smem = sizeof(GemmNtSmem<PARTS>) + 128;
    AL3D_CHECK_CUDA(cudaFuncSetAttribute(split_gemm_nt_kernel<PARTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    split_gemm_nt_kernel<PARTS><<<grid, kThreads, smem, st>>>(p);
    AL3D_CHECK_LAUNCH("split_gemm_nt_kernel");
    return 0;
}

extern "C" int al3d_gemm_split_nt(const float *a, int64_t lda, int M, int K, const float *b, int64_t ldb, int b_transposed,
                                  const float *bias, const float *rowbias, int rows_per_group, int N, int accumulate,
                                  float *c, int64_t ldc, int parts, void *ws, void *stream)
{
    using namespace al3d::split;
    AL3D_CHECK_ARG(a && b && c && ws, "al3d_gemm_split_nt: null pointer");
    AL3D_CHECK_ARG(parts == 2 || parts == 3, "al3d_gemm_split_nt: parts=%d must be 2 (bf16x3) or 3 (bf16x6)", parts);
    AL3D_CHECK_ARG(M >= 1 && K >= 64 && K % 64 == 0, "al3d_gemm_split_nt: M=%d K=%d (K must be a multiple of 64)", M, K);
    AL3D_CHECK_ARG(N == 64 || N == 128 || (N % 256 == 0 && N <= 4096), "al3d_gemm_split_nt: N=%d must be 64, 128 or a multiple of 256", N);
    AL3D_CHECK_ARG(lda % 4 == 0 && ldc % 4 == 0 && ((uintptr_t)a & 15) == 0 && ((uintptr_t)c & 15) == 0, "al3d_gemm_split_nt: A / C rows must be 16-byte aligned");
    AL3D_CHECK_ARG(!rowbias || rows_per_group >= 1, "al3d_gemm_split_nt: rowbias needs rows_per_group");
    AL3D_CHECK_ARG(!(bias && rowbias), "al3d_gemm_split_nt: bias and rowbias are exclusive");
    GemmNtParams p;
    p.a = a; p.lda = lda; p.M = M; p.K = K; p.wimg = (const uint8_t *)ws;
    p.bias = bias; p.rowbias = rowbias; p.rows_per_group = rows_per_group > 0 ? rows_per_group : 1;
    p.c = c; p.ldc = ldc; p.N = N; p.accumulate = accumulate ? 1 : 0;
    p.np = std::min(N, 256); p.rows = std::min(N, 128); p.n_pass = N / p.np; p.n_mtiles = (M + kTile - 1) / kTile;
    if (tc_launch_status(&p.wd)) return 1;
    return parts == 2 ? launch_nt<2>(p, b, ldb, b_transposed, ws, (cudaStream_t)stream) : launch_nt<3>(p, b, ldb, b_transposed, ws, (cudaStream_t)stream);
}

static void tn_plan(int64_t M, int N, int K, int parts, int sms, int &kc, int &n_ntiles, int &n_ktiles, int &splits, int64_t &rows_per_split)
{
    kc = std::min(K, parts == 2 ? 256 : 128);
    n_ntiles = (N + 127) / 128; n_ktiles = K / kc;
    const int tiles = n_ntiles * n_ktiles;
    const int64_t slabs = (M + 63) / 64;
    int64_t want = std::max<int64_t>(1, (2 * sms) / tiles);
    want = std::min<int64_t>(want, slabs);
    rows_per_split = ((slabs + want - 1) / want) * 64;
    splits = (int)((M + rows_per_split - 1) / rows_per_split);
}

extern "C" int64_t al3d_gemm_split_tn_ws_bytes(int64_t M, int N, int K, int parts)
{
    int kc, nn, nk, splits; int64_t rps;
    tn_plan(M, N, K, parts, tc_num_sms(), kc, nn, nk, splits, rps);
    return (int64_t)splits * N * K * 4;
}

extern "C" int al3d_gemm_split_tn(const float *a, int64_t lda, const float *b, int64_t ldb, int64_t M, int N, int K, int parts,
                                  void *ws, float *c, int64_t ldc, int accumulate, void *stream)
{
    using namespace al3d::split;
    AL3D_CHECK_ARG(a && b && c && ws, "al3d_gemm_split_tn: null pointer");
    AL3D_CHECK_ARG(parts == 2 || parts == 3, "al3d_gemm_split_tn: parts=%d must be 2 (bf16x3) or 3 (bf16x6)", parts);
    const int max_kc = parts == 2 ? 256 : 128;
    AL3D_CHECK_ARG(M >= 1 && N >= 8 && N % 8 == 0 && K >= 64 && (K <= max_kc ? K % 32 == 0 : K % max_kc == 0),
                   "al3d_gemm_split_tn: M=%lld N=%d K=%d (N %% 8; K 64..%d in steps of 32 or a multiple of %d)", (long long)M, N, K, max_kc, max_kc);
    AL3D_CHECK_ARG(lda % 4 == 0 && ldb % 4 == 0 && ((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0, "al3d_gemm_split_tn: operand rows must be 16-byte aligned");
    GemmTnParams p;
    p.a = a; p.lda = lda; p.b = b; p.ldb = ldb; p.M = M; p.N = N; p.K = K; p.part = (float *)ws;
    tn_plan(M, N, K, parts, tc_num_sms(), p.kc, p.n_ntiles, p.n_ktiles, p.splits, p.rows_per_split);
    if (tc_launch_status(&p.wd)) return 1;
    const int grid = p.n_ntiles * p.n_ktiles * p.splits;
    const size_t smem = sizeof(GemmTnSmem) + 128;
    if (parts == 2) {
        AL3D_CHECK_CUDA(cudaFuncSetAttribute(split_gemm_tn_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        split_gemm_tn_kernel<2><<<grid, kThreads, smem, (cudaStream_t)stream>>>(p);
    } else {
        AL3D_CHECK_CUDA(cudaFuncSetAttribute(split_gemm_tn_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        split_gemm_tn_kernel<3><<<grid, kThreads, smem, (cudaStream_t)stream>>>(p);
    }
    AL3D_CHECK_LAUNCH("split_gemm_tn_kernel");
    const int64_t NK = (int64_t)N * K;
    split_gemm_tn_reduce_kernel<<<(unsigned)std::min<int64_t>((NK + 255) / 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(
        (const float *)ws, p.splits, NK, K, c, ldc, accumulate ? 1 : 0);
    AL3D_CHECK_LAUNCH("split_gemm_tn_reduce_kernel");
    return 0;
}
