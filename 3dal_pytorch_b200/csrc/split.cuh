// Shared pieces of the split-precision ("bf16x3") tensor-core kernels (chain_split.cu, gemm_split.cu): the hi/lo split,
// the epilogue that converts an accumulator row into split operand planes, the weight ring as the MMA issuer sees it,
// and the three-MMA block.  See chain_split.cu for the arithmetic.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "umma.cuh"

namespace al3d {
namespace split {
using namespace umma;

constexpr int kTile = 128;                 // points per tile == TMEM lanes
constexpr int kStage = 16384;              // one weight block: <= 128 rows x 64 K bf16
constexpr int kPlane = kTile * 16;         // one K-plane of a 128-row operand
constexpr int kEpiThreads = 256;
constexpr int kThreads = 64 + kEpiThreads;
constexpr int kMaxStages = 8;

#ifdef AL3D_TC_STRESS_HOOKS
#define SPLIT_STRESS(st, salt) stress_delay(st, salt)
#define SPLIT_STRESS_WARP(st, salt) stress_delay_warp(st, salt)
#else
#define SPLIT_STRESS(st, salt) do { } while (0)
#define SPLIT_STRESS_WARP(st, salt) do { } while (0)
#endif

__device__ __forceinline__ int epi_row() { return ((threadIdx.x >> 5) & 3) * 32 + (threadIdx.x & 31); }
__device__ __forceinline__ int epi_half() { return ((threadIdx.x >> 5) - 2) >> 2; }

__device__ __forceinline__ float fmax3(float a, float b, float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// (a, b) fp32 -> hi = {bf16(a), bf16(b)}, lo = {bf16(a - hi_a), bf16(b - hi_b)}; a sits in the low half (even k).
__device__ __forceinline__ void split2(float a, float b, uint32_t &hi, uint32_t &lo)
{
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - hb), "f"(a - ha));
}

// Three-way split: x = hi + mid + lo, 24 significant bits (all of an fp32 mantissa).
__device__ __forceinline__ void split3(float a, float b, uint32_t &hi, uint32_t &mid, uint32_t &lo)
{
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(mid) : "f"(rb), "f"(ra));
    const float sa = ra - __uint_as_float(mid << 16), sb = rb - __uint_as_float(mid & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(sb), "f"(sa));
}

// IEEE fp16 operands for the layers of the "mixed" mode (see chain_split.cu): 11 significant bits per number instead
// of 8.  (a, b) fp32 >= 0 -> {f16(a), f16(b)}, a in the low half; values above the fp16 range saturate to 65504.
__device__ __forceinline__ uint32_t pack_f16(float a, float b)
{
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(fminf(b, 65504.f)), "f"(fminf(a, 65504.f)));
    return r;
}
__device__ __forceinline__ float f16_lo_f32(uint32_t h) { return __half2float(__ushort_as_half((unsigned short)(h & 0xFFFFu))); }
__device__ __forceinline__ float f16_hi_f32(uint32_t h) { return __half2float(__ushort_as_half((unsigned short)(h >> 16))); }
// hi = f16(x), lo = f16(x - hi): 22 significant bits together (less below 2^-14, where lo is subnormal: the absolute
// error stays under 3e-8).
__device__ __forceinline__ void split2_f16(float a, float b, uint32_t &hi, uint32_t &lo)
{
    hi = pack_f16(a, b);
    const float ra = fminf(a, 65504.f) - f16_lo_f32(hi), rb = fminf(b, 65504.f) - f16_hi_f32(hi);
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}

// Operand formats an epilogue can write.
constexpr int kFmtBf16x2 = 0;      // hi | lo planes of bf16 (the bf16x3 layers)
constexpr int kFmtF16 = 1;         // one plane set of fp16 (hi half of the buffer; the lo half is not touched)
constexpr int kFmtF16x2 = 2;       // hi | lo planes of fp16

// 32 accumulator columns (channels c .. c+31 of this thread's row) -> + bias, ReLU, hi / lo split -> four 16-byte
// plane rows in each half of the operand buffer.  dst: address of (plane c/8, row) in the hi half.
template <int kFmt = kFmtBf16x2>
__device__ __forceinline__ void store_split32(uint8_t *dst, uint32_t lo_off, uint32_t plane_stride, const uint32_t (&v)[32],
                                              const float *bias)
{
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 b0 = *reinterpret_cast<const float4 *>(bias + j * 8);
        const float4 b1 = *reinterpret_cast<const float4 *>(bias + j * 8 + 4);
        const float y0 = fmaxf(__uint_as_float(v[j * 8 + 0]) + b0.x, 0.f), y1 = fmaxf(__uint_as_float(v[j * 8 + 1]) + b0.y, 0.f);
        const float y2 = fmaxf(__uint_as_float(v[j * 8 + 2]) + b0.z, 0.f), y3 = fmaxf(__uint_as_float(v[j * 8 + 3]) + b0.w, 0.f);
        const float y4 = fmaxf(__uint_as_float(v[j * 8 + 4]) + b1.x, 0.f), y5 = fmaxf(__uint_as_float(v[j * 8 + 5]) + b1.y, 0.f);
        const float y6 = fmaxf(__uint_as_float(v[j * 8 + 6]) + b1.z, 0.f), y7 = fmaxf(__uint_as_float(v[j * 8 + 7]) + b1.w, 0.f);
        uint4 h, l;
        if (kFmt == kFmtF16) {
            h.x = pack_f16(y0, y1); h.y = pack_f16(y2, y3); h.z = pack_f16(y4, y5); h.w = pack_f16(y6, y7);
            *reinterpret_cast<uint4 *>(dst + (size_t)j * plane_stride) = h;
        } else {
            if (kFmt == kFmtF16x2) {
                split2_f16(y0, y1, h.x, l.x); split2_f16(y2, y3, h.y, l.y); split2_f16(y4, y5, h.z, l.z); split2_f16(y6, y7, h.w, l.w);
            } else {
                split2(y0, y1, h.x, l.x); split2(y2, y3, h.y, l.y); split2(y4, y5, h.z, l.z); split2(y6, y7, h.w, l.w);
            }
            *reinterpret_cast<uint4 *>(dst + (size_t)j * plane_stride) = h;
            *reinterpret_cast<uint4 *>(dst + lo_off + (size_t)j * plane_stride) = l;
        }
    }
}

// Accumulator columns [c0, c0 + ncols) of this thread's TMEM lane -> split operand planes (ncols multiple of 32).
template <int kFmt = kFmtBf16x2>
__device__ __forceinline__ void epilogue_split(uint32_t taddr, int c0, int ncols, uint8_t *buf, uint32_t lo_off, uint32_t plane_stride,
                                               int buf_row, const float *bias)
{
    for (int c = c0; c < c0 + ncols; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c, v);
        tmem_ld_wait();
        store_split32<kFmt>(buf + (size_t)(c >> 3) * plane_stride + (size_t)buf_row * 16, lo_off, plane_stride, v, bias + c);
    }
}

// First layer on CUDA cores: x[c] (c < c_in) -> output channels [ch0, ch0 + nch) -> split planes.  sw[c * w0 + ch].
__device__ __forceinline__ void first_layer_split(uint8_t *buf, uint32_t lo_off, int row, const float *xv, int c_in, int w0, int ch0, int nch,
                                                  const float *sw, const float *sb, uint32_t plane_stride = kPlane)
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
        uint4 h, l;
        split2(fmaxf(a0.x, 0.f), fmaxf(a0.y, 0.f), h.x, l.x); split2(fmaxf(a0.z, 0.f), fmaxf(a0.w, 0.f), h.y, l.y);
        split2(fmaxf(a1.x, 0.f), fmaxf(a1.y, 0.f), h.z, l.z); split2(fmaxf(a1.z, 0.f), fmaxf(a1.w, 0.f), h.w, l.w);
        uint8_t *dst = buf + (size_t)(ch >> 3) * plane_stride + (size_t)row * 16;
        *reinterpret_cast<uint4 *>(dst) = h;
        *reinterpret_cast<uint4 *>(dst + lo_off) = l;
    }
}

// The weight ring as the MMA issuer sees it.
struct RingView {
    uint32_t base;                 // shared-memory address of stage 0
    uint64_t *full, *empty;
    int n, stage; uint32_t phase;
};
#define SPLIT_RING_NEXT(r, code)                                                   \
    if (!mbar_wait(&(r).full[(r).stage], (r).phase, (code) + (r).stage, wd)) goto done; \
    tc_fence_after();                                                              \
    const uint32_t wst_ = (r).base + (uint32_t)(r).stage * kStage;
#define SPLIT_RING_RELEASE(r)                                                      \
    mma_commit(&(r).empty[(r).stage]);                                             \
    if (++(r).stage == (r).n) { (r).stage = 0; (r).phase ^= 1; }

// One K = 64 block of D[128 x rows] (+)= A[128 x 64] * W[rows x 64]^T in split precision: the hi weight block feeds
// A_hi and A_lo, the lo weight block A_hi.  a_hi: address of the first of the block's 8 planes (plane stride a_plane).
// Expands to two ring stages; `first` = the very first MMA of this accumulator (overwrite instead of accumulate).
#define SPLIT_MMA_BLOCK(r, d_tmem, a_hi, a_lo_off, a_plane, a_rows, w_rows, idesc, first, code)                      \
    {                                                                                                                \
        { SPLIT_RING_NEXT(r, code)                                                                                   \
          _Pragma("unroll")                                                                                          \
          for (int k_ = 0; k_ < 4; ++k_) {                                                                           \
              const uint64_t db_ = make_desc(wst_ + k_ * 2 * (w_rows) * 16, (w_rows));                               \
              mma_bf16((d_tmem), make_desc((a_hi) + k_ * 2 * (a_plane), (a_rows)), db_, (idesc), ((first) && k_ == 0) ? 0u : 1u); \
              mma_bf16((d_tmem), make_desc((a_hi) + (a_lo_off) + k_ * 2 * (a_plane), (a_rows)), db_, (idesc), 1u);   \
          }                                                                                                          \
          SPLIT_RING_RELEASE(r) }                                                                                    \
        { SPLIT_RING_NEXT(r, code)                                                                                   \
          _Pragma("unroll")                                                                                          \
          for (int k_ = 0; k_ < 4; ++k_)                                                                             \
              mma_bf16((d_tmem), make_desc((a_hi) + k_ * 2 * (a_plane), (a_rows)),                                   \
                       make_desc(wst_ + k_ * 2 * (w_rows) * 16, (w_rows)), (idesc), 1u);                             \
          SPLIT_RING_RELEASE(r) }                                                                                    \
    }

// Transposed form for the max-pooled last layer: D^T[128 channels x n_pts] (+)= W[128 x 64] * Act[n_pts x 64]^T.
#define SPLIT_MMA_BLOCK_T(r, d_tmem, b_hi, b_lo_off, b_plane, b_rows, idesc, first, code)                            \
    {                                                                                                                \
        { SPLIT_RING_NEXT(r, code)                                                                                   \
          _Pragma("unroll")                                                                                          \
          for (int k_ = 0; k_ < 4; ++k_) {                                                                           \
              const uint64_t da_ = make_desc(wst_ + k_ * 4096, 128);                                                 \
              mma_bf16((d_tmem), da_, make_desc((b_hi) + k_ * 2 * (b_plane), (b_rows)), (idesc), ((first) && k_ == 0) ? 0u : 1u); \
              mma_bf16((d_tmem), da_, make_desc((b_hi) + (b_lo_off) + k_ * 2 * (b_plane), (b_rows)), (idesc), 1u);   \
          }                                                                                                          \
          SPLIT_RING_RELEASE(r) }                                                                                    \
        { SPLIT_RING_NEXT(r, code)                                                                                   \
          _Pragma("unroll")                                                                                          \
          for (int k_ = 0; k_ < 4; ++k_)                                                                             \
              mma_bf16((d_tmem), make_desc(wst_ + k_ * 4096, 128), make_desc((b_hi) + k_ * 2 * (b_plane), (b_rows)), (idesc), 1u); \
          SPLIT_RING_RELEASE(r) }                                                                                    \
    }

// fp16 forms (one weight slot per block, fp16 idesc).  kTerms = 1: D (+)= A_hi * W; kTerms = 2: D (+)= (A_hi + A_lo) * W.
#define SPLIT_MMA_BLOCK_F16(r, d_tmem, a_hi, a_lo_off, a_plane, a_rows, w_rows, idesc, first, two_terms, code)       \
    {                                                                                                                \
        SPLIT_RING_NEXT(r, code)                                                                                     \
        _Pragma("unroll")                                                                                            \
        for (int k_ = 0; k_ < 4; ++k_) {                                                                             \
            const uint64_t db_ = make_desc(wst_ + k_ * 2 * (w_rows) * 16, (w_rows));                                 \
            mma_bf16((d_tmem), make_desc((a_hi) + k_ * 2 * (a_plane), (a_rows)), db_, (idesc), ((first) && k_ == 0) ? 0u : 1u); \
            if (two_terms) mma_bf16((d_tmem), make_desc((a_hi) + (a_lo_off) + k_ * 2 * (a_plane), (a_rows)), db_, (idesc), 1u); \
        }                                                                                                            \
        SPLIT_RING_RELEASE(r)                                                                                        \
    }
// Transposed, one term: D^T[128 channels x n_pts] (+)= W[128 x 64] * Act_hi[n_pts x 64]^T.
#define SPLIT_MMA_BLOCK_T_F16(r, d_tmem, b_hi, b_plane, b_rows, idesc, first, code)                                  \
    {                                                                                                                \
        SPLIT_RING_NEXT(r, code)                                                                                     \
        _Pragma("unroll")                                                                                            \
        for (int k_ = 0; k_ < 4; ++k_)                                                                               \
            mma_bf16((d_tmem), make_desc(wst_ + k_ * 4096, 128), make_desc((b_hi) + k_ * 2 * (b_plane), (b_rows)), (idesc), \
                     ((first) && k_ == 0) ? 0u : 1u);                                                                \
        SPLIT_RING_RELEASE(r)                                                                                        \
    }

}  // namespace split
}  // namespace al3d
