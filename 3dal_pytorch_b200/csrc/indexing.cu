// Integer / indexing stages of the auto-label forward: foreground mask, ordered compaction, gather,
// head parsing, box decode and the two-stage canonical re-transform.  Bit-exact by construction:
// ascending-index ordering comes from ballot + popc prefixes, never from atomics.
//
// Reference: point_cloud_masking / gather_object_pts tools/static_model.py:23-62
// (tools/dynamic_model.py:24-63), parse_output_to_tensors :64-96, StaticModelTwoBoxEst.forward
// :177-205, class2angle / class2size / angle2class tools/utils.py:53-79.
#include "common.cuh"
#include "../../include/al3d.h"

namespace al3d {

__constant__ float  c_mean_size_f32[9] = {4.8f, 1.8f, 1.5f, 10.0f, 2.6f, 3.2f, 2.0f, 1.0f, 1.6f};
__constant__ double c_mean_size_f64[9] = {4.8, 1.8, 1.5, 10.0, 2.6, 3.2, 2.0, 1.0, 1.6};

// ---------------------------------------------------------------------------------------------
// mask + ordered compaction: one CTA of 128 threads per object, 4096 points per sweep -- 32 consecutive points per thread
// (vector loads, all issued before the first is used; a 32-bit foreground word), one block-wide exclusive scan of the
// per-thread counts, then every thread writes its own foreground indices.  Two barriers per sweep (the first cut swept
// 256 points at a time with three barriers each: 0.65 TB/s on 8192 x 4096 points).
// ---------------------------------------------------------------------------------------------
constexpr int kCompactThreads = 128;
constexpr int kCompactPer = 32;                        // points per thread and sweep

__global__ void __launch_bounds__(kCompactThreads)
mask_compact_kernel(const float *__restrict__ logits, uint8_t *__restrict__ mask, int n,
                    int32_t *__restrict__ pos, int32_t *__restrict__ count)
{
    __shared__ int warp_cnt[kCompactThreads / 32];
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const float2 *lg = logits ? reinterpret_cast<const float2 *>(logits) + (int64_t)b * n : nullptr;
    uint8_t *mk = mask + (int64_t)b * n;
    int32_t *ps = pos + (int64_t)b * n;
    const bool vec = (n & 15) == 0 && (reinterpret_cast<uintptr_t>(mask) & 15) == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0;
    int base = 0;                                     // foreground points of the sweeps done (the same in every thread)
    for (int p0 = 0; p0 < n; p0 += kCompactThreads * kCompactPer) {
        const int t0 = p0 + threadIdx.x * kCompactPer;
        unsigned bits = 0;
        if (vec && t0 + kCompactPer <= n) {
            if (lg) {
                const float4 *q = reinterpret_cast<const float4 *>(lg + t0);
                float4 v[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = __ldg(q + k);
#pragma unroll
                for (int k = 0; k < 16; ++k)          // strict; false when either logit is NaN
                    bits |= (v[k].x < v[k].y ? 1u << (2 * k) : 0u) | (v[k].z < v[k].w ? 2u << (2 * k) : 0u);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t w[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const unsigned x = (bits >> (16 * h + 4 * j)) & 0xFu;
                        w[j] = (x & 1u) | ((x & 2u) << 7) | ((x & 4u) << 14) | ((x & 8u) << 21);
                    }
                    *reinterpret_cast<uint4 *>(mk + t0 + 16 * h) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            } else {
                const uint4 m0 = *reinterpret_cast<const uint4 *>(mk + t0), m1 = *reinterpret_cast<const uint4 *>(mk + t0 + 16);
                const uint32_t w[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    // four mask bytes -> four bits: non-zero bytes to 0x01 each, then one multiply gathers them in bits 24..27
                    const uint32_t nz = __vcmpne4(w[j], 0u) & 0x01010101u;
                    bits |= ((nz * 0x01020408u) >> 24) << (4 * j);
                }
            }
        } else {
            for (int k = 0; k < kCompactPer; ++k) {
                const int p = t0 + k;
                if (p >= n) break;
                bool fg;
                if (lg) {
                    const float2 l = __ldg(lg + p);
                    fg = l.x < l.y;
                    mk[p] = fg ? 1 : 0;
                } else fg = mk[p] != 0;
                bits |= fg ? 1u << k : 0u;
            }
        }
        const int cnt = __popc(bits);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) warp_cnt[wid] = incl;
        __syncthreads();
        int before = base + incl - cnt, total = 0;
#pragma unroll
        for (int w = 0; w < kCompactThreads / 32; ++w) { const int c = warp_cnt[w]; before += w < wid ? c : 0; total += c; }
        while (bits) {                                 // ascending: lowest set bit first
            const int k = __ffs(bits) - 1;
            bits &= bits - 1;
            ps[before++] = t0 + k;
        }
        base += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) count[b] = base;
}

// ---------------------------------------------------------------------------------------------
// gather: one CTA of 128 threads per object, four output slots per thread and round -- the four index loads, then the
// twelve coordinate loads are all in flight together.
// ---------------------------------------------------------------------------------------------
constexpr int kGatherThreads = 128;

__global__ void __launch_bounds__(kGatherThreads)
gather_fg_kernel(const float *__restrict__ x, int64_t sb, int64_t sc, int64_t sp, int C, int n,
                 const int32_t *__restrict__ pos, const int32_t *__restrict__ count, int policy,
                 const int32_t *__restrict__ choice, int n_pts, float *__restrict__ out, int64_t *__restrict__ indices)
{
    const int b = blockIdx.x;
    const int L = count[b];
    const int32_t *ps = pos + (int64_t)b * n;
    const float *xb = x + b * sb;
    const bool small = (uint64_t)n_pts * (uint64_t)max(L, 1) < (1ull << 32);      // j * L fits 32 bits: no 64-bit division
    for (int j0 = threadIdx.x; j0 < n_pts; j0 += 4 * kGatherThreads) {
        int64_t src[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * kGatherThreads;
            src[u] = 0;
            if (L > 0 && j < n_pts) {
                int sel;
                if (policy == AL3D_GATHER_TABLE) sel = choice[(int64_t)b * n_pts + j];
                else if (L < n_pts) sel = j % L;
                else sel = small ? (int)(((unsigned)j * (unsigned)L) / (unsigned)n_pts) : (int)(((uint64_t)j * (uint64_t)L) / (uint64_t)n_pts);
                src[u] = ps[sel];
            }
        }
        if (C == 3) {
            float v[4][3];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float *q = xb + src[u] * sp;
                const bool ok = L > 0 && j0 + u * kGatherThreads < n_pts;
                v[u][0] = ok ? __ldg(q) : 0.f; v[u][1] = ok ? __ldg(q + sc) : 0.f; v[u][2] = ok ? __ldg(q + 2 * sc) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u * kGatherThreads;
                if (j >= n_pts) continue;
                if (indices) indices[(int64_t)b * n_pts + j] = src[u];
                float *o = out + (int64_t)b * 3 * n_pts + j;
                o[0] = v[u][0]; o[n_pts] = v[u][1]; o[2 * (int64_t)n_pts] = v[u][2];
            }
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u * kGatherThreads;
                if (j >= n_pts) continue;
                if (indices) indices[(int64_t)b * n_pts + j] = src[u];
                for (int c = 0; c < C; ++c)
                    out[((int64_t)b * C + c) * n_pts + j] = (L > 0) ? __ldg(xb + c * sc + src[u] * sp) : 0.f;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// parse heads: one thread per (object, element of the 39-vector)
// ---------------------------------------------------------------------------------------------
__global__ void parse_heads_kernel(const float *__restrict__ bp, int bs, const float *__restrict__ add, int64_t add_stride,
                                   float *center_boxnet, float *center, float *hs, float *hrn, float *hr,
                                   float *ss, float *srn, float *sr)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= bs * 39) return;
    const int b = t / 39, e = t - b * 39;
    const float v = bp[t];
    if (e < 3) {
        if (center_boxnet) center_boxnet[b * 3 + e] = v;
        if (center) center[b * 3 + e] = add ? v + add[b * add_stride + e] : v;
    } else if (e < 15) {
        if (hs) hs[b * 12 + (e - 3)] = v;
    } else if (e < 27) {
        if (hrn) hrn[b * 12 + (e - 15)] = v;
        if (hr) hr[b * 12 + (e - 15)] = v * 0.2617993877991494f;            // (float)(pi/12)
    } else if (e < 30) {
        if (ss) ss[b * 3 + (e - 27)] = v;
    } else {
        if (srn) srn[b * 9 + (e - 30)] = v;
        if (sr) sr[b * 9 + (e - 30)] = v * c_mean_size_f32[e - 30];
    }
}

// first-maximum argmax (numpy semantics; NaN handling: a NaN compares false and is never picked
// unless it is element 0 -- numpy would propagate NaN; heads never produce NaN on finite inputs)
__device__ __forceinline__ int argmax_first(const float *v, int n)
{
    int best = 0;
    float bv = v[0];
    for (int i = 1; i < n; ++i) if (v[i] > bv) { bv = v[i]; best = i; }
    return best;
}

__global__ void decode_boxes_kernel(const float *__restrict__ center, const float *__restrict__ hs,
                                    const float *__restrict__ hr, const float *__restrict__ ss,
                                    const float *__restrict__ sr, const float *__restrict__ base, int64_t base_stride,
                                    int bs, float *__restrict__ box, int32_t *__restrict__ cls)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= bs) return;
    const int hc = argmax_first(hs + b * 12, 12);
    const int sc = argmax_first(ss + b * 3, 3);
    const double per = 2.0 * 3.141592653589793 / 12.0;
    double ang = (double)hc * per + (double)hr[b * 12 + hc];
    if (ang > 3.141592653589793) ang = ang - 2.0 * 3.141592653589793;
    if (base) ang += (double)base[b * base_stride];
    for (int k = 0; k < 3; ++k) {
        box[b * 7 + k] = center[b * 3 + k];
        box[b * 7 + 3 + k] = (float)(c_mean_size_f64[sc * 3 + k] + (double)sr[b * 9 + sc * 3 + k]);
    }
    box[b * 7 + 6] = (float)ang;
    if (cls) { cls[b * 2] = hc; cls[b * 2 + 1] = sc; }
}

// torch.remainder on f32 (fmod + sign fix-up), the `%` of tools/utils.py:54,57 on 0-dim tensors
__device__ __forceinline__ float torch_remainder(float x, float m)
{
    float r = fmodf(x, m);
    if (r != 0.f && ((r < 0.f) != (m < 0.f))) r = __fadd_rn(r, m);
    return r;
}

__global__ void __launch_bounds__(128)
twostage_retransform_kernel(const float *__restrict__ obj, int m, const float *__restrict__ init_box,
                            const float *__restrict__ box_one, const float *__restrict__ gt,
                            float *__restrict__ obj2, int64_t *__restrict__ cls_label, float *__restrict__ res_label)
{
    const int b = blockIdx.x;
    const float h0 = init_box[b * 7 + 6], h1 = box_one[b * 7 + 6];
    const float c0 = cosf(h0), s0 = sinf(h0);
    const float c1 = cosf(-h1), s1 = sinf(-h1);
    const float tx = init_box[b * 7 + 0], ty = init_box[b * 7 + 1], tz = init_box[b * 7 + 2];
    const float ux = box_one[b * 7 + 0], uy = box_one[b * 7 + 1], uz = box_one[b * 7 + 2];
    const float *px = obj + (int64_t)b * 3 * m;
    float *qx = obj2 + (int64_t)b * 3 * m;
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const float x = px[j], y = px[m + j], z = px[2 * m + j];
        // P <- Rz(h0) P + c0 - c1 ; P <- Rz(-h1) P      (tools/static_model.py:195-200)
        float ax = c0 * x - s0 * y, ay = s0 * x + c0 * y, az = z;
        ax = (ax + tx) - ux; ay = (ay + ty) - uy; az = (az + tz) - uz;
        qx[j] = c1 * ax - s1 * ay;
        qx[m + j] = s1 * ax + c1 * ay;
        qx[2 * m + j] = az;
    }
    if (threadIdx.x == 0 && cls_label) {
        // angle2class(gt_heading - h1, 12) evaluated in f32 (tools/utils.py:53-60)
        const float two_pi = 6.283185307179586f;
        const float per = (float)(6.283185307179586 / 12.0);
        const float half = (float)((6.283185307179586 / 12.0) / 2.0);
        const float a = torch_remainder(__fsub_rn(gt[b * 7 + 6], h1), two_pi);
        const float shifted = torch_remainder(__fadd_rn(a, half), two_pi);
        const int cid = (int)__fdiv_rn(shifted, per);
        const float centre = (float)((double)cid * (6.283185307179586 / 12.0) + (6.283185307179586 / 12.0) / 2.0);
        cls_label[b] = cid;
        res_label[b] = __fsub_rn(shifted, centre);
    }
}

}  // namespace al3d

using namespace al3d;

extern "C" int al3d_mask_compact(const float *logits, uint8_t *mask, int bs, int n, int32_t *pos, int32_t *count, void *stream)
{
    AL3D_CHECK_ARG(mask && pos && count, "al3d_mask_compact: null pointer");
    AL3D_CHECK_ARG(bs >= 0 && n >= 0, "al3d_mask_compact: negative size");
    if (bs == 0) return 0;
    mask_compact_kernel<<<bs, kCompactThreads, 0, (cudaStream_t)stream>>>(logits, mask, n, pos, count);
    AL3D_CHECK_LAUNCH("mask_compact_kernel");
    return 0;
}

extern "C" int al3d_gather_fg(const float *x, int64_t sb, int64_t sc, int64_t sp, int bs, int C, int n,
                              const int32_t *pos, const int32_t *count, int policy, const int32_t *choice, int n_pts,
                              float *out, int64_t *indices, void *stream)
{
    AL3D_CHECK_ARG(x && pos && count && out, "al3d_gather_fg: null pointer");
    AL3D_CHECK_ARG(policy == AL3D_GATHER_STRIDED || (policy == AL3D_GATHER_TABLE && choice),
                   "al3d_gather_fg: bad policy %d (or missing choice table)", policy);
    AL3D_CHECK_ARG(bs >= 0 && C > 0 && n >= 0 && n_pts > 0, "al3d_gather_fg: bad shape");
    if (bs == 0) return 0;
    gather_fg_kernel<<<bs, kGatherThreads, 0, (cudaStream_t)stream>>>(x, sb, sc, sp, C, n, pos, count, policy, choice, n_pts, out, indices);
    AL3D_CHECK_LAUNCH("gather_fg_kernel");
    return 0;
}

extern "C" int al3d_parse_heads(const float *box_pred, int bs, const float *add, int64_t add_stride,
                                float *center_boxnet, float *center, float *heading_scores, float *heading_res_norm,
                                float *heading_res, float *size_scores, float *size_res_norm, float *size_res, void *stream)
{
    AL3D_CHECK_ARG(box_pred, "al3d_parse_heads: null box_pred");
    if (bs <= 0) return 0;
    const int total = bs * 39;
    parse_heads_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(box_pred, bs, add, add_stride, center_boxnet, center,
                                                                             heading_scores, heading_res_norm, heading_res,
                                                                             size_scores, size_res_norm, size_res);
    AL3D_CHECK_LAUNCH("parse_heads_kernel");
    return 0;
}

extern "C" int al3d_decode_boxes(const float *center, const float *heading_scores, const float *heading_res,
                                 const float *size_scores, const float *size_res, const float *base_heading,
                                 int64_t base_stride, int bs, float *box_out, int32_t *cls_out, void *stream)
{
    AL3D_CHECK_ARG(center && heading_scores && heading_res && size_scores && size_res && box_out, "al3d_decode_boxes: null pointer");
    if (bs <= 0) return 0;
    decode_boxes_kernel<<<(bs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(center, heading_scores, heading_res, size_scores, size_res,
                                                                           base_heading, base_stride, bs, box_out, cls_out);
    AL3D_CHECK_LAUNCH("decode_boxes_kernel");
    return 0;
}

extern "C" int al3d_twostage_retransform(const float *obj_pts, int bs, int m, const float *init_box, const float *box_one,
                                         const float *bbox_gt, float *obj_pts_two, int64_t *heading_cls_label,
                                         float *heading_res_label, void *stream)
{
    AL3D_CHECK_ARG(obj_pts && init_box && box_one && obj_pts_two, "al3d_twostage_retransform: null pointer");
    AL3D_CHECK_ARG(!heading_cls_label || (bbox_gt && heading_res_label), "al3d_twostage_retransform: labels need bbox_gt");
    if (bs <= 0) return 0;
    twostage_retransform_kernel<<<bs, 128, 0, (cudaStream_t)stream>>>(obj_pts, m, init_box, box_one, bbox_gt, obj_pts_two,
                                                                     heading_cls_label, heading_res_label);
    AL3D_CHECK_LAUNCH("twostage_retransform_kernel");
    return 0;
}
