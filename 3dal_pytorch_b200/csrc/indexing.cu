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
// mask + ordered compaction: one CTA per object, 256 points per sweep.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mask_compact_kernel(const float *__restrict__ logits, uint8_t *__restrict__ mask, int n,
                    int32_t *__restrict__ pos, int32_t *__restrict__ count)
{
    __shared__ int warp_cnt[8];
    __shared__ int base_s;
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) base_s = 0;
    __syncthreads();
    const float2 *lg = logits ? reinterpret_cast<const float2 *>(logits) + (int64_t)b * n : nullptr;
    uint8_t *mk = mask + (int64_t)b * n;
    int32_t *ps = pos + (int64_t)b * n;
    for (int p0 = 0; p0 < n; p0 += 256) {
        const int p = p0 + threadIdx.x;
        bool fg = false;
        if (p < n) {
            if (lg) {
                const float2 l = __ldg(lg + p);
                fg = l.x < l.y;                       // strict; false when either is NaN
                mk[p] = fg ? 1 : 0;
            } else {
                fg = mk[p] != 0;
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, fg);
        if (lane == 0) warp_cnt[wid] = __popc(bal);
        __syncthreads();
        int before = base_s;
        for (int w = 0; w < wid; ++w) before += warp_cnt[w];
        if (fg) ps[before + __popc(bal & ((1u << lane) - 1u))] = p;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 8; ++w) t += warp_cnt[w];
            base_s += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) count[b] = base_s;
}

// ---------------------------------------------------------------------------------------------
// gather: one CTA per object.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather_fg_kernel(const float *__restrict__ x, int64_t sb, int64_t sc, int64_t sp, int C, int n,
                 const int32_t *__restrict__ pos, const int32_t *__restrict__ count, int policy,
                 const int32_t *__restrict__ choice, int n_pts, float *__restrict__ out, int64_t *__restrict__ indices)
{
    const int b = blockIdx.x;
    const int L = count[b];
    const int32_t *ps = pos + (int64_t)b * n;
    for (int j = threadIdx.x; j < n_pts; j += blockDim.x) {
        int64_t src = 0;
        if (L > 0) {
            int sel;
            if (policy == AL3D_GATHER_TABLE) sel = choice[(int64_t)b * n_pts + j];
            else sel = (L >= n_pts) ? (int)(((uint64_t)j * (uint64_t)L) / (uint64_t)n_pts) : (j % L);
            src = ps[sel];
        }
        if (indices) indices[(int64_t)b * n_pts + j] = src;
        for (int c = 0; c < C; ++c) {
            const float v = (L > 0) ? __ldg(x + b * sb + c * sc + src * sp) : 0.f;
            out[((int64_t)b * C + c) * n_pts + j] = v;
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
    mask_compact_kernel<<<bs, 256, 0, (cudaStream_t)stream>>>(logits, mask, n, pos, count);
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
    gather_fg_kernel<<<bs, 256, 0, (cudaStream_t)stream>>>(x, sb, sc, sp, C, n, pos, count, policy, choice, n_pts, out, indices);
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
