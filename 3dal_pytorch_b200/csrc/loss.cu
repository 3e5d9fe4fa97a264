// Fused forward of the auto-label losses (eval / logging path; no backward yet).
//
// Reference: FrustumPointNetLossOneBoxEst.forward tools/static_model.py:348-425, ...TwoBoxEst :427-517,
// DynamicModelLoss tools/dynamic_model.py:321-398, huber_loss tools/static_model.py:341-346.  One launch
// computes the segmentation NLL over all bs*n points (deterministic two-level reduction), a second one the
// five box-head terms of one head set; the weighted sums are formed by the caller.
#include "common.cuh"
#include "../../include/al3d.h"

namespace al3d {

__constant__ float c_loss_mean_size[9] = {4.8f, 1.8f, 1.5f, 10.0f, 2.6f, 3.2f, 2.0f, 1.0f, 1.6f};

__device__ __forceinline__ float block_sum_256(float v, float *red)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];      // fixed order
    return t;
}

__device__ __forceinline__ float huber(float err, float delta)
{
    const float a = fabsf(err), q = fminf(a, delta);
    return 0.5f * q * q + delta * (a - q);
}

// partial[blockIdx.x] = sum over this block's points of -log_softmax(logits)[label]
__global__ void __launch_bounds__(256)
loss_mask_partial_kernel(const float *__restrict__ logits, const float *__restrict__ label, int64_t M, float *__restrict__ partial)
{
    __shared__ float red[8];
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < M; i += (int64_t)gridDim.x * 256) {
        const float2 l = __ldg(reinterpret_cast<const float2 *>(logits) + i);
        const float m = fmaxf(l.x, l.y);
        const float lse = m + logf(expf(l.x - m) + expf(l.y - m));
        const int lab = (int)(long long)__ldg(label + i);           // mask_label.long()
        acc += lse - (lab == 0 ? l.x : l.y);
    }
    const float t = block_sum_256(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// out[0] = mean mask NLL (from the partials); out[1..5] = center, heading-class, size-class, heading-residual,
// size-residual losses of one head set (means over bs).  Single block.
__global__ void __launch_bounds__(256)
loss_heads_kernel(const float *__restrict__ partial, int n_partial, int64_t M, const float *__restrict__ center,
                  const float *__restrict__ center_label, const float *__restrict__ hs, const int64_t *__restrict__ hcls,
                  const float *__restrict__ hrn, const float *__restrict__ hres_label, const float *__restrict__ ss,
                  const int64_t *__restrict__ scls, const float *__restrict__ srn, const float *__restrict__ sres_label, int bs,
                  float *__restrict__ out)
{
    __shared__ float red[8];
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (partial) for (int i = threadIdx.x; i < n_partial; i += 256) acc[0] += partial[i];
    for (int b = threadIdx.x; b < bs; b += 256) {
        float dx = center[b * 3] - center_label[b * 3], dy = center[b * 3 + 1] - center_label[b * 3 + 1], dz = center[b * 3 + 2] - center_label[b * 3 + 2];
        acc[1] += huber(sqrtf(dx * dx + dy * dy + dz * dz), 2.0f);
        {
            const float *v = hs + b * 12;
            float m = v[0];
            for (int i = 1; i < 12; ++i) m = fmaxf(m, v[i]);
            float se = 0.f;
            for (int i = 0; i < 12; ++i) se += expf(v[i] - m);
            const int c = (int)hcls[b];
            acc[2] += (m + logf(se)) - v[c];
            const float lab = hres_label[b] / 0.2617993877991494f;      // / (pi / 12)
            acc[4] += huber(hrn[b * 12 + c] - lab, 1.0f);
        }
        {
            const float *v = ss + b * 3;
            const float m = fmaxf(v[0], fmaxf(v[1], v[2]));
            const float se = expf(v[0] - m) + expf(v[1] - m) + expf(v[2] - m);
            const int c = (int)scls[b];
            acc[3] += (m + logf(se)) - v[c];
            float d2 = 0.f;
            for (int k = 0; k < 3; ++k) {
                const float e = sres_label[b * 3 + k] / c_loss_mean_size[c * 3 + k] - srn[b * 9 + c * 3 + k];
                d2 += e * e;
            }
            acc[5] += huber(sqrtf(d2), 1.0f);
        }
    }
    for (int k = 0; k < 6; ++k) {
        const float t = block_sum_256(acc[k], red);
        if (threadIdx.x == 0) out[k] = (k == 0) ? (partial ? t / (float)M : 0.f) : t / (float)bs;
    }
}

}  // namespace al3d

using namespace al3d;

extern "C" int al3d_loss_forward(const float *logits, const float *mask_label, int64_t M, const float *center,
                                 const float *center_label, const float *heading_scores, const int64_t *heading_cls_label,
                                 const float *heading_res_norm, const float *heading_res_label, const float *size_scores,
                                 const int64_t *size_cls_label, const float *size_res_norm, const float *size_res_label, int bs,
                                 float *partial_ws, int n_partial, float *out6, void *stream)
{
    AL3D_CHECK_ARG(center && center_label && heading_scores && heading_cls_label && heading_res_norm && heading_res_label &&
                   size_scores && size_cls_label && size_res_norm && size_res_label && out6, "al3d_loss_forward: null pointer");
    AL3D_CHECK_ARG(bs > 0, "al3d_loss_forward: bs must be positive");
    const bool with_mask = logits != nullptr;
    AL3D_CHECK_ARG(!with_mask || (mask_label && partial_ws && n_partial >= 1 && n_partial <= 4096 && M > 0),
                   "al3d_loss_forward: mask term needs labels and 1..4096 partial slots");
    if (with_mask) {
        loss_mask_partial_kernel<<<n_partial, 256, 0, (cudaStream_t)stream>>>(logits, mask_label, M, partial_ws);
        AL3D_CHECK_LAUNCH("loss_mask_partial_kernel");
    }
    loss_heads_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(with_mask ? partial_ws : nullptr, n_partial, M, center, center_label,
                                                           heading_scores, heading_cls_label, heading_res_norm, heading_res_label,
                                                           size_scores, size_cls_label, size_res_norm, size_res_label, bs, out6);
    AL3D_CHECK_LAUNCH("loss_heads_kernel");
    return 0;
}
