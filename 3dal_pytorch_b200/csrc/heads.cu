// Fused FC heads: the whole fully-connected chain of a box head (or embedding) in ONE launch, with head parsing, the
// centre residual add and the box decode in its epilogue.
//
// Replaces, per model: static PointNetEstimation fc1 -> fc2 -> fc3 (tools/static_model.py:336-338) +
// parse_output_to_tensors (:64-96) + the centre add (:132,174,211) + the eval-side decode (tools/static_eval.py:270-288);
// dynamic PointEmbedding / BoxEmbedding fc1 -> fc2 (tools/dynamic_model.py:247-248, 284-285) and the dynamic head on
// cat[point_e, box_e] (:137, 307-311).  BatchNorm is folded by the caller; weights arrive TRANSPOSED (K, N) so that a
// warp reads 32 consecutive output channels of one k.
//
// One CTA takes kFcObjs objects: their activations ping-pong between two shared-memory buffers, every thread owns one
// output channel of the current layer for all kFcObjs objects (register accumulators, fixed k order -> the result of
// an object does not depend on how the batch is cut), weights stream from L2.  fp32 SIMT on purpose: these layers are
// 0.01 % of the model's FLOPs and feed arg-max decisions.
#include "common.cuh"
#include "../../include/al3d.h"

namespace al3d {
namespace heads {

constexpr int kFcObjs = 16;
constexpr int kFcThreads = 256;
constexpr int kFcMaxW = 512;           // widest activation row

__constant__ float c_mean_size_f[9] = {4.8f, 1.8f, 1.5f, 10.0f, 2.6f, 3.2f, 2.0f, 1.0f, 1.6f};
__constant__ double c_mean_size_d[9] = {4.8, 1.8, 1.5, 10.0, 2.6, 3.2, 2.0, 1.0, 1.6};

struct FcChainParams {
    const float *x0; int k0; int64_t ld0;
    const float *x1; int k1; int64_t ld1;            // optional second input, concatenated after x0
    int n_layers; int width[3]; int relu[3];
    const float *wt[3]; const float *bias[3];        // W^T (K, N) row-major, bias (N)
    int bs;
    float *out; int64_t ldo;                         // output of the last layer (may be NULL)
    int heads;                                       // 1: the last layer is the 39-wide head vector -> parse (+ decode)
    const float *add; int64_t add_stride;            // centre residual base (init_box / centre of stage one), may be NULL
    const float *base_heading; int64_t base_stride;  // heading base of the decode, may be NULL
    float *center_boxnet, *center, *hs, *hrn, *hr, *ss, *srn, *sr, *box;     // any may be NULL
    int32_t *cls;
};

__device__ __forceinline__ int argmax_first(const float *v, int n)
{
    int best = 0;
    for (int i = 1; i < n; ++i) if (v[i] > v[best]) best = i;
    return best;
}

__global__ void __launch_bounds__(kFcThreads)
fc_chain_kernel(const FcChainParams p)
{
    extern __shared__ __align__(16) float sm[];
    float *a = sm, *b = sm + kFcObjs * kFcMaxW;
    const int tid = threadIdx.x;
    const int o0 = blockIdx.x * kFcObjs;
    const int nobj = min(kFcObjs, p.bs - o0);
    int K = p.k0 + p.k1;
    for (int i = tid; i < kFcObjs * K; i += kFcThreads) {
        const int o = i / K, k = i - o * K;
        float v = 0.f;
        if (o < nobj) v = k < p.k0 ? __ldg(p.x0 + (int64_t)(o0 + o) * p.ld0 + k) : __ldg(p.x1 + (int64_t)(o0 + o) * p.ld1 + (k - p.k0));
        a[o * kFcMaxW + k] = v;
    }
    __syncthreads();
    for (int l = 0; l < p.n_layers; ++l) {
        const int N = p.width[l];
        const float *wt = p.wt[l];
        for (int ch = tid; ch < N; ch += kFcThreads) {
            float acc[kFcObjs];
            const float bv = __ldg(p.bias[l] + ch);
#pragma unroll
            for (int o = 0; o < kFcObjs; ++o) acc[o] = bv;
            for (int k = 0; k < K; k += 4) {
                const float w0 = __ldg(wt + (int64_t)k * N + ch), w1 = __ldg(wt + (int64_t)(k + 1) * N + ch);
                const float w2 = __ldg(wt + (int64_t)(k + 2) * N + ch), w3 = __ldg(wt + (int64_t)(k + 3) * N + ch);
#pragma unroll
                for (int o = 0; o < kFcObjs; ++o) {
                    const float4 x = *reinterpret_cast<const float4 *>(a + o * kFcMaxW + k);      // broadcast read
                    acc[o] = fmaf(x.x, w0, acc[o]); acc[o] = fmaf(x.y, w1, acc[o]);
                    acc[o] = fmaf(x.z, w2, acc[o]); acc[o] = fmaf(x.w, w3, acc[o]);
                }
            }
#pragma unroll
            for (int o = 0; o < kFcObjs; ++o) b[o * kFcMaxW + ch] = p.relu[l] ? fmaxf(acc[o], 0.f) : acc[o];
        }
        __syncthreads();
        float *t = a; a = b; b = t;
        K = N;
    }
    // a: (kFcObjs, K) outputs of the last layer
    if (p.out)
        for (int i = tid; i < nobj * K; i += kFcThreads) {
            const int o = i / K, k = i - o * K;
            p.out[(int64_t)(o0 + o) * p.ldo + k] = a[o * kFcMaxW + k];
        }
    if (!p.heads) return;
    // ---- parse_output_to_tensors + centre add
    for (int i = tid; i < nobj * 39; i += kFcThreads) {
        const int o = i / 39, e = i - o * 39;
        const int64_t g = o0 + o;
        const float v = a[o * kFcMaxW + e];
        if (e < 3) {
            const float c = p.add ? v + __ldg(p.add + g * p.add_stride + e) : v;
            if (p.center_boxnet) p.center_boxnet[g * 3 + e] = v;
            if (p.center) p.center[g * 3 + e] = c;
            b[o * 4 + e] = c;                                   // kept for the decode below
        } else if (e < 15) {
            if (p.hs) p.hs[g * 12 + (e - 3)] = v;
        } else if (e < 27) {
            if (p.hrn) p.hrn[g * 12 + (e - 15)] = v;
            if (p.hr) p.hr[g * 12 + (e - 15)] = v * 0.2617993877991494f;            // (float)(pi / 12)
        } else if (e < 30) {
            if (p.ss) p.ss[g * 3 + (e - 27)] = v;
        } else {
            if (p.srn) p.srn[g * 9 + (e - 30)] = v;
            if (p.sr) p.sr[g * 9 + (e - 30)] = v * c_mean_size_f[e - 30];
        }
    }
    if (!p.box) return;
    __syncthreads();
    // ---- decode (tools/static_eval.py:270-288, tools/utils.py:69-79): arg-max heads, class2angle / class2size in float64
    if (tid < nobj) {
        const int64_t g = o0 + tid;
        const float *v = a + tid * kFcMaxW;
        const int hc = argmax_first(v + 3, 12), sc = argmax_first(v + 27, 3);
        const double per = 2.0 * 3.141592653589793 / 12.0;
        double ang = (double)hc * per + (double)(v[15 + hc] * 0.2617993877991494f);
        if (ang > 3.141592653589793) ang = ang - 2.0 * 3.141592653589793;
        if (p.base_heading) ang += (double)__ldg(p.base_heading + g * p.base_stride);
        for (int k = 0; k < 3; ++k) {
            p.box[g * 7 + k] = b[tid * 4 + k];
            p.box[g * 7 + 3 + k] = (float)(c_mean_size_d[sc * 3 + k] + (double)(v[30 + sc * 3 + k] * c_mean_size_f[sc * 3 + k]));
        }
        p.box[g * 7 + 6] = (float)ang;
        if (p.cls) { p.cls[g * 2] = hc; p.cls[g * 2 + 1] = sc; }
    }
}

}  // namespace heads
}  // namespace al3d

using namespace al3d;

extern "C" int al3d_fc_chain(const al3d_fc_chain_desc *d, int bs, void *stream)
{
    using namespace al3d::heads;
    AL3D_CHECK_ARG(d && d->x0, "al3d_fc_chain: null pointer");
    AL3D_CHECK_ARG(d->n_layers >= 1 && d->n_layers <= 3, "al3d_fc_chain: n_layers=%d", d->n_layers);
    AL3D_CHECK_ARG(d->k0 >= 0 && d->k1 >= 0 && (d->k0 + d->k1) % 4 == 0 && d->k0 + d->k1 >= 4 && d->k0 + d->k1 <= kFcMaxW &&
                   (d->k1 == 0 || d->x1), "al3d_fc_chain: bad input widths %d + %d", d->k0, d->k1);
    FcChainParams p;
    p.x0 = d->x0; p.k0 = d->k0; p.ld0 = d->ld0; p.x1 = d->x1; p.k1 = d->k1; p.ld1 = d->ld1;
    p.n_layers = d->n_layers;
    for (int l = 0; l < 3; ++l) {
        p.width[l] = d->width[l]; p.relu[l] = d->relu[l]; p.wt[l] = d->wt[l]; p.bias[l] = d->bias[l];
        if (l < d->n_layers) {
            AL3D_CHECK_ARG(d->wt[l] && d->bias[l] && d->width[l] >= 1 && d->width[l] <= kFcMaxW, "al3d_fc_chain: layer %d", l);
            AL3D_CHECK_ARG(l == d->n_layers - 1 || d->width[l] % 4 == 0, "al3d_fc_chain: inner width %d must be a multiple of 4", d->width[l]);
        }
    }
    p.bs = bs; p.out = d->out; p.ldo = d->ldo; p.heads = d->heads;
    AL3D_CHECK_ARG(!d->heads || d->width[d->n_layers - 1] == 39, "al3d_fc_chain: the head epilogue needs a 39-wide last layer");
    AL3D_CHECK_ARG(d->heads || d->out, "al3d_fc_chain: no output");
    p.add = d->add; p.add_stride = d->add_stride; p.base_heading = d->base_heading; p.base_stride = d->base_stride;
    p.center_boxnet = d->center_boxnet; p.center = d->center; p.hs = d->heading_scores; p.hrn = d->heading_res_norm; p.hr = d->heading_res;
    p.ss = d->size_scores; p.srn = d->size_res_norm; p.sr = d->size_res; p.box = d->box; p.cls = d->cls;
    if (bs <= 0) return 0;
    const size_t smem = (size_t)2 * kFcObjs * kFcMaxW * sizeof(float);
    AL3D_CHECK_CUDA(cudaFuncSetAttribute(fc_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fc_chain_kernel<<<(unsigned)ceil_div(bs, kFcObjs), kFcThreads, smem, (cudaStream_t)stream>>>(p);
    AL3D_CHECK_LAUNCH("fc_chain_kernel");
    return 0;
}
