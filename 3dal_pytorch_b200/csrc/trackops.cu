// Glue around the models, on the device: integer / index work is exact.
//
//   track_regroup     per-frame detections -> per-track observation lists in first-appearance order
//                     (tools/trackData.py:25-45: a dict keyed by tracking id, lists appended in frame order)
//   motion_features   per-track [ ||b_first - b_last||, ||var(b)|| ] for the static / dynamic split
//                     (tools/motionState.py:30-67; the linear SVC itself is w.x + b)
//   track_labels      training labels of a track: per-point mask from the float64 points-in-GT-box test, centre,
//                     heading class / residual, size class / residual
//                     (tools/static_model.py:549-566, tools/utils.py:53-67)
//   box_writeback     refined box of a track -> every frame of the track (tools/static_eval.py:84-92)
#include "common.cuh"
#include "../../include/al3d.h"

namespace al3d {
namespace trackops {

constexpr long long kEmptyKey = (long long)0x8000000000000000ull;

__device__ __forceinline__ uint32_t hash64(long long k)
{
    unsigned long long x = (unsigned long long)k;
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return (uint32_t)x;
}

__global__ void fill_i64_kernel(long long *p, int64_t n, long long v)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void fill_i32_kernel(int *p, int64_t n, int v)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// open-addressing table: first observation index of every id (atomicMin), any insertion order
__global__ void regroup_insert_kernel(const long long *__restrict__ ids, int n_obs, long long *__restrict__ keys, int *__restrict__ first,
                                      uint32_t mask)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_obs) return;
    const long long id = ids[i];
    uint32_t h = hash64(id) & mask;
    for (;;) {
        const long long prev = (long long)atomicCAS(reinterpret_cast<unsigned long long *>(keys + h), (unsigned long long)kEmptyKey,
                                                    (unsigned long long)id);
        if (prev == kEmptyKey || prev == id) { atomicMin(first + h, i); return; }
        h = (h + 1) & mask;
    }
}

__device__ __forceinline__ int regroup_lookup(long long id, const long long *keys, uint32_t mask)
{
    uint32_t h = hash64(id) & mask;
    while (keys[h] != id) h = (h + 1) & mask;
    return (int)h;
}

// single CTA: rank of every first appearance (exclusive scan of the "is first" flags), track ids, per-observation track
__global__ void __launch_bounds__(1024)
regroup_rank_kernel(const long long *__restrict__ ids, int n_obs, const long long *__restrict__ keys, const int *__restrict__ first,
                    uint32_t mask, int *__restrict__ slot_rank, int *__restrict__ track_of_obs, long long *__restrict__ track_id,
                    int track_cap, int *__restrict__ n_tracks)
{
    __shared__ int warp_sum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < n_obs; base += 1024) {
        const int i = base + threadIdx.x;
        int slot = -1, flag = 0;
        if (i < n_obs) { slot = regroup_lookup(ids[i], keys, mask); flag = (first[slot] == i) ? 1 : 0; }
        int incl = flag;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) warp_sum[w] = incl;
        __syncthreads();
        int before = carry;
        for (int k = 0; k < w; ++k) before += warp_sum[k];
        if (flag) {
            const int r = before + incl - 1;
            slot_rank[slot] = r;
            if (r < track_cap) track_id[r] = ids[i];
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int k = 0; k < 32; ++k) t += warp_sum[k]; carry += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_tracks = carry;
    // every first appearance precedes (in observation order) all other observations of its id, but not necessarily in
    // this loop's chunk order for LATER ids: resolve the per-observation track in a second sweep
    __syncthreads();
    for (int i = threadIdx.x; i < n_obs; i += 1024) track_of_obs[i] = slot_rank[regroup_lookup(ids[i], keys, mask)];
}

// presence[t, f] = observation index of track t in frame f (the first one if an id repeats within a frame)
__global__ void regroup_presence_kernel(const int *__restrict__ track_of_obs, const int *__restrict__ frame_of_obs, int n_obs, int n_frames,
                                        int track_cap, int *__restrict__ presence, int *__restrict__ error)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_obs) return;
    const int t = track_of_obs[i], f = frame_of_obs[i];
    if (t >= track_cap || f < 0 || f >= n_frames) { atomicExch(error, 1); return; }
    const int prev = atomicMin(presence + (int64_t)t * n_frames + f, i);
    if (prev != 0x7fffffff) atomicExch(error, 2);                   // the same id twice in one frame
}

// one warp per track: compact its row of `presence` (frame order) into track_obs[t, 0:len], pad with -1
__global__ void regroup_compact_kernel(const int *__restrict__ presence, const int *__restrict__ n_tracks, int n_frames, int track_cap,
                                       int *__restrict__ track_obs, int *__restrict__ track_len)
{
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= track_cap) return;
    const bool live = t < *n_tracks;
    int count = 0;
    for (int f0 = 0; f0 < n_frames; f0 += 32) {
        const int f = f0 + lane;
        const int v = (live && f < n_frames) ? presence[(int64_t)t * n_frames + f] : 0x7fffffff;
        const bool has = v != 0x7fffffff;
        const unsigned bal = __ballot_sync(0xffffffffu, has);
        if (has) track_obs[(int64_t)t * n_frames + count + __popc(bal & ((1u << lane) - 1u))] = v;
        count += __popc(bal);
    }
    for (int k = count + lane; k < n_frames; k += 32) track_obs[(int64_t)t * n_frames + k] = -1;
    if (lane == 0) track_len[t] = live ? count : 0;
}

// features of one track over the n_cols leading box columns: ||b_first - b_last||, ||var over observations|| (ddof 0)
__global__ void motion_features_kernel(const int *__restrict__ track_obs, const int *__restrict__ track_len, int n_tracks, int n_frames,
                                       const double *__restrict__ boxes, int box_stride, int n_cols, double *__restrict__ feat)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tracks) return;
    const int L = track_len[t];
    if (L <= 0) { feat[t * 2] = 0.0; feat[t * 2 + 1] = 0.0; return; }
    const int *obs = track_obs + (int64_t)t * n_frames;
    double d2 = 0.0, v2 = 0.0;
    for (int c = 0; c < n_cols; ++c) {
        const double first = boxes[(int64_t)obs[0] * box_stride + c], last = boxes[(int64_t)obs[L - 1] * box_stride + c];
        d2 += (first - last) * (first - last);
        double s = 0.0;
        for (int k = 0; k < L; ++k) s += boxes[(int64_t)obs[k] * box_stride + c];
        const double mean = s / (double)L;
        double q = 0.0;
        for (int k = 0; k < L; ++k) { const double e = boxes[(int64_t)obs[k] * box_stride + c] - mean; q += e * e; }
        const double var = q / (double)L;
        v2 += var * var;
    }
    feat[t * 2] = sqrt(d2);
    feat[t * 2 + 1] = sqrt(v2);
}

__constant__ double c_mean_size_d[9] = {4.8, 1.8, 1.5, 10.0, 2.6, 3.2, 2.0, 1.0, 1.6};

// mask label of every resampled point: float64 points against the float32 plane equations of the GT box, evaluated in
// float64 like numba's float64 specialisation of _points_in_convex_polygon_3d_jit (geometry.py:241-276): no FMA.
__global__ void __launch_bounds__(256)
track_mask_label_kernel(const double *__restrict__ src, const int64_t *__restrict__ choice, int n_out, const double *__restrict__ inv_pose,
                        const float *__restrict__ gt_planes, float *__restrict__ mask_label)
{
    const int b = blockIdx.y;
    const double *P = inv_pose + (int64_t)b * 16;
    double pn[6][4];
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
        for (int c = 0; c < 4; ++c) pn[k][c] = (double)gt_planes[((int64_t)b * 6 + k) * 4 + c];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_out; j += gridDim.x * blockDim.x) {
        const int64_t r = choice[(int64_t)b * n_out + j];
        double x = 0.0, y = 0.0, z = 0.0;
        if (r >= 0) { x = src[r * 3]; y = src[r * 3 + 1]; z = src[r * 3 + 2]; }
        const double vx = ((P[0] * x + P[1] * y) + P[2] * z) + P[3];
        const double vy = ((P[4] * x + P[5] * y) + P[6] * z) + P[7];
        const double vz = ((P[8] * x + P[9] * y) + P[10] * z) + P[11];
        bool in = true;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const double sgn = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(vx, pn[k][0]), __dmul_rn(vy, pn[k][1])), __dmul_rn(vz, pn[k][2])), pn[k][3]);
            in = in && !(sgn >= 0.0);
        }
        mask_label[(int64_t)b * n_out + j] = in ? 1.f : 0.f;
    }
}

// per track: centre label, angle2class(gt_heading - init_heading, 12), size2class(gt_lwh)  (float64, tools/utils.py:53-67)
__global__ void track_box_label_kernel(const float *__restrict__ gt_box, const double *__restrict__ init_heading, int bs, int n_bins,
                                       float *__restrict__ center_label, int64_t *__restrict__ heading_cls, float *__restrict__ heading_res,
                                       int64_t *__restrict__ size_cls, float *__restrict__ size_res)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= bs) return;
    const float *g = gt_box + (int64_t)b * 7;
    center_label[b * 3] = g[0]; center_label[b * 3 + 1] = g[1]; center_label[b * 3 + 2] = g[2];
    {
        const double two_pi = 2.0 * 3.141592653589793;
        double a = fmod((double)g[6] - init_heading[b], two_pi);
        if (a < 0.0) a += two_pi;                                       // Python's % for a positive modulus
        if (a >= two_pi) a -= two_pi;
        const double per = two_pi / (double)n_bins;
        double sh = fmod(a + per / 2.0, two_pi);
        if (sh < 0.0) sh += two_pi;
        const int cls = (int)(sh / per);
        heading_cls[b] = cls;
        heading_res[b] = (float)(sh - ((double)cls * per + per / 2.0));
    }
    {
        int best = 0;
        double bd = INFINITY;
        for (int c = 0; c < 3; ++c) {
            double d2 = 0.0;
            for (int k = 0; k < 3; ++k) { const double e = (double)g[3 + k] - c_mean_size_d[c * 3 + k]; d2 += e * e; }
            const double d = sqrt(d2);
            if (d < bd) { bd = d; best = c; }                             // np.argmin: first minimum
        }
        size_cls[b] = best;
        for (int k = 0; k < 3; ++k) size_res[b * 3 + k] = (float)((double)g[3 + k] - c_mean_size_d[best * 3 + k]);
    }
}

// transform_box (tools/static_eval.py:30-45): heading += atan2(T[1,0], T[0,0]); centre = R c + t.  Two hops per
// observation: best frame's vehicle frame -> global -> the observation's frame.
__device__ __forceinline__ void transform_box_d(const double *T, double (&b)[7])
{
    const double x = b[0], y = b[1], z = b[2];
    b[0] = ((T[0] * x + T[1] * y) + T[2] * z) + T[3];
    b[1] = ((T[4] * x + T[5] * y) + T[6] * z) + T[7];
    b[2] = ((T[8] * x + T[9] * y) + T[10] * z) + T[11];
    b[6] = b[6] + atan2(T[4], T[0]);
}

__global__ void box_writeback_kernel(const float *__restrict__ final_box, const double *__restrict__ best_pose, const int *__restrict__ track_obs,
                                     const int *__restrict__ track_len, int n_tracks, int n_frames, const double *__restrict__ obs_inv_pose,
                                     double *__restrict__ out)
{
    const int t = blockIdx.x;
    if (t >= n_tracks) return;
    const int L = track_len[t];
    for (int k = threadIdx.x; k < L; k += blockDim.x) {
        const int o = track_obs[(int64_t)t * n_frames + k];
        double b[7];
#pragma unroll
        for (int c = 0; c < 7; ++c) b[c] = (double)final_box[(int64_t)t * 7 + c];
        transform_box_d(best_pose + (int64_t)t * 16, b);
        transform_box_d(obs_inv_pose + (int64_t)o * 16, b);
#pragma unroll
        for (int c = 0; c < 7; ++c) out[(int64_t)o * 7 + c] = b[c];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// det <-> GT matching by rotated 3-D IoU (det3d/datasets/waymo/waymo_common.py:173-188: boxes_iou3d_gpu of one detection
// against the frame's GT boxes, arg-max, threshold 0.75 applied by the caller).  IoU3D = BEV overlap x height overlap /
// (vol_a + vol_b - overlap) as in det3d/ops/iou3d_nms/iou3d_nms_utils.py:35-72.  The BEV overlap is computed by
// clipping rectangle A, expressed in B's local frame, against B's four sides (Sutherland-Hodgman) and the shoelace
// formula; float32 like the reference kernel.  Footprint convention as in the crop: length along the direction -heading.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int clip_axis(const float (*in)[2], int n, float (*out)[2], int axis, float bound, bool keep_less)
{
    int m = 0;
    for (int i = 0; i < n; ++i) {
        const float *p = in[i], *q = in[(i + 1) % n];
        const bool pin = keep_less ? p[axis] <= bound : p[axis] >= bound;
        const bool qin = keep_less ? q[axis] <= bound : q[axis] >= bound;
        if (pin) { out[m][0] = p[0]; out[m][1] = p[1]; ++m; }
        if (pin != qin) {
            const float t = (bound - p[axis]) / (q[axis] - p[axis]);
            out[m][axis] = bound;
            out[m][1 - axis] = p[1 - axis] + t * (q[1 - axis] - p[1 - axis]);
            ++m;
        }
    }
    return m;
}

__device__ float bev_overlap(const float *a, const float *b)
{
    // corners of A (local (+-l/2, +-w/2), world = centre + [[c, s], [-s, c]] local), then into B's local frame
    const float ca = cosf(a[6]), sa = sinf(a[6]), cb = cosf(b[6]), sb = sinf(b[6]);
    float poly[2][10][2];
    const float sx[4] = {0.5f, 0.5f, -0.5f, -0.5f}, sy[4] = {0.5f, -0.5f, -0.5f, 0.5f};
    for (int k = 0; k < 4; ++k) {
        const float lx = sx[k] * a[3], ly = sy[k] * a[4];
        const float wx = a[0] + lx * ca + ly * sa - b[0], wy = a[1] - lx * sa + ly * ca - b[1];
        poly[0][k][0] = wx * cb - wy * sb;                       // local = [[c, -s], [s, c]] world
        poly[0][k][1] = wx * sb + wy * cb;
    }
    int n = 4;
    n = clip_axis(poly[0], n, poly[1], 0, 0.5f * b[3], true);
    if (n < 3) return 0.f;
    n = clip_axis(poly[1], n, poly[0], 0, -0.5f * b[3], false);
    if (n < 3) return 0.f;
    n = clip_axis(poly[0], n, poly[1], 1, 0.5f * b[4], true);
    if (n < 3) return 0.f;
    n = clip_axis(poly[1], n, poly[0], 1, -0.5f * b[4], false);
    if (n < 3) return 0.f;
    float area = 0.f;
    for (int i = 0; i < n; ++i) {
        const float *p = poly[0][i], *q = poly[0][(i + 1) % n];
        area += p[0] * q[1] - q[0] * p[1];
    }
    return 0.5f * fabsf(area);
}

__global__ void match_iou3d_kernel(const float *__restrict__ det, const int32_t *__restrict__ det_frame, int64_t n_det,
                                   const float *__restrict__ gt, const int64_t *__restrict__ gt_off, int32_t *__restrict__ best_idx,
                                   float *__restrict__ best_iou)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_det) return;
    const float *a = det + i * 7;
    const int f = det_frame[i];
    const int64_t g0 = gt_off[f], g1 = gt_off[f + 1];
    const float a_max = a[2] + a[5] / 2.f, a_min = a[2] - a[5] / 2.f, vol_a = a[3] * a[4] * a[5];
    int best = -1;
    float bi = -1.f;
    for (int64_t j = g0; j < g1; ++j) {
        const float *b = gt + j * 7;
        const float oh = fmaxf(fminf(a_max, b[2] + b[5] / 2.f) - fmaxf(a_min, b[2] - b[5] / 2.f), 0.f);
        float iou = 0.f;
        if (oh > 0.f) {
            const float o3 = bev_overlap(a, b) * oh;
            iou = o3 / fmaxf(vol_a + b[3] * b[4] * b[5] - o3, 1e-6f);
        }
        if (iou > bi) { bi = iou; best = (int)(j - g0); }          // np.argmax: first maximum
    }
    best_idx[i] = best;
    best_iou[i] = best < 0 ? 0.f : bi;
}

}  // namespace trackops
}  // namespace al3d

using namespace al3d;
using namespace al3d::trackops;

extern "C" int al3d_track_regroup(const int64_t *ids, const int32_t *frame_of_obs, int n_obs, int n_frames, int track_cap,
                                  int64_t *hash_keys, int32_t *hash_first, int32_t *hash_rank, int hash_size, int32_t *presence,
                                  int32_t *track_of_obs, int64_t *track_id, int32_t *track_obs, int32_t *track_len,
                                  int32_t *n_tracks, int32_t *error, void *stream)
{
    AL3D_CHECK_ARG(n_obs >= 0 && n_frames >= 1 && track_cap >= 1, "al3d_track_regroup: bad sizes");
    AL3D_CHECK_ARG(n_tracks && error && track_len && track_obs && presence, "al3d_track_regroup: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    AL3D_CHECK_CUDA(cudaMemsetAsync(n_tracks, 0, sizeof(int32_t), st));
    AL3D_CHECK_CUDA(cudaMemsetAsync(error, 0, sizeof(int32_t), st));
    if (n_obs > 0) {
        AL3D_CHECK_ARG(ids && frame_of_obs && hash_keys && hash_first && hash_rank && track_of_obs && track_id,
                       "al3d_track_regroup: null pointer");
        AL3D_CHECK_ARG(hash_size >= 2 * n_obs && (hash_size & (hash_size - 1)) == 0,
                       "al3d_track_regroup: hash_size must be a power of two >= 2 * n_obs");
        fill_i64_kernel<<<(unsigned)ceil_div(hash_size, 256), 256, 0, st>>>(reinterpret_cast<long long *>(hash_keys), hash_size, kEmptyKey);
        fill_i32_kernel<<<(unsigned)ceil_div(hash_size, 256), 256, 0, st>>>(hash_first, hash_size, 0x7fffffff);
        regroup_insert_kernel<<<(unsigned)ceil_div(n_obs, 256), 256, 0, st>>>(reinterpret_cast<const long long *>(ids), n_obs,
                                                                             reinterpret_cast<long long *>(hash_keys), hash_first,
                                                                             (uint32_t)(hash_size - 1));
        AL3D_CHECK_LAUNCH("regroup_insert_kernel");
        regroup_rank_kernel<<<1, 1024, 0, st>>>(reinterpret_cast<const long long *>(ids), n_obs, reinterpret_cast<const long long *>(hash_keys),
                                                hash_first, (uint32_t)(hash_size - 1), hash_rank, track_of_obs,
                                                reinterpret_cast<long long *>(track_id), track_cap, n_tracks);
        AL3D_CHECK_LAUNCH("regroup_rank_kernel");
    }
    fill_i32_kernel<<<(unsigned)ceil_div((int64_t)track_cap * n_frames, 256), 256, 0, st>>>(presence, (int64_t)track_cap * n_frames, 0x7fffffff);
    if (n_obs > 0) {
        regroup_presence_kernel<<<(unsigned)ceil_div(n_obs, 256), 256, 0, st>>>(track_of_obs, frame_of_obs, n_obs, n_frames, track_cap, presence, error);
        AL3D_CHECK_LAUNCH("regroup_presence_kernel");
    }
    regroup_compact_kernel<<<(unsigned)ceil_div(track_cap, 8), 256, 0, st>>>(presence, n_tracks, n_frames, track_cap, track_obs, track_len);
    AL3D_CHECK_LAUNCH("regroup_compact_kernel");
    return 0;
}

extern "C" int al3d_motion_features(const int32_t *track_obs, const int32_t *track_len, int n_tracks, int n_frames, const double *boxes,
                                    int box_stride, int n_cols, double *feat, void *stream)
{
    AL3D_CHECK_ARG(n_tracks >= 0 && n_frames >= 1 && n_cols >= 1 && box_stride >= n_cols, "al3d_motion_features: bad sizes");
    if (n_tracks == 0) return 0;
    AL3D_CHECK_ARG(track_obs && track_len && boxes && feat, "al3d_motion_features: null pointer");
    motion_features_kernel<<<(unsigned)ceil_div(n_tracks, 128), 128, 0, (cudaStream_t)stream>>>(track_obs, track_len, n_tracks, n_frames, boxes,
                                                                                               box_stride, n_cols, feat);
    AL3D_CHECK_LAUNCH("motion_features_kernel");
    return 0;
}

extern "C" int al3d_track_labels(const double *src_xyz, const int64_t *choice, int bs, int n_out, const double *inv_pose,
                                 const float *gt_planes, const float *gt_box, const double *init_heading, float *mask_label,
                                 float *center_label, int64_t *heading_cls, float *heading_res, int64_t *size_cls, float *size_res,
                                 void *stream)
{
    AL3D_CHECK_ARG(bs >= 0 && n_out >= 0, "al3d_track_labels: bad sizes");
    if (bs == 0) return 0;
    AL3D_CHECK_ARG(gt_box && init_heading && center_label && heading_cls && heading_res && size_cls && size_res, "al3d_track_labels: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_out > 0 && mask_label) {
        AL3D_CHECK_ARG(src_xyz && choice && inv_pose && gt_planes, "al3d_track_labels: the mask label needs points, poses and planes");
        AL3D_CHECK_ARG(bs <= 65535, "al3d_track_labels: bs too large for one launch");
        track_mask_label_kernel<<<dim3((unsigned)std::min<int64_t>(ceil_div(n_out, 256), 64), (unsigned)bs), 256, 0, st>>>(
            src_xyz, choice, n_out, inv_pose, gt_planes, mask_label);
        AL3D_CHECK_LAUNCH("track_mask_label_kernel");
    }
    track_box_label_kernel<<<(unsigned)ceil_div(bs, 128), 128, 0, st>>>(gt_box, init_heading, bs, 12, center_label, heading_cls, heading_res,
                                                                        size_cls, size_res);
    AL3D_CHECK_LAUNCH("track_box_label_kernel");
    return 0;
}

extern "C" int al3d_box_writeback(const float *final_box, const double *best_pose, const int32_t *track_obs, const int32_t *track_len,
                                  int n_tracks, int n_frames, const double *obs_inv_pose, double *out_boxes, void *stream)
{
    AL3D_CHECK_ARG(n_tracks >= 0 && n_frames >= 1, "al3d_box_writeback: bad sizes");
    if (n_tracks == 0) return 0;
    AL3D_CHECK_ARG(final_box && best_pose && track_obs && track_len && obs_inv_pose && out_boxes, "al3d_box_writeback: null pointer");
    box_writeback_kernel<<<n_tracks, 64, 0, (cudaStream_t)stream>>>(final_box, best_pose, track_obs, track_len, n_tracks, n_frames, obs_inv_pose,
                                                                    out_boxes);
    AL3D_CHECK_LAUNCH("box_writeback_kernel");
    return 0;
}

extern "C" int al3d_match_iou3d(const float *det, const int32_t *det_frame, int64_t n_det, const float *gt, const int64_t *gt_off,
                                int32_t *best_idx, float *best_iou, void *stream)
{
    AL3D_CHECK_ARG(n_det >= 0, "al3d_match_iou3d: negative size");
    if (n_det == 0) return 0;
    AL3D_CHECK_ARG(det && det_frame && gt_off && best_idx && best_iou, "al3d_match_iou3d: null pointer");
    match_iou3d_kernel<<<(unsigned)ceil_div(n_det, 128), 128, 0, (cudaStream_t)stream>>>(det, det_frame, n_det, gt, gt_off, best_idx, best_iou);
    AL3D_CHECK_LAUNCH("match_iou3d_kernel");
    return 0;
}
