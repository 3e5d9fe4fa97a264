// Device-side track preparation: the dataset-side merge / resample / canonical-frame transform that feeds the
// models.
//
// Reference: STATICTRACK.__getitem__ tools/static_model.py:529-572 (inverse-pose transform of the merged
// crops :541-543, resample with replacement :546-547, canonicalisation by the initial box :569-570) and
// DYNAMICTRACK.__getitem__ tools/dynamic_model.py:419-509 (5 x 1024 resampled points + time channel
// :429-439, 101-step box window + time channel :441-447, pose transform :452-453, canonicalisation by the
// centre box :503-507).  All arithmetic is float64 like the reference's numpy code; the result is rounded
// to float32 once at the end (the eval loops call .float() on it, tools/static_eval.py:265).
#include "common.cuh"
#include "../../include/al3d.h"

namespace al3d {

// out[b, j, :] = Rz(-h_b) * (P_b * [p; 1] - c_b)   with p = src[choice[b, j]] or 0 when choice < 0
__global__ void __launch_bounds__(256)
track_points_prep_kernel(const double *__restrict__ src, const int64_t *__restrict__ choice, int n_out,
                         const double *__restrict__ inv_pose, const double *__restrict__ init_box, int box_stride,
                         int heading_col, int c_out, int time_block, int time_center, float *__restrict__ out)
{
    const int b = blockIdx.y;
    const double *P = inv_pose + (int64_t)b * 16;
    const double cx = init_box[(int64_t)b * box_stride], cy = init_box[(int64_t)b * box_stride + 1], cz = init_box[(int64_t)b * box_stride + 2];
    const double h = -init_box[(int64_t)b * box_stride + heading_col];
    const double c = cos(h), s = sin(h);
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_out; j += gridDim.x * blockDim.x) {
        const int64_t r = choice[(int64_t)b * n_out + j];
        double x = 0.0, y = 0.0, z = 0.0;
        if (r >= 0) { x = src[r * 3]; y = src[r * 3 + 1]; z = src[r * 3 + 2]; }
        // homogeneous transform, row by row, left to right (numpy's 4x4 @ 4xN dot product order)
        double vx = ((P[0] * x + P[1] * y) + P[2] * z) + P[3];
        double vy = ((P[4] * x + P[5] * y) + P[6] * z) + P[7];
        double vz = ((P[8] * x + P[9] * y) + P[10] * z) + P[11];
        vx -= cx; vy -= cy; vz -= cz;
        float *o = out + ((int64_t)b * n_out + j) * c_out;
        o[0] = (float)(c * vx - s * vy);
        o[1] = (float)(s * vx + c * vy);
        o[2] = (float)vz;
        if (c_out == 4) o[3] = (float)(0.1 * (double)(j / time_block - time_center));
    }
}

// boxes (bs, steps, 8) f64 [x y z l w h heading dt] in the global frame -> vehicle frame of inv_pose, then
// relative to the centre step (tools/dynamic_model.py:452,506-507); init_box (bs, 8) = the centre step before
// the subtraction.
__global__ void boxseq_prep_kernel(const double *__restrict__ box, int steps, int center_step, const double *__restrict__ inv_pose,
                                   float *__restrict__ out, double *__restrict__ init_box)
{
    const int b = blockIdx.x;
    const double *P = inv_pose + (int64_t)b * 16;
    const double dh = atan2(P[4], P[0]);
    __shared__ double ctr[4];
    const double *bc = box + ((int64_t)b * steps + center_step) * 8;
    if (threadIdx.x == 0) {
        ctr[0] = ((P[0] * bc[0] + P[1] * bc[1]) + P[2] * bc[2]) + P[3];
        ctr[1] = ((P[4] * bc[0] + P[5] * bc[1]) + P[6] * bc[2]) + P[7];
        ctr[2] = ((P[8] * bc[0] + P[9] * bc[1]) + P[10] * bc[2]) + P[11];
        ctr[3] = bc[6] + dh;
        if (init_box) {
            double *ib = init_box + (int64_t)b * 8;
            ib[0] = ctr[0]; ib[1] = ctr[1]; ib[2] = ctr[2]; ib[3] = bc[3]; ib[4] = bc[4]; ib[5] = bc[5]; ib[6] = ctr[3]; ib[7] = bc[7];
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < steps; t += blockDim.x) {
        const double *bi = box + ((int64_t)b * steps + t) * 8;
        float *o = out + ((int64_t)b * steps + t) * 8;
        const double x = ((P[0] * bi[0] + P[1] * bi[1]) + P[2] * bi[2]) + P[3];
        const double y = ((P[4] * bi[0] + P[5] * bi[1]) + P[6] * bi[2]) + P[7];
        const double z = ((P[8] * bi[0] + P[9] * bi[1]) + P[10] * bi[2]) + P[11];
        o[0] = (float)(x - ctr[0]); o[1] = (float)(y - ctr[1]); o[2] = (float)(z - ctr[2]);
        o[3] = (float)bi[3]; o[4] = (float)bi[4]; o[5] = (float)bi[5];
        o[6] = (float)((bi[6] + dh) - ctr[3]);
        o[7] = (float)bi[7];
    }
}

}  // namespace al3d

using namespace al3d;

extern "C" int al3d_track_points_prep(const double *src_xyz, const int64_t *choice, int bs, int n_out, const double *inv_pose,
                                      const double *init_box, int box_stride, int heading_col, int c_out, int time_block,
                                      int time_center, float *out, void *stream)
{
    AL3D_CHECK_ARG(src_xyz && choice && inv_pose && init_box && out, "al3d_track_points_prep: null pointer");
    AL3D_CHECK_ARG(c_out == 3 || c_out == 4, "al3d_track_points_prep: c_out=%d must be 3 or 4", c_out);
    AL3D_CHECK_ARG(c_out == 3 || time_block > 0, "al3d_track_points_prep: time_block must be positive");
    AL3D_CHECK_ARG(box_stride > heading_col && heading_col >= 3, "al3d_track_points_prep: bad box layout");
    if (bs <= 0 || n_out <= 0) return 0;
    dim3 grid((unsigned)std::min<int64_t>(ceil_div(n_out, 256), 64), (unsigned)bs);
    track_points_prep_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src_xyz, choice, n_out, inv_pose, init_box, box_stride, heading_col,
                                                                     c_out, time_block > 0 ? time_block : 1, time_center, out);
    AL3D_CHECK_LAUNCH("track_points_prep_kernel");
    return 0;
}

extern "C" int al3d_boxseq_prep(const double *box, int bs, int steps, int center_step, const double *inv_pose, float *out,
                                double *init_box, void *stream)
{
    AL3D_CHECK_ARG(box && inv_pose && out, "al3d_boxseq_prep: null pointer");
    AL3D_CHECK_ARG(steps > 0 && center_step >= 0 && center_step < steps, "al3d_boxseq_prep: bad step counts");
    if (bs <= 0) return 0;
    boxseq_prep_kernel<<<bs, 128, 0, (cudaStream_t)stream>>>(box, steps, center_step, inv_pose, out, init_box);
    AL3D_CHECK_LAUNCH("boxseq_prep_kernel");
    return 0;
}
