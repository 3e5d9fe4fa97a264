/* Plain-C restatement of the reference's points-in-rotated-box inner loops.
 * TEST INFRASTRUCTURE (oracle) -- never linked into the product library.
 *
 *   al3d_ref_planes   <- surface_equ_3d_jitv2        det3d/core/bbox/geometry.py:351-377
 *   al3d_ref_inside   <- _points_in_convex_polygon_3d_jit   det3d/core/bbox/geometry.py:241-276
 *
 * The reference compiles these with numba: for float32 inputs the generated code is scalar
 * mulss/addss/subss with NO fused multiply-add, evaluated left to right.
 * Build with -ffp-contract=off (oracle/Makefile) so gcc keeps the same roundings.
 */
#include <stdint.h>
#include <stddef.h>

/* surfaces: (B,6,4,3) f32 quads whose normals point inward (only the first 3 corners are read).
 * planes out: (B,6,4) f32 = (nx, ny, nz, d). */
void al3d_ref_planes(const float *surfaces, int64_t n_boxes, float *planes)
{
    for (int64_t b = 0; b < n_boxes; ++b) {
        for (int j = 0; j < 6; ++j) {
            const float *s = surfaces + ((b * 6 + j) * 4) * 3;
            float a0 = s[0] - s[3], a1 = s[1] - s[4], a2 = s[2] - s[5];
            float b0 = s[3] - s[6], b1 = s[4] - s[7], b2 = s[5] - s[8];
            float nx = a1 * b2 - a2 * b1;
            float ny = a2 * b0 - a0 * b2;
            float nz = a0 * b1 - a1 * b0;
            float d = -s[0] * nx - s[1] * ny - s[2] * nz;
            float *o = planes + (b * 6 + j) * 4;
            o[0] = nx; o[1] = ny; o[2] = nz; o[3] = d;
        }
    }
}

/* points: (N, stride) f32 rows (xyz first); out: (N,B) uint8, 1 = inside all six planes. */
void al3d_ref_inside(const float *points, int64_t n_points, int64_t stride,
                     const float *planes, int64_t n_boxes, uint8_t *out)
{
    for (int64_t i = 0; i < n_points; ++i) {
        const float *p = points + i * stride;
        for (int64_t b = 0; b < n_boxes; ++b) {
            uint8_t in = 1;
            for (int k = 0; k < 6; ++k) {
                const float *q = planes + (b * 6 + k) * 4;
                float sign = p[0] * q[0] + p[1] * q[1] + p[2] * q[2] + q[3];
                if (sign >= 0) { in = 0; break; }
            }
            out[i * n_boxes + b] = in;
        }
    }
}

/* Same test for float64 points (the reference's f64 specialisation, used by the dataset label
 * path tools/static_model.py:556 where points are float64 and the box float32 -> planes f32).
 * numba promotes the f32 plane coefficients to f64 for the arithmetic. */
void al3d_ref_inside_f64(const double *points, int64_t n_points, int64_t stride,
                         const float *planes, int64_t n_boxes, uint8_t *out)
{
    for (int64_t i = 0; i < n_points; ++i) {
        const double *p = points + i * stride;
        for (int64_t b = 0; b < n_boxes; ++b) {
            uint8_t in = 1;
            for (int k = 0; k < 6; ++k) {
                const float *q = planes + (b * 6 + k) * 4;
                double sign = p[0] * (double)q[0] + p[1] * (double)q[1] + p[2] * (double)q[2] + (double)q[3];
                if (sign >= 0) { in = 0; break; }
            }
            out[i * n_boxes + b] = in;
        }
    }
}
