// Points-in-rotated-box crop for batches of lidar frames: exact reference predicate, ascending index
// lists per (frame, box), fused gather + float64 pose transform.  Integer / indexing work: bit-exact.
//
// Reference: box_np_ops.points_in_rbbox det3d/core/bbox/box_np_ops.py:641-647 ->
// _points_in_convex_polygon_3d_jit det3d/core/bbox/geometry.py:241-276 (a point is inside iff for all six
// planes ((px*nx + py*ny) + pz*nz) + d is NOT >= 0, evaluated in float32 without FMA), and the crop
// materialisation det3d/datasets/waymo/waymo_common.py:168-171 (gather in ascending point order, then
// pose(4x4 f64) @ [x y z 1]).  The plane equations are computed on the device (crop_box_setup_kernel) in numpy's own
// float32 operation order; only sin / cos of the headings come from the host.
//
// The reference tests every point against every box (N*B*6 plane evaluations per frame).  Here a fine BEV occupancy
// bitmap per frame discards the points that are near no box, a coarse BEV grid maps a survivor to the few boxes whose
// (padded) bounding rectangle covers its cell, and a (point, box) pair evaluates the six plane equations only when a
// classification in the box's own frame cannot decide it with a rounding margin to spare -- same result, bit for bit.
//
// Pipeline (all launches cover the whole batch of frames; frames are read where they lie, through a table of pointers):
//   crop_box_setup / crop_box_local   per box: plane equations + padded rectangle, box-frame record
//   crop_grid_kernel   kGridBands CTAs per frame: grid extent from the rectangles, CSR cell -> box lists + packed cell
//                      entries, fine occupancy bitmap
//   crop_hits_kernel   one CTA per chunk of <= 16384 points: filter -> expand -> test in dense warp batches, ordered
//                      (x y z index | box, rank) hit records per chunk and per-(chunk, box) counts
//   crop_scan_kernel   per frame: exclusive scan over chunks for every box; crop_offsets_kernel: global offsets
//   crop_fill_kernel   scatters each hit's slot to its final, ascending position; crop_materialise_kernel turns the slots
//                      into point indices, coordinates and float64 global coordinates, front to back
#include "common.cuh"
#include "../../include/al3d.h"

namespace al3d {

constexpr int kCropChunk = 32768;      // most points one CTA of the hits kernel takes (the caller may cut smaller chunks)
constexpr int kCropThreads = 256;
// measured on the 200-frame sweep (bands x threads): 1 x 256 0.067 ms, 1 x 1024 0.050, 4 x 256 0.067, 2 x 512 0.044, 8 x 128 0.145
#ifndef CROP_GRID_THREADS
#define CROP_GRID_THREADS 512
#endif
#ifndef CROP_GRID_BANDS
#define CROP_GRID_BANDS 2
#endif
constexpr int kGridThreads = CROP_GRID_THREADS;
constexpr int kGridBands = CROP_GRID_BANDS;          // crop_grid_kernel: CTAs per frame, each owns a band of rows of both grids
constexpr int kOccRes = 256;           // fine occupancy bitmap per frame: kOccRes x kOccRes bits (8 KB)
constexpr int kOccWords = kOccRes * kOccRes / 32;

// per frame: BEV grid origin / scale of the coarse cell lists (x0, y0, inv_x, inv_y), scale of the fine occupancy
// bitmap (inv_fx, inv_fy) and the z range covered by any (padded) box
struct CropGridMeta { float x0, y0, inv_x, inv_y, inv_fx, inv_fy, zmin, zmax; };

// cell index of a coordinate; the SAME expression is used for points and for box rectangles, and it is
// monotone in v, so a point inside a rectangle always lands in a cell the rectangle was registered in.
__device__ __forceinline__ int crop_cell(float v, float v0, float inv, int G)
{
#ifdef CROP_CELL_FLOOR
    const float c = floorf((v - v0) * inv);
    return c < 0.f ? -1 : (c >= (float)G ? G : (int)c);
#else
    // float -> int with round-down saturates at the int range; clamp to [-1, G] ("left of / right of the grid")
    return max(-1, min(__float2int_rd((v - v0) * inv), G));
#endif
}

__global__ void __launch_bounds__(kGridThreads)
crop_grid_kernel(const float *__restrict__ aabb, const float *__restrict__ boxes, const float *__restrict__ sincos,
                 const int64_t *__restrict__ box_off, int G, CropGridMeta *__restrict__ meta, int32_t *__restrict__ cell_start,
                 int32_t *__restrict__ cell_boxes, int cell_cap, uint2 *__restrict__ cell4, int packed32, uint32_t *__restrict__ occ,
                 int32_t *__restrict__ overflow)
{
    // A frame is built by kGridBands CTAs.  Each of them repeats the cheap, global parts (extent of the boxes, per-cell counts,
    // exclusive scan: ~12 shared-memory atomics per box and 4096 cells) and does the other parts -- rasterising the fine
    // bitmap, filling / sorting / packing the cell lists -- only for its own band of rows, so that 200 frames make 400 CTAs
    // that are all resident at once instead of 200 long ones on 148 SMs (two waves).
    extern __shared__ int32_t s_cnt[];            // G*G counts -> exclusive offsets; then G*G fill cursors; then the band of the occupancy bitmap
    const int band = blockIdx.y;
    const int fr0 = band * (kOccRes / kGridBands), fr1 = fr0 + kOccRes / kGridBands - 1;          // fine rows of this band
    const int gpb = (G + kGridBands - 1) / kGridBands;
    const int gr0 = band * gpb, gr1 = min(G, gr0 + gpb) - 1;                                      // coarse rows of this band
    constexpr int kBandWords = kOccWords / kGridBands;
    __shared__ float red[6][kGridThreads / 32];
    __shared__ int32_t s_warp[kGridThreads / 32];
    const int f = blockIdx.x;
    const int64_t b0 = box_off[f];
    const int B = (int)(box_off[f + 1] - b0);
    const int cells = G * G;
    int32_t *s_cur = s_cnt + cells;
    // ---- extent of the boxes of this frame
    float xmin = INFINITY, ymin = INFINITY, xmax = -INFINITY, ymax = -INFINITY, zmin = INFINITY, zmax = -INFINITY;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float *a = aabb + (b0 + b) * 6;
        xmin = fminf(xmin, a[0]); ymin = fminf(ymin, a[1]); zmin = fminf(zmin, a[2]);
        xmax = fmaxf(xmax, a[3]); ymax = fmaxf(ymax, a[4]); zmax = fmaxf(zmax, a[5]);
    }
    for (int o = 16; o > 0; o >>= 1) {
        xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, o)); ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o)); ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
        zmin = fminf(zmin, __shfl_xor_sync(0xffffffffu, zmin, o)); zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { red[0][wid] = xmin; red[1][wid] = ymin; red[2][wid] = xmax; red[3][wid] = ymax; red[4][wid] = zmin; red[5][wid] = zmax; }
    uint32_t *s_occ = reinterpret_cast<uint32_t *>(s_cnt + 2 * cells);
    for (int c = threadIdx.x; c < 2 * cells + kBandWords; c += blockDim.x) s_cnt[c] = 0;
    __syncthreads();
    for (int w = 0; w < kGridThreads / 32; ++w) {
        xmin = fminf(xmin, red[0][w]); ymin = fminf(ymin, red[1][w]);
        xmax = fmaxf(xmax, red[2][w]); ymax = fmaxf(ymax, red[3][w]);
        zmin = fminf(zmin, red[4][w]); zmax = fmaxf(zmax, red[5][w]);
    }
    CropGridMeta m;
    float fsx = 0.f, fsy = 0.f;                                       // fine cell size
    if (B == 0 || !(xmax >= xmin) || !(ymax >= ymin)) {
        m.x0 = 0.f; m.y0 = 0.f; m.inv_x = 0.f; m.inv_y = 0.f; m.inv_fx = 0.f; m.inv_fy = 0.f; m.zmin = INFINITY; m.zmax = -INFINITY;
    } else {
        const float ex = fmaxf(xmax - xmin, 1e-3f), ey = fmaxf(ymax - ymin, 1e-3f);
        m.x0 = xmin; m.y0 = ymin;
        m.inv_x = (float)G / ex * 0.999f;                            // keep xmax inside the last cell
        m.inv_y = (float)G / ey * 0.999f;
        m.inv_fx = (float)kOccRes / ex * 0.999f;
        m.inv_fy = (float)kOccRes / ey * 0.999f;
        m.zmin = zmin; m.zmax = zmax;
        fsx = 1.f / m.inv_fx; fsy = 1.f / m.inv_fy;
    }
    if (threadIdx.x == 0 && band == 0) meta[f] = m;
    // ---- fine occupancy bitmap: a bit is set if the cell can contain a point of some box.  Conservative: the cell
    //      centre lies inside the box footprint grown by the cell's half diagonal plus the rectangle padding (which
    //      already exceeds the rounding slack of the exact test by orders of magnitude) plus 1 cm.
    if (occ != nullptr && m.inv_fx > 0.f) {
        // one warp per box, the lanes over the fine cells of its rectangle
        const float hd = 0.5f * sqrtf(fsx * fsx + fsy * fsy);
        for (int b = wid; b < B; b += kGridThreads / 32) {
            const float *a = aabb + (b0 + b) * 6;
            const float *bx = boxes + (b0 + b) * 7;
            const float sn = sincos[(b0 + b) * 2], cs = sincos[(b0 + b) * 2 + 1];
            const float grow = hd + 0.5f * ((a[3] - a[0]) - (fabsf(bx[3] * cs) + fabsf(bx[4] * sn))) + 0.01f;   // half diagonal + rectangle pad + 1 cm
            // |l|, |w|: two negative dimensions mirror the box onto itself, and the reference then still finds points inside
            const float hl = 0.5f * fabsf(bx[3]) + fmaxf(grow, hd + 0.06f), hw = 0.5f * fabsf(bx[4]) + fmaxf(grow, hd + 0.06f);
            const int cx0 = max(crop_cell(a[0], m.x0, m.inv_fx, kOccRes), 0), cx1 = min(crop_cell(a[3], m.x0, m.inv_fx, kOccRes), kOccRes - 1);
            const int cy0 = max(crop_cell(a[1], m.y0, m.inv_fy, kOccRes), fr0), cy1 = min(crop_cell(a[4], m.y0, m.inv_fy, kOccRes), fr1);
            const int ncx = cx1 - cx0 + 1, n = ncx > 0 && cy1 >= cy0 ? ncx * (cy1 - cy0 + 1) : 0;
            for (int t = lane; t < n; t += 32) {
                const int cy = cy0 + t / ncx, cx = cx0 + t % ncx;
                const float dx = m.x0 + ((float)cx + 0.5f) * fsx - bx[0], dy = m.y0 + ((float)cy + 0.5f) * fsy - bx[1];
                // world = [[c, s], [-s, c]] local  (rotation_3d_in_axis)  ->  local = [[c, -s], [s, c]] world
                const float lx = dx * cs - dy * sn, ly = dx * sn + dy * cs;
                if (fabsf(lx) <= hl && fabsf(ly) <= hw) atomicOr(&s_occ[((cy - fr0) * kOccRes + cx) >> 5], 1u << (cx & 31));
            }
        }
    }
    // ---- pass 0: every box adds itself to the counters of the cells its rectangle covers (integer adds commute);
    //      pass 1: it appends its id through a per-cell cursor; the short lists are sorted afterwards, so the
    //      final cell -> box lists are in ascending box order whatever the order of the atomics.
    for (int pass = 0; pass < 2; ++pass) {
        for (int b = threadIdx.x; b < B; b += blockDim.x) {
            const float *a = aabb + (b0 + b) * 6;
            int x0c = crop_cell(a[0], m.x0, m.inv_x, G), x1c = crop_cell(a[3], m.x0, m.inv_x, G);
            int y0c = crop_cell(a[1], m.y0, m.inv_y, G), y1c = crop_cell(a[4], m.y0, m.inv_y, G);
            x0c = max(x0c, 0); y0c = max(y0c, 0); x1c = min(x1c, G - 1); y1c = min(y1c, G - 1);
            if (pass == 1) { y0c = max(y0c, gr0); y1c = min(y1c, gr1); }      // lists: only the rows of this band
            for (int cy = y0c; cy <= y1c; ++cy)
                for (int cx = x0c; cx <= x1c; ++cx) {
                    const int c = cy * G + cx;
                    if (pass == 0) atomicAdd(&s_cnt[c], 1);
                    else {
                        const int at = s_cnt[c] + atomicAdd(&s_cur[c], 1);
                        if (at < cell_cap) cell_boxes[(int64_t)f * cell_cap + at] = b;
                    }
                }
        }
        __syncthreads();
        if (pass == 0) {
            // block-wide exclusive scan of the per-cell counts: contiguous strips per thread
            const int per = (cells + kGridThreads - 1) / kGridThreads;
            const int lo = min(threadIdx.x * per, cells), hi = min(lo + per, cells);
            int sum = 0;
            for (int c = lo; c < hi; ++c) sum += s_cnt[c];
            int incl = sum;
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            if (lane == 31) s_warp[wid] = incl;
            __syncthreads();
            int base = incl - sum;
            for (int w = 0; w < wid; ++w) base += s_warp[w];
            for (int c = lo; c < hi; ++c) { const int v = s_cnt[c]; s_cnt[c] = base; base += v; }
            __syncthreads();
            int total = 0;
            for (int w = 0; w < kGridThreads / 32; ++w) total += s_warp[w];
            if (threadIdx.x == 0 && band == 0) {
                cell_start[(int64_t)f * (cells + 1) + cells] = total;
                if (total > cell_cap) atomicExch(overflow, 1);
            }
            if (gr0 <= gr1)
                for (int c = gr0 * G + threadIdx.x; c < (gr1 + 1) * G; c += blockDim.x) cell_start[(int64_t)f * (cells + 1) + c] = s_cnt[c];
        }
    }
    if (occ != nullptr) for (int w = threadIdx.x; w < kBandWords; w += blockDim.x) occ[(int64_t)f * kOccWords + band * kBandWords + w] = s_occ[w];
    // ---- sort every cell's list (a handful of entries): insertion sort by one thread per cell
    __threadfence_block();
    for (int c = gr0 * G + threadIdx.x; c < (gr1 + 1) * G; c += blockDim.x) {
        const int lo = s_cnt[c], n = min(s_cur[c], max(cell_cap - lo, 0));
        int32_t *l = cell_boxes + (int64_t)f * cell_cap + lo;
        if (n <= 8) {
            // the usual case, in registers: one round trip to read the list and one to write it (sorting it in place in
            // global memory is a chain of dependent L2 accesses -- it was most of this kernel's time)
            int v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = k < n ? l[k] : 0x7fffffff;
#pragma unroll
            for (int i = 1; i < 8; ++i)
#pragma unroll
                for (int j = i; j > 0; --j) {
                    const int a = min(v[j - 1], v[j]), b = max(v[j - 1], v[j]);
                    v[j - 1] = a; v[j] = b;
                }
#pragma unroll
            for (int k = 0; k < 8; ++k) if (k < n) l[k] = v[k];
        } else {
            for (int i = 1; i < n; ++i) {
                const int v = l[i];
                int j = i - 1;
                while (j >= 0 && l[j] > v) { l[j + 1] = l[j]; --j; }
                l[j + 1] = v;
            }
        }
        // packed entry of the cell for the hits kernel: up to three box ids in one 8-byte word (id0 | id1 << 16,
        // id2 | count << 16); count = 0xFFFF sends the reader to the CSR list (longer or truncated lists)
        // (packed32, for frames of at most 1023 boxes: one 4-byte word id0 | id1 << 10 | id2 << 20 | count << 30, all ones =
        // use the CSR list -- small enough for the hits kernel to keep a frame's table in shared memory)
        if (cell4 != nullptr) {
            const int full = s_cur[c];
            const bool ok = full <= 3 && full == n;
            const unsigned i0 = ok && n > 0 ? (unsigned)l[0] : 0u, i1 = ok && n > 1 ? (unsigned)l[1] : 0u, i2 = ok && n > 2 ? (unsigned)l[2] : 0u;
            if (packed32)
                reinterpret_cast<uint32_t *>(cell4)[(int64_t)f * cells + c] = ok ? (i0 | (i1 << 10) | (i2 << 20) | ((unsigned)n << 30)) : 0xFFFFFFFFu;
            else
                cell4[(int64_t)f * cells + c] = ok ? make_uint2(i0 | (i1 << 16), i2 | ((unsigned)n << 16)) : make_uint2(0u, 0xFFFFu << 16);
        }
    }
}

// Box -> plane equations + padded bounding rectangle, on the device, in the reference's own float32 operation order
// (corners_nd / rotation_3d_in_axis / center_to_corner_box3d det3d/core/bbox/box_np_ops.py:55-85,146-179,241-262;
// corner_to_surfaces_3d :650-670; surface_equ_3d_jitv2 det3d/core/bbox/geometry.py:351-377).  Only sin / cos of the
// heading come from the host (numpy's float32 sin / cos are not bit-identical to CUDA's sinf / cosf): everything else is
// explicit round-to-nearest multiplies / adds without contraction, so the planes equal numpy's bit for bit.
__constant__ float c_corner_sign[8][3] = {{-.5f, -.5f, -.5f}, {-.5f, -.5f, .5f}, {-.5f, .5f, .5f}, {-.5f, .5f, -.5f},
                                          {.5f, -.5f, -.5f},  {.5f, -.5f, .5f},  {.5f, .5f, .5f},  {.5f, .5f, -.5f}};
__constant__ int c_quad[6][4] = {{0, 1, 2, 3}, {7, 6, 5, 4}, {0, 3, 7, 4}, {1, 5, 6, 2}, {0, 4, 5, 1}, {3, 2, 6, 7}};

__global__ void crop_box_setup_kernel(const float *__restrict__ boxes, const float *__restrict__ sincos, int64_t n_boxes, float pad_abs,
                                      float pad_rel, float *__restrict__ planes, float *__restrict__ aabb)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_boxes) return;
    const float *bx = boxes + b * 7;
    const float s = sincos[b * 2], c = sincos[b * 2 + 1], ns = -s;
    float cx[8], cy[8], cz[8];
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float x = __fmul_rn(bx[3], c_corner_sign[k][0]), y = __fmul_rn(bx[4], c_corner_sign[k][1]), z = __fmul_rn(bx[5], c_corner_sign[k][2]);
        // einsum("aij,jka->aik") with rot_mat_T = [[c,-s,0],[s,c,0],[0,0,1]]: three-term sums, left to right
        const float rx = __fadd_rn(__fadd_rn(__fmul_rn(x, c), __fmul_rn(y, s)), __fmul_rn(z, 0.f));
        const float ry = __fadd_rn(__fadd_rn(__fmul_rn(x, ns), __fmul_rn(y, c)), __fmul_rn(z, 0.f));
        const float rz = __fadd_rn(__fadd_rn(__fmul_rn(x, 0.f), __fmul_rn(y, 0.f)), __fmul_rn(z, 1.f));
        cx[k] = __fadd_rn(rx, bx[0]); cy[k] = __fadd_rn(ry, bx[1]); cz[k] = __fadd_rn(rz, bx[2]);
        mn[0] = fminf(mn[0], cx[k]); mn[1] = fminf(mn[1], cy[k]); mn[2] = fminf(mn[2], cz[k]);
        mx[0] = fmaxf(mx[0], cx[k]); mx[1] = fmaxf(mx[1], cy[k]); mx[2] = fmaxf(mx[2], cz[k]);
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const int i0 = c_quad[q][0], i1 = c_quad[q][1], i2 = c_quad[q][2];
        const float ax = __fsub_rn(cx[i0], cx[i1]), ay = __fsub_rn(cy[i0], cy[i1]), az = __fsub_rn(cz[i0], cz[i1]);
        const float bxv = __fsub_rn(cx[i1], cx[i2]), byv = __fsub_rn(cy[i1], cy[i2]), bzv = __fsub_rn(cz[i1], cz[i2]);
        const float nx = __fsub_rn(__fmul_rn(ay, bzv), __fmul_rn(az, byv));
        const float ny = __fsub_rn(__fmul_rn(az, bxv), __fmul_rn(ax, bzv));
        const float nz = __fsub_rn(__fmul_rn(ax, byv), __fmul_rn(ay, bxv));
        // d = -q0.x * nx - q0.y * ny - q0.z * nz, left to right
        const float d = __fsub_rn(__fsub_rn(__fmul_rn(-cx[i0], nx), __fmul_rn(cy[i0], ny)), __fmul_rn(cz[i0], nz));
        reinterpret_cast<float4 *>(planes)[b * 6 + q] = make_float4(nx, ny, nz, d);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float pad = __fadd_rn(pad_abs, __fmul_rn(pad_rel, fmaxf(fabsf(mn[k]), fabsf(mx[k]))));
        aabb[b * 6 + k] = __fsub_rn(mn[k], pad);
        aabb[b * 6 + 3 + k] = __fadd_rn(mx[k], pad);
    }
}

// exact reference predicate.  The reference rejects at the first plane with sign >= 0; "inside" is therefore the
// AND over the six planes of !(sign >= 0) (a NaN sign never rejects), which lets all six plane equations be
// loaded up front (one memory round trip instead of six dependent ones) without changing any result.
__device__ __forceinline__ bool crop_inside(float px, float py, float pz, const float4 *__restrict__ pl)
{
    float4 q[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) q[k] = __ldg(pl + k);
    bool in = true;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const float sgn = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(px, q[k].x), __fmul_rn(py, q[k].y)), __fmul_rn(pz, q[k].z)), q[k].w);
        in = in && !(sgn >= 0.f);
    }
    return in;
}

struct CropChunk { int32_t frame; int32_t first_pt; int32_t n_pts; int32_t chunk_in_frame; };

// one hit = a 16-byte record (x, y, z, point index in its frame) + a 4-byte word box | rank << 16 (rank among the chunk's
// hits of that box), kept in two arrays over the same slots: the ranking only touches the 4-byte words, and the
// materialise pass reads the coordinates from the records (60 MB, written moments earlier) instead of gathering them
// from the point cloud again -- a gather of 10 % of the points touches nearly every DRAM line of it (measured: 450 MB).
constexpr int kCropHitBytes = 20;
constexpr int kCropWarps = kCropThreads / 32;
constexpr int kCropIter = 128;                            // points per warp iteration: point base + j * 32 + lane, j < 4
#ifndef CROP_PER_CHECK
#define CROP_PER_CHECK 2                                  // measured (hits pass): a look at the queue every 32 points 0.306 ms, 64: 0.231, 128: 0.229
#endif
constexpr int kPerCheck = CROP_PER_CHECK;                 // filters (of 32 points) between two looks at the candidate queue
constexpr int kCropCQ = 32 + 32 * kPerCheck;              // per-warp candidate queue: < 32 carried over + <= 32 kPerCheck new
constexpr int kLocSmemBoxes = 256;                        // frames with at most this many boxes keep their box records in shared memory
// The packed 4-byte cell table of such frames can be staged in shared memory too.  Measured on the 200-frame sweep
// (profiles/r2_crop_v3_ablation.txt): it takes 0.026 ms off the expand stage, but the 16 KB cost the fourth resident
// CTA per SM and the streaming filter then loses 0.04 ms -- off.
#ifndef CROP_C4_SMEM
#define CROP_C4_SMEM 0
#endif
constexpr bool kC4Smem = CROP_C4_SMEM != 0;
constexpr int kCropPQ = 128;                              // per-warp (point, box) pair ring: < 32 left over + <= 96 new
constexpr int kCellIds = 3;                               // box ids held by one packed coarse-cell entry
constexpr int kIdxBits = 15;                              // a pair's tag: chunk-relative point index | box << 15 | weird << 30 (box < 2^14)
static_assert(kCropChunk <= (1 << kIdxBits), "a pair carries the chunk-relative point index in kIdxBits bits");
static_assert(31 + 32 * kCellIds <= kCropPQ, "pair ring");

// Box in its own frame, for the conservative classification of a (point, box) pair (crop_box_local_kernel):
//   a = (cx, cy, cz, cos)   b = (sin, l/2, w/2, h/2)   c.x = margin m (NaN: always run the exact predicate)
// With (lx, ly) = R^T (p - c) and t = max(|lx| - l/2, |ly| - w/2, |dz| - h/2):  t < -m  =>  the reference predicate is
// true, t > m  =>  it is false; only pairs with |t| <= m (a millimetre-thin shell around the faces, or an irregular box)
// evaluate the six float32 plane equations.  m bounds the rounding of BOTH computations with a wide factor:
//   reference sign / |n|: <= ~8 eps (|p| + R) k,  k = 1 + D / min(l, w, h)  (corner rounding eps R turned into an
//   angular error of the normal eps R / min_dim, acting over the box extent D);  local coordinates: <= ~4 eps (t + D);
//   m = 32 eps (R + D) k  with R = |cx| + |cy| + |cz| + D, D = l + w + h, eps = 2^-23.
struct CropBoxLocal { float4 a, b, c; };
static_assert(sizeof(CropBoxLocal) == 48, "CropBoxLocal");

__global__ void crop_box_local_kernel(const float *__restrict__ boxes, const float *__restrict__ sincos, int64_t n_boxes,
                                      CropBoxLocal *__restrict__ local)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_boxes) return;
    const float *bx = boxes + b * 7;
    const float l = bx[3], w = bx[4], h = bx[5];
    const float D = l + w + h, mind = fminf(l, fminf(w, h));
    const float R = fabsf(bx[0]) + fabsf(bx[1]) + fabsf(bx[2]) + D;
    float m = 32.f * 1.1920929e-7f * (R + D) * (1.f + D / mind);
    // irregular: a non-positive or non-finite dimension, or a margin that is not small against the box
    if (!(l > 0.f) || !(w > 0.f) || !(h > 0.f) || !(m < 0.25f * mind) || !(R < 1e18f)) m = __int_as_float(0x7fc00000);
    CropBoxLocal o;
    o.a = make_float4(bx[0], bx[1], bx[2], sincos[b * 2 + 1]);
    o.b = make_float4(sincos[b * 2], 0.5f * l, 0.5f * w, 0.5f * h);
    o.c = make_float4(m, 0.f, 0.f, 0.f);
    local[b] = o;
}

// TMA staging of the points (one 1536-byte bulk copy per warp iteration, completion on an mbarrier of the warp's own)
__device__ __forceinline__ uint32_t crop_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void crop_mbar_init(uint64_t *bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(crop_smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void crop_bulk_load(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the generic-proxy reads of the buffer come before the copy
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(crop_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(crop_smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(crop_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void crop_mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(crop_smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

// NaN / infinite (or absurdly large) coordinate: the reference's float32 arithmetic decides (a NaN sign never rejects, so
// a NaN point is inside every box; inf * 0 products do the same for some boxes) -> exact predicate against EVERY box
__device__ __forceinline__ bool crop_weird(float x, float y, float z) { return !(fabsf(x) + fabsf(y) + fabsf(z) < 1e18f); }

#ifndef CROP_MINB
#define CROP_MINB 3
#endif
#ifdef CROP_STATS
// diagnostic build only (scripts/gpu_r2_w.sh): [0] candidates, [1] pairs, [2] expand batches, [3] serial expand batches,
// [4] test batches, [5] test batches that ran the exact predicate, [6] pairs inside the margin, [7] hits
__device__ unsigned long long g_crop_stats[8];
#define CROP_STAT(i, v) do { if (lane == 0) atomicAdd(&g_crop_stats[i], (unsigned long long)(v)); } while (0)
#else
#define CROP_STAT(i, v) do { } while (0)
#endif
template <bool kSmemTables>
__global__ void __launch_bounds__(kCropThreads, CROP_MINB)
crop_hits_kernel(const float *const *__restrict__ frame_points, int64_t pt_stride,
                 const float *__restrict__ planes, const CropBoxLocal *__restrict__ local, const int64_t *__restrict__ box_off, int G,
                 const CropGridMeta *__restrict__ meta, const int32_t *__restrict__ cell_start,
                 const int32_t *__restrict__ cell_boxes, int cell_cap, const uint2 *__restrict__ cell4, const uint32_t *__restrict__ occ,
                 const CropChunk *__restrict__ chunks, float4 *__restrict__ hits_xyz, int32_t *__restrict__ hits_br, int hit_cap,
                 int32_t *__restrict__ n_hits,
                 int32_t *__restrict__ chunk_box_count, int max_boxes, int rank_boxes, int32_t *__restrict__ overflow)
{
    // Each warp owns a run of CONSECUTIVE points of the chunk and streams it 128 points at a time (point = base + j * 32 +
    // lane: lane order is point order, so a ballot compacts in order).  Three stages, each run on DENSE batches (every
    // lane busy), connected by per-warp queues in shared memory that keep point order:
    //   1  filter: z range of the frame's boxes, then one bit of the frame's fine occupancy bitmap (staged in shared
    //      memory).  ~80 % of the points end here; survivors go to the candidate queue.
    //   2  expand: one 8-byte load gives the (up to three) boxes registered in the candidate's coarse BEV cell; every
    //      (point, box) pair goes to the pair ring.  A lane whose cell holds more boxes reads the CSR list; NaN / inf points
    //      (tested against every box, like the reference does) take a serial form.
    //   3  test: classification of the pair in the box's own frame with a rounding margin; only pairs inside the margin
    //      evaluate the exact six-plane float32 predicate.  Hits are appended, in point order, to the warp's segment.
    // No block barrier inside the point loop; no limit on the number of boxes a point is in.  The hit list of a chunk is
    // kept as eight per-warp segments of hit_cap entries in global memory (L2-resident while the chunk is processed):
    // hits[(chunk * 8 + warp) * hit_cap + i], n_hits[chunk * 8 + warp].
    extern __shared__ float4 s_dyn4[];
    float4 *s_cq = s_dyn4;                                             // kCropWarps x kCropCQ
    float4 *s_pq = s_cq + kCropWarps * kCropCQ;                        // kCropWarps x kCropPQ
    float *s_pt = reinterpret_cast<float *>(s_pq + kCropWarps * kCropPQ);    // kCropWarps x 384 floats: the points of one warp iteration (TMA)
    float4 *s_loc = reinterpret_cast<float4 *>(s_pt + kCropWarps * kCropIter * 3);   // kSmemTables: the frame's box records (3 x 16 B each)
    __shared__ __align__(8) uint64_t s_bar[kCropWarps];
    uint32_t *s_c4 = reinterpret_cast<uint32_t *>(s_loc + (kSmemTables ? 3 * max_boxes : 0));   // kSmemTables && kC4Smem: packed cell entries (G x G)
    uint32_t *s_occ = s_c4 + (kSmemTables && kC4Smem ? G * G : 0);    // kOccWords
    int32_t *s_box_cnt = reinterpret_cast<int32_t *>(s_occ + kOccWords);           // per-box hit counters (max_boxes)
    int32_t *s_wcnt = s_box_cnt + ((max_boxes + 3) & ~3);              // kCropWarps x rank_boxes per-warp box counts
    __shared__ int warp_total[kCropWarps];
    const int stage_cap = hit_cap;
    const CropChunk ck = chunks[blockIdx.x];
    const int f = ck.frame;
    const int64_t b0 = box_off[f];
    const int B = (int)(box_off[f + 1] - b0);
    const CropGridMeta m = meta[f];
    const int cells = G * G;
    const int32_t *cs = cell_start + (int64_t)f * (cells + 1);
    const int32_t *cb = cell_boxes + (int64_t)f * cell_cap;
    const uint2 *c4 = cell4 + (int64_t)f * cells;
    const float *pts = frame_points[f] + (int64_t)ck.first_pt * pt_stride;
    const float4 *pl = reinterpret_cast<const float4 *>(planes) + b0 * 6;
    const CropBoxLocal *loc = local + b0;
    {
        const uint32_t *oc = occ + (int64_t)f * kOccWords;
        for (int w = threadIdx.x; w < kOccWords; w += blockDim.x) s_occ[w] = __ldg(oc + w);
    }
    if (kSmemTables) {
        // the expand stage gathers one cell entry per lane and the pair test a 48-byte box record per lane: scattered, so
        // from L2 they cost a full round trip per batch (measured: 0.085 ms of a 0.27 ms sweep for the cell entries alone)
        const float4 *g = reinterpret_cast<const float4 *>(loc);
        for (int t = threadIdx.x; t < 3 * B; t += blockDim.x) s_loc[t] = __ldg(g + t);
        if (kC4Smem) {
            const uint32_t *g4 = reinterpret_cast<const uint32_t *>(cell4) + (int64_t)f * cells;
            for (int t = threadIdx.x; t < cells; t += blockDim.x) s_c4[t] = __ldg(g4 + t);
        }
    }
    for (int b = threadIdx.x; b < B; b += blockDim.x) s_box_cnt[b] = 0;
    const bool par_rank = B <= rank_boxes;                            // ranking by all warps in parallel (else one warp, serially)
    if (par_rank) for (int t = threadIdx.x; t < kCropWarps * rank_boxes; t += blockDim.x) s_wcnt[t] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    float4 *cq = s_cq + wid * kCropCQ;
    int32_t *mine = s_wcnt + wid * rank_boxes;                        // this warp's per-box hit counts (par_rank)
    float4 *pq = s_pq + wid * kCropPQ;
    float4 *stage_xyz = hits_xyz + ((int64_t)blockIdx.x * kCropWarps + wid) * hit_cap;
    int32_t *stage = hits_br + ((int64_t)blockIdx.x * kCropWarps + wid) * hit_cap;
    int wcount = 0, phead = 0, pcount = 0, cqn = 0;
#ifdef CROP_ABLATE       // timing experiments only (scripts/gpu_r2_abl.sh): 1 prologue, 2 + filter, 3 + expand, 4 + test (no ranking)
    if (CROP_ABLATE == 1) return;
#endif

    // stage 3 on n <= 32 queued pairs (one per lane, in point order)
    auto test = [&](int n) {
        bool hit = false, unc = false;
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
        int pidx = 0, box = 0;
        if (lane < n) {
            e = pq[(phead + lane) & (kCropPQ - 1)];
            const int w = __float_as_int(e.w);
            pidx = w & ((1 << kIdxBits) - 1); box = (w >> kIdxBits) & 0x3FFF;
            float4 A, Bq;
            float mm;
            if (kSmemTables) { A = s_loc[box * 3]; Bq = s_loc[box * 3 + 1]; mm = s_loc[box * 3 + 2].x; }
            else {
                const float4 *L = reinterpret_cast<const float4 *>(loc + box);
                A = __ldg(L); Bq = __ldg(L + 1); mm = __ldg(reinterpret_cast<const float *>(L + 2));
            }
            const float dx = e.x - A.x, dy = e.y - A.y, dz = e.z - A.z;
            // world = [[c, s], [-s, c]] local  (rotation_3d_in_axis)  ->  local = [[c, -s], [s, c]] world
            const float lx = dx * A.w - dy * Bq.x, ly = dx * Bq.x + dy * A.w;
            const float t = fmaxf(fmaxf(fabsf(lx) - Bq.y, fabsf(ly) - Bq.z), fabsf(dz) - Bq.w);
            hit = t < -mm;
            unc = (w >> 30) != 0 || !(hit || t > mm);                  // weird point, irregular box (m = NaN), or inside the margin
        }
        CROP_STAT(4, 1);
        const unsigned um = __ballot_sync(0xffffffffu, unc);
        if (um) {
            CROP_STAT(5, 1); CROP_STAT(6, __popc(um));
            if (unc) hit = crop_inside(e.x, e.y, e.z, pl + box * 6);
        }
        const unsigned hm = __ballot_sync(0xffffffffu, hit);
        const int total = __popc(hm);
        CROP_STAT(7, total);
        if (total) {
            // a batch that does not fit is dropped as a whole (warp-uniform decision): the segment never has holes
            const bool fits = wcount + total <= stage_cap;
            if (!fits) { if (lane == 0) atomicExch(overflow, 3); }
            else {
                // (ranking the hit here among the warp's hits of its box -- match.any + the per-warp counters -- instead of in a
                // pass over the finished segment was measured: hits pass 0.246 -> 0.270 ms, the chain of a batch gets longer)
                if (hit) {
                    const int at = wcount + __popc(hm & lt_mask);
                    stage_xyz[at] = make_float4(e.x, e.y, e.z, __int_as_float(ck.first_pt + pidx));
                    stage[at] = box;
                }
                wcount += total;
            }
        }
        phead = (phead + n) & (kCropPQ - 1);
        pcount -= n;
        __syncwarp();
    };

    // stage 2, serial form: candidates cq[base .. base + n) one after the other, lanes over the boxes of the list
    auto expand_slow = [&](int base, int n) {
        CROP_STAT(3, 1);
        for (int c = 0; c < n; ++c) {
            const float4 e = cq[base + c];                               // same address in every lane: broadcast
            const bool weird = crop_weird(e.x, e.y, e.z);
            int e0 = 0, len = 0;
            if (weird) len = B;
            else {
                const int cx = crop_cell(e.x, m.x0, m.inv_x, G), cy = crop_cell(e.y, m.y0, m.inv_y, G);
                if (cx >= 0 && cx < G && cy >= 0 && cy < G) {
                    e0 = __ldg(cs + cy * G + cx);
                    len = min(__ldg(cs + cy * G + cx + 1), cell_cap) - e0;
                }
            }
            const int tag = __float_as_int(e.w) | (weird ? (1 << 30) : 0);
            for (int k0 = 0; k0 < len; k0 += 32) {
                const int k = k0 + lane;
                if (k < len) {
                    const int b = weird ? k : __ldg(cb + e0 + k);
                    pq[(phead + pcount + lane) & (kCropPQ - 1)] = make_float4(e.x, e.y, e.z, __int_as_float(tag | (b << kIdxBits)));
                }
                pcount += min(32, len - k0);
                CROP_STAT(1, min(32, len - k0));
                __syncwarp();
                while (pcount >= 32) test(32);
            }
        }
    };

    // stage 2 on n <= 32 candidates cq[base .. base + n) (one per lane)
    auto expand = [&](int base, int n) {
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
        uint2 ce = make_uint2(0u, 0u);
        int cnt = 0, e0 = -1;
        bool weird = false;
        CROP_STAT(0, n); CROP_STAT(2, 1);
        if (lane < n) {
            e = cq[base + lane];
            weird = crop_weird(e.x, e.y, e.z);
            // the same cell function the grid kernel registered the rectangles with (crop_cell, unclamped: the unsigned
            // comparison rejects everything outside [0, G))
            const int cx = __float2int_rd((e.x - m.x0) * m.inv_x), cy = __float2int_rd((e.y - m.y0) * m.inv_y);
            if (!weird && (unsigned)cx < (unsigned)G && (unsigned)cy < (unsigned)G) {
                if (kSmemTables && kC4Smem) {
                    const uint32_t w = s_c4[cy * G + cx];              // id0 | id1 << 10 | id2 << 20 | count << 30, all ones = long list
                    ce = make_uint2((w & 1023u) | (((w >> 10) & 1023u) << 16), ((w >> 20) & 1023u) | ((w == 0xFFFFFFFFu ? 0xFFFFu : (w >> 30)) << 16));
                } else ce = __ldg(c4 + cy * G + cx);                   // id0 | id1 << 16,  id2 | count << 16
                cnt = (int)(ce.y >> 16);
                if (cnt > kCellIds) {                                  // a longer list: this lane reads it from the CSR arrays
                    e0 = __ldg(cs + cy * G + cx);
                    cnt = min(__ldg(cs + cy * G + cx + 1), cell_cap) - e0;
                }
            }
        }
        // exclusive prefix of the pair counts over the lanes: counts of 0..3 (all but a few per cent of the batches) from two
        // ballots of the count's bits, anything longer from a shuffle scan
        int excl, total;
        if (!__any_sync(0xffffffffu, cnt > 3)) {
            const unsigned b0m = __ballot_sync(0xffffffffu, cnt & 1), b1m = __ballot_sync(0xffffffffu, cnt & 2);
            excl = __popc(b0m & lt_mask) + 2 * __popc(b1m & lt_mask);
            total = __popc(b0m) + 2 * __popc(b1m);
        } else {
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            total = __shfl_sync(0xffffffffu, incl, 31);
            excl = incl - cnt;
        }
        // weird points (every box) and batches that would not fit the ring go through the serial form
        if (__any_sync(0xffffffffu, weird) || total > kCropPQ - 32) { expand_slow(base, n); return; }
        if (total == 0) return;
        const int at = phead + pcount + excl;
        const int tag = __float_as_int(e.w);
        if (e0 < 0) {
            if (cnt > 0) pq[at & (kCropPQ - 1)] = make_float4(e.x, e.y, e.z, __int_as_float(tag | (int)((ce.x & 0xFFFFu) << kIdxBits)));
            if (cnt > 1) pq[(at + 1) & (kCropPQ - 1)] = make_float4(e.x, e.y, e.z, __int_as_float(tag | (int)((ce.x >> 16) << kIdxBits)));
            if (cnt > 2) pq[(at + 2) & (kCropPQ - 1)] = make_float4(e.x, e.y, e.z, __int_as_float(tag | (int)((ce.y & 0xFFFFu) << kIdxBits)));
        } else {
            for (int k = 0; k < cnt; ++k)
                pq[(at + k) & (kCropPQ - 1)] = make_float4(e.x, e.y, e.z, __int_as_float(tag | (__ldg(cb + e0 + k) << kIdxBits)));
        }
        pcount += total;
        CROP_STAT(1, total);
        __syncwarp();
#ifdef CROP_ABLATE
        if (CROP_ABLATE == 3) { phead = (phead + pcount) & (kCropPQ - 1); pcount = 0; }
#endif
        while (pcount >= 32) test(32);
    };

    // the warps split the chunk into runs of whole iterations
    const int wp = ((ck.n_pts + kCropWarps * kCropIter - 1) / (kCropWarps * kCropIter)) * kCropIter;
    const int w_lo = min(wid * wp, ck.n_pts), w_hi = min(w_lo + wp, ck.n_pts);
    const int ps = (int)pt_stride;
    const float ofx = -m.x0 * m.inv_fx, ofy = -m.y0 * m.inv_fy;      // fine cell = floor(x * inv + of): the bitmap is conservative
    const bool any_box = B > 0;                                       // by centimetres, so this need not be crop_cell's rounding
    // Full iterations of contiguous xyz points whose run starts on a 16-byte boundary arrive by TMA: lane 0 requests the
    // next 1536 bytes as soon as the warp has copied the current ones to registers, so the copy is in flight during the
    // whole iteration, costs no load instructions and no L1 wavefronts, and the twelve shared-memory reads at a stride of
    // three words are conflict-free.  Anything else (strided points, odd alignment, the tail of a run) uses plain loads.
    float *buf = s_pt + wid * (kCropIter * 3);
    uint64_t *bar = s_bar + wid;
    const bool tma_ok = ps == 3 && (reinterpret_cast<uintptr_t>(pts + (int64_t)w_lo * 3) & 15) == 0;
    uint32_t tma_phase = 0;
    if (lane == 0) {
        crop_mbar_init(bar);
        if (tma_ok && w_lo + kCropIter <= w_hi) crop_bulk_load(buf, pts + (int64_t)w_lo * 3, kCropIter * 12, bar);
    }
    __syncwarp();
    for (int base = w_lo; base < w_hi; base += kCropIter) {
        const bool full = base + kCropIter <= w_hi;
        float px[4], py[4], pz[4];
        if (full && tma_ok) {
            crop_mbar_wait(bar, tma_phase);
            tma_phase ^= 1u;
#pragma unroll
            for (int j = 0; j < 4; ++j) { px[j] = buf[(j * 32 + lane) * 3]; py[j] = buf[(j * 32 + lane) * 3 + 1]; pz[j] = buf[(j * 32 + lane) * 3 + 2]; }
            __syncwarp();
            if (lane == 0 && base + 2 * kCropIter <= w_hi) crop_bulk_load(buf, pts + (int64_t)(base + kCropIter) * 3, kCropIter * 12, bar);
        } else if (full && ps == 3) {
            const float *q = pts + (base + lane) * 3;                  // 12-byte stride: immediate offsets, any alignment
#pragma unroll
            for (int j = 0; j < 4; ++j) { px[j] = __ldg(q + j * 96); py[j] = __ldg(q + j * 96 + 1); pz[j] = __ldg(q + j * 96 + 2); }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float *q = pts + (int64_t)min(base + j * 32 + lane, w_hi - 1) * ps;
                px[j] = __ldg(q); py[j] = __ldg(q + 1); pz[j] = __ldg(q + 2);
            }
        }
        // (Requesting a survivor's coarse cell entry here, right after its filter, and queueing it behind the four filters was
        // measured: four sparse L2 round trips per iteration instead of 0.7 dense ones, hits pass 0.25 -> 0.30 ms.  So was
        // expanding two batches of 32 at a time with both batches' cell entries in flight together: 0.23 -> 0.29 ms.)
#pragma unroll
        for (int half = 0; half < 4 / kPerCheck; ++half) {
#pragma unroll
            for (int jj = 0; jj < kPerCheck; ++jj) {
                const int j = half * kPerCheck + jj;
                const float x = px[j], y = py[j], z = pz[j];
                const int ix = __float2int_rd(fmaf(x, m.inv_fx, ofx)), iy = __float2int_rd(fmaf(y, m.inv_fy, ofy));
                const int bit = iy * kOccRes + ix;
                const uint32_t word = s_occ[(bit >> 5) & (kOccWords - 1)];
                const bool near_box = (unsigned)(ix | iy) < (unsigned)kOccRes && z >= m.zmin && z <= m.zmax && ((word >> (bit & 31)) & 1u);
                bool ok = crop_weird(x, y, z) ? any_box : near_box;    // empty frame: zmin = +inf, nothing passes
                if (!full) ok = ok && base + j * 32 + lane < w_hi;
                const unsigned bm = __ballot_sync(0xffffffffu, ok);
                if (ok) cq[cqn + __popc(bm & lt_mask)] = make_float4(x, y, z, __int_as_float(base + j * 32 + lane));
                cqn += __popc(bm);
            }
            __syncwarp();
#ifdef CROP_ABLATE
            if (CROP_ABLATE == 2) cqn = 0;
#endif
            if (cqn >= 32) {
                int q = 0;
                for (; q + 32 <= cqn; q += 32) expand(q, 32);
                // carry the remainder (< 32 candidates) to the front of the queue
                const int rem = cqn - q;
                float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
                if (lane < rem) e = cq[q + lane];
                __syncwarp();
                if (lane < rem) cq[lane] = e;
                cqn = rem;
                __syncwarp();
            }
        }
    }
    if (cqn > 0) expand(0, cqn);
    while (pcount > 0) test(min(pcount, 32));
#ifdef CROP_ABLATE
    if (CROP_ABLATE >= 2) return;
#endif
    if (lane == 0) { warp_total[wid] = wcount; n_hits[(int64_t)blockIdx.x * kCropWarps + wid] = wcount; }
    __syncwarp();
    // ---- rank of every hit among the hits of the same box in this chunk (point order).
    //      Parallel form: every warp first ranks its own (ordered) segment against its own per-box counters -- 32 hits at a
    //      time, equal boxes inside a group ranked by lane -- then adds the counts of the warps before it.
    if (par_rank) {
        for (int h0 = 0; h0 < wcount; h0 += 32) {
            const int h = h0 + lane;
            const bool act = h < wcount;
            const int box = act ? stage[h] : -1 - lane;
            const unsigned same = __match_any_sync(0xffffffffu, box);
            if (act) stage[h] = box | ((mine[box] + __popc(same & lt_mask)) << 16);
            __syncwarp();
            if (act && (same >> lane) == 1u) mine[box] += __popc(same);          // highest lane of each group updates
            __syncwarp();
        }
        __syncthreads();
        if (wid > 0)
            for (int h = lane; h < wcount; h += 32) {
                const int br = stage[h];
                const int box = br & 0xFFFF;
                int prefix = 0;
                for (int w = 0; w < wid; ++w) prefix += s_wcnt[w * rank_boxes + box];
                if (prefix) stage[h] = br + (prefix << 16);
            }
        for (int b = threadIdx.x; b < B; b += blockDim.x) {
            int sum = 0;
#pragma unroll
            for (int w = 0; w < kCropWarps; ++w) sum += s_wcnt[w * rank_boxes + b];
            chunk_box_count[(int64_t)blockIdx.x * max_boxes + b] = sum;
        }
        return;
    }
    // Serial form (frames with more boxes than the per-warp counter table holds): one warp walks the segments in order.
    __threadfence_block();
    __syncthreads();
    if (wid == 0) {
        for (int w = 0; w < kCropWarps; ++w) {
            const int cnt = warp_total[w];
            int32_t *l = hits_br + ((int64_t)blockIdx.x * kCropWarps + w) * hit_cap;
            for (int h0 = 0; h0 < cnt; h0 += 32) {
                const int h = h0 + lane;
                const bool act = h < cnt;
                const int box = act ? l[h] : -1 - lane;
                const unsigned same = __match_any_sync(0xffffffffu, box);
                if (act) l[h] = box | ((s_box_cnt[box] + __popc(same & ((1u << lane) - 1u))) << 16);
                __syncwarp();
                if (act && (same >> lane) == 1u) s_box_cnt[box] += __popc(same);  // highest lane of each group updates
                __syncwarp();
            }
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += blockDim.x) chunk_box_count[(int64_t)blockIdx.x * max_boxes + b] = s_box_cnt[b];
}

// per frame: exclusive scan over the frame's chunks for every box, and the box totals.  One thread per box (a warp reads
// 32 consecutive boxes of a chunk's row: coalesced), the frame's chunks walked in order.
__global__ void crop_scan_kernel(const int64_t *__restrict__ box_off, const int64_t *__restrict__ frame_chunk_off,
                                 int32_t *__restrict__ chunk_box_count, int max_boxes, int32_t *__restrict__ box_total)
{
    const int f = blockIdx.x;
    const int64_t b0 = box_off[f];
    const int B = (int)(box_off[f + 1] - b0);
    const int64_t c0 = frame_chunk_off[f], c1 = frame_chunk_off[f + 1];
    for (int b = blockIdx.y * blockDim.x + threadIdx.x; b < B; b += gridDim.y * blockDim.x) {
        int run = 0;
        // sixteen chunks at a time, all loads issued before the first is used (the walk is a chain of L2 round trips otherwise:
        // four at a time took 0.022 ms for eleven chunks per frame)
        for (int64_t c = c0; c < c1; c += 16) {
            int v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = c + k < c1 ? chunk_box_count[(c + k) * max_boxes + b] : 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                if (c + k < c1) chunk_box_count[(c + k) * max_boxes + b] = run;
                run += v[k];
            }
        }
        box_total[b0 + b] = run;
    }
}

// global exclusive offsets over all boxes of all frames.  CTA k owns the 1024 boxes [1024 k, 1024 k + 1024): it sums the
// totals in front of its tile itself (n_boxes is tens of thousands: a few coalesced loads per thread) and scans its tile.
__global__ void __launch_bounds__(1024)
crop_offsets_kernel(const int32_t *__restrict__ box_total, int64_t n_boxes, int64_t *__restrict__ offsets)
{
    __shared__ long long red[32];
    __shared__ long long warp_incl[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t t0 = (int64_t)blockIdx.x * 1024;
    long long before = 0;
    for (int64_t i = threadIdx.x; i < t0; i += 1024) before += box_total[i];
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    if (lane == 0) red[w] = before;
    const int64_t i = t0 + threadIdx.x;
    const long long v = i < n_boxes ? box_total[i] : 0;
    long long incl = v;
    for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_incl[w] = incl;
    __syncthreads();
    long long base = 0;
    for (int k = 0; k < 32; ++k) base += red[k];                 // fixed order
    for (int k = 0; k < w; ++k) base += warp_incl[k];
    if (i < n_boxes) offsets[i] = base + incl - v;
    if (i == n_boxes - 1) offsets[n_boxes] = base + incl;
    if (n_boxes == 0 && blockIdx.x == 0 && threadIdx.x == 0) offsets[0] = 0;
}

__global__ void __launch_bounds__(kCropThreads)
crop_fill_kernel(const int64_t *__restrict__ box_off, const CropChunk *__restrict__ chunks, const float4 *__restrict__ hits_xyz,
                 const int32_t *__restrict__ hits_br, int hit_cap, const int32_t *__restrict__ n_hits,
                 const int32_t *__restrict__ chunk_box_count, int max_boxes, const int64_t *__restrict__ offsets, int64_t capacity,
                 int write_slots, int32_t *__restrict__ out_idx, int32_t *__restrict__ overflow)
{
    // destination of a hit = global offset of its box + hits of that box in earlier chunks of the frame + rank in this chunk.
    // Only four bytes are scattered here: the hit's SLOT when crop_materialise_kernel follows (it writes every box's list
    // front to back from the hit records; seven scattered stores per hit in this kernel were what bounded it), the point
    // index itself when only the index lists are wanted.
    const CropChunk ck = chunks[blockIdx.x];
    const int64_t b0 = box_off[ck.frame];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = min(n_hits[(int64_t)blockIdx.x * kCropWarps + wid], hit_cap);          // this warp's segment
    const int64_t seg = ((int64_t)blockIdx.x * kCropWarps + wid) * hit_cap;
    for (int h = lane; h < total; h += 32) {
        const int br = hits_br[seg + h];
        const int box = br & 0xFFFF;
        const int64_t dst = offsets[b0 + box] + chunk_box_count[(int64_t)blockIdx.x * max_boxes + box] + (br >> 16);
        if (dst >= capacity) { atomicExch(overflow, 4); continue; }
        out_idx[dst] = write_slots ? (int32_t)(seg + h) : __float_as_int(hits_xyz[seg + h].w);
    }
}

// The lists of a frame's boxes are one contiguous range of the output, so the threads simply stride over it (coalesced,
// every element independent of the others): idx[k] holds the slot of the k-th hit; its record gives idx[k] = point index,
// out_xyz[k] = the point and out_xyz_global[k] = pose_f [x y z 1]; waymo_common.py:169-171.
__global__ void __launch_bounds__(256)
crop_materialise_kernel(const float4 *__restrict__ hits_xyz, const int64_t *__restrict__ box_off, const int64_t *__restrict__ offsets,
                        const double *__restrict__ poses, int64_t capacity, int32_t *__restrict__ idx, float *__restrict__ out_xyz,
                        double *__restrict__ out_xyz_global)
{
    const int f = blockIdx.x;
    const int64_t lo = offsets[box_off[f]], hi = min(offsets[box_off[f + 1]], capacity);
    const double *P = poses ? poses + (int64_t)f * 16 : nullptr;
    double Pm[12];
    if (P) {
#pragma unroll
        for (int i = 0; i < 12; ++i) Pm[i] = P[i];
    }
    for (int64_t k = lo + blockIdx.y * blockDim.x + threadIdx.x; k < hi; k += (int64_t)gridDim.y * blockDim.x) {
        const float4 r = __ldg(hits_xyz + idx[k]);
        idx[k] = __float_as_int(r.w);
        if (out_xyz) { out_xyz[k * 3] = r.x; out_xyz[k * 3 + 1] = r.y; out_xyz[k * 3 + 2] = r.z; }
        if (out_xyz_global && P) {
            const double dx = r.x, dy = r.y, dz = r.z;
#pragma unroll
            for (int q = 0; q < 3; ++q)
                out_xyz_global[k * 3 + q] = ((Pm[q * 4] * dx + Pm[q * 4 + 1] * dy) + Pm[q * 4 + 2] * dz) + Pm[q * 4 + 3];
        }
    }
}

// dense (N, B) uint8 mask from the index lists of ONE frame (drop-in for box_np_ops.points_in_rbbox)
__global__ void crop_dense_mask_kernel(const int32_t *__restrict__ idx, const int64_t *__restrict__ offsets, int n_boxes,
                                       uint8_t *__restrict__ mask)
{
    const int b = blockIdx.y;
    const int64_t lo = offsets[b], hi = offsets[b + 1];
    for (int64_t i = lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x)
        mask[(int64_t)idx[i] * n_boxes + b] = 1;
}

}  // namespace al3d

using namespace al3d;

extern "C" int al3d_crop_chunk_points(void) { return kCropChunk; }

extern "C" int al3d_crop_box_setup(const float *boxes, const float *sincos, int64_t n_boxes, float pad_abs, float pad_rel, float *planes,
                                   float *aabb, void *stream)
{
    AL3D_CHECK_ARG(n_boxes >= 0, "al3d_crop_box_setup: negative size");
    if (n_boxes == 0) return 0;
    AL3D_CHECK_ARG(boxes && sincos && planes && aabb, "al3d_crop_box_setup: null pointer");
    crop_box_setup_kernel<<<(unsigned)ceil_div(n_boxes, 128), 128, 0, (cudaStream_t)stream>>>(boxes, sincos, n_boxes, pad_abs, pad_rel, planes, aabb);
    AL3D_CHECK_LAUNCH("crop_box_setup_kernel");
    return 0;
}

extern "C" int al3d_crop_box_local(const float *boxes, const float *sincos, int64_t n_boxes, float *local, void *stream)
{
    AL3D_CHECK_ARG(n_boxes >= 0, "al3d_crop_box_local: negative size");
    if (n_boxes == 0) return 0;
    AL3D_CHECK_ARG(boxes && sincos && local, "al3d_crop_box_local: null pointer");
    AL3D_CHECK_ARG((reinterpret_cast<uintptr_t>(local) & 15) == 0, "al3d_crop_box_local: local must be 16-byte aligned");
    crop_box_local_kernel<<<(unsigned)ceil_div(n_boxes, 128), 128, 0, (cudaStream_t)stream>>>(boxes, sincos, n_boxes,
                                                                                              reinterpret_cast<CropBoxLocal *>(local));
    AL3D_CHECK_LAUNCH("crop_box_local_kernel");
    return 0;
}

extern "C" int al3d_crop_occ_words(void) { return kOccWords; }
#ifdef CROP_STATS
extern "C" int al3d_crop_stats(unsigned long long *out8, int reset)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out8, g_crop_stats, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_crop_stats, z, sizeof(z)); }
    return 0;
}
#endif

extern "C" int al3d_crop_build_grid(const float *aabb, const float *boxes, const float *sincos, const int64_t *box_off, int n_frames, int G,
                                    float *grid_meta, int32_t *cell_start, int32_t *cell_boxes, int cell_cap, uint32_t *cell4,
                                    int max_boxes, uint32_t *occ, int32_t *overflow, void *stream)
{
    AL3D_CHECK_ARG(aabb && box_off && grid_meta && cell_start && cell_boxes && overflow, "al3d_crop_build_grid: null pointer");
    AL3D_CHECK_ARG(!occ || (boxes && sincos), "al3d_crop_build_grid: the occupancy bitmap needs boxes and sincos");
    AL3D_CHECK_ARG(G >= 1 && G <= 64, "al3d_crop_build_grid: G=%d not in [1,64]", G);
    if (n_frames <= 0) return 0;
    const size_t smem = (size_t)(2 * G * G + kOccWords / kGridBands) * sizeof(int32_t);
    crop_grid_kernel<<<dim3(n_frames, kGridBands), kGridThreads, smem, (cudaStream_t)stream>>>(
        aabb, boxes, sincos, box_off, G, reinterpret_cast<CropGridMeta *>(grid_meta), cell_start, cell_boxes, cell_cap,
        reinterpret_cast<uint2 *>(cell4), kC4Smem && max_boxes <= kLocSmemBoxes, occ, overflow);
    AL3D_CHECK_LAUNCH("crop_grid_kernel");
    return 0;
}

extern "C" int al3d_crop_hits(const float *const *frame_points, int64_t pt_stride, const float *planes,
                              const float *local, const int64_t *box_off, int G, const float *grid_meta, const int32_t *cell_start,
                              const int32_t *cell_boxes, int cell_cap, const uint32_t *cell4, const uint32_t *occ, const int32_t *chunks,
                              int n_chunks, void *hits, int hit_cap, int32_t *n_hits, int32_t *chunk_box_count, int max_boxes,
                              int32_t *overflow, void *stream)
{
    AL3D_CHECK_ARG(frame_points && planes && local && box_off && grid_meta && cell_start && cell_boxes && cell4 && occ && chunks && hits &&
                   n_hits && chunk_box_count && overflow, "al3d_crop_hits: null pointer");
    AL3D_CHECK_ARG(pt_stride >= 3, "al3d_crop_hits: pt_stride=%lld", (long long)pt_stride);
    AL3D_CHECK_ARG(max_boxes >= 1 && max_boxes <= 12288, "al3d_crop_hits: max_boxes=%d not in [1,12288]", max_boxes);
    AL3D_CHECK_ARG(hit_cap >= 1 && hit_cap <= (1 << 20), "al3d_crop_hits: hit_cap=%d not in [1, 2^20]", hit_cap);
    AL3D_CHECK_ARG((reinterpret_cast<uintptr_t>(local) & 15) == 0 && (reinterpret_cast<uintptr_t>(cell4) & 7) == 0 &&
                   (reinterpret_cast<uintptr_t>(hits) & 15) == 0, "al3d_crop_hits: local / cell4 / hits misaligned");
    AL3D_CHECK_ARG((int64_t)n_chunks * kCropWarps * hit_cap < (int64_t)1 << 31, "al3d_crop_hits: more than 2^31 hit slots");
    if (n_chunks <= 0) return 0;
    const int rank_boxes = max_boxes <= 512 ? max_boxes : 0;          // per-warp box counters for the parallel ranking, if modest
    const bool loc_smem = max_boxes <= kLocSmemBoxes;                 // must agree with the format al3d_crop_build_grid chose
    const size_t smem = (size_t)kCropWarps * (kCropCQ + kCropPQ) * sizeof(float4) + (size_t)kCropWarps * kCropIter * 12 +
                        (loc_smem ? (size_t)max_boxes * sizeof(CropBoxLocal) + (kC4Smem ? (size_t)G * G * sizeof(uint32_t) : 0) : 0) +
                        (size_t)kOccWords * sizeof(uint32_t) + (size_t)((max_boxes + 3) & ~3) * sizeof(int32_t) +
                        (size_t)kCropWarps * rank_boxes * sizeof(int32_t);
    AL3D_CHECK_ARG(smem <= 200 * 1024, "al3d_crop_hits: max_boxes=%d needs too much shared memory", max_boxes);
    auto kern = loc_smem ? crop_hits_kernel<true> : crop_hits_kernel<false>;
    if (smem > 48 * 1024) AL3D_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<n_chunks, kCropThreads, smem, (cudaStream_t)stream>>>(
        frame_points, pt_stride, planes, reinterpret_cast<const CropBoxLocal *>(local), box_off, G,
        reinterpret_cast<const CropGridMeta *>(grid_meta), cell_start, cell_boxes, cell_cap, reinterpret_cast<const uint2 *>(cell4), occ,
        reinterpret_cast<const CropChunk *>(chunks), reinterpret_cast<float4 *>(hits),
        reinterpret_cast<int32_t *>(reinterpret_cast<float4 *>(hits) + (int64_t)n_chunks * kCropWarps * hit_cap), hit_cap, n_hits, chunk_box_count, max_boxes,
        rank_boxes, overflow);
    AL3D_CHECK_LAUNCH("crop_hits_kernel");
    return 0;
}

extern "C" int al3d_crop_scan(const int64_t *box_off, const int64_t *frame_chunk_off, int n_frames, int64_t n_boxes,
                              int32_t *chunk_box_count, int max_boxes, int32_t *box_total, int64_t *offsets, void *stream)
{
    AL3D_CHECK_ARG(box_off && frame_chunk_off && chunk_box_count && box_total && offsets, "al3d_crop_scan: null pointer");
    if (n_frames > 0) {
        crop_scan_kernel<<<dim3(n_frames, 1), 256, 0, (cudaStream_t)stream>>>(box_off, frame_chunk_off, chunk_box_count, max_boxes, box_total);
        AL3D_CHECK_LAUNCH("crop_scan_kernel");
    }
    crop_offsets_kernel<<<(unsigned)std::max<int64_t>(ceil_div(n_boxes, 1024), 1), 1024, 0, (cudaStream_t)stream>>>(box_total, n_boxes, offsets);
    AL3D_CHECK_LAUNCH("crop_offsets_kernel");
    return 0;
}

extern "C" int al3d_crop_hit_bytes(void) { return kCropHitBytes; }

extern "C" int al3d_crop_fill(const int64_t *box_off, int n_frames,
                              const int32_t *chunks, int n_chunks, const void *hits, int hit_cap, const int32_t *n_hits,
                              const int32_t *chunk_box_count, int max_boxes, const int64_t *offsets, const double *poses,
                              int64_t capacity, int32_t *out_idx, float *out_xyz, double *out_xyz_global, int32_t *overflow,
                              void *stream)
{
    AL3D_CHECK_ARG(box_off && chunks && hits && n_hits && chunk_box_count && offsets && out_idx && overflow, "al3d_crop_fill: null pointer");
    if (n_chunks <= 0 || n_frames <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    AL3D_CHECK_ARG((reinterpret_cast<uintptr_t>(hits) & 15) == 0, "al3d_crop_fill: hits must be 16-byte aligned");
    AL3D_CHECK_ARG((int64_t)n_chunks * kCropWarps * hit_cap < (int64_t)1 << 31, "al3d_crop_fill: more than 2^31 hit slots");
    const float4 *hx = reinterpret_cast<const float4 *>(hits);
    const int32_t *hb = reinterpret_cast<const int32_t *>(hx + (int64_t)n_chunks * kCropWarps * hit_cap);
    const bool materialise = out_xyz || (out_xyz_global && poses);
    crop_fill_kernel<<<n_chunks, kCropThreads, 0, st>>>(box_off, reinterpret_cast<const CropChunk *>(chunks), hx, hb, hit_cap, n_hits,
                                                        chunk_box_count, max_boxes, offsets, capacity, materialise ? 1 : 0, out_idx, overflow);
    AL3D_CHECK_LAUNCH("crop_fill_kernel");
    if (materialise) {
        // enough CTAs per frame to fill the GPU a few times over whatever the number of frames
        const unsigned per_frame = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(148 * 16, n_frames), 256));
        crop_materialise_kernel<<<dim3(n_frames, per_frame), 256, 0, st>>>(
            hx, box_off, offsets, poses, capacity, out_idx, out_xyz, out_xyz_global);
        AL3D_CHECK_LAUNCH("crop_materialise_kernel");
    }
    return 0;
}

extern "C" int al3d_crop_dense_mask(const int32_t *idx, const int64_t *offsets, int n_boxes, uint8_t *mask, void *stream)
{
    AL3D_CHECK_ARG(idx && offsets && mask, "al3d_crop_dense_mask: null pointer");
    if (n_boxes <= 0) return 0;
    crop_dense_mask_kernel<<<dim3(8, n_boxes), 256, 0, (cudaStream_t)stream>>>(idx, offsets, n_boxes, mask);
    AL3D_CHECK_LAUNCH("crop_dense_mask_kernel");
    return 0;
}
