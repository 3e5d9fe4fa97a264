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
// The reference tests every point against every box (N*B*6 plane evaluations per frame).  Here a BEV grid
// per frame maps a point to the few boxes whose (padded) bounding rectangle covers its cell, so each
// point does O(1) exact tests and the kernels are bound by reading the points once from HBM.
//
// Pipeline (all launches cover the whole batch of frames):
//   crop_grid_kernel   one CTA per frame: grid extent from the box AABBs, CSR cell -> box lists
//   crop_hits_kernel   one CTA per chunk of 2048 points: exact tests, ordered (point, box, rank) hit list
//                      per chunk and per-(chunk, box) counts
//   crop_scan_kernel   per frame: exclusive scan over chunks for every box; crop_offsets_kernel: global offsets
//   crop_fill_kernel   scatter indices / xyz / global xyz to their final, ascending positions
#include "common.cuh"
#include "../../include/al3d.h"

namespace al3d {

constexpr int kCropChunk = 2048;       // points per CTA of the hits kernel
constexpr int kCropThreads = 256;
constexpr int kMaxHitsPerPoint = 8;    // a point inside more boxes than this raises overflow code 2 (rectangle candidates do not count)

struct CropGridMeta { float x0, y0, inv_x, inv_y; };

// cell index of a coordinate; the SAME expression is used for points and for box rectangles, and it is
// monotone in v, so a point inside a rectangle always lands in a cell the rectangle was registered in.
__device__ __forceinline__ int crop_cell(float v, float v0, float inv, int G)
{
    const float c = floorf((v - v0) * inv);
    return c < 0.f ? -1 : (c >= (float)G ? G : (int)c);
}

__global__ void __launch_bounds__(kCropThreads)
crop_grid_kernel(const float *__restrict__ aabb, const int64_t *__restrict__ box_off, int G, CropGridMeta *__restrict__ meta,
                 int32_t *__restrict__ cell_start, int32_t *__restrict__ cell_boxes, int cell_cap, int32_t *__restrict__ overflow)
{
    extern __shared__ int32_t s_cnt[];            // G*G counts -> exclusive offsets; then G*G fill cursors
    __shared__ float red[4][kCropThreads / 32];
    __shared__ int32_t s_warp[kCropThreads / 32];
    const int f = blockIdx.x;
    const int64_t b0 = box_off[f];
    const int B = (int)(box_off[f + 1] - b0);
    const int cells = G * G;
    int32_t *s_cur = s_cnt + cells;
    // ---- extent of the boxes of this frame
    float xmin = INFINITY, ymin = INFINITY, xmax = -INFINITY, ymax = -INFINITY;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float *a = aabb + (b0 + b) * 6;
        xmin = fminf(xmin, a[0]); ymin = fminf(ymin, a[1]);
        xmax = fmaxf(xmax, a[3]); ymax = fmaxf(ymax, a[4]);
    }
    for (int o = 16; o > 0; o >>= 1) {
        xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, o)); ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o)); ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { red[0][wid] = xmin; red[1][wid] = ymin; red[2][wid] = xmax; red[3][wid] = ymax; }
    for (int c = threadIdx.x; c < 2 * cells; c += blockDim.x) s_cnt[c] = 0;
    __syncthreads();
    for (int w = 0; w < kCropThreads / 32; ++w) {
        xmin = fminf(xmin, red[0][w]); ymin = fminf(ymin, red[1][w]);
        xmax = fmaxf(xmax, red[2][w]); ymax = fmaxf(ymax, red[3][w]);
    }
    CropGridMeta m;
    if (B == 0 || !(xmax >= xmin) || !(ymax >= ymin)) { m.x0 = 0.f; m.y0 = 0.f; m.inv_x = 0.f; m.inv_y = 0.f; }
    else {
        m.x0 = xmin; m.y0 = ymin;
        m.inv_x = (float)G / fmaxf(xmax - xmin, 1e-3f) * 0.999f;     // keep xmax inside the last cell
        m.inv_y = (float)G / fmaxf(ymax - ymin, 1e-3f) * 0.999f;
    }
    if (threadIdx.x == 0) meta[f] = m;
    // ---- pass 0: every box adds itself to the counters of the cells its rectangle covers (integer adds commute);
    //      pass 1: it appends its id through a per-cell cursor; the short lists are sorted afterwards, so the
    //      final cell -> box lists are in ascending box order whatever the order of the atomics.
    for (int pass = 0; pass < 2; ++pass) {
        for (int b = threadIdx.x; b < B; b += blockDim.x) {
            const float *a = aabb + (b0 + b) * 6;
            int x0c = crop_cell(a[0], m.x0, m.inv_x, G), x1c = crop_cell(a[3], m.x0, m.inv_x, G);
            int y0c = crop_cell(a[1], m.y0, m.inv_y, G), y1c = crop_cell(a[4], m.y0, m.inv_y, G);
            x0c = max(x0c, 0); y0c = max(y0c, 0); x1c = min(x1c, G - 1); y1c = min(y1c, G - 1);
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
            const int per = (cells + kCropThreads - 1) / kCropThreads;
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
            for (int w = 0; w < kCropThreads / 32; ++w) total += s_warp[w];
            if (threadIdx.x == 0) {
                cell_start[(int64_t)f * (cells + 1) + cells] = total;
                if (total > cell_cap) atomicExch(overflow, 1);
            }
            for (int c = threadIdx.x; c < cells; c += blockDim.x) cell_start[(int64_t)f * (cells + 1) + c] = s_cnt[c];
        }
    }
    // ---- sort every cell's list (a handful of entries): insertion sort by one thread per cell
    __threadfence_block();
    for (int c = threadIdx.x; c < cells; c += blockDim.x) {
        const int lo = s_cnt[c], n = min(s_cur[c], max(cell_cap - lo, 0));
        int32_t *l = cell_boxes + (int64_t)f * cell_cap + lo;
        for (int i = 1; i < n; ++i) {
            const int v = l[i];
            int j = i - 1;
            while (j >= 0 && l[j] > v) { l[j + 1] = l[j]; --j; }
            l[j + 1] = v;
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

constexpr int kCropWarps = kCropThreads / 32;
constexpr int kCropWarpPts = kCropChunk / kCropWarps;     // consecutive points owned by one warp
constexpr int kCropQueue = 320;                           // per-warp candidate queue (>= 31 + 32 * kMaxHitsPerPoint)
constexpr int kCropLook = 2;                              // cell-list entries fetched ahead per point

__global__ void __launch_bounds__(kCropThreads)
crop_hits_kernel(const float *__restrict__ points, int64_t pt_stride, const int64_t *__restrict__ pt_off,
                 const float *__restrict__ planes, const float *__restrict__ aabb, const int64_t *__restrict__ box_off, int G,
                 const CropGridMeta *__restrict__ meta, const int32_t *__restrict__ cell_start,
                 const int32_t *__restrict__ cell_boxes, int cell_cap, const CropChunk *__restrict__ chunks,
                 int2 *__restrict__ hits, int hit_cap, int32_t *__restrict__ n_hits, int32_t *__restrict__ chunk_box_count,
                 int max_boxes, int rank_boxes, int32_t *__restrict__ overflow)
{
    // Each warp owns kCropWarpPts CONSECUTIVE points and walks them 32 at a time.  A lane looks up the boxes
    // registered in its point's BEV cell and keeps those whose padded bounding box contains the point; these
    // (point, box) pairs go, in point order, into a per-warp queue.  Whenever 32 pairs are queued the warp runs
    // the exact six-plane predicate on them in lock-step (no divergence) and appends the hits, still in point
    // order, to its staging list.  One block-wide prefix over the eight warp totals then places every warp's
    // list in the chunk's ordered hit list.  No block barrier inside the point loop.
    extern __shared__ int32_t s_dyn[];
    int32_t *s_box_cnt = s_dyn;                                        // per-box hit counters (max_boxes)
    int2 *s_queue = reinterpret_cast<int2 *>(s_dyn + ((max_boxes + 1) & ~1));   // kCropWarps x kCropQueue
    int2 *s_stage = s_queue + kCropWarps * kCropQueue;                  // kCropWarps x stage_cap
    __shared__ int warp_total[kCropWarps];
    const int stage_cap = min(hit_cap, kCropWarpPts * kMaxHitsPerPoint); // per warp (its worst case); the chunk total is capped at hit_cap
    int32_t *s_wcnt = reinterpret_cast<int32_t *>(s_stage + (size_t)kCropWarps * stage_cap);   // kCropWarps x rank_boxes per-warp box counts
    const CropChunk ck = chunks[blockIdx.x];
    const int f = ck.frame;
    const int64_t b0 = box_off[f];
    const int B = (int)(box_off[f + 1] - b0);
    const CropGridMeta m = meta[f];
    const int cells = G * G;
    const int32_t *cs = cell_start + (int64_t)f * (cells + 1);
    const int32_t *cb = cell_boxes + (int64_t)f * cell_cap;
    const float *pts = points + (pt_off[f] + ck.first_pt) * pt_stride;
    const float4 *pl = reinterpret_cast<const float4 *>(planes) + b0 * 6;
    const float2 *bb = reinterpret_cast<const float2 *>(aabb) + b0 * 3;
    int2 *my_hits = hits + (int64_t)blockIdx.x * hit_cap;      // .x = point index in frame, .y = box | rank << 16
    for (int b = threadIdx.x; b < B; b += blockDim.x) s_box_cnt[b] = 0;
    const bool par_rank = B <= rank_boxes;                            // ranking by all warps in parallel (else one warp, serially)
    if (par_rank) for (int t = threadIdx.x; t < kCropWarps * rank_boxes; t += blockDim.x) s_wcnt[t] = 0;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int2 *queue = s_queue + wid * kCropQueue;
    int2 *stage = s_stage + (size_t)wid * stage_cap;
    int wcount = 0, qhead = 0, qcount = 0;

    // exact test of up to 32 queued pairs (one per lane), hits appended in queue order
    auto drain = [&](int n) {
        bool hit = false;
        int2 e = make_int2(0, 0);
        if (lane < n) {
            int at = qhead + lane;
            if (at >= kCropQueue) at -= kCropQueue;
            e = queue[at];
            if (e.y & (1 << 30)) { hit = true; e.y &= ~(1 << 30); }     // NaN point: already decided
            else {
                const float *q = pts + (int64_t)e.x * pt_stride;
                hit = crop_inside(__ldg(q), __ldg(q + 1), __ldg(q + 2), pl + e.y * 6);
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            const int at = wcount + __popc(bal & ((1u << lane) - 1u));
            if (at < stage_cap) stage[at] = make_int2(ck.first_pt + e.x, e.y);
            else atomicExch(overflow, 3);
        }
        wcount += __popc(bal);
        qhead += n; if (qhead >= kCropQueue) qhead -= kCropQueue;
        qcount -= n;
    };

    const int w_lo = wid * kCropWarpPts, w_hi = min(w_lo + kCropWarpPts, ck.n_pts);
    constexpr int U = 4;                       // points per lane per outer iteration: U independent load chains in flight
    for (int i0 = w_lo; i0 < w_hi; i0 += 32 * U) {
        float px[U], py[U], pz[U];
        int e0[U], e1[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int i = i0 + k * 32 + lane;
            px[k] = py[k] = pz[k] = 0.f;
            if (i < w_hi) {
                px[k] = __ldg(pts + i * pt_stride); py[k] = __ldg(pts + i * pt_stride + 1); pz[k] = __ldg(pts + i * pt_stride + 2);
            }
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int i = i0 + k * 32 + lane;
            e0[k] = e1[k] = 0;
            const bool isnan_k = (px[k] != px[k] || py[k] != py[k] || pz[k] != pz[k]);
            if (i < w_hi && !isnan_k) {
                const int cx = crop_cell(px[k], m.x0, m.inv_x, G), cy = crop_cell(py[k], m.y0, m.inv_y, G);
                if (cx >= 0 && cx < G && cy >= 0 && cy < G && m.inv_x > 0.f) {
                    const int c = cy * G + cx;
                    e0[k] = __ldg(cs + c);
                    e1[k] = min(__ldg(cs + c + 1), cell_cap);
                }
            }
        }
        // Candidate list of one point (kMaxHitsPerPoint slots).  When it fills up -- many overlapping padded rectangles -- the
        // stored candidates are put through the exact test right away and only the real hits stay (flagged as decided), so
        // the limit applies to the boxes a point is INSIDE, as the name says, not to its rectangle candidates.
        auto push_cand = [&](int (&cand)[kMaxHitsPerPoint], int &nc, int b, float x, float y, float z) {
            if (nc == kMaxHitsPerPoint) {
                int m = 0;
                for (int q = 0; q < kMaxHitsPerPoint; ++q) {
                    const int cb = cand[q];
                    if ((cb & (1 << 30)) || crop_inside(x, y, z, pl + (cb & 0xFFFF) * 6)) cand[m++] = (cb & 0xFFFF) | (1 << 30);
                }
                nc = m;
                if (nc == kMaxHitsPerPoint) { if (crop_inside(x, y, z, pl + b * 6)) atomicExch(overflow, 2); return; }
            }
            cand[nc++] = b;
        };
        // the first kCropLook entries of every point's cell list: U * kCropLook independent loads, one round trip
        // (the per-point loop below used to walk its list one dependent load pair at a time)
        int bj[U][kCropLook];
#pragma unroll
        for (int k = 0; k < U; ++k)
#pragma unroll
            for (int j = 0; j < kCropLook; ++j) bj[k][j] = (e0[k] + j < e1[k]) ? __ldg(cb + e0[k] + j) : -1;
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int i = i0 + k * 32 + lane;
            int cand[kMaxHitsPerPoint];
            int nc = 0;
            if (i < w_hi) {
                if (px[k] != px[k] || py[k] != py[k] || pz[k] != pz[k]) {
                    // NaN never satisfies `sign >= 0`: the reference reports such a point inside every box.  Rare:
                    // decided in place, flagged so that the queue does not test it again.
                    for (int b = 0; b < B; ++b)
                        if (crop_inside(px[k], py[k], pz[k], pl + b * 6)) {
                            if (nc < kMaxHitsPerPoint) cand[nc++] = b | (1 << 30); else atomicExch(overflow, 2);
                        }
                } else {
                    // padded boxes of the prefetched entries: their loads are independent of each other
                    float2 lo[kCropLook], mid[kCropLook], hi[kCropLook];
#pragma unroll
                    for (int j = 0; j < kCropLook; ++j) {
                        const int b = max(bj[k][j], 0);
                        lo[j] = __ldg(bb + b * 3); mid[j] = __ldg(bb + b * 3 + 1); hi[j] = __ldg(bb + b * 3 + 2);
                    }
#pragma unroll
                    for (int j = 0; j < kCropLook; ++j)
                        if (bj[k][j] >= 0 && px[k] >= lo[j].x && py[k] >= lo[j].y && pz[k] >= mid[j].x && px[k] <= mid[j].y &&
                            py[k] <= hi[j].x && pz[k] <= hi[j].y) {
                            push_cand(cand, nc, bj[k][j], px[k], py[k], pz[k]);
                        }
                    for (int e = e0[k] + kCropLook; e < e1[k]; ++e) {               // longer lists (rare): the rest, one at a time
                        const int b = __ldg(cb + e);
                        const float2 lo = __ldg(bb + b * 3), mid = __ldg(bb + b * 3 + 1), hi = __ldg(bb + b * 3 + 2);
                        // aabb = [xmin ymin | zmin xmax | ymax zmax], padded: never rejects a point the exact test accepts
                        if (px[k] >= lo.x && py[k] >= lo.y && pz[k] >= mid.x && px[k] <= mid.y && py[k] <= hi.x && pz[k] <= hi.y) {
                            push_cand(cand, nc, b, px[k], py[k], pz[k]);
                        }
                    }
                }
            }
            const unsigned any = __ballot_sync(0xffffffffu, nc > 0);
            if (any == 0) continue;                                    // the common case: no candidate at all
            int incl = nc;
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            const int at = qhead + qcount + incl - nc;
#pragma unroll
            for (int j = 0; j < kMaxHitsPerPoint; ++j)
                if (j < nc) {
                    int a2 = at + j;
                    while (a2 >= kCropQueue) a2 -= kCropQueue;
                    queue[a2] = make_int2(i, cand[j]);
                }
            qcount += __shfl_sync(0xffffffffu, incl, 31);
            __syncwarp();
            while (qcount >= 32) drain(32);
        }
    }
    __syncwarp();
    while (qcount > 0) drain(min(qcount, 32));
    if (wcount > stage_cap) wcount = stage_cap;
    if (lane == 0) warp_total[wid] = wcount;
    // ---- rank of every hit among the hits of the same box in this chunk (point order).
    //      Parallel form: every warp first ranks its own (ordered) list against its own per-box counters -- 32 hits at a
    //      time, equal boxes inside a group ranked by lane -- then adds the counts of the warps before it.
    if (par_rank) {
        int32_t *mine = s_wcnt + wid * rank_boxes;
        for (int h0 = 0; h0 < wcount; h0 += 32) {
            const int h = h0 + lane;
            const bool act = h < wcount;
            const int2 e = act ? stage[h] : make_int2(0, -1 - lane);
            const unsigned same = __match_any_sync(0xffffffffu, e.y);
            if (act) stage[h] = make_int2(e.x, e.y | ((mine[e.y] + __popc(same & ((1u << lane) - 1u))) << 16));
            __syncwarp();
            if (act && (same >> lane) == 1u) mine[e.y] += __popc(same);          // highest lane of each group updates
            __syncwarp();
        }
    }
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < kCropWarps; ++w) { if (w < wid) before += warp_total[w]; total += warp_total[w]; }
    if (before + wcount > hit_cap) atomicExch(overflow, 3);
    total = min(total, hit_cap);
    if (threadIdx.x == 0) n_hits[blockIdx.x] = total;
    if (par_rank) {
        for (int h = lane; h < wcount; h += 32) {
            const int2 e = stage[h];
            const int box = e.y & 0xFFFF;
            int prefix = 0;
            for (int w = 0; w < wid; ++w) prefix += s_wcnt[w * rank_boxes + box];
            if (before + h < hit_cap) my_hits[before + h] = make_int2(e.x, e.y + (prefix << 16));
        }
        for (int b = threadIdx.x; b < B; b += blockDim.x) {
            int sum = 0;
#pragma unroll
            for (int w = 0; w < kCropWarps; ++w) sum += s_wcnt[w * rank_boxes + b];
            chunk_box_count[(int64_t)blockIdx.x * max_boxes + b] = sum;
        }
        return;
    }
    // Serial form (frames with more boxes than the per-warp counter table holds): one warp walks the whole ordered list.
    if (wid == 0) {
        int g0 = 0;                                                    // global position of the current warp list
        for (int w = 0; w < kCropWarps; ++w) {
            const int cnt = warp_total[w];
            int2 *l = s_stage + (size_t)w * stage_cap;
            for (int h0 = 0; h0 < cnt; h0 += 32) {
                const int h = h0 + lane;
                const bool act = h < cnt;
                const int2 e = act ? l[h] : make_int2(0, -1 - lane);
                const unsigned same = __match_any_sync(0xffffffffu, e.y);
                if (act) {
                    const int rank = s_box_cnt[e.y] + __popc(same & ((1u << lane) - 1u));
                    if (g0 + h < hit_cap) my_hits[g0 + h] = make_int2(e.x, e.y | (rank << 16));
                }
                __syncwarp();
                if (act && (same >> lane) == 1u) s_box_cnt[e.y] += __popc(same);  // highest lane of each group updates
                __syncwarp();
            }
            g0 += cnt;
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += blockDim.x) chunk_box_count[(int64_t)blockIdx.x * max_boxes + b] = s_box_cnt[b];
}

// per frame: for every box an exclusive scan over the frame's chunks; box totals.  One WARP per box: the lanes
// take 32 consecutive chunks at a time (shuffle scan), so a frame's ~90 chunks cost three round trips, not ninety.
__global__ void crop_scan_kernel(const int64_t *__restrict__ box_off, const int64_t *__restrict__ frame_chunk_off,
                                 int32_t *__restrict__ chunk_box_count, int max_boxes, int32_t *__restrict__ box_total)
{
    const int f = blockIdx.x;
    const int64_t b0 = box_off[f];
    const int B = (int)(box_off[f + 1] - b0);
    const int64_t c0 = frame_chunk_off[f], c1 = frame_chunk_off[f + 1];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (int b = blockIdx.y * n_warps + wid; b < B; b += gridDim.y * n_warps) {
        int run = 0;
        for (int64_t cb = c0; cb < c1; cb += 32) {
            const int64_t c = cb + lane;
            const int v = c < c1 ? chunk_box_count[c * max_boxes + b] : 0;
            int incl = v;
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            if (c < c1) chunk_box_count[c * max_boxes + b] = run + incl - v;
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) box_total[b0 + b] = run;
    }
}

// global exclusive offsets over all boxes of all frames (single CTA)
__global__ void __launch_bounds__(1024)
crop_offsets_kernel(const int32_t *__restrict__ box_total, int64_t n_boxes, int64_t *__restrict__ offsets)
{
    __shared__ long long part[1024];
    const int64_t per = (n_boxes + blockDim.x - 1) / blockDim.x;
    const int64_t lo = (int64_t)threadIdx.x * per, hi = min(lo + per, n_boxes);
    long long sum = 0;
    for (int64_t i = lo; i < hi; ++i) sum += box_total[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long run = 0;
        for (int t = 0; t < (int)blockDim.x; ++t) { const long long v = part[t]; part[t] = run; run += v; }
        offsets[n_boxes] = run;
    }
    __syncthreads();
    long long run = part[threadIdx.x];
    for (int64_t i = lo; i < hi; ++i) { offsets[i] = run; run += box_total[i]; }
}

__global__ void __launch_bounds__(kCropThreads)
crop_fill_kernel(const float *__restrict__ points, int64_t pt_stride, const int64_t *__restrict__ pt_off,
                 const int64_t *__restrict__ box_off, const CropChunk *__restrict__ chunks, const int2 *__restrict__ hits,
                 int hit_cap, const int32_t *__restrict__ n_hits, const int32_t *__restrict__ chunk_box_count, int max_boxes,
                 const int64_t *__restrict__ offsets, const double *__restrict__ poses, int64_t capacity,
                 int32_t *__restrict__ out_idx, float *__restrict__ out_xyz, double *__restrict__ out_xyz_global,
                 int32_t *__restrict__ overflow)
{
    const CropChunk ck = chunks[blockIdx.x];
    const int f = ck.frame;
    const int64_t b0 = box_off[f];
    const int total = n_hits[blockIdx.x];
    const int2 *my_hits = hits + (int64_t)blockIdx.x * hit_cap;
    const double *P = poses ? poses + (int64_t)f * 16 : nullptr;
    for (int h = threadIdx.x; h < total; h += blockDim.x) {
        const int2 hr = my_hits[h];
        const int3 hv = make_int3(hr.x, hr.y & 0xFFFF, hr.y >> 16);
        const int64_t dst = offsets[b0 + hv.y] + chunk_box_count[(int64_t)blockIdx.x * max_boxes + hv.y] + hv.z;
        if (dst >= capacity) { atomicExch(overflow, 4); continue; }
        const float *p = points + (pt_off[f] + hv.x) * pt_stride;
        const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
        out_idx[dst] = hv.x;
        if (out_xyz) { out_xyz[dst * 3] = x; out_xyz[dst * 3 + 1] = y; out_xyz[dst * 3 + 2] = z; }
        if (out_xyz_global && P) {
            const double dx = x, dy = y, dz = z;
#pragma unroll
            for (int r = 0; r < 3; ++r)
                out_xyz_global[dst * 3 + r] = ((P[r * 4] * dx + P[r * 4 + 1] * dy) + P[r * 4 + 2] * dz) + P[r * 4 + 3];
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

extern "C" int al3d_crop_build_grid(const float *aabb, const int64_t *box_off, int n_frames, int G, float *grid_meta,
                                    int32_t *cell_start, int32_t *cell_boxes, int cell_cap, int32_t *overflow, void *stream)
{
    AL3D_CHECK_ARG(aabb && box_off && grid_meta && cell_start && cell_boxes && overflow, "al3d_crop_build_grid: null pointer");
    AL3D_CHECK_ARG(G >= 1 && G <= 64, "al3d_crop_build_grid: G=%d not in [1,64]", G);
    if (n_frames <= 0) return 0;
    crop_grid_kernel<<<n_frames, kCropThreads, (size_t)2 * G * G * sizeof(int32_t), (cudaStream_t)stream>>>(
        aabb, box_off, G, reinterpret_cast<CropGridMeta *>(grid_meta), cell_start, cell_boxes, cell_cap, overflow);
    AL3D_CHECK_LAUNCH("crop_grid_kernel");
    return 0;
}

extern "C" int al3d_crop_hits(const float *points, int64_t pt_stride, const int64_t *pt_off, const float *planes,
                              const float *aabb, const int64_t *box_off, int G, const float *grid_meta, const int32_t *cell_start,
                              const int32_t *cell_boxes, int cell_cap, const int32_t *chunks, int n_chunks, void *hits,
                              int hit_cap, int32_t *n_hits, int32_t *chunk_box_count, int max_boxes, int32_t *overflow, void *stream)
{
    AL3D_CHECK_ARG(points && pt_off && planes && aabb && box_off && grid_meta && cell_start && cell_boxes && chunks && hits && n_hits &&
                   chunk_box_count && overflow, "al3d_crop_hits: null pointer");
    AL3D_CHECK_ARG(pt_stride >= 3, "al3d_crop_hits: pt_stride=%lld", (long long)pt_stride);
    AL3D_CHECK_ARG(max_boxes >= 1 && max_boxes <= 12288, "al3d_crop_hits: max_boxes=%d not in [1,12288]", max_boxes);
    AL3D_CHECK_ARG(hit_cap >= 1 && hit_cap <= 16384, "al3d_crop_hits: hit_cap=%d not in [1,16384]", hit_cap);
    if (n_chunks <= 0) return 0;
    const int rank_boxes = max_boxes <= 512 ? max_boxes : 0;          // per-warp box counters for the parallel ranking, if modest
    const size_t smem = (size_t)((max_boxes + 1) & ~1) * sizeof(int32_t) + (size_t)kCropWarps * kCropQueue * sizeof(int2) +
                        (size_t)kCropWarps * std::min(hit_cap, kCropWarpPts * kMaxHitsPerPoint) * sizeof(int2) +
                        (size_t)kCropWarps * rank_boxes * sizeof(int32_t);
    AL3D_CHECK_ARG(smem <= 200 * 1024, "al3d_crop_hits: hit_cap=%d x max_boxes=%d needs too much shared memory", hit_cap, max_boxes);
    if (smem > 48 * 1024) AL3D_CHECK_CUDA(cudaFuncSetAttribute(crop_hits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    crop_hits_kernel<<<n_chunks, kCropThreads, smem, (cudaStream_t)stream>>>(
        points, pt_stride, pt_off, planes, aabb, box_off, G, reinterpret_cast<const CropGridMeta *>(grid_meta), cell_start, cell_boxes,
        cell_cap, reinterpret_cast<const CropChunk *>(chunks), reinterpret_cast<int2 *>(hits), hit_cap, n_hits, chunk_box_count,
        max_boxes, rank_boxes, overflow);
    AL3D_CHECK_LAUNCH("crop_hits_kernel");
    return 0;
}

extern "C" int al3d_crop_scan(const int64_t *box_off, const int64_t *frame_chunk_off, int n_frames, int64_t n_boxes,
                              int32_t *chunk_box_count, int max_boxes, int32_t *box_total, int64_t *offsets, void *stream)
{
    AL3D_CHECK_ARG(box_off && frame_chunk_off && chunk_box_count && box_total && offsets, "al3d_crop_scan: null pointer");
    if (n_frames > 0) {
        crop_scan_kernel<<<dim3(n_frames, 4), 256, 0, (cudaStream_t)stream>>>(box_off, frame_chunk_off, chunk_box_count, max_boxes, box_total);
        AL3D_CHECK_LAUNCH("crop_scan_kernel");
    }
    crop_offsets_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(box_total, n_boxes, offsets);
    AL3D_CHECK_LAUNCH("crop_offsets_kernel");
    return 0;
}

extern "C" int al3d_crop_fill(const float *points, int64_t pt_stride, const int64_t *pt_off, const int64_t *box_off,
                              const int32_t *chunks, int n_chunks, const void *hits, int hit_cap, const int32_t *n_hits,
                              const int32_t *chunk_box_count, int max_boxes, const int64_t *offsets, const double *poses,
                              int64_t capacity, int32_t *out_idx, float *out_xyz, double *out_xyz_global, int32_t *overflow,
                              void *stream)
{
    AL3D_CHECK_ARG(points && pt_off && box_off && chunks && hits && n_hits && chunk_box_count && offsets && out_idx && overflow,
                   "al3d_crop_fill: null pointer");
    if (n_chunks <= 0) return 0;
    crop_fill_kernel<<<n_chunks, kCropThreads, 0, (cudaStream_t)stream>>>(
        points, pt_stride, pt_off, box_off, reinterpret_cast<const CropChunk *>(chunks), reinterpret_cast<const int2 *>(hits), hit_cap,
        n_hits, chunk_box_count, max_boxes, offsets, poses, capacity, out_idx, out_xyz, out_xyz_global, overflow);
    AL3D_CHECK_LAUNCH("crop_fill_kernel");
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
