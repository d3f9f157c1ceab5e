// mapgrid.cu -- device build of the uniform-grid map index (one per local map per frame).
//
// Replaces kdtree{Corner,Surf}FromMap->setInputCloud (mapOptmization.h:1413-1414), which the
// reference re-runs every frame on one thread.  Here: bbox reduce -> cell id per point +
// histogram (atomics) -> exclusive scan over cells -> scatter into cell-contiguous order
// (counting sort).  Order inside a cell is arbitrary; exactness of the k-NN does not depend on
// it because candidates are ranked by (d^2, original index).
// HBM-bound: reads each map point three times (16 B: bbox, count, scatter) and writes it once; the cell array costs
// 16 B per cell (zero, sum, scan in place).  blockIdx.y = segment (2 maps x frame slots).
#include "internal.cuh"
#include "mapgrid.cuh"

namespace {

constexpr int TPB = 256;
constexpr int TILE = 2048;
constexpr int IPT = TILE / TPB;
constexpr int SCAN_TILE = 4096;   // cells per CTA in the cell scan

__device__ inline unsigned warp_min_u(unsigned v) { for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }
__device__ inline unsigned warp_max_u(unsigned v) { for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }

__global__ void grid_init(const GridSeg* segs) {
    const GridSeg& s = segs[blockIdx.x];
    if (threadIdx.x < 3) s.bbox[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6) s.bbox[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(TPB) grid_minmax(const GridSeg* segs) {
    const GridSeg& s = segs[blockIdx.y];
    int n = min(*s.n, s.cap);
    int base = blockIdx.x * TILE;
    if (base >= n) return;
    unsigned mn[3] = { 0xffffffffu, 0xffffffffu, 0xffffffffu }, mx[3] = { 0u, 0u, 0u };
    for (int k = 0; k < IPT; k++) {
        int i = base + k * TPB + threadIdx.x;
        if (i < n) {
            float4 p = s.pts[i];
            if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) continue;  // KdTreeFLANN::setInputCloud indexes finite points only (Appendix B-2); such a point lands in a clamped cell and never ranks
            unsigned ex = f2ord(p.x), ey = f2ord(p.y), ez = f2ord(p.z);
            mn[0] = min(mn[0], ex); mn[1] = min(mn[1], ey); mn[2] = min(mn[2], ez);
            mx[0] = max(mx[0], ex); mx[1] = max(mx[1], ey); mx[2] = max(mx[2], ez);
        }
    }
    __shared__ unsigned sm[6][TPB / 32];
    for (int c = 0; c < 3; c++) { mn[c] = warp_min_u(mn[c]); mx[c] = warp_max_u(mx[c]); }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) for (int c = 0; c < 3; c++) { sm[c][w] = mn[c]; sm[3 + c][w] = mx[c]; }
    __syncthreads();
    if (threadIdx.x < 6) {
        unsigned v = sm[threadIdx.x][0];
        for (int k = 1; k < TPB / 32; k++) v = threadIdx.x < 3 ? min(v, sm[threadIdx.x][k]) : max(v, sm[threadIdx.x][k]);
        if (threadIdx.x < 3) atomicMin(&s.bbox[threadIdx.x], v); else atomicMax(&s.bbox[threadIdx.x], v);
    }
}

// choose the cell edge (>= requested, doubled until the dense grid fits cells_cap) and zero the counters
__global__ void __launch_bounds__(TPB) grid_setup_zero(const GridSeg* segs) {
    const GridSeg& s = segs[blockIdx.y];
    __shared__ GridDesc gd;
    if (threadIdx.x == 0) {
        int n = min(*s.n, s.cap); if (n < 0) n = 0;
        GridDesc g;
        g.n = n;
        float mn[3] = { 0, 0, 0 }, mx[3] = { 0, 0, 0 };
        if (n > 0 && s.bbox[0] <= s.bbox[3]) for (int c = 0; c < 3; c++) { mn[c] = ord2f(s.bbox[c]); mx[c] = ord2f(s.bbox[3 + c]); }   // (no finite point: a one-cell grid at the origin)
        float h = s.h0;
        while (true) {
            float inv = 1.0f / h;
            long long dx = (long long)((mx[0] - mn[0]) * inv) + 1, dy = (long long)((mx[1] - mn[1]) * inv) + 1, dz = (long long)((mx[2] - mn[2]) * inv) + 1;
            if (dx < 2000000 && dy < 2000000 && dz < 2000000 && dx * dy * dz <= (long long)s.cells_cap) {
                g.dx = (int)dx; g.dy = (int)dy; g.dz = (int)dz; g.h = h; g.inv_h = inv; break;
            }
            h *= 2.0f;
        }
        g.ox = mn[0]; g.oy = mn[1]; g.oz = mn[2];
        g.ncells = g.dx * g.dy * g.dz;
        int rmax = 1; while ((float)rmax * g.h * 0.9995f < 1.0f) rmax++;
        g.rmax = rmax; g.pad = 0;
        gd = g;
        if (blockIdx.x == 0) *s.desc = g;
    }
    __syncthreads();
    const int nc = gd.ncells;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c <= nc; c += gridDim.x * blockDim.x) s.cell_start[c] = 0;
}

__device__ inline int cell_of_point(const GridDesc& g, float4 p) {
    int cx = (int)floorf((p.x - g.ox) * g.inv_h), cy = (int)floorf((p.y - g.oy) * g.inv_h), cz = (int)floorf((p.z - g.oz) * g.inv_h);
    cx = min(max(cx, 0), g.dx - 1); cy = min(max(cy, 0), g.dy - 1); cz = min(max(cz, 0), g.dz - 1);
    return (cz * g.dy + cy) * g.dx + cx;
}

// histogram of the cells; the value the atomic returns is the point's rank inside its cell, which makes the scatter
// below atomic-free (position = cell_start[cell] + rank)
__global__ void __launch_bounds__(TPB) grid_count(const GridSeg* segs) {
    const GridSeg& s = segs[blockIdx.y];
    const GridDesc g = *s.desc;
    int base = blockIdx.x * TILE;
    if (base >= g.n) return;
    // all loads, then all atomics, then all stores: eight independent L2 round trips in flight per thread instead of one
    int c[IPT], r[IPT];
    #pragma unroll
    for (int k = 0; k < IPT; k++) {
        const int i = base + k * TPB + threadIdx.x;
        c[k] = i < g.n ? cell_of_point(g, s.pts[i]) : -1;
    }
    #pragma unroll
    for (int k = 0; k < IPT; k++) r[k] = c[k] >= 0 ? atomicAdd(&s.cell_start[c[k]], 1) : 0;
    #pragma unroll
    for (int k = 0; k < IPT; k++) {
        const int i = base + k * TPB + threadIdx.x;
        if (c[k] >= 0) s.cell_of[i] = r[k];
    }
}

// exclusive scan over the cells of every segment in two passes (12 B of traffic per cell): per-tile sums, then every
// tile adds up the sums of the tiles before it (at most cells_cap / SCAN_TILE values) and scans itself in place
__device__ inline int block_sum_1024(int v, int* ws) {
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (l == 0) ws[w] = v;
    __syncthreads();
    int t = ws[l];
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    __syncthreads();
    return t;
}
// both scan kernels walk the tiles of their segment with a small fixed number of CTAs: the cell capacity is a worst case
// (1 M cells) and a launch sized for it would consist mostly of CTAs that find nothing to do
constexpr int SCAN_CTAS = 16;
__global__ void __launch_bounds__(1024) grid_tile_sums(const GridSeg* segs) {
    const GridSeg& s = segs[blockIdx.y];
    const int nc = s.desc->ncells + 1;
    __shared__ int ws[32];
    for (int tile = blockIdx.x; tile * SCAN_TILE < nc; tile += gridDim.x) {
        const int base = tile * SCAN_TILE;
        int sum = 0;
        const int c0 = base + threadIdx.x * 4;
        if (c0 + 3 < nc) { const int4 v = *reinterpret_cast<const int4*>(s.cell_start + c0); sum = v.x + v.y + v.z + v.w; }
        else for (int k = 0; k < 4; k++) if (c0 + k < nc) sum += s.cell_start[c0 + k];
        sum = block_sum_1024(sum, ws);
        if (threadIdx.x == 0) s.tile_sum[tile] = sum;
    }
}
__global__ void __launch_bounds__(1024) grid_scan_apply(const GridSeg* segs) {
    const GridSeg& s = segs[blockIdx.y];
    const int nc = s.desc->ncells + 1;
    __shared__ int ws[32];
    for (int tile = blockIdx.x; tile * SCAN_TILE < nc; tile += gridDim.x) {
        const int base = tile * SCAN_TILE;
        int before = 0;
        for (int t = threadIdx.x; t < tile; t += 1024) before += s.tile_sum[t];
        before = block_sum_1024(before, ws);
        int v[4] = { 0, 0, 0, 0 }, sum = 0;
        const int c0 = base + threadIdx.x * 4;
        if (c0 + 3 < nc) { const int4 q = *reinterpret_cast<const int4*>(s.cell_start + c0); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
        else for (int k = 0; k < 4; k++) if (c0 + k < nc) v[k] = s.cell_start[c0 + k];
        sum = v[0] + v[1] + v[2] + v[3];
        int incl = sum, l = threadIdx.x & 31, w = threadIdx.x >> 5;
        for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, o); if (l >= o) incl += u; }
        if (l == 31) ws[w] = incl;
        __syncthreads();
        if (w == 0) {
            int a = ws[l], ia = a;
            for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, ia, o); if (l >= o) ia += u; }
            ws[l] = ia - a;
        }
        __syncthreads();
        int run = before + ws[w] + incl - sum;
        int4 o4; o4.x = run; o4.y = run + v[0]; o4.z = o4.y + v[1]; o4.w = o4.z + v[2];
        if (c0 + 3 < nc) *reinterpret_cast<int4*>(s.cell_start + c0) = o4;
        else { const int o[4] = { o4.x, o4.y, o4.z, o4.w }; for (int k = 0; k < 4; k++) if (c0 + k < nc) s.cell_start[c0 + k] = o[k]; }
        __syncthreads();                                  // ws is reused by the next tile
    }
}

__global__ void __launch_bounds__(TPB) grid_scatter(const GridSeg* segs) {
    const GridSeg& s = segs[blockIdx.y];
    const GridDesc g = *s.desc;
    int base = blockIdx.x * TILE;
    if (base >= g.n) return;
    float4 p[IPT]; int pos[IPT];
    #pragma unroll
    for (int k = 0; k < IPT; k++) {
        const int i = base + k * TPB + threadIdx.x;
        if (i < g.n) { p[k] = s.pts[i]; pos[k] = s.cell_of[i]; } else pos[k] = -1;
    }
    #pragma unroll
    for (int k = 0; k < IPT; k++) if (pos[k] >= 0) pos[k] += __ldg(s.cell_start + cell_of_point(g, p[k]));
    #pragma unroll
    for (int k = 0; k < IPT; k++) {
        const int i = base + k * TPB + threadIdx.x;
        if (pos[k] >= 0) s.sorted[pos[k]] = make_float4(p[k].x, p[k].y, p[k].z, __int_as_float(i));
    }
}

// stand-alone query kernel behind fbpr_knn5 (parity tests of the index itself): the same routine the LM kernel
// uses; rad0 = first cube radius in cells (the LM kernel derives it per point from the previous iteration)
__global__ void __launch_bounds__(128) knn5_query(const GridSeg* segs, const float* __restrict__ q, int nq, int rad0, int* idx, float* d2) {
    __shared__ GridDesc gd[2];
    const GridSeg& s = segs[0];
    if (threadIdx.x == 0) { gd[0] = *s.desc; gd[1] = gd[0]; }
    __syncthreads();
    KnnMaps M; M.gd = gd; M.cell_start[0] = M.cell_start[1] = s.cell_start; M.pts[0] = M.pts[1] = s.sorted;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < nq && gd[0].n >= 5;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) { qx = q[3 * i]; qy = q[3 * i + 1]; qz = q[3 * i + 2]; }
    ThreadKnn5 r;
    #pragma unroll
    for (int k = 0; k < 5; k++) r.key[k] = ~0ull;
    warp_knn5(M, 0, qx, qy, qz, (float)rad0 * gd[0].h, active, r, nullptr, nullptr);     // rad0 cells -> metres
    const bool ok = active && knn_d5(r) < 1.0f;
    if (i >= nq) return;
    #pragma unroll
    for (int k = 0; k < 5; k++) {
        idx[5 * i + k] = ok ? knn_index(r, k) : -1;
        d2[5 * i + k] = r.key[k] == ~0ull ? 3.0e38f : __uint_as_float((unsigned)(r.key[k] >> 32));
    }
}

}  // namespace

int fbpr_launch_grid_build(const GridSeg* d_segs, int nsegs, int max_n, int cells_cap, cudaStream_t st, long long* launches) {
    if (nsegs <= 0) return 0;
    int tiles = (max_n + TILE - 1) / TILE; if (tiles < 1) tiles = 1;
    int stiles = (cells_cap + 1 + SCAN_TILE - 1) / SCAN_TILE;
    // grid-stride zeroing / scanning with about two waves of CTAs over all segments (every zeroing CTA repeats the serial choice of
    // the cell size, and the cell capacity is a worst case: a launch sized for it is mostly CTAs that find nothing to do)
    int zb = (cells_cap + TPB * 8) / (TPB * 8);
    const int zmax = 24 > 2368 / nsegs ? 24 : 2368 / nsegs;
    if (zb > zmax) zb = zmax;
    const int smax = SCAN_CTAS > 592 / nsegs ? SCAN_CTAS : 592 / nsegs;
    dim3 g(tiles, nsegs), gs(stiles < smax ? stiles : smax, nsegs);
    grid_init<<<nsegs, 32, 0, st>>>(d_segs);
    grid_minmax<<<g, TPB, 0, st>>>(d_segs);
    grid_setup_zero<<<dim3(zb, nsegs), TPB, 0, st>>>(d_segs);
    grid_count<<<g, TPB, 0, st>>>(d_segs);
    grid_tile_sums<<<gs, 1024, 0, st>>>(d_segs);
    grid_scan_apply<<<gs, 1024, 0, st>>>(d_segs);
    grid_scatter<<<g, TPB, 0, st>>>(d_segs);
    if (launches) *launches += 7;
    return fbpr_launch_ok("map index (grid_*)");
}

int fbpr_launch_knn5(const GridSeg* d_seg, const float* d_q, int nq, int* d_idx, float* d_d2, int rad0, cudaStream_t st, long long* launches) {
    if (nq <= 0) return 0;
    knn5_query<<<(nq + 127) / 128, 128, 0, st>>>(d_seg, d_q, nq, rad0, d_idx, d_d2);
    if (launches) *launches += 1;
    return fbpr_launch_ok("knn5_query");
}
