// mapgrid.cuh -- device-side exact 5-NN query on the uniform-grid map index.
//
// Replaces pcl::KdTreeFLANN::nearestKSearch(p, 5, ...) (mapOptmization.h:1020, :1143; FLANN
// KDTreeSingleIndex, L2_Simple, exact, sorted -- SURVEY.md Appendix B-2).  Only neighbours
// with d^2 < 1.0 matter to the caller (:1027, :1154), so the query returns the exact 5-NN inside
// the 1 m ball or "reject".  Distance is ((dx*dx)+(dy*dy))+(dz*dz) in f32 without FMA (the
// library is built with -fmad=false) so ordering and the < 1.0 gate match the reference; ties
// are broken by the original point index.
#pragma once
#include "internal.cuh"

struct Knn5 {
    float d[5];
    int id[5];     // original map index
    int pos[5];    // position in the cell-sorted array (to re-fetch coordinates)
};

__device__ __forceinline__ bool knn_better(float dd, int ii, float d, int i) { return dd < d || (dd == d && ii < i); }

__device__ __forceinline__ void knn_offer(Knn5& r, float dd, int ii, int pp) {
    if (!knn_better(dd, ii, r.d[4], r.id[4])) return;
    // insertion keeping (d, id) ascending; fully unrolled so the set stays in registers
    #pragma unroll
    for (int k = 4; k >= 0; k--) {
        if (k > 0 && knn_better(dd, ii, r.d[k - 1], r.id[k - 1])) { r.d[k] = r.d[k - 1]; r.id[k] = r.id[k - 1]; r.pos[k] = r.pos[k - 1]; }
        else { r.d[k] = dd; r.id[k] = ii; r.pos[k] = pp; break; }
    }
}

// Returns true when 5 neighbours with d^2 < 1.0 exist (then r is exact and sorted).
// When false, r holds whatever was found inside the covered ball (exact for every entry < 1.0).
__device__ inline bool grid_knn5(const GridDesc& g, const int* __restrict__ cell_start, const int* __restrict__ cell_end,
                                 const float4* __restrict__ pts, float qx, float qy, float qz, Knn5& r) {
    const int cx = (int)floorf((qx - g.ox) * g.inv_h);
    const int cy = (int)floorf((qy - g.oy) * g.inv_h);
    const int cz = (int)floorf((qz - g.oz) * g.inv_h);
    int rad = 1;
    while (true) {
        #pragma unroll
        for (int k = 0; k < 5; k++) { r.d[k] = 3.0e38f; r.id[k] = 0x7fffffff; r.pos[k] = -1; }
        const int x0 = max(cx - rad, 0), x1 = min(cx + rad, g.dx - 1);
        const int y0 = max(cy - rad, 0), y1 = min(cy + rad, g.dy - 1);
        const int z0 = max(cz - rad, 0), z1 = min(cz + rad, g.dz - 1);
        if (x0 <= x1) {
            for (int z = z0; z <= z1; z++) {
                for (int y = y0; y <= y1; y++) {
                    const int row = (z * g.dy + y) * g.dx;
                    const int a = cell_start[row + x0], b = cell_end[row + x1];   // x-adjacent cells are contiguous
                    for (int p = a; p < b; p++) {
                        const float4 m = pts[p];
                        const float ddx = qx - m.x, ddy = qy - m.y, ddz = qz - m.z;
                        float dd = ddx * ddx; dd += ddy * ddy; dd += ddz * ddz;
                        knn_offer(r, dd, __float_as_int(m.w), p);
                    }
                }
            }
        }
        // every point closer than rad*h (minus a rounding guard) has been seen
        const float guard = (float)rad * g.h * 0.9995f;
        if (r.d[4] < guard * guard) break;          // exact 5-NN found inside the covered ball
        if (rad >= g.rmax) break;                   // the whole 1 m ball is covered
        rad = min(rad * 2, g.rmax);
    }
    return r.d[4] < 1.0f;
}


// ---------------------------------------------------------------------------------------------
// Warp-cooperative form of the same query: ONE WARP per query point.
//   - lanes fetch the (start, end) bounds of up to 32 cell rows at once (x-adjacent cells are
//     contiguous in the cell-sorted array, so a row of the search cube is one range);
//   - the concatenated candidates of those rows are walked 32 at a time: consecutive lanes read
//     consecutive float4 points (coalesced 512 B), compute d^2 and keep a private sorted top-5;
//   - five rounds of warp-min over packed (d^2 bits, index) keys merge the 32 private lists.
// Every global load of a batch is independent, so a query costs two or three memory round trips
// instead of one per candidate.  Ranking key = (d^2 as ordered u32) << 32 | original index: the
// (d^2, index) total order of the oracle.  Result is replicated in all lanes.
struct WarpKnn5 {
    unsigned long long key[5];   // ascending; 0xffff... = empty
    int pos[5];                  // position in the cell-sorted array
};

__device__ __forceinline__ void lane_insert5(unsigned long long* k, int* ps, unsigned long long key, int pos) {
    if (key >= k[4]) return;
    #pragma unroll
    for (int i = 4; i >= 0; i--) {
        if (i > 0 && key < k[i - 1]) { k[i] = k[i - 1]; ps[i] = ps[i - 1]; }
        else { k[i] = key; ps[i] = pos; break; }
    }
}

__device__ inline bool warp_knn5(const GridDesc& g, const int* __restrict__ cell_start, const int* __restrict__ cell_end,
                                 const float4* __restrict__ pts, float qx, float qy, float qz, WarpKnn5& out) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int cx = (int)floorf((qx - g.ox) * g.inv_h);
    const int cy = (int)floorf((qy - g.oy) * g.inv_h);
    const int cz = (int)floorf((qz - g.oz) * g.inv_h);
    int rad = 1;
    while (true) {
        unsigned long long k[5]; int ps[5];
        #pragma unroll
        for (int i = 0; i < 5; i++) { k[i] = ~0ull; ps[i] = -1; }
        const int x0 = max(cx - rad, 0), x1 = min(cx + rad, g.dx - 1);
        const int y0 = max(cy - rad, 0), y1 = min(cy + rad, g.dy - 1);
        const int z0 = max(cz - rad, 0), z1 = min(cz + rad, g.dz - 1);
        const int ny = y1 - y0 + 1, nz = z1 - z0 + 1;
        const int nrows = (x0 <= x1 && ny > 0 && nz > 0) ? ny * nz : 0;
        for (int rbase = 0; rbase < nrows; rbase += 32) {
            const int rr = rbase + lane;
            int a = 0, len = 0;
            if (rr < nrows) {
                const int zz = z0 + rr / ny, yy = y0 + rr % ny;
                const int row = (zz * g.dy + yy) * g.dx;
                a = cell_start[row + x0];
                len = cell_end[row + x1] - a;
            }
            int incl = len;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
            const int total = __shfl_sync(FULL, incl, 31);
            const int excl = incl - len;
            // four batches of 32 candidates per trip: all 128 loads are in flight before any is consumed
            for (int t0 = 0; t0 < total; t0 += 128) {
                float4 m[4]; int p[4]; bool v[4];
                #pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int t = t0 + u * 32 + lane;
                    int j = 0;
                    #pragma unroll
                    for (int s = 16; s >= 1; s >>= 1) {
                        const int cand = j + s;
                        const int e = __shfl_sync(FULL, excl, cand & 31);
                        if (e <= t) j = cand;
                    }
                    const int aj = __shfl_sync(FULL, a, j), ej = __shfl_sync(FULL, excl, j);
                    v[u] = t < total;
                    p[u] = aj + (t - ej);
                    if (v[u]) m[u] = pts[p[u]];
                }
                #pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (v[u]) {
                        const float ddx = qx - m[u].x, ddy = qy - m[u].y, ddz = qz - m[u].z;
                        float dd = ddx * ddx; dd += ddy * ddy; dd += ddz * ddz;
                        lane_insert5(k, ps, ((unsigned long long)__float_as_uint(dd) << 32) | (unsigned)__float_as_int(m[u].w), p[u]);
                    }
                }
            }
        }
        // merge the 32 private lists: five rounds of (hardware warp-min on d^2 bits, then on the index among the ties)
        #pragma unroll
        for (int r = 0; r < 5; r++) {
            const unsigned hi = (unsigned)(k[0] >> 32), lo = (unsigned)k[0];
            const unsigned mhi = __reduce_min_sync(FULL, hi);
            const unsigned mlo = __reduce_min_sync(FULL, hi == mhi ? lo : 0xffffffffu);
            const unsigned who = __ballot_sync(FULL, hi == mhi && lo == mlo);
            const int src = __ffs(who) - 1;
            out.key[r] = ((unsigned long long)mhi << 32) | mlo;
            out.pos[r] = __shfl_sync(FULL, ps[0], src);
            if (lane == src) {
                k[0] = k[1]; k[1] = k[2]; k[2] = k[3]; k[3] = k[4]; k[4] = ~0ull;
                ps[0] = ps[1]; ps[1] = ps[2]; ps[2] = ps[3]; ps[3] = ps[4]; ps[4] = -1;
            }
        }
        const float d5 = __uint_as_float((unsigned)(out.key[4] >> 32));
        const bool have5 = out.key[4] != ~0ull;
        const float guard = (float)rad * g.h * 0.9995f;
        if (have5 && d5 < guard * guard) break;      // exact 5-NN found inside the covered ball
        if (rad >= g.rmax) break;                    // the whole 1 m ball is covered
        rad = min(rad * 2, g.rmax);
    }
    return out.key[4] != ~0ull && __uint_as_float((unsigned)(out.key[4] >> 32)) < 1.0f;
}


// ---------------------------------------------------------------------------------------------
// Thread-per-query form with memory-level parallelism: ONE THREAD per query point.
// The bounds of up to 9 cell rows (18 loads) are issued together, then the concatenated candidates
// are fetched eight at a time (8 independent 16-byte loads in flight) before any is consumed, so a
// query costs ~2 + T/8 memory round trips instead of one per candidate, and a warp keeps 32 queries
// in flight.  Used where the map is sparse around the query (planar points); dense linear features
// go through warp_knn5.  Every array index below is a compile-time constant after unrolling.
struct ThreadKnn5 {
    unsigned long long key[5];
    int pos[5];
};

__device__ inline bool thread_knn5(const GridDesc& g, const int* __restrict__ cell_start, const int* __restrict__ cell_end,
                                   const float4* __restrict__ pts, float qx, float qy, float qz, ThreadKnn5& out) {
    constexpr int RC = 9;        // rows per chunk
    constexpr int GB = 8;        // candidates per batch
    const int cx = (int)floorf((qx - g.ox) * g.inv_h);
    const int cy = (int)floorf((qy - g.oy) * g.inv_h);
    const int cz = (int)floorf((qz - g.oz) * g.inv_h);
    int rad = 1;
    while (true) {
        #pragma unroll
        for (int i = 0; i < 5; i++) { out.key[i] = ~0ull; out.pos[i] = -1; }
        const int x0 = max(cx - rad, 0), x1 = min(cx + rad, g.dx - 1);
        const int y0 = max(cy - rad, 0), y1 = min(cy + rad, g.dy - 1);
        const int z0 = max(cz - rad, 0), z1 = min(cz + rad, g.dz - 1);
        const int ny = y1 - y0 + 1, nz = z1 - z0 + 1;
        const int nrows = (x0 <= x1 && ny > 0 && nz > 0) ? ny * nz : 0;
        for (int rbase = 0; rbase < nrows; rbase += RC) {
            int a[RC], b[RC];
            int yy = y0 + rbase % ny, zz = z0 + rbase / ny;
            #pragma unroll
            for (int r = 0; r < RC; r++) {
                a[r] = 0; b[r] = 0;
                if (rbase + r < nrows) {
                    const int row = (zz * g.dy + yy) * g.dx;
                    a[r] = cell_start[row + x0];
                    b[r] = cell_end[row + x1];
                }
                if (++yy > y1) { yy = y0; zz++; }
            }
            int s[RC];                                // exclusive prefix of the row lengths
            int T = 0;
            #pragma unroll
            for (int r = 0; r < RC; r++) { s[r] = T; T += b[r] - a[r]; }
            for (int t0 = 0; t0 < T; t0 += GB) {
                float4 m[GB]; int p[GB];
                #pragma unroll
                for (int u = 0; u < GB; u++) {
                    const int t = t0 + u;
                    int ar = a[0], sr = 0;
                    #pragma unroll
                    for (int r = 1; r < RC; r++) if (t >= s[r]) { ar = a[r]; sr = s[r]; }
                    p[u] = ar + (t - sr);
                    if (t < T) m[u] = pts[p[u]];
                }
                #pragma unroll
                for (int u = 0; u < GB; u++) {
                    if (t0 + u < T) {
                        const float ddx = qx - m[u].x, ddy = qy - m[u].y, ddz = qz - m[u].z;
                        float dd = ddx * ddx; dd += ddy * ddy; dd += ddz * ddz;
                        lane_insert5(out.key, out.pos, ((unsigned long long)__float_as_uint(dd) << 32) | (unsigned)__float_as_int(m[u].w), p[u]);
                    }
                }
            }
        }
        const float d5 = __uint_as_float((unsigned)(out.key[4] >> 32));
        const bool have5 = out.key[4] != ~0ull;
        const float guard = (float)rad * g.h * 0.9995f;
        if (have5 && d5 < guard * guard) break;
        if (rad >= g.rmax) break;
        rad = min(rad * 2, g.rmax);
    }
    return out.key[4] != ~0ull && __uint_as_float((unsigned)(out.key[4] >> 32)) < 1.0f;
}
