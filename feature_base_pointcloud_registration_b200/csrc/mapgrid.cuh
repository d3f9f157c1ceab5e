// mapgrid.cuh -- device-side exact 5-NN query on the uniform-grid map index.
//
// Replaces pcl::KdTreeFLANN::nearestKSearch(p, 5, ...) (mapOptmization.h:1020, :1143; FLANN
// KDTreeSingleIndex, L2_Simple, exact, sorted -- SURVEY.md Appendix B-2).  Only neighbours
// with d^2 < 1.0 matter to the caller (:1027, :1154), so the query returns the exact 5-NN inside
// the 1 m ball or "reject".  Distance is ((dx*dx)+(dy*dy))+(dz*dz) in f32 without FMA (the
// library is built with -fmad=false) so ordering and the < 1.0 gate match the reference; ties
// are broken by the original point index.
//
// ONE WARP per query, 32 queries per warp taken one after the other.  The map is stored
// cell-contiguous (counting sort, mapgrid.cu) with x the fastest cell axis, so the (2R+1)^3 cells
// around a query are (2R+1)^2 contiguous ranges of the sorted array ("rows"):
//   - the lanes fetch the bounds of up to 32 rows at once and prefix-sum their lengths;
//   - the concatenated candidates are walked 64 at a time: consecutive lanes read consecutive
//     float4 points (coalesced), every load of a trip is issued before any is consumed, each lane
//     keeps a private sorted top-5 of packed keys in registers (compare-exchange chain, no local
//     memory), and five redux.sync rounds merge the 32 private lists;
//   - the radius of a query's first pass (metres) is chosen by the caller.  The LM kernel derives it
//     from the previous iteration: the old 5 neighbours lie within sqrt(d5_old) + |p_new - p_old| of
//     the new position, so a cube covering that radius contains the new exact 5-NN and ONE pass is
//     enough.  If the covered ball does not yet certify the result (first iteration, R too small)
//     the radius is doubled and the (larger) ball is scanned again.
//   - a pass with radius R certifies its result only inside the ball of radius R * h around the query,
//     so each row is trimmed to the chord of that ball: about a sixth of the cube's points are read.
//   - temporal coherence across LM iterations: a full search also leaves a CANDIDATE CACHE for the point -- up
//     to FBPR_KNN_CACHE (16) map indices and a radius tau such that every map point NOT in the cache is at least tau away from
//     the search position.  At the next iteration the point has moved by delta, so every uncached map point
//     is at least tau - delta away; one thread re-ranks the cached candidates, and if their 5th distance is
//     below tau - delta (or tau - delta covers the whole 1 m ball) that is the exact 5-NN and no search is
//     needed.  Otherwise the warp searches again, inside the radius the cached 5th distance bounds.
// A query therefore costs two or three dependent memory round trips whatever its candidate count
// (25 for a surface point in a dense map, ~1000 for a corner point next to several edges).
// Ranking key = (d^2 bits as u32) << 32 | original index: the (d^2, index) total order of the oracle
// (d^2 >= 0, so the f32 bit pattern is monotonic).
#pragma once
#include "internal.cuh"

#ifdef FBPR_KNN_STATS
// diagnostics build only (make EXTRA=-DFBPR_KNN_STATS, scripts/knn_stats.py; never shipped): [0] searches [1] passes [3] candidates
// [4] rows [5..12] passes by candidate count <=8, <=16, <=32, <=64, <=128, <=256, <=512, more  [13..16] passes by rows <=9, <=25, <=32,
// more [17] cache re-ranks [18] re-ranks that certify [19] point-iterations [20..27] searches at iteration 0..7+
static __device__ unsigned long long g_knn_stats[32];
#define KNN_STAT(i, v) do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_knn_stats[i], (unsigned long long)(v)); } while (0)
#else
#define KNN_STAT(i, v) do { } while (0)
#endif

// the (up to) two map indices a warp's queries may refer to: kind 0 = corner map, kind 1 = surface map
struct KnnMaps {
    const GridDesc* gd;           // [2], usually in shared memory
    const int* cell_start[2];     // [ncells + 1] exclusive prefix of the cell populations
    const float4* pts[2];         // cell-sorted copies (w = original index bits)
};

struct ThreadKnn5 {
    unsigned long long key[5];   // ascending; 0xffff... = empty.  Only ever indexed with compile-time constants (registers).
};

// Bounds that only decide WHICH cells are scanned (always with explicit margins, never a result) use the one-instruction
// approximations (relative error <= 2^-22) instead of the IEEE sequences the library is otherwise compiled for (-prec-div / -prec-sqrt).
#ifndef FBPR_KNN_APPROX
#define FBPR_KNN_APPROX 1
#endif
__device__ __forceinline__ float knn_sqrt_bound(float x) {
#if FBPR_KNN_APPROX
    float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
    return sqrtf(x);
#endif
}
__device__ __forceinline__ float knn_rcp_bound(float x) {
#if FBPR_KNN_APPROX
    return __fdividef(1.0f, x);
#else
    return 1.0f / x;
#endif
}

__device__ __forceinline__ void knn_cswap(unsigned long long& lo, unsigned long long& hi) {   // order a pair
    const unsigned long long a = lo, b = hi;
    const bool sw = b < a;
    lo = sw ? b : a; hi = sw ? a : b;
}

__device__ __forceinline__ void knn_offer(ThreadKnn5& r, const float4 m, float qx, float qy, float qz) {
    const float ddx = qx - m.x, ddy = qy - m.y, ddz = qz - m.z;
    float dd = ddx * ddx; dd += ddy * ddy; dd += ddz * ddz;
    if (__float_as_uint(dd) > (unsigned)(r.key[4] >> 32)) return;          // cannot enter the top-5 (the common case)
    const unsigned long long key = ((unsigned long long)__float_as_uint(dd) << 32) | (unsigned)__float_as_int(m.w);
    if (key >= r.key[4]) return;
    r.key[4] = key;                                                         // replace the worst, then bubble it down
    knn_cswap(r.key[3], r.key[4]); knn_cswap(r.key[2], r.key[3]); knn_cswap(r.key[1], r.key[2]); knn_cswap(r.key[0], r.key[1]);
}

__device__ __forceinline__ void knn_offer_idx(ThreadKnn5& r, float mx, float my, float mz, int idx, float qx, float qy, float qz) {
    knn_offer(r, make_float4(mx, my, mz, __int_as_float(idx)), qx, qy, qz);
}

__device__ __forceinline__ float knn_d5(const ThreadKnn5& r) {      // 5th-best squared distance (huge when fewer than 5 were found)
    return r.key[4] == ~0ull ? 3.0e38f : __uint_as_float((unsigned)(r.key[4] >> 32));
}
__device__ __forceinline__ int knn_index(const ThreadKnn5& r, int k) { return (int)(unsigned)(r.key[k] & 0xffffffffu); }   // original map index

// Merge the private sorted lists of the 32 lanes into the warp's top-5, replicated in every lane.  Keys are unique
// (the low word is the map index), so in each of the five rounds exactly one lane owns the minimum and pops it.
__device__ __forceinline__ void knn5_merge_warp(ThreadKnn5& p, ThreadKnn5& m) {
    const unsigned FULL = 0xffffffffu;
    #pragma unroll
    for (int k = 0; k < 5; k++) {
        const unsigned hi = (unsigned)(p.key[0] >> 32), lo = (unsigned)p.key[0];
        const unsigned mhi = __reduce_min_sync(FULL, hi);
        const unsigned mlo = __reduce_min_sync(FULL, hi == mhi ? lo : 0xffffffffu);
        m.key[k] = ((unsigned long long)mhi << 32) | mlo;
        if (hi == mhi && lo == mlo && m.key[k] != ~0ull) { p.key[0] = p.key[1]; p.key[1] = p.key[2]; p.key[2] = p.key[3]; p.key[3] = p.key[4]; p.key[4] = ~0ull; }
    }
}

// Stream the concatenation of 32 ranges (lane i owns [a, a + len)) through the lanes' private top-5 lists:
// 64 candidates per trip, consecutive lanes on consecutive points, both loads of a trip issued before use.
__device__ __forceinline__ int knn5_scan_ranges(ThreadKnn5& p, const float4* __restrict__ pts, int a, int len, float qx, float qy, float qz) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int incl = len;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
    const int total = __shfl_sync(FULL, incl, 31);
    const int excl = incl - len;
    for (int t0 = 0; t0 < total; t0 += 64) {
        float4 m[2]; bool v[2];
        #pragma unroll
        for (int u = 0; u < 2; u++) {
            if (u == 1 && t0 + 32 >= total) { v[1] = false; break; }   // (uniform) the trip's second half is empty: most searches scan < 32 candidates
            const int t = t0 + u * 32 + lane;
            int j = 0;                                  // the last lane whose exclusive offset is <= t owns candidate t
            #pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
                const int e = __shfl_sync(FULL, excl, (j + s) & 31);
                if (e <= t) j += s;
            }
            const int aj = __shfl_sync(FULL, a, j), ej = __shfl_sync(FULL, excl, j);
            v[u] = t < total;
            if (v[u]) m[u] = __ldg(pts + aj + (t - ej));
        }
        #pragma unroll
        for (int u = 0; u < 2; u++) if (v[u]) knn_offer(p, m[u], qx, qy, qz);
    }
    return total;
}

#ifndef FBPR_KNN_CACHE
#define FBPR_KNN_CACHE 16           // cached candidates per feature point (multiple of 8, <= 32; 16 measured best: 8 -> 6.1 ms, 16 -> 5.9, 24 -> 6.2, 32 -> 6.7 per 128 frames)
#endif

// Exact 5-NN of ONE query by the whole warp (all lanes pass the same query).  rad0 = first cube radius in cells.
// r (replicated) = exact sorted 5-NN among all map points within the covered ball; the caller rejects when
// knn_d5(r) >= 1.0.  Exactness: every point closer than rad * h (minus a rounding guard) lies in the cube, so the
// result is final once the 5th distance is inside that ball or the cube covers the whole 1 m ball (rad = rmax).
// cache (may be null) receives up to FBPR_KNN_CACHE original indices (-1 = unused slot); the return value is tau (metres):
// every map point whose index is not in the cache is at distance >= tau from the query (0 = cache not usable).
__device__ __forceinline__ float warp_query_knn5(const GridDesc& g, const int* __restrict__ cell_start, const float4* __restrict__ pts,
                                                 float qx, float qy, float qz, float ball0, ThreadKnn5& r, int* cache) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int cx = (int)floorf((qx - g.ox) * g.inv_h);
    const int cy = (int)floorf((qy - g.oy) * g.inv_h);
    const int cz = (int)floorf((qz - g.oz) * g.inv_h);
    ThreadKnn5 p;                                            // private list of this lane
    // ball = radius (metres) of the pass: everything within it is scanned, the result is final when the 5th distance lies inside
    // it.  The last ball covers the 1 m gate of the caller.
    const float ballMax = 1.35f;                             // beyond the 1 m gate: a point with nothing within 1.35 m keeps a cache that says so while it moves < 0.3 m
    float ball = fminf(fmaxf(ball0, 0.05f), ballMax);
    float guard;
    const float slack = g.h * 1.0e-3f + 1.0e-5f;             // cell edges as the f32 cell_of_point sees them
    KNN_STAT(0, 1);
    while (true) {
        int st_tot = 0;                                      // candidates of this pass (diagnostics build only)
        // A pass certifies its result only inside a ball around the query, and the candidate cache can never reach farther either,
        // so only the part of every cell row that can intersect that ball is scanned -- about a sixth of the points of the cube of
        // cells around it -- and the ball's radius need not be a multiple of the cell edge.
        guard = ball * 0.9995f;
        const float ballR = ball * 1.0001f + 1.0e-5f;
        const int rad = (int)ceilf(ballR * g.inv_h * 1.0001f);       // cells beyond cx +- rad are farther than the ball (the query lies in cell cx)
        #pragma unroll
        for (int i = 0; i < 5; i++) p.key[i] = ~0ull;        // a widening pass rescans its whole (larger) ball
        const int x0 = max(cx - rad, 0), x1 = min(cx + rad, g.dx - 1);
        const int y0 = max(cy - rad, 0), y1 = min(cy + rad, g.dy - 1);
        const int z0 = max(cz - rad, 0), z1 = min(cz + rad, g.dz - 1);
        const int ny = y1 - y0 + 1, nz = z1 - z0 + 1;
        const int nrows = (x0 <= x1 && ny > 0 && nz > 0) ? ny * nz : 0;
        const float inv_ny = knn_rcp_bound((float)max(ny, 1));
        for (int rbase = 0; rbase < nrows; rbase += 32) {
            const int i = rbase + lane;
            int a0 = 0, l0 = 0;
            if (i < nrows) {
                const int zi = (int)(((float)i + 0.5f) * inv_ny);            // i / ny for the small integers involved
                const int zz = z0 + zi, yy = y0 + (i - zi * ny);
                const int row = (zz * g.dy + yy) * g.dx;
                // distance from the query to the row's cell column in y and z, then the chord of the ball along x
                const float ylo = g.oy + (float)yy * g.h, zlo = g.oz + (float)zz * g.h;
                const float dy = fmaxf(0.f, fmaxf(ylo - qy, qy - (ylo + g.h)) - slack);
                const float dz = fmaxf(0.f, fmaxf(zlo - qz, qz - (zlo + g.h)) - slack);
                const float rem = ballR * ballR - dy * dy - dz * dz;
                if (rem >= 0.f) {
                    const float hc = knn_sqrt_bound(rem) * 1.0001f + slack;
                    const int xa = max(x0, (int)floorf((qx - hc - g.ox) * g.inv_h));
                    const int xb = min(x1, (int)floorf((qx + hc - g.ox) * g.inv_h));
                    if (xa <= xb) { a0 = __ldg(cell_start + row + xa); l0 = __ldg(cell_start + row + xb + 1) - a0; }
                }
            }
            st_tot += knn5_scan_ranges(p, pts, a0, l0, qx, qy, qz);
        }
#ifdef FBPR_KNN_STATS
        KNN_STAT(1, 1); KNN_STAT(3, st_tot); KNN_STAT(4, nrows);
        KNN_STAT(st_tot <= 8 ? 5 : st_tot <= 16 ? 6 : st_tot <= 32 ? 7 : st_tot <= 64 ? 8 : st_tot <= 128 ? 9 : st_tot <= 256 ? 10 : st_tot <= 512 ? 11 : 12, 1);
        KNN_STAT(nrows <= 9 ? 13 : nrows <= 25 ? 14 : nrows <= 32 ? 15 : 16, 1);
#endif
        (void)st_tot;
        ThreadKnn5 pm;                                       // the merge consumes a copy; the private lists feed the cache below
        #pragma unroll
        for (int i = 0; i < 5; i++) pm.key[i] = p.key[i];
        knn5_merge_warp(pm, r);
        if (knn_d5(r) < guard * guard || ball >= ballMax) break;
        ball = fminf(ball * 2.0f, ballMax);
    }
    if (!cache) return 0.f;
    // ---- candidate cache.  Points outside it are: never scanned (>= guard away), pushed out of a full private
    // list (>= that list's last key), or kept but >= the threshold chosen below.
    const unsigned d5b = (unsigned)(r.key[4] >> 32);                                  // 0xffffffff when fewer than 5 exist
    unsigned tb = __float_as_uint(guard * guard);                                     // threshold on d^2 bits, exclusive
    tb = min(tb, __reduce_min_sync(FULL, (unsigned)(p.key[4] >> 32)));                // tails of the full lists (empty = 0xffffffff)
    int count = 0, c = 0;                                                             // c = this lane's cached entries at the final threshold
    for (int tries = 0; tries < 8; tries++) {
        const unsigned long long tk = (unsigned long long)tb << 32;
        c = 0;
        #pragma unroll
        for (int i = 0; i < 5; i++) c += p.key[i] < tk ? 1 : 0;
        count = __reduce_add_sync(FULL, c);
        if (count <= FBPR_KNN_CACHE) break;
        // too many: shrink towards the 5th distance in proportion to the surplus (point counts grow ~ with d^2)
        const float t2 = __uint_as_float(tb), d5f = __uint_as_float(d5b);
        tb = __float_as_uint(d5f + (t2 - d5f) * ((float)(FBPR_KNN_CACHE - 3) / (float)count));
    }
    const bool usable = count <= FBPR_KNN_CACHE && (d5b == 0xffffffffu || tb > d5b);   // must hold the whole top-5
    {
        if (!usable) c = 0;
        int incl = c;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
        const int total = __shfl_sync(FULL, incl, 31);
        int off = incl - c;
        #pragma unroll
        for (int i = 0; i < 5; i++) if (i < c) cache[off++] = (int)(unsigned)(p.key[i] & 0xffffffffu);   // the list is sorted: its first c keys qualify
        if (lane >= total && lane < FBPR_KNN_CACHE) cache[lane] = -1;
    }
    return usable ? sqrtf(__uint_as_float(tb)) : 0.f;
}

// Full searches for the lanes of a warp that `need` one (one query per lane; `kind` = which of the two maps it
// searches, `ball0` = the radius of its first pass in metres); must be called by all 32 lanes.  A lane that does not need a
// search keeps its r.  anchor / cache (per lane: this lane's record, may be null) receive the candidate cache.
__device__ __forceinline__ void warp_knn5(const KnnMaps& M, int kind, float qx, float qy, float qz, float ball0, bool need, ThreadKnn5& r,
                                          float4* anchor, int* cache) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    unsigned todo = __ballot_sync(FULL, need);
    while (todo) {
        const int src = __ffs(todo) - 1; todo &= todo - 1;
        const float bx = __shfl_sync(FULL, qx, src), by = __shfl_sync(FULL, qy, src), bz = __shfl_sync(FULL, qz, src);
        const int bk = __shfl_sync(FULL, kind, src); const float br = __shfl_sync(FULL, ball0, src);
        int* bc = reinterpret_cast<int*>(__shfl_sync(FULL, reinterpret_cast<unsigned long long>(cache), src));
        ThreadKnn5 m;
        const float tau = warp_query_knn5(M.gd[bk], bk ? M.cell_start[1] : M.cell_start[0], bk ? M.pts[1] : M.pts[0], bx, by, bz, br, m, bc);
        if (lane == src) {
            #pragma unroll
            for (int i = 0; i < 5; i++) r.key[i] = m.key[i];
            if (anchor) *anchor = make_float4(qx, qy, qz, tau);
        }
    }
}

// First ball radius (metres) that certainly contains the exact 5-NN when 5 map points are known to lie within squared distance
// d5_known of the query (huge = unknown: cover the whole 1 m gate), never below `floor_m` -- the scan of a ball also decides how
// far the candidate cache reaches, and a cache that reaches less than the first iteration's ball only brings the searches back.
__device__ __forceinline__ float knn5_ball_from_bound(float d5_known, float floor_m) {
    if (!(d5_known < 1.0e30f)) return 2.0f;
    return fmaxf(sqrtf(d5_known) * 1.001f + 1.0e-5f, floor_m);
}
