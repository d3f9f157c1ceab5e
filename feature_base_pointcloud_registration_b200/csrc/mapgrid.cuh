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
