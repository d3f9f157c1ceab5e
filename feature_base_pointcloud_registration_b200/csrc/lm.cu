// lm.cu -- the scan-to-map Gauss-Newton ("LM") loop, all iterations in ONE kernel launch.
//
// Replaces, per frame (mapOptmization.h): the loop of scan2MapOptimization (:1417-1436) with
// cornerOptimization (:1002-1124), surfOptimization (:1126-1215), combineOptimizationCoeffs
// (:1218-1243), LMOptimization (:1246-1401) and transformUpdate (:1444-1479).
//
// B200 mapping: one thread-block CLUSTER per frame.  Warps own chunks of 32 feature points
// (transform -> exact 5-NN in the grid index, warp-cooperative -> line / plane fit, coefficient and
// Jacobian row, one thread per point); rows are staged in shared memory and folded warp-wise into the 21+6 unique entries of J^T J / J^T r
// in f64; a shared-memory block reduce produces one partial per CTA; after ONE hardware cluster
// barrier every CTA reads all partials in rank order and redundantly solves the 6x6 system (QR,
// degeneracy projection at iteration 0, pose update, convergence test), so a frame needs one
// cluster.sync() per iteration and no grid-wide or host synchronisation at all.
// Independent frames are independent clusters (blockIdx.x / cluster size); a single frame can
// instead take the whole GPU as one cooperative grid (grid.sync() per iteration).
//
// Bound: latency (<= 30 dependent iterations) and L1/L2 gathers of map cells; never tensor cores
// (contractions are K=3 and 6x6).  Compulsory traffic per iteration: 16 B per feature point +
// 5 x 16 B neighbours (B_iter = 96 * (n_c + n_s)).
#include <cooperative_groups.h>

#include "internal.cuh"
#include "mapgrid.cuh"
#include "smallmat.cuh"

namespace cg = cooperative_groups;


namespace {

// Launch shapes.  Batched calls: one cluster per frame, 512-thread CTAs (2 resident per SM at 64 registers), normally 2 CTAs per
// frame.  Measured on 1024 config-4 frames, pipelined 256-frame launches (ms per 1024 frames): 256-thread CTAs x cluster 4:
// 48.06; 1024 x 1: 46.78; 1024 x 2: 47.27; 512 x 4: 47.43; 512 x 2: 46.08 -- the fewer DIFFERENT frames share an SM's L1 the
// better (with 256-thread CTAs an SM serves four frames at once), and one 1024-thread CTA per frame loses more to its coarse
// launch granularity than it wins.  Single-frame calls: one 768-thread CTA per SM over the whole GPU.
#ifndef LM_CTAS
#define LM_CTAS 2                  // resident CTAs per SM the batched shape is compiled for (register budget = 65536 / (LM_BATCH_TPB * LM_CTAS))
#endif
#ifndef LM_BATCH_TPB
#define LM_BATCH_TPB 512           // threads per CTA of the batched shape
#endif
constexpr int LM_TPB_CLUSTER = LM_BATCH_TPB, LM_CTAS_CLUSTER = LM_CTAS;
constexpr int LM_TPB_GRID = 768, LM_CTAS_GRID = 1;
#ifndef LM_CARVEOUT
#define LM_CARVEOUT 16
#endif
constexpr int LM_PART_SLOTS = 64;      // chunks of one CTA whose partial sums have their own shared-memory slot
constexpr int NACC = 28;          // 21 (upper triangle of A^T A) + 6 (A^T b) + 1 (row count)

__device__ __forceinline__ void transform_point(const float* T, float4 p, float& x, float& y, float& z) {
    x = T[0] * p.x + T[1] * p.y + T[2] * p.z + T[3];
    y = T[4] * p.x + T[5] * p.y + T[6] * p.z + T[7];
    z = T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11];
}

// cornerOptimization body for one point (mapOptmization.h:1026-1121)
__device__ __forceinline__ bool corner_fit(const float4 (&nb)[5], float x0, float y0, float z0, float4& coeff) {
    float px[5], py[5], pz[5];
    #pragma unroll
    for (int j = 0; j < 5; j++) { px[j] = nb[j].x; py[j] = nb[j].y; pz[j] = nb[j].z; }
    float cx = 0, cy = 0, cz = 0;
    #pragma unroll
    for (int j = 0; j < 5; j++) { cx += px[j]; cy += py[j]; cz += pz[j]; }
    cx /= 5; cy /= 5; cz /= 5;
    float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
    #pragma unroll
    for (int j = 0; j < 5; j++) {
        float ax = px[j] - cx, ay = py[j] - cy, az = pz[j] - cz;
        a11 += ax * ax; a12 += ax * ay; a13 += ax * az;
        a22 += ay * ay; a23 += ay * az;
        a33 += az * az;
    }
    a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;
    float D1[3], V1[9];
    dev_jacobi3(a11, a12, a13, a22, a23, a33, D1, V1);
    if (!(D1[0] > 3 * D1[1])) return false;
    float x1 = (float)((double)cx + 0.1 * (double)V1[0]);
    float y1 = (float)((double)cy + 0.1 * (double)V1[1]);
    float z1 = (float)((double)cz + 0.1 * (double)V1[2]);
    float x2 = (float)((double)cx - 0.1 * (double)V1[0]);
    float y2 = (float)((double)cy - 0.1 * (double)V1[1]);
    float z2 = (float)((double)cz - 0.1 * (double)V1[2]);
    float a012 = sqrtf(((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1))
                     + ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1)) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1))
                     + ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1)) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1)));
    float l12 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
    float la = ((y1 - y2) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1))
              + (z1 - z2) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1))) / a012 / l12;
    float lb = -((x1 - x2) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1))
               - (z1 - z2) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1))) / a012 / l12;
    float lc = -((x1 - x2) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1))
               + (y1 - y2) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1))) / a012 / l12;
    float ld2 = a012 / l12;
    float s = (float)(1.0 - 0.9 * (double)fabsf(ld2));
    coeff = make_float4(s * la, s * lb, s * lc, s * ld2);
    return (double)s > 0.1;
}

// surfOptimization body for one point (mapOptmization.h:1153-1212)
__device__ __forceinline__ bool surf_fit(const float4 (&nb)[5], float x0, float y0, float z0, float4& coeff) {
    float A0[15];
    #pragma unroll
    for (int j = 0; j < 5; j++) { A0[3 * j] = nb[j].x; A0[3 * j + 1] = nb[j].y; A0[3 * j + 2] = nb[j].z; }
    float X0[3];
    dev_plane_solve(A0, X0);
    float pa = X0[0], pb = X0[1], pc = X0[2], pd = 1;
    float ps = sqrtf(pa * pa + pb * pb + pc * pc);
    pa /= ps; pb /= ps; pc /= ps; pd /= ps;
    #pragma unroll
    for (int j = 0; j < 5; j++)
        if ((double)fabsf(pa * A0[3 * j] + pb * A0[3 * j + 1] + pc * A0[3 * j + 2] + pd) > 0.2) return false;
    float pd2 = pa * x0 + pb * y0 + pc * z0 + pd;
    float s = (float)(1.0 - 0.9 * (double)fabsf(pd2) / (double)sqrtf(sqrtf(x0 * x0 + y0 * y0 + z0 * z0)));
    coeff = make_float4(s * pa, s * pb, s * pc, s * pd2);
    return (double)s > 0.1;
}

// tf::Quaternion / tf::Matrix3x3 helpers (f64) for transformUpdate (mapOptmization.h:1459-1472)
__device__ inline void q_set_rpy(double roll, double pitch, double yaw, double q[4]) {
    double hy = yaw * 0.5, hp = pitch * 0.5, hr = roll * 0.5;
    double cy = cos(hy), sy = sin(hy), cp = cos(hp), sp = sin(hp), cr = cos(hr), sr = sin(hr);
    q[0] = sr * cp * cy - cr * sp * sy; q[1] = cr * sp * cy + sr * cp * sy;
    q[2] = cr * cp * sy - sr * sp * cy; q[3] = cr * cp * cy + sr * sp * sy;
}
__device__ inline double q_dot(const double a[4], const double b[4]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3]; }
__device__ inline void q_slerp(const double a[4], const double b[4], double t, double o[4]) {
    double s = sqrt(q_dot(a, a) * q_dot(b, b));
    double d = q_dot(a, b);
    double theta = (d < 0 ? acos(-d / s) * 2.0 : acos(d / s) * 2.0) / 2.0;
    if (theta != 0.0) {
        double dd = 1.0 / sin(theta), s0 = sin((1.0 - t) * theta), s1 = sin(t * theta);
        double sg = d < 0 ? -1.0 : 1.0;
        for (int k = 0; k < 4; k++) o[k] = (a[k] * s0 + sg * b[k] * s1) * dd;
    } else {
        for (int k = 0; k < 4; k++) o[k] = a[k];
    }
}
__device__ inline void q_get_rpy(const double q[4], double& roll, double& pitch, double& yaw) {
    const double PI = 3.14159265358979323846;
    double d = q_dot(q, q), s = 2.0 / d;
    double xs = q[0] * s, ys = q[1] * s, zs = q[2] * s;
    double wx = q[3] * xs, wy = q[3] * ys, wz = q[3] * zs;
    double xx = q[0] * xs, xy = q[0] * ys, xz = q[0] * zs, yy = q[1] * ys, yz = q[1] * zs, zz = q[2] * zs;
    double m00 = 1.0 - (yy + zz), m10 = xy + wz, m20 = xz - wy, m21 = yz + wx, m22 = 1.0 - (xx + yy);
    double m01 = xy - wz, m02 = xz + wy;
    if (fabs(m20) >= 1) {
        yaw = 0;
        double delta = atan2(m01, m02);
        if (m20 < 0) { pitch = PI / 2.0; roll = delta; } else { pitch = -PI / 2.0; roll = delta; }
    } else {
        pitch = -asin(m20);
        roll = atan2(m21 / cos(pitch), m22 / cos(pitch));
        yaw = atan2(m10 / cos(pitch), m00 / cos(pitch));
    }
}
__device__ inline float clampf(float v, float lim) { if (v < -lim) v = -lim; if (v > lim) v = lim; return v; }

__device__ __noinline__ void transform_update(float* pose, long long imuAvailable, float imuRollInit, float imuPitchInit, float rot_tol, float z_tol) {
    if (imuAvailable == 1 && fabs((double)imuPitchInit) < 1.4) {
        double q0[4], q1[4], qm[4], r, p, y;
        q_set_rpy((double)pose[0], 0, 0, q0); q_set_rpy((double)imuRollInit, 0, 0, q1);
        q_slerp(q0, q1, 0.05, qm); q_get_rpy(qm, r, p, y);
        pose[0] = (float)r;
        q_set_rpy(0, (double)pose[1], 0, q0); q_set_rpy(0, (double)imuPitchInit, 0, q1);
        q_slerp(q0, q1, 0.05, qm); q_get_rpy(qm, r, p, y);
        pose[1] = (float)p;
    }
    pose[0] = clampf(pose[0], rot_tol);
    pose[1] = clampf(pose[1], rot_tol);
    pose[5] = clampf(pose[5], z_tol);
}

// LMOptimization after the reduction (mapOptmization.h:1336-1400).  The 6x6 QR solve (cv::solve, :1343) is run by the whole
// first warp (dev_qr_solve6_warp: one matrix column per lane, registers only) -- it is on the critical path of every iteration
// of a frame, and one thread walking local-memory arrays needed ~10 us for it; the rest (degeneracy test at iteration 0, pose
// update, convergence) is one thread.  Returns 1 when converged.  matP is the reference's LOCAL zero-initialised matrix (:1278).
__device__ __noinline__ int lm_solve_step(const float* AtA, const float* Xqr, int iter, int& isDegenerate, float* pose, float* Xout) {
    float Aw[36], X[6];
    for (int k = 0; k < 6; k++) X[k] = Xqr[k];
    float matP[36];
    for (int k = 0; k < 36; k++) matP[k] = 0.f;
    if (iter == 0 && dev_surely_not_degenerate(AtA)) {
        isDegenerate = 0;                            // every eigenvalue is provably >= 100: matP is never used
    } else if (iter == 0) {
        float E[6], V[36], V2[36], Vinv[36];
        for (int k = 0; k < 36; k++) Aw[k] = AtA[k];
        dev_jacobi<6>(Aw, E, V);
        for (int k = 0; k < 36; k++) V2[k] = V[k];
        isDegenerate = 0;
        for (int i = 5; i >= 0; i--) {
            if (E[i] < 100.f) { for (int j = 0; j < 6; j++) V2[i * 6 + j] = 0; isDegenerate = 1; }
            else break;
        }
        for (int k = 0; k < 36; k++) Aw[k] = V[k];
        dev_lu_invert6(Aw, Vinv);
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < 6; j++) {
                double s = 0.0;
                for (int q = 0; q < 6; q++) s += (double)Vinv[i * 6 + q] * (double)V2[q * 6 + j];
                matP[i * 6 + j] = (float)s;
            }
    }
    if (isDegenerate) {
        float X2[6];
        for (int k = 0; k < 6; k++) X2[k] = X[k];
        for (int i = 0; i < 6; i++) {
            double s = 0.0;
            for (int q = 0; q < 6; q++) s += (double)matP[i * 6 + q] * (double)X2[q];
            X[i] = (float)s;
        }
    }
    for (int k = 0; k < 6; k++) { pose[k] += X[k]; Xout[k] = X[k]; }
    double r0 = (double)(X[0] * 57.29578f), r1 = (double)(X[1] * 57.29578f), r2 = (double)(X[2] * 57.29578f);
    double t0 = (double)(X[3] * 100), t1 = (double)(X[4] * 100), t2 = (double)(X[5] * 100);
    float deltaR = (float)sqrt(r0 * r0 + r1 * r1 + r2 * r2);
    float deltaT = (float)sqrt(t0 * t0 + t1 * t1 + t2 * t2);
    return ((double)deltaR < 0.05 && (double)deltaT < 0.05) ? 1 : 0;
}

// One LM iteration = association (a warp takes 32 consecutive feature points: transform, then the exact 5-NN
// of each point on the grid index by the WHOLE WARP, one point after the other (mapgrid.cuh), then ONE
// THREAD per point: line / plane fit -> coefficient -> Jacobian row, staged as f64 in shared memory) -> every
// warp folds its 32 staged rows into the 27 unique entries of J^T J / J^T r (+ the row count), one
// entry per lane, f64 -> CTA reduce in fixed warp order -> one partial per CTA -> team barrier ->
// every CTA sums all partials in the same fixed order and solves redundantly.
// The TEAM of one frame is a thread-block cluster (GRID = false, many frames per launch) or the whole
// cooperative grid (GRID = true, one frame).  A warp takes 32 CONSECUTIVE points of the scan (voxel order),
// so successive queries touch neighbouring map cells and find their candidates in L1.
template <bool GRID>
__global__ void __launch_bounds__(GRID ? LM_TPB_GRID : LM_TPB_CLUSTER, GRID ? LM_CTAS_GRID : LM_CTAS_CLUSTER) lm_kernel(LmArgs a) {
    constexpr int LM_TPB = GRID ? LM_TPB_GRID : LM_TPB_CLUSTER;
    cg::cluster_group cluster = cg::this_cluster();
    cg::grid_group grid = cg::this_grid();
    const int C = GRID ? (int)gridDim.x : (int)cluster.num_blocks();        // CTAs in the team
    const int rank = GRID ? (int)blockIdx.x : (int)cluster.block_rank();
    const int slot = GRID ? a.first : a.first + (int)(blockIdx.x / C);
    FrameMeta& M = a.meta[slot];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int WPB = LM_TPB / 32;

    // per-warp staging of the 32 Jacobian rows (6) and -residual (1) of the warp's current 32 points, as f32 -- widened exactly when folded (dynamic shared memory)
    extern __shared__ float s_dyn[];
    float (*s_rows)[32][7] = reinterpret_cast<float (*)[32][7]>(s_dyn);      // staged as f32 (what they are); widened exactly when folded
    // one 28-double partial per CHUNK (dynamic dispatch, summed in chunk order: the result does not depend on which warp
    // took which chunk) or, when this CTA has more chunks than slots, one per WARP (static dispatch)
    // Batched (cluster) shape: the per-chunk partials live in global memory (L2) instead, so that a CTA needs only ~8 KB of
    // shared memory and the rest of the SM's 228 KB serves as L1 for the neighbour searches.
    constexpr int SLOTS = GRID ? (LM_PART_SLOTS > WPB ? LM_PART_SLOTS : WPB) : 1;
    __shared__ double s_part[SLOTS][NACC];
    __shared__ int s_next;
    __shared__ double sh_tot[NACC];
    __shared__ GridDesc sh_gd[2];
    __shared__ float sh_pose[6], sh_T[12], sh_trig[6], sh_AtA[36], sh_AtB[6];
    __shared__ int sh_stop, sh_nsel;

    const int nC = min(M.n_corner_ds, a.cornerCap), nS = min(M.n_surf_ds, a.surfCap);
    if (!(nC > a.edgeMin && nS > a.surfMin)) {     // mapOptmization.h:1410 / :1439-1441
        if (rank == 0 && tid == 0) { M.flags = FBPR_FLAG_NOT_ENOUGH_FEATURES | (M.mapTruncated ? FBPR_FLAG_MAP_TRUNCATED : 0u); M.iters = 0; M.nSel = 0; M.isDegenerate = 0; }
        return;
    }
    const int nQ = nC + nS;
    const GridSeg gc = a.gsegs[2 * slot], gs = a.gsegs[2 * slot + 1];
    const float4* cpts = a.cornerDS + (size_t)slot * a.cornerCap;
    const float4* spts = a.surfDS + (size_t)slot * a.surfCap;
    double* part = GRID ? a.partialsGrid + (size_t)slot * 2 * a.gridMax * NACC : a.partials + (size_t)slot * 2 * a.teamMax * NACC;
    const int teamStride = GRID ? a.gridMax : a.teamMax;
    if (tid < 6) sh_pose[tid] = M.pose[tid];
    if (tid == 32) sh_gd[0] = *gc.desc;
    if (tid == 64) sh_gd[1] = *gs.desc;
    if (tid == 0) s_next = 0;
    // the accumulator entry this lane owns: 0..20 = upper triangle (ei <= ej), 21..26 = (ei, 6) = A^T b, 27 = row count
    int ei = 0, ej = 6;
    if (lane < 21) { int rr = 0, qx = lane; while (qx >= 6 - rr) { qx -= 6 - rr; rr++; } ei = rr; ej = rr + qx; }
    else if (lane < 27) ei = lane - 21;
    __syncthreads();

    KnnMaps maps; maps.gd = sh_gd; maps.cell_start[0] = gc.cell_start; maps.cell_start[1] = gs.cell_start; maps.pts[0] = gc.sorted; maps.pts[1] = gs.sorted;
    // points per warp chunk: 32 when the team has fewer warps than chunks (batched calls); when a whole GPU works
    // on one frame there are more warps than that, so chunks shrink until every warp has a few points to search
    int CS = 32;
    while (CS > 1 && nQ <= (CS / 2) * C * WPB) CS >>= 1;
    const int nChunks = (nQ + CS - 1) / CS;
    const int myChunks = rank < nChunks ? (nChunks - rank + C - 1) / C : 0;      // chunks k * C + rank of this CTA
    const bool dynamic = !GRID || myChunks <= LM_PART_SLOTS;
    double* cpart = GRID ? nullptr : a.chunkPart + (size_t)slot * a.chunkCap * NACC;        // [chunk][28], chunk = k * C + rank
    unsigned flags = 0; int isDegenerate = 0; int iters = 0;
    for (int iter = 0; iter < FBPR_MAX_ITERS; iter++) {
        // --- pose -> rigid transform + the six sines/cosines LMOptimization needs (:1259-1264)
        if (tid < 6) {
            float ang = sh_pose[tid % 3];            // 0 roll, 1 pitch, 2 yaw
            sh_trig[tid] = tid < 3 ? sinf_c(ang) : cosf_c(ang);
        }
        __syncthreads();
        if (tid == 0) {
            // pcl::getTransformation from the shared trig values (same expression order as Appendix B-4)
            float F = sh_trig[0], D = sh_trig[1], B = sh_trig[2], E = sh_trig[3], Cc = sh_trig[4], A = sh_trig[5];
            float DE = D * E, DF = D * F;
            sh_T[0] = A * Cc; sh_T[1] = A * DF - B * E; sh_T[2]  = B * F + A * DE; sh_T[3]  = sh_pose[3];
            sh_T[4] = B * Cc; sh_T[5] = A * E + B * DF; sh_T[6]  = B * DE - A * F; sh_T[7]  = sh_pose[4];
            sh_T[8] = -D;     sh_T[9] = Cc * F;         sh_T[10] = Cc * E;         sh_T[11] = sh_pose[5];
        }
        __syncthreads();
        const bool cap = iter == a.debug_iter && slot < a.dbgSlots;

        // --- association: one thread per feature point
        double acc = 0.0;
        // chunks of CS consecutive points are dealt round-robin over the team's CTAs (chunk = k * C + rank), so the (more
        // expensive) corner chunks at the front of the index range spread evenly; inside the CTA the warps take the
        // CTA's chunks from a shared counter, because a chunk costs anything between 0 and 32 full searches
        int k = warp;
        if (dynamic) { if (lane == 0) k = atomicAdd(&s_next, 1); k = __shfl_sync(0xffffffffu, k, 0); }
        while (k < myChunks) {
            const int chunk = k * C + rank;
            const int q = chunk * CS + lane;
            bool ok = false;
            // corners occupy [0, nC), surface points follow; each lane searches the map of its own kind
            const bool inRange = lane < CS && q < nQ;
            const bool isCorner = q < nC;
            const int li = isCorner ? q : q - nC;
            float4 pOri = make_float4(0, 0, 0, 0);
            float x0 = 0.f, y0 = 0.f, z0 = 0.f;
            if (inRange) { pOri = __ldcs(isCorner ? cpts + li : spts + li); transform_point(sh_T, pOri, x0, y0, z0); }
            const int kind = isCorner ? 0 : 1;
            const bool active = inRange && sh_gd[kind].n >= 5;
            float4* anchor = a.qanchor + (size_t)slot * a.qCap + q;
            int* cache = a.qcache + ((size_t)slot * a.qCap + q) * FBPR_KNN_CACHE;
            const float4* mo = isCorner ? gc.pts : gs.pts;   // the map in original order (XYZI)
            ThreadKnn5 r;
            #pragma unroll
            for (int i = 0; i < 5; i++) r.key[i] = ~0ull;
            // (1) candidate cache of this point's last full search: re-rank the cached map points (one thread); the
            //     result is the exact 5-NN when every uncached map point is provably farther (mapgrid.cuh)
            bool need = active;
            float ball0 = a.firstRadius;             // radius of the first search pass (metres)
            if (active) {
                if (iter > 0) {
                    // the point's record (anchor + cached indices) in ONE round trip: the index loads do not wait for the anchor's test
                    const int4* c4 = reinterpret_cast<const int4*>(cache);
                    const float4 an = __ldcs(anchor);
                    int4 cv[FBPR_KNN_CACHE / 4];
                    #pragma unroll
                    for (int k = 0; k < FBPR_KNN_CACHE / 4; k++) cv[k] = __ldcs(c4 + k);
                    if (an.w > 0.f) {
                        #pragma unroll
                        for (int half = 0; half < FBPR_KNN_CACHE / 8; half++) {     // 8 independent gathers in flight
                            const int4 va = cv[2 * half], vb = cv[2 * half + 1];
                            const int ci[8] = { va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w };
                            float4 cm[8];
                            #pragma unroll
                            for (int k = 0; k < 8; k++) if (ci[k] >= 0) cm[k] = __ldcg(mo + ci[k]);
                            #pragma unroll
                            for (int k = 0; k < 8; k++) if (ci[k] >= 0) knn_offer_idx(r, cm[k].x, cm[k].y, cm[k].z, ci[k], x0, y0, z0);
                        }
                        const float mx = x0 - an.x, my = y0 - an.y, mz = z0 - an.z;
                        const float moved = sqrtf(mx * mx + my * my + mz * mz);
                        const float lim = an.w * 0.9999f - moved * 1.0001f - 1.0e-5f;     // every uncached point is farther than this
                        const float lim2 = lim > 0.f ? lim * lim * 0.9999f : 0.f;
                        need = !(knn_d5(r) < lim2 || lim2 >= 1.0f);                       // 2nd case: the cache decides the whole 1 m ball
                        ball0 = knn5_ball_from_bound(knn_d5(r), a.firstRadius);
                    } else {
                        ball0 = 2.0f;
                    }
                }
            }
#ifdef FBPR_KNN_STATS
            {
                const unsigned ba = __ballot_sync(0xffffffffu, active), bn = __ballot_sync(0xffffffffu, need), bi = __ballot_sync(0xffffffffu, active && iter > 0);
                KNN_STAT(19, __popc(ba)); KNN_STAT(17, __popc(bi)); KNN_STAT(18, __popc(bi & ~bn)); KNN_STAT(20 + min(iter, 7), __popc(bn));
            }
#endif
            // (2) full search by the whole warp for the points that need one; it refreshes their cache
            warp_knn5(maps, kind, x0, y0, z0, ball0, need, r, anchor, cache);
            ok = active && knn_d5(r) < 1.0f;
            if (inRange) {
                const GridDesc& gd = sh_gd[isCorner ? 0 : 1];
                if (cap) {
                    const size_t o = isCorner ? (size_t)slot * a.cornerCap + li : (size_t)slot * a.surfCap + li;
                    int* kd = isCorner ? a.knnC : a.knnS; float* dd = isCorner ? a.d2C : a.d2S;
                    #pragma unroll
                    for (int k = 0; k < 5; k++) {
                        kd[5 * o + k] = ok ? (int)(unsigned)(r.key[k] & 0xffffffffu) : -1;
                        dd[5 * o + k] = (gd.n >= 5 && r.key[k] != ~0ull) ? __uint_as_float((unsigned)(r.key[k] >> 32)) : 3.0e38f;
                    }
                }
                float4 coeff = make_float4(0, 0, 0, 0);
                if (ok) {
                    // the five neighbours' coordinates, by original map index (mapOptmization.h:1028-1036, :1157-1163)
                    float4 nb[5];
                    #pragma unroll
                    for (int k = 0; k < 5; k++) nb[k] = __ldcg(mo + knn_index(r, k));
                    ok = isCorner ? corner_fit(nb, x0, y0, z0, coeff) : surf_fit(nb, x0, y0, z0, coeff);
                }
                if (cap) {
                    if (isCorner) { size_t o = (size_t)slot * a.cornerCap + li; a.coeffC[o] = coeff; a.flagC[o] = ok ? 1 : 0; }
                    else          { size_t o = (size_t)slot * a.surfCap + li;   a.coeffS[o] = coeff; a.flagS[o] = ok ? 1 : 0; }
                }
                if (ok) {
                    // Jacobian row, lidar -> "camera" axis permutation (mapOptmization.h:1286-1332)
                    const float srz = sh_trig[0], srx = sh_trig[1], sry = sh_trig[2], crz = sh_trig[3], crx = sh_trig[4], cry = sh_trig[5];
                    const float ox = pOri.y, oy = pOri.z, oz = pOri.x;
                    const float kx = coeff.y, ky = coeff.z, kz = coeff.x;
                    float arx = (crx * sry * srz * ox + crx * crz * sry * oy - srx * sry * oz) * kx
                              + (-srx * srz * ox - crz * srx * oy - crx * oz) * ky
                              + (crx * cry * srz * ox + crx * cry * crz * oy - cry * srx * oz) * kz;
                    float ary = ((cry * srx * srz - crz * sry) * ox
                              + (sry * srz + cry * crz * srx) * oy + crx * cry * oz) * kx
                              + ((-cry * crz - srx * sry * srz) * ox
                              + (cry * srz - crz * srx * sry) * oy - crx * sry * oz) * kz;
                    float arz = ((crz * srx * sry - cry * srz) * ox + (-cry * crz - srx * sry * srz) * oy) * kx
                              + (crx * crz * ox - crx * srz * oy) * ky
                              + ((sry * srz + cry * crz * srx) * ox + (crz * sry - cry * srx * srz) * oy) * kz;
                    float* row = s_rows[warp][lane];
                    row[0] = arz; row[1] = arx; row[2] = ary;
                    row[3] = kz;  row[4] = kx;  row[5] = ky;
                    row[6] = -coeff.w;
                }
            }
            // fold the warp's staged rows into the lane-owned entries (rows in point order, f64)
            unsigned m = __ballot_sync(0xffffffffu, ok);
            __syncwarp();
            if (lane == 27) acc += (double)__popc(m);
            else if (lane < 27) {
                while (m) {
                    const int rr = __ffs(m) - 1; m &= m - 1;
                    acc += (double)s_rows[warp][rr][ei] * (double)s_rows[warp][rr][ej];
                }
            }
            __syncwarp();
            if (dynamic) {
                if (lane < NACC) { if (GRID) s_part[k][lane] = acc; else cpart[(size_t)chunk * NACC + lane] = acc; }
                acc = 0.0;
                if (lane == 0) k = atomicAdd(&s_next, 1);
                k = __shfl_sync(0xffffffffu, k, 0);
            } else {
                k += WPB;
            }
        }
        // --- CTA reduce (shared memory, fixed order) -> one partial per CTA in global memory
        if (GRID && !dynamic && lane < NACC) s_part[warp][lane] = acc;
        __syncthreads();
        const int buf = iter & 1;
        double* mypart = part + ((size_t)buf * teamStride + rank) * NACC;
        if (tid < NACC) {
            double v = 0.0;
            if (GRID) {
                const int nslots = dynamic ? myChunks : WPB;
                for (int w = 0; w < nslots; w++) v += s_part[w][tid];
            } else {
                // chunk order, four loads in flight (L2: the lines were written by other warps of this CTA before the barrier)
                int w = 0;
                for (; w + 4 <= myChunks; w += 4) {
                    const double p0 = __ldcg(cpart + (size_t)((w + 0) * C + rank) * NACC + tid), p1 = __ldcg(cpart + (size_t)((w + 1) * C + rank) * NACC + tid);
                    const double p2 = __ldcg(cpart + (size_t)((w + 2) * C + rank) * NACC + tid), p3 = __ldcg(cpart + (size_t)((w + 3) * C + rank) * NACC + tid);
                    v += p0; v += p1; v += p2; v += p3;
                }
                for (; w < myChunks; w++) v += __ldcg(cpart + (size_t)(w * C + rank) * NACC + tid);
            }
            mypart[tid] = v;
            __threadfence();
        }
        if (tid == 0) s_next = 0;                    // next iteration's dispatch counter (ordered by the barriers below)
        if (GRID) grid.sync(); else cluster.sync();
        // --- every CTA sums all partials in the same fixed order (bitwise identical everywhere), then solves redundantly
        if (tid < NACC * 8) {                        // 224 threads <= LM_TPB in both shapes
            const int v = tid >> 3, s8 = tid & 7;
            double sum = 0.0;
            const double* base = part + (size_t)buf * teamStride * NACC + v;
            for (int rk = s8; rk < C; rk += 8) sum += __ldcg(base + (size_t)rk * NACC);
            sum += __shfl_xor_sync(0xffffffffu, sum, 4);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            if (s8 == 0) sh_tot[v] = sum;
        }
        __syncthreads();
        if (tid < NACC) {
            const double v = sh_tot[tid];
            if (tid < 21) {
                int rr = 0, qx = tid;                 // unpack the upper-triangle index
                while (qx >= 6 - rr) { qx -= 6 - rr; rr++; }
                int cc = rr + qx;
                float fv = (float)v;
                sh_AtA[rr * 6 + cc] = fv; sh_AtA[cc * 6 + rr] = fv;
            } else if (tid < 27) {
                sh_AtB[tid - 21] = (float)v;
            } else {
                sh_nsel = (int)(v + 0.5);
            }
        }
        __syncthreads();
        float Xqr[6];
        if (tid < 32 && sh_nsel >= 50) dev_qr_solve6_warp(sh_AtA, sh_AtB, Xqr);      // cv::solve(matAtA, matAtB, matX, DECOMP_QR), :1343
        if (tid == 0) {
            int stop = 0;
            iters = iter + 1;
            if (sh_nsel < 50) {                     // :1267-1270: LMOptimization returns false, pose unchanged;
                flags |= FBPR_FLAG_TOO_FEW_CORRESPONDENCES;   // every remaining iteration repeats identically
                iters = FBPR_MAX_ITERS; stop = 1;
            } else {
                float pose[6], X[6];
                for (int k = 0; k < 6; k++) pose[k] = sh_pose[k];
                int conv = lm_solve_step(sh_AtA, Xqr, iter, isDegenerate, pose, X);
                for (int k = 0; k < 6; k++) sh_pose[k] = pose[k];
                if (conv) { flags |= FBPR_FLAG_CONVERGED; stop = 1; }
                if (rank == 0 && cap) {
                    for (int k = 0; k < 36; k++) a.dbgAtA[(size_t)slot * 36 + k] = sh_AtA[k];
                    for (int k = 0; k < 6; k++) { a.dbgAtB[(size_t)slot * 6 + k] = sh_AtB[k]; a.dbgX[(size_t)slot * 6 + k] = X[k]; }
                }
            }
            if (rank == 0 && a.poseTrace)
                for (int k = 0; k < 6; k++) a.poseTrace[((size_t)slot * FBPR_MAX_ITERS + iter) * 6 + k] = sh_pose[k];
            sh_stop = stop;
        }
        __syncthreads();
        if (sh_stop) break;
    }
    if (rank == 0 && tid == 0) {
        float pose[6];
        for (int k = 0; k < 6; k++) pose[k] = sh_pose[k];
        if (isDegenerate) flags |= FBPR_FLAG_DEGENERATE;
        if (M.mapTruncated) flags |= FBPR_FLAG_MAP_TRUNCATED;
        transform_update(pose, M.imuAvailable, M.imuRollInit, M.imuPitchInit, a.rot_tol, a.z_tol);
        for (int k = 0; k < 6; k++) M.pose[k] = pose[k];
        M.iters = iters; M.flags = flags; M.nSel = sh_nsel; M.isDegenerate = isDegenerate;
    }
}

__global__ void transform_update_kernel(FrameMeta* meta, int first, int count, float rot_tol, float z_tol) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    FrameMeta& M = meta[first + i];
    float pose[6];
    for (int k = 0; k < 6; k++) pose[k] = M.pose[k];
    transform_update(pose, M.imuAvailable, M.imuRollInit, M.imuPitchInit, rot_tol, z_tol);
    for (int k = 0; k < 6; k++) M.pose[k] = pose[k];
}

}  // namespace

#ifndef LM_EXTRA_SMEM
#define LM_EXTRA_SMEM 0            // experiments only: unused dynamic shared memory per batched CTA (limits the CTAs resident per SM)
#endif
static size_t lm_dyn_smem(int tpb) { return (size_t)(tpb / 32) * 32 * 7 * sizeof(float) + (tpb == LM_TPB_CLUSTER ? LM_EXTRA_SMEM : 0); }    // s_rows

// one-time function attributes of both variants (cluster sizes above 8, dynamic shared memory above the default limit)
static int lm_configure() {
    static int configured = 0;
    if (configured) return 0;
    cudaError_t e = cudaFuncSetAttribute(lm_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return fbpr_fail(e, "cudaFuncSetAttribute(NonPortableClusterSizeAllowed)", __FILE__, __LINE__);
    // The batched shape keeps its per-chunk partial sums in global memory, so a CTA needs ~8 KB of shared memory; ask for the
    // smallest carve-out that holds 4 CTAs and leave the rest of the SM's 228 KB to L1, which the neighbour searches live on.
    // LM ms per 128 frames: partials in shared memory (22 KB per CTA) at carve-out 25 / 40 / 55 / 100 %: 7.52 / 5.15 / 5.23 / 6.76;
    // partials in global memory at 8 / 14 / 16 / 20 / 28 / 40 %: 5.08 / 4.81 / 4.50 / 4.52 / 4.58 / 4.54.
    cudaFuncSetAttribute(lm_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, LM_CARVEOUT);
    e = cudaFuncSetAttribute(lm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lm_dyn_smem(LM_TPB_GRID));
    if (e != cudaSuccess) return fbpr_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)", __FILE__, __LINE__);
    configured = 1;
    return 0;
}

int fbpr_knn_cache_slots() { return FBPR_KNN_CACHE; }

#ifdef FBPR_KNN_STATS
extern "C" __attribute__((visibility("default"))) int fbpr_debug_knn_stats(unsigned long long* out, int reset) {     // diagnostics build only
    cudaDeviceSynchronize();
    if (out) cudaMemcpyFromSymbol(out, g_knn_stats, sizeof(unsigned long long) * 32);
    if (reset) { unsigned long long z[32] = { 0 }; cudaMemcpyToSymbol(g_knn_stats, z, sizeof(z)); }
    return 0;
}
#endif

int fbpr_lm_grid_blocks(int device) {
    // co-resident CTAs of the cooperative (one frame on the whole GPU) variant: one per SM
    if (lm_configure()) return 0;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
    int per = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, lm_kernel<true>, LM_TPB_GRID, lm_dyn_smem(LM_TPB_GRID)) != cudaSuccess || per < 1) return 0;
    return sms;
}

// co-resident clusters of `c` CTAs of the batched variant (0 if the size cannot be launched)
static int lm_max_active_clusters(int c) {
    static int cached[17] = { 0 }, known[17] = { 0 };
    if (c < 1 || c > 16) return 0;
    if (!known[c]) {
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(LM_TPB_CLUSTER); cfg.gridDim = dim3((unsigned)(c * 64)); cfg.dynamicSmemBytes = lm_dyn_smem(LM_TPB_CLUSTER);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)c; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, lm_kernel<false>, &cfg) != cudaSuccess) { n = 0; cudaGetLastError(); }
        cached[c] = n; known[c] = 1;
    }
    return cached[c];
}

// cluster size for a batch of `count` frames: the time of the launch is ~ waves / (CTAs per frame); among the sizes whose
// estimate is within 15 % of the best one the SMALLEST wins (cheaper barrier, fewer redundant solves, and fewer warps than chunks
// keeps the warps of a frame busy) -- a batch that fills the GPU gets 2 CTAs per frame, a small batch big clusters
int fbpr_lm_auto_cluster(int count) {
    // powers of two only: odd sizes are allowed when asked for (lm_cluster_size = 1..16) but measured no better; one CTA per
    // frame (16 warps for ~220 chunks) measured 1.46x slower than two and is only taken when asked for
    const int sizes[4] = { 2, 4, 8, 16 };
    double score[4]; double bestScore = 1e30;
    for (int k = 0; k < 4; k++) {
        const int c = sizes[k], n = lm_max_active_clusters(c);
        score[k] = n > 0 ? (double)((count + n - 1) / n) / (double)c : 1e30;
        if (score[k] < bestScore) bestScore = score[k];
    }
    for (int k = 0; k < 4; k++) if (score[k] <= bestScore * 1.15) return sizes[k];
    return 2;
}

int fbpr_launch_lm(const LmArgs& args, int count, int cluster_size, int grid_blocks, cudaStream_t st, long long* launches) {
    if (count <= 0) return 0;
    { int rc = lm_configure(); if (rc) return rc; }
    cudaLaunchConfig_t cfg = {};
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e;
    if (count == 1 && grid_blocks > 0) {           // one frame: the whole GPU cooperates, grid-wide barrier per iteration
        cfg.blockDim = dim3(LM_TPB_GRID);
        cfg.dynamicSmemBytes = lm_dyn_smem(LM_TPB_GRID);
        cfg.gridDim = dim3((unsigned)grid_blocks);
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        e = cudaLaunchKernelEx(&cfg, lm_kernel<true>, args);
    } else {                                        // many frames: one cluster per frame, hardware cluster barrier per iteration
        if (cluster_size <= 0) cluster_size = fbpr_lm_auto_cluster(count);
        cfg.blockDim = dim3(LM_TPB_CLUSTER);
        cfg.dynamicSmemBytes = lm_dyn_smem(LM_TPB_CLUSTER);
        cfg.gridDim = dim3((unsigned)(count * cluster_size));
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cluster_size; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        e = cudaLaunchKernelEx(&cfg, lm_kernel<false>, args);
    }
    if (e != cudaSuccess) return fbpr_fail(e, "cudaLaunchKernelEx(lm_kernel)", __FILE__, __LINE__);
    if (launches) *launches += 1;
    return fbpr_launch_ok("lm_kernel");
}

int fbpr_launch_transform_update(FrameMeta* meta, int first, int count, float rot_tol, float z_tol, cudaStream_t st, long long* launches) {
    if (count <= 0) return 0;
    transform_update_kernel<<<(count + 63) / 64, 64, 0, st>>>(meta, first, count, rot_tol, z_tol);
    if (launches) *launches += 1;
    return fbpr_launch_ok("transform_update_kernel");
}
