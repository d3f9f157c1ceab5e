// lm.cu -- the scan-to-map Gauss-Newton ("LM") loop, all iterations in ONE kernel launch.
//
// Replaces, per frame (mapOptmization.h): the loop of scan2MapOptimization (:1417-1436) with
// cornerOptimization (:1002-1124), surfOptimization (:1126-1215), combineOptimizationCoeffs
// (:1218-1243), LMOptimization (:1246-1401) and transformUpdate (:1444-1479).
//
// B200 mapping.  The team of one frame is a thread-block CLUSTER (batched calls: many frames per
// launch, one hardware cluster barrier per iteration) or the whole cooperative GRID (one frame on
// the whole GPU).  Inside an iteration a CTA walks MACRO-CHUNKS of spatially sorted feature
// points (lm_order below sorts the down-sampled scan along a Morton curve once per frame):
//   1. every thread transforms its point and finds its map cell; a block reduce gives the box of
//      cells the chunk's search cubes cover;
//   2. the cell_start entries of that box are copied to shared memory (they are the row bounds),
//      a block scan turns the row lengths into tile offsets, and every cell ROW of the box -- one
//      contiguous run of the cell-sorted map -- is fetched with ONE cp.async.bulk (1-D TMA) that
//      completes on an mbarrier: the chunk's candidate points now sit in a shared-memory TILE;
//   3. ONE THREAD PER POINT (G = 1; the single-frame shape uses G = 8 lanes per point) scans the
//      cells of its own cube out of the tile -- exact 5-NN by (d^2, index) keys in registers, no
//      dependent global loads, no cross-lane traffic -- and certifies the result against the
//      cube's inscribed ball exactly like the warp-cooperative search (mapgrid.cuh), which
//      remains the fallback for the few points whose cube must grow beyond the tile;
//   4. the five neighbours are read back from the tile, line / plane fit, coefficient, Jacobian
//      row; rows are folded warp-wise into the 21 + 6 unique entries of J^T J / J^T r in f64.
// After the team barrier every CTA sums the per-CTA partials in the same order and redundantly
// solves the 6x6 system, so a frame needs one barrier per iteration and no host round trip.
//
// Bound: instruction issue + shared-memory bandwidth of the tile scans and L2 -> shared-memory
// bulk copies (about 20 staged map points per feature point and iteration); never tensor cores
// (contractions are K=3 and 6x6).  Compulsory traffic per iteration: 16 B per feature point +
// 5 x 16 B neighbours (B_iter = 96 * (n_c + n_s)).
#include <cooperative_groups.h>

#include "internal.cuh"
#include "mapgrid.cuh"
#include "smallmat.cuh"

namespace cg = cooperative_groups;


namespace {

// Launch shapes.  Batched calls: one cluster per frame, 256-thread CTAs, one thread per feature point (4 CTAs resident per
// SM at 64 registers and ~53 KB of shared memory each).  Single-frame calls: one 512-thread CTA per SM over the whole GPU,
// 8 lanes per feature point (a frame has only ~50 points per SM; the lanes split the rows of the point's search cube).
constexpr int LM_TPB_CLUSTER = 256, LM_CTAS_CLUSTER = 4, LM_G_CLUSTER = 1;
constexpr int LM_TPB_GRID = 512, LM_CTAS_GRID = 1, LM_G_GRID = 8;
constexpr int NACC = 28;          // 21 (upper triangle of A^T A) + 6 (A^T b) + 1 (row count)
// shared-memory tile of one macro-chunk
constexpr int LM_TILE_PTS = 2048; // staged map points (the usable count is 2^posBits - 1 <= 2047: the top position marks "not in the tile")
constexpr int LM_CS_CAP = 4096;   // staged cell_start entries = rows * (cells per row + 1)
constexpr int LM_ROW_CAP = 768;   // cell rows of the box (3 per thread of the batched shape)
constexpr size_t LM_DYN_SMEM = (size_t)LM_TILE_PTS * 16 + (size_t)LM_CS_CAP * 4 + (size_t)(LM_ROW_CAP + 8) * 4;

// ---- 1-D bulk async copy (TMA) + mbarrier, sm_90+ PTX ------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");      // make the initialised barrier visible to the async proxy
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// global -> this CTA's shared memory, `bytes` a multiple of 16, both addresses 16-byte aligned; completes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

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

// LMOptimization after the reduction (mapOptmization.h:1336-1400), one thread.
// Returns 1 when converged.  matP is the reference's LOCAL zero-initialised matrix (:1278).
__device__ __noinline__ int lm_solve_step(const float* AtA, const float* AtB, int iter, int& isDegenerate, float* pose, float* Xout) {
    float Aw[36], bw[6], X[6];
    for (int k = 0; k < 36; k++) Aw[k] = AtA[k];
    for (int k = 0; k < 6; k++) bw[k] = AtB[k];
    dev_qr_solve6(Aw, bw, X);
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

// ---- lm_order: spatial order of the down-sampled feature points, once per frame ---------------------------------------
// The association works on macro-chunks of consecutive feature points and stages the map cells around a chunk in shared
// memory, so consecutive points must be close in space.  laserCloud{Corner,Surf}LastDS come out of the VoxelGrid in voxel-key
// order (x fastest: long thin strips).  One CTA per window of LM_ORDER_WIN points sorts them along a Morton curve of their
// position in the map frame (initial guess) with a shared-memory bitonic sort of (morton18 << 13 | index) keys and writes the
// permuted points with their original index in w.  The ORDER affects only the speed of the association and the association
// order of the f64 normal-equation sums, never a per-point result.
constexpr int LM_ORDER_WIN = 8192, LM_ORDER_TPB = 1024;

__device__ __forceinline__ unsigned morton_spread6(unsigned v) {     // abcdef -> a00b00c00d00e00f
    v &= 63u;
    v = (v | (v << 8)) & 0x300Fu;
    v = (v | (v << 4)) & 0x30C3u;
    v = (v | (v << 2)) & 0x9249u;
    return v;
}

__global__ void __launch_bounds__(LM_ORDER_TPB) lm_order(LmArgs a) {
    const int slot = a.first + blockIdx.z, kind = blockIdx.y, win = blockIdx.x;
    const FrameMeta& M = a.meta[slot];
    const int nC = min(M.n_corner_ds, a.cornerCap), nS = min(M.n_surf_ds, a.surfCap);
    const int n = kind == 0 ? nC : nS;
    const int base = win * LM_ORDER_WIN;
    if (base >= n) return;
    const int cnt = min(LM_ORDER_WIN, n - base);
    const float4* src = (kind == 0 ? a.cornerDS + (size_t)slot * a.cornerCap : a.surfDS + (size_t)slot * a.surfCap) + base;
    float4* dst = a.qpts + (size_t)slot * a.qCap + (kind == 0 ? 0 : nC) + base;
    __shared__ unsigned s_key[LM_ORDER_WIN];
    __shared__ float s_T[12];
    __shared__ unsigned s_bb[6];
    const int tid = threadIdx.x;
    if (tid == 0) get_transformation(M.pose[3], M.pose[4], M.pose[5], M.pose[0], M.pose[1], M.pose[2], s_T);
    if (tid < 3) s_bb[tid] = 0xffffffffu; else if (tid < 6) s_bb[tid] = 0u;
    __syncthreads();
    constexpr int IPT = LM_ORDER_WIN / LM_ORDER_TPB;
    float px[IPT], py[IPT], pz[IPT];
    unsigned mn[3] = { 0xffffffffu, 0xffffffffu, 0xffffffffu }, mx[3] = { 0u, 0u, 0u };
    #pragma unroll
    for (int k = 0; k < IPT; k++) {
        const int i = k * LM_ORDER_TPB + tid;
        px[k] = py[k] = pz[k] = 0.f;
        if (i < cnt) {
            const float4 p = src[i];
            px[k] = s_T[0] * p.x + s_T[1] * p.y + s_T[2] * p.z + s_T[3];
            py[k] = s_T[4] * p.x + s_T[5] * p.y + s_T[6] * p.z + s_T[7];
            pz[k] = s_T[8] * p.x + s_T[9] * p.y + s_T[10] * p.z + s_T[11];
            if (isfinite(px[k]) && isfinite(py[k]) && isfinite(pz[k])) {
                const unsigned ex = f2ord(px[k]), ey = f2ord(py[k]), ez = f2ord(pz[k]);
                mn[0] = min(mn[0], ex); mn[1] = min(mn[1], ey); mn[2] = min(mn[2], ez);
                mx[0] = max(mx[0], ex); mx[1] = max(mx[1], ey); mx[2] = max(mx[2], ez);
            }
        }
    }
    #pragma unroll
    for (int c = 0; c < 3; c++) { mn[c] = __reduce_min_sync(0xffffffffu, mn[c]); mx[c] = __reduce_max_sync(0xffffffffu, mx[c]); }
    if ((tid & 31) == 0) for (int c = 0; c < 3; c++) { atomicMin(&s_bb[c], mn[c]); atomicMax(&s_bb[3 + c], mx[c]); }
    __syncthreads();
    int npad = 32; while (npad < cnt) npad <<= 1;
    {
        float lo[3] = { 0.f, 0.f, 0.f }, sc[3] = { 0.f, 0.f, 0.f };
        if (s_bb[0] <= s_bb[3]) {
            #pragma unroll
            for (int c = 0; c < 3; c++) {
                lo[c] = ord2f(s_bb[c]);
                const float ext = ord2f(s_bb[3 + c]) - lo[c];
                sc[c] = ext > 0.f ? 63.999f / ext : 0.f;
            }
            // one scale for all axes (cubic Morton cells): the largest extent decides
            const float smin = fminf(sc[0] > 0.f ? sc[0] : 3.0e38f, fminf(sc[1] > 0.f ? sc[1] : 3.0e38f, sc[2] > 0.f ? sc[2] : 3.0e38f));
            sc[0] = sc[1] = sc[2] = smin < 3.0e38f ? smin : 0.f;
        }
        #pragma unroll
        for (int k = 0; k < IPT; k++) {
            const int i = k * LM_ORDER_TPB + tid;
            if (i < npad) {
                unsigned key = 0xffffffffu;
                if (i < cnt) {
                    const float fx = (px[k] - lo[0]) * sc[0], fy = (py[k] - lo[1]) * sc[1], fz = (pz[k] - lo[2]) * sc[2];
                    const unsigned qx = (unsigned)min(max((int)fx, 0), 63), qy = (unsigned)min(max((int)fy, 0), 63), qz = (unsigned)min(max((int)fz, 0), 63);
                    const unsigned mort = morton_spread6(qx) | (morton_spread6(qy) << 1) | (morton_spread6(qz) << 2);      // 18 bits
                    key = (mort << 13) | (unsigned)i;
                }
                s_key[i] = key;
            }
        }
    }
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (npad >> 1); t += LM_ORDER_TPB) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), p = i | j;
                const unsigned x = s_key[i], y = s_key[p];
                const bool up = (i & k) == 0;
                if ((x > y) == up) { s_key[i] = y; s_key[p] = x; }
            }
            __syncthreads();
        }
    }
    for (int r = tid; r < cnt; r += LM_ORDER_TPB) {
        const int i = (int)(s_key[r] & (LM_ORDER_WIN - 1));
        const float4 p = src[i];
        dst[r] = make_float4(p.x, p.y, p.z, __int_as_float(base + i));          // w = index inside its cloud (corner or surface)
    }
}

// ---- exact 5-NN of one point out of the shared-memory tile ---------------------------------------------------------------
// Keys are (d^2 bits << 32) | (original map index << posBits) | position in the tile: the (d^2, index) total order of the
// oracle (a map point occupies one tile position, so the low bits never decide), and the position lets the fit read the
// neighbour's coordinates back from shared memory.  Position 2^posBits - 1 means "not in the tile" (fallback search).
__device__ __forceinline__ void tile_offer(ThreadKnn5& r, const float4 m, int pos, int posBits, float qx, float qy, float qz) {
    const float ddx = qx - m.x, ddy = qy - m.y, ddz = qz - m.z;
    float dd = ddx * ddx; dd += ddy * ddy; dd += ddz * ddz;
    if (__float_as_uint(dd) > (unsigned)(r.key[4] >> 32)) return;          // cannot enter the top-5 (the common case)
    const unsigned long long key = ((unsigned long long)__float_as_uint(dd) << 32) | (((unsigned)__float_as_int(m.w) << posBits) | (unsigned)pos);
    if (key >= r.key[4]) return;
    r.key[4] = key;
    knn_cswap(r.key[3], r.key[4]); knn_cswap(r.key[2], r.key[3]); knn_cswap(r.key[1], r.key[2]); knn_cswap(r.key[0], r.key[1]);
}

struct TileBox { int x0, y0, z0, nx1, ny; };     // first cell of the staged box, entries per staged cell_start row (cells + 1), rows per z layer

// scan the cells of the cube of radius `rad` around cell (cx, cy, cz); lane `sub` of the point's G lanes takes every G-th row
template <int G>
__device__ __forceinline__ void tile_scan(ThreadKnn5& p, const float4* __restrict__ s_tile, const int* __restrict__ s_cs, const int* __restrict__ s_rowOff,
                                          const TileBox& B, const GridDesc& g, int cx, int cy, int cz, int rad, int sub, int posBits,
                                          float qx, float qy, float qz) {
    const int xa = max(cx - rad, 0), xb = min(cx + rad, g.dx - 1);
    const int ya = max(cy - rad, 0), yb = min(cy + rad, g.dy - 1);
    const int za = max(cz - rad, 0), zb = min(cz + rad, g.dz - 1);
    if (xa > xb || ya > yb || za > zb) return;
    const int wy = yb - ya + 1, nrow = wy * (zb - za + 1);
    for (int i = sub; i < nrow; i += G) {
        const int zi = i / wy, yi = i - zi * wy;
        const int row = (za + zi - B.z0) * B.ny + (ya + yi - B.y0);
        const int* cs = s_cs + row * B.nx1;
        const int off = s_rowOff[row] - cs[0];
        int j = off + cs[xa - B.x0];
        const int hi = off + cs[xb + 1 - B.x0];
        for (; j + 4 <= hi; j += 4) {                       // four shared-memory loads in flight
            const float4 m0 = s_tile[j], m1 = s_tile[j + 1], m2 = s_tile[j + 2], m3 = s_tile[j + 3];
            tile_offer(p, m0, j, posBits, qx, qy, qz); tile_offer(p, m1, j + 1, posBits, qx, qy, qz);
            tile_offer(p, m2, j + 2, posBits, qx, qy, qz); tile_offer(p, m3, j + 3, posBits, qx, qy, qz);
        }
        for (; j < hi; j++) tile_offer(p, s_tile[j], j, posBits, qx, qy, qz);
    }
}

// merge the private lists of the G lanes of one point (G consecutive lanes of a warp); every lane ends with the merged top-5
template <int G>
__device__ __forceinline__ void tile_merge(ThreadKnn5& p) {
    if (G == 1) return;
    // only the G lanes of this point are known to be here together (other points of the warp may have no search to do)
    const unsigned gmask = (G >= 32 ? 0xffffffffu : ((1u << G) - 1u)) << ((threadIdx.x & 31) & ~(G - 1));
    ThreadKnn5 m;
    #pragma unroll
    for (int k = 0; k < 5; k++) {
        unsigned long long best = p.key[0];
        #pragma unroll
        for (int o = G >> 1; o > 0; o >>= 1) { const unsigned long long other = __shfl_xor_sync(gmask, best, o); best = other < best ? other : best; }
        m.key[k] = best;
        if (p.key[0] == best && best != ~0ull) { p.key[0] = p.key[1]; p.key[1] = p.key[2]; p.key[2] = p.key[3]; p.key[3] = p.key[4]; p.key[4] = ~0ull; }
    }
    #pragma unroll
    for (int k = 0; k < 5; k++) p.key[k] = m.key[k];
}

// One LM iteration = association over the CTA's macro-chunks (see the file header) -> every warp folds the Jacobian rows of
// its points into the 27 unique entries of J^T J / J^T r (+ the row count), one entry per lane, f64 -> CTA reduce in fixed warp
// order -> one partial per CTA -> team barrier -> every CTA sums all partials in the same fixed order and solves redundantly.
// The TEAM of one frame is a thread-block cluster (GRID = false, many frames per launch) or the whole cooperative grid
// (GRID = true, one frame).  Macro-chunk k * C + rank belongs to CTA `rank` of the team in every iteration (static, so
// the summation order is fixed); corner chunks come first, chunks never mix the two maps.
template <bool GRID>
__global__ void __launch_bounds__(GRID ? LM_TPB_GRID : LM_TPB_CLUSTER, GRID ? LM_CTAS_GRID : LM_CTAS_CLUSTER) lm_kernel(LmArgs a) {
    constexpr int LM_TPB = GRID ? LM_TPB_GRID : LM_TPB_CLUSTER;
    constexpr int G = GRID ? LM_G_GRID : LM_G_CLUSTER;       // lanes per feature point
    constexpr int MC = LM_TPB / G;                           // feature points per macro-chunk
    constexpr int WPB = LM_TPB / 32;
    constexpr int QPW = 32 / G;                              // feature points per warp
    cg::cluster_group cluster = cg::this_cluster();
    cg::grid_group grid = cg::this_grid();
    const int C = GRID ? (int)gridDim.x : (int)cluster.num_blocks();        // CTAs in the team
    const int rank = GRID ? (int)blockIdx.x : (int)cluster.block_rank();
    const int slot = GRID ? a.first : a.first + (int)(blockIdx.x / C);
    FrameMeta& M = a.meta[slot];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qi = tid / G, sub = tid % G;                   // point of the macro-chunk this thread works for, lane among that point's G

    extern __shared__ __align__(128) unsigned char s_dyn[];
    float4* s_tile = reinterpret_cast<float4*>(s_dyn);                                           // [LM_TILE_PTS]
    int* s_cs = reinterpret_cast<int*>(s_dyn + (size_t)LM_TILE_PTS * 16);                        // [LM_CS_CAP]
    int* s_rowOff = s_cs + LM_CS_CAP;                                                            // [LM_ROW_CAP + 1]
    float (*s_rows)[QPW][7] = reinterpret_cast<float (*)[QPW][7]>(s_cs);                         // aliases s_cs once the scans are done
    __shared__ double s_part[WPB][NACC];
    __shared__ double sh_tot[NACC];
    __shared__ GridDesc sh_gd[2];
    __shared__ float sh_pose[6], sh_T[12], sh_trig[6], sh_AtA[36], sh_AtB[6];
    __shared__ int sh_stop, sh_nsel, s_bbox[2][6], s_wsum[WPB + 1];
    __shared__ unsigned long long s_mbar;

    const int nC = min(M.n_corner_ds, a.cornerCap), nS = min(M.n_surf_ds, a.surfCap);
    if (!(nC > a.edgeMin && nS > a.surfMin)) {     // mapOptmization.h:1410 / :1439-1441
        if (rank == 0 && tid == 0) { M.flags = FBPR_FLAG_NOT_ENOUGH_FEATURES | (M.mapTruncated ? FBPR_FLAG_MAP_TRUNCATED : 0u); M.iters = 0; M.nSel = 0; M.isDegenerate = 0; }
        return;
    }
    const GridSeg gc = a.gsegs[2 * slot], gs = a.gsegs[2 * slot + 1];
    const float4* qpts = a.qpts + (size_t)slot * a.qCap;     // corners [0, nC), surface points [nC, nC + nS), each Morton-ordered (lm_order)
    double* part = GRID ? a.partialsGrid + (size_t)slot * 2 * a.gridMax * NACC : a.partials + (size_t)slot * 2 * a.teamMax * NACC;
    const int teamStride = GRID ? a.gridMax : a.teamMax;
    if (tid < 6) sh_pose[tid] = M.pose[tid];
    if (tid == 32) sh_gd[0] = *gc.desc;
    if (tid == 64) sh_gd[1] = *gs.desc;
    if (tid == 0) mbar_init(&s_mbar, 1);
    // the accumulator entry this lane owns: 0..20 = upper triangle (ei <= ej), 21..26 = (ei, 6) = A^T b, 27 = row count
    int ei = 0, ej = 6;
    if (lane < 21) { int rr = 0, qx = lane; while (qx >= 6 - rr) { qx -= 6 - rr; rr++; } ei = rr; ej = rr + qx; }
    else if (lane < 27) ei = lane - 21;
    __syncthreads();

    KnnMaps maps; maps.gd = sh_gd; maps.cell_start[0] = gc.cell_start; maps.cell_start[1] = gs.cell_start; maps.pts[0] = gc.sorted; maps.pts[1] = gs.sorted;
    const int posBits = a.posBits;
    const unsigned posMask = (1u << posBits) - 1u;           // also the "not in the tile" position
    const int tileCap = min((int)posMask, LM_TILE_PTS);
    // points per macro-chunk: the smallest multiple of a warp's points that needs no more rounds over the team than MC would
    // (balances the CTAs: e.g. one frame on the whole GPU gets exactly one chunk per CTA)
    int mcSize = MC;
    {
        const int rounds = max(1, ((nC + MC - 1) / MC + (nS + MC - 1) / MC + C - 1) / C);
        for (int s = QPW; s < MC; s += QPW) if ((nC + s - 1) / s + (nS + s - 1) / s <= rounds * C) { mcSize = s; break; }
    }
    const int chunksC = (nC + mcSize - 1) / mcSize, chunksS = (nS + mcSize - 1) / mcSize;
    const int nChunks = chunksC + chunksS;
    unsigned mbarPhase = 0;
    int bbSel = 0;
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

        double acc = 0.0;
        for (int mc = rank; mc < nChunks; mc += C) {
            // ---- the macro-chunk's points: corners occupy chunks [0, chunksC), surface points follow
            const bool isCorner = mc < chunksC;
            const int kind = isCorner ? 0 : 1;
            const int li0 = (isCorner ? mc : mc - chunksC) * mcSize;                 // first point of the chunk inside its cloud
            const int nk = isCorner ? nC : nS;
            const int mcCount = min(mcSize, nk - li0);
            const GridDesc& gd = sh_gd[kind];
            const bool inRange = qi < mcCount;
            float4 pOri = make_float4(0.f, 0.f, 0.f, 0.f);
            float x0 = 0.f, y0 = 0.f, z0 = 0.f;
            if (inRange) { pOri = __ldg(qpts + (isCorner ? 0 : nC) + li0 + qi); transform_point(sh_T, pOri, x0, y0, z0); }
            const int li = __float_as_int(pOri.w);                               // index in laserCloud{Corner,Surf}LastDS
            const int cx = (int)floorf((x0 - gd.ox) * gd.inv_h), cy = (int)floorf((y0 - gd.oy) * gd.inv_h), cz = (int)floorf((z0 - gd.oz) * gd.inv_h);
            // a point whose 1 m ball cannot reach the map's bounding box has no neighbour at all (the search would find nothing)
            const bool reach = cx + gd.rmax >= 0 && cx - gd.rmax < gd.dx && cy + gd.rmax >= 0 && cy - gd.rmax < gd.dy && cz + gd.rmax >= 0 && cz - gd.rmax < gd.dz;
            const bool active = inRange && gd.n >= 5 && reach;
            const float4* mo = isCorner ? gc.pts : gs.pts;                       // the map in original order (only the fallback reads it)
            const float4* sorted = isCorner ? gc.sorted : gs.sorted;
            const int* cell_start = isCorner ? gc.cell_start : gs.cell_start;
            // first cube radius: one cell (after the first iteration 99.9 % of the surface points are certified there); the very
            // first iteration starts from knn_first_radius.  The tile is staged with a margin of up to two cells so that a point
            // may grow its cube once without leaving it.
            const int radStart = iter == 0 ? min(max(1, (int)ceilf(a.firstRadius / (gd.h * 0.9995f))), gd.rmax) : 1;
            int margin = min((iter == 0 || isCorner) ? max(2, radStart) : 1, gd.rmax);

            // ---- adaptive sub-passes over [pa, pa + plen) of the chunk's points: the whole chunk if its tile fits
            int pa = 0, plen = MC;
            while (pa < mcCount) {
                const bool mine = active && qi >= pa && qi < pa + plen;
                // (1) box of cells covered by the cubes of radius `margin` (two copies used in turn: a retry re-initialises the box
                //     while slower threads may still be reading the previous one)
                bbSel ^= 1;
                int* s_bb = s_bbox[bbSel];
                if (tid < 3) s_bb[tid] = 0x7fffffff; else if (tid < 6) s_bb[tid] = -0x7fffffff;
                __syncthreads();                                                 // also: everyone is done with the previous tile / s_rows
                {
                    int lo0 = 0x7fffffff, lo1 = 0x7fffffff, lo2 = 0x7fffffff, hi0 = -0x7fffffff, hi1 = -0x7fffffff, hi2 = -0x7fffffff;
                    if (mine) {
                        const int xa = max(cx - margin, 0), xb = min(cx + margin, gd.dx - 1);
                        const int ya = max(cy - margin, 0), yb = min(cy + margin, gd.dy - 1);
                        const int za = max(cz - margin, 0), zb = min(cz + margin, gd.dz - 1);
                        if (xa <= xb && ya <= yb && za <= zb) { lo0 = xa; hi0 = xb; lo1 = ya; hi1 = yb; lo2 = za; hi2 = zb; }
                    }
                    lo0 = __reduce_min_sync(0xffffffffu, lo0); lo1 = __reduce_min_sync(0xffffffffu, lo1); lo2 = __reduce_min_sync(0xffffffffu, lo2);
                    hi0 = __reduce_max_sync(0xffffffffu, hi0); hi1 = __reduce_max_sync(0xffffffffu, hi1); hi2 = __reduce_max_sync(0xffffffffu, hi2);
                    if (lane == 0 && lo0 <= hi0) {
                        atomicMin(&s_bb[0], lo0); atomicMin(&s_bb[1], lo1); atomicMin(&s_bb[2], lo2);
                        atomicMax(&s_bb[3], hi0); atomicMax(&s_bb[4], hi1); atomicMax(&s_bb[5], hi2);
                    }
                }
                __syncthreads();
                TileBox B; B.x0 = s_bb[0]; B.y0 = s_bb[1]; B.z0 = s_bb[2];
                const bool anyBox = s_bb[0] <= s_bb[3];
                const int nx1 = anyBox ? s_bb[3] - s_bb[0] + 2 : 0, ny = anyBox ? s_bb[4] - s_bb[1] + 1 : 0, nz = anyBox ? s_bb[5] - s_bb[2] + 1 : 0;
                B.nx1 = nx1; B.ny = ny;
                const int rows = ny * nz;
                bool fits = rows <= LM_ROW_CAP && (long long)rows * nx1 <= LM_CS_CAP;
                int total = 0;
                if (fits && anyBox) {
                    // (2) cell_start entries of the box, one warp per cell row (coalesced), then the rows' lengths -> tile offsets
                    for (int r = warp; r < rows; r += WPB) {
                        const int zi = r / ny, yi = r - zi * ny;
                        const int* src = cell_start + ((size_t)(B.z0 + zi) * gd.dy + (B.y0 + yi)) * gd.dx + B.x0;
                        for (int xi = lane; xi < nx1; xi += 32) s_cs[r * nx1 + xi] = __ldg(src + xi);
                    }
                    __syncthreads();
                    constexpr int RPT = (LM_ROW_CAP + LM_TPB - 1) / LM_TPB;           // rows per thread in the block scan
                    int len[RPT], sum = 0;
                    #pragma unroll
                    for (int k = 0; k < RPT; k++) {
                        const int r = tid * RPT + k;
                        len[k] = r < rows ? s_cs[r * nx1 + nx1 - 1] - s_cs[r * nx1] : 0;
                        sum += len[k];
                    }
                    int incl = sum;
                    #pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                    if (lane == 31) s_wsum[warp] = incl;
                    __syncthreads();
                    if (tid == 0) { int run = 0; for (int w = 0; w < WPB; w++) { const int v = s_wsum[w]; s_wsum[w] = run; run += v; } s_wsum[WPB] = run; }
                    __syncthreads();
                    total = s_wsum[WPB];
                    fits = total <= tileCap;
                    if (fits) {
                        int run = s_wsum[warp] + incl - sum;
                        #pragma unroll
                        for (int k = 0; k < RPT; k++) { const int r = tid * RPT + k; if (r < rows) s_rowOff[r] = run; run += len[k]; }
                        // (3) one bulk copy per non-empty cell row into the tile; thread 0 arms the barrier with the byte count
                        if (tid == 0 && total > 0) mbar_arrive_expect_tx(&s_mbar, (unsigned)total * 16u);
                        run = s_wsum[warp] + incl - sum;
                        #pragma unroll
                        for (int k = 0; k < RPT; k++) {
                            const int r = tid * RPT + k;
                            if (r < rows && len[k] > 0) bulk_g2s(s_tile + run, sorted + s_cs[r * nx1], (unsigned)len[k] * 16u, &s_mbar);
                            run += len[k];
                        }
                    }
                }
                if (!fits) {                        // uniform decision: fewer points first (down to two warps' worth), then the margin, then fewer points again
                    if (a.stats && tid == 0) atomicAdd(a.stats + 2, 1ull);
                    if (plen > 2 * QPW) { plen >>= 1; continue; }
                    if (margin > 1) { margin = 1; continue; }
                    if (plen > QPW) { plen >>= 1; continue; }
                }
                const bool haveTile = fits && anyBox;
                if (a.stats && tid == 0) {
                    atomicAdd(a.stats + 0, 1ull); atomicAdd(a.stats + 1, (unsigned long long)total);
                    atomicAdd(a.stats + 4, (unsigned long long)rows); atomicAdd(a.stats + 5, (unsigned long long)rows * nx1);
                    if (!haveTile) atomicAdd(a.stats + 6, 1ull);
                }
                if (haveTile) {
                    __syncthreads();                                             // s_rowOff complete
                    if (total > 0) { mbar_wait(&s_mbar, mbarPhase); mbarPhase ^= 1u; }
                }
                // (4) exact 5-NN out of the tile: cube of radStart cells, grown cell by cell while the result is not certified and
                //     the cube stays inside the staged margin
                ThreadKnn5 r;
                #pragma unroll
                for (int k = 0; k < 5; k++) r.key[k] = ~0ull;
                bool need = mine;                                                // still needs a (fallback) search
                int radNext = 1;
                if (mine && haveTile) {
                    int rad = min(radStart, margin);
                    while (true) {
                        #pragma unroll
                        for (int k = 0; k < 5; k++) r.key[k] = ~0ull;
                        tile_scan<G>(r, s_tile, s_cs, s_rowOff, B, gd, cx, cy, cz, rad, sub, posBits, x0, y0, z0);
                        tile_merge<G>(r);
                        const float guard = (float)rad * gd.h * 0.9995f;
                        if (knn_d5(r) < guard * guard || rad >= gd.rmax) { need = false; break; }
                        if (rad >= margin) break;
                        rad++;
                    }
                    radNext = min(rad * 2, gd.rmax);
                }
                __syncthreads();                                                 // every warp is done with s_cs: s_rows may overwrite it
                // (5) fallback: warp-cooperative search on the global index for the points the tile could not certify
                {
                    const bool needW = need && sub == 0;
                    if (a.stats) { const unsigned nb_ = __ballot_sync(0xffffffffu, needW); if (lane == 0 && nb_) atomicAdd(a.stats + 3, (unsigned long long)__popc(nb_)); }
                    if (__any_sync(0xffffffffu, needW)) {
                        ThreadKnn5 rf;
                        #pragma unroll
                        for (int k = 0; k < 5; k++) rf.key[k] = ~0ull;
                        warp_knn5(maps, kind, x0, y0, z0, radNext, needW, rf, nullptr, nullptr);
                        if (G > 1) {
                            #pragma unroll
                            for (int k = 0; k < 5; k++) rf.key[k] = __shfl_sync(0xffffffffu, rf.key[k], lane & ~(G - 1));
                        }
                        if (need) {
                            #pragma unroll
                            for (int k = 0; k < 5; k++)
                                r.key[k] = rf.key[k] == ~0ull ? ~0ull : ((rf.key[k] & 0xffffffff00000000ull) | (((unsigned)rf.key[k] << posBits) | posMask));
                        }
                    }
                }
                // (6) fit, coefficient, Jacobian row: one lane per point
                const bool worker = inRange && qi >= pa && qi < pa + plen && sub == 0;
                bool ok = mine && knn_d5(r) < 1.0f;
                if (worker) {
                    if (cap) {
                        const size_t o = isCorner ? (size_t)slot * a.cornerCap + li : (size_t)slot * a.surfCap + li;
                        int* kd = isCorner ? a.knnC : a.knnS; float* dd = isCorner ? a.d2C : a.d2S;
                        #pragma unroll
                        for (int k = 0; k < 5; k++) {
                            kd[5 * o + k] = ok ? (int)((unsigned)r.key[k] >> posBits) : -1;
                            dd[5 * o + k] = (gd.n >= 5 && r.key[k] != ~0ull) ? __uint_as_float((unsigned)(r.key[k] >> 32)) : 3.0e38f;
                        }
                    }
                    float4 coeff = make_float4(0, 0, 0, 0);
                    if (ok) {
                        // the five neighbours' coordinates (mapOptmization.h:1028-1036, :1157-1163): from the tile, or by original index
                        float4 nb[5];
                        #pragma unroll
                        for (int k = 0; k < 5; k++) {
                            const unsigned low = (unsigned)r.key[k], pos = low & posMask;
                            nb[k] = pos != posMask ? s_tile[pos] : __ldcg(mo + (low >> posBits));
                        }
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
                        float* row = s_rows[warp][lane / G];
                        row[0] = arz; row[1] = arx; row[2] = ary;
                        row[3] = kz;  row[4] = kx;  row[5] = ky;
                        row[6] = -coeff.w;
                    }
                } else {
                    ok = false;
                }
                // fold the warp's staged rows into the lane-owned entries (rows in point order, f64)
                unsigned m = __ballot_sync(0xffffffffu, ok);
                __syncwarp();
                if (lane == 27) acc += (double)__popc(m);
                else if (lane < 27) {
                    while (m) {
                        const int rr = (__ffs(m) - 1) / G; m &= m - 1;
                        acc += (double)s_rows[warp][rr][ei] * (double)s_rows[warp][rr][ej];
                    }
                }
                pa += plen;
            }
            __syncthreads();                         // the next macro-chunk's box reduce reuses s_bb, its tile overwrites s_tile / s_rows
        }
        // --- CTA reduce (shared memory, fixed warp order) -> one partial per CTA in global memory
        if (lane < NACC) s_part[warp][lane] = acc;
        __syncthreads();
        const int buf = iter & 1;
        double* mypart = part + ((size_t)buf * teamStride + rank) * NACC;
        if (tid < NACC) {
            double v = 0.0;
            for (int w = 0; w < WPB; w++) v += s_part[w][tid];
            mypart[tid] = v;
            __threadfence();
        }
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
        if (tid == 0) {
            int stop = 0;
            iters = iter + 1;
            if (sh_nsel < 50) {                     // :1267-1270: LMOptimization returns false, pose unchanged;
                flags |= FBPR_FLAG_TOO_FEW_CORRESPONDENCES;   // every remaining iteration repeats identically
                iters = FBPR_MAX_ITERS; stop = 1;
            } else {
                float pose[6], X[6];
                for (int k = 0; k < 6; k++) pose[k] = sh_pose[k];
                int conv = lm_solve_step(sh_AtA, sh_AtB, iter, isDegenerate, pose, X);
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

// one-time function attributes of both variants (cluster sizes above 8, dynamic shared memory above the default limit)
static int lm_configure() {
    static int configured = 0;
    if (configured) return 0;
    cudaError_t e = cudaFuncSetAttribute(lm_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return fbpr_fail(e, "cudaFuncSetAttribute(NonPortableClusterSizeAllowed)", __FILE__, __LINE__);
    e = cudaFuncSetAttribute(lm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LM_DYN_SMEM);
    if (e != cudaSuccess) return fbpr_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize, batched)", __FILE__, __LINE__);
    e = cudaFuncSetAttribute(lm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LM_DYN_SMEM);
    if (e != cudaSuccess) return fbpr_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize, single frame)", __FILE__, __LINE__);
    // four CTAs of ~53 KB per SM: ask for the whole shared-memory carve-out (the tiles replaced L1 as the neighbour cache)
    cudaFuncSetAttribute(lm_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured = 1;
    return 0;
}

int fbpr_lm_tile_points() { return LM_TILE_PTS; }

int fbpr_lm_grid_blocks(int device) {
    // co-resident CTAs of the cooperative (one frame on the whole GPU) variant: one per SM
    if (lm_configure()) return 0;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
    int per = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, lm_kernel<true>, LM_TPB_GRID, LM_DYN_SMEM) != cudaSuccess || per < 1) return 0;
    return sms;
}

// co-resident clusters of `c` CTAs of the batched variant (0 if the size cannot be launched)
static int lm_max_active_clusters(int c) {
    static int cached[17] = { 0 }, known[17] = { 0 };
    if (c < 1 || c > 16) return 0;
    if (!known[c]) {
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(LM_TPB_CLUSTER); cfg.gridDim = dim3((unsigned)(c * 64)); cfg.dynamicSmemBytes = LM_DYN_SMEM;
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

// cluster size for a batch of `count` frames: the time of the launch is ~ waves / (CTAs per frame), so take the size
// that minimises it (a small batch gets big clusters to fill the GPU, a big batch small ones for fewer waves);
// ties go to the smaller cluster (cheaper barrier, fewer redundant solves)
int fbpr_lm_auto_cluster(int count) {
    const int sizes[5] = { 1, 2, 4, 8, 16 };
    int best = 8; double bestScore = 1e30;
    for (int k = 0; k < 5; k++) {
        const int c = sizes[k], n = lm_max_active_clusters(c);
        if (n <= 0) continue;
        const double score = (double)((count + n - 1) / n) / (double)c;
        if (score < bestScore * 0.999) { bestScore = score; best = c; }
    }
    return best;
}

int fbpr_launch_lm(const LmArgs& args, int count, int cluster_size, int grid_blocks, cudaStream_t st, long long* launches) {
    if (count <= 0) return 0;
    { int rc = lm_configure(); if (rc) return rc; }
    // spatial order of the down-sampled feature points (once per frame, from the initial guess)
    {
        const int maxPts = args.cornerCap > args.surfCap ? args.cornerCap : args.surfCap;
        lm_order<<<dim3((unsigned)((maxPts + LM_ORDER_WIN - 1) / LM_ORDER_WIN), 2, (unsigned)count), LM_ORDER_TPB, 0, st>>>(args);
        if (launches) *launches += 1;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.dynamicSmemBytes = LM_DYN_SMEM;
    cudaError_t e;
    if (count == 1 && grid_blocks > 0) {           // one frame: the whole GPU cooperates, grid-wide barrier per iteration
        cfg.blockDim = dim3(LM_TPB_GRID);
        cfg.gridDim = dim3((unsigned)grid_blocks);
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        e = cudaLaunchKernelEx(&cfg, lm_kernel<true>, args);
    } else {                                        // many frames: one cluster per frame, hardware cluster barrier per iteration
        if (cluster_size <= 0) cluster_size = fbpr_lm_auto_cluster(count);
        cfg.blockDim = dim3(LM_TPB_CLUSTER);
        cfg.gridDim = dim3((unsigned)(count * cluster_size));
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cluster_size; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        e = cudaLaunchKernelEx(&cfg, lm_kernel<false>, args);
    }
    if (e != cudaSuccess) return fbpr_fail(e, "cudaLaunchKernelEx(lm_kernel)", __FILE__, __LINE__);
    if (launches) *launches += 1;
    return fbpr_launch_ok("lm_order / lm_kernel");
}

int fbpr_launch_transform_update(FrameMeta* meta, int first, int count, float rot_tol, float z_tol, cudaStream_t st, long long* launches) {
    if (count <= 0) return 0;
    transform_update_kernel<<<(count + 63) / 64, 64, 0, st>>>(meta, first, count, rot_tol, z_tol);
    if (launches) *launches += 1;
    return fbpr_launch_ok("transform_update_kernel");
}
