// TEST INFRASTRUCTURE -- CPU oracle (see ref_smallmat.hpp header).
//
// oracle_capi.cpp -- extern "C" surface over the restatement so that tests/ and bench.py's
// cpu_baseline leg can drive it through ctypes.  Nothing in the product links this.
#include <cstdio>
#include <cstring>

#include "ref_pipeline.hpp"

using namespace orc;

// memcpy that is a no-op for empty ranges (an empty std::vector's data() may be null, which memcpy must not be given)
static inline void cpy(void* dst, const void* src, size_t bytes) { if (bytes) std::memcpy(dst, src, bytes); }

extern "C" {

struct orc_params {
    int N_SCAN, Horizon_SCAN;
    float edgeThreshold, surfThreshold;
    int edgeFeatureMinValidNum, surfFeatureMinValidNum;
    float odometrySurfLeafSize, mappingCornerLeafSize, mappingSurfLeafSize;
    float z_tollerance, rotation_tollerance;
    int numberOfCores;
    float surroundingKeyframeSearchRadius;
};

static Params to_params(const orc_params* p) {
    Params P;
    P.N_SCAN = p->N_SCAN; P.Horizon_SCAN = p->Horizon_SCAN;
    P.edgeThreshold = p->edgeThreshold; P.surfThreshold = p->surfThreshold;
    P.edgeFeatureMinValidNum = p->edgeFeatureMinValidNum; P.surfFeatureMinValidNum = p->surfFeatureMinValidNum;
    P.odometrySurfLeafSize = p->odometrySurfLeafSize; P.mappingCornerLeafSize = p->mappingCornerLeafSize;
    P.mappingSurfLeafSize = p->mappingSurfLeafSize;
    P.z_tollerance = p->z_tollerance; P.rotation_tollerance = p->rotation_tollerance;
    P.numberOfCores = p->numberOfCores; P.surroundingKeyframeSearchRadius = p->surroundingKeyframeSearchRadius;
    return P;
}

// ---------------------------------------------------------------- small matrices
void orc_eigen_sym(int n, const float* A, float* W, float* V) {
    float Aw[36]; cpy(Aw, A, sizeof(float) * n * n);
    jacobi_eigen_sym(n, Aw, W, V);
}
int orc_qr_solve(int n, const float* A, const float* b, float* x) {
    float Aw[36], bw[6]; cpy(Aw, A, sizeof(float) * n * n); cpy(bw, b, sizeof(float) * n);
    return qr_solve(n, Aw, bw, x);
}
int orc_lu_invert(int n, const float* A, float* Ainv) {
    float Aw[36]; cpy(Aw, A, sizeof(float) * n * n);
    return lu_invert(n, Aw, Ainv);
}
void orc_matmul_f64acc(int r, int k, int c, const float* A, const float* B, float* C) { matmul_f64acc(r, k, c, A, B, C); }
void orc_colpiv_solve_5x3(const float* A, const float* b, float* x) { colpiv_householder_solve_5x3(A, b, x); }
void orc_get_transformation(const float pose6[6], float T[12]) {
    get_transformation(pose6[3], pose6[4], pose6[5], pose6[0], pose6[1], pose6[2], T);
}
void orc_get_translation_and_euler(const float T[12], float pose6[6]) {
    get_translation_and_euler(T, pose6[3], pose6[4], pose6[5], pose6[0], pose6[1], pose6[2]);
}

// ---------------------------------------------------------------- cloud primitives
int orc_voxel_grid(const float* xyzi, int n, float leaf, float* out_xyzi, int* point_keys, int* out_keys, int* overflow) {
    std::vector<P4> out; std::vector<int> pk, ok;
    int m = voxel_grid(reinterpret_cast<const P4*>(xyzi), n, leaf, out, point_keys ? &pk : nullptr, out_keys ? &ok : nullptr, overflow);
    cpy(out_xyzi, out.data(), sizeof(P4) * m);
    if (point_keys && !pk.empty()) cpy(point_keys, pk.data(), sizeof(int) * n);
    if (out_keys && !ok.empty()) cpy(out_keys, ok.data(), sizeof(int) * ok.size());
    return m;
}
int orc_crop_box(const float* xyzi, int n, const float mn[3], const float mx[3], float* out_xyzi) {
    std::vector<P4> out; crop_box(reinterpret_cast<const P4*>(xyzi), n, mn, mx, out);
    cpy(out_xyzi, out.data(), sizeof(P4) * out.size());
    return (int)out.size();
}
void orc_kdtree_knn5(const float* map_xyzi, int M, const float* q_xyz, int nq, int* idx, float* d2, int threads) {
    KdTree5 t; t.build(reinterpret_cast<const P4*>(map_xyzi), M);
    #pragma omp parallel for num_threads(threads)
    for (int i = 0; i < nq; i++) t.knn5(q_xyz + 3 * (size_t)i, idx + 5 * (size_t)i, d2 + 5 * (size_t)i);
}
void orc_brute_knn5(const float* map_xyzi, int M, const float* q_xyz, int nq, int* idx, float* d2, int threads) {
    #pragma omp parallel for num_threads(threads)
    for (int i = 0; i < nq; i++) brute_knn5(reinterpret_cast<const P4*>(map_xyzi), M, q_xyz + 3 * (size_t)i, idx + 5 * (size_t)i, d2 + 5 * (size_t)i);
}

// ---------------------------------------------------------------- projection
// outputs sized N_SCAN*Horizon_SCAN (cloud: x4 floats); returns N_v
int orc_project(const orc_params* p, const float* x, const float* y, const float* z, const float* intensity,
                const int32_t* ring, const float* time, int n_raw,
                int64_t imuAvailable, int deskewFlag, double timeScanCur,
                const double* imuTime, const double* imuRotX, const double* imuRotY, const double* imuRotZ, int imuPointerCur,
                int* startRingIndex, int* endRingIndex, int* pointColInd, float* pointRange, float* cloud_xyzi, int* winner_raw) {
    Params P = to_params(p);
    RawScan raw{ x, y, z, intensity, ring, time, n_raw };
    ImuRamp imu{ imuTime, imuRotX, imuRotY, imuRotZ, imuPointerCur, timeScanCur };
    CloudInfo ci; std::vector<int> win;
    project(P, raw, imuAvailable, deskewFlag, imu, ci, winner_raw ? &win : nullptr);
    int nv = (int)ci.cloud_deskewed.size();
    cpy(startRingIndex, ci.startRingIndex.data(), sizeof(int) * P.N_SCAN);
    cpy(endRingIndex, ci.endRingIndex.data(), sizeof(int) * P.N_SCAN);
    cpy(pointColInd, ci.pointColInd.data(), sizeof(int) * nv);
    cpy(pointRange, ci.pointRange.data(), sizeof(float) * nv);
    cpy(cloud_xyzi, ci.cloud_deskewed.data(), sizeof(P4) * nv);
    if (winner_raw) cpy(winner_raw, win.data(), sizeof(int) * nv);
    return nv;
}

// ---------------------------------------------------------------- features
// corner_xyzi / surface_xyzi / *_index sized >= n_valid.  counts[0]=corners, [1]=surface DS, [2]=surface raw
void orc_extract_features(const orc_params* p, const int* startRingIndex, const int* endRingIndex,
                          const int* pointColInd, const float* pointRange, const float* cloud_xyzi, int n_valid,
                          float* corner_xyzi, int* corner_index, float* surface_xyzi, int* surface_raw_index,
                          int* ring_surf_count, int* ring_surf_count_ds,
                          float* curvature, int* picked, int* label, int* counts) {
    Params P = to_params(p);
    CloudInfo ci;
    ci.startRingIndex.assign(startRingIndex, startRingIndex + P.N_SCAN);
    ci.endRingIndex.assign(endRingIndex, endRingIndex + P.N_SCAN);
    ci.pointColInd.assign(pointColInd, pointColInd + n_valid);
    ci.pointRange.assign(pointRange, pointRange + n_valid);
    ci.cloud_deskewed.assign(reinterpret_cast<const P4*>(cloud_xyzi), reinterpret_cast<const P4*>(cloud_xyzi) + n_valid);
    FeatureOut fo; extract_features(P, ci, fo);
    counts[0] = (int)fo.cornerCloud.size(); counts[1] = (int)fo.surfaceCloud.size(); counts[2] = (int)fo.surfaceRawIndex.size();
    if (corner_xyzi) cpy(corner_xyzi, fo.cornerCloud.data(), sizeof(P4) * fo.cornerCloud.size());
    if (corner_index) cpy(corner_index, fo.cornerIndex.data(), sizeof(int) * fo.cornerIndex.size());
    if (surface_xyzi) cpy(surface_xyzi, fo.surfaceCloud.data(), sizeof(P4) * fo.surfaceCloud.size());
    if (surface_raw_index) cpy(surface_raw_index, fo.surfaceRawIndex.data(), sizeof(int) * fo.surfaceRawIndex.size());
    if (ring_surf_count) cpy(ring_surf_count, fo.surfaceRingCount.data(), sizeof(int) * P.N_SCAN);
    if (ring_surf_count_ds) cpy(ring_surf_count_ds, fo.surfaceRingCountDS.data(), sizeof(int) * P.N_SCAN);
    if (curvature) cpy(curvature, fo.cloudCurvature.data(), sizeof(float) * n_valid);
    if (picked) cpy(picked, fo.cloudNeighborPicked.data(), sizeof(int) * n_valid);
    if (label) cpy(label, fo.cloudLabel.data(), sizeof(int) * n_valid);
}

// ---------------------------------------------------------------- mapOptimization handle
struct orc_mo { MapOptimization mo; IterDebug dbg; };

orc_mo* orc_mo_create(const orc_params* p) { auto* h = new orc_mo(); h->mo.P = to_params(p); return h; }
void orc_mo_destroy(orc_mo* h) { delete h; }
void orc_mo_set_threads(orc_mo* h, int n) { h->mo.P.numberOfCores = n; }
void orc_mo_set_scan(orc_mo* h, const float* corner, int nC, const float* surf, int nS) {
    h->mo.laserCloudCornerLast.assign(reinterpret_cast<const P4*>(corner), reinterpret_cast<const P4*>(corner) + nC);
    h->mo.laserCloudSurfLast.assign(reinterpret_cast<const P4*>(surf), reinterpret_cast<const P4*>(surf) + nS);
}
void orc_mo_set_map(orc_mo* h, const float* corner, int nC, const float* surf, int nS) {
    h->mo.laserCloudCornerFromMapDS.assign(reinterpret_cast<const P4*>(corner), reinterpret_cast<const P4*>(corner) + nC);
    h->mo.laserCloudSurfFromMapDS.assign(reinterpret_cast<const P4*>(surf), reinterpret_cast<const P4*>(surf) + nS);
}
void orc_mo_set_imu(orc_mo* h, int64_t imuAvailable, float imuRollInit, float imuPitchInit) {
    h->mo.imuAvailable = imuAvailable; h->mo.imuRollInit = imuRollInit; h->mo.imuPitchInit = imuPitchInit;
}
// 1: the path's two std::sort calls compare only what the reference compares (see oracle_literal_sort); returns the old value
int orc_set_literal_sort(int on) { int old = oracle_literal_sort(); oracle_literal_sort() = on ? 1 : 0; return old; }
// kd-tree split rule: 1 = FLANN middleSplit_ (default), 0 = median; returns the old value
int orc_set_kdtree_flann_split(int on) { int old = oracle_kdtree_flann_split(); oracle_kdtree_flann_split() = on ? 1 : 0; return old; }
// keyframe clouds are given concatenated with CSR offsets (K+1 entries)
void orc_mo_extract_cloud(orc_mo* h, const float* keyPoses6, int K, const float* corner_all, const int* corner_off,
                          const float* surf_all, const int* surf_off, const float* lastKeyXYZ, int* counts) {
    std::vector<const P4*> cf(K), sf(K); std::vector<int> cn(K), sn(K);
    for (int i = 0; i < K; i++) {
        cf[i] = reinterpret_cast<const P4*>(corner_all) + corner_off[i]; cn[i] = corner_off[i + 1] - corner_off[i];
        sf[i] = reinterpret_cast<const P4*>(surf_all) + surf_off[i];     sn[i] = surf_off[i + 1] - surf_off[i];
    }
    h->mo.extractCloud(keyPoses6, K, cf.data(), cn.data(), sf.data(), sn.data(), lastKeyXYZ);
    counts[0] = (int)h->mo.laserCloudCornerFromMap.size(); counts[1] = (int)h->mo.laserCloudSurfFromMap.size();
    counts[2] = (int)h->mo.laserCloudCornerFromMapDS.size(); counts[3] = (int)h->mo.laserCloudSurfFromMapDS.size();
}
// extractSurroundingKeyFrames as the reference runs it: extractNearby (:872-907) + extractCloud (:909-955) over the whole
// keyframe store (key pose i = row i of keyPoses6All, cloudKeyPoses3D[i].intensity = i).  ds_out = surroundingKeyPosesDS.
// loopClosureEnableFlag != 0: extractForLoopClosure (:857-870) with surroundingKeyframeSize instead of extractNearby (:970-977).
int orc_mo_extract_surrounding(orc_mo* h, const float* keyPoses6All, const double* keyTime, int nKeys, float density, double timeLast,
                               const float* corner_all, const int* corner_off, const float* surf_all, const int* surf_off,
                               float* ds_out, int ds_cap, int* counts, int loopClosureEnableFlag, int surroundingKeyframeSize) {
    std::vector<P4> k3(nKeys);
    for (int i = 0; i < nKeys; i++) k3[i] = P4{ keyPoses6All[6 * i + 3], keyPoses6All[6 * i + 4], keyPoses6All[6 * i + 5], (float)i };
    std::vector<P4> ds;
    if (loopClosureEnableFlag) MapOptimization::extractForLoopClosure(k3.data(), nKeys, surroundingKeyframeSize, ds);
    else MapOptimization::extractNearby(k3.data(), keyTime, nKeys, h->mo.P.surroundingKeyframeSearchRadius, density, timeLast, ds);
    std::vector<const P4*> cf(nKeys), sf(nKeys); std::vector<int> cn(nKeys), sn(nKeys);
    for (int i = 0; i < nKeys; i++) {
        cf[i] = reinterpret_cast<const P4*>(corner_all) + corner_off[i]; cn[i] = corner_off[i + 1] - corner_off[i];
        sf[i] = reinterpret_cast<const P4*>(surf_all) + surf_off[i];     sn[i] = surf_off[i + 1] - surf_off[i];
    }
    h->mo.extractCloudIndexed(ds, keyPoses6All, nKeys, cf.data(), cn.data(), sf.data(), sn.data());
    counts[0] = (int)h->mo.laserCloudCornerFromMap.size(); counts[1] = (int)h->mo.laserCloudSurfFromMap.size();
    counts[2] = (int)h->mo.laserCloudCornerFromMapDS.size(); counts[3] = (int)h->mo.laserCloudSurfFromMapDS.size();
    const int m = (int)ds.size();
    for (int i = 0; i < m && i < ds_cap; i++) cpy(ds_out + 4 * i, &ds[i], 16);
    return m;
}
// ImageProjection::imuDeskewInfo (imageProjection.cpp:323-393); out5 = imuAvailable, imuPointerCur, roll, pitch, yaw
int orc_imu_deskew_info(const double* q8, int nq, double timeScanCur, double timeScanNext, int queueLength,
                        double* imuTime, double* imuRotX, double* imuRotY, double* imuRotZ, double* out5) {
    ImuDeskewOut o;
    int popped = imu_deskew_info(q8, nq, timeScanCur, timeScanNext, queueLength, imuTime, imuRotX, imuRotY, imuRotZ, &o);
    out5[0] = (double)o.imuAvailable; out5[1] = o.imuPointerCur; out5[2] = o.roll; out5[3] = o.pitch; out5[4] = o.yaw;
    return popped;
}
void orc_mo_downsample(orc_mo* h, int* counts) {
    h->mo.downsampleCurrentScan();
    counts[0] = (int)h->mo.laserCloudCornerLastDS.size(); counts[1] = (int)h->mo.laserCloudSurfLastDS.size();
}
// which: 0 cornerLastDS, 1 surfLastDS, 2 cornerFromMapDS, 3 surfFromMapDS
int orc_mo_get_cloud(orc_mo* h, int which, float* out, int cap) {
    const std::vector<P4>* v = which == 0 ? &h->mo.laserCloudCornerLastDS : which == 1 ? &h->mo.laserCloudSurfLastDS
                             : which == 2 ? &h->mo.laserCloudCornerFromMapDS : &h->mo.laserCloudSurfFromMapDS;
    int n = (int)v->size(); if (out && n <= cap) cpy(out, v->data(), sizeof(P4) * n);
    return n;
}
void orc_mo_scan2map(orc_mo* h, float pose6[6], int debug_iter, int* iters, unsigned* flags, double* seconds /*[2] build, loop*/) {
    cpy(h->mo.transformTobeMapped, pose6, sizeof(float) * 6);
    h->mo.debug = debug_iter >= 0 ? &h->dbg : nullptr; h->mo.debugIter = debug_iter; h->dbg.iter = -1;
    h->mo.scan2MapOptimization();
    cpy(pose6, h->mo.transformTobeMapped, sizeof(float) * 6);
    if (iters) *iters = h->mo.itersDone;
    if (flags) *flags = h->mo.flags;
    if (seconds) { seconds[0] = h->mo.buildSeconds; seconds[1] = h->mo.loopSeconds; }
}
void orc_mo_transform_update(orc_mo* h, float pose6[6]) {
    cpy(h->mo.transformTobeMapped, pose6, sizeof(float) * 6);
    h->mo.transformUpdate();
    cpy(pose6, h->mo.transformTobeMapped, sizeof(float) * 6);
}
void orc_mo_registration(orc_mo* h, const float* corner_global, int nCg, const float* surf_global, int nSg, float pose12[12],
                         int* iters, unsigned* flags) {
    h->mo.debug = nullptr;
    h->mo.registration(reinterpret_cast<const P4*>(corner_global), nCg, reinterpret_cast<const P4*>(surf_global), nSg, pose12);
    if (iters) *iters = h->mo.itersDone;
    if (flags) *flags = h->mo.flags;
}
int orc_mo_pose_trace(orc_mo* h, float* out, int cap_iters) {
    int n = (int)h->mo.poseTrace.size() / 6;
    if (out) cpy(out, h->mo.poseTrace.data(), sizeof(float) * 6 * (n < cap_iters ? n : cap_iters));
    return n;
}
// debug capture of the iteration chosen in orc_mo_scan2map
int orc_mo_debug(orc_mo* h, int* cornerKnn, float* cornerD2, float* cornerCoeff, uint8_t* cornerFlag,
                 int* surfKnn, float* surfD2, float* surfCoeff, uint8_t* surfFlag, float* AtA, float* AtB, float* X, int* nSel) {
    const IterDebug& d = h->dbg;
    if (d.iter < 0) return -1;
    if (cornerKnn) cpy(cornerKnn, d.cornerKnn.data(), sizeof(int) * d.cornerKnn.size());
    if (cornerD2) cpy(cornerD2, d.cornerD2.data(), sizeof(float) * d.cornerD2.size());
    if (cornerCoeff) cpy(cornerCoeff, d.cornerCoeff.data(), sizeof(P4) * d.cornerCoeff.size());
    if (cornerFlag) cpy(cornerFlag, d.cornerFlag.data(), d.cornerFlag.size());
    if (surfKnn) cpy(surfKnn, d.surfKnn.data(), sizeof(int) * d.surfKnn.size());
    if (surfD2) cpy(surfD2, d.surfD2.data(), sizeof(float) * d.surfD2.size());
    if (surfCoeff) cpy(surfCoeff, d.surfCoeff.data(), sizeof(P4) * d.surfCoeff.size());
    if (surfFlag) cpy(surfFlag, d.surfFlag.data(), d.surfFlag.size());
    if (AtA) cpy(AtA, d.AtA, sizeof(float) * 36);
    if (AtB) cpy(AtB, d.AtB, sizeof(float) * 6);
    if (X) cpy(X, d.X, sizeof(float) * 6);
    if (nSel) *nSel = d.nSel;
    return d.iter;
}

}  // extern "C"
