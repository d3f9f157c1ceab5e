// TEST INFRASTRUCTURE -- CPU oracle (see ref_smallmat.hpp header).
//
// ref_pipeline.hpp -- ROS-free restatement of the reference's hot path, structure lifted
// from (same loop order, same float/double promotions, same expression association):
//   projection   : src/imageProjection.cpp:494-526 (findRotation), :545-580 (deskewPoint),
//                  :583-640 (projectPointCloud), :642-670 (cloudExtraction)
//   features     : src/featureExtraction.h:109-131, :134-176, :178-294
//   registration : src/mapOptmization.h:263-343, :397-425, :909-955, :981-1000, :1002-1489
// Oracle contract where the reference has UB or implementation-defined order
// (hard parts 5 and 6): fresh zero-initialised scratch per frame,
// cloudSmoothness slots never written hold {0.0f, ind 0}, an index < 0 breaks the
// suppression loop, all sorts are total orders with the point index as the last key,
// f32 trig = (float)trig((double)x), A^T A / A^T b accumulate in f64 and round once.
#pragma once
#include <omp.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <vector>

#include "ref_cloud.hpp"
#include "ref_smallmat.hpp"

namespace orc {

struct Params {                       // include/utility.h:146-212, values of config/params.yaml
    int N_SCAN = 16, Horizon_SCAN = 1800;
    float edgeThreshold = 1.0f, surfThreshold = 0.1f;
    int edgeFeatureMinValidNum = 10, surfFeatureMinValidNum = 100;
    float odometrySurfLeafSize = 0.4f, mappingCornerLeafSize = 0.2f, mappingSurfLeafSize = 0.4f;
    float z_tollerance = 1000.f, rotation_tollerance = 1000.f;
    int numberOfCores = 4;
    float surroundingKeyframeSearchRadius = 50.f;
};

struct CloudInfo {                    // msg/cloud_info.msg:1-34 as a POD
    std::vector<int> startRingIndex, endRingIndex, pointColInd;
    std::vector<float> pointRange;
    std::vector<P4> cloud_deskewed;
    int64_t imuAvailable = 0;
    float imuRollInit = 0, imuPitchInit = 0, imuYawInit = 0;
};

// ===================================================================== projection
struct RawScan { const float *x, *y, *z, *intensity; const int32_t* ring; const float* time; int n; };
struct ImuRamp { const double *imuTime, *imuRotX, *imuRotY, *imuRotZ; int imuPointerCur; double timeScanCur; };

static inline void find_rotation(const ImuRamp& imu, double pointTime, float* rx, float* ry, float* rz) {
    *rx = 0; *ry = 0; *rz = 0;
    int front = 0;
    while (front < imu.imuPointerCur) { if (pointTime < imu.imuTime[front]) break; ++front; }
    if (pointTime > imu.imuTime[front] || front == 0) {
        *rx = (float)imu.imuRotX[front]; *ry = (float)imu.imuRotY[front]; *rz = (float)imu.imuRotZ[front];
    } else {
        int back = front - 1;
        double ratioFront = (pointTime - imu.imuTime[back]) / (imu.imuTime[front] - imu.imuTime[back]);
        double ratioBack = (imu.imuTime[front] - pointTime) / (imu.imuTime[front] - imu.imuTime[back]);
        *rx = (float)(imu.imuRotX[front] * ratioFront + imu.imuRotX[back] * ratioBack);
        *ry = (float)(imu.imuRotY[front] * ratioFront + imu.imuRotY[back] * ratioBack);
        *rz = (float)(imu.imuRotZ[front] * ratioFront + imu.imuRotZ[back] * ratioBack);
    }
}

// Inverse of a rigid 3x4 whose translation is zero here (findPosition returns 0,
// imageProjection.cpp:528-542).  Eigen's Affine inverse inverts the 3x3 by cofactors; for a
// rotation the oracle DEFINES it as the cofactor/determinant form below (unpinned).
static inline void affine_inverse(const float T[12], float Ti[12]) {
    const float a = T[0], b = T[1], c = T[2], d = T[4], e = T[5], f = T[6], g = T[8], h = T[9], i = T[10];
    float c00 = e * i - f * h, c01 = f * g - d * i, c02 = d * h - e * g;
    float det = a * c00 + b * c01 + c * c02;
    float inv = 1.0f / det;
    float M[9];
    M[0] = c00 * inv;             M[1] = (c * h - b * i) * inv; M[2] = (b * f - c * e) * inv;
    M[3] = c01 * inv;             M[4] = (a * i - c * g) * inv; M[5] = (c * d - a * f) * inv;
    M[6] = c02 * inv;             M[7] = (b * g - a * h) * inv; M[8] = (a * e - b * d) * inv;
    for (int r = 0; r < 3; r++) {
        Ti[4 * r] = M[3 * r]; Ti[4 * r + 1] = M[3 * r + 1]; Ti[4 * r + 2] = M[3 * r + 2];
        Ti[4 * r + 3] = -(M[3 * r] * T[3] + M[3 * r + 1] * T[7] + M[3 * r + 2] * T[11]);
    }
}
static inline void affine_mul(const float A[12], const float B[12], float C[12]) {
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) C[4 * r + c] = A[4 * r] * B[c] + A[4 * r + 1] * B[4 + c] + A[4 * r + 2] * B[8 + c];
        C[4 * r + 3] = A[4 * r] * B[3] + A[4 * r + 1] * B[7] + A[4 * r + 2] * B[11] + A[4 * r + 3];
    }
}

// projectPointCloud + cloudExtraction.  deskewFlag: 1 when the cloud has a time field.
static inline void project(const Params& P, const RawScan& raw, int64_t imuAvailable, int deskewFlag,
                           const ImuRamp& imu, CloudInfo& out, std::vector<int>* winner_raw_index = nullptr) {
    const int H = P.Horizon_SCAN, N = P.N_SCAN;
    std::vector<float> rangeMat((size_t)N * H, FLT_MAX);
    std::vector<P4> fullCloud((size_t)N * H, P4{ 0, 0, 0, 0 });
    std::vector<int> winner((size_t)N * H, -1);
    bool firstPointFlag = true;
    float transStartInverse[12] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0 };
    for (int i = 0; i < raw.n; ++i) {
        P4 thisPoint{ raw.x[i], raw.y[i], raw.z[i], raw.intensity[i] };
        int rowIdn = raw.ring[i];
        if (rowIdn < 0 || rowIdn >= N) continue;
        float horizonAngle = (float)((double)(atan2f_c(thisPoint.x, thisPoint.y) * 180) / M_PI);
        float ang_res_x = (float)(360.0 / (double)(float)H);
        int columnIdn = (int)(-std::round(((double)horizonAngle - 90.0) / (double)ang_res_x) + (double)(H / 2));
        if (columnIdn >= H) columnIdn -= H;
        if (columnIdn < 0 || columnIdn >= H) continue;
        float range = std::sqrt(thisPoint.x * thisPoint.x + thisPoint.y * thisPoint.y + thisPoint.z * thisPoint.z);
        if ((double)range < 1.0) continue;
        if (rangeMat[(size_t)rowIdn * H + columnIdn] != FLT_MAX) continue;
        rangeMat[(size_t)rowIdn * H + columnIdn] = range;
        // deskewPoint
        if (!(deskewFlag == -1 || imuAvailable == 0)) {
            double pointTime = imu.timeScanCur + (double)raw.time[i];
            float rx, ry, rz; find_rotation(imu, pointTime, &rx, &ry, &rz);
            float T[12]; get_transformation(0.f, 0.f, 0.f, rx, ry, rz, T);
            if (firstPointFlag) { affine_inverse(T, transStartInverse); firstPointFlag = false; }
            float Bt[12]; affine_mul(transStartInverse, T, Bt);
            P4 np;
            np.x = Bt[0] * thisPoint.x + Bt[1] * thisPoint.y + Bt[2] * thisPoint.z + Bt[3];
            np.y = Bt[4] * thisPoint.x + Bt[5] * thisPoint.y + Bt[6] * thisPoint.z + Bt[7];
            np.z = Bt[8] * thisPoint.x + Bt[9] * thisPoint.y + Bt[10] * thisPoint.z + Bt[11];
            np.i = thisPoint.i;
            thisPoint = np;
        }
        fullCloud[(size_t)columnIdn + (size_t)rowIdn * H] = thisPoint;
        winner[(size_t)columnIdn + (size_t)rowIdn * H] = i;
    }
    // cloudExtraction
    out.startRingIndex.assign(N, 0); out.endRingIndex.assign(N, 0);
    out.pointColInd.clear(); out.pointRange.clear(); out.cloud_deskewed.clear();
    if (winner_raw_index) winner_raw_index->clear();
    out.imuAvailable = imuAvailable;
    int count = 0;
    for (int i = 0; i < N; ++i) {
        out.startRingIndex[i] = count - 1 + 5;
        for (int j = 0; j < H; ++j) {
            if (rangeMat[(size_t)i * H + j] != FLT_MAX) {
                out.pointColInd.push_back(j);
                out.pointRange.push_back(rangeMat[(size_t)i * H + j]);
                out.cloud_deskewed.push_back(fullCloud[(size_t)j + (size_t)i * H]);
                if (winner_raw_index) winner_raw_index->push_back(winner[(size_t)j + (size_t)i * H]);
                ++count;
            }
        }
        out.endRingIndex[i] = count - 1 - 5;
    }
}

// ImageProjection::imuDeskewInfo  imageProjection.cpp:323-393 (queue of sensor_msgs::Imu as 8 doubles per sample:
// stamp, angular velocity xyz, orientation xyzw).  Returns how many samples the reference pops from the queue front.
struct ImuDeskewOut { int64_t imuAvailable = 0; int imuPointerCur = 0; float roll = 0, pitch = 0, yaw = 0; };
static inline void tf_msg_quat_rpy(double x, double y, double z, double w, double* roll, double* pitch, double* yaw) {
    // tf::quaternionMsgToTF (normalise when |len^2 - 1| > 0.1) + tf::Matrix3x3(q).getRPY  (utility.h:293-303)
    double l2 = x * x + y * y + z * z + w * w;
    if (std::fabs(l2 - 1.0) > 0.1) { double l = std::sqrt(l2); x /= l; y /= l; z /= l; w /= l; l2 = x * x + y * y + z * z + w * w; }
    const double s = 2.0 / l2;
    const double xs = x * s, ys = y * s, zs = z * s, wx = w * xs, wy = w * ys, wz = w * zs;
    const double xx = x * xs, xy = x * ys, xz = x * zs, yy = y * ys, yz = y * zs, zz = z * zs;
    const double m00 = 1.0 - (yy + zz), m01 = xy - wz, m02 = xz + wy, m10 = xy + wz, m20 = xz - wy, m21 = yz + wx, m22 = 1.0 - (xx + yy);
    if (std::fabs(m20) >= 1) {
        *yaw = 0; *roll = std::atan2(m01, m02); *pitch = m20 < 0 ? M_PI / 2.0 : -M_PI / 2.0;
    } else {
        *pitch = -std::asin(m20);
        *roll = std::atan2(m21 / std::cos(*pitch), m22 / std::cos(*pitch));
        *yaw = std::atan2(m10 / std::cos(*pitch), m00 / std::cos(*pitch));
    }
}
static inline int imu_deskew_info(const double* q8, int nq, double timeScanCur, double timeScanNext, int queueLength,
                                  double* imuTime, double* imuRotX, double* imuRotY, double* imuRotZ, ImuDeskewOut* out) {
    *out = ImuDeskewOut();
    int popped = 0;
    while (popped < nq) { if (q8[8 * popped] < timeScanCur - 0.01) ++popped; else break; }
    if (popped == nq) return popped;
    int cur = 0;
    for (int i = popped; i < nq && cur < queueLength; ++i) {
        const double* m = q8 + 8 * i;
        const double t = m[0];
        if (t <= timeScanCur) {
            double r, p, y; tf_msg_quat_rpy(m[4], m[5], m[6], m[7], &r, &p, &y);
            out->roll = (float)r; out->pitch = (float)p; out->yaw = (float)y;
        }
        if (t > timeScanNext + 0.01) break;
        if (cur == 0) { imuRotX[0] = 0; imuRotY[0] = 0; imuRotZ[0] = 0; imuTime[0] = t; ++cur; continue; }
        const double timeDiff = t - imuTime[cur - 1];
        imuRotX[cur] = imuRotX[cur - 1] + m[1] * timeDiff;
        imuRotY[cur] = imuRotY[cur - 1] + m[2] * timeDiff;
        imuRotZ[cur] = imuRotZ[cur - 1] + m[3] * timeDiff;
        imuTime[cur] = t;
        ++cur;
    }
    --cur;
    out->imuPointerCur = cur;
    if (cur <= 0) return popped;
    out->imuAvailable = 1;
    return popped;
}

// ===================================================================== features
struct FeatureOut {
    std::vector<P4> cornerCloud, surfaceCloud;        // surfaceCloud = per-ring VoxelGrid output, rings concatenated
    std::vector<int> cornerIndex;                     // index into cloud_deskewed of every corner, in push order
    std::vector<int> surfaceRawIndex;                 // index of every surface candidate (before the per-ring VoxelGrid)
    std::vector<int> surfaceRingCount;                // per ring: surface candidates / voxel outputs
    std::vector<int> surfaceRingCountDS;
    std::vector<float> cloudCurvature;
    std::vector<int> cloudNeighborPicked, cloudLabel;
};

static inline void extract_features(const Params& P, const CloudInfo& ci, FeatureOut& fo) {
    const int cap = P.N_SCAN * P.Horizon_SCAN;
    const int cloudSize = (int)ci.cloud_deskewed.size();
    struct Smooth { float value; int ind; };
    std::vector<Smooth> cloudSmoothness(cap, Smooth{ 0.f, 0 });
    fo.cloudCurvature.assign(cap, 0.f); fo.cloudNeighborPicked.assign(cap, 0); fo.cloudLabel.assign(cap, 0);
    float* cloudCurvature = fo.cloudCurvature.data();
    int* cloudNeighborPicked = fo.cloudNeighborPicked.data();
    int* cloudLabel = fo.cloudLabel.data();
    const float* pointRange = ci.pointRange.data();
    auto colInd = [&](int k) -> int { return (k >= 0 && k < cloudSize) ? ci.pointColInd[k] : 0; };
    // calculateSmoothness  featureExtraction.h:109-131
    for (int i = 5; i < cloudSize - 5; i++) {
        float diffRange = pointRange[i - 5] + pointRange[i - 4] + pointRange[i - 3] + pointRange[i - 2] + pointRange[i - 1]
                        - pointRange[i] * 10
                        + pointRange[i + 1] + pointRange[i + 2] + pointRange[i + 3] + pointRange[i + 4] + pointRange[i + 5];
        cloudCurvature[i] = diffRange * diffRange;
        cloudNeighborPicked[i] = 0;
        cloudLabel[i] = 0;
        cloudSmoothness[i].value = cloudCurvature[i];
        cloudSmoothness[i].ind = i;
    }
    // markOccludedPoints  featureExtraction.h:134-176
    for (int i = 5; i < cloudSize - 6; ++i) {
        float depth1 = pointRange[i], depth2 = pointRange[i + 1];
        int columnDiff = std::abs(int(ci.pointColInd[i + 1] - ci.pointColInd[i]));
        if (columnDiff < 10) {
            if ((double)(depth1 - depth2) > 0.3) {
                for (int q = 5; q >= 0; q--) cloudNeighborPicked[i - q] = 1;
            } else if ((double)(depth2 - depth1) > 0.3) {
                for (int q = 1; q <= 6; q++) cloudNeighborPicked[i + q] = 1;
            }
        }
        float diff1 = std::abs(float(pointRange[i - 1] - pointRange[i]));
        float diff2 = std::abs(float(pointRange[i + 1] - pointRange[i]));
        if ((double)diff1 > 0.02 * (double)pointRange[i] && (double)diff2 > 0.02 * (double)pointRange[i])
            cloudNeighborPicked[i] = 1;
    }
    // extractFeatures  featureExtraction.h:178-294
    fo.cornerCloud.clear(); fo.surfaceCloud.clear(); fo.cornerIndex.clear(); fo.surfaceRawIndex.clear();
    fo.surfaceRingCount.assign(P.N_SCAN, 0); fo.surfaceRingCountDS.assign(P.N_SCAN, 0);
    std::vector<P4> surfaceCloudScan, surfaceCloudScanDS;
    auto suppress = [&](int ind) {
        for (int l = 1; l <= 5; l++) {
            if (ind + l >= cap) break;
            int columnDiff = std::abs(int(colInd(ind + l) - colInd(ind + l - 1)));
            if (columnDiff > 10) break;
            cloudNeighborPicked[ind + l] = 1;
        }
        for (int l = -1; l >= -5; l--) {
            if (ind + l < 0) break;                           // UB guard (oracle contract)
            int columnDiff = std::abs(int(colInd(ind + l) - colInd(ind + l + 1)));
            if (columnDiff > 10) break;
            cloudNeighborPicked[ind + l] = 1;
        }
    };
    for (int i = 0; i < P.N_SCAN; i++) {
        surfaceCloudScan.clear();
        for (int j = 0; j < 6; j++) {
            int sp = (ci.startRingIndex[i] * (6 - j) + ci.endRingIndex[i] * j) / 6;
            int ep = (ci.startRingIndex[i] * (5 - j) + ci.endRingIndex[i] * (j + 1)) / 6 - 1;
            if (sp >= ep) continue;
            if (oracle_literal_sort())                        // featureExtraction.h:203 with by_value (:13-17): value only
                std::sort(cloudSmoothness.begin() + sp, cloudSmoothness.begin() + ep, [](const Smooth& l, const Smooth& r) { return l.value < r.value; });
            else
            std::sort(cloudSmoothness.begin() + sp, cloudSmoothness.begin() + ep,
                      [](const Smooth& l, const Smooth& r) { return l.value < r.value || (l.value == r.value && l.ind < r.ind); });
            int largestPickedNum = 0;
            for (int k = ep; k >= sp; k--) {
                int ind = cloudSmoothness[k].ind;
                if (cloudNeighborPicked[ind] == 0 && cloudCurvature[ind] > P.edgeThreshold) {
                    largestPickedNum++;
                    if (largestPickedNum <= 20) {
                        cloudLabel[ind] = 1;
                        fo.cornerCloud.push_back(ci.cloud_deskewed[ind]);
                        fo.cornerIndex.push_back(ind);
                    } else {
                        break;
                    }
                    cloudNeighborPicked[ind] = 1;
                    suppress(ind);
                }
            }
            for (int k = sp; k <= ep; k++) {
                int ind = cloudSmoothness[k].ind;
                if (cloudNeighborPicked[ind] == 0 && cloudCurvature[ind] < P.surfThreshold) {
                    cloudLabel[ind] = -1;
                    cloudNeighborPicked[ind] = 1;
                    suppress(ind);
                }
            }
            for (int k = sp; k <= ep; k++) {
                if (cloudLabel[k] <= 0) { surfaceCloudScan.push_back(ci.cloud_deskewed[k]); fo.surfaceRawIndex.push_back(k); }
            }
        }
        fo.surfaceRingCount[i] = (int)surfaceCloudScan.size();
        voxel_grid(surfaceCloudScan.data(), (int)surfaceCloudScan.size(), P.odometrySurfLeafSize, surfaceCloudScanDS);
        fo.surfaceRingCountDS[i] = (int)surfaceCloudScanDS.size();
        fo.surfaceCloud.insert(fo.surfaceCloud.end(), surfaceCloudScanDS.begin(), surfaceCloudScanDS.end());
    }
}

// ===================================================================== scan-to-map
enum : unsigned {                       // semantic outcomes, mirrored by the C ABI (include/fbpr_b200.h)
    FLAG_NOT_ENOUGH_FEATURES = 1u,      // mapOptmization.h:1410 gate failed, pose unchanged, transformUpdate skipped
    FLAG_TOO_FEW_CORRESPONDENCES = 2u,  // some iteration had < 50 rows (mapOptmization.h:1267-1270)
    FLAG_DEGENERATE = 4u,               // isDegenerate set at iteration 0 (mapOptmization.h:1346-1371)
    FLAG_CONVERGED = 8u,                // LMOptimization returned true before iteration 30
};

struct IterDebug {                      // captured for one chosen iteration (parity tests)
    int iter = -1;
    std::vector<int> cornerKnn, surfKnn;          // 5 per point
    std::vector<float> cornerD2, surfD2;          // 5 per point
    std::vector<P4> cornerCoeff, surfCoeff;       // per DS point (valid only where flag)
    std::vector<uint8_t> cornerFlag, surfFlag;
    float AtA[36], AtB[6], X[6];
    int nSel = 0;
};

class MapOptimization {
public:
    Params P;
    std::vector<P4> laserCloudCornerLast, laserCloudSurfLast, laserCloudCornerLastDS, laserCloudSurfLastDS;
    std::vector<P4> laserCloudCornerFromMap, laserCloudSurfFromMap, laserCloudCornerFromMapDS, laserCloudSurfFromMapDS;
    float transformTobeMapped[6] = { 0, 0, 0, 0, 0, 0 };
    bool isDegenerate = false;
    int64_t imuAvailable = 0; float imuRollInit = 0, imuPitchInit = 0;
    int itersDone = 0; unsigned flags = 0;
    double buildSeconds = 0, loopSeconds = 0;
    IterDebug* debug = nullptr; int debugIter = -1;
    std::vector<float> poseTrace;                  // pose after every executed iteration (6 each)

    // extractCloud  mapOptmization.h:909-955 : transform K keyframes, concat, VoxelGrid x2
    void extractCloud(const float* keyPoses6 /*K x (roll,pitch,yaw,x,y,z)*/, int K,
                      const P4* const* cornerFrames, const int* cornerN,
                      const P4* const* surfFrames, const int* surfN,
                      const float* lastKeyXYZ) {
        std::vector<std::vector<P4>> cv(K), sv(K);
        #pragma omp parallel for num_threads(P.numberOfCores)
        for (int i = 0; i < K; ++i) {
            const float* kp = keyPoses6 + 6 * i;
            float ddx = kp[3] - lastKeyXYZ[0], ddy = kp[4] - lastKeyXYZ[1], ddz = kp[5] - lastKeyXYZ[2];
            if (std::sqrt(ddx * ddx + ddy * ddy + ddz * ddz) > P.surroundingKeyframeSearchRadius) continue;
            float T[12]; get_transformation(kp[3], kp[4], kp[5], kp[0], kp[1], kp[2], T);
            transformPointCloud(cornerFrames[i], cornerN[i], T, cv[i]);
            transformPointCloud(surfFrames[i], surfN[i], T, sv[i]);
        }
        laserCloudCornerFromMap.clear(); laserCloudSurfFromMap.clear();
        for (int i = 0; i < K; ++i) {
            laserCloudCornerFromMap.insert(laserCloudCornerFromMap.end(), cv[i].begin(), cv[i].end());
            laserCloudSurfFromMap.insert(laserCloudSurfFromMap.end(), sv[i].begin(), sv[i].end());
        }
        voxel_grid(laserCloudCornerFromMap.data(), (int)laserCloudCornerFromMap.size(), P.mappingCornerLeafSize, laserCloudCornerFromMapDS);
        voxel_grid(laserCloudSurfFromMap.data(), (int)laserCloudSurfFromMap.size(), P.mappingSurfLeafSize, laserCloudSurfFromMapDS);
    }
    // extractNearby  mapOptmization.h:872-907.  cloudKeyPoses3D[i] = (x, y, z, intensity = i); radiusSearch is FLANN's
    // RadiusResultSet: d^2 < (float)(r*r) (strict -- confirmed against cv2.flann's radiusSearch, tests/golden/flann_cv2.npz), sorted by (d^2, index);
    // VoxelGrid(surroundingKeyframeDensity) averages xyz AND the intensity; then the key poses of the last 10 s, newest first.
    static void extractNearby(const P4* cloudKeyPoses3D, const double* keyTime, int n, float searchRadius, float density,
                              double timeLaserCloudInfoLast, std::vector<P4>& surroundingKeyPosesDS) {
        surroundingKeyPosesDS.clear();
        if (n <= 0) return;
        const P4& q = cloudKeyPoses3D[n - 1];
        const float r2 = (float)((double)searchRadius * (double)searchRadius);
        std::vector<std::pair<float, int>> hit;
        for (int i = 0; i < n; ++i) {
            const float dx = q.x - cloudKeyPoses3D[i].x, dy = q.y - cloudKeyPoses3D[i].y, dz = q.z - cloudKeyPoses3D[i].z;
            float d = dx * dx; d += dy * dy; d += dz * dz;                       // L2_Simple, x, y, z order
            if (d < r2) hit.push_back({ d, i });
        }
        std::sort(hit.begin(), hit.end());
        std::vector<P4> surroundingKeyPoses;
        for (auto& h : hit) surroundingKeyPoses.push_back(cloudKeyPoses3D[h.second]);
        voxel_grid(surroundingKeyPoses.data(), (int)surroundingKeyPoses.size(), density, surroundingKeyPosesDS);
        for (int i = n - 1; i >= 0; --i) {
            if (timeLaserCloudInfoLast - keyTime[i] < 10.0) surroundingKeyPosesDS.push_back(cloudKeyPoses3D[i]);
            else break;
        }
    }
    // extractForLoopClosure  mapOptmization.h:857-870: key poses from the newest backwards while the list holds
    // <= surroundingKeyframeSize entries, i.e. surroundingKeyframeSize + 1 of them when the store is large enough
    static void extractForLoopClosure(const P4* cloudKeyPoses3D, int n, int surroundingKeyframeSize, std::vector<P4>& cloudToExtract) {
        cloudToExtract.clear();
        for (int i = n - 1; i >= 0; --i) {
            if ((int)cloudToExtract.size() <= surroundingKeyframeSize) cloudToExtract.push_back(cloudKeyPoses3D[i]);
            else break;
        }
    }
    // extractCloud as the reference calls it (:909-955): entry i of cloudToExtract is re-checked at ITS OWN position (:924) and
    // names keyframe (int)intensity (:927), whose pose and clouds are used
    void extractCloudIndexed(const std::vector<P4>& cloudToExtract, const float* keyPoses6All, int nKeys,
                             const P4* const* cornerFramesAll, const int* cornerNAll,
                             const P4* const* surfFramesAll, const int* surfNAll) {
        const int K = (int)cloudToExtract.size();
        std::vector<std::vector<P4>> cv(K), sv(K);
        const float* lk = keyPoses6All + 6 * (nKeys - 1) + 3;                    // cloudKeyPoses3D->back()
        #pragma omp parallel for num_threads(P.numberOfCores)
        for (int i = 0; i < K; ++i) {
            const P4& c = cloudToExtract[i];
            if (std::sqrt((c.x - lk[0]) * (c.x - lk[0]) + (c.y - lk[1]) * (c.y - lk[1]) + (c.z - lk[2]) * (c.z - lk[2])) > P.surroundingKeyframeSearchRadius) continue;
            const int thisKeyInd = (int)c.i;
            const float* kp = keyPoses6All + 6 * thisKeyInd;
            float T[12]; get_transformation(kp[3], kp[4], kp[5], kp[0], kp[1], kp[2], T);
            transformPointCloud(cornerFramesAll[thisKeyInd], cornerNAll[thisKeyInd], T, cv[i]);
            transformPointCloud(surfFramesAll[thisKeyInd], surfNAll[thisKeyInd], T, sv[i]);
        }
        laserCloudCornerFromMap.clear(); laserCloudSurfFromMap.clear();
        for (int i = 0; i < K; ++i) {
            laserCloudCornerFromMap.insert(laserCloudCornerFromMap.end(), cv[i].begin(), cv[i].end());
            laserCloudSurfFromMap.insert(laserCloudSurfFromMap.end(), sv[i].begin(), sv[i].end());
        }
        voxel_grid(laserCloudCornerFromMap.data(), (int)laserCloudCornerFromMap.size(), P.mappingCornerLeafSize, laserCloudCornerFromMapDS);
        voxel_grid(laserCloudSurfFromMap.data(), (int)laserCloudSurfFromMap.size(), P.mappingSurfLeafSize, laserCloudSurfFromMapDS);
    }
    static void transformPointCloud(const P4* in, int n, const float T[12], std::vector<P4>& out) {   // :405-425
        out.resize(n);
        for (int i = 0; i < n; ++i) {
            const P4& p = in[i];
            out[i].x = T[0] * p.x + T[1] * p.y + T[2] * p.z + T[3];
            out[i].y = T[4] * p.x + T[5] * p.y + T[6] * p.z + T[7];
            out[i].z = T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11];
            out[i].i = p.i;
        }
    }

    void downsampleCurrentScan() {                                                   // :981-993
        voxel_grid(laserCloudCornerLast.data(), (int)laserCloudCornerLast.size(), P.mappingCornerLeafSize, laserCloudCornerLastDS);
        voxel_grid(laserCloudSurfLast.data(), (int)laserCloudSurfLast.size(), P.mappingSurfLeafSize, laserCloudSurfLastDS);
    }

    void scan2MapOptimization() {                                                    // :1403-1442
        const int nC = (int)laserCloudCornerLastDS.size(), nS = (int)laserCloudSurfLastDS.size();
        flags = 0; itersDone = 0; isDegenerate = false; poseTrace.clear();
        if (nC > P.edgeFeatureMinValidNum && nS > P.surfFeatureMinValidNum) {
            double t0 = omp_get_wtime();
            kdCorner.build(laserCloudCornerFromMapDS.data(), (int)laserCloudCornerFromMapDS.size());
            kdSurf.build(laserCloudSurfFromMapDS.data(), (int)laserCloudSurfFromMapDS.size());
            double t1 = omp_get_wtime();
            oriC.assign(nC, P4{}); coeffC.assign(nC, P4{}); flagC.assign(nC, 0);
            oriS.assign(nS, P4{}); coeffS.assign(nS, P4{}); flagS.assign(nS, 0);
            for (int iterCount = 0; iterCount < 30; iterCount++) {
                laserCloudOri.clear(); coeffSel.clear();
                bool cap = debug && iterCount == debugIter;
                if (cap) { debug->iter = iterCount; debug->cornerKnn.assign(5 * (size_t)nC, -1); debug->surfKnn.assign(5 * (size_t)nS, -1);
                           debug->cornerD2.assign(5 * (size_t)nC, 0.f); debug->surfD2.assign(5 * (size_t)nS, 0.f); }
                cornerOptimization(cap);
                surfOptimization(cap);
                if (cap) { debug->cornerCoeff = coeffC; debug->surfCoeff = coeffS; debug->cornerFlag = flagC; debug->surfFlag = flagS; }
                combineOptimizationCoeffs();
                itersDone = iterCount + 1;
                bool conv = LMOptimization(iterCount, cap);
                poseTrace.insert(poseTrace.end(), transformTobeMapped, transformTobeMapped + 6);
                if (conv) { flags |= FLAG_CONVERGED; break; }
            }
            if (isDegenerate) flags |= FLAG_DEGENERATE;
            double t2 = omp_get_wtime();
            buildSeconds = t1 - t0; loopSeconds = t2 - t1;
            transformUpdate();
        } else {
            flags |= FLAG_NOT_ENOUGH_FEATURES;
        }
    }

    void transformUpdate() {                                                          // :1444-1479
        if (imuAvailable == 1) {
            if (std::abs(imuPitchInit) < 1.4) {
                double imuWeight = 0.05;
                double q0[4], q1[4], qm[4], r, p, y;
                set_rpy((double)transformTobeMapped[0], 0, 0, q0); set_rpy((double)imuRollInit, 0, 0, q1);
                slerp(q0, q1, imuWeight, qm); get_rpy(qm, r, p, y);
                transformTobeMapped[0] = (float)r;
                set_rpy(0, (double)transformTobeMapped[1], 0, q0); set_rpy(0, (double)imuPitchInit, 0, q1);
                slerp(q0, q1, imuWeight, qm); get_rpy(qm, r, p, y);
                transformTobeMapped[1] = (float)p;
            }
        }
        transformTobeMapped[0] = constraintTransformation(transformTobeMapped[0], P.rotation_tollerance);
        transformTobeMapped[1] = constraintTransformation(transformTobeMapped[1], P.rotation_tollerance);
        transformTobeMapped[5] = constraintTransformation(transformTobeMapped[5], P.z_tollerance);
    }

    // registration()  mapOptmization.h:263-343 : CropBox local map, pose decompose, downsample, LM, recompose.
    // pose is a 3x4 row-major rigid transform (Eigen::Affine3f), in/out.
    void registration(const P4* cornerGlobal, int nCg, const P4* surfGlobal, int nSg, float pose[12]) {
        float mn[3] = { -30.0f + pose[3], -30.0f + pose[7], -10.0f + pose[11] };
        float mx[3] = { 30.0f + pose[3], 30.0f + pose[7], 10.0f + pose[11] };
        crop_box(cornerGlobal, nCg, mn, mx, laserCloudCornerFromMapDS);
        crop_box(surfGlobal, nSg, mn, mx, laserCloudSurfFromMapDS);
        get_translation_and_euler(pose, transformTobeMapped[3], transformTobeMapped[4], transformTobeMapped[5],
                                  transformTobeMapped[0], transformTobeMapped[1], transformTobeMapped[2]);
        downsampleCurrentScan();
        scan2MapOptimization();
        get_transformation(transformTobeMapped[3], transformTobeMapped[4], transformTobeMapped[5],
                           transformTobeMapped[0], transformTobeMapped[1], transformTobeMapped[2], pose);
    }

private:
    KdTree5 kdCorner, kdSurf;
    std::vector<P4> oriC, coeffC, oriS, coeffS, laserCloudOri, coeffSel;
    std::vector<uint8_t> flagC, flagS;
    float T_[12];

    static float constraintTransformation(float value, float limit) {
        if (value < -limit) value = -limit;
        if (value > limit) value = limit;
        return value;
    }
    // tf::Quaternion::setRPY / slerp, tf::Matrix3x3::getRPY (all f64)   mapOptmization.h:1459-1472
    static void set_rpy(double roll, double pitch, double yaw, double q[4]) {
        double hy = yaw * 0.5, hp = pitch * 0.5, hr = roll * 0.5;
        double cy = std::cos(hy), sy = std::sin(hy), cp = std::cos(hp), sp = std::sin(hp), cr = std::cos(hr), sr = std::sin(hr);
        q[0] = sr * cp * cy - cr * sp * sy; q[1] = cr * sp * cy + sr * cp * sy;
        q[2] = cr * cp * sy - sr * sp * cy; q[3] = cr * cp * cy + sr * sp * sy;
    }
    static double qdot(const double a[4], const double b[4]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3]; }
    static void slerp(const double a[4], const double b[4], double t, double o[4]) {
        double s = std::sqrt(qdot(a, a) * qdot(b, b));
        double d = qdot(a, b);
        double theta = (d < 0 ? std::acos(-d / s) * 2.0 : std::acos(d / s) * 2.0) / 2.0;
        if (theta != 0.0) {
            double dd = 1.0 / std::sin(theta), s0 = std::sin((1.0 - t) * theta), s1 = std::sin(t * theta);
            double sg = d < 0 ? -1.0 : 1.0;
            for (int k = 0; k < 4; k++) o[k] = (a[k] * s0 + sg * b[k] * s1) * dd;
        } else {
            for (int k = 0; k < 4; k++) o[k] = a[k];
        }
    }
    static void get_rpy(const double q[4], double& roll, double& pitch, double& yaw) {
        double d = qdot(q, q), s = 2.0 / d;
        double xs = q[0] * s, ys = q[1] * s, zs = q[2] * s;
        double wx = q[3] * xs, wy = q[3] * ys, wz = q[3] * zs;
        double xx = q[0] * xs, xy = q[0] * ys, xz = q[0] * zs, yy = q[1] * ys, yz = q[1] * zs, zz = q[2] * zs;
        double m00 = 1.0 - (yy + zz), m10 = xy + wz, m20 = xz - wy, m21 = yz + wx, m22 = 1.0 - (xx + yy);
        double m01 = xy - wz, m02 = xz + wy;
        if (std::fabs(m20) >= 1) {
            yaw = 0;
            double delta = std::atan2(m01, m02);
            if (m20 < 0) { pitch = M_PI / 2.0; roll = delta; }
            else { pitch = -M_PI / 2.0; roll = delta; }
        } else {
            pitch = -std::asin(m20);
            roll = std::atan2(m21 / std::cos(pitch), m22 / std::cos(pitch));
            yaw = std::atan2(m10 / std::cos(pitch), m00 / std::cos(pitch));
        }
    }

    void updatePointAssociateToMap() {                                               // :995-1000, :444-448
        get_transformation(transformTobeMapped[3], transformTobeMapped[4], transformTobeMapped[5],
                           transformTobeMapped[0], transformTobeMapped[1], transformTobeMapped[2], T_);
    }
    inline void pointAssociateToMap(const P4& pi, P4& po) const {                    // :397-403
        po.x = T_[0] * pi.x + T_[1] * pi.y + T_[2] * pi.z + T_[3];
        po.y = T_[4] * pi.x + T_[5] * pi.y + T_[6] * pi.z + T_[7];
        po.z = T_[8] * pi.x + T_[9] * pi.y + T_[10] * pi.z + T_[11];
        po.i = pi.i;
    }

    void cornerOptimization(bool cap) {                                              // :1002-1124
        updatePointAssociateToMap();
        const int n = (int)laserCloudCornerLastDS.size();
        const P4* map = laserCloudCornerFromMapDS.data();
        #pragma omp parallel for num_threads(P.numberOfCores)
        for (int i = 0; i < n; i++) {
            P4 pointOri = laserCloudCornerLastDS[i], pointSel, coeff;
            pointAssociateToMap(pointOri, pointSel);
            int ind[5]; float sq[5];
            const float q[3] = { pointSel.x, pointSel.y, pointSel.z };
            kdCorner.knn5(q, ind, sq);
            if (cap) for (int j = 0; j < 5; j++) { debug->cornerKnn[5 * (size_t)i + j] = ind[j]; debug->cornerD2[5 * (size_t)i + j] = sq[j]; }
            if (kdCorner.size() >= 5 && (double)sq[4] < 1.0) {
                float cx = 0, cy = 0, cz = 0;
                for (int j = 0; j < 5; j++) { cx += map[ind[j]].x; cy += map[ind[j]].y; cz += map[ind[j]].z; }
                cx /= 5; cy /= 5; cz /= 5;
                float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
                for (int j = 0; j < 5; j++) {
                    float ax = map[ind[j]].x - cx, ay = map[ind[j]].y - cy, az = map[ind[j]].z - cz;
                    a11 += ax * ax; a12 += ax * ay; a13 += ax * az;
                    a22 += ay * ay; a23 += ay * az;
                    a33 += az * az;
                }
                a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;
                float A1[9] = { a11, a12, a13, a12, a22, a23, a13, a23, a33 }, D1[3], V1[9];
                jacobi_eigen_sym(3, A1, D1, V1);
                if (D1[0] > 3 * D1[1]) {
                    float x0 = pointSel.x, y0 = pointSel.y, z0 = pointSel.z;
                    float x1 = (float)((double)cx + 0.1 * (double)V1[0]);
                    float y1 = (float)((double)cy + 0.1 * (double)V1[1]);
                    float z1 = (float)((double)cz + 0.1 * (double)V1[2]);
                    float x2 = (float)((double)cx - 0.1 * (double)V1[0]);
                    float y2 = (float)((double)cy - 0.1 * (double)V1[1]);
                    float z2 = (float)((double)cz - 0.1 * (double)V1[2]);
                    float a012 = std::sqrt(((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1))
                                         + ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1)) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1))
                                         + ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1)) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1)));
                    float l12 = std::sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
                    float la = ((y1 - y2) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1))
                              + (z1 - z2) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1))) / a012 / l12;
                    float lb = -((x1 - x2) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1))
                               - (z1 - z2) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1))) / a012 / l12;
                    float lc = -((x1 - x2) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1))
                               + (y1 - y2) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1))) / a012 / l12;
                    float ld2 = a012 / l12;
                    float s = (float)(1.0 - 0.9 * (double)std::fabs(ld2));
                    coeff.x = s * la; coeff.y = s * lb; coeff.z = s * lc; coeff.i = s * ld2;
                    if ((double)s > 0.1) { oriC[i] = pointOri; coeffC[i] = coeff; flagC[i] = 1; }
                }
            }
        }
    }

    void surfOptimization(bool cap) {                                                // :1126-1215
        updatePointAssociateToMap();
        const int n = (int)laserCloudSurfLastDS.size();
        const P4* map = laserCloudSurfFromMapDS.data();
        #pragma omp parallel for num_threads(P.numberOfCores)
        for (int i = 0; i < n; i++) {
            P4 pointOri = laserCloudSurfLastDS[i], pointSel, coeff;
            pointAssociateToMap(pointOri, pointSel);
            int ind[5]; float sq[5];
            const float q[3] = { pointSel.x, pointSel.y, pointSel.z };
            kdSurf.knn5(q, ind, sq);
            if (cap) for (int j = 0; j < 5; j++) { debug->surfKnn[5 * (size_t)i + j] = ind[j]; debug->surfD2[5 * (size_t)i + j] = sq[j]; }
            if (kdSurf.size() >= 5 && (double)sq[4] < 1.0) {
                float A0[15], B0[5] = { -1, -1, -1, -1, -1 }, X0[3];
                for (int j = 0; j < 5; j++) { A0[3 * j] = map[ind[j]].x; A0[3 * j + 1] = map[ind[j]].y; A0[3 * j + 2] = map[ind[j]].z; }
                colpiv_householder_solve_5x3(A0, B0, X0);
                float pa = X0[0], pb = X0[1], pc = X0[2], pd = 1;
                float ps = std::sqrt(pa * pa + pb * pb + pc * pc);
                pa /= ps; pb /= ps; pc /= ps; pd /= ps;
                bool planeValid = true;
                for (int j = 0; j < 5; j++) {
                    if ((double)std::fabs(pa * map[ind[j]].x + pb * map[ind[j]].y + pc * map[ind[j]].z + pd) > 0.2) { planeValid = false; break; }
                }
                if (planeValid) {
                    float pd2 = pa * pointSel.x + pb * pointSel.y + pc * pointSel.z + pd;
                    float s = (float)(1.0 - 0.9 * (double)std::fabs(pd2)
                                            / (double)std::sqrt(std::sqrt(pointSel.x * pointSel.x + pointSel.y * pointSel.y + pointSel.z * pointSel.z)));
                    coeff.x = s * pa; coeff.y = s * pb; coeff.z = s * pc; coeff.i = s * pd2;
                    if ((double)s > 0.1) { oriS[i] = pointOri; coeffS[i] = coeff; flagS[i] = 1; }
                }
            }
        }
    }

    void combineOptimizationCoeffs() {                                               // :1218-1243
        for (size_t i = 0; i < flagC.size(); ++i) if (flagC[i]) { laserCloudOri.push_back(oriC[i]); coeffSel.push_back(coeffC[i]); }
        for (size_t i = 0; i < flagS.size(); ++i) if (flagS[i]) { laserCloudOri.push_back(oriS[i]); coeffSel.push_back(coeffS[i]); }
        std::fill(flagC.begin(), flagC.end(), 0); std::fill(flagS.begin(), flagS.end(), 0);
    }

    bool LMOptimization(int iterCount, bool cap) {                                   // :1246-1401
        float srx = sinf_c(transformTobeMapped[1]), crx = cosf_c(transformTobeMapped[1]);
        float sry = sinf_c(transformTobeMapped[2]), cry = cosf_c(transformTobeMapped[2]);
        float srz = sinf_c(transformTobeMapped[0]), crz = cosf_c(transformTobeMapped[0]);
        int laserCloudSelNum = (int)laserCloudOri.size();
        if (cap) debug->nSel = laserCloudSelNum;
        if (laserCloudSelNum < 50) { flags |= FLAG_TOO_FEW_CORRESPONDENCES; return false; }
        double AtA64[36], AtB64[6];
        for (int k = 0; k < 36; k++) AtA64[k] = 0.0;
        for (int k = 0; k < 6; k++) AtB64[k] = 0.0;
        for (int i = 0; i < laserCloudSelNum; i++) {
            P4 pointOri, coeff;
            pointOri.x = laserCloudOri[i].y; pointOri.y = laserCloudOri[i].z; pointOri.z = laserCloudOri[i].x;
            coeff.x = coeffSel[i].y; coeff.y = coeffSel[i].z; coeff.z = coeffSel[i].x; coeff.i = coeffSel[i].i;
            float arx = (crx * sry * srz * pointOri.x + crx * crz * sry * pointOri.y - srx * sry * pointOri.z) * coeff.x
                      + (-srx * srz * pointOri.x - crz * srx * pointOri.y - crx * pointOri.z) * coeff.y
                      + (crx * cry * srz * pointOri.x + crx * cry * crz * pointOri.y - cry * srx * pointOri.z) * coeff.z;
            float ary = ((cry * srx * srz - crz * sry) * pointOri.x
                      + (sry * srz + cry * crz * srx) * pointOri.y + crx * cry * pointOri.z) * coeff.x
                      + ((-cry * crz - srx * sry * srz) * pointOri.x
                      + (cry * srz - crz * srx * sry) * pointOri.y - crx * sry * pointOri.z) * coeff.z;
            float arz = ((crz * srx * sry - cry * srz) * pointOri.x + (-cry * crz - srx * sry * srz) * pointOri.y) * coeff.x
                      + (crx * crz * pointOri.x - crx * srz * pointOri.y) * coeff.y
                      + ((sry * srz + cry * crz * srx) * pointOri.x + (crz * sry - cry * srx * srz) * pointOri.y) * coeff.z;
            float row[6] = { arz, arx, ary, coeff.z, coeff.x, coeff.y };
            float b = -coeff.i;
            for (int r = 0; r < 6; r++) {
                for (int c = 0; c < 6; c++) AtA64[r * 6 + c] += (double)row[r] * (double)row[c];
                AtB64[r] += (double)row[r] * (double)b;
            }
        }
        float matAtA[36], matAtB[6], matX[6];
        for (int k = 0; k < 36; k++) matAtA[k] = (float)AtA64[k];
        for (int k = 0; k < 6; k++) matAtB[k] = (float)AtB64[k];
        if (cap) { for (int k = 0; k < 36; k++) debug->AtA[k] = matAtA[k]; for (int k = 0; k < 6; k++) debug->AtB[k] = matAtB[k]; }
        { float Aw[36], bw[6]; for (int k = 0; k < 36; k++) Aw[k] = matAtA[k]; for (int k = 0; k < 6; k++) bw[k] = matAtB[k];
          qr_solve(6, Aw, bw, matX); }
        float matP[36];                                  // LOCAL, zero-initialised: shadows the member (:1278)
        for (int k = 0; k < 36; k++) matP[k] = 0.f;
        if (iterCount == 0) {
            float Aw[36], matE[6], matV[36], matV2[36];
            for (int k = 0; k < 36; k++) Aw[k] = matAtA[k];
            jacobi_eigen_sym(6, Aw, matE, matV);
            for (int k = 0; k < 36; k++) matV2[k] = matV[k];
            isDegenerate = false;
            float eignThre[6] = { 100, 100, 100, 100, 100, 100 };
            for (int i = 5; i >= 0; i--) {
                if (matE[i] < eignThre[i]) { for (int j = 0; j < 6; j++) matV2[i * 6 + j] = 0; isDegenerate = true; }
                else break;
            }
            float Vw[36], Vinv[36];
            for (int k = 0; k < 36; k++) Vw[k] = matV[k];
            lu_invert(6, Vw, Vinv);
            matmul_f64acc(6, 6, 6, Vinv, matV2, matP);
        }
        if (isDegenerate) {
            float matX2[6]; for (int k = 0; k < 6; k++) matX2[k] = matX[k];
            matmul_f64acc(6, 6, 1, matP, matX2, matX);
        }
        if (cap) for (int k = 0; k < 6; k++) debug->X[k] = matX[k];
        for (int k = 0; k < 6; k++) transformTobeMapped[k] += matX[k];
        float deltaR = (float)std::sqrt(std::pow((double)(matX[0] * 57.29578f), 2) + std::pow((double)(matX[1] * 57.29578f), 2)
                                      + std::pow((double)(matX[2] * 57.29578f), 2));
        float deltaT = (float)std::sqrt(std::pow((double)(matX[3] * 100), 2) + std::pow((double)(matX[4] * 100), 2)
                                      + std::pow((double)(matX[5] * 100), 2));
        if ((double)deltaR < 0.05 && (double)deltaT < 0.05) return true;
        return false;
    }
};

}  // namespace orc
