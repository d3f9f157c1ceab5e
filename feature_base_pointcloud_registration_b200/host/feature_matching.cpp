// feature_matching.cpp -- the reference's operator surface over the C ABI (see feature_matching.hpp).
// Host code only: argument marshalling, the params.yaml reader and the keyframe bookkeeping the
// reference keeps on the host.  Every numeric step is a call into libfbpr_b200.so.
#include "feature_matching.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace feature_matching_b200 {

static void check(int rc, const char* what) {
    if (rc < 0) throw std::runtime_error(std::string(what) + ": " + fbpr_last_error());
}

// ------------------------------------------------------------------ DeviceContext
DeviceContext::DeviceContext(const fbpr_params& p, int device) { check(fbpr_create(&p, device, &h), "fbpr_create"); }
DeviceContext::~DeviceContext() { if (h) fbpr_destroy(h); }

// ------------------------------------------------------------------ ParamServer
bool ParamServer::loadYaml(const std::string& path) {
    std::ifstream f(path);
    if (!f) return false;
    std::string line;
    while (std::getline(f, line)) {
        size_t hash = line.find('#'); if (hash != std::string::npos) line = line.substr(0, hash);
        size_t colon = line.find(':'); if (colon == std::string::npos) continue;
        std::string key = line.substr(0, colon), val = line.substr(colon + 1);
        auto trim = [](std::string& s) { size_t a = s.find_first_not_of(" \t\r\""), b = s.find_last_not_of(" \t\r\""); s = a == std::string::npos ? "" : s.substr(a, b - a + 1); };
        trim(key); trim(val);
        if (val.empty()) continue;
        std::istringstream is(val);
        if (key == "N_SCAN") is >> N_SCAN; else if (key == "Horizon_SCAN") is >> Horizon_SCAN;
        else if (key == "edgeThreshold") is >> edgeThreshold; else if (key == "surfThreshold") is >> surfThreshold;
        else if (key == "edgeFeatureMinValidNum") is >> edgeFeatureMinValidNum; else if (key == "surfFeatureMinValidNum") is >> surfFeatureMinValidNum;
        else if (key == "odometrySurfLeafSize") is >> odometrySurfLeafSize; else if (key == "mappingCornerLeafSize") is >> mappingCornerLeafSize;
        else if (key == "mappingSurfLeafSize") is >> mappingSurfLeafSize;
        else if (key == "z_tollerance") is >> z_tollerance; else if (key == "rotation_tollerance") is >> rotation_tollerance;
        else if (key == "numberOfCores") is >> numberOfCores; else if (key == "mappingProcessInterval") is >> mappingProcessInterval;
        else if (key == "surroundingKeyframeSearchRadius") is >> surroundingKeyframeSearchRadius;
        else if (key == "loopClosureEnableFlag") loopClosureEnableFlag = (val == "true" || val == "True" || val == "1");
    }
    return true;
}

fbpr_params ParamServer::toAbi(int max_frames, int max_map_corner, int max_map_surf, int max_keyframe_points) const {
    fbpr_params p; std::memset(&p, 0, sizeof(p));
    p.N_SCAN = N_SCAN; p.Horizon_SCAN = Horizon_SCAN; p.edgeThreshold = edgeThreshold; p.surfThreshold = surfThreshold;
    p.edgeFeatureMinValidNum = edgeFeatureMinValidNum; p.surfFeatureMinValidNum = surfFeatureMinValidNum;
    p.odometrySurfLeafSize = odometrySurfLeafSize; p.mappingCornerLeafSize = mappingCornerLeafSize; p.mappingSurfLeafSize = mappingSurfLeafSize;
    p.z_tollerance = z_tollerance; p.rotation_tollerance = rotation_tollerance; p.numberOfCores = numberOfCores;
    p.surroundingKeyframeSearchRadius = surroundingKeyframeSearchRadius;
    p.max_frames = max_frames; p.max_map_corner = max_map_corner; p.max_map_surf = max_map_surf; p.max_keyframe_points = max_keyframe_points;
    return p;
}

static void download(fbpr_handle* h, int which, int n, PointCloud& out) {
    out.resize((size_t)(n > 0 ? n : 0));
    if (n > 0) check((int)fbpr_get_buffer(h, 0, which, out.data(), (int64_t)out.size() * sizeof(PointType)), "fbpr_get_buffer");
}

// ------------------------------------------------------------------ FeatureExtraction
FeatureExtraction::FeatureExtraction(const ParamServer& params, std::shared_ptr<DeviceContext> ctx) : ParamServer(params), ctx_(ctx) {}

void FeatureExtraction::featureExtra(const cloud_info& in) {
    cloudInfo = in;                                              // featureExtraction.h:90
    extractedCloud = in.cloud_deskewed;                          // :92
    fbpr_cloud_info_view v;
    v.startRingIndex = in.startRingIndex.data(); v.endRingIndex = in.endRingIndex.data();
    v.pointColInd = in.pointColInd.data(); v.pointRange = in.pointRange.data();
    v.cloud_deskewed = reinterpret_cast<const float*>(in.cloud_deskewed.data());
    v.n_valid = (int)in.cloud_deskewed.size();
    v.imuAvailable = in.imuAvailable; v.imuRollInit = in.imuRollInit; v.imuPitchInit = in.imuPitchInit; v.imuYawInit = in.imuYawInit;
    check(fbpr_set_cloud_info(ctx_->h, 0, &v, FBPR_MEM_HOST), "fbpr_set_cloud_info");
    check(fbpr_feature_extract(ctx_->h, 0, 1), "fbpr_feature_extract");   // calculateSmoothness, markOccludedPoints, extractFeatures
    // freeCloudInfoMemory (:296-303) + publishFeatureCloud (:306-315)
    cloudInfo.startRingIndex.clear(); cloudInfo.endRingIndex.clear(); cloudInfo.pointColInd.clear(); cloudInfo.pointRange.clear();
    cloudInfo.device_token = ++ctx_->generation;
    if (downloadClouds) {
        int32_t c[8]; check(fbpr_get_counts(ctx_->h, 0, c), "fbpr_get_counts");
        download(ctx_->h, FBPR_BUF_CORNER, c[2], cornerCloud);
        download(ctx_->h, FBPR_BUF_SURF, c[3], surfaceCloud);
        cloudInfo.cloud_corner = cornerCloud; cloudInfo.cloud_surface = surfaceCloud;
    } else {
        check(fbpr_sync(ctx_->h), "fbpr_sync");
    }
}

// ------------------------------------------------------------------ mapOptimization
mapOptimization::mapOptimization(const ParamServer& params, std::shared_ptr<DeviceContext> ctx) : ParamServer(params), ctx_(ctx) {}

void mapOptimization::setGlobalMap(const PointCloud& corner, const PointCloud& surf) {
    corner_GlobalMap = corner; surf_GlobalMap = surf;
    check(fbpr_set_global_map(ctx_->h, reinterpret_cast<const float*>(corner.data()), (int)corner.size(),
                              reinterpret_cast<const float*>(surf.data()), (int)surf.size(), FBPR_MEM_HOST), "fbpr_set_global_map");
}

void mapOptimization::setCurrentScan(const cloud_info& ci) {
    cloudInfo = ci;
    if (ci.device_token != 0 && ci.device_token == ctx_->generation) {
        // the feature clouds are already laserCloud{Corner,Surf}Last of slot 0 in HBM: nothing to upload
    } else {
        laserCloudCornerLast = ci.cloud_corner; laserCloudSurfLast = ci.cloud_surface;        // fromROSMsg, :272-273
        check(fbpr_set_feature_clouds(ctx_->h, 0, reinterpret_cast<const float*>(ci.cloud_corner.data()), (int)ci.cloud_corner.size(),
                                      reinterpret_cast<const float*>(ci.cloud_surface.data()), (int)ci.cloud_surface.size(), FBPR_MEM_HOST),
              "fbpr_set_feature_clouds");
    }
}

void mapOptimization::pushPose() { check(fbpr_set_pose(ctx_->h, 0, transformTobeMapped), "fbpr_set_pose"); }
void mapOptimization::pullPose() {
    int32_t it = 0; uint32_t fl = 0;
    check(fbpr_get_pose(ctx_->h, 0, transformTobeMapped, &it, &fl), "fbpr_get_pose");
    iterCount = it; flags = fl; isDegenerate = (fl & FBPR_FLAG_DEGENERATE) != 0;
    int32_t c[8]; check(fbpr_get_counts(ctx_->h, 0, c), "fbpr_get_counts");
    laserCloudCornerLastDSNum = c[4]; laserCloudSurfLastDSNum = c[5]; laserCloudCornerFromMapDSNum = c[6]; laserCloudSurfFromMapDSNum = c[7];
}

void mapOptimization::registration(const cloud_info& cloud_info_, Affine3f& pose_guess_, double stamp) {
    timeLaserCloudInfoLast = stamp;
    setCurrentScan(cloud_info_);
    if (timeLaserCloudInfoLast - timeLastProcessing >= mappingProcessInterval) {              // :279
        timeLastProcessing = timeLaserCloudInfoLast;
        // CropBox local map, pose decompose, downsampleCurrentScan, scan2MapOptimization, recompose: all on device (:284-326)
        check(fbpr_registration(ctx_->h, 0, nullptr, 0, nullptr, 0, FBPR_MEM_DEVICE, pose_guess_.m), "fbpr_registration");
        pullPose();
    }
}

void mapOptimization::extractSurroundingKeyFrames() {                                          // :964-978
    if (cloudKeyPoses6D.empty()) return;
    const PointTypePose& last = cloudKeyPoses6D.back();
    std::vector<int> sel = surroundingKeyframeIndices;
    if (sel.empty()) {
        for (int i = 0; i < (int)cloudKeyPoses6D.size(); i++) {
            const PointTypePose& p = cloudKeyPoses6D[i];
            float dx = p.x - last.x, dy = p.y - last.y, dz = p.z - last.z;
            if (std::sqrt(dx * dx + dy * dy + dz * dz) <= surroundingKeyframeSearchRadius) sel.push_back(i);
        }
    }
    std::vector<float> poses; std::vector<int32_t> coff(1, 0), soff(1, 0); PointCloud call, sall;
    for (int i : sel) {
        const PointTypePose& p = cloudKeyPoses6D[i];
        const float v[6] = { p.roll, p.pitch, p.yaw, p.x, p.y, p.z };
        poses.insert(poses.end(), v, v + 6);
        call.insert(call.end(), cornerCloudKeyFrames[i].begin(), cornerCloudKeyFrames[i].end());
        sall.insert(sall.end(), surfCloudKeyFrames[i].begin(), surfCloudKeyFrames[i].end());
        coff.push_back((int32_t)call.size()); soff.push_back((int32_t)sall.size());
    }
    const float lk[3] = { last.x, last.y, last.z };
    check(fbpr_extract_surrounding_keyframes(ctx_->h, 0, (int)sel.size(), poses.data(), reinterpret_cast<const float*>(call.data()), coff.data(),
                                             reinterpret_cast<const float*>(sall.data()), soff.data(), lk, FBPR_MEM_HOST),
          "fbpr_extract_surrounding_keyframes");
    check(fbpr_sync(ctx_->h), "fbpr_sync");
    int32_t c[8]; check(fbpr_get_counts(ctx_->h, 0, c), "fbpr_get_counts");
    laserCloudCornerFromMapDSNum = c[6]; laserCloudSurfFromMapDSNum = c[7];
}

void mapOptimization::downsampleCurrentScan() {                                                // :981-993
    check(fbpr_downsample_current_scan(ctx_->h, 0, 1), "fbpr_downsample_current_scan");
    int32_t c[8]; check(fbpr_get_counts(ctx_->h, 0, c), "fbpr_get_counts");
    laserCloudCornerLastDSNum = c[4]; laserCloudSurfLastDSNum = c[5];
}

void mapOptimization::scan2MapOptimization() {                                                 // :1403-1442 (includes transformUpdate, :1438)
    pushPose();
    check(fbpr_scan2map_optimization(ctx_->h, 0, 1), "fbpr_scan2map_optimization");
    pullPose();
}

void mapOptimization::transformUpdate() {                                                      // :1444-1479
    pushPose();
    check(fbpr_transform_update(ctx_->h, 0, 1), "fbpr_transform_update");
    pullPose();
}

void mapOptimization::syncHostClouds() {
    int32_t c[8]; check(fbpr_get_counts(ctx_->h, 0, c), "fbpr_get_counts");
    download(ctx_->h, FBPR_BUF_CORNER_DS, c[4], laserCloudCornerLastDS);
    download(ctx_->h, FBPR_BUF_SURF_DS, c[5], laserCloudSurfLastDS);
    download(ctx_->h, FBPR_BUF_MAP_CORNER, c[6], laserCloudCornerFromMapDS);
    download(ctx_->h, FBPR_BUF_MAP_SURF, c[7], laserCloudSurfFromMapDS);
}

}  // namespace feature_matching_b200

// ---- C hooks so the Python tests can drive the C++ classes exactly as ImageProjection::cloudHandler does
//      (imageProjection.cpp:203, :218): featureExtra(cloudInfo) then registration(extractor.cloudInfo, pose).
using namespace feature_matching_b200;

extern "C" __attribute__((visibility("default")))
int fm_cloud_handler(const char* params_yaml, int N_SCAN, int Horizon_SCAN,
                     const int32_t* startRing, const int32_t* endRing, const int32_t* colInd, const float* range, const float* cloud, int n_valid,
                     const float* corner_global, int nCg, const float* surf_global, int nSg,
                     float pose12[12], int share_device_clouds, int* iters, unsigned* flags, int* counts4, char* err, int errlen) {
    try {
        ParamServer ps;
        if (params_yaml && params_yaml[0] && !ps.loadYaml(params_yaml)) throw std::runtime_error("cannot read params yaml");
        ps.N_SCAN = N_SCAN; ps.Horizon_SCAN = Horizon_SCAN;
        auto ctx = std::make_shared<DeviceContext>(ps.toAbi(1, nCg + 64, nSg + 64, 0), 0);
        FeatureExtraction extrator_(ps, ctx);
        mapOptimization matcher_(ps, ctx);
        matcher_.setGlobalMap(PointCloud(reinterpret_cast<const PointType*>(corner_global), reinterpret_cast<const PointType*>(corner_global) + nCg),
                              PointCloud(reinterpret_cast<const PointType*>(surf_global), reinterpret_cast<const PointType*>(surf_global) + nSg));
        cloud_info ci;
        ci.startRingIndex.assign(startRing, startRing + N_SCAN); ci.endRingIndex.assign(endRing, endRing + N_SCAN);
        ci.pointColInd.assign(colInd, colInd + n_valid); ci.pointRange.assign(range, range + n_valid);
        ci.cloud_deskewed.assign(reinterpret_cast<const PointType*>(cloud), reinterpret_cast<const PointType*>(cloud) + n_valid);
        extrator_.downloadClouds = !share_device_clouds;
        extrator_.featureExtra(ci);
        cloud_info out = extrator_.cloudInfo;
        if (!share_device_clouds) out.device_token = 0;          // force the host round trip the ROS message implies
        Affine3f pose; std::memcpy(pose.m, pose12, sizeof(pose.m));
        matcher_.registration(out, pose, 1.0);
        std::memcpy(pose12, pose.m, sizeof(pose.m));
        if (iters) *iters = matcher_.iterCount;
        if (flags) *flags = matcher_.flags;
        if (counts4) { counts4[0] = matcher_.laserCloudCornerLastDSNum; counts4[1] = matcher_.laserCloudSurfLastDSNum;
                       counts4[2] = matcher_.laserCloudCornerFromMapDSNum; counts4[3] = matcher_.laserCloudSurfFromMapDSNum; }
        return 0;
    } catch (const std::exception& e) {
        if (err && errlen > 0) { std::strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; }
        return -1;
    }
}
