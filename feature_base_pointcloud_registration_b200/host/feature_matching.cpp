// feature_matching.cpp -- the reference's operator surface over the C ABI (see feature_matching.hpp).
// Host code only: argument marshalling, the params.yaml reader and the keyframe bookkeeping the
// reference keeps on the host.  Every numeric step is a call into libfbpr_b200.so.
#include "feature_matching.hpp"

#include <algorithm>
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
        else if (key == "surroundingKeyframeDensity") is >> surroundingKeyframeDensity;
        else if (key == "surroundingKeyframeSize") is >> surroundingKeyframeSize;
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

// ------------------------------------------------------------------ PCD IO
bool loadPCDFile(const std::string& path, PointCloud& cloud, std::string* error) {
    auto fail = [&](const char* m) { if (error) *error = m; return false; };
    std::ifstream f(path, std::ios::binary);
    if (!f) return fail("cannot open file");
    std::vector<std::string> fields; std::vector<int> sizes, counts; std::vector<char> types;
    long long points = -1, width = -1, height = 1; std::string data;
    std::string line;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty() || line[0] == '#') continue;
        std::istringstream is(line); std::string key; is >> key;
        if (key == "FIELDS") { std::string v; while (is >> v) fields.push_back(v); }
        else if (key == "SIZE") { int v; while (is >> v) sizes.push_back(v); }
        else if (key == "TYPE") { char v; while (is >> v) types.push_back(v); }
        else if (key == "COUNT") { int v; while (is >> v) counts.push_back(v); }
        else if (key == "WIDTH") is >> width; else if (key == "HEIGHT") is >> height; else if (key == "POINTS") is >> points;
        else if (key == "DATA") { is >> data; break; }
    }
    if (points < 0) points = width * height;
    if (fields.empty() || sizes.size() != fields.size() || types.size() != fields.size() || points < 0) return fail("bad PCD header");
    if (counts.empty()) counts.assign(fields.size(), 1);
    int ix = -1, iy = -1, iz = -1, ii = -1; std::vector<int> off(fields.size()); int step = 0, col = 0; std::vector<int> colOf(fields.size());
    for (size_t k = 0; k < fields.size(); k++) {
        off[k] = step; colOf[k] = col; step += sizes[k] * counts[k]; col += counts[k];
        if (fields[k] == "x") ix = (int)k; else if (fields[k] == "y") iy = (int)k; else if (fields[k] == "z") iz = (int)k; else if (fields[k] == "intensity") ii = (int)k;
    }
    if (ix < 0 || iy < 0 || iz < 0) return fail("PCD has no x y z fields");
    for (int k : { ix, iy, iz, ii }) if (k >= 0 && !(types[k] == 'F' && sizes[k] == 4)) return fail("x y z intensity must be 4-byte floats");
    cloud.assign((size_t)points, PointType{ 0, 0, 0, 0 });
    if (data == "ascii") {
        std::vector<double> row(col);
        for (long long n = 0; n < points; n++) {
            if (!std::getline(f, line)) return fail("PCD ascii data truncated");
            std::istringstream is(line);
            for (int c = 0; c < col; c++) { std::string tok; if (!(is >> tok)) return fail("PCD ascii row too short"); row[c] = std::strtod(tok.c_str(), nullptr); }
            PointType& p = cloud[(size_t)n];
            p.x = (float)row[colOf[ix]]; p.y = (float)row[colOf[iy]]; p.z = (float)row[colOf[iz]]; p.intensity = ii >= 0 ? (float)row[colOf[ii]] : 0.f;
        }
    } else if (data == "binary") {
        std::vector<char> buf((size_t)points * step);
        f.read(buf.data(), (std::streamsize)buf.size());
        if ((size_t)f.gcount() != buf.size()) return fail("PCD binary data truncated");
        for (long long n = 0; n < points; n++) {
            const char* r = buf.data() + (size_t)n * step; PointType& p = cloud[(size_t)n];
            std::memcpy(&p.x, r + off[ix], 4); std::memcpy(&p.y, r + off[iy], 4); std::memcpy(&p.z, r + off[iz], 4);
            if (ii >= 0) std::memcpy(&p.intensity, r + off[ii], 4);
        }
    } else return fail("unsupported PCD DATA mode (ascii and binary are supported)");
    return true;
}

static void pcd_header(std::ostream& o, size_t n, const char* mode) {
    o << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
      << "WIDTH " << n << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA " << mode << "\n";
}
bool savePCDFileASCII(const std::string& path, const PointCloud& cloud) {
    std::ofstream f(path, std::ios::binary);
    if (!f) return false;
    pcd_header(f, cloud.size(), "ascii");
    char buf[128];
    for (const PointType& p : cloud) {          // PCL writes 8 significant digits
        int len = std::snprintf(buf, sizeof(buf), "%.8g %.8g %.8g %.8g\n", (double)p.x, (double)p.y, (double)p.z, (double)p.intensity);
        f.write(buf, len);
    }
    return (bool)f;
}
bool savePCDFileBinary(const std::string& path, const PointCloud& cloud) {
    std::ofstream f(path, std::ios::binary);
    if (!f) return false;
    pcd_header(f, cloud.size(), "binary");
    f.write(reinterpret_cast<const char*>(cloud.data()), (std::streamsize)(cloud.size() * sizeof(PointType)));
    return (bool)f;
}

// ------------------------------------------------------------------ IMU deskew inputs
static void quat_to_rpy(double x, double y, double z, double w, double& roll, double& pitch, double& yaw) {   // tf::Matrix3x3(q).getRPY
    double d = x * x + y * y + z * z + w * w;
    if (std::fabs(d - 1.0) > 0.1) {                            // tf::quaternionMsgToTF normalises (with a warning) beyond QUATERNION_TOLERANCE
        const double l = std::sqrt(d); x /= l; y /= l; z /= l; w /= l; d = x * x + y * y + z * z + w * w;
    }
    const double s = 2.0 / d;
    const double xs = x * s, ys = y * s, zs = z * s, wx = w * xs, wy = w * ys, wz = w * zs;
    const double xx = x * xs, xy = x * ys, xz = x * zs, yy = y * ys, yz = y * zs, zz = z * zs;
    const double m00 = 1.0 - (yy + zz), m01 = xy - wz, m02 = xz + wy, m10 = xy + wz, m20 = xz - wy, m21 = yz + wx, m22 = 1.0 - (xx + yy);
    if (std::fabs(m20) >= 1) {
        yaw = 0; const double delta = std::atan2(m01, m02);
        if (m20 < 0) { pitch = M_PI / 2.0; roll = delta; } else { pitch = -M_PI / 2.0; roll = delta; }
    } else {
        pitch = -std::asin(m20);
        roll = std::atan2(m21 / std::cos(pitch), m22 / std::cos(pitch));
        yaw = std::atan2(m10 / std::cos(pitch), m00 / std::cos(pitch));
    }
}

ImuDeskewInfo imuDeskewInfo(std::vector<ImuSample>& q, double timeScanCur, double timeScanNext, int capacity) {
    ImuDeskewInfo r;                                                        // imuAvailable = false (:325)
    size_t drop = 0;
    while (drop < q.size() && q[drop].time < timeScanCur - 0.01) drop++;    // :328-335
    q.erase(q.begin(), q.begin() + (long)drop);
    if (q.empty()) return r;
    r.imuTime.assign((size_t)capacity, 0.0); r.imuRotX.assign((size_t)capacity, 0.0); r.imuRotY.assign((size_t)capacity, 0.0); r.imuRotZ.assign((size_t)capacity, 0.0);
    int cur = 0;
    for (size_t i = 0; i < q.size() && cur < capacity; i++) {
        const ImuSample& s = q[i];
        if (s.time <= timeScanCur) {                                        // :354-355
            double ro, pi, ya; quat_to_rpy(s.qx, s.qy, s.qz, s.qw, ro, pi, ya);
            r.imuRollInit = (float)ro; r.imuPitchInit = (float)pi; r.imuYawInit = (float)ya;
        }
        if (s.time > timeScanNext + 0.01) break;                            // :358-359
        if (cur == 0) { r.imuRotX[0] = r.imuRotY[0] = r.imuRotZ[0] = 0; r.imuTime[0] = s.time; ++cur; continue; }
        const double dt = s.time - r.imuTime[(size_t)cur - 1];              // :376-381
        r.imuRotX[(size_t)cur] = r.imuRotX[(size_t)cur - 1] + s.gx * dt;
        r.imuRotY[(size_t)cur] = r.imuRotY[(size_t)cur - 1] + s.gy * dt;
        r.imuRotZ[(size_t)cur] = r.imuRotZ[(size_t)cur - 1] + s.gz * dt;
        r.imuTime[(size_t)cur] = s.time;
        ++cur;
    }
    --cur;                                                                  // :385
    r.imuPointerCur = cur;
    if (cur <= 0) return r;
    r.imuAvailable = true;
    return r;
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

// mirror cloudKeyPoses6D / cornerCloudKeyFrames / surfCloudKeyFrames into the device-resident store: append the keyframes that
// are new since the last call, start over when the history shrank, refresh the poses when correctPoses (:1735-1766) moved them
void mapOptimization::syncKeyframeStore() {
    const int n = (int)cloudKeyPoses6D.size();
    int have = fbpr_keyframes_count(ctx_->h);
    check(have, "fbpr_keyframes_count");
    if (have > n || devicePoses_.size() != 6 * (size_t)have) {
        check(fbpr_keyframes_clear(ctx_->h), "fbpr_keyframes_clear");
        have = 0; devicePoses_.clear();
    }
    std::vector<float> cur(6 * (size_t)n);
    for (int i = 0; i < n; i++) {
        const PointTypePose& p = cloudKeyPoses6D[i];
        const float v[6] = { p.roll, p.pitch, p.yaw, p.x, p.y, p.z };
        std::memcpy(&cur[6 * (size_t)i], v, sizeof(v));
    }
    if (have > 0 && std::memcmp(cur.data(), devicePoses_.data(), 6 * sizeof(float) * (size_t)have) != 0)
        check(fbpr_keyframes_set_poses(ctx_->h, 0, have, cur.data()), "fbpr_keyframes_set_poses");
    for (int i = have; i < n; i++)
        check(fbpr_keyframe_push(ctx_->h, &cur[6 * (size_t)i], cloudKeyPoses6D[i].time,
                                 reinterpret_cast<const float*>(cornerCloudKeyFrames[i].data()), (int)cornerCloudKeyFrames[i].size(),
                                 reinterpret_cast<const float*>(surfCloudKeyFrames[i].data()), (int)surfCloudKeyFrames[i].size(), FBPR_MEM_HOST),
              "fbpr_keyframe_push");
    devicePoses_.swap(cur);
}

void mapOptimization::extractResident(bool loopClosure) {
    syncKeyframeStore();
    check(fbpr_extract_surrounding_keyframes_resident(ctx_->h, 0, timeLaserCloudInfoLast, surroundingKeyframeDensity, loopClosure ? 1 : 0,
                                                      surroundingKeyframeSize), "fbpr_extract_surrounding_keyframes_resident");
}
void mapOptimization::extractNearby() { extractResident(false); }                              // :872-907
void mapOptimization::extractForLoopClosure() { extractResident(true); }                       // :857-870

std::vector<int> mapOptimization::downloadSelection() {
    const int cap = 2 * (int)cloudKeyPoses6D.size() + 8;
    std::vector<float> lst(4 * (size_t)cap); std::vector<int32_t> idx((size_t)cap);
    const int K = fbpr_get_keyframe_selection(ctx_->h, lst.data(), idx.data(), cap);
    check(K, "fbpr_get_keyframe_selection");
    surroundingKeyPosesDS.clear();
    std::vector<int> out;
    for (int k = 0; k < K && k < cap; k++) {
        surroundingKeyPosesDS.push_back(PointType{ lst[4 * k], lst[4 * k + 1], lst[4 * k + 2], lst[4 * k + 3] });
        out.push_back(idx[k]);
    }
    return out;
}

void mapOptimization::extractSurroundingKeyFrames() {                                          // :964-978
    if (cloudKeyPoses6D.empty()) return;
    if (surroundingKeyframeIndices.empty()) {
        if (loopClosureEnableFlag) extractForLoopClosure(); else extractNearby();              // :970-977, all on the device
    } else {
        // caller-side selection: extractCloud (:909-955) over the named keyframes
        const PointTypePose& last = cloudKeyPoses6D.back();
        std::vector<float> poses; std::vector<int32_t> coff(1, 0), soff(1, 0); PointCloud call, sall;
        for (int i : surroundingKeyframeIndices) {
            const PointTypePose& p = cloudKeyPoses6D[i];
            const float v[6] = { p.roll, p.pitch, p.yaw, p.x, p.y, p.z };
            poses.insert(poses.end(), v, v + 6);
            call.insert(call.end(), cornerCloudKeyFrames[i].begin(), cornerCloudKeyFrames[i].end());
            sall.insert(sall.end(), surfCloudKeyFrames[i].begin(), surfCloudKeyFrames[i].end());
            coff.push_back((int32_t)call.size()); soff.push_back((int32_t)sall.size());
        }
        const float lk[3] = { last.x, last.y, last.z };
        check(fbpr_extract_cloud(ctx_->h, 0, (int)surroundingKeyframeIndices.size(), poses.data(), nullptr,
                                 reinterpret_cast<const float*>(call.data()), coff.data(),
                                 reinterpret_cast<const float*>(sall.data()), soff.data(), lk, FBPR_MEM_HOST),
              "fbpr_extract_cloud");
    }
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

// extractSurroundingKeyFrames with the class's own extractNearby (mapOptmization.h:872-955) over a keyframe store given as
// CSR arrays; returns surroundingKeyPosesDS and the resulting local map sizes (the maps stay in slot 0 of the handle).
extern "C" __attribute__((visibility("default")))
int fm_extract_surrounding(const char* params_yaml, int N_SCAN, int Horizon_SCAN, const float* keyPoses6, const double* keyTime, int nKeys,
                           double timeLaserCloudInfoLast, const float* corner_all, const int32_t* corner_off, const float* surf_all, const int32_t* surf_off,
                           float* ds_out, int ds_cap, int* n_ds, float* map_corner, int capC, float* map_surf, int capS, int* counts2, char* err, int errlen,
                           int loopClosureEnableFlag, int surroundingKeyframeSize) {
    try {
        ParamServer ps;
        if (params_yaml && params_yaml[0] && !ps.loadYaml(params_yaml)) throw std::runtime_error("cannot read params yaml");
        ps.N_SCAN = N_SCAN; ps.Horizon_SCAN = Horizon_SCAN;
        auto ctx = std::make_shared<DeviceContext>(ps.toAbi(1, capC + 64, capS + 64, corner_off[nKeys] + surf_off[nKeys] + 64), 0);
        mapOptimization m(ps, ctx);
        for (int i = 0; i < nKeys; i++) {
            PointTypePose p{}; p.roll = keyPoses6[6 * i]; p.pitch = keyPoses6[6 * i + 1]; p.yaw = keyPoses6[6 * i + 2];
            p.x = keyPoses6[6 * i + 3]; p.y = keyPoses6[6 * i + 4]; p.z = keyPoses6[6 * i + 5]; p.intensity = (float)i; p.time = keyTime[i];
            m.cloudKeyPoses6D.push_back(p);
            const PointType* c = reinterpret_cast<const PointType*>(corner_all); const PointType* s = reinterpret_cast<const PointType*>(surf_all);
            m.cornerCloudKeyFrames.emplace_back(c + corner_off[i], c + corner_off[i + 1]);
            m.surfCloudKeyFrames.emplace_back(s + surf_off[i], s + surf_off[i + 1]);
        }
        m.timeLaserCloudInfoLast = timeLaserCloudInfoLast;
        m.loopClosureEnableFlag = loopClosureEnableFlag != 0;
        if (surroundingKeyframeSize >= 0) m.surroundingKeyframeSize = surroundingKeyframeSize;
        m.extractSurroundingKeyFrames();
        m.syncHostClouds();
        m.downloadSelection();
        *n_ds = (int)m.surroundingKeyPosesDS.size();
        for (int i = 0; i < *n_ds && i < ds_cap; i++) std::memcpy(ds_out + 4 * i, &m.surroundingKeyPosesDS[i], 16);
        counts2[0] = (int)m.laserCloudCornerFromMapDS.size(); counts2[1] = (int)m.laserCloudSurfFromMapDS.size();
        if (counts2[0] > capC || counts2[1] > capS) throw std::runtime_error("output capacity too small");
        std::memcpy(map_corner, m.laserCloudCornerFromMapDS.data(), 16 * (size_t)counts2[0]);
        std::memcpy(map_surf, m.laserCloudSurfFromMapDS.data(), 16 * (size_t)counts2[1]);
        return 0;
    } catch (const std::exception& e) {
        if (err && errlen > 0) { std::strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; }
        return -1;
    }
}

// imuDeskewInfo over a queue given as 8 doubles per sample (stamp, gyro xyz, orientation xyzw); out5 = imuAvailable,
// imuPointerCur, roll, pitch, yaw; returns how many samples were popped from the front
extern "C" __attribute__((visibility("default")))
int fm_imu_deskew_info(const double* q8, int nq, double timeScanCur, double timeScanNext, int capacity,
                       double* imuTime, double* imuRotX, double* imuRotY, double* imuRotZ, double* out5) {
    std::vector<ImuSample> q((size_t)nq);
    for (int i = 0; i < nq; i++) { const double* m = q8 + 8 * i; q[(size_t)i] = ImuSample{ m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7] }; }
    ImuDeskewInfo r = imuDeskewInfo(q, timeScanCur, timeScanNext, capacity);
    for (size_t i = 0; i < r.imuTime.size(); i++) { imuTime[i] = r.imuTime[i]; imuRotX[i] = r.imuRotX[i]; imuRotY[i] = r.imuRotY[i]; imuRotZ[i] = r.imuRotZ[i]; }
    out5[0] = r.imuAvailable ? 1 : 0; out5[1] = r.imuPointerCur; out5[2] = r.imuRollInit; out5[3] = r.imuPitchInit; out5[4] = r.imuYawInit;
    return nq - (int)q.size();
}

// PCD IO hooks: mode 0 = ascii, 1 = binary
extern "C" __attribute__((visibility("default")))
int fm_save_pcd(const char* path, const float* xyzi, int n, int mode) {
    PointCloud c(reinterpret_cast<const PointType*>(xyzi), reinterpret_cast<const PointType*>(xyzi) + n);
    return (mode ? savePCDFileBinary(path, c) : savePCDFileASCII(path, c)) ? 0 : -1;
}
extern "C" __attribute__((visibility("default")))
int fm_load_pcd(const char* path, float* xyzi, int cap, char* err, int errlen) {
    PointCloud c; std::string e;
    if (!loadPCDFile(path, c, &e)) { if (err && errlen > 0) { std::strncpy(err, e.c_str(), errlen - 1); err[errlen - 1] = 0; } return -1; }
    if ((int)c.size() > cap) return -2;
    std::memcpy(xyzi, c.data(), 16 * c.size());
    return (int)c.size();
}
