// feature_matching.hpp -- C++ host layer that keeps the reference's operator surface on top of
// the C ABI (include/fbpr_b200.h).  ROS/PCL/Eigen-free stand-ins carry the same names:
//
//   reference (file:line)                                      here
//   ParamServer                 include/utility.h:61-317        feature_matching_b200::ParamServer
//   feature_matching::cloud_info msg/cloud_info.msg:1-34        feature_matching_b200::cloud_info
//   FeatureExtraction           src/featureExtraction.h:19-316  feature_matching_b200::FeatureExtraction
//     void featureExtra(const cloud_info&)            :79-81      same name, results in cloudInfo.cloud_corner/.cloud_surface,
//                                                                 cornerCloud, surfaceCloud (:30-36, :311-312)
//   mapOptimization             src/mapOptmization.h:54-1849    feature_matching_b200::mapOptimization
//     void registration(const cloud_info&, Affine3f&) :263-343    same name; pose in/out by reference
//     void extractSurroundingKeyFrames()              :964-978    same name (keyframe SELECTION = all key poses within
//                                                                 surroundingKeyframeSearchRadius of the last one unless
//                                                                 surroundingKeyframeIndices is set by the caller)
//     void downsampleCurrentScan()                    :981-993    same name
//     void scan2MapOptimization()                     :1403-1442  same name
//     void transformUpdate()                          :1444-1479  same name
//   public state: laserCloud{Corner,Surf}Last[DS] (:90-93), laserCloud{Corner,Surf}FromMapDS (:107-108),
//   transformTobeMapped[6] (:131), isDegenerate (:137), *Num counters (:140-143), cloudKeyPoses6D,
//   corner/surfCloudKeyFrames (:84-88).
// All compute runs on the GPU; clouds are mirrored to these host members only when asked
// (syncHostClouds) so a live loop pays no extra copies.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "fbpr_b200.h"

namespace feature_matching_b200 {

struct PointType { float x, y, z, intensity; };              // pcl::PointXYZI (utility.h:55)
typedef std::vector<PointType> PointCloud;
struct PointTypePose { float x, y, z, intensity, roll, pitch, yaw; double time; };   // mapOptmization.h:34-51
struct Affine3f { float m[12]; };                             // 3x4 row-major rigid transform (Eigen::Affine3f stand-in)

struct cloud_info {                                           // msg/cloud_info.msg
    std::vector<int32_t> startRingIndex, endRingIndex, pointColInd;
    std::vector<float> pointRange;
    int64_t imuAvailable = 0, odomAvailable = 0;
    float imuRollInit = 0, imuPitchInit = 0, imuYawInit = 0;
    float initialGuessX = 0, initialGuessY = 0, initialGuessZ = 0, initialGuessRoll = 0, initialGuessPitch = 0, initialGuessYaw = 0;
    int64_t imuPreintegrationResetId = 0;
    PointCloud cloud_deskewed, cloud_corner, cloud_surface;
    // not in the message: set by FeatureExtraction so that a mapOptimization sharing the same
    // device context can take the feature clouds where they already are (HBM) instead of re-uploading
    uint64_t device_token = 0;
};

// ---- PCD map IO (pcl::io::loadPCDFile / savePCDFileASCII of PointXYZI clouds; mapOptmization.h:247-257, :495-519) ----------
// FIELDS x y z intensity, SIZE 4 4 4 4, TYPE F F F F; DATA ascii (8 significant digits, as PCL writes) or binary.
bool loadPCDFile(const std::string& path, PointCloud& cloud, std::string* error = nullptr);
bool savePCDFileASCII(const std::string& path, const PointCloud& cloud);
bool savePCDFileBinary(const std::string& path, const PointCloud& cloud);

// ---- IMU-side deskew inputs (ImageProjection::imuDeskewInfo, imageProjection.cpp:323-393) ----------------------------------
struct ImuSample { double time; double gx, gy, gz; double qx, qy, qz, qw; };   // stamp, angular velocity, orientation (already in the lidar frame)
struct ImuDeskewInfo {
    bool imuAvailable = false;
    float imuRollInit = 0, imuPitchInit = 0, imuYawInit = 0;   // attitude of the last sample at or before the sweep start (imuRPY2rosRPY, utility.h:293-303)
    int imuPointerCur = 0;                                     // index of the last valid entry of the ramps
    std::vector<double> imuTime, imuRotX, imuRotY, imuRotZ;    // integrated rotation since the first kept sample
};
// queue: IMU samples in time order; entries older than timeScanCur - 0.01 are popped from its front, as the reference does
ImuDeskewInfo imuDeskewInfo(std::vector<ImuSample>& queue, double timeScanCur, double timeScanNext, int capacity = 2000);

class DeviceContext {                                         // one fbpr handle (one GPU, one stream), shared by both stages
public:
    DeviceContext(const fbpr_params& p, int device);
    ~DeviceContext();
    fbpr_handle* h = nullptr;
    uint64_t generation = 0;
};

class ParamServer {                                           // utility.h:61-317, only the knobs the path reads
public:
    int N_SCAN = 16, Horizon_SCAN = 1800;
    float edgeThreshold = 0.1f, surfThreshold = 0.1f;
    int edgeFeatureMinValidNum = 10, surfFeatureMinValidNum = 100;
    float odometrySurfLeafSize = 0.2f, mappingCornerLeafSize = 0.2f, mappingSurfLeafSize = 0.2f;
    float z_tollerance = 3.4028235e38f, rotation_tollerance = 3.4028235e38f;
    int numberOfCores = 2;
    double mappingProcessInterval = 0.15;
    float surroundingKeyframeSearchRadius = 50.0f;
    float surroundingKeyframeDensity = 1.0f;                  // utility.h:197 (params.yaml: 2.0)
    bool loopClosureEnableFlag = false;                       // utility.h:199 -> extractForLoopClosure instead of extractNearby (:970-977)
    int surroundingKeyframeSize = 50;                         // utility.h:201 (params.yaml:71: 25)
    ParamServer() {}
    explicit ParamServer(const std::string& params_yaml) { loadYaml(params_yaml); }
    bool loadYaml(const std::string& path);                   // flat `key: value` reader of config/params.yaml
    fbpr_params toAbi(int max_frames, int max_map_corner, int max_map_surf, int max_keyframe_points) const;
};

class FeatureExtraction : public ParamServer {
public:
    cloud_info cloudInfo;
    PointCloud extractedCloud, cornerCloud, surfaceCloud;
    FeatureExtraction(const ParamServer& params, std::shared_ptr<DeviceContext> ctx);
    void featureExtra(const cloud_info& cloud_info_);         // featureExtraction.h:79-81
    bool downloadClouds = true;                               // mirror corner/surface clouds to the host members
private:
    std::shared_ptr<DeviceContext> ctx_;
};

class mapOptimization : public ParamServer {
public:
    cloud_info cloudInfo;
    PointCloud laserCloudCornerLast, laserCloudSurfLast, laserCloudCornerLastDS, laserCloudSurfLastDS;
    PointCloud laserCloudCornerFromMapDS, laserCloudSurfFromMapDS;
    PointCloud corner_GlobalMap, surf_GlobalMap;              // the fork's pre-built feature maps (mapOptmization.h:245-260)
    std::vector<PointTypePose> cloudKeyPoses6D;               // keyframe store for extractSurroundingKeyFrames
    std::vector<PointCloud> cornerCloudKeyFrames, surfCloudKeyFrames;
    std::vector<int> surroundingKeyframeIndices;              // optional caller-side selection (overrides the device-side selection)
    PointCloud surroundingKeyPosesDS;                         // cloudToExtract of the last extraction (filled by downloadSelection)
    // The keyframe containers above are mirrored into a store that is RESIDENT in HBM (appended to as they grow, poses refreshed
    // when correctPoses (:1735-1766) moved them); selection AND extraction then run on the device with no host round trip:
    // extractNearby (mapOptmization.h:872-907): key poses within surroundingKeyframeSearchRadius of the last one (ascending
    // distance, ties by index), VoxelGrid(surroundingKeyframeDensity) of those poses INCLUDING the averaged intensity that the
    // reference then truncates to a keyframe index (:927), plus the key poses of the last 10 s, newest first, then extractCloud;
    // extractForLoopClosure (:857-870): the newest surroundingKeyframeSize + 1 key poses, then extractCloud.
    void extractNearby();
    void extractForLoopClosure();
    std::vector<int> downloadSelection();                     // keyframe index of every entry of cloudToExtract (-1: dropped by the re-check, :924)
    float transformTobeMapped[6] = { 0, 0, 0, 0, 0, 0 };
    bool isDegenerate = false;
    int laserCloudCornerFromMapDSNum = 0, laserCloudSurfFromMapDSNum = 0, laserCloudCornerLastDSNum = 0, laserCloudSurfLastDSNum = 0;
    int iterCount = 0;                                        // iterations executed by the last scan2MapOptimization
    unsigned flags = 0;                                       // FBPR_FLAG_*
    double timeLaserCloudInfoLast = 0, timeLastProcessing = -1;

    mapOptimization(const ParamServer& params, std::shared_ptr<DeviceContext> ctx);
    void setGlobalMap(const PointCloud& corner, const PointCloud& surf);   // replaces loadPCDFile (:247-257), keeps them in HBM
    void registration(const cloud_info& cloud_info_, Affine3f& pose_guess_, double stamp = 0.0);   // :263-343
    void extractSurroundingKeyFrames();                       // :964-978
    void downsampleCurrentScan();                             // :981-993
    void scan2MapOptimization();                              // :1403-1442
    void transformUpdate();                                   // :1444-1479
    void setCurrentScan(const cloud_info& ci);                // the fromROSMsg part of the callbacks (:271-273, :355-357)
    void syncHostClouds();                                    // fill the *DS / FromMapDS host members from HBM
private:
    std::shared_ptr<DeviceContext> ctx_;
    void pushPose();
    void pullPose();
    void syncKeyframeStore();
    void extractResident(bool loopClosure);
    std::vector<float> devicePoses_;                          // poses as uploaded (6 per keyframe)
};

}  // namespace feature_matching_b200
