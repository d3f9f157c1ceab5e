/*
 * fbpr_b200.h -- C ABI of the B200-native scan-to-map registration hot path.
 *
 * This is the drop-in boundary for the LOAM/LIO-SAM hot path of
 * qpc001/Feature_Base_Pointcloud_Registration.  The reference has no FFI layer: the path sits
 * behind the public member functions of two classes.  Each entry
 * point below names the reference interface it replaces (file:line under the reference tree).
 * The C++ host classes in feature_base_pointcloud_registration_b200/host/ keep the reference's
 * method and member names on top of this ABI; INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions
 *   - plain C types only; every pointer is caller-owned; `mem` says where it lives;
 *   - a handle owns `max_frames` independent FRAME SLOTS (one lidar frame + its local map +
 *     its pose each) on ONE device and ONE stream; every operator works on the slot range
 *     [first, first+count) in one batch of launches, so slot 0 / count 1 is the reference's
 *     single-frame call and count = F is the batched, independent-frames mode;
 *   - operators are asynchronous on the handle's stream; getters and fbpr_sync() synchronise;
 *   - return 0 on success, < 0 on a usage/CUDA error (fbpr_last_error() has the text);
 *     semantic outcomes of a frame are reported in its `flags` word;
 *   - a handle is not thread-safe (mirrors the reference's std::mutex mtx, mapOptmization.h:133);
 *     distinct handles may be used concurrently on distinct GPUs;
 *   - there is NO CPU fallback: every entry point fails if the CUDA device is unusable.
 *   - clouds must be dense (finite coordinates): the reference refuses non-dense sweeps (imageProjection.cpp:250-254);
 *     here a NaN / Inf raw point is dropped by the projection and a non-finite map point is never a neighbour;
 * Points are XYZI float4 (x, y, z, intensity), 16 bytes, the GPU layout of pcl::PointXYZI
 * (utility.h:55).  Poses are float[6] = (roll, pitch, yaw, x, y, z) = transformTobeMapped
 * (mapOptmization.h:131) unless stated otherwise.
 */
#ifndef FBPR_B200_H
#define FBPR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FBPR_API __attribute__((visibility("default")))

typedef struct fbpr_handle fbpr_handle;

enum { FBPR_MEM_HOST = 0, FBPR_MEM_DEVICE = 1 };

/* semantic outcomes of scan2MapOptimization for one frame (mapOptmization.h:1403-1442) */
enum {
    FBPR_FLAG_NOT_ENOUGH_FEATURES      = 1u,  /* :1410 gate failed; pose unchanged; transformUpdate skipped (:1439-1441) */
    FBPR_FLAG_TOO_FEW_CORRESPONDENCES  = 2u,  /* an iteration had < 50 rows (:1267-1270)                               */
    FBPR_FLAG_DEGENERATE               = 4u,  /* isDegenerate set at iteration 0 (:1346-1371)                          */
    FBPR_FLAG_CONVERGED                = 8u,  /* LMOptimization returned true before iteration 30 (:1397-1399)         */
    FBPR_FLAG_MAP_TRUNCATED            = 16u  /* no reference counterpart: the slot's local map (CropBox :284-304 or extractCloud :948-954) had more points than
                                                 max_map_corner / max_map_surf and was cut there -- the reference keeps them all, so the pose may differ  */
};

/* The knobs the path reads: include/utility.h:164-198, values of config/params.yaml:19-67. */
typedef struct fbpr_params {
    int32_t N_SCAN;
    int32_t Horizon_SCAN;
    float   edgeThreshold;
    float   surfThreshold;
    int32_t edgeFeatureMinValidNum;
    int32_t surfFeatureMinValidNum;
    float   odometrySurfLeafSize;
    float   mappingCornerLeafSize;
    float   mappingSurfLeafSize;
    float   z_tollerance;
    float   rotation_tollerance;
    int32_t numberOfCores;                   /* accepted for drop-in compatibility; unused on the GPU */
    float   surroundingKeyframeSearchRadius;
    /* capacities and device-side tuning (no reference counterpart) */
    int32_t max_frames;                      /* frame slots                                             */
    int32_t max_raw_points;                  /* raw PointXYZIRT records per frame (0 = N_SCAN*Horizon_SCAN*1.02+64) */
    int32_t max_map_corner;                  /* local-map capacity per frame (after VoxelGrid)          */
    int32_t max_map_surf;
    int32_t max_keyframe_points;             /* extractCloud concat capacity per frame and kind (0 = no keyframe operator) */
    float   knn_cell_corner;                 /* uniform-grid cell edge for the corner / surf map index (0 = 0.5 / 0.4); results do not depend on it */
    float   knn_cell_surf;
    int32_t grid_cells_corner;               /* dense-grid cell budget per map index (0 = 262144 / 1048576); the cell */
    int32_t grid_cells_surf;                 /*   edge is doubled until the map's bounding box fits the budget        */
    int32_t lm_cluster_size;                 /* CTAs cooperating on one frame's LM loop in batched calls: 1..16 (0 = per call, from the batch size) */
    int32_t lm_single_frame_mode;            /* count == 1 calls: 0 = whole GPU cooperates (grid barrier), 1 = one cluster   */
    float   knn_first_radius;                /* radius (metres) of the FIRST LM iteration's neighbour search around each point (0 = 0.35):
                                                about the expected error of the initial guess; later iterations derive it exactly */
} fbpr_params;

/* one frame's outcome: transformTobeMapped after transformUpdate, iterations executed, FBPR_FLAG_* */
typedef struct fbpr_result {
    float    pose[6];
    int32_t  iters;
    uint32_t flags;
} fbpr_result;

/* PointXYZIRT as the projection consumes it (imageProjection.cpp:8-21), packed to 24 bytes. */
typedef struct fbpr_raw_point {
    float   x, y, z, intensity;
    int32_t ring;
    float   time;
} fbpr_raw_point;

/* feature_matching::cloud_info (msg/cloud_info.msg:1-34) as a view over caller memory. */
typedef struct fbpr_cloud_info_view {
    const int32_t* startRingIndex;           /* [N_SCAN]  */
    const int32_t* endRingIndex;             /* [N_SCAN]  */
    const int32_t* pointColInd;              /* [n_valid] */
    const float*   pointRange;               /* [n_valid] */
    const float*   cloud_deskewed;           /* [n_valid] XYZI */
    int32_t        n_valid;
    int64_t        imuAvailable;
    float          imuRollInit, imuPitchInit, imuYawInit;
} fbpr_cloud_info_view;

/* ---- lifecycle ------------------------------------------------------------------------ */
/* replaces: the constructors / allocateMemory of FeatureExtraction (featureExtraction.h:43-77)
   and mapOptimization (mapOptmization.h:153-261), minus ROS and PCD IO. */
FBPR_API int  fbpr_create(const fbpr_params* params, int device, fbpr_handle** out);
FBPR_API void fbpr_destroy(fbpr_handle* h);
FBPR_API int  fbpr_sync(fbpr_handle* h);
FBPR_API const char* fbpr_last_error(void);
FBPR_API void* fbpr_stream(fbpr_handle* h);               /* cudaStream_t the handle launches on   */
FBPR_API int64_t fbpr_kernel_launches(fbpr_handle* h);    /* kernels launched (or graph-replayed) so far */
FBPR_API int  fbpr_use_graphs(fbpr_handle* h, int on);    /* capture each operator sequence once, replay after */

/* ---- inputs ---------------------------------------------------------------------------- */
/* replaces: cachePointCloud + deskewInfo outputs (imageProjection.cpp:229-301, :303-393): the raw
   cloud and the integrated IMU rotation ramp of one sweep.  imu_* may be NULL when imuAvailable == 0. */
FBPR_API int fbpr_set_raw_scan(fbpr_handle* h, int slot, const fbpr_raw_point* pts, int n, int mem,
                               int64_t imuAvailable, int deskewFlag, double timeScanCur,
                               const double* imuTime, const double* imuRotX, const double* imuRotY,
                               const double* imuRotZ, int imuPointerCur,
                               float imuRollInit, float imuPitchInit);
/* byte layout of one sensor_msgs/PointCloud2 record (the fields cachePointCloud checks, imageProjection.cpp:262-297) */
typedef struct fbpr_pc2_layout {
    int32_t point_step;                      /* bytes per point (22 for the Velodyne driver, 32 for pcl::toROSMsg<PointXYZIRT>) */
    int32_t off_x, off_y, off_z;             /* float32 fields                                                                  */
    int32_t off_intensity;                   /* float32, -1 = absent (0 is used)                                                */
    int32_t off_ring, ring_bytes;            /* unsigned ring: 1, 2 or 4 bytes; ring_bytes = 0 means absent (the reference refuses such clouds) */
    int32_t off_time;                        /* float32 seconds since sweep start, -1 = absent (deskew disabled, deskewFlag = -1) */
} fbpr_pc2_layout;
/* replaces: pcl::fromROSMsg(currentCloudMsg, *laserCloudIn) + the field checks of cachePointCloud (imageProjection.cpp:252-297):
   the message bytes are uploaded as they are and repacked on the device.  Other arguments as fbpr_set_raw_scan
   (deskewFlag is derived: 1 with a time field, -1 without). */
FBPR_API int fbpr_set_raw_scan_pc2(fbpr_handle* h, int slot, const void* data, int n, const fbpr_pc2_layout* layout, int mem,
                                   int64_t imuAvailable, double timeScanCur,
                                   const double* imuTime, const double* imuRotX, const double* imuRotY,
                                   const double* imuRotZ, int imuPointerCur,
                                   float imuRollInit, float imuPitchInit);
/* replaces: the cloud_info message handed to featureExtra() (featureExtraction.h:79-92). */
FBPR_API int fbpr_set_cloud_info(fbpr_handle* h, int slot, const fbpr_cloud_info_view* ci, int mem);
/* replaces: fromROSMsg(cloud_corner / cloud_surface) into laserCloud{Corner,Surf}Last (mapOptmization.h:272-273). */
FBPR_API int fbpr_set_feature_clouds(fbpr_handle* h, int slot, const float* corner_xyzi, int n_corner,
                                     const float* surf_xyzi, int n_surf, int mem);
/* replaces: laserCloud{Corner,Surf}FromMapDS as produced by CropBox (mapOptmization.h:284-304) or extractCloud (:948-954). */
FBPR_API int fbpr_set_local_map(fbpr_handle* h, int slot, const float* corner_xyzi, int n_corner,
                                const float* surf_xyzi, int n_surf, int mem);
/* replaces: transformTobeMapped[] initialisation (mapOptmization.h:309-310 / updateInitialGuess). */
FBPR_API int fbpr_set_pose(fbpr_handle* h, int slot, const float pose6[6]);
FBPR_API int fbpr_set_poses(fbpr_handle* h, int first, int count, const float* pose6, int mem);

/* one frame's inputs for the batched upload below; pointers may be NULL when a stage is not used */
typedef struct fbpr_frame_input {
    const fbpr_raw_point* raw;  int32_t n_raw;
    int32_t        deskewFlag;               /* -1: cloud has no time field (imageProjection.cpp:548) */
    int64_t        imuAvailable;
    double         timeScanCur;
    const double*  imuTime; const double* imuRotX; const double* imuRotY; const double* imuRotZ;
    int32_t        imuPointerCur;
    float          imuRollInit, imuPitchInit;
    const float*   map_corner_xyzi; int32_t n_map_corner;
    const float*   map_surf_xyzi;   int32_t n_map_surf;
    float          pose[6];
    int32_t        raw_format;               /* FBPR_RAW_*: layout of the records behind `raw`                              */
    int32_t        map_format;               /* FBPR_MAP_*: layout of the points behind map_corner_xyzi / map_surf_xyzi     */
} fbpr_frame_input;
/* wire formats of the batched uploads: what crosses PCIe is the caller's layout, the repack runs on the device.
   FBPR_RAW_VELODYNE22 = the Velodyne driver's PointCloud2 record of PointXYZIRT (imageProjection.cpp:8-21): x, y, z, intensity
   float32, ring uint16, time float32, 22 bytes, no padding.  FBPR_MAP_XYZ12 = x, y, z float32, 12 bytes: the registration never
   reads a map point's intensity (mapOptmization.h:1028-1036, :1157-1163), so pose-only callers need not ship it (it reads as 0). */
enum { FBPR_RAW_PACKED24 = 0, FBPR_RAW_VELODYNE22 = 1 };
enum { FBPR_MAP_XYZI16 = 0, FBPR_MAP_XYZ12 = 1,
       FBPR_MAP_FROM_GLOBAL = 2 };           /* no map is uploaded: the slot's local map is the CropBox (+-30/+-30/+-10 m around the frame's
                                                pose guess, mapOptmization.h:284-304) of the maps made resident by fbpr_set_global_map */
/* batched form of fbpr_set_raw_scan + fbpr_set_local_map + fbpr_set_pose for `count` independent frames
   (BASELINE config 4: 1024 frames, each against its own local map): one async copy per cloud, one packed
   copy for all the scalars.  The slots' counters and results are reset. */
FBPR_API int fbpr_set_frames(fbpr_handle* h, int first, int count, const fbpr_frame_input* frames, int mem);

/* ---- operators (slot range [first, first+count)) ----------------------------------------- */
/* replaces: ImageProjection::projectPointCloud + cloudExtraction (imageProjection.cpp:583-670). */
FBPR_API int fbpr_project(fbpr_handle* h, int first, int count);
/* replaces: FeatureExtraction::featureExtra (featureExtraction.h:79-103): calculateSmoothness,
   markOccludedPoints, extractFeatures; results become laserCloud{Corner,Surf}Last of the slot. */
FBPR_API int fbpr_feature_extract(fbpr_handle* h, int first, int count);
/* replaces: mapOptimization::extractSurroundingKeyFrames -> extractCloud (mapOptmization.h:909-978):
   transform K selected keyframes by their poses, concatenate in list order, VoxelGrid both kinds.
   The caller names the keyframes here (fbpr_extract_surrounding_keyframes_resident below selects them on the device).
   Clouds are concatenated with
   CSR offsets (K+1 entries). */
FBPR_API int fbpr_extract_surrounding_keyframes(fbpr_handle* h, int slot, int K, const float* key_poses6,
                                                const float* corner_xyzi, const int32_t* corner_off,
                                                const float* surf_xyzi, const int32_t* surf_off,
                                                const float last_key_xyz[3], int mem);
/* the same with explicit positions for the distance re-check of extractCloud (mapOptmization.h:924): after extractNearby
   (:872-907) entry i of the list is a VoxelGrid-averaged key pose at check_xyz[3i..3i+2] whose truncated averaged intensity
   names the keyframe whose pose (key_poses6) and clouds are used.  check_xyz = NULL: the keyframes' own positions. */
FBPR_API int fbpr_extract_cloud(fbpr_handle* h, int slot, int K, const float* key_poses6, const float* check_xyz,
                                const float* corner_xyzi, const int32_t* corner_off,
                                const float* surf_xyzi, const int32_t* surf_off,
                                const float last_key_xyz[3], int mem);
/* replaces: the keyframe containers cloudKeyPoses3D / cloudKeyPoses6D / cornerCloudKeyFrames / surfCloudKeyFrames
   (mapOptmization.h:84-88) by a store that is RESIDENT in HBM.  fbpr_keyframe_push = the push_backs of saveKeyFramesAndFactor
   (:1690, :1700, :1725-1726): pose6 = (roll, pitch, yaw, x, y, z), time = cloudKeyPoses6D[i].time, the two clouds in the lidar
   frame; returns the keyframe's index (= cloudKeyPoses3D[i].intensity, :1689).  fbpr_keyframes_set_poses = correctPoses (:1735-1766)
   writing optimised poses back.  fbpr_keyframes_clear empties the store (capacity is kept). */
FBPR_API int fbpr_keyframes_clear(fbpr_handle* h);
FBPR_API int fbpr_keyframes_count(fbpr_handle* h);
FBPR_API int fbpr_keyframe_push(fbpr_handle* h, const float pose6[6], double time, const float* corner_xyzi, int n_corner,
                                const float* surf_xyzi, int n_surf, int mem);
FBPR_API int fbpr_keyframes_set_poses(fbpr_handle* h, int first, int count, const float* pose6);
/* replaces: mapOptimization::extractSurroundingKeyFrames (mapOptmization.h:964-978) END TO END on the device over the resident
   store: extractNearby (:872-907: radius search around the newest key pose, VoxelGrid(surroundingKeyframeDensity) of the hits,
   the key poses of the last 10 s) or, when loopClosureEnableFlag != 0, extractForLoopClosure (:857-870: the newest
   surroundingKeyframeSize + 1 key poses), then extractCloud (:909-955) into the slot's local map.  No host round trip; an
   empty store leaves the local map untouched (:966-967).  Needs max_keyframe_points > 0. */
FBPR_API int fbpr_extract_surrounding_keyframes_resident(fbpr_handle* h, int slot, double timeLaserCloudInfoLast,
                                                         float surroundingKeyframeDensity, int loopClosureEnableFlag,
                                                         int surroundingKeyframeSize);
/* parity getter: cloudToExtract of the last resident extraction (surroundingKeyPosesDS + the last-10-s poses, or the loop-closure
   list) and, per entry, the keyframe it names (-1: dropped by the distance re-check, :924).  Returns the list length. */
FBPR_API int fbpr_get_keyframe_selection(fbpr_handle* h, float* list_xyzi, int32_t* key_index, int cap);
/* debugging aid (no reference counterpart): when the handle was created with FBPR_GUARD=1 in the environment every device
   array it owns sits between two 256-byte guard zones; returns how many of them were overwritten since (0 = no out-of-bounds
   write past any array end), < 0 on error. */
FBPR_API int fbpr_debug_check_guards(fbpr_handle* h);
/* replaces: mapOptimization::downsampleCurrentScan (mapOptmization.h:981-993). */
FBPR_API int fbpr_downsample_current_scan(fbpr_handle* h, int first, int count);
/* replaces: mapOptimization::scan2MapOptimization (mapOptmization.h:1403-1442) including the two
   per-frame kd-tree builds (:1413-1414, here: uniform-grid index builds), the <= 30 iteration loop
   of cornerOptimization / surfOptimization / combineOptimizationCoeffs / LMOptimization, and
   transformUpdate (:1444-1479).  No host round trip inside. */
FBPR_API int fbpr_scan2map_optimization(fbpr_handle* h, int first, int count);
/* replaces: mapOptimization::transformUpdate on its own (mapOptmization.h:1444-1479). */
FBPR_API int fbpr_transform_update(fbpr_handle* h, int first, int count);
/* replaces: mapOptimization::registration (mapOptmization.h:263-343) for one slot: CropBox of the
   given GLOBAL maps around pose (+-30/+-30/+-10 m), Affine -> (rpy,xyz), downsampleCurrentScan,
   scan2MapOptimization, back to Affine.  pose12 is a 3x4 row-major rigid transform, in/out.
   The slot's feature clouds must be present (fbpr_feature_extract or fbpr_set_feature_clouds). */
/* replaces: the map load of the fork's constructor (mapOptmization.h:245-260, minus PCD IO and its VoxelGrid):
   keeps corner_GlobalMap / surf_GlobalMap resident in HBM; fbpr_registration() then takes NULL maps. */
FBPR_API int fbpr_set_global_map(fbpr_handle* h, const float* corner_xyzi, int n_corner,
                                 const float* surf_xyzi, int n_surf, int mem);
/* replaces: the CropBox block of registration() (mapOptmization.h:284-304) for a RANGE of slots at once: every slot's local map
   becomes the crop of the resident global maps (fbpr_set_global_map) around that slot's current pose translation. */
FBPR_API int fbpr_crop_local_maps(fbpr_handle* h, int first, int count);
FBPR_API int fbpr_registration(fbpr_handle* h, int slot, const float* corner_global_xyzi, int n_corner,
                               const float* surf_global_xyzi, int n_surf, int mem, float pose12[12]);
/* the whole per-frame path in one call: project -> feature_extract -> downsample -> scan2map */
FBPR_API int fbpr_run_frames(fbpr_handle* h, int first, int count, int with_projection, int with_features);
/* the same whole path for `count` resident frames in batches of `batch_frames` (0 = 128), with the front-end of batch k+1 on the
   handle's stream overlapping the map index + LM loop of batch k on a second stream.  Results are identical to fbpr_run_frames
   batch by batch; work queued on the handle's stream afterwards is ordered behind it. */
FBPR_API int fbpr_run_frames_pipelined(fbpr_handle* h, int first, int count, int batch_frames);

/* the batch form of the caller's per-frame loop (cloudHandler -> featureExtra -> registration, imageProjection.cpp:182-226)
   for `count` INDEPENDENT frames given in HOST memory (pinned for full PCIe speed): uploads are issued in chunks of
   `chunk_frames` (0 = 32) on a second stream, so the copies of chunk k+1 run under the kernels of chunk k; the whole path
   (projection, features, downsample, map index, LM, transformUpdate) runs per chunk; results land in `out` (host) and the
   call returns when they are there. */
FBPR_API int fbpr_register_frames(fbpr_handle* h, int first, int count, const fbpr_frame_input* frames, int chunk_frames,
                                  fbpr_result* out);
/* the same call split in two so that a caller streaming batches keeps PCIe busy: _begin enqueues the uploads and the whole
   path of `count` frames into slots [first, first+count) and returns a ticket (>= 0) at once; _end blocks until that batch's
   results are on the host, copies them to `out` and returns the frame count.  Up to FBPR_MAX_TICKETS batches may be in flight
   on DISJOINT slot ranges (double buffering: the uploads of batch k+1 run under the last kernels of batch k); the frames'
   host buffers must stay valid until _end.  No other operator may touch those slots between _begin and _end.
   chunk_frames: frames per upload chunk; 0 = 32. */
#define FBPR_MAX_TICKETS 4
FBPR_API int fbpr_register_frames_begin(fbpr_handle* h, int first, int count, const fbpr_frame_input* frames, int chunk_frames);
FBPR_API int fbpr_register_frames_end(fbpr_handle* h, int ticket, fbpr_result* out);

/* ---- results ----------------------------------------------------------------------------- */
FBPR_API int fbpr_get_pose(fbpr_handle* h, int slot, float pose6[6], int32_t* iters, uint32_t* flags);
FBPR_API int fbpr_get_results(fbpr_handle* h, int first, int count, fbpr_result* out, int mem);
/* counts[8] = n_raw, n_valid, n_corner, n_surf, n_corner_ds, n_surf_ds, n_map_corner, n_map_surf */
FBPR_API int fbpr_get_counts(fbpr_handle* h, int slot, int32_t counts[8]);

/* per-stage device time, measured with CUDA events on the handle's stream around each stage's launches
   (replaces the reference's TicToc around scan2MapOptimization, mapOptmization.h:315-318, tic_toc.hpp:14-33).
   Only active when graphs are off.  ms / calls are accumulated since the last reset. */
enum { FBPR_STAGE_PROJECT = 0, FBPR_STAGE_FEATURES, FBPR_STAGE_DOWNSAMPLE, FBPR_STAGE_MAP_INDEX, FBPR_STAGE_LM, FBPR_STAGE_COUNT };
FBPR_API int fbpr_enable_stage_timing(fbpr_handle* h, int on);
FBPR_API int fbpr_get_stage_ms(fbpr_handle* h, float ms[FBPR_STAGE_COUNT], int32_t calls[FBPR_STAGE_COUNT], int reset);

/* ---- parity / debug getters (host destinations) --------------------------------------------- */
enum {
    FBPR_BUF_START_RING = 0, FBPR_BUF_END_RING, FBPR_BUF_COL_IND, FBPR_BUF_RANGE, FBPR_BUF_CLOUD,
    FBPR_BUF_CURVATURE, FBPR_BUF_PICKED, FBPR_BUF_LABEL,
    FBPR_BUF_CORNER, FBPR_BUF_CORNER_INDEX, FBPR_BUF_SURF, FBPR_BUF_RING_SURF_COUNT, FBPR_BUF_RING_SURF_COUNT_DS,
    FBPR_BUF_CORNER_DS, FBPR_BUF_SURF_DS, FBPR_BUF_MAP_CORNER, FBPR_BUF_MAP_SURF,
    FBPR_BUF_KNN_CORNER, FBPR_BUF_KNN_SURF, FBPR_BUF_KNN_D2_CORNER, FBPR_BUF_KNN_D2_SURF,
    FBPR_BUF_COEFF_CORNER, FBPR_BUF_COEFF_SURF, FBPR_BUF_FLAG_CORNER, FBPR_BUF_FLAG_SURF,
    FBPR_BUF_ATA, FBPR_BUF_ATB, FBPR_BUF_X, FBPR_BUF_POSE_TRACE, FBPR_BUF_WINNER_RAW
};
/* copies buffer `which` of `slot` to dst (at most cap_bytes); returns the byte count, < 0 on error */
FBPR_API int64_t fbpr_get_buffer(fbpr_handle* h, int slot, int which, void* dst, int64_t cap_bytes);
/* buffer `which` (a point cloud: CLOUD, CORNER, SURF, *_DS, MAP_*) as 32-byte pcl::PointXYZI records, the payload of the
   PointCloud2 that publishCloud / pcl::toROSMsg produce (utility.h:255-264); repacked on the device.  Returns bytes. */
FBPR_API int64_t fbpr_get_buffer_xyzi32(fbpr_handle* h, int slot, int which, void* dst, int64_t cap_bytes);
/* the inverse for inputs: 32-byte pcl::PointXYZI records (pcl::fromROSMsg, mapOptmization.h:272-273; a loaded PCD map) as the
   slot's feature clouds / local map.  kind: 0 = feature clouds (corner, surf), 1 = local map (corner, surf). */
FBPR_API int fbpr_set_clouds_xyzi32(fbpr_handle* h, int slot, int kind, const void* corner32, int n_corner, const void* surf32, int n_surf);
/* capture per-point kNN / coefficients / AtA / AtB / X of LM iteration `iter` on the next scan2map (-1 = off) */
FBPR_API int fbpr_set_debug_iteration(fbpr_handle* h, int iter);

/* stand-alone VoxelGrid (pcl::VoxelGrid<PointXYZI>::filter; call sites featureExtraction.h:289-290,
   mapOptmization.h:251-257,:948-953,:985-991).  out_xyzi capacity n; point_keys (n) and out_keys
   (n) may be NULL.  Returns the number of output points, < 0 on error. */
FBPR_API int fbpr_voxel_grid(fbpr_handle* h, const float* xyzi, int n, float leaf, float* out_xyzi,
                             int32_t* point_keys, int32_t* out_keys, int mem);
/* stand-alone exact 5-NN (pcl::KdTreeFLANN::nearestKSearch k=5, mapOptmization.h:1020,:1143) of nq
   XYZ queries against an XYZI map: idx/d2 are nq x 5, ascending (d2, idx); neighbours are exact
   inside the 1 m ball, entries whose 5th distance is >= 1 m^2 are marked by idx = -1. */
FBPR_API int fbpr_knn5(fbpr_handle* h, const float* map_xyzi, int n_map, float cell, const float* q_xyz,
                       int nq, int32_t* idx, float* d2, int mem);
/* first search-cube radius (grid cells) fbpr_knn5 starts from (default 1).  Results do not depend on it; the LM
   kernel derives the radius per point from the previous iteration, and the knob lets tests pin several starts. */
FBPR_API int fbpr_knn5_first_radius(fbpr_handle* h, int cells);

/* test hook: runs the DEVICE small-matrix routines of the LM kernel on n caller-supplied problems (host buffers), one thread each,
   so that the tests can compare them directly with OpenCV's own results (cv::eigen mapOptmization.h:1060 / :1353,
   cv::solve(DECOMP_QR) :1343, cv::Mat::inv :1370) and with Eigen's colPivHouseholderQr (:1169).  Row widths (floats) in / out:
   JACOBI3 9 / 12 (W, V rows), JACOBI6 36 / 42, QR6 42 (A then b) / 6, LU6 36 / 36, PLANE5X3 15 / 3, NOT_DEGENERATE 36 / 1. */
enum { FBPR_SELFTEST_JACOBI3 = 0, FBPR_SELFTEST_JACOBI6, FBPR_SELFTEST_QR6, FBPR_SELFTEST_LU6, FBPR_SELFTEST_PLANE5X3, FBPR_SELFTEST_NOT_DEGENERATE,
       FBPR_SELFTEST_QR6_WARP /* the warp-parallel form of QR6 that the LM kernel runs */,
       FBPR_SELFTEST_SINCOS   /* in: one angle, out: sin, cos under the f32 trig contract (float)sin((double)x) -- pcl::getTransformation's trig */ };
FBPR_API int fbpr_selftest_smallmat(fbpr_handle* h, int which, const float* in, int n, float* out);

#ifdef __cplusplus
}
#endif
#endif /* FBPR_B200_H */
