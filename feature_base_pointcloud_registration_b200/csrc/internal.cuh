// internal.cuh -- device-side data layout shared by the kernels of libfbpr_b200.so.
//
// Layout in HBM (DESIGN.md "Data layout"): every per-frame array is stored slot-major with a
// fixed per-slot stride (capacity), points are float4 XYZI (16 B), indices / keys are int32.
// Sizes that depend on the data (n_valid, n_corner, ...) live in FrameMeta in device memory and
// are read by the kernels themselves, so no operator ever needs a host round trip.
//
// Arithmetic contract (mirrors the CPU reference op-for-op): this
// library is compiled with -fmad=false -prec-div=true -prec-sqrt=true and never uses fast
// intrinsics on the parity path; f32 trig is (float)sin((double)x).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fbpr_b200.h"

#define FBPR_IMU_CAP 512          // IMU ramp samples kept per frame
#define FBPR_MAX_ITERS 30         // mapOptmization.h:1417
#define FBPR_CORNERS_PER_SEG 20   // featureExtraction.h:217
#define FBPR_SEGS 6               // featureExtraction.h:192

struct FrameMeta {
    int n_raw, n_valid, n_corner, n_surf, n_corner_ds, n_surf_ds, n_map_corner, n_map_surf;
    int mapTruncated;             // the slot's local map lost points to max_map_corner / max_map_surf (CropBox, extractCloud); sits behind the map counts so one copy resets all three
    int first_valid_raw;          // min raw index that passed the projection gates (defines transStartInverse)
    int deskewFlag, imuPointerCur;
    long long imuAvailable;
    double timeScanCur;
    float imuRollInit, imuPitchInit;
    float pose[6];                // transformTobeMapped: roll, pitch, yaw, x, y, z
    int iters;
    unsigned flags;
    int nSel;                     // rows of the last executed iteration
    int isDegenerate;
};

// uniform-grid index of one local map (replaces the per-frame FLANN kd-tree, mapOptmization.h:1413-1414)
struct GridDesc {
    float ox, oy, oz;             // origin = bbox min
    float h, inv_h;               // cell edge
    int dx, dy, dz;               // cells per axis
    int ncells;
    int n;                        // points indexed
    int rmax;                     // shells needed to cover the 1 m ball
    int pad;
};

// one VoxelGrid problem (pcl::VoxelGrid restatement)
struct VoxDesc {
    int n;                        // input points
    int overflow;                 // dx*dy*dz > INT_MAX: output = input
    int npass;                    // 8-bit radix passes needed for the key range
    int min_b[3];
    int div_b[3];
    float inv_leaf;
    int n_out;
    int pad;
};

struct VoxSeg {                   // host-built descriptor of one VoxelGrid segment
    const float4* in;             // input cloud
    const int* n_in;              // device pointer to its size
    float4* out;
    int* n_out;
    float leaf;
    int cap;                      // capacity of in / scratch
    int out_cap;                  // capacity of out (voxels beyond it are dropped and *truncated is raised)
    int* truncated;               // optional flag word (FrameMeta::mapTruncated of the slot that owns `out`)
    unsigned* key[2];             // ping-pong keys
    unsigned* val[2];             // ping-pong point indices
    unsigned* tile_hist;          // [256][tiles_cap]
    unsigned* bbox;               // 6 ordered-uint encoded floats: min xyz, max xyz
    int* run_tile;                // [tiles_cap+1] run starts per tile / scanned
    VoxDesc* desc;
    int* point_keys;              // optional: per input point key (parity getter)
    int* out_keys;                // optional: per output voxel key
};

struct GridSeg {                  // host-built descriptor of one map-index segment
    const float4* pts;            // map points (XYZI, original order)
    const int* n;                 // device pointer to count
    float4* sorted;               // cell-contiguous copy, w = original index bits
    int* cell_start;              // [cells_cap+1]
    int* cell_of;                 // [cap] rank of the point inside its cell (the value its counting atomic returned)
    int* tile_sum;                // scan scratch
    unsigned* bbox;               // 6 encoded floats
    GridDesc* desc;
    float h0;                     // requested cell edge
    int cap, cells_cap;
};


// resident keyframe store: cloudKeyPoses3D / cloudKeyPoses6D + cornerCloudKeyFrames / surfCloudKeyFrames (mapOptmization.h:84-88),
// appended to by fbpr_keyframe_push (saveKeyFramesAndFactor, :1690-1726); key pose i has intensity = i (:1689)
struct KfStoreView {
    const float* pose6;           // [n][6] roll, pitch, yaw, x, y, z
    const double* time;           // [n] cloudKeyPoses6D[i].time
    const int* off[2];            // [n+1] CSR offsets into the pools: 0 corner, 1 surf
    const float4* pool[2];        // keyframe clouds, lidar frame
    int n;
};
// scratch of one keyframe selection (extractNearby :872-907 / extractForLoopClosure :857-870), sized for the store's pose capacity
struct KfSelect {
    unsigned long long* keys;     // [pow2 >= cap] radius hits as (d^2 bits << 32 | index)
    int* counters;                // [0] radius hits, [1] surroundingKeyPosesDS size (VoxelGrid output), [2] K = entries of cloudToExtract
    float4* hitPts;               // [cap] surroundingKeyPoses: the hits in ascending (d^2, index) order
    float4* list;                 // [2 cap] cloudToExtract: surroundingKeyPosesDS followed by the key poses of the last 10 s
    int* selIdx;                  // [2 cap] keyframe named by entry k ((int)intensity, :927), -1 = dropped by the distance re-check (:924)
    int* outoff[2];               // [2 cap + 1] output offsets of the entries' clouds in the concatenation
    float* T;                     // [2 cap][12] pcl::getTransformation of the named keyframes' poses
    int cap;
};

// pieces of one dense host span that was uploaded with a single copy and is now scattered to its slots (capi.cu, mapops.cu)
#define FBPR_SCATTER_MAX 112
// kind (bits 60..63 of `bytes`): 0 = plain copy (offsets / sizes multiples of 4), 1 = 22-byte Velodyne wire records -> 24-byte
// fbpr_raw_point, 2 = 12-byte XYZ -> float4 XYZI with intensity 0.  `bytes` counts SOURCE bytes.
#define FBPR_PIECE_COPY 0
#define FBPR_PIECE_WIRE22 1
#define FBPR_PIECE_XYZ12 2
struct ScatterPiece { unsigned long long src_off; void* dst; unsigned long long bytes; };
struct ScatterTable { const unsigned char* stage; int n; int pad; ScatterPiece p[FBPR_SCATTER_MAX]; };

// ---- kernel argument blocks (passed by value) ----------------------------------------------
struct ProjArgs {
    FrameMeta* meta;
    const fbpr_raw_point* raw; int rawCap;
    const double* imuTime; const double* imuRotX; const double* imuRotY; const double* imuRotZ;   // [slot][FBPR_IMU_CAP]
    int* pix;                     // [slot][P] winning raw index per pixel
    int* ringCount;               // [slot][N_SCAN]
    int* startRing; int* endRing; // [slot][N_SCAN]
    int* colInd; float* range; float4* cloud; int* winner;   // [slot][P]
    int N_SCAN, H, P;
    int first;
};

struct FeatArgs {
    FrameMeta* meta;
    const int* startRing; const int* endRing;     // [slot][N_SCAN]
    const int* colInd; const float* range; const float4* cloud;   // [slot][P]
    float* curv; int* picked; int* label;         // [slot][P]
    int* ringCorner;                              // [slot][N_SCAN] corners per ring
    int* cornerStage;                             // [slot][N_SCAN][120] indices
    int* ringSurf; int* ringSurfDS;               // [slot][N_SCAN]
    float4* surfStage;                            // [slot][P] ring r at r*H
    float4* corner; int* cornerIndex; int cornerCap;   // [slot][cornerCap]
    float4* surf;                                 // [slot][P]
    int N_SCAN, H, P;
    float edgeThreshold, surfThreshold, leaf;
    int segPad, voxPad, wcap;                     // pow2 paddings, window capacity
    int first;
};

struct LmArgs {
    FrameMeta* meta;
    const float4* cornerDS; int cornerCap;
    const float4* surfDS; int surfCap;
    const GridSeg* gsegs;         // [2*slot + kind]
    float4* qanchor; int* qcache; int qCap;   // [slot][qCap] per feature point: position + bound of its last full search, [..][16] cached map indices
    float firstRadius;            // metres the first iteration's search cube must cover
    double* partials; int teamMax;     // [slot][2][teamMax][28] per-CTA partial sums, double-buffered by iteration parity
    double* partialsGrid; int gridMax; // [slot][2][gridMax][28] same for the whole-GPU single-frame variant
    double* chunkPart; int chunkCap;   // [slot][chunkCap][28] per-chunk partial sums of the batched variant (chunk = 32 feature points)
    int first;
    int edgeMin, surfMin;
    float z_tol, rot_tol;
    // debug capture (slots < dbgSlots only)
    int debug_iter, dbgSlots;
    int* knnC; float* d2C; float4* coeffC; unsigned char* flagC;
    int* knnS; float* d2S; float4* coeffS; unsigned char* flagS;
    float* dbgAtA; float* dbgAtB; float* dbgX;      // [slot][36], [slot][6], [slot][6]
    float* poseTrace;                               // [slot][30][6]
};

// ordered-uint encoding of floats for atomicMin/atomicMax
__host__ __device__ inline unsigned f2ord(float f) {
    unsigned u;
#ifdef __CUDA_ARCH__
    u = __float_as_uint(f);
#else
    union { float f; unsigned u; } c; c.f = f; u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float ord2f(unsigned u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; unsigned u; } c; c.u = u; return c.f;
#endif
}

// f32 trig contract
__device__ inline float sinf_c(float x) { return (float)sin((double)x); }
__device__ inline float cosf_c(float x) { return (float)cos((double)x); }
// Inside |x| <= pi/4 (every deskew angle of a sweep, most pose angles) no argument reduction is needed and the two kernels of
// fdlibm (k_sin.c / k_cos.c: minimax polynomials in x^2, error < 1 ulp of f64) give the same f32 after rounding as glibc's sin / cos --
// checked on the CPU with the same fused multiply-add sequence (explicit __fma_rn here, so -fmad=false does not touch it) over 4e8
// arguments, uniform and log-uniform in [-pi/4, pi/4]: 0 differences.  About 20 f64
// operations per pair instead of the general sincos(double)'s reduction + selection; proj_compact evaluates three pairs per point.
#ifndef FBPR_FAST_TRIG
#define FBPR_FAST_TRIG 1
#endif
__device__ inline void sincosf_c(float x, float& s, float& c) {      // both at once: one argument reduction instead of two
    const double xd = (double)x;
#if FBPR_FAST_TRIG
    if (fabsf(x) <= 0.78539816f) {
        const double z = xd * xd, v = z * xd;
        const double rs = __fma_rn(z, __fma_rn(z, __fma_rn(z, __fma_rn(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08), 2.75573137070700676789e-06),
                                               -1.98412698298579493134e-04), 8.33333333332248946124e-03);
        s = fabsf(x) < 7.4505806e-9f ? x : (float)__fma_rn(v, __fma_rn(z, rs, -1.66666666666666324348e-01), xd);      // |x| < 2^-27: sin x = x (keeps -0, as fdlibm and glibc do)
        const double rc = z * __fma_rn(z, __fma_rn(z, __fma_rn(z, __fma_rn(z, __fma_rn(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09), -2.75573143513906633035e-07),
                                                               2.48015872894767294178e-05), -1.38888888888741095749e-03), 4.16666666666666019037e-02);
        const double ax = fabs(xd), zr = z * rc;
        if (ax < 0.3) c = (float)(1.0 - __fma_rn(0.5, z, -zr));
        else {
            const double q0 = ax > 0.78125 ? 0.28125 : ax * 0.25;
            const double qx = __hiloint2double(__double2hiint(q0), 0);           // ~|x| / 4 with a short mantissa: 1 - qx is exact
            const double hz = __fma_rn(0.5, z, -qx), a = 1.0 - qx;
            c = (float)(a - (hz - zr));
        }
        return;
    }
#endif
    double sd, cd; sincos(xd, &sd, &cd); s = (float)sd; c = (float)cd;
}

// pcl::getTransformation(x,y,z,roll,pitch,yaw) -> 3x4 row-major, f32
__device__ inline void get_transformation(float x, float y, float z, float roll, float pitch, float yaw, float T[12]) {
    float A, B, C, D, E, F;
    sincosf_c(yaw, B, A); sincosf_c(pitch, D, C); sincosf_c(roll, F, E);
    float DE = D * E, DF = D * F;
    T[0] = A * C; T[1] = A * DF - B * E; T[2]  = B * F + A * DE; T[3]  = x;
    T[4] = B * C; T[5] = A * E + B * DF; T[6]  = B * DE - A * F; T[7]  = y;
    T[8] = -D;    T[9] = C * F;          T[10] = C * E;          T[11] = z;
}

// ---- launchers (one per operator; each returns 0 or the launch error, named) ------------------
int fbpr_launch_projection(const ProjArgs& a, int count, cudaStream_t st, long long* launches);
int fbpr_launch_features(const FeatArgs& a, int count, cudaStream_t st, long long* launches);
size_t fbpr_feat_ring_smem(const FeatArgs& a);
int fbpr_launch_voxel(const VoxSeg* d_segs, int nsegs, int max_n, int tiles_cap, cudaStream_t st, long long* launches);
int fbpr_voxel_tile();
int fbpr_launch_grid_build(const GridSeg* d_segs, int nsegs, int max_n, int cells_cap, cudaStream_t st, long long* launches);
int fbpr_launch_knn5(const GridSeg* d_seg, const float* d_q, int nq, int* d_idx, float* d_d2, int rad0, cudaStream_t st, long long* launches);
int fbpr_launch_lm(const LmArgs& args, int count, int cluster_size, int grid_blocks, cudaStream_t st, long long* launches);
int fbpr_lm_grid_blocks(int device);
int fbpr_knn_cache_slots();
int fbpr_launch_transform_update(FrameMeta* meta, int first, int count, float rot_tol, float z_tol, cudaStream_t st, long long* launches);
int fbpr_launch_keyframe_transform(const float* d_poses6, int K, const float4* d_in, const int* d_off, float4* d_out, int* d_n_out,
                                   const float* d_last_xyz, float radius, const float* d_check_xyz, int max_pts, int* d_outoff, float* d_T,
                                   cudaStream_t st, long long* launches);
int fbpr_launch_keyframe_select(const KfStoreView& store, const KfSelect& sel, const VoxSeg* d_poseSeg, int poseTilesCap,
                                double timeLast, float radius, float density, int loopClosure, int keyframeSize,
                                float4* d_outCorner, float4* d_outSurf, int kfCap, int* d_kfCount, int* d_truncated, cudaStream_t st, long long* launches);
int fbpr_launch_crop_box(const float4* d_in, int n, const float* d_pose12, float4* d_out, int cap, int* d_n_out, int* d_truncated, int* d_tile,
                         cudaStream_t st, long long* launches);
int fbpr_launch_crop_box_batched(const float4* d_gc, int nC, const float4* d_gs, int nS, FrameMeta* meta, int first, int count,
                                 float4* d_mapCorner, int capC, float4* d_mapSurf, int capS, int* d_tile, int tilesPer, cudaStream_t st, long long* launches);
int fbpr_launch_pose_decompose(const float* d_pose12, FrameMeta* meta, int slot, cudaStream_t st, long long* launches);
int fbpr_launch_pose_compose(const FrameMeta* meta, int slot, float* d_pose12, cudaStream_t st, long long* launches);
int fbpr_launch_pc2_to_raw(const unsigned char* d_src, int n, const fbpr_pc2_layout& L, fbpr_raw_point* d_dst, cudaStream_t st, long long* launches);
int fbpr_launch_xyzi_repack(const float4* d_in, int n, float4* d_out, int to32, cudaStream_t st, long long* launches);
int fbpr_launch_stage_scatter(const ScatterTable& t, cudaStream_t st, long long* launches);

#define FBPR_CUDA_OK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return fbpr_fail(e_, #expr, __FILE__, __LINE__); } while (0)
int fbpr_fail(cudaError_t e, const char* what, const char* file, int line);
int fbpr_fail_msg(const char* msg);
int fbpr_launch_ok(const char* what);        // cudaGetLastError() after a group of launches: 0, or the error with the operator's name
