// mapops.cu -- local-map providers around the registration path.
//
//   keyframe transform + concat : extractCloud / transformPointCloud (mapOptmization.h:909-944, :405-425)
//   CropBox                     : the fork's registration() local map (mapOptmization.h:284-304),
//                                 pcl::CropBox = inclusive AABB test, input order preserved
//   pose decompose / compose    : pcl::getTranslationAndEulerAngles (:309-310), pcl::getTransformation (:326)
// All are HBM-bound streaming kernels (16 B in, 16 B out per kept point).
#include "internal.cuh"

namespace {

constexpr int TPB = 256;
constexpr int TILE = 2048;
constexpr int IPT = TILE / TPB;

// one CTA: keep flags (distance re-check, :924), output offsets, one 3x4 transform per keyframe
// check_xyz (may be null) = the positions the re-check looks at when they differ from the keyframes' own poses: after
// extractNearby they are the VoxelGrid-averaged key poses, whose averaged intensity picks the keyframe (:927)
__global__ void kf_prepare(const float* poses6, int K, const int* off, const float* last_xyz, float radius, const float* check_xyz,
                           int* outoff, float* T, int* n_out) {
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        const float* p = poses6 + 6 * i;
        get_transformation(p[3], p[4], p[5], p[0], p[1], p[2], T + 12 * i);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < K; i++) {
            const float* p = check_xyz ? check_xyz + 3 * i - 3 : poses6 + 6 * i;       // p[3..5] = position
            float ddx = p[3] - last_xyz[0], ddy = p[4] - last_xyz[1], ddz = p[5] - last_xyz[2];
            bool keep = !(sqrtf(ddx * ddx + ddy * ddy + ddz * ddz) > radius);
            outoff[i] = keep ? run : -1;
            if (keep) run += off[i + 1] - off[i];
        }
        outoff[K] = run;
        *n_out = run;
    }
}

__global__ void __launch_bounds__(TPB) kf_transform(int K, const float4* in, const int* off, const int* outoff, const float* T, float4* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = 0, hi = K;                      // keyframe k with off[k] <= i < off[k+1]
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (off[mid] <= i) lo = mid; else hi = mid; }
    int base = outoff[lo];
    if (base < 0) return;
    const float* t = T + 12 * lo;
    float4 p = in[i], q;
    q.x = t[0] * p.x + t[1] * p.y + t[2] * p.z + t[3];
    q.y = t[4] * p.x + t[5] * p.y + t[6] * p.z + t[7];
    q.z = t[8] * p.x + t[9] * p.y + t[10] * p.z + t[11];
    q.w = p.w;
    out[base + (i - off[lo])] = q;
}

__device__ inline bool in_box(float4 p, const float* pose12) {
    const float mnx = -30.0f + pose12[3], mxx = 30.0f + pose12[3];
    const float mny = -30.0f + pose12[7], mxy = 30.0f + pose12[7];
    const float mnz = -10.0f + pose12[11], mxz = 10.0f + pose12[11];
    return !(p.x < mnx || p.y < mny || p.z < mnz || p.x > mxx || p.y > mxy || p.z > mxz);
}

__global__ void __launch_bounds__(TPB) crop_count(const float4* in, int n, const float* pose12, int* tile) {
    int base = blockIdx.x * TILE, cnt = 0;
    for (int k = 0; k < IPT; k++) { int i = base + k * TPB + threadIdx.x; if (i < n && in_box(in[i], pose12)) cnt++; }
    __shared__ int ws[TPB / 32];
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int k = 0; k < TPB / 32; k++) t += ws[k]; tile[blockIdx.x] = t; }
}

__global__ void __launch_bounds__(1024) crop_scan(int* tile, int ntiles, int* n_out, int cap, int* truncated) {
    __shared__ int ws[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b = 0; b < ntiles; b += 1024) {
        int e = b + threadIdx.x;
        int v = e < ntiles ? tile[e] : 0;
        int incl = v, l = threadIdx.x & 31, w = threadIdx.x >> 5;
        for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, o); if (l >= o) incl += u; }
        if (l == 31) ws[w] = incl;
        __syncthreads();
        if (w == 0) {
            int a = ws[l], ia = a;
            for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, ia, o); if (l >= o) ia += u; }
            ws[l] = ia - a;
        }
        __syncthreads();
        int excl = carry + ws[w] + incl - v;
        if (e < ntiles) tile[e] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {                              // points beyond the slot's map capacity are dropped -- and reported:
        *n_out = min(carry, cap);                        // the reference keeps every cropped point, so the frame's flags say so
        if (carry > cap && truncated) atomicOr(truncated, 1);
    }
}

__global__ void __launch_bounds__(TPB) crop_emit(const float4* in, int n, const float* pose12, const int* tile, float4* out, int cap) {
    // blocked layout (thread t owns IPT consecutive points) keeps the input order
    int start = blockIdx.x * TILE + threadIdx.x * IPT;
    float4 p[IPT]; int flags = 0, cnt = 0;
    for (int k = 0; k < IPT; k++) { int i = start + k; if (i < n) { p[k] = in[i]; if (in_box(p[k], pose12)) { flags |= 1 << k; cnt++; } } }
    __shared__ int ws[TPB / 32];
    int incl = cnt, l = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, o); if (l >= o) incl += u; }
    if (l == 31) ws[w] = incl;
    __syncthreads();
    int woff = 0; for (int q = 0; q < w; q++) woff += ws[q];
    int slot = tile[blockIdx.x] + woff + incl - cnt;
    for (int k = 0; k < IPT; k++) if (flags & (1 << k)) { if (slot < cap) out[slot] = p[k]; slot++; }
}

// ---- batched CropBox: every slot of a range crops the SAME resident global maps around its own pose (blockIdx.y = frame,
// blockIdx.z = kind).  The fork's live registration() (mapOptmization.h:284-304) for many independent frames at once: only the
// sweeps cross PCIe, the local maps are cut on the device.  Same inclusive box and order-preserving compaction as above.
__device__ inline bool in_box_t(float4 p, float tx, float ty, float tz) {
    const float mnx = -30.0f + tx, mxx = 30.0f + tx, mny = -30.0f + ty, mxy = 30.0f + ty, mnz = -10.0f + tz, mxz = 10.0f + tz;
    return !(p.x < mnx || p.y < mny || p.z < mnz || p.x > mxx || p.y > mxy || p.z > mxz);
}
__global__ void __launch_bounds__(TPB) crop_count_b(const float4* gc, int nC, const float4* gs, int nS, const FrameMeta* meta, int first, int* tile, int tilesPer) {
    const int kind = blockIdx.z, n = kind ? nS : nC;
    const int base = blockIdx.x * TILE;
    if (base >= n) return;
    const float4* in = kind ? gs : gc;
    const FrameMeta& M = meta[first + blockIdx.y];
    const float tx = M.pose[3], ty = M.pose[4], tz = M.pose[5];
    int cnt = 0;
    for (int k = 0; k < IPT; k++) { int i = base + k * TPB + threadIdx.x; if (i < n && in_box_t(in[i], tx, ty, tz)) cnt++; }
    __shared__ int ws[TPB / 32];
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int k = 0; k < TPB / 32; k++) t += ws[k]; tile[((size_t)blockIdx.y * 2 + kind) * tilesPer + blockIdx.x] = t; }
}
__global__ void __launch_bounds__(1024) crop_scan_b(int* tile, int tilesPer, int nC, int nS, FrameMeta* meta, int first, int capC, int capS) {
    const int kind = blockIdx.y, n = kind ? nS : nC, ntiles = (n + TILE - 1) / TILE, cap = kind ? capS : capC;
    int* t = tile + ((size_t)blockIdx.x * 2 + kind) * tilesPer;
    FrameMeta& M = meta[first + blockIdx.x];
    __shared__ int ws[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b = 0; b < ntiles; b += 1024) {
        int e = b + threadIdx.x;
        int v = e < ntiles ? t[e] : 0;
        int incl = v, l = threadIdx.x & 31, w = threadIdx.x >> 5;
        for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, o); if (l >= o) incl += u; }
        if (l == 31) ws[w] = incl;
        __syncthreads();
        if (w == 0) {
            int a = ws[l], ia = a;
            for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, ia, o); if (l >= o) ia += u; }
            ws[l] = ia - a;
        }
        __syncthreads();
        int excl = carry + ws[w] + incl - v;
        if (e < ntiles) t[e] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (kind) M.n_map_surf = min(carry, cap); else M.n_map_corner = min(carry, cap);
        if (carry > cap) atomicOr(&M.mapTruncated, 1);
    }
}
__global__ void __launch_bounds__(TPB) crop_emit_b(const float4* gc, int nC, const float4* gs, int nS, const FrameMeta* meta, int first, const int* tile, int tilesPer,
                                                   float4* outC, int capC, float4* outS, int capS) {
    const int kind = blockIdx.z, n = kind ? nS : nC, cap = kind ? capS : capC;
    if (blockIdx.x * TILE >= n) return;
    const float4* in = kind ? gs : gc;
    const int slot = first + blockIdx.y;
    float4* out = kind ? outS + (size_t)slot * capS : outC + (size_t)slot * capC;
    const FrameMeta& M = meta[slot];
    const float tx = M.pose[3], ty = M.pose[4], tz = M.pose[5];
    int start = blockIdx.x * TILE + threadIdx.x * IPT;
    float4 p[IPT]; int flags = 0, cnt = 0;
    for (int k = 0; k < IPT; k++) { int i = start + k; if (i < n) { p[k] = in[i]; if (in_box_t(p[k], tx, ty, tz)) { flags |= 1 << k; cnt++; } } }
    __shared__ int ws[TPB / 32];
    int incl = cnt, l = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, o); if (l >= o) incl += u; }
    if (l == 31) ws[w] = incl;
    __syncthreads();
    int woff = 0; for (int q = 0; q < w; q++) woff += ws[q];
    int at = tile[((size_t)blockIdx.y * 2 + kind) * tilesPer + blockIdx.x] + woff + incl - cnt;
    for (int k = 0; k < IPT; k++) if (flags & (1 << k)) { if (at < cap) out[at] = p[k]; at++; }
}

__global__ void pose_decompose(const float* T, FrameMeta* meta, int slot) {
    if (threadIdx.x != 0) return;
    FrameMeta& M = meta[slot];
    M.pose[3] = T[3]; M.pose[4] = T[7]; M.pose[5] = T[11];
    M.pose[0] = (float)atan2((double)T[9], (double)T[10]);
    M.pose[1] = (float)asin((double)(-T[8]));
    M.pose[2] = (float)atan2((double)T[4], (double)T[0]);
}
__global__ void pose_compose(const FrameMeta* meta, int slot, float* T) {
    if (threadIdx.x != 0) return;
    const FrameMeta& M = meta[slot];
    float t[12];
    get_transformation(M.pose[3], M.pose[4], M.pose[5], M.pose[0], M.pose[1], M.pose[2], t);
    for (int k = 0; k < 12; k++) T[k] = t[k];
}

// ---- wire formats -----------------------------------------------------------------------------
// sensor_msgs/PointCloud2 records of any point_step / field offsets (the Velodyne driver's 22-byte x,y,z,intensity,ring,time
// or pcl::toROSMsg's padded 32-byte PointXYZIRT, imageProjection.cpp:8-21, :252) -> the packed 24-byte record of the projection
__device__ __forceinline__ unsigned load_u32_unaligned(const unsigned char* p) {
    return (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24);
}
__global__ void __launch_bounds__(TPB) pc2_to_raw(const unsigned char* __restrict__ src, int n, fbpr_pc2_layout L, fbpr_raw_point* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned char* r = src + (size_t)i * L.point_step;
    fbpr_raw_point o;
    o.x = __uint_as_float(load_u32_unaligned(r + L.off_x));
    o.y = __uint_as_float(load_u32_unaligned(r + L.off_y));
    o.z = __uint_as_float(load_u32_unaligned(r + L.off_z));
    o.intensity = L.off_intensity >= 0 ? __uint_as_float(load_u32_unaligned(r + L.off_intensity)) : 0.f;
    int ring = 0;
    if (L.ring_bytes == 1) ring = r[L.off_ring];
    else if (L.ring_bytes == 2) ring = (int)r[L.off_ring] | ((int)r[L.off_ring + 1] << 8);
    else if (L.ring_bytes == 4) ring = (int)load_u32_unaligned(r + L.off_ring);
    o.ring = ring;
    o.time = L.off_time >= 0 ? __uint_as_float(load_u32_unaligned(r + L.off_time)) : 0.f;
    dst[i] = o;
}
// float4 XYZI <-> the 32-byte pcl::PointXYZI of pcl::toROSMsg / fromROSMsg (x,y,z,pad | intensity,pad,pad,pad; utility.h:55, :255-264)
__global__ void __launch_bounds__(TPB) xyzi16_to_32(const float4* __restrict__ in, int n, float4* __restrict__ out2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = in[i];
    out2[2 * i] = make_float4(p.x, p.y, p.z, 1.0f);          // PCL initialises the padding word data[3] to 1
    out2[2 * i + 1] = make_float4(p.w, 0.f, 0.f, 0.f);
}
__global__ void __launch_bounds__(TPB) xyzi32_to_16(const float4* __restrict__ in2, int n, float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = in2[2 * i], b = in2[2 * i + 1];
    out[i] = make_float4(a.x, a.y, a.z, b.x);
}

// one CTA column per piece (blockIdx.y); 16-byte copies when source and destination allow, 4-byte otherwise.  Pieces of kind
// WIRE22 / XYZ12 are repacked on the way: the bytes that crossed PCIe are the wire format's, the slot gets the kernels' layout.
__global__ void __launch_bounds__(TPB) stage_scatter(const __grid_constant__ ScatterTable t) {
    const ScatterPiece pc = t.p[blockIdx.y];
    const int kind = (int)(pc.bytes >> 60);
    const unsigned long long bytes = pc.bytes & 0x0fffffffffffffffull;
    const unsigned char* src = t.stage + pc.src_off;
    unsigned char* dst = reinterpret_cast<unsigned char*>(pc.dst);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    if (kind == FBPR_PIECE_WIRE22) {
        // the Velodyne driver's PointXYZIRT record: x, y, z, intensity (f32), ring (u16), time (f32) = 22 bytes, any alignment.
        // A tile of 256 records (5632 bytes) is staged in shared memory with aligned word loads, then one thread per record
        constexpr int REC = 22, TILE_BYTES = TPB * REC;
        __shared__ unsigned sw[TILE_BYTES / 4 + 2];
        const unsigned char* sb = reinterpret_cast<const unsigned char*>(sw);
        const long long n = (long long)(bytes / REC);
        fbpr_raw_point* out = reinterpret_cast<fbpr_raw_point*>(dst);
        for (long long base = (long long)blockIdx.x * TPB; base < n; base += (long long)gridDim.x * TPB) {
            const unsigned char* g = src + base * REC;
            const unsigned shift = (unsigned)((size_t)g & 3);
            const unsigned* gw = reinterpret_cast<const unsigned*>(g - shift);
            const long long left = (n - base) * REC;
            const int nb = (int)(left < TILE_BYTES ? left : TILE_BYTES);
            const int nw = (int)((shift + nb + 3) >> 2);
            for (int w = threadIdx.x; w < nw; w += TPB) sw[w] = gw[w];
            __syncthreads();
            if ((long long)threadIdx.x < n - base) {
                const unsigned char* r = sb + shift + threadIdx.x * REC;
                fbpr_raw_point o;
                o.x = __uint_as_float(load_u32_unaligned(r)); o.y = __uint_as_float(load_u32_unaligned(r + 4));
                o.z = __uint_as_float(load_u32_unaligned(r + 8)); o.intensity = __uint_as_float(load_u32_unaligned(r + 12));
                o.ring = (int)r[16] | ((int)r[17] << 8);
                o.time = __uint_as_float(load_u32_unaligned(r + 18));
                out[base + threadIdx.x] = o;
            }
            __syncthreads();
        }
        return;
    }
    if (kind == FBPR_PIECE_XYZ12) {
        const size_t n = (size_t)(bytes / 12);
        const float* s3 = reinterpret_cast<const float*>(src); float4* d4 = reinterpret_cast<float4*>(dst);
        for (size_t i = tid; i < n; i += nth) d4[i] = make_float4(s3[3 * i], s3[3 * i + 1], s3[3 * i + 2], 0.f);
        return;
    }
    if ((((size_t)src | (size_t)dst) & 15) == 0) {
        const size_t n16 = bytes >> 4;
        const uint4* s4 = reinterpret_cast<const uint4*>(src); uint4* d4 = reinterpret_cast<uint4*>(dst);
        for (size_t i = tid; i < n16; i += nth) d4[i] = s4[i];
        const size_t done = n16 << 4;
        for (size_t i = done + 4 * tid; i < bytes; i += 4 * nth) *reinterpret_cast<unsigned*>(dst + i) = *reinterpret_cast<const unsigned*>(src + i);
    } else {
        const size_t n4 = bytes >> 2;
        const unsigned* s1 = reinterpret_cast<const unsigned*>(src); unsigned* d1 = reinterpret_cast<unsigned*>(dst);
        for (size_t i = tid; i < n4; i += nth) d1[i] = s1[i];
    }
}

}  // namespace

int fbpr_launch_stage_scatter(const ScatterTable& t, cudaStream_t st, long long* launches) {
    if (t.n <= 0) return 0;
    stage_scatter<<<dim3(24, t.n), TPB, 0, st>>>(t);
    if (launches) *launches += 1;
    return fbpr_launch_ok("stage_scatter");
}

int fbpr_launch_pc2_to_raw(const unsigned char* d_src, int n, const fbpr_pc2_layout& L, fbpr_raw_point* d_dst, cudaStream_t st, long long* launches) {
    if (n <= 0) return 0;
    pc2_to_raw<<<(n + TPB - 1) / TPB, TPB, 0, st>>>(d_src, n, L, d_dst);
    if (launches) *launches += 1;
    return fbpr_launch_ok("pc2_to_raw");
}
int fbpr_launch_xyzi_repack(const float4* d_in, int n, float4* d_out, int to32, cudaStream_t st, long long* launches) {
    if (n <= 0) return 0;
    if (to32) xyzi16_to_32<<<(n + TPB - 1) / TPB, TPB, 0, st>>>(d_in, n, d_out);
    else xyzi32_to_16<<<(n + TPB - 1) / TPB, TPB, 0, st>>>(d_in, n, d_out);
    if (launches) *launches += 1;
    return fbpr_launch_ok("xyzi repack");
}

int fbpr_launch_keyframe_transform(const float* d_poses6, int K, const float4* d_in, const int* d_off, float4* d_out, int* d_n_out,
                                    const float* d_last_xyz, float radius, const float* d_check_xyz, int max_pts, int* d_outoff, float* d_T,
                                    cudaStream_t st, long long* launches) {
    kf_prepare<<<1, 256, 0, st>>>(d_poses6, K, d_off, d_last_xyz, radius, d_check_xyz, d_outoff, d_T, d_n_out);
    if (max_pts > 0 && K > 0) kf_transform<<<(max_pts + TPB - 1) / TPB, TPB, 0, st>>>(K, d_in, d_off, d_outoff, d_T, d_out, max_pts);
    if (launches) *launches += (max_pts > 0 && K > 0) ? 2 : 1;
    return fbpr_launch_ok("kf_prepare / kf_transform");
}

int fbpr_launch_crop_box(const float4* d_in, int n, const float* d_pose12, float4* d_out, int cap, int* d_n_out, int* d_truncated, int* d_tile,
                         cudaStream_t st, long long* launches) {
    int tiles = (n + TILE - 1) / TILE;
    if (tiles > 0) crop_count<<<tiles, TPB, 0, st>>>(d_in, n, d_pose12, d_tile);
    crop_scan<<<1, 1024, 0, st>>>(d_tile, tiles, d_n_out, cap, d_truncated);
    if (tiles > 0) crop_emit<<<tiles, TPB, 0, st>>>(d_in, n, d_pose12, d_tile, d_out, cap);
    if (launches) *launches += tiles > 0 ? 3 : 1;
    return fbpr_launch_ok("CropBox (crop_count / crop_scan / crop_emit)");
}

int fbpr_launch_crop_box_batched(const float4* d_gc, int nC, const float4* d_gs, int nS, FrameMeta* meta, int first, int count,
                                 float4* d_mapCorner, int capC, float4* d_mapSurf, int capS, int* d_tile, int tilesPer, cudaStream_t st, long long* launches) {
    if (count <= 0) return 0;
    const int need = nC > nS ? nC : nS;
    const int tiles = (need + TILE - 1) / TILE;
    if (tiles > tilesPer) return fbpr_fail_msg("CropBox scratch too small");
    if (tiles > 0) crop_count_b<<<dim3(tiles, count, 2), TPB, 0, st>>>(d_gc, nC, d_gs, nS, meta, first, d_tile, tilesPer);
    crop_scan_b<<<dim3(count, 2), 1024, 0, st>>>(d_tile, tilesPer, nC, nS, meta, first, capC, capS);
    if (tiles > 0) crop_emit_b<<<dim3(tiles, count, 2), TPB, 0, st>>>(d_gc, nC, d_gs, nS, meta, first, d_tile, tilesPer, d_mapCorner, capC, d_mapSurf, capS);
    if (launches) *launches += tiles > 0 ? 3 : 1;
    return fbpr_launch_ok("batched CropBox (crop_*_b)");
}

int fbpr_launch_pose_decompose(const float* d_pose12, FrameMeta* meta, int slot, cudaStream_t st, long long* launches) {
    pose_decompose<<<1, 32, 0, st>>>(d_pose12, meta, slot);
    if (launches) *launches += 1;
    return fbpr_launch_ok("pose_decompose");
}
int fbpr_launch_pose_compose(const FrameMeta* meta, int slot, float* d_pose12, cudaStream_t st, long long* launches) {
    pose_compose<<<1, 32, 0, st>>>(meta, slot, d_pose12);
    if (launches) *launches += 1;
    return fbpr_launch_ok("pose_compose");
}
