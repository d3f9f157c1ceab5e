// keyframes.cu -- extractSurroundingKeyFrames end to end on the device, over a RESIDENT keyframe store.
//
// Replaces (mapOptmization.h): extractSurroundingKeyFrames (:964-978) = extractNearby (:872-907) or, with
// loopClosureEnableFlag, extractForLoopClosure (:857-870), followed by extractCloud (:909-955).  The store
// (cloudKeyPoses3D / 6D, cornerCloudKeyFrames / surfCloudKeyFrames, :84-88) lives in HBM and is appended to by
// fbpr_keyframe_push, so a frame of a live sequence needs no host round trip to build its local map:
//
//   kfs_hits      radiusSearch(cloudKeyPoses3D->back(), r): d^2 < (float)(r*r) over all key poses, L2_Simple order of operations
//   kfs_sort      hits ascending by (d^2, index) -- FLANN's sorted result set -- one CTA, bitonic network
//   voxel.cu      VoxelGrid(surroundingKeyframeDensity) of the hit poses, xyz AND intensity averaged (:887-888)
//   kfs_list      + the key poses of the last 10 s, newest first (:897-904); per entry the keyframe it names ((int)intensity, :927),
//                 the distance re-check at the entry's own position (:924), its transform and its place in the concatenation
//   kfs_transform transformPointCloud (:405-425) of the named keyframes' clouds into the concatenation, list order (:939-944)
//   voxel.cu      the two VoxelGrids of extractCloud (:948-954) into the slot's local map
// All sizes stay in device memory.  HBM-bound streaming (16 B in, 16 B out per keyframe point); the selection itself is
// a few thousand poses and is latency bound.
#include "internal.cuh"

namespace {

constexpr int TPB = 256;

__global__ void __launch_bounds__(TPB) kfs_hits(KfStoreView s, float r2, KfSelect q) {
    const float* last = s.pose6 + 6 * (size_t)(s.n - 1);
    const float qx = last[3], qy = last[4], qz = last[5];
    for (int i = blockIdx.x * TPB + threadIdx.x; i < s.n; i += gridDim.x * TPB) {
        const float* p = s.pose6 + 6 * (size_t)i;
        const float dx = qx - p[3], dy = qy - p[4], dz = qz - p[5];
        float d = dx * dx; d += dy * dy; d += dz * dz;
        if (d < r2) {                                          // flann::RadiusResultSet: strictly inside
            const int at = atomicAdd(q.counters, 1);           // any order: sorted next
            q.keys[at] = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)i;
        }
    }
}

// one CTA: sort the hit keys (d^2 >= 0, so the f32 bit pattern orders like the value; low word = index breaks ties) and
// emit surroundingKeyPoses in that order
__global__ void __launch_bounds__(1024) kfs_sort(KfStoreView s, KfSelect q) {
    __shared__ unsigned long long sk[4096];
    const int nh = q.counters[0];
    int npad = 1; while (npad < nh) npad <<= 1;
    unsigned long long* a = npad <= 4096 ? sk : q.keys;
    for (int i = threadIdx.x; i < npad; i += 1024) {
        if (npad <= 4096) sk[i] = i < nh ? q.keys[i] : ~0ull;
        else if (i >= nh) q.keys[i] = ~0ull;
    }
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npad; i += 1024) {
                const int p = i ^ j;
                if (p > i) {
                    const unsigned long long x = a[i], y = a[p];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { a[i] = y; a[p] = x; }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < nh; i += 1024) {
        const int idx = (int)(unsigned)(a[i] & 0xffffffffu);
        const float* p = s.pose6 + 6 * (size_t)idx;
        q.hitPts[i] = make_float4(p[3], p[4], p[5], (float)idx);   // cloudKeyPoses3D[idx]: intensity = index (:1689)
    }
}

__device__ inline int block_excl_scan(int v, int* ws, int& total) {      // 1024 threads; returns the exclusive prefix, total = sum
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, o); if (l >= o) incl += u; }
    __syncthreads();
    if (l == 31) ws[w] = incl;
    __syncthreads();
    if (w == 0) {
        int a = ws[l], ia = a;
        for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, ia, o); if (l >= o) ia += u; }
        ws[l] = ia - a;
        if (l == 31) ws[32] = ia;
    }
    __syncthreads();
    total = ws[32];
    return ws[w] + incl - v;
}

// one CTA: finish cloudToExtract and lay out the concatenation
__global__ void __launch_bounds__(1024) kfs_list(KfStoreView s, KfSelect q, double timeLast, float radius, int loopClosure, int keyframeSize,
                                                  int kfCap, int* kfCount, int* truncated) {
    __shared__ int ws[33];
    __shared__ int s_fail;
    const int n = s.n;
    int K;
    if (loopClosure) {
        // extractForLoopClosure (:857-870): poses from the newest backwards while size() <= surroundingKeyframeSize
        K = keyframeSize + 1; if (K > n) K = n; if (K < 0) K = 0;
        for (int k = threadIdx.x; k < K; k += 1024) {
            const int i = n - 1 - k; const float* p = s.pose6 + 6 * (size_t)i;
            q.list[k] = make_float4(p[3], p[4], p[5], (float)i);
        }
    } else {
        // :897-904: newest first until the first key pose that is 10 s old or older
        if (threadIdx.x == 0) s_fail = -1;
        __syncthreads();
        int f = -1;
        for (int i = threadIdx.x; i < n; i += 1024) if (!(timeLast - s.time[i] < 10.0)) f = i;     // ascending i: the last assignment is this thread's max
        for (int o = 16; o; o >>= 1) f = max(f, __shfl_xor_sync(0xffffffffu, f, o));
        if ((threadIdx.x & 31) == 0 && f >= 0) atomicMax(&s_fail, f);
        __syncthreads();
        const int nDS = q.counters[1], R = n - 1 - s_fail;
        for (int k = threadIdx.x; k < R; k += 1024) {
            const int i = n - 1 - k; const float* p = s.pose6 + 6 * (size_t)i;
            q.list[nDS + k] = make_float4(p[3], p[4], p[5], (float)i);
        }
        K = nDS + R;
    }
    __syncthreads();
    const float* lk = s.pose6 + 6 * (size_t)(n - 1) + 3;         // cloudKeyPoses3D->back()
    int carry[2] = { 0, 0 };
    for (int b = 0; b < K; b += 1024) {
        const int k = b + threadIdx.x;
        int idx = -1, len[2] = { 0, 0 };
        if (k < K) {
            const float4 c = q.list[k];
            const float d = sqrtf((c.x - lk[0]) * (c.x - lk[0]) + (c.y - lk[1]) * (c.y - lk[1]) + (c.z - lk[2]) * (c.z - lk[2]));   // pointDistance, :924
            const int cand = (int)c.w;                               // thisKeyInd = (int)intensity (:927): the AVERAGED index, truncated
            if (!(d > radius) && cand >= 0 && cand < n) {
                idx = cand;
                len[0] = s.off[0][idx + 1] - s.off[0][idx]; len[1] = s.off[1][idx + 1] - s.off[1][idx];
                const float* p = s.pose6 + 6 * (size_t)idx;
                get_transformation(p[3], p[4], p[5], p[0], p[1], p[2], q.T + 12 * (size_t)k);
            }
            q.selIdx[k] = idx;
        }
        for (int kind = 0; kind < 2; kind++) {
            int total;
            const int ex = block_excl_scan(len[kind], ws, total);
            if (k < K) q.outoff[kind][k] = carry[kind] + ex;
            carry[kind] += total;
        }
    }
    if (threadIdx.x == 0) {
        q.counters[2] = K;
        for (int kind = 0; kind < 2; kind++) {
            q.outoff[kind][K] = carry[kind];
            kfCount[kind] = min(carry[kind], kfCap);              // the concatenation is cut at max_keyframe_points -- and reported
            if (carry[kind] > kfCap && truncated) atomicOr(truncated, 1);
        }
    }
}

// transformPointCloud (:405-425) of every kept entry's clouds, straight into the concatenation (blockIdx.y = kind)
__global__ void __launch_bounds__(TPB) kfs_transform(KfStoreView s, KfSelect q, float4* outCorner, float4* outSurf, const int* kfCount) {
    const int kind = blockIdx.y;
    const int K = q.counters[2], total = kfCount[kind];
    const int* outoff = q.outoff[kind];
    const int* off = s.off[kind];
    const float4* pool = s.pool[kind];
    float4* out = kind ? outSurf : outCorner;
    for (int o = blockIdx.x * TPB + threadIdx.x; o < total; o += gridDim.x * TPB) {
        int lo = 0, hi = K;                                      // last entry whose offset is <= o (dropped entries have length 0)
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (outoff[mid] <= o) lo = mid; else hi = mid; }
        const int idx = q.selIdx[lo];
        const float* t = q.T + 12 * (size_t)lo;
        const float4 p = pool[off[idx] + (o - outoff[lo])];
        float4 r;
        r.x = t[0] * p.x + t[1] * p.y + t[2] * p.z + t[3];
        r.y = t[4] * p.x + t[5] * p.y + t[6] * p.z + t[7];
        r.z = t[8] * p.x + t[9] * p.y + t[10] * p.z + t[11];
        r.w = p.w;
        out[o] = r;
    }
}

}  // namespace

int fbpr_launch_keyframe_select(const KfStoreView& store, const KfSelect& sel, const VoxSeg* d_poseSeg, int poseTilesCap,
                                double timeLast, float radius, float density, int loopClosure, int keyframeSize,
                                float4* d_outCorner, float4* d_outSurf, int kfCap, int* d_kfCount, int* d_truncated, cudaStream_t st, long long* launches) {
    (void)density;        // the leaf of d_poseSeg
    if (store.n <= 0) return 0;
    cudaError_t e = cudaMemsetAsync(sel.counters, 0, 4 * sizeof(int), st);
    if (e != cudaSuccess) return fbpr_fail(e, "cudaMemsetAsync(keyframe selection counters)", __FILE__, __LINE__);
    if (!loopClosure) {
        const float r2 = (float)((double)radius * (double)radius);
        int g = (store.n + TPB - 1) / TPB; if (g > 592) g = 592;
        kfs_hits<<<g, TPB, 0, st>>>(store, r2, sel);
        kfs_sort<<<1, 1024, 0, st>>>(store, sel);
        if (launches) *launches += 2;
        int rc = fbpr_launch_voxel(d_poseSeg, 1, store.n, poseTilesCap, st, launches);
        if (rc) return rc;
    }
    kfs_list<<<1, 1024, 0, st>>>(store, sel, timeLast, radius, loopClosure, keyframeSize, kfCap, d_kfCount, d_truncated);
    kfs_transform<<<dim3(592, 2), TPB, 0, st>>>(store, sel, d_outCorner, d_outSurf, d_kfCount);
    if (launches) *launches += 2;
    return fbpr_launch_ok("keyframe selection (kfs_*)");
}
