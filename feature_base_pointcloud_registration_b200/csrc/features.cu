// features.cu -- LOAM feature front-end: smoothness, occlusion masks, per-ring corner / surface
// selection and the per-ring VoxelGrid of surface points.
//
// Replaces FeatureExtraction::calculateSmoothness (featureExtraction.h:109-131),
// markOccludedPoints (:134-176) and extractFeatures (:178-294).
//
// Kernel 1 (feat_smooth): coalesced +-5 stencil on the flat ring-major range array; the
//   occlusion marks are evaluated in GATHER form (each index looks at the <= 13 source
//   positions that could mark it), so there are no scattered writes and no ordering issue.
// Kernel 2 (feat_ring): one CTA per ring, everything in shared memory:
//   - the 6 segment sorts (featureExtraction.h:203) run as ONE bitonic network over
//     (curvature bits, index) keys -- a total order, ties broken by point index;
//   - the corner loop (:208-242) is walked by one thread but only over keys > edgeThreshold;
//   - the "flat" loop (:245-276) is a lexicographically-first maximal independent set in
//     sorted order; it is evaluated exactly in parallel rounds (a candidate is decided once all
//     lower-ranked candidates that could suppress it are decided), reproducing the sequential
//     result including the marks that leak into the next segment;
//   - surface = every in-segment index that is not a corner (:279-284), stream-compacted;
//   - the ring's pcl::VoxelGrid (:287-292): bbox, keys, bitonic sort of (key, seq), runs,
//     sequential in-order centroids.
// Kernel 3 (feat_gather): ring-order concatenation of corners and down-sampled surface points.
// Quirks reproduced on purpose: slot `ep` is never sorted and is
// visited first by the corner loop and last by the flat loop; cloudSmoothness slots outside
// [5, n-5) hold {0.0f, ind 0}; the 21st corner breaks before marking.
#include "internal.cuh"


namespace {

constexpr int TPB1 = 256;
constexpr int RING_TPB = 512;
constexpr int CORNERS_PER_RING = FBPR_SEGS * FBPR_CORNERS_PER_SEG;

// occlusion rule of source position i (featureExtraction.h:142-166): bit 0 = "mark i-5..i", bit 1 = "mark i+1..i+6"
__device__ __forceinline__ int occlusion_source(const float* __restrict__ r, const int* __restrict__ col, int i, int n) {
    if (i < 5 || i >= n - 6) return 0;
    if (abs(col[i + 1] - col[i]) >= 10) return 0;
    const float d1 = r[i], d2 = r[i + 1];
    if ((double)(d1 - d2) > 0.3) return 1;
    if ((double)(d2 - d1) > 0.3) return 2;
    return 0;
}

// One tile of TPB1 consecutive points per CTA.  Every source position is evaluated once (plus a 32-point halo on each side),
// its two flags go into per-warp ballot words, and a point's mark is a bit-window test: any "mark back" source in j..j+5 or
// any "mark forward" source in j-6..j-1.  No scattered writes, no ordering issue.
__global__ void __launch_bounds__(TPB1) feat_smooth(FeatArgs a) {
    const int slot = a.first + blockIdx.y;
    const int n = a.meta[slot].n_valid;
    const int base = blockIdx.x * TPB1;
    if (base >= n) return;
    const float* r = a.range + (size_t)slot * a.P;
    const int* col = a.colInd + (size_t)slot * a.P;
    float* curv = a.curv + (size_t)slot * a.P;
    int* picked = a.picked + (size_t)slot * a.P;
    int* label = a.label + (size_t)slot * a.P;
    constexpr int NW = TPB1 / 32;
    __shared__ unsigned s_back[NW + 2], s_fwd[NW + 2];          // word 0 = halo before the tile, word NW + 1 = halo after it
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j = base + tid;
    {
        const int f = occlusion_source(r, col, j, n);
        const unsigned wb = __ballot_sync(0xffffffffu, f == 1), wf = __ballot_sync(0xffffffffu, f == 2);
        if (lane == 0) { s_back[warp + 1] = wb; s_fwd[warp + 1] = wf; }
        if (warp < 2) {                                           // halos: 32 positions before / after the tile
            const int i = warp == 0 ? base - 32 + lane : base + TPB1 + lane;
            const int fh = occlusion_source(r, col, i, n);
            const unsigned hb = __ballot_sync(0xffffffffu, fh == 1), hf = __ballot_sync(0xffffffffu, fh == 2);
            if (lane == 0) { const int wd = warp == 0 ? 0 : NW + 1; s_back[wd] = hb; s_fwd[wd] = hf; }
        }
    }
    __syncthreads();
    if (j >= n) return;
    float c = 0.f;
    if (j >= 5 && j < n - 5) {
        float d = r[j - 5] + r[j - 4] + r[j - 3] + r[j - 2] + r[j - 1] - r[j] * 10
                + r[j + 1] + r[j + 2] + r[j + 3] + r[j + 4] + r[j + 5];
        c = d * d;
    }
    // "mark back" sources j .. j+5: bits lane .. lane+5 of (next:this); "mark forward" sources j-6 .. j-1: bits lane+26 .. lane+31 of (this:prev)
    const unsigned long long backBits = (((unsigned long long)s_back[warp + 2] << 32) | s_back[warp + 1]) >> lane;
    const unsigned long long fwdBits = (((unsigned long long)s_fwd[warp + 1] << 32) | s_fwd[warp]) >> (lane + 26);
    int pk = ((backBits & 0x3full) | (fwdBits & 0x3full)) ? 1 : 0;
    if (j >= 5 && j < n - 6) {
        float diff1 = fabsf(r[j - 1] - r[j]), diff2 = fabsf(r[j + 1] - r[j]);
        if ((double)diff1 > 0.02 * (double)r[j] && (double)diff2 > 0.02 * (double)r[j]) pk = 1;
    }
    curv[j] = c; picked[j] = pk; label[j] = 0;
}

// ---- block helpers -----------------------------------------------------------------------
// exclusive scan of a 0 / 1 flag over the CTA (every call site scans a flag): ballot + popc inside the warp, and the
// RING_TPB / 32 warp counts are scanned by every warp for itself with shuffles
__device__ inline int block_excl_scan(int v, int* total, int* ws) {
    static_assert(RING_TPB / 32 <= 32, "one lane per warp count");
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned b = __ballot_sync(0xffffffffu, v != 0);
    __syncthreads();                                      // ws may still be read by the previous scan
    if (l == 0) ws[w] = __popc(b);
    __syncthreads();
    const int c = l < RING_TPB / 32 ? ws[l] : 0;
    int incl = c;
    #pragma unroll
    for (int o = 1; o < RING_TPB / 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (l >= o) incl += u; }
    *total = __shfl_sync(0xffffffffu, incl, RING_TPB / 32 - 1);
    const int before = __shfl_sync(0xffffffffu, incl - c, w);
    return before + __popc(b & ((1u << l) - 1u));
}

// bitonic network over `count` keys made of aligned blocks of `blk` (pow2); every block ends ascending.
// Generic shared-memory form (any sizes); the register forms below are used for the common shapes.
template <typename K>
__device__ inline void bitonic_blocks(K* keys, int count, int blk) {
    for (int k = 2; k <= blk; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < count; t += RING_TPB) {
                int p = t ^ j;
                if (p > t) {
                    bool asc = ((t & k) == 0) || (k == blk);
                    K x = keys[t], y = keys[p];
                    if ((x > y) == asc) { keys[t] = y; keys[p] = x; }
                }
            }
            __syncthreads();
        }
    }
}

template <typename K>
__device__ __forceinline__ void key_cswap(K& a, K& b, bool asc) {
    const K x = a, y = b;
    const bool sw = (x > y) == asc;
    a = sw ? y : x; b = sw ? x : y;
}

// One WARP sorts 32 * IPT keys ascending, IPT keys per lane in registers (element e = lane * IPT + i): strides below
// IPT are register-to-register compare-exchanges, the others one shuffle per key.  No shared memory, no barriers.
template <int IPT>
__device__ __forceinline__ void warp_bitonic_sort(unsigned long long (&k)[IPT], int lane) {
    #pragma unroll
    for (int kk = 2; kk <= 32 * IPT; kk <<= 1) {
        #pragma unroll
        for (int j = kk >> 1; j > 0; j >>= 1) {
            if (j >= IPT) {
                const int lj = j / IPT;
                const bool up = ((lane * IPT) & kk) == 0 || kk == 32 * IPT;
                const bool keepMin = ((lane & lj) == 0) == up;
                #pragma unroll
                for (int i = 0; i < IPT; i++) {
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, k[i], lj);
                    const bool less = o < k[i];
                    k[i] = (less == keepMin) ? o : k[i];
                }
            } else {
                #pragma unroll
                for (int i = 0; i < IPT; i++) {
                    if ((i & j) == 0) {
                        const bool up = (((lane * IPT + i) & kk) == 0) || kk == 32 * IPT;
                        key_cswap(k[i], k[i | j], up);
                    }
                }
            }
        }
    }
}

// The whole CTA (RING_TPB threads) sorts RING_TPB * IPT keys of shared memory ascending, IPT keys per thread in
// registers (element e = tid * IPT + i); only the strides that cross warps go through shared memory.
template <int IPT, typename K>
__device__ __forceinline__ void cta_bitonic_sort(K* keys) {
    const int tid = threadIdx.x, lane = tid & 31;
    constexpr int N = RING_TPB * IPT;
    K k[IPT];
    #pragma unroll
    for (int i = 0; i < IPT; i++) k[i] = keys[tid * IPT + i];
    #pragma unroll
    for (int kk = 2; kk <= N; kk <<= 1) {
        #pragma unroll
        for (int j = kk >> 1; j > 0; j >>= 1) {
            if (j >= 32 * IPT) {                     // partner in another warp: exchange through shared memory
                __syncthreads();
                #pragma unroll
                for (int i = 0; i < IPT; i++) keys[tid * IPT + i] = k[i];
                __syncthreads();
                const int tj = j / IPT;
                const bool up = ((tid * IPT) & kk) == 0 || kk == N;
                const bool keepMin = ((tid & tj) == 0) == up;
                #pragma unroll
                for (int i = 0; i < IPT; i++) {
                    const K o = keys[(tid ^ tj) * IPT + i];
                    const bool less = o < k[i];
                    k[i] = (less == keepMin) ? o : k[i];
                }
            } else if (j >= IPT) {
                const int lj = j / IPT;
                const bool up = ((tid * IPT) & kk) == 0 || kk == N;
                const bool keepMin = ((lane & lj) == 0) == up;
                #pragma unroll
                for (int i = 0; i < IPT; i++) {
                    const K o = __shfl_xor_sync(0xffffffffu, k[i], lj);
                    const bool less = o < k[i];
                    k[i] = (less == keepMin) ? o : k[i];
                }
            } else {
                #pragma unroll
                for (int i = 0; i < IPT; i++) {
                    if ((i & j) == 0) {
                        const bool up = (((tid * IPT + i) & kk) == 0) || kk == N;
                        key_cswap(k[i], k[i | j], up);
                    }
                }
            }
        }
    }
    __syncthreads();
    #pragma unroll
    for (int i = 0; i < IPT; i++) keys[tid * IPT + i] = k[i];
    __syncthreads();
}

// the six segment sorts: warp w < 6 sorts keys [w * segPad, (w + 1) * segPad) in registers
template <int IPT>
__device__ __forceinline__ void segment_sorts(unsigned long long* keys) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (w < FBPR_SEGS) {
        unsigned long long k[IPT];
        unsigned long long* base = keys + w * 32 * IPT;
        #pragma unroll
        for (int i = 0; i < IPT; i++) k[i] = base[lane * IPT + i];
        warp_bitonic_sort<IPT>(k, lane);
        #pragma unroll
        for (int i = 0; i < IPT; i++) base[lane * IPT + i] = k[i];
    }
    __syncthreads();
}

template <int IPT>
__device__ __forceinline__ void warp_sort_list(unsigned long long* base, int lane) {   // 32 * IPT keys of shared memory, one warp
    unsigned long long k[IPT];
    #pragma unroll
    for (int i = 0; i < IPT; i++) k[i] = base[lane * IPT + i];
    warp_bitonic_sort<IPT>(k, lane);
    #pragma unroll
    for (int i = 0; i < IPT; i++) base[lane * IPT + i] = k[i];
}

// bits e-5 .. e+5 (bit 5 = e itself) of a bit vector stored one word per warp, for element e = 32 * w + l
__device__ __forceinline__ unsigned bit_window3(unsigned lo, unsigned mid, unsigned hi, int l) {
    return l >= 5 ? __funnelshift_r(mid, hi, l - 5) : __funnelshift_r(lo, mid, l + 27);     // low word of (hi:mid) >> (l - 5) / (mid:lo) >> (l + 27)
}
__device__ __forceinline__ unsigned bit_window(const unsigned* words, int w, int l) {
    return bit_window3(w > 0 ? words[w - 1] : 0u, words[w], w + 1 < RING_TPB / 32 ? words[w + 1] : 0u, l);
}

enum : unsigned char { ST_NONE = 0, ST_UNDECIDED = 1, ST_PICKED = 2, ST_DEAD = 3 };

__host__ __device__ inline bool feat_sort_free(const FeatArgs& a) { return a.segPad <= RING_TPB && a.segPad >= 32; }

#ifndef FEAT_RING_CTAS
#define FEAT_RING_CTAS 3
#endif
__global__ void __launch_bounds__(RING_TPB, FEAT_RING_CTAS) feat_ring(FeatArgs a) {
    extern __shared__ unsigned char smem_raw[];
    const int slot = a.first + blockIdx.y, ring = blockIdx.x;
    const int n = a.meta[slot].n_valid;
    const int s = a.startRing[slot * a.N_SCAN + ring], e = a.endRing[slot * a.N_SCAN + ring];
    const int tid = threadIdx.x;
    // carve shared memory
    unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(smem_raw);
    const int keyCount = max(FBPR_SEGS * a.segPad, a.voxPad);
    float* s_curv = reinterpret_cast<float*>(s_keys + keyCount);
    int* s_col = reinterpret_cast<int*>(s_curv + a.wcap);
    int* s_list = s_col + a.wcap;                         // surface candidate list (global indices)
    unsigned* s_meta = reinterpret_cast<unsigned*>(s_list + a.wcap);          // rank << 16 | forward reach << 8 | backward reach
    unsigned char* s_picked = reinterpret_cast<unsigned char*>(s_meta + a.wcap);
    signed char* s_label = reinterpret_cast<signed char*>(s_picked + a.wcap);
    unsigned char* s_state0 = reinterpret_cast<unsigned char*>(s_label + a.wcap);   // flat-loop state, double-buffered by round
    unsigned char* s_state1 = s_state0 + a.wcap;
    __shared__ int s_corner[CORNERS_PER_RING];
    __shared__ int s_ncorner, s_ws[RING_TPB / 32];
    __shared__ int s_sp[FBPR_SEGS], s_ep[FBPR_SEGS], s_ccnt[FBPR_SEGS];
    __shared__ unsigned s_bb[6];
    __shared__ int s_vox[8];          // overflow, min_b[3], m1, m2
    __shared__ unsigned s_ubits[2][RING_TPB / 32], s_pbits[2][RING_TPB / 32];   // flat loop: undecided / picked bit per element, by round parity
    __shared__ float s_inv;

    const float* g_curv = a.curv + (size_t)slot * a.P;
    const int* g_col = a.colInd + (size_t)slot * a.P;
    int* g_picked = a.picked + (size_t)slot * a.P;
    int* g_label = a.label + (size_t)slot * a.P;
    const float4* g_cloud = a.cloud + (size_t)slot * a.P;

    if (tid == 0) s_ncorner = 0;
    if (tid < FBPR_SEGS) {
        int j = tid;
        s_sp[j] = (s * (6 - j) + e * j) / 6;
        s_ep[j] = (s * (5 - j) + e * (j + 1)) / 6 - 1;
    }
    // window [w0, w1] of flat indices this ring may read or mark
    const int w0 = max(s - 6, 0), w1 = min(e + 5, n - 1);
    const int W = w1 - w0 + 1;
    const bool any = (e - 1 > s) && W > 0 && W <= a.wcap && n > 0;
    __syncthreads();
    int nsurf = 0;
    if (any) {
        for (int t = tid; t < W; t += RING_TPB) {
            int g = w0 + t;
            s_curv[t] = g_curv[g]; s_col[t] = g_col[g];
            s_picked[t] = (unsigned char)(g_picked[g] != 0); s_label[t] = 0;
            if (!feat_sort_free(a)) { s_state0[t] = ST_NONE; s_state1[t] = ST_NONE; }     // only the full-sort path has state arrays
        }
        __syncthreads();
        // static suppression reach of every index (featureExtraction.h:226-240)
        for (int t = tid; t < W; t += RING_TPB) {
            int g = w0 + t, f = 0, b = 0;
            for (int l = 1; l <= 5; l++) {
                int q = g + l; if (q > w1) break;
                if (abs(s_col[q - w0] - s_col[q - 1 - w0]) > 10) break;
                f = l;
            }
            for (int l = -1; l >= -5; l--) {
                int q = g + l; if (q < 0 || q < w0) break;
                if (abs(s_col[q - w0] - s_col[q + 1 - w0]) > 10) break;
                b = -l;
            }
            s_meta[t] = ((unsigned)f << 8) | (unsigned)b;
        }
        // Two ways through the six segments.  The usual one needs NO full sort of cloudSmoothness (:203): the corner loop only
        // ever reaches the entries above edgeThreshold (it stops at the first one that is not, :208-242 over an ascending
        // array walked from the top), so only those few are listed and sorted; and the flat loop's outcome depends only on the
        // ORDER BETWEEN NEIGHBOURS within the +-5 suppression reach, which is a direct comparison of their
        // (curvature, index) keys -- slot `ep` is never sorted (:203 sorts [sp, ep)), so it ranks after everything else.
        const bool sortFree = feat_sort_free(a);                       // one thread per segment element is possible
        if (sortFree) {
            const int warp = tid >> 5, lane = tid & 31;
            if (tid < FBPR_SEGS) s_ccnt[tid] = 0;
            __syncthreads();
            for (int t = tid; t < W; t += RING_TPB) {
                const int g = w0 + t;
                if (!(g >= 5 && g < n - 5) || !(s_curv[t] > a.edgeThreshold)) continue;
                for (int j = 0; j < FBPR_SEGS; j++) {
                    if (s_sp[j] < s_ep[j] && g >= s_sp[j] && g < s_ep[j]) {
                        const int pos = atomicAdd(&s_ccnt[j], 1);
                        s_keys[j * a.segPad + pos] = ((unsigned long long)__float_as_uint(s_curv[t]) << 32) | (unsigned)g;
                    }
                }
            }
            __syncthreads();
            const int segShift = 31 - __clz(a.segPad);                       // segPad is a power of two (fbpr_launch_features)
            for (int t = tid; t < FBPR_SEGS * a.segPad; t += RING_TPB) {      // pad every list to a power of two with +inf keys
                const int j = t >> segShift, q = t - (j << segShift), c = s_ccnt[j];
                int pad = 32; while (pad < c) pad <<= 1;
                if (q >= c && q < pad) s_keys[t] = ~0ull;
            }
            __syncthreads();
            if (warp < FBPR_SEGS && s_ccnt[warp] > 1) {                      // warp j sorts list j ascending, in registers
                const int c = s_ccnt[warp];
                unsigned long long* base = s_keys + warp * a.segPad;
                if (c <= 32) { unsigned long long k[1] = { base[lane] }; warp_bitonic_sort<1>(k, lane); base[lane] = k[0]; }
                else if (c <= 64) warp_sort_list<2>(base, lane);
                else if (c <= 128) warp_sort_list<4>(base, lane);
                else if (c <= 256) warp_sort_list<8>(base, lane);
                else warp_sort_list<16>(base, lane);
            }
            __syncthreads();
#ifdef FEAT_SKIP_SEGLOOP                                                  // phase timing only (results are wrong)
            for (int j = 0; j < 0; j++) {
#else
            for (int j = 0; j < FBPR_SEGS; j++) {
#endif
                const int sp = s_sp[j], ep = s_ep[j];
                if (sp >= ep) continue;                                   // uniform across the CTA
                const unsigned long long* keys = s_keys + j * a.segPad;
                const int ncand = s_ccnt[j];
                // ---- corner loop (:208-242) by warp 0: slot ep first, then the listed entries from the largest down, 32 visiting
                // positions per batch.  Lanes hold one entry each; the picks happen strictly in visiting order (a uniform loop over
                // the eligible lanes), every pick marks its reach in the other lanes' registers and in shared memory.
                if (warp == 0) {
                    int cnt = 0, nc = s_ncorner;
                    bool stop = false;
                    for (int base = 0; base <= ncand && !stop; base += 32) {
                        const int pos = base + lane;
                        int ind = -1;
                        if (pos <= ncand) ind = pos == 0 ? ((ep >= 5 && ep < n - 5) ? ep : 0) : (int)(unsigned)(keys[ncand - pos] & 0xffffffffu);
                        const bool inwin = ind >= w0 && ind <= w1;
                        int taken = 1, f = 0, b = 0;
                        bool elig = false;
                        if (inwin) {
                            taken = s_picked[ind - w0];
                            elig = s_curv[ind - w0] > a.edgeThreshold;
                            const unsigned m = s_meta[ind - w0];
                            f = (int)((m >> 8) & 0xff); b = (int)(m & 0xff);
                        }
                        unsigned todo = __ballot_sync(0xffffffffu, elig);
                        while (todo) {
                            const int i = __ffs(todo) - 1; todo &= todo - 1;
                            if (__shfl_sync(0xffffffffu, taken, i)) continue;
                            cnt++;
                            if (cnt > FBPR_CORNERS_PER_SEG) { stop = true; break; }            // the 21st breaks before marking
                            const int pi = __shfl_sync(0xffffffffu, ind, i), pf = __shfl_sync(0xffffffffu, f, i), pb = __shfl_sync(0xffffffffu, b, i);
                            if (inwin && ind >= pi - pb && ind <= pi + pf) taken = 1;
                            if (lane <= pf + pb) s_picked[pi - pb + lane - w0] = 1;
                            if (lane == 0) { s_label[pi - w0] = 1; s_corner[nc] = pi; }
                            nc++;
                        }
                        __syncwarp();
                    }
                    if (lane == 0) {
                        s_ncorner = nc;
                        // quirk: a slot of [sp, ep) outside [5, n-5) holds {0.0f, ind 0}; it sorts first, so the flat loop starts at index 0
                        if ((sp < 5 || ep - 1 >= n - 5) && (0 < sp || 0 > ep) && 0 >= w0 && 0 <= w1) {
                            if (s_picked[0 - w0] == 0 && s_curv[0 - w0] < a.surfThreshold) {
                                s_label[0 - w0] = -1; s_picked[0 - w0] = 1;
                                int f = (s_meta[0 - w0] >> 8) & 0xff, b = s_meta[0 - w0] & 0xff;
                                for (int l = 1; l <= f; l++) s_picked[0 + l - w0] = 1;
                                for (int l = 1; l <= b; l++) s_picked[0 - l - w0] = 1;
                            }
                        }
                    }
                }
                __syncthreads();
                // ---- flat loop (:245-276) = the lexicographically-first maximal independent set in visiting order, one thread
                // per element g = sp + tid.  The elements that can decide g ("dominators": undecided at the start, visited
                // before g, covering g with their reach) never change, so they are found once as an 11-bit mask over g-5..g+5; a
                // round is then two bit-window tests against the picked / undecided bit vectors of the segment (one word per
                // warp, written with a ballot, double-buffered so a round needs one barrier).
                const int g = sp + tid;
                const bool mine = g <= ep;
                const float cg = mine ? s_curv[g - w0] : 0.f;
                bool und = mine && g >= 5 && g < n - 5 && cg < a.surfThreshold && s_picked[g - w0] == 0, pk = false;
                {
                    const unsigned bu = __ballot_sync(0xffffffffu, und);
                    if (lane == 0) { s_ubits[0][warp] = bu; s_pbits[0][warp] = 0u; }
                }
                __syncthreads();
                unsigned dom = 0;
                if (und) {
                    const unsigned uw0 = bit_window(s_ubits[0], warp, lane) & ~(1u << 5);
                    const unsigned cgb = __float_as_uint(cg);
                    #pragma unroll
                    for (int d = 0; d < 11; d++) {
                        if (!((uw0 >> d) & 1u)) continue;
                        const int q = g + d - 5;
                        const unsigned cqb = __float_as_uint(s_curv[q - w0]);
                        const bool before = g == ep ? true : (q == ep ? false : (cqb < cgb || (cqb == cgb && q < g)));
                        if (!before) continue;
                        const unsigned mq = s_meta[q - w0];
                        const bool covers = q < g ? (g - q <= (int)((mq >> 8) & 0xff)) : (q - g <= (int)(mq & 0xff));
                        if (covers) dom |= 1u << d;
                    }
                }
                // A block round: the neighbouring warps' words are read once (they may be stale -- states only move from
                // undecided to picked / suppressed, so a stale word can delay a decision but never change it), then the warp
                // iterates on its own fresh ballots until nothing changes; only chains that cross warps need another round.
                int cur = 0;
                while (true) {
                    const unsigned ulo = warp > 0 ? s_ubits[cur][warp - 1] : 0u, uhi = warp + 1 < RING_TPB / 32 ? s_ubits[cur][warp + 1] : 0u;
                    const unsigned plo = warp > 0 ? s_pbits[cur][warp - 1] : 0u, phi = warp + 1 < RING_TPB / 32 ? s_pbits[cur][warp + 1] : 0u;
                    unsigned bu = __ballot_sync(0xffffffffu, und), bp = __ballot_sync(0xffffffffu, pk);
                    while (bu) {
                        bool changed = false;
                        if (und) {
                            const unsigned uw = bit_window3(ulo, bu, uhi, lane), pw = bit_window3(plo, bp, phi, lane);
                            if (pw & dom) { und = false; changed = true; }                     // a dominator was picked: suppressed
                            else if (!(uw & dom)) { und = false; pk = true; changed = true; }  // all dominators decided, none picked
                        }
                        if (!__any_sync(0xffffffffu, changed)) break;
                        bu = __ballot_sync(0xffffffffu, und); bp = __ballot_sync(0xffffffffu, pk);
                    }
                    if (lane == 0) { s_ubits[cur ^ 1][warp] = bu; s_pbits[cur ^ 1][warp] = bp; }
                    cur ^= 1;
                    if (!__syncthreads_or(und ? 1 : 0)) break;
                }
                if (pk) {                                                  // apply picks: label -1, mark self and reach
                    s_label[g - w0] = -1; s_picked[g - w0] = 1;
                    const int f = (s_meta[g - w0] >> 8) & 0xff, bb = s_meta[g - w0] & 0xff;
                    for (int q = 1; q <= f; q++) s_picked[g + q - w0] = 1;
                    for (int q = 1; q <= bb; q++) s_picked[g - q - w0] = 1;
                }
                __syncthreads();
            }
        } else {
            // sort keys of all six segments: cloudSmoothness[k] = {curv[k], k} inside [5, n-5), else {0.0f, 0}
            for (int t = tid; t < FBPR_SEGS * a.segPad; t += RING_TPB) {
                int j = t / a.segPad, q = t - j * a.segPad;
                int sp = s_sp[j], ep = s_ep[j];
                unsigned long long key = ~0ull;
                if (sp < ep && q < ep - sp) {
                    int k = sp + q;
                    if (k >= 5 && k < n - 5) key = ((unsigned long long)__float_as_uint(s_curv[k - w0]) << 32) | (unsigned)k;
                    else key = 0ull;
                }
                s_keys[t] = key;
            }
            __syncthreads();
            if (a.segPad == 512) segment_sorts<16>(s_keys);
            else if (a.segPad == 256) segment_sorts<8>(s_keys);
            else if (a.segPad == 128) segment_sorts<4>(s_keys);
            else bitonic_blocks(s_keys, FBPR_SEGS * a.segPad, a.segPad);

            for (int j = 0; j < FBPR_SEGS; j++) {
                const int sp = s_sp[j], ep = s_ep[j];
                if (sp >= ep) continue;                                   // uniform across the CTA
                const int len = ep - sp;
                const unsigned long long* keys = s_keys + j * a.segPad;
                // ---- corner loop (:208-242), one thread, early exit once sorted curvature <= edgeThreshold
                if (tid == 0) {
                    int cnt = 0;
                    for (int v = len; v >= 0; v--) {
                        int ind; float cv;
                        if (v == len) { ind = (ep >= 5 && ep < n - 5) ? ep : 0; }
                        else { ind = (int)(unsigned)(keys[v] & 0xffffffffu); }
                        cv = (ind >= w0 && ind <= w1) ? s_curv[ind - w0] : 0.f;
                        if (v < len && !(cv > a.edgeThreshold)) break;   // sorted ascending: nothing below can qualify
                        if (ind < w0 || ind > w1) continue;
                        if (s_picked[ind - w0] == 0 && cv > a.edgeThreshold) {
                            cnt++;
                            if (cnt <= FBPR_CORNERS_PER_SEG) { s_label[ind - w0] = 1; s_corner[s_ncorner++] = ind; }
                            else break;
                            s_picked[ind - w0] = 1;
                            int f = (s_meta[ind - w0] >> 8) & 0xff, b = s_meta[ind - w0] & 0xff;
                            for (int l = 1; l <= f; l++) s_picked[ind + l - w0] = 1;
                            for (int l = 1; l <= b; l++) s_picked[ind - l - w0] = 1;
                        }
                    }
                    // quirk slot: a sorted entry whose index lies outside the segment (cloudSmoothness slot < 5 -> ind 0)
                    // is ranked first in the flat loop; handle it sequentially before the parallel rounds
                    int ind0 = (int)(unsigned)(keys[0] & 0xffffffffu);
                    if (len > 0 && (ind0 < sp || ind0 > ep) && ind0 >= w0 && ind0 <= w1) {
                        if (s_picked[ind0 - w0] == 0 && s_curv[ind0 - w0] < a.surfThreshold) {
                            s_label[ind0 - w0] = -1; s_picked[ind0 - w0] = 1;
                            int f = (s_meta[ind0 - w0] >> 8) & 0xff, b = s_meta[ind0 - w0] & 0xff;
                            for (int l = 1; l <= f; l++) s_picked[ind0 + l - w0] = 1;
                            for (int l = 1; l <= b; l++) s_picked[ind0 - l - w0] = 1;
                        }
                    }
                }
                __syncthreads();
                // ---- flat loop (:245-276) as parallel lexicographically-first MIS
                for (int v = tid; v <= len; v += RING_TPB) {
                    int ind = v == len ? ((ep >= 5 && ep < n - 5) ? ep : 0) : (int)(unsigned)(keys[v] & 0xffffffffu);
                    if (ind < sp || ind > ep) continue;
                    s_meta[ind - w0] = (s_meta[ind - w0] & 0xffffu) | ((unsigned)v << 16);
                    unsigned char st = ST_NONE;
                    if (s_curv[ind - w0] < a.surfThreshold) st = s_picked[ind - w0] ? ST_DEAD : ST_UNDECIDED;
                    s_state0[ind - w0] = st;
                }
                __syncthreads();
                if (len + 1 <= RING_TPB) {
                    // One thread per element g = sp + tid.  The candidates that can decide g ("dominators": undecided at the start,
                    // ranked before g, and covering g with their reach) never change, so they are found once as an 11-bit mask over
                    // g-5..g+5; a round is then two bit-window tests against the picked / undecided bit vectors of the segment
                    // (one word per warp, written with a ballot, double-buffered so a round needs one barrier).
                    const int g = sp + tid, w = tid >> 5, l = tid & 31;
                    const bool mine = g <= ep;
                    bool und = mine && s_state0[g - w0] == ST_UNDECIDED, pk = false;
                    unsigned dom = 0;
                    if (und) {
                        const unsigned rk = s_meta[g - w0] >> 16;
                        const int qlo = max(g - 5, sp), qhi = min(g + 5, ep);
                        for (int q = qlo; q <= qhi; q++) {
                            if (q == g || s_state0[q - w0] != ST_UNDECIDED) continue;
                            const unsigned mq = s_meta[q - w0];
                            if ((mq >> 16) >= rk) continue;
                            const bool covers = q < g ? (g - q <= (int)((mq >> 8) & 0xff)) : (q - g <= (int)(mq & 0xff));
                            if (covers) dom |= 1u << (q - g + 5);
                        }
                    }
                    int cur = 0;
                    {
                        const unsigned bu = __ballot_sync(0xffffffffu, und);
                        if (l == 0) { s_ubits[0][w] = bu; s_pbits[0][w] = 0u; }
                    }
                    __syncthreads();
                    while (true) {
                        int pending = 0;
                        if (und) {
                            // bits g-5 .. g+5 of both vectors (bit 5 = g itself)
                            const unsigned ulo = w > 0 ? s_ubits[cur][w - 1] : 0u, umid = s_ubits[cur][w], uhi = w + 1 < RING_TPB / 32 ? s_ubits[cur][w + 1] : 0u;
                            const unsigned plo = w > 0 ? s_pbits[cur][w - 1] : 0u, pmid = s_pbits[cur][w], phi = w + 1 < RING_TPB / 32 ? s_pbits[cur][w + 1] : 0u;
                            unsigned uw, pw;
                            if (l >= 5) {
                                uw = (unsigned)((((unsigned long long)uhi << 32) | umid) >> (l - 5));
                                pw = (unsigned)((((unsigned long long)phi << 32) | pmid) >> (l - 5));
                            } else {
                                uw = (unsigned)((((unsigned long long)umid << 32) | ulo) >> (l + 27));
                                pw = (unsigned)((((unsigned long long)pmid << 32) | plo) >> (l + 27));
                            }
                            if (pw & dom) und = false;                         // a dominator was picked: suppressed
                            else if (!(uw & dom)) { und = false; pk = true; }  // every dominator is decided and none picked: picked
                            else pending = 1;
                        }
                        const unsigned bu = __ballot_sync(0xffffffffu, und), bp = __ballot_sync(0xffffffffu, pk);
                        if (l == 0) { s_ubits[cur ^ 1][w] = bu; s_pbits[cur ^ 1][w] = bp; }
                        cur ^= 1;
                        if (!__syncthreads_or(pending)) break;
                    }
                    if (pk) {                                                  // apply picks: label -1, mark self and reach
                        s_label[g - w0] = -1; s_picked[g - w0] = 1;
                        const int f = (s_meta[g - w0] >> 8) & 0xff, bb = s_meta[g - w0] & 0xff;
                        for (int q = 1; q <= f; q++) s_picked[g + q - w0] = 1;
                        for (int q = 1; q <= bb; q++) s_picked[g - q - w0] = 1;
                    }
                    __syncthreads();
                    continue;
                }
                // rounds: every element of [sp, ep] is re-written into the other buffer each round, so one barrier per round
                unsigned char* cur = s_state0; unsigned char* nxt = s_state1;
                while (true) {
                    int pending = 0;
                    for (int g = sp + tid; g <= ep; g += RING_TPB) {
                        unsigned char st = cur[g - w0];
                        if (st == ST_UNDECIDED) {
                            const unsigned rk = s_meta[g - w0] >> 16;
                            bool dead = false, wait = false;
                            const int qlo = max(g - 5, sp), qhi = min(g + 5, ep);
                            for (int q = qlo; q <= qhi; q++) {
                                if (q == g) continue;
                                const unsigned char sq = cur[q - w0];
                                if (sq != ST_PICKED && sq != ST_UNDECIDED) continue;
                                const unsigned mq = s_meta[q - w0];
                                if ((mq >> 16) >= rk) continue;
                                const bool covers = q < g ? (g - q <= (int)((mq >> 8) & 0xff)) : (q - g <= (int)(mq & 0xff));
                                if (!covers) continue;
                                if (sq == ST_PICKED) dead = true; else wait = true;
                            }
                            if (dead) st = ST_DEAD;
                            else if (!wait) st = ST_PICKED;
                            else pending = 1;
                        }
                        nxt[g - w0] = st;
                    }
                    unsigned char* t_ = cur; cur = nxt; nxt = t_;
                    if (!__syncthreads_or(pending)) break;
                }
                // apply picks: label -1, mark self and reach
                for (int g = sp + tid; g <= ep; g += RING_TPB) {
                    if (cur[g - w0] == ST_PICKED) {
                        s_label[g - w0] = -1; s_picked[g - w0] = 1;
                        const int f = (s_meta[g - w0] >> 8) & 0xff, b = s_meta[g - w0] & 0xff;
                        for (int l = 1; l <= f; l++) s_picked[g + l - w0] = 1;
                        for (int l = 1; l <= b; l++) s_picked[g - l - w0] = 1;
                    }
                }
                __syncthreads();
            }
        }
        // ---- write back labels / marks (1s only: neighbouring rings' windows overlap)
        for (int t = tid; t < W; t += RING_TPB) {
            if (s_picked[t]) g_picked[w0 + t] = 1;
            if (s_label[t] != 0) g_label[w0 + t] = s_label[t];
        }
        // ---- surface candidates: every in-segment k with label <= 0, ascending (:279-284)
        int carry = 0;
        for (int base = 0; base < W; base += RING_TPB) {
            int t = base + tid, g = w0 + t;
            int flag = 0;
            if (t < W && s_label[t] <= 0) {
                for (int j = 0; j < FBPR_SEGS; j++) if (s_sp[j] < s_ep[j] && g >= s_sp[j] && g <= s_ep[j]) flag = 1;
            }
            int tot, off = block_excl_scan(flag, &tot, s_ws);
            if (flag) s_list[carry + off] = g;
            carry += tot;
        }
        nsurf = carry;
        __syncthreads();
    }
    if (tid == 0) {
        a.ringCorner[slot * a.N_SCAN + ring] = s_ncorner;
        a.ringSurf[slot * a.N_SCAN + ring] = nsurf;
    }
    for (int t = tid; t < s_ncorner; t += RING_TPB) a.cornerStage[((size_t)slot * a.N_SCAN + ring) * CORNERS_PER_RING + t] = s_corner[t];

    // ---- per-ring VoxelGrid (featureExtraction.h:287-292)
    float4* stage = a.surfStage + (size_t)slot * a.P + (size_t)ring * a.H;
    int nout = 0;
#ifdef FEAT_SKIP_VOXEL
    nsurf = 0;
#endif
#ifdef FEAT_SKIP_SEGMENTS
    if (false)
#endif
    if (nsurf > 0) {
        // the ring's surface points, staged once in the shared memory the segment phase no longer needs (s_keys + s_curv)
        float4* s_pts = reinterpret_cast<float4*>(smem_raw);
        const bool staged = (size_t)nsurf * 16 <= (size_t)keyCount * 8 + (size_t)a.wcap * 4;
        if (staged) {
            for (int t = tid; t < nsurf; t += RING_TPB) s_pts[t] = g_cloud[s_list[t]];
            __syncthreads();
        }
        unsigned mn[3] = { 0xffffffffu, 0xffffffffu, 0xffffffffu }, mx[3] = { 0u, 0u, 0u };
        for (int t = tid; t < nsurf; t += RING_TPB) {
            float4 p = staged ? s_pts[t] : g_cloud[s_list[t]];
            unsigned ex = f2ord(p.x), ey = f2ord(p.y), ez = f2ord(p.z);
            mn[0] = min(mn[0], ex); mn[1] = min(mn[1], ey); mn[2] = min(mn[2], ez);
            mx[0] = max(mx[0], ex); mx[1] = max(mx[1], ey); mx[2] = max(mx[2], ez);
        }
        if (tid < 3) s_bb[tid] = 0xffffffffu; else if (tid < 6) s_bb[tid] = 0u;
        __syncthreads();
        for (int c = 0; c < 3; c++) {
            for (int o = 16; o; o >>= 1) { mn[c] = min(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o)); mx[c] = max(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o)); }
            if ((tid & 31) == 0) { atomicMin(&s_bb[c], mn[c]); atomicMax(&s_bb[3 + c], mx[c]); }
        }
        __syncthreads();
        if (tid == 0) {
            const float inv = 1.0f / a.leaf;
            s_inv = inv;
            float fmn[3], fmx[3];
            for (int c = 0; c < 3; c++) { fmn[c] = ord2f(s_bb[c]); fmx[c] = ord2f(s_bb[3 + c]); }
            long long ex = (long long)((fmx[0] - fmn[0]) * inv) + 1, ey = (long long)((fmx[1] - fmn[1]) * inv) + 1, ez = (long long)((fmx[2] - fmn[2]) * inv) + 1;
            bool over = ex > 0x7fffffffLL || ey > 0x7fffffffLL || ez > 0x7fffffffLL;
            if (!over) { over = ex * ey > 0x7fffffffLL; if (!over) over = ex * ey * ez > 0x7fffffffLL; }
            s_vox[0] = over ? 1 : 0;
            int div[3];
            for (int c = 0; c < 3; c++) { s_vox[1 + c] = (int)floorf(fmn[c] * inv); div[c] = (int)floorf(fmx[c] * inv) - s_vox[1 + c] + 1; }
            s_vox[4] = div[0]; s_vox[5] = div[0] * div[1]; s_vox[6] = div[2];
        }
        __syncthreads();
        // 32-bit (voxel << tbits | sequence) keys and the surface points staged ONCE in shared memory, when the ring's voxel range and
        // the ring's size allow (they do for every real sweep: the bounding box of one ring at the odometry leaf has < 2^21 voxels);
        // otherwise 64-bit keys and gathers from global memory.  Same order, same sums.
        int tbits = 0; while ((1 << tbits) < a.voxPad) tbits++;
        const long long nvox = (long long)s_vox[5] * (long long)s_vox[6];
        const bool small = staged && !s_vox[0] && nvox < (1ll << (32 - tbits)) && a.voxPad <= 2 * a.wcap;
        if (s_vox[0]) {                                   // leaf too small: output = input
            for (int t = tid; t < nsurf; t += RING_TPB) stage[t] = g_cloud[s_list[t]];
            nout = nsurf;
        } else if (small) {
            unsigned* s_k32 = reinterpret_cast<unsigned*>(s_col);        // spans s_col + s_list (2 * wcap words >= voxPad)
            int* s_run = reinterpret_cast<int*>(s_meta);
            const float inv = s_inv;
            const unsigned tmask = (1u << tbits) - 1u;
            // voxel index of every staged point (s_list is dead from here on: its words are reused)
            for (int t = tid; t < nsurf; t += RING_TPB) {
                const float4 p = s_pts[t];
                const int i0 = (int)(floorf(p.x * inv) - (float)s_vox[1]);
                const int i1 = (int)(floorf(p.y * inv) - (float)s_vox[2]);
                const int i2 = (int)(floorf(p.z * inv) - (float)s_vox[3]);
                s_k32[t] = (unsigned)(i0 + i1 * s_vox[4] + i2 * s_vox[5]);
            }
            __syncthreads();
            // Consecutive points of a ring mostly fall into the same voxel (3 cm between neighbours at 10 m against a 0.4 m leaf), so
            // the ring is a few hundred RUNS of equal voxel index.  Sorting the runs by (voxel, ordinal) instead of the points by
            // (voxel, sequence) gives the same order -- a voxel's points are visited run by run, each run in sequence order -- with a
            // sort a quarter of the size.  Rings with more than 1024 runs take the per-point sort below.
            int R = 0;
            for (int base = 0; base < nsurf; base += RING_TPB) {
                const int t = base + tid;
                const int flag = (t < nsurf) && (t == 0 || s_k32[t] != s_k32[t - 1]);
                int tot; const int off = block_excl_scan(flag, &tot, s_ws);
                if (flag) s_run[R + off] = t;
                R += tot;
            }
            __syncthreads();
            const int A = (nsurf + 31) & ~31, RP = R <= RING_TPB ? RING_TPB : 2 * RING_TPB;
            if (R <= 2 * RING_TPB && A + 2 * RP <= 2 * a.wcap) {
                unsigned* s_rk = s_k32 + A;                               // run keys: voxel << tbits | run ordinal
                int* s_out = reinterpret_cast<int*>(s_rk + RP);           // first run of every output voxel
                for (int r = tid; r < RP; r += RING_TPB) s_rk[r] = r < R ? ((s_k32[s_run[r]] << tbits) | (unsigned)r) : 0xffffffffu;
                __syncthreads();
                if (RP == RING_TPB) cta_bitonic_sort<1>(s_rk); else cta_bitonic_sort<2>(s_rk);
                int V = 0;
                for (int base = 0; base < R; base += RING_TPB) {
                    const int i = base + tid;
                    const int flag = (i < R) && (i == 0 || (s_rk[i] >> tbits) != (s_rk[i - 1] >> tbits));
                    int tot; const int off = block_excl_scan(flag, &tot, s_ws);
                    if (flag) s_out[V + off] = i;
                    V += tot;
                }
                __syncthreads();
                for (int v = tid; v < V; v += RING_TPB) {                 // one thread per output voxel, sequential f32 sums in sequence order
                    const int i0 = s_out[v], i1 = v + 1 < V ? s_out[v + 1] : R;
                    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f; int cnt = 0;
                    for (int i = i0; i < i1; i++) {
                        const int r = (int)(s_rk[i] & tmask);
                        const int t0 = s_run[r], t1 = r + 1 < R ? s_run[r + 1] : nsurf;
                        for (int t = t0; t < t1; t++) { const float4 p = s_pts[t]; sx += p.x; sy += p.y; sz += p.z; si += p.w; }
                        cnt += t1 - t0;
                    }
                    const float fc = (float)cnt;
                    stage[v] = make_float4(sx / fc, sy / fc, sz / fc, si / fc);
                }
                nout = V;
            } else {
            __syncthreads();
            for (int t = tid; t < a.voxPad; t += RING_TPB) s_k32[t] = t < nsurf ? ((s_k32[t] << tbits) | (unsigned)t) : 0xffffffffu;
            __syncthreads();
            if (a.voxPad == 4 * RING_TPB) cta_bitonic_sort<4>(s_k32);
            else if (a.voxPad == 2 * RING_TPB) cta_bitonic_sort<2>(s_k32);
            else if (a.voxPad == 8 * RING_TPB) cta_bitonic_sort<8>(s_k32);
            else bitonic_blocks(s_k32, a.voxPad, a.voxPad);
            int carry = 0;
            for (int base = 0; base < nsurf; base += RING_TPB) {
                int t = base + tid;
                int flag = (t < nsurf) && (t == 0 || (s_k32[t] >> tbits) != (s_k32[t - 1] >> tbits));
                int tot, off = block_excl_scan(flag, &tot, s_ws);
                if (flag) s_run[carry + off] = t;
                carry += tot;
            }
            __syncthreads();
            const unsigned tmask = (1u << tbits) - 1u;
            for (int r = tid; r < carry; r += RING_TPB) {
                const int t = s_run[r], end = r + 1 < carry ? s_run[r + 1] : nsurf;
                float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
                for (int q = t; q < end; q++) {
                    const float4 p = s_pts[s_k32[q] & tmask];
                    sx += p.x; sy += p.y; sz += p.z; si += p.w;
                }
                const float fc = (float)(end - t);
                stage[r] = make_float4(sx / fc, sy / fc, sz / fc, si / fc);
            }
            nout = carry;
            }
        } else {
            const float inv = s_inv;
            for (int t = tid; t < a.voxPad; t += RING_TPB) {
                unsigned long long key = ~0ull;
                if (t < nsurf) {
                    float4 p = g_cloud[s_list[t]];
                    int i0 = (int)(floorf(p.x * inv) - (float)s_vox[1]);
                    int i1 = (int)(floorf(p.y * inv) - (float)s_vox[2]);
                    int i2 = (int)(floorf(p.z * inv) - (float)s_vox[3]);
                    int k = i0 + i1 * s_vox[4] + i2 * s_vox[5];
                    key = ((unsigned long long)(unsigned)k << 32) | (unsigned)t;
                }
                s_keys[t] = key;
            }
            __syncthreads();
            if (a.voxPad == 4 * RING_TPB) cta_bitonic_sort<4>(s_keys);
            else if (a.voxPad == 2 * RING_TPB) cta_bitonic_sort<2>(s_keys);
            else if (a.voxPad == 8 * RING_TPB) cta_bitonic_sort<8>(s_keys);
            else bitonic_blocks(s_keys, a.voxPad, a.voxPad);
            // runs of equal voxel index: list the run heads (s_col is free by now), then ONE THREAD PER RUN sums its points in
            // sorted order (sequential f32 sums, as pcl::VoxelGrid does) -- whole warps stay busy instead of the few head lanes
            int* s_run = s_col;
            int carry = 0;
            for (int base = 0; base < nsurf; base += RING_TPB) {
                int t = base + tid;
                int flag = (t < nsurf) && (t == 0 || (s_keys[t] >> 32) != (s_keys[t - 1] >> 32));
                int tot, off = block_excl_scan(flag, &tot, s_ws);
                if (flag) s_run[carry + off] = t;
                carry += tot;
            }
            __syncthreads();
            for (int r = tid; r < carry; r += RING_TPB) {
                const int t = s_run[r], end = r + 1 < carry ? s_run[r + 1] : nsurf;
                float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
                for (int q = t; q < end; q++) {
                    const float4 p = g_cloud[s_list[(unsigned)(s_keys[q] & 0xffffffffu)]];
                    sx += p.x; sy += p.y; sz += p.z; si += p.w;
                }
                const float fc = (float)(end - t);
                stage[r] = make_float4(sx / fc, sy / fc, sz / fc, si / fc);
            }
            nout = carry;
        }
    }
    if (tid == 0) a.ringSurfDS[slot * a.N_SCAN + ring] = nout;
}

__global__ void __launch_bounds__(256) feat_gather(FeatArgs a) {
    const int slot = a.first + blockIdx.y, ring = blockIdx.x;
    __shared__ int s_cb, s_sb, s_ct, s_st;
    if (threadIdx.x < 32) {
        int cb = 0, sb = 0, ct = 0, st = 0;
        for (int r = threadIdx.x; r < a.N_SCAN; r += 32) {
            int c = a.ringCorner[slot * a.N_SCAN + r], sdc = a.ringSurfDS[slot * a.N_SCAN + r];
            ct += c; st += sdc; if (r < ring) { cb += c; sb += sdc; }
        }
        for (int o = 16; o; o >>= 1) {
            cb += __shfl_xor_sync(0xffffffffu, cb, o); sb += __shfl_xor_sync(0xffffffffu, sb, o);
            ct += __shfl_xor_sync(0xffffffffu, ct, o); st += __shfl_xor_sync(0xffffffffu, st, o);
        }
        if (threadIdx.x == 0) { s_cb = cb; s_sb = sb; s_ct = ct; s_st = st; }
    }
    __syncthreads();
    const float4* cloud = a.cloud + (size_t)slot * a.P;
    const int nc = a.ringCorner[slot * a.N_SCAN + ring], ns = a.ringSurfDS[slot * a.N_SCAN + ring];
    float4* corner = a.corner + (size_t)slot * a.cornerCap;
    int* cidx = a.cornerIndex + (size_t)slot * a.cornerCap;
    const int* cst = a.cornerStage + ((size_t)slot * a.N_SCAN + ring) * CORNERS_PER_RING;
    for (int t = threadIdx.x; t < nc; t += blockDim.x) { int g = cst[t]; corner[s_cb + t] = cloud[g]; cidx[s_cb + t] = g; }
    float4* surf = a.surf + (size_t)slot * a.P;
    const float4* stage = a.surfStage + (size_t)slot * a.P + (size_t)ring * a.H;
    for (int t = threadIdx.x; t < ns; t += blockDim.x) surf[s_sb + t] = stage[t];
    if (ring == 0 && threadIdx.x == 0) { a.meta[slot].n_corner = s_ct; a.meta[slot].n_surf = s_st; }
}

}  // namespace

size_t fbpr_feat_ring_smem(const FeatArgs& a) {
    size_t keyCount = (size_t)(FBPR_SEGS * a.segPad > a.voxPad ? FBPR_SEGS * a.segPad : a.voxPad);
    // the two flat-loop state arrays exist only on the full-sort path; without them three CTAs fit the 196 KB carve-out and L1 doubles
    return keyCount * 8 + (size_t)a.wcap * (4 + 4 + 4 + 4 + 1 + 1 + (feat_sort_free(a) ? 0 : 2)) + 64;
}

int fbpr_launch_features(const FeatArgs& a, int count, cudaStream_t st, long long* launches) {
    if (count <= 0) return 0;
    static size_t configured = 0;
    size_t smem = fbpr_feat_ring_smem(a);
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(feat_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fbpr_fail(e, "cudaFuncSetAttribute(feat_ring smem)", __FILE__, __LINE__);
        configured = smem;
    }
    int b = (a.P + TPB1 - 1) / TPB1;
    feat_smooth<<<dim3(b, count), TPB1, 0, st>>>(a);
    feat_ring<<<dim3(a.N_SCAN, count), RING_TPB, smem, st>>>(a);
    feat_gather<<<dim3(a.N_SCAN, count), 256, 0, st>>>(a);
    if (launches) *launches += 3;
    return fbpr_launch_ok("features (feat_smooth / feat_ring / feat_gather)");
}
