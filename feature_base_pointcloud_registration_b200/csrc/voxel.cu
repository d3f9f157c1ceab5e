// voxel.cu -- batched pcl::VoxelGrid<PointXYZI>::filter for sm_100a.
//
// Replaces every VoxelGrid call site on the path (featureExtraction.h:289-290 is done inside
// the per-ring feature kernel; mapOptmization.h:251-257, :948-953, :985-991 come here).
// Algorithm: bbox -> int32 cell key -> stable LSD radix sort of
// (key, point index) -> one output point per run of equal keys = sequential f32 sum of the
// members IN POINT-INDEX ORDER / (float)count, runs emitted in ascending key order.
// Because the sort is stable and starts from index order, ties are broken by point index --
// the total order the oracle uses.
//
// Many independent problems ("segments": corner + surf cloud of every frame slot) are solved by
// one sequence of launches: blockIdx.y = segment, sizes are read from device memory.
// All kernels are HBM/latency bound integer + f32 work; no tensor cores by design.
#include "internal.cuh"

namespace {

constexpr int TILE = 2048;        // items per CTA in the tiled kernels
constexpr int TPB = 256;
constexpr int IPT = TILE / TPB;   // 8

__device__ inline unsigned warp_min_u(unsigned v) { for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }
__device__ inline unsigned warp_max_u(unsigned v) { for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }

__global__ void vox_init(const VoxSeg* segs) {
    const VoxSeg& s = segs[blockIdx.x];
    if (threadIdx.x < 3) s.bbox[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6) s.bbox[threadIdx.x] = 0u;
    if (threadIdx.x == 0) {
        int n = *s.n_in; if (n > s.cap) n = s.cap; if (n < 0) n = 0;
        s.desc->n = n; s.desc->n_out = 0; s.desc->overflow = 0; s.desc->npass = 0;
    }
}

__global__ void __launch_bounds__(TPB) vox_minmax(const VoxSeg* segs) {
    const VoxSeg& s = segs[blockIdx.y];
    int n = *s.n_in; if (n > s.cap) n = s.cap;
    unsigned mn[3] = { 0xffffffffu, 0xffffffffu, 0xffffffffu }, mx[3] = { 0u, 0u, 0u };
    for (int i = blockIdx.x * TPB + threadIdx.x; i < n; i += gridDim.x * TPB) {
        float4 p = s.in[i];
        if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) continue;      // PCL's getMinMax3D skips non-finite points of a non-dense cloud
        unsigned ex = f2ord(p.x), ey = f2ord(p.y), ez = f2ord(p.z);
        mn[0] = min(mn[0], ex); mn[1] = min(mn[1], ey); mn[2] = min(mn[2], ez);
        mx[0] = max(mx[0], ex); mx[1] = max(mx[1], ey); mx[2] = max(mx[2], ez);
    }
    __shared__ unsigned sm[6][TPB / 32];
    for (int c = 0; c < 3; c++) { mn[c] = warp_min_u(mn[c]); mx[c] = warp_max_u(mx[c]); }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) for (int c = 0; c < 3; c++) { sm[c][w] = mn[c]; sm[3 + c][w] = mx[c]; }
    __syncthreads();
    if (threadIdx.x < 6) {
        unsigned v = sm[threadIdx.x][0];
        for (int k = 1; k < TPB / 32; k++) v = threadIdx.x < 3 ? min(v, sm[threadIdx.x][k]) : max(v, sm[threadIdx.x][k]);
        if (threadIdx.x < 3) atomicMin(&s.bbox[threadIdx.x], v); else atomicMax(&s.bbox[threadIdx.x], v);
    }
}

__global__ void vox_setup(const VoxSeg* segs) {
    if (threadIdx.x != 0) return;
    const VoxSeg& s = segs[blockIdx.x];
    VoxDesc& d = *s.desc;
    if (d.n <= 0) return;
    const float inv = 1.0f / s.leaf;
    d.inv_leaf = inv;
    float mn[3], mx[3];
    for (int c = 0; c < 3; c++) { mn[c] = ord2f(s.bbox[c]); mx[c] = ord2f(s.bbox[3 + c]); }
    if (s.bbox[0] > s.bbox[3]) for (int c = 0; c < 3; c++) mn[c] = mx[c] = 0.f;      // no finite point at all
    long long ex = (long long)((mx[0] - mn[0]) * inv) + 1;
    long long ey = (long long)((mx[1] - mn[1]) * inv) + 1;
    long long ez = (long long)((mx[2] - mn[2]) * inv) + 1;
    // PCL multiplies in int64; guard the triple product against int64 wrap as well
    bool over = ex > 0x7fffffffLL || ey > 0x7fffffffLL || ez > 0x7fffffffLL;
    if (!over) { over = ex * ey > 0x7fffffffLL; if (!over) over = ex * ey * ez > 0x7fffffffLL; }
    if (over) { d.overflow = 1; return; }
    long long total = 1;
    for (int c = 0; c < 3; c++) {
        d.min_b[c] = (int)floorf(mn[c] * inv);
        int max_b = (int)floorf(mx[c] * inv);
        d.div_b[c] = max_b - d.min_b[c] + 1;
        total *= d.div_b[c];
    }
    int bits = 0; while (bits < 32 && (1LL << bits) < total) bits++;
    d.npass = (bits + 7) / 8; if (d.npass < 1) d.npass = 1;
}

__global__ void __launch_bounds__(TPB) vox_keys(const VoxSeg* segs) {
    const VoxSeg& s = segs[blockIdx.y];
    const VoxDesc d = *s.desc;
    if (d.overflow) return;
    const float inv = d.inv_leaf;
    const int m1 = d.div_b[0], m2 = d.div_b[0] * d.div_b[1];
    {
        for (int i = blockIdx.x * TPB + threadIdx.x; i < d.n; i += gridDim.x * TPB) {
            float4 p = s.in[i];
            int i0 = (int)(floorf(p.x * inv) - (float)d.min_b[0]);
            int i1 = (int)(floorf(p.y * inv) - (float)d.min_b[1]);
            int i2 = (int)(floorf(p.z * inv) - (float)d.min_b[2]);
            int key = i0 + i1 * m1 + i2 * m2;
            s.key[0][i] = (unsigned)key; s.val[0][i] = (unsigned)i;
            if (s.point_keys) s.point_keys[i] = key;
        }
    }
}

// ---- stable LSD radix sort, 8-bit digits, one (histogram, scan, scatter) triple per pass ----
__global__ void __launch_bounds__(TPB) rs_hist(const VoxSeg* segs, int pass, int tiles_cap) {
    const VoxSeg& s = segs[blockIdx.y];
    const VoxDesc d = *s.desc;
    int ntiles = (d.n + TILE - 1) / TILE;
    if ((int)blockIdx.x >= ntiles || d.overflow || pass >= d.npass) return;
    __shared__ unsigned hist[256];
    const unsigned* key = s.key[pass & 1];
    const int shift = pass * 8;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {       // a few CTAs walk the tiles of the segment
        const int base = tile * TILE;
        hist[threadIdx.x] = 0;
        __syncthreads();
        for (int k = 0; k < IPT; k++) {
            int i = base + k * TPB + threadIdx.x;
            if (i < d.n) atomicAdd(&hist[(key[i] >> shift) & 255u], 1u);
        }
        __syncthreads();
        s.tile_hist[(size_t)threadIdx.x * tiles_cap + tile] = hist[threadIdx.x];
        __syncthreads();
    }
}

// exclusive scan of tile_hist over (digit-major, tile) order; one CTA per segment
__global__ void __launch_bounds__(1024) rs_scan(const VoxSeg* segs, int pass, int tiles_cap) {
    const VoxSeg& s = segs[blockIdx.x];
    const VoxDesc d = *s.desc;
    if (d.overflow || pass >= d.npass || d.n <= 0) return;
    const int ntiles = (d.n + TILE - 1) / TILE;
    const int total = 256 * ntiles;
    __shared__ unsigned wsum[32];
    const int per = (total + 1023) / 1024;            // consecutive entries per thread
    int lo = threadIdx.x * per, hi = min(lo + per, total);
    unsigned sum = 0;
    for (int e = lo; e < hi; e++) { int dg = e / ntiles, t = e - dg * ntiles; sum += s.tile_hist[(size_t)dg * tiles_cap + t]; }
    // block exclusive scan of `sum`
    unsigned incl = sum;
    int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if (l >= o) incl += v; }
    if (l == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
        unsigned v = wsum[l], iv = v;
        for (int o = 1; o < 32; o <<= 1) { unsigned u = __shfl_up_sync(0xffffffffu, iv, o); if (l >= o) iv += u; }
        wsum[l] = iv - v;
    }
    __syncthreads();
    unsigned run = wsum[w] + incl - sum;
    for (int e = lo; e < hi; e++) {
        int dg = e / ntiles, t = e - dg * ntiles;
        size_t a = (size_t)dg * tiles_cap + t;
        unsigned v = s.tile_hist[a]; s.tile_hist[a] = run; run += v;
    }
}

__global__ void __launch_bounds__(TPB) rs_scatter(const VoxSeg* segs, int pass, int tiles_cap) {
    const VoxSeg& s = segs[blockIdx.y];
    const VoxDesc d = *s.desc;
    const int ntiles = (d.n + TILE - 1) / TILE;
    if ((int)blockIdx.x >= ntiles || d.overflow || pass >= d.npass) return;
    constexpr int NW = TPB / 32;
    __shared__ unsigned wcount[NW][256];
    const unsigned* key = s.key[pass & 1]; const unsigned* val = s.val[pass & 1];
    unsigned* okey = s.key[(pass + 1) & 1]; unsigned* oval = s.val[(pass + 1) & 1];
    const int shift = pass * 8;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {       // a few CTAs walk the tiles of the segment
    for (int k = threadIdx.x; k < NW * 256; k += TPB) (&wcount[0][0])[k] = 0;
    __syncthreads();
    // each warp owns a contiguous chunk of TILE/NW items, walked in order 32 at a time
    const int chunk = TILE / NW;
    const int wbase = tile * TILE + w * chunk;
    unsigned mykey[chunk / 32], myval[chunk / 32];
    #pragma unroll
    for (int r = 0; r < chunk / 32; r++) {
        int i = wbase + r * 32 + l;
        bool ok = i < d.n;
        mykey[r] = ok ? key[i] : 0xffffffffu; myval[r] = ok ? val[i] : 0u;
        unsigned dg = (mykey[r] >> shift) & 255u;
        unsigned act = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            unsigned peers = __match_any_sync(act, dg);
            if ((peers & ((1u << l) - 1)) == 0) wcount[w][dg] += __popc(peers);   // leader adds (one writer per digit per warp)
        }
        __syncwarp();
    }
    __syncthreads();
    // digit `threadIdx.x`: exclusive prefix over warps + global base
    {
        unsigned base = s.tile_hist[(size_t)threadIdx.x * tiles_cap + tile];
        for (int q = 0; q < NW; q++) { unsigned c = wcount[q][threadIdx.x]; wcount[q][threadIdx.x] = base; base += c; }
    }
    __syncthreads();
    #pragma unroll
    for (int r = 0; r < chunk / 32; r++) {
        int i = wbase + r * 32 + l;
        bool ok = i < d.n;
        unsigned dg = (mykey[r] >> shift) & 255u;
        unsigned act = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            unsigned peers = __match_any_sync(act, dg);
            unsigned rank = __popc(peers & ((1u << l) - 1));
            unsigned dst = wcount[w][dg] + rank;
            okey[dst] = mykey[r]; oval[dst] = myval[r];
            __syncwarp(act);
            if (rank == 0) wcount[w][dg] += __popc(peers);
        }
        __syncwarp();
    }
    __syncthreads();
    }
}

// ---- runs -> centroids -----------------------------------------------------------------
__global__ void __launch_bounds__(TPB) vox_runs_count(const VoxSeg* segs) {
    const VoxSeg& s = segs[blockIdx.y];
    const VoxDesc d = *s.desc;
    const int ntiles = (d.n + TILE - 1) / TILE;
    if ((int)blockIdx.x >= ntiles || d.overflow) return;
    const unsigned* key = s.key[d.npass & 1];
    __shared__ int ws[TPB / 32];
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int base = tile * TILE, cnt = 0;
        for (int k = 0; k < IPT; k++) {
            int i = base + k * TPB + threadIdx.x;
            if (i < d.n && (i == 0 || key[i] != key[i - 1])) cnt++;
        }
        for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = cnt;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int k = 0; k < TPB / 32; k++) t += ws[k]; s.run_tile[tile] = t; }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) vox_runs_scan(const VoxSeg* segs) {
    const VoxSeg& s = segs[blockIdx.x];
    VoxDesc& d = *s.desc;
    if (d.overflow) {
        if (threadIdx.x == 0) { d.n_out = d.n; *s.n_out = min(d.n, s.out_cap); if (d.n > s.out_cap && s.truncated) atomicOr(s.truncated, 1); }
        return;
    }
    if (d.n <= 0) { if (threadIdx.x == 0) { d.n_out = 0; *s.n_out = 0; } return; }
    const int ntiles = (d.n + TILE - 1) / TILE;
    __shared__ int wsum[32];
    const int per = (ntiles + 1023) / 1024;
    int lo = threadIdx.x * per, hi = min(lo + per, ntiles);
    int sum = 0;
    for (int e = lo; e < hi; e++) sum += s.run_tile[e];
    int incl = sum, l = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (l >= o) incl += v; }
    if (l == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
        int v = wsum[l], iv = v;
        for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, iv, o); if (l >= o) iv += u; }
        wsum[l] = iv - v;
        if (l == 31) {                                     // d.n_out keeps the true voxel count; the cloud is cut at the output capacity
            d.n_out = iv; *s.n_out = min(iv, s.out_cap);
            if (iv > s.out_cap && s.truncated) atomicOr(s.truncated, 1);
        }
    }
    __syncthreads();
    int run = wsum[w] + incl - sum;
    for (int e = lo; e < hi; e++) { int v = s.run_tile[e]; s.run_tile[e] = run; run += v; }
}

__global__ void __launch_bounds__(TPB) vox_emit(const VoxSeg* segs) {
    const VoxSeg& s = segs[blockIdx.y];
    const VoxDesc d = *s.desc;
    const int ntiles = (d.n + TILE - 1) / TILE;
    if ((int)blockIdx.x >= ntiles) return;
    __shared__ int ws[TPB / 32];
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int base = tile * TILE;
    if (d.overflow) {             // PCL's "leaf size too small" path: output = input
        for (int k = 0; k < IPT; k++) { int i = base + k * TPB + threadIdx.x; if (i < d.n && i < s.out_cap) s.out[i] = s.in[i]; }
        continue;
    }
    const unsigned* key = s.key[d.npass & 1]; const unsigned* val = s.val[d.npass & 1];
    // thread t owns items [base + t*IPT, base + (t+1)*IPT): blocked layout keeps run order
    int start = base + threadIdx.x * IPT;
    int flags = 0, cnt = 0;
    for (int k = 0; k < IPT; k++) {
        int i = start + k;
        if (i < d.n && (i == 0 || key[i] != key[i - 1])) { flags |= 1 << k; cnt++; }
    }
    int incl = cnt, l = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (l >= o) incl += v; }
    if (l == 31) ws[w] = incl;
    __syncthreads();
    int woff = 0; for (int q = 0; q < w; q++) woff += ws[q];
    int slot = s.run_tile[tile] + woff + incl - cnt;
    for (int k = 0; k < IPT; k++) {
        if (!(flags & (1 << k))) continue;
        int i = start + k;
        unsigned kk = key[i];
        float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f; int c = 0;
        for (int j = i; j < d.n && key[j] == kk; j++) {      // members in point-index order (stable sort)
            float4 p = s.in[val[j]];
            sx += p.x; sy += p.y; sz += p.z; si += p.w; c++;
        }
        float fc = (float)c;
        if (slot < s.out_cap) {
            s.out[slot] = make_float4(sx / fc, sy / fc, sz / fc, si / fc);
            if (s.out_keys) s.out_keys[slot] = (int)kk;
        }
        slot++;
    }
    __syncthreads();                                      // ws is reused by the next tile
    }
}

}  // namespace

// Enqueue the VoxelGrid of `nsegs` segments (device descriptor array) whose sizes are all <= max_n.
int fbpr_launch_voxel(const VoxSeg* d_segs, int nsegs, int max_n, int tiles_cap, cudaStream_t st, long long* launches) {
    if (nsegs <= 0) return 0;
    int tiles = (max_n + TILE - 1) / TILE; if (tiles < 1) tiles = 1;
    // the capacity is a worst case (every pixel a surface point): a few CTAs per segment walk the tiles that exist instead of a
    // launch that is mostly CTAs with nothing to do; small batches get more CTAs per segment
    const int gmax = 16 > 2368 / nsegs ? 16 : 2368 / nsegs;
    dim3 g(tiles < gmax ? tiles : gmax, nsegs);
    vox_init<<<nsegs, 32, 0, st>>>(d_segs);
    vox_minmax<<<g, TPB, 0, st>>>(d_segs);
    vox_setup<<<nsegs, 32, 0, st>>>(d_segs);
    vox_keys<<<g, TPB, 0, st>>>(d_segs);
    for (int pass = 0; pass < 4; pass++) {
        rs_hist<<<g, TPB, 0, st>>>(d_segs, pass, tiles_cap);
        rs_scan<<<nsegs, 1024, 0, st>>>(d_segs, pass, tiles_cap);
        rs_scatter<<<g, TPB, 0, st>>>(d_segs, pass, tiles_cap);
    }
    vox_runs_count<<<g, TPB, 0, st>>>(d_segs);
    vox_runs_scan<<<nsegs, 1024, 0, st>>>(d_segs);
    vox_emit<<<g, TPB, 0, st>>>(d_segs);
    if (launches) *launches += 19;
    return fbpr_launch_ok("VoxelGrid (vox_* / rs_*)");
}

int fbpr_voxel_tile() { return TILE; }
