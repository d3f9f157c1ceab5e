// capi.cu -- the C ABI (include/fbpr_b200.h): handle, HBM layout, operator sequencing.
// Plain pointers and sizes only; no torch types.  Every operator enqueues its kernels on the
// handle's stream and returns; nothing here computes on the CPU and nothing falls back.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <utility>
#include <string>
#include <tuple>
#include <vector>

#include "internal.cuh"


// ---- error reporting -------------------------------------------------------------------------
static thread_local std::string g_err;
int fbpr_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "%s:%d: %s -> %s", file, line, what, cudaGetErrorString(e));
    g_err = buf;
    return -2;
}
int fbpr_fail_msg(const char* msg) { g_err = msg; return -1; }
int fbpr_launch_ok(const char* what) {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fbpr_fail(e, what, "kernel launch", 0);
}

// ---- handle ------------------------------------------------------------------------------------
struct fbpr_handle {
    fbpr_params p;
    int device = 0;
    cudaStream_t stream = nullptr;
    long long launches = 0;
    int F = 0, P = 0, rawCap = 0, cornerCap = 0, mapCornerCap = 0, mapSurfCap = 0, kfCap = 0;
    int tilesCap = 0, cellsCorner = 0, cellsSurf = 0, cluster = 0;
    float cellCorner = 0.5f, cellSurf = 0.33f;
    std::vector<void*> allocs;
    size_t bytes = 0;
    bool guard = false; std::vector<std::pair<unsigned char*, size_t>> guards;     // FBPR_GUARD=1: (allocation base, payload bytes)
    // per-slot arrays
    FrameMeta* meta = nullptr;
    fbpr_raw_point* raw = nullptr;
    double *imuTime = nullptr, *imuRotX = nullptr, *imuRotY = nullptr, *imuRotZ = nullptr;
    int *pix = nullptr, *ringCount = nullptr, *startRing = nullptr, *endRing = nullptr, *colInd = nullptr, *winner = nullptr;
    float *range = nullptr, *curv = nullptr;
    float4 *cloud = nullptr, *surfStage = nullptr, *surf = nullptr, *surfDS = nullptr, *corner = nullptr, *cornerDS = nullptr;
    int *picked = nullptr, *label = nullptr, *ringCorner = nullptr, *cornerStage = nullptr, *ringSurf = nullptr, *ringSurfDS = nullptr, *cornerIndex = nullptr;
    float4 *mapCorner = nullptr, *mapSurf = nullptr;
    float* poseTrace = nullptr;
    float4* qanchor = nullptr; int* qcache = nullptr; double* partials = nullptr; double* partialsGrid = nullptr; double* chunkPart = nullptr; int chunkCap = 0; int lmGridBlocks = 0; bool lmWholeGpu = true;
    // descriptors
    VoxSeg* d_scanSegs = nullptr;      // [2F]  downsampleCurrentScan
    GridSeg* d_gridSegs = nullptr;     // [2F]  map index
    std::vector<GridSeg> h_gridSegs;
    VoxSeg* d_kfSegs = nullptr;        // [2F]  extractCloud VoxelGrid (keyframe concat -> local map)
    float4 *kfCorner = nullptr, *kfSurf = nullptr;   // [F][kfCap] transformed keyframe concat
    int* kfCount = nullptr;            // [F][2]
    // resident keyframe store + the scratch of one selection (keyframes.cu); grown on demand
    struct KfStore {
        float* pose6 = nullptr; double* time = nullptr; int* off[2] = { nullptr, nullptr }; float4* pool[2] = { nullptr, nullptr };
        int n = 0, poseCap = 0; long long poolCap[2] = { 0, 0 }, poolUsed[2] = { 0, 0 };
        KfSelect sel = {}; VoxSeg* d_poseSeg = nullptr; int poseTilesCap = 0; float leaf = 0.f; std::vector<void*> selAllocs;
    } kfs;
    // stand-alone ops scratch (grown on demand)
    VoxSeg* d_soloVox = nullptr; int soloVoxCap = 0; std::vector<void*> soloVoxAllocs;
    float4 *soloIn = nullptr, *soloOut = nullptr; int *soloN = nullptr, *soloNout = nullptr, *soloPK = nullptr, *soloOK = nullptr;
    GridSeg* d_soloGrid = nullptr; int soloGridCap = 0; float soloGridCell = 0; std::vector<void*> soloGridAllocs;
    float4* soloMap = nullptr; int* soloMapN = nullptr; int knnRad0 = 1;
    // debug capture (slots < dbgSlots)
    int debugIter = -1, dbgSlots = 0;
    int *knnC = nullptr, *knnS = nullptr; float *d2C = nullptr, *d2S = nullptr; float4 *coeffC = nullptr, *coeffS = nullptr;
    unsigned char *flagC = nullptr, *flagS = nullptr; float *dbgAtA = nullptr, *dbgAtB = nullptr, *dbgX = nullptr;
    // registration() scratch
    float4 *regGlobal = nullptr; int regGlobalCap = 0; int* regTile = nullptr; float* regPose = nullptr;
    int* cropTile = nullptr; size_t cropTileInts = 0;            // batched CropBox scratch: [frames][2][tiles]
    int globalCornerN = -1, globalSurfN = 0;   // resident global maps (fbpr_set_global_map)
    // batched input staging (pinned) + stage timing
    FrameMeta* h_metaStage = nullptr; double* h_imuStage = nullptr; cudaEvent_t stageDone = nullptr; bool stagePending = false;
    unsigned char* wireStage = nullptr; size_t wireStageBytes = 0;          // PointCloud2 / 32-byte PCL staging (grown on demand)
    cudaStream_t copyStream = nullptr, lmStream = nullptr, scatterStream = nullptr; std::vector<cudaEvent_t> pipeEvents;      // fbpr_register_frames: uploads overlap compute
    struct Ticket { bool busy = false; int first = 0, count = 0; cudaEvent_t done = nullptr; fbpr_result* h_res = nullptr; int cap = 0;
                    unsigned char* stage = nullptr; size_t stageBytes = 0, stageUsed = 0; };   // stage: landing area of merged uploads
    Ticket tickets[FBPR_MAX_TICKETS];                                       // batches between _begin and _end
    bool timing = false;
    bool streamTouched = true;    // an operator other than the pipelined batch calls has queued work on `stream` since the upload streams last waited for it
    struct TimedSpan { int stage; cudaEvent_t a, b; };
    std::vector<TimedSpan> spans; size_t spansUsed = 0;
    float stageMs[FBPR_STAGE_COUNT] = { 0, 0, 0, 0, 0 }; int stageCalls[FBPR_STAGE_COUNT] = { 0, 0, 0, 0, 0 };
    // graphs
    bool useGraphs = false;
    std::map<std::tuple<int, int, int, int>, cudaGraphExec_t> graphs;
    std::map<std::tuple<int, int, int, int>, long long> graphLaunches;
};

// Guarded allocations (FBPR_GUARD=1 in the environment at fbpr_create): compute-sanitizer is closed on the B200 pool this library
// is developed on, so an out-of-bounds WRITE past either end of any per-handle array is caught by 256-byte guard zones filled with
// a pattern in front of and behind every allocation, verified by fbpr_debug_check_guards (scripts/sanitize_case.py, tests).
static const size_t GUARD_BYTES = 256;
static const unsigned char GUARD_PATTERN = 0xA5;
template <typename T>
static int dev_alloc(fbpr_handle* h, T** p, size_t count, bool zero = true) {
    void* q = nullptr;
    size_t bytes = count * sizeof(T); if (bytes == 0) bytes = sizeof(T);
    const size_t g = h->guard ? GUARD_BYTES : 0;
    const size_t body = (bytes + 255) & ~(size_t)255;            // the rear guard begins right behind the payload and runs through the padding
    cudaError_t e = cudaMalloc(&q, g ? body + 2 * g : bytes);
    if (e != cudaSuccess) return fbpr_fail(e, "cudaMalloc", __FILE__, __LINE__);
    if (g) {
        e = cudaMemsetAsync(q, GUARD_PATTERN, body + 2 * g, h->stream);
        if (e != cudaSuccess) return fbpr_fail(e, "cudaMemsetAsync", __FILE__, __LINE__);
        h->guards.push_back({ static_cast<unsigned char*>(q), bytes });
    }
    h->allocs.push_back(q); h->bytes += bytes;
    q = static_cast<unsigned char*>(q) + g;
    if (zero || g) { e = cudaMemsetAsync(q, 0, bytes, h->stream); if (e != cudaSuccess) return fbpr_fail(e, "cudaMemsetAsync", __FILE__, __LINE__); }
    *p = reinterpret_cast<T*>(q);
    return 0;
}
#define ALLOC(ptr, count) do { int rc_ = dev_alloc(h, &(ptr), (size_t)(count)); if (rc_) return rc_; } while (0)

static size_t raw_src_bytes(int fmt, int n) { return (size_t)n * (fmt == FBPR_RAW_VELODYNE22 ? 22 : sizeof(fbpr_raw_point)); }
static size_t map_src_bytes(int fmt, int n) { return (size_t)n * (fmt == FBPR_MAP_XYZ12 ? 12 : sizeof(float4)); }
static int raw_kind(int fmt) { return fmt == FBPR_RAW_VELODYNE22 ? FBPR_PIECE_WIRE22 : FBPR_PIECE_COPY; }
static int map_kind(int fmt) { return fmt == FBPR_MAP_XYZ12 ? FBPR_PIECE_XYZ12 : FBPR_PIECE_COPY; }
static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }
static cudaMemcpyKind kind_in(int mem) { return mem == FBPR_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice; }

static int check_range(fbpr_handle* h, int first, int count) {
    if (!h) return fbpr_fail_msg("null handle");
    if (first < 0 || count < 0 || first + count > h->F) return fbpr_fail_msg("slot range out of bounds");
    h->streamTouched = true;      // every slot operator passes here; fbpr_register_frames_begin clears it again for its own call
    return 0;
}

static int enqueue_crop_local_maps(fbpr_handle* h, int first, int count);
// the registration stream carries the longest chain of a pipelined batch (the LM loop): it gets the highest priority, so that a
// ready LM launch is scheduled ahead of the next batch's front-end kernels
static cudaError_t create_lm_stream(fbpr_handle* h) {
    int lo = 0, hi = 0;
    cudaError_t e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (e != cudaSuccess) return e;
    const char* p = getenv("FBPR_LM_STREAM_PRIORITY");          // "0" switches the priority off (A/B measurements)
    return cudaStreamCreateWithPriority(&h->lmStream, cudaStreamNonBlocking, (p && p[0] == '0') ? lo : hi);
}

static int build_vox_segs(fbpr_handle* h, std::vector<VoxSeg>& segs, VoxSeg** d_out) {
    for (auto& s : segs) {
        for (int b = 0; b < 2; b++) { ALLOC(s.key[b], s.cap); ALLOC(s.val[b], s.cap); }
        ALLOC(s.tile_hist, (size_t)256 * h->tilesCap);
        ALLOC(s.bbox, 8);
        ALLOC(s.run_tile, h->tilesCap + 1);
        ALLOC(s.desc, 1);
    }
    ALLOC(*d_out, segs.size());
    FBPR_CUDA_OK(cudaMemcpyAsync(*d_out, segs.data(), segs.size() * sizeof(VoxSeg), cudaMemcpyHostToDevice, h->stream));
    FBPR_CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" {

const char* fbpr_last_error(void) { return g_err.c_str(); }

int fbpr_create(const fbpr_params* params, int device, fbpr_handle** out) {
    if (!params || !out) return fbpr_fail_msg("null argument");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) return fbpr_fail(e == cudaSuccess ? cudaErrorNoDevice : e, "no usable CUDA device (this library has no CPU fallback)", __FILE__, __LINE__);
    if (device < 0 || device >= ndev) return fbpr_fail_msg("device index out of range");
    FBPR_CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    FBPR_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fbpr_fail_msg("libfbpr_b200 is built for sm_100a (B200) only");
    if (params->N_SCAN <= 0 || params->Horizon_SCAN <= 0 || params->max_frames <= 0) return fbpr_fail_msg("bad N_SCAN / Horizon_SCAN / max_frames");
    fbpr_handle* h = new fbpr_handle();
    h->p = *params; h->device = device;
    { const char* g = getenv("FBPR_GUARD"); h->guard = g && g[0] == '1'; }
    FBPR_CUDA_OK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    const int F = h->F = params->max_frames;
    const int N = params->N_SCAN, H = params->Horizon_SCAN;
    const int P = h->P = N * H;
    h->rawCap = params->max_raw_points > 0 ? params->max_raw_points : (int)(P * 1.02) + 64;
    h->cornerCap = N * FBPR_SEGS * FBPR_CORNERS_PER_SEG;
    h->mapCornerCap = params->max_map_corner > 0 ? params->max_map_corner : 65536;
    h->mapSurfCap = params->max_map_surf > 0 ? params->max_map_surf : 262144;
    h->kfCap = params->max_keyframe_points;
    h->cellCorner = params->knn_cell_corner > 0 ? params->knn_cell_corner : 0.5f;
    h->cellSurf = params->knn_cell_surf > 0 ? params->knn_cell_surf : 0.4f;      // measured on config 3/4 (scripts/knn_param_sweep.py, profiles/r02_ncu_summary.md): map index + LM per 128 frames 4.95 ms at 0.33 m, 4.76 at 0.4 m (first radius 0.35 m)
    h->cellsCorner = params->grid_cells_corner > 0 ? params->grid_cells_corner : 262144;
    h->cellsSurf = params->grid_cells_surf > 0 ? params->grid_cells_surf : 1048576;
    h->cluster = params->lm_cluster_size;              // 0 = chosen per call from the batch size
    if (h->cluster < 0 || h->cluster > 16) return fbpr_fail_msg("lm_cluster_size must be 0 (auto) or 1..16");
    int maxVox = P; if (h->kfCap > maxVox) maxVox = h->kfCap;
    h->tilesCap = (maxVox + fbpr_voxel_tile() - 1) / fbpr_voxel_tile() + 1;

    ALLOC(h->meta, F);
    ALLOC(h->raw, (size_t)F * h->rawCap);
    ALLOC(h->imuTime, (size_t)F * FBPR_IMU_CAP); ALLOC(h->imuRotX, (size_t)F * FBPR_IMU_CAP);
    ALLOC(h->imuRotY, (size_t)F * FBPR_IMU_CAP); ALLOC(h->imuRotZ, (size_t)F * FBPR_IMU_CAP);
    ALLOC(h->pix, (size_t)F * P); ALLOC(h->colInd, (size_t)F * P); ALLOC(h->winner, (size_t)F * P);
    ALLOC(h->picked, (size_t)F * P); ALLOC(h->label, (size_t)F * P);
    ALLOC(h->range, (size_t)F * P); ALLOC(h->curv, (size_t)F * P);
    ALLOC(h->cloud, (size_t)F * P); ALLOC(h->surfStage, (size_t)F * P); ALLOC(h->surf, (size_t)F * P); ALLOC(h->surfDS, (size_t)F * P);
    ALLOC(h->ringCount, (size_t)F * N); ALLOC(h->startRing, (size_t)F * N); ALLOC(h->endRing, (size_t)F * N);
    ALLOC(h->ringCorner, (size_t)F * N); ALLOC(h->ringSurf, (size_t)F * N); ALLOC(h->ringSurfDS, (size_t)F * N);
    ALLOC(h->cornerStage, (size_t)F * h->cornerCap);
    ALLOC(h->corner, (size_t)F * h->cornerCap); ALLOC(h->cornerDS, (size_t)F * h->cornerCap); ALLOC(h->cornerIndex, (size_t)F * h->cornerCap);
    ALLOC(h->mapCorner, (size_t)F * h->mapCornerCap); ALLOC(h->mapSurf, (size_t)F * h->mapSurfCap);
    ALLOC(h->poseTrace, (size_t)F * FBPR_MAX_ITERS * 6);
    ALLOC(h->qanchor, (size_t)F * (h->cornerCap + P));
    ALLOC(h->qcache, (size_t)F * (h->cornerCap + P) * fbpr_knn_cache_slots());
    h->chunkCap = (h->cornerCap + P) / 32 + 260;
    ALLOC(h->chunkPart, (size_t)F * h->chunkCap * 28);
    ALLOC(h->partials, (size_t)F * 2 * 16 * 28);
    h->lmGridBlocks = fbpr_lm_grid_blocks(device);
    if (h->lmGridBlocks > 1024) h->lmGridBlocks = 1024;
    ALLOC(h->partialsGrid, (size_t)F * 2 * (h->lmGridBlocks > 0 ? h->lmGridBlocks : 1) * 28);
    h->lmWholeGpu = params->lm_single_frame_mode == 0;
    if (h->lmWholeGpu && h->lmGridBlocks <= 0) return fbpr_fail_msg("the cooperative single-frame LM kernel cannot be made resident on this device (set lm_single_frame_mode = 1 to use one cluster)");
    ALLOC(h->regPose, 16);

    // downsampleCurrentScan segments: (slot, corner), (slot, surf)
    {
        std::vector<VoxSeg> segs(2 * (size_t)F);
        for (int f = 0; f < F; f++) {
            VoxSeg c = {}; c.in = h->corner + (size_t)f * h->cornerCap; c.n_in = &h->meta[f].n_corner;
            c.out = h->cornerDS + (size_t)f * h->cornerCap; c.n_out = &h->meta[f].n_corner_ds;
            c.leaf = params->mappingCornerLeafSize; c.cap = h->cornerCap; c.out_cap = h->cornerCap;
            VoxSeg s = {}; s.in = h->surf + (size_t)f * P; s.n_in = &h->meta[f].n_surf;
            s.out = h->surfDS + (size_t)f * P; s.n_out = &h->meta[f].n_surf_ds;
            s.leaf = params->mappingSurfLeafSize; s.cap = P; s.out_cap = P;
            segs[2 * f] = c; segs[2 * f + 1] = s;
        }
        int rc = build_vox_segs(h, segs, &h->d_scanSegs); if (rc) return rc;
    }
    // map index segments
    {
        h->h_gridSegs.resize(2 * (size_t)F);
        for (int f = 0; f < F; f++) {
            for (int k = 0; k < 2; k++) {
                GridSeg g = {};
                g.cap = k == 0 ? h->mapCornerCap : h->mapSurfCap;
                g.cells_cap = k == 0 ? h->cellsCorner : h->cellsSurf;
                g.h0 = k == 0 ? h->cellCorner : h->cellSurf;
                g.pts = (k == 0 ? h->mapCorner + (size_t)f * h->mapCornerCap : h->mapSurf + (size_t)f * h->mapSurfCap);
                g.n = k == 0 ? &h->meta[f].n_map_corner : &h->meta[f].n_map_surf;
                ALLOC(g.sorted, g.cap); ALLOC(g.cell_start, g.cells_cap + 2);
                ALLOC(g.cell_of, g.cap); ALLOC(g.tile_sum, g.cells_cap / 4096 + 4); ALLOC(g.bbox, 8); ALLOC(g.desc, 1);
                h->h_gridSegs[2 * f + k] = g;
            }
        }
        ALLOC(h->d_gridSegs, 2 * (size_t)F);
        FBPR_CUDA_OK(cudaMemcpyAsync(h->d_gridSegs, h->h_gridSegs.data(), h->h_gridSegs.size() * sizeof(GridSeg), cudaMemcpyHostToDevice, h->stream));
    }
    // extractCloud: transformed keyframe concat -> VoxelGrid -> local map
    if (h->kfCap > 0) {
        ALLOC(h->kfCorner, (size_t)F * h->kfCap); ALLOC(h->kfSurf, (size_t)F * h->kfCap); ALLOC(h->kfCount, (size_t)F * 2);
        std::vector<VoxSeg> segs(2 * (size_t)F);
        for (int f = 0; f < F; f++) {
            VoxSeg c = {}; c.in = h->kfCorner + (size_t)f * h->kfCap; c.n_in = h->kfCount + 2 * f;
            c.out = h->mapCorner + (size_t)f * h->mapCornerCap; c.n_out = &h->meta[f].n_map_corner;
            c.leaf = params->mappingCornerLeafSize; c.cap = h->kfCap; c.out_cap = h->mapCornerCap; c.truncated = &h->meta[f].mapTruncated;
            VoxSeg s = {}; s.in = h->kfSurf + (size_t)f * h->kfCap; s.n_in = h->kfCount + 2 * f + 1;
            s.out = h->mapSurf + (size_t)f * h->mapSurfCap; s.n_out = &h->meta[f].n_map_surf;
            s.leaf = params->mappingSurfLeafSize; s.cap = h->kfCap; s.out_cap = h->mapSurfCap; s.truncated = &h->meta[f].mapTruncated;
            segs[2 * f] = c; segs[2 * f + 1] = s;
        }
        int rc = build_vox_segs(h, segs, &h->d_kfSegs); if (rc) return rc;
    }
    FBPR_CUDA_OK(cudaStreamSynchronize(h->stream));
    *out = h;
    return 0;
}

void fbpr_destroy(fbpr_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (auto& g : h->graphs) cudaGraphExecDestroy(g.second);
    for (auto& sp : h->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    if (h->stageDone) cudaEventDestroy(h->stageDone);
    for (auto& e : h->pipeEvents) cudaEventDestroy(e);
    for (auto& t : h->tickets) { if (t.done) cudaEventDestroy(t.done); if (t.h_res) cudaFreeHost(t.h_res); if (t.stage) cudaFree(t.stage); }
    if (h->wireStage) cudaFree(h->wireStage);
    if (h->cropTile) cudaFree(h->cropTile);
    if (h->copyStream) cudaStreamDestroy(h->copyStream);
    if (h->scatterStream) cudaStreamDestroy(h->scatterStream);
    if (h->lmStream) cudaStreamDestroy(h->lmStream);
    if (h->h_metaStage) cudaFreeHost(h->h_metaStage);
    if (h->h_imuStage) cudaFreeHost(h->h_imuStage);
    for (void* p : h->allocs) cudaFree(p);
    for (void* p : h->soloVoxAllocs) cudaFree(p);
    for (void* p : h->soloGridAllocs) cudaFree(p);
    for (void* p : h->kfs.selAllocs) cudaFree(p);
    cudaFree(h->kfs.pose6); cudaFree(h->kfs.time);
    for (int k = 0; k < 2; k++) { cudaFree(h->kfs.off[k]); cudaFree(h->kfs.pool[k]); }
    cudaStreamDestroy(h->stream);
    delete h;
}

int fbpr_debug_check_guards(fbpr_handle* h) {
    if (!h) return fbpr_fail_msg("null handle");
    if (!h->guard) return fbpr_fail_msg("handle was not created with FBPR_GUARD=1");
    cudaSetDevice(h->device);
    FBPR_CUDA_OK(cudaDeviceSynchronize());
    int bad = 0;
    std::vector<unsigned char> buf(3 * GUARD_BYTES);
    for (size_t k = 0; k < h->guards.size(); k++) {
        unsigned char* base = h->guards[k].first; const size_t bytes = h->guards[k].second;
        const size_t body = (bytes + 255) & ~(size_t)255, rear = body - bytes + GUARD_BYTES;
        FBPR_CUDA_OK(cudaMemcpy(buf.data(), base, GUARD_BYTES, cudaMemcpyDeviceToHost));
        FBPR_CUDA_OK(cudaMemcpy(buf.data() + GUARD_BYTES, base + GUARD_BYTES + bytes, rear, cudaMemcpyDeviceToHost));
        bool hit = false;
        for (size_t b = 0; b < GUARD_BYTES + rear; b++) if (buf[b] != GUARD_PATTERN) { hit = true; break; }
        if (hit) {
            bad++;
            char msg[160]; snprintf(msg, sizeof(msg), "guard zone of allocation #%zu (%zu bytes) was overwritten", k, bytes);
            g_err = msg;
        }
    }
    return bad;
}

int fbpr_sync(fbpr_handle* h) {
    if (!h) return fbpr_fail_msg("null handle");
    FBPR_CUDA_OK(cudaStreamSynchronize(h->stream));
    FBPR_CUDA_OK(cudaGetLastError());
    return 0;
}
void* fbpr_stream(fbpr_handle* h) { return h ? (void*)h->stream : nullptr; }
int64_t fbpr_kernel_launches(fbpr_handle* h) { return h ? h->launches : 0; }
int fbpr_use_graphs(fbpr_handle* h, int on) { if (!h) return fbpr_fail_msg("null handle"); h->useGraphs = on != 0; return 0; }

// ---- inputs ----------------------------------------------------------------------------------
int fbpr_set_raw_scan(fbpr_handle* h, int slot, const fbpr_raw_point* pts, int n, int mem,
                      int64_t imuAvailable, int deskewFlag, double timeScanCur,
                      const double* imuTime, const double* imuRotX, const double* imuRotY, const double* imuRotZ,
                      int imuPointerCur, float imuRollInit, float imuPitchInit) {
    int rc = check_range(h, slot, 1); if (rc) return rc;
    if (n < 0 || n > h->rawCap) return fbpr_fail_msg("raw scan larger than max_raw_points");
    if (imuAvailable != 0 && deskewFlag != -1) {
        if (!imuTime || !imuRotX || !imuRotY || !imuRotZ) return fbpr_fail_msg("IMU ramp missing while imuAvailable != 0");
        if (imuPointerCur < 0 || imuPointerCur >= FBPR_IMU_CAP) return fbpr_fail_msg("imuPointerCur out of range");
    }
    cudaSetDevice(h->device);
    if (n > 0) FBPR_CUDA_OK(cudaMemcpyAsync(h->raw + (size_t)slot * h->rawCap, pts, (size_t)n * sizeof(fbpr_raw_point), kind_in(mem), h->stream));
    if (imuAvailable != 0 && deskewFlag != -1) {
        size_t nb = (size_t)(imuPointerCur + 1) * sizeof(double);
        FBPR_CUDA_OK(cudaMemcpyAsync(h->imuTime + (size_t)slot * FBPR_IMU_CAP, imuTime, nb, cudaMemcpyHostToDevice, h->stream));
        FBPR_CUDA_OK(cudaMemcpyAsync(h->imuRotX + (size_t)slot * FBPR_IMU_CAP, imuRotX, nb, cudaMemcpyHostToDevice, h->stream));
        FBPR_CUDA_OK(cudaMemcpyAsync(h->imuRotY + (size_t)slot * FBPR_IMU_CAP, imuRotY, nb, cudaMemcpyHostToDevice, h->stream));
        FBPR_CUDA_OK(cudaMemcpyAsync(h->imuRotZ + (size_t)slot * FBPR_IMU_CAP, imuRotZ, nb, cudaMemcpyHostToDevice, h->stream));
    }
    FrameMeta m = {};
    m.n_raw = n; m.deskewFlag = deskewFlag; m.imuPointerCur = imuPointerCur; m.imuAvailable = imuAvailable;
    m.timeScanCur = timeScanCur; m.imuRollInit = imuRollInit; m.imuPitchInit = imuPitchInit;
    char* base = reinterpret_cast<char*>(h->meta + slot);
    FBPR_CUDA_OK(cudaMemcpyAsync(base + offsetof(FrameMeta, n_raw), &m.n_raw, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    FBPR_CUDA_OK(cudaMemcpyAsync(base + offsetof(FrameMeta, deskewFlag), &m.deskewFlag,
                                 offsetof(FrameMeta, pose) - offsetof(FrameMeta, deskewFlag), cudaMemcpyHostToDevice, h->stream));
    return 0;
}

static int wire_stage(fbpr_handle* h, size_t bytes) {
    if (bytes <= h->wireStageBytes) return 0;
    FBPR_CUDA_OK(cudaStreamSynchronize(h->stream));
    if (h->wireStage) cudaFree(h->wireStage);
    h->wireStage = nullptr; h->wireStageBytes = 0;
    FBPR_CUDA_OK(cudaMalloc((void**)&h->wireStage, bytes + 4096));
    h->wireStageBytes = bytes + 4096;
    return 0;
}

int fbpr_set_raw_scan_pc2(fbpr_handle* h, int slot, const void* data, int n, const fbpr_pc2_layout* L, int mem,
                          int64_t imuAvailable, double timeScanCur,
                          const double* imuTime, const double* imuRotX, const double* imuRotY, const double* imuRotZ,
                          int imuPointerCur, float imuRollInit, float imuPitchInit) {
    int rc = check_range(h, slot, 1); if (rc) return rc;
    if (!L || (!data && n > 0)) return fbpr_fail_msg("null PointCloud2 data / layout");
    if (n < 0 || n > h->rawCap) return fbpr_fail_msg("raw scan larger than max_raw_points");
    if (L->ring_bytes != 1 && L->ring_bytes != 2 && L->ring_bytes != 4)
        return fbpr_fail_msg("Point cloud ring channel not available (imageProjection.cpp:262-280)");
    const int step = L->point_step;
    auto inside = [&](int off, int sz) { return off >= 0 && off + sz <= step; };
    if (step <= 0 || !inside(L->off_x, 4) || !inside(L->off_y, 4) || !inside(L->off_z, 4) || !inside(L->off_ring, L->ring_bytes) ||
        (L->off_intensity >= 0 && !inside(L->off_intensity, 4)) || (L->off_time >= 0 && !inside(L->off_time, 4)))
        return fbpr_fail_msg("PointCloud2 field offsets outside point_step");
    cudaSetDevice(h->device);
    const unsigned char* d_src = reinterpret_cast<const unsigned char*>(data);
    if (mem == FBPR_MEM_HOST && n > 0) {
        rc = wire_stage(h, (size_t)n * step); if (rc) return rc;
        FBPR_CUDA_OK(cudaMemcpyAsync(h->wireStage, data, (size_t)n * step, cudaMemcpyHostToDevice, h->stream));
        d_src = h->wireStage;
    }
    const int deskewFlag = L->off_time >= 0 ? 1 : -1;          // imageProjection.cpp:283-297
    rc = fbpr_set_raw_scan(h, slot, nullptr, 0, FBPR_MEM_DEVICE, imuAvailable, deskewFlag, timeScanCur, imuTime, imuRotX, imuRotY, imuRotZ,
                           imuPointerCur, imuRollInit, imuPitchInit);
    if (rc) return rc;
    rc = fbpr_launch_pc2_to_raw(d_src, n, *L, h->raw + (size_t)slot * h->rawCap, h->stream, &h->launches); if (rc) return rc;
    FBPR_CUDA_OK(cudaMemcpyAsync(reinterpret_cast<char*>(h->meta + slot) + offsetof(FrameMeta, n_raw), &n, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    FBPR_CUDA_OK(cudaStreamSynchronize(h->stream));            // `n` lives on the caller's stack
    return 0;
}

int fbpr_set_clouds_xyzi32(fbpr_handle* h, int slot, int kind, const void* corner32, int nC, const void* surf32, int nS) {
    int rc = check_range(h, slot, 1); if (rc) return rc;
    if (kind != 0 && kind != 1) return fbpr_fail_msg("kind must be 0 (feature clouds) or 1 (local map)");
    const int capC = kind == 0 ? h->cornerCap : h->mapCornerCap, capS = kind == 0 ? h->P : h->mapSurfCap;
    if (nC < 0 || nC > capC || nS < 0 || nS > capS) return fbpr_fail_msg("cloud exceeds capacity");
    cudaSetDevice(h->device);
    rc = wire_stage(h, ((size_t)nC + (size_t)nS) * 32); if (rc) return rc;
    float4* dC = kind == 0 ? h->corner + (size_t)slot * h->cornerCap : h->mapCorner + (size_t)slot * h->mapCornerCap;
    float4* dS = kind == 0 ? h->surf + (size_t)slot * h->P : h->mapSurf + (size_t)slot * h->mapSurfCap;
    if (nC) FBPR_CUDA_OK(cudaMemcpyAsync(h->wireStage, corner32, (size_t)nC * 32, cudaMemcpyHostToDevice, h->stream));
    if (nS) FBPR_CUDA_OK(cudaMemcpyAsync(h->wireStage + (size_t)nC * 32, surf32, (size_t)nS * 32, cudaMemcpyHostToDevice, h->stream));
    rc = fbpr_launch_xyzi_repack(reinterpret_cast<const float4*>(h->wireStage), nC, dC, 0, h->stream, &h->launches); if (rc) return rc;
    rc = fbpr_launch_xyzi_repack(reinterpret_cast<const float4*>(h->wireStage + (size_t)nC * 32), nS, dS, 0, h->stream, &h->launches); if (rc) return rc;
    int v[3] = { nC, nS, 0 };                                  // a new local map also clears mapTruncated, which sits behind the two counts
    const size_t off = kind == 0 ? offsetof(FrameMeta, n_corner) : offsetof(FrameMeta, n_map_corner);
    FBPR_CUDA_OK(cudaMemcpyAsync(reinterpret_cast<char*>(h->meta + slot) + off, v, (kind == 0 ? 2 : 3) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    FBPR_CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}

int fbpr_set_cloud_info(fbpr_handle* h, int slot, const fbpr_cloud_info_view* ci, int mem) {
    int rc = check_range(h, slot, 1); if (rc) return rc;
    if (!ci || ci->n_valid < 0 || ci->n_valid > h->P) return fbpr_fail_msg("bad cloud_info view");
    cudaSetDevice(h->device);
    const int N = h->p.N_SCAN; const size_t nv = (size_t)ci->n_valid;
    cudaMemcpyKind k = kind_in(mem);
    FBPR_CUDA_OK(cudaMemcpyAsync(h->startRing + (size_t)slot * N, ci->startRingIndex, N * sizeof(int), k, h->stream));
    FBPR_CUDA_OK(cudaMemcpyAsync(h->endRing + (size_t)slot * N, ci->endRingIndex, N * sizeof(int), k, h->stream));
    if (nv) {
        FBPR_CUDA_OK(cudaMemcpyAsync(h->colInd + (size_t)slot * h->P, ci->pointColInd, nv * sizeof(int), k, h->stream));
        FBPR_CUDA_OK(cudaMemcpyAsync(h->range + (size_t)slot * h->P, ci->pointRange, nv * sizeof(float), k, h->stream));
        FBPR_CUDA_OK(cudaMemcpyAsync(h->cloud + (size_t)slot * h->P, ci->cloud_deskewed, nv * sizeof(float4), k, h->stream));
    }
    FrameMeta m = {};
    m.n_valid = ci->n_valid; m.imuAvailable = ci->imuAvailable; m.imuRollInit = ci->imuRollInit; m.imuPitchInit = ci->imuPitchInit;
    char* base = reinterpret_cast<char*>(h->meta + slot);
    FBPR_CUDA_OK(cudaMemcpyAsync(base + offsetof(FrameMeta, n_valid), &m.n_valid, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    FBPR_CUDA_OK(cudaMemcpyAsync(base + offsetof(FrameMeta, imuAvailable), &m.imuAvailable, sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    FBPR_CUDA_OK(cudaMemcpyAsync(base + offsetof(FrameMeta, imuRollInit), &m.imuRollInit, 2 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    return 0;
}

int fbpr_set_feature_clouds(fbpr_handle* h, int slot, const float* corner, int nC, const float* surf, int nS, int mem) {
    int rc = check_range(h, slot, 1); if (rc) return rc;
    if (nC < 0 || nC > h->cornerCap || nS < 0 || nS > h->P) return fbpr_fail_msg("feature cloud exceeds capacity");
    cudaSetDevice(h->device);
    if (nC) FBPR_CUDA_OK(cudaMemcpyAsync(h->corner + (size_t)slot * h->cornerCap, corner, (size_t)nC * sizeof(float4), kind_in(mem), h->stream));
    if (nS) FBPR_CUDA_OK(cudaMemcpyAsync(h->surf + (size_t)slot * h->P, surf, (size_t)nS * sizeof(float4), kind_in(mem), h->stream));
    int v[2] = { nC, nS };
    FBPR_CUDA_OK(cudaMemcpyAsync(reinterpret_cast<char*>(h->meta + slot) + offsetof(FrameMeta, n_corner), v, 2 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    return 0;
}

int fbpr_set_local_map(fbpr_handle* h, int slot, const float* corner, int nC, const float* surf, int nS, int mem) {
    int rc = check_range(h, slot, 1); if (rc) return rc;
    if (nC < 0 || nC > h->mapCornerCap || nS < 0 || nS > h->mapSurfCap) return fbpr_fail_msg("local map exceeds max_map_corner / max_map_surf");
    cudaSetDevice(h->device);
    if (nC) FBPR_CUDA_OK(cudaMemcpyAsync(h->mapCorner + (size_t)slot * h->mapCornerCap, corner, (size_t)nC * sizeof(float4), kind_in(mem), h->stream));
    if (nS) FBPR_CUDA_OK(cudaMemcpyAsync(h->mapSurf + (size_t)slot * h->mapSurfCap, surf, (size_t)nS * sizeof(float4), kind_in(mem), h->stream));
    static_assert(offsetof(FrameMeta, mapTruncated) == offsetof(FrameMeta, n_map_corner) + 8, "mapTruncated must follow the map counts");
    int v[3] = { nC, nS, 0 };
    FBPR_CUDA_OK(cudaMemcpyAsync(reinterpret_cast<char*>(h->meta + slot) + offsetof(FrameMeta, n_map_corner), v, 3 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    return 0;
}

int fbpr_set_pose(fbpr_handle* h, int slot, const float pose6[6]) { return fbpr_set_poses(h, slot, 1, pose6, FBPR_MEM_HOST); }

// validate `count` frames and stage their scalars (and IMU ramps) in pinned memory; *anyImu reports whether ramps were staged
static int stage_frames(fbpr_handle* h, int count, const fbpr_frame_input* fr, bool* anyImu) {
    if (!h->h_metaStage) {
        FBPR_CUDA_OK(cudaHostAlloc((void**)&h->h_metaStage, sizeof(FrameMeta) * (size_t)h->F, cudaHostAllocDefault));
        FBPR_CUDA_OK(cudaHostAlloc((void**)&h->h_imuStage, sizeof(double) * 4 * (size_t)h->F * FBPR_IMU_CAP, cudaHostAllocDefault));
        FBPR_CUDA_OK(cudaEventCreateWithFlags(&h->stageDone, cudaEventDisableTiming));
    }
    if (h->stagePending) { FBPR_CUDA_OK(cudaEventSynchronize(h->stageDone)); h->stagePending = false; }
    *anyImu = false;
    for (int i = 0; i < count; i++) {
        const fbpr_frame_input& f = fr[i];
        if (f.n_raw < 0 || f.n_raw > h->rawCap) return fbpr_fail_msg("raw scan larger than max_raw_points");
        if (f.n_map_corner < 0 || f.n_map_corner > h->mapCornerCap || f.n_map_surf < 0 || f.n_map_surf > h->mapSurfCap)
            return fbpr_fail_msg("local map exceeds max_map_corner / max_map_surf");
        if (f.raw_format != FBPR_RAW_PACKED24 && f.raw_format != FBPR_RAW_VELODYNE22) return fbpr_fail_msg("unknown raw_format");
        if (f.map_format != FBPR_MAP_XYZI16 && f.map_format != FBPR_MAP_XYZ12 && f.map_format != FBPR_MAP_FROM_GLOBAL) return fbpr_fail_msg("unknown map_format");
        if (f.map_format == FBPR_MAP_FROM_GLOBAL && h->globalCornerN < 0) return fbpr_fail_msg("map_format FROM_GLOBAL needs fbpr_set_global_map");
        if ((f.map_format == FBPR_MAP_FROM_GLOBAL) != (fr[0].map_format == FBPR_MAP_FROM_GLOBAL)) return fbpr_fail_msg("FROM_GLOBAL must be used by all frames of a batch or by none");
        if (f.map_format == FBPR_MAP_XYZ12 && ((((uintptr_t)f.map_corner_xyzi) | ((uintptr_t)f.map_surf_xyzi)) & 3)) return fbpr_fail_msg("XYZ12 maps must be 4-byte aligned");
        FrameMeta m = {};
        const bool own = f.map_format != FBPR_MAP_FROM_GLOBAL;
        m.n_raw = f.raw ? f.n_raw : 0; m.n_map_corner = own && f.map_corner_xyzi ? f.n_map_corner : 0; m.n_map_surf = own && f.map_surf_xyzi ? f.n_map_surf : 0;
        m.deskewFlag = f.deskewFlag; m.imuAvailable = f.imuAvailable; m.timeScanCur = f.timeScanCur; m.imuPointerCur = f.imuPointerCur;
        m.imuRollInit = f.imuRollInit; m.imuPitchInit = f.imuPitchInit;
        for (int q = 0; q < 6; q++) m.pose[q] = f.pose[q];
        h->h_metaStage[i] = m;
        if (f.imuAvailable != 0 && f.deskewFlag != -1) {
            if (!f.imuTime || !f.imuRotX || !f.imuRotY || !f.imuRotZ) return fbpr_fail_msg("IMU ramp missing while imuAvailable != 0");
            if (f.imuPointerCur < 0 || f.imuPointerCur >= FBPR_IMU_CAP) return fbpr_fail_msg("imuPointerCur out of range");
            const size_t nb = (size_t)(f.imuPointerCur + 1) * sizeof(double);
            const double* src[4] = { f.imuTime, f.imuRotX, f.imuRotY, f.imuRotZ };
            for (int a = 0; a < 4; a++) memcpy(h->h_imuStage + ((size_t)a * count + i) * FBPR_IMU_CAP, src[a], nb);
            *anyImu = true;
        }
    }
    return 0;
}

// enqueue the copies of the staged scalars / IMU ramps of `count` frames starting at slot `first`
static int upload_staged(fbpr_handle* h, int first, int count, bool anyImu, cudaStream_t st) {
    FBPR_CUDA_OK(cudaMemcpyAsync(h->meta + first, h->h_metaStage, sizeof(FrameMeta) * (size_t)count, cudaMemcpyHostToDevice, st));
    if (anyImu) {
        double* dst[4] = { h->imuTime, h->imuRotX, h->imuRotY, h->imuRotZ };
        for (int a = 0; a < 4; a++)
            FBPR_CUDA_OK(cudaMemcpyAsync(dst[a] + (size_t)first * FBPR_IMU_CAP, h->h_imuStage + (size_t)a * count * FBPR_IMU_CAP,
                                         sizeof(double) * (size_t)count * FBPR_IMU_CAP, cudaMemcpyHostToDevice, st));
    }
    return 0;
}

int fbpr_set_frames(fbpr_handle* h, int first, int count, const fbpr_frame_input* fr, int mem) {
    int rc = check_range(h, first, count); if (rc) return rc;
    if (count == 0) return 0;
    if (!fr) return fbpr_fail_msg("null frames");
    cudaSetDevice(h->device);
    bool anyImu = false;
    rc = stage_frames(h, count, fr, &anyImu); if (rc) return rc;
    const cudaMemcpyKind k = kind_in(mem);
    // wire-format pieces (22-byte sweeps, 12-byte maps) are repacked by stage_scatter: host buffers land in the handle's wire
    // staging area first, device buffers are read where they are
    size_t wire = 0;
    for (int i = 0; i < count; i++) {
        const fbpr_frame_input& f = fr[i]; const FrameMeta& m = h->h_metaStage[i];
        if (f.raw_format != FBPR_RAW_PACKED24) wire += raw_src_bytes(f.raw_format, m.n_raw) + 512;
        if (f.map_format != FBPR_MAP_XYZI16) wire += map_src_bytes(f.map_format, m.n_map_corner) + map_src_bytes(f.map_format, m.n_map_surf) + 1024;
    }
    if (wire && mem == FBPR_MEM_HOST) { rc = wire_stage(h, wire); if (rc) return rc; }
    size_t used = 0;
    ScatterTable t; t.stage = nullptr; t.n = 0; t.pad = 0;
    auto flush = [&]() -> int { if (t.n == 0) return 0; int r = fbpr_launch_stage_scatter(t, h->stream, &h->launches); t.n = 0; return r; };
    auto put = [&](const void* src, void* dst, size_t bytes, int kind) -> int {
        if (!bytes) return 0;
        if (kind == FBPR_PIECE_COPY) { FBPR_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, k, h->stream)); return 0; }
        const unsigned char* dsrc = static_cast<const unsigned char*>(src);
        if (mem == FBPR_MEM_HOST) {
            const size_t at = ((used + 255) & ~(size_t)255) + ((uintptr_t)src & 15);
            FBPR_CUDA_OK(cudaMemcpyAsync(h->wireStage + at, src, bytes, cudaMemcpyHostToDevice, h->stream));
            dsrc = h->wireStage + at; used = at + bytes;
        }
        t.p[t.n++] = ScatterPiece{ (unsigned long long)(uintptr_t)dsrc, dst, (unsigned long long)bytes | ((unsigned long long)kind << 60) };   // t.stage = 0: absolute addresses
        return t.n == FBPR_SCATTER_MAX ? flush() : 0;
    };
    for (int i = 0; i < count; i++) {
        const fbpr_frame_input& f = fr[i];
        const FrameMeta& m = h->h_metaStage[i];
        const int slot = first + i;
        rc = put(f.raw, h->raw + (size_t)slot * h->rawCap, raw_src_bytes(f.raw_format, m.n_raw), raw_kind(f.raw_format)); if (rc) return rc;
        rc = put(f.map_corner_xyzi, h->mapCorner + (size_t)slot * h->mapCornerCap, map_src_bytes(f.map_format, m.n_map_corner), map_kind(f.map_format)); if (rc) return rc;
        rc = put(f.map_surf_xyzi, h->mapSurf + (size_t)slot * h->mapSurfCap, map_src_bytes(f.map_format, m.n_map_surf), map_kind(f.map_format)); if (rc) return rc;
    }
    rc = flush(); if (rc) return rc;
    rc = upload_staged(h, first, count, anyImu, h->stream); if (rc) return rc;
    FBPR_CUDA_OK(cudaEventRecord(h->stageDone, h->stream));
    h->stagePending = true;
    if (fr[0].map_format == FBPR_MAP_FROM_GLOBAL) { rc = enqueue_crop_local_maps(h, first, count); if (rc) return rc; }
    return 0;
}

int fbpr_set_poses(fbpr_handle* h, int first, int count, const float* pose6, int mem) {
    int rc = check_range(h, first, count); if (rc) return rc;
    if (count == 0) return 0;
    cudaSetDevice(h->device);
    FBPR_CUDA_OK(cudaMemcpy2DAsync(reinterpret_cast<char*>(h->meta + first) + offsetof(FrameMeta, pose), sizeof(FrameMeta),
                                   pose6, 6 * sizeof(float), 6 * sizeof(float), count, kind_in(mem), h->stream));
    return 0;
}

}  // extern "C"

// ---- graph capture helper ------------------------------------------------------------------------
template <typename Fn>
static int run_op(fbpr_handle* h, int op, int first, int count, Fn&& enqueue) {
    cudaSetDevice(h->device);
    if (!h->useGraphs) return enqueue();
    auto key = std::make_tuple(op, first, count, h->debugIter);
    auto it = h->graphs.find(key);
    if (it == h->graphs.end()) {
        long long before = h->launches;
        FBPR_CUDA_OK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        int rc = enqueue();
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamEndCapture(h->stream, &g);
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) return fbpr_fail(e, "cudaStreamEndCapture", __FILE__, __LINE__);
        cudaGraphExec_t ex = nullptr;
        e = cudaGraphInstantiate(&ex, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return fbpr_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__);
        h->graphs[key] = ex;
        h->graphLaunches[key] = h->launches - before;   // kernels one replay stands for
        h->launches = before;
        it = h->graphs.find(key);
    }
    FBPR_CUDA_OK(cudaGraphLaunch(it->second, h->stream));
    h->launches += h->graphLaunches[key];
    return 0;
}

// ---- operators -------------------------------------------------------------------------------
static ProjArgs proj_args(fbpr_handle* h, int first) {
    ProjArgs a = {};
    a.meta = h->meta; a.raw = h->raw; a.rawCap = h->rawCap;
    a.imuTime = h->imuTime; a.imuRotX = h->imuRotX; a.imuRotY = h->imuRotY; a.imuRotZ = h->imuRotZ;
    a.pix = h->pix; a.ringCount = h->ringCount; a.startRing = h->startRing; a.endRing = h->endRing;
    a.colInd = h->colInd; a.range = h->range; a.cloud = h->cloud; a.winner = h->winner;
    a.N_SCAN = h->p.N_SCAN; a.H = h->p.Horizon_SCAN; a.P = h->P; a.first = first;
    return a;
}
static FeatArgs feat_args(fbpr_handle* h, int first) {
    FeatArgs a = {};
    a.meta = h->meta; a.startRing = h->startRing; a.endRing = h->endRing;
    a.colInd = h->colInd; a.range = h->range; a.cloud = h->cloud;
    a.curv = h->curv; a.picked = h->picked; a.label = h->label;
    a.ringCorner = h->ringCorner; a.cornerStage = h->cornerStage; a.ringSurf = h->ringSurf; a.ringSurfDS = h->ringSurfDS;
    a.surfStage = h->surfStage; a.corner = h->corner; a.cornerIndex = h->cornerIndex; a.cornerCap = h->cornerCap; a.surf = h->surf;
    a.N_SCAN = h->p.N_SCAN; a.H = h->p.Horizon_SCAN; a.P = h->P;
    a.edgeThreshold = h->p.edgeThreshold; a.surfThreshold = h->p.surfThreshold; a.leaf = h->p.odometrySurfLeafSize;
    a.segPad = next_pow2(a.H / 6 + 2); a.voxPad = next_pow2(a.H); a.wcap = a.H + 32;
    a.first = first;
    return a;
}
static LmArgs lm_args(fbpr_handle* h, int first) {
    LmArgs a = {};
    a.meta = h->meta; a.cornerDS = h->cornerDS; a.cornerCap = h->cornerCap; a.surfDS = h->surfDS; a.surfCap = h->P;
    a.gsegs = h->d_gridSegs; a.first = first;
    a.qanchor = h->qanchor; a.qcache = h->qcache; a.qCap = h->cornerCap + h->P; a.firstRadius = h->p.knn_first_radius > 0 ? h->p.knn_first_radius : 0.35f;
    a.partials = h->partials; a.teamMax = 16; a.partialsGrid = h->partialsGrid; a.gridMax = h->lmGridBlocks > 0 ? h->lmGridBlocks : 1; a.chunkPart = h->chunkPart; a.chunkCap = h->chunkCap;
    a.edgeMin = h->p.edgeFeatureMinValidNum; a.surfMin = h->p.surfFeatureMinValidNum;
    a.z_tol = h->p.z_tollerance; a.rot_tol = h->p.rotation_tollerance;
    a.debug_iter = h->debugIter; a.dbgSlots = h->dbgSlots;      // captured for the slots of the launch that are < dbgSlots
    a.knnC = h->knnC; a.d2C = h->d2C; a.coeffC = h->coeffC; a.flagC = h->flagC;
    a.knnS = h->knnS; a.d2S = h->d2S; a.coeffS = h->coeffS; a.flagS = h->flagS;
    a.dbgAtA = h->dbgAtA; a.dbgAtB = h->dbgAtB; a.dbgX = h->dbgX; a.poseTrace = h->poseTrace;

    return a;
}

struct StageTimer {                      // records an event pair around one stage when timing is on (never inside a capture)
    fbpr_handle* h; int idx = -1;
    StageTimer(fbpr_handle* h_, int stage) : h(h_) {
        if (!h->timing || h->useGraphs) return;
        if (h->spansUsed == h->spans.size()) {
            fbpr_handle::TimedSpan sp; sp.stage = stage;
            if (cudaEventCreate(&sp.a) != cudaSuccess || cudaEventCreate(&sp.b) != cudaSuccess) return;
            h->spans.push_back(sp);
        }
        idx = (int)h->spansUsed++;
        h->spans[idx].stage = stage;
        cudaEventRecord(h->spans[idx].a, h->stream);
    }
    ~StageTimer() { if (idx >= 0) cudaEventRecord(h->spans[idx].b, h->stream); }
};

static int enqueue_project(fbpr_handle* h, int first, int count) {
    StageTimer t(h, FBPR_STAGE_PROJECT);
    return fbpr_launch_projection(proj_args(h, first), count, h->stream, &h->launches);
}
static int enqueue_features(fbpr_handle* h, int first, int count) {
    StageTimer t(h, FBPR_STAGE_FEATURES);
    return fbpr_launch_features(feat_args(h, first), count, h->stream, &h->launches);
}
static int enqueue_downsample(fbpr_handle* h, int first, int count) {
    StageTimer t(h, FBPR_STAGE_DOWNSAMPLE);
    return fbpr_launch_voxel(h->d_scanSegs + 2 * (size_t)first, 2 * count, h->P, h->tilesCap, h->stream, &h->launches);
}
static int enqueue_map_index(fbpr_handle* h, int first, int count) {
    int maxMap = h->mapCornerCap > h->mapSurfCap ? h->mapCornerCap : h->mapSurfCap;
    int maxCells = h->cellsCorner > h->cellsSurf ? h->cellsCorner : h->cellsSurf;
    StageTimer t(h, FBPR_STAGE_MAP_INDEX);
    return fbpr_launch_grid_build(h->d_gridSegs + 2 * (size_t)first, 2 * count, maxMap, maxCells, h->stream, &h->launches);
}
static int enqueue_lm(fbpr_handle* h, int first, int count) {
    LmArgs a = lm_args(h, first);
    StageTimer t(h, FBPR_STAGE_LM);
    return fbpr_launch_lm(a, count, h->cluster, h->lmWholeGpu ? h->lmGridBlocks : 0, h->stream, &h->launches);
}
// batched CropBox of the resident global maps into the slots' local maps (on h->stream)
static int enqueue_crop_local_maps(fbpr_handle* h, int first, int count) {
    if (h->globalCornerN < 0) return fbpr_fail_msg("no global map: call fbpr_set_global_map first");
    const int nC = h->globalCornerN, nS = h->globalSurfN, need = nC > nS ? nC : nS;
    const int tilesPer = need / 2048 + 2;
    const size_t ints = (size_t)h->F * 2 * tilesPer;
    if (ints > h->cropTileInts) {
        FBPR_CUDA_OK(cudaDeviceSynchronize());
        if (h->cropTile) cudaFree(h->cropTile);
        h->cropTile = nullptr; h->cropTileInts = 0;
        FBPR_CUDA_OK(cudaMalloc((void**)&h->cropTile, ints * sizeof(int)));
        h->cropTileInts = ints;
    }
    // the scratch is indexed by slot, so batches in flight on disjoint slot ranges do not share tiles
    return fbpr_launch_crop_box_batched(h->regGlobal, nC, h->regGlobal + nC, nS, h->meta, first, count, h->mapCorner, h->mapCornerCap, h->mapSurf, h->mapSurfCap,
                                        h->cropTile + (size_t)first * 2 * tilesPer, tilesPer, h->stream, &h->launches);
}
static int enqueue_scan2map(fbpr_handle* h, int first, int count) {
    int rc = enqueue_map_index(h, first, count);
    return rc ? rc : enqueue_lm(h, first, count);
}

extern "C" {

int fbpr_project(fbpr_handle* h, int first, int count) {
    int rc = check_range(h, first, count); if (rc) return rc;
    return run_op(h, 1, first, count, [&] { return enqueue_project(h, first, count); });
}
int fbpr_feature_extract(fbpr_handle* h, int first, int count) {
    int rc = check_range(h, first, count); if (rc) return rc;
    if (fbpr_feat_ring_smem(feat_args(h, first)) > 200 * 1024) return fbpr_fail_msg("Horizon_SCAN too large for the per-ring shared-memory kernel");
    return run_op(h, 2, first, count, [&] { return enqueue_features(h, first, count); });
}
int fbpr_downsample_current_scan(fbpr_handle* h, int first, int count) {
    int rc = check_range(h, first, count); if (rc) return rc;
    return run_op(h, 3, first, count, [&] { return enqueue_downsample(h, first, count); });
}
int fbpr_scan2map_optimization(fbpr_handle* h, int first, int count) {
    int rc = check_range(h, first, count); if (rc) return rc;
    return run_op(h, 4, first, count, [&] { return enqueue_scan2map(h, first, count); });
}
int fbpr_transform_update(fbpr_handle* h, int first, int count) {
    int rc = check_range(h, first, count); if (rc) return rc;
    cudaSetDevice(h->device);
    return fbpr_launch_transform_update(h->meta, first, count, h->p.rotation_tollerance, h->p.z_tollerance, h->stream, &h->launches);
}
int fbpr_run_frames(fbpr_handle* h, int first, int count, int with_projection, int with_features) {
    int rc = check_range(h, first, count); if (rc) return rc;
    int op = 8 + (with_projection ? 1 : 0) + (with_features ? 2 : 0);
    return run_op(h, op, first, count, [&] {
        int r = 0;
        if (with_projection) r = enqueue_project(h, first, count);
        if (!r && with_features) r = enqueue_features(h, first, count);
        if (!r) r = enqueue_downsample(h, first, count);
        if (!r) r = enqueue_scan2map(h, first, count);
        return r;
    });
}

// The resident form of the pipelined schedule: the slots [first, first + count) already hold their inputs; batches of
// `batch_frames` go through the front-end (projection, features, downsample) on the handle's stream and through the map index +
// LM loop on the registration stream, so that the front-end of batch k+1 fills the SMs that the latency-bound LM kernel of
// batch k leaves idle (its tail, where the frames with few iterations have already finished).  Same kernels, same results as
// fbpr_run_frames batch by batch.
int fbpr_run_frames_pipelined(fbpr_handle* h, int first, int count, int batch_frames) {
    int rc = check_range(h, first, count); if (rc) return rc;
    if (count == 0) return 0;
    if (batch_frames <= 0) batch_frames = 128;
    if (fbpr_feat_ring_smem(feat_args(h, first)) > 200 * 1024) return fbpr_fail_msg("Horizon_SCAN too large for the per-ring shared-memory kernel");
    for (int t = 0; t < FBPR_MAX_TICKETS; t++) if (h->tickets[t].busy) return fbpr_fail_msg("a pipelined upload batch is in flight (call fbpr_register_frames_end first)");
    cudaSetDevice(h->device);
    if (!h->lmStream) FBPR_CUDA_OK(create_lm_stream(h));
    if (!h->scatterStream) FBPR_CUDA_OK(cudaStreamCreateWithFlags(&h->scatterStream, cudaStreamNonBlocking));
    cudaStream_t mapStream = h->scatterStream;                   // third stream: the map index needs only the map, not the sweep
    const int nb = (count + batch_frames - 1) / batch_frames;
    while ((int)h->pipeEvents.size() < 2 * nb + 2) {
        cudaEvent_t e; FBPR_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); h->pipeEvents.push_back(e);
    }
    const bool timing = h->timing; h->timing = false;           // stage event pairs would serialise the streams
    cudaEvent_t evStart = h->pipeEvents[2 * nb], evDone = h->pipeEvents[2 * nb + 1];
    FBPR_CUDA_OK(cudaEventRecord(evStart, h->stream));           // everything queued so far on the handle's stream
    FBPR_CUDA_OK(cudaStreamWaitEvent(h->lmStream, evStart, 0));
    FBPR_CUDA_OK(cudaStreamWaitEvent(mapStream, evStart, 0));
    cudaStream_t front = h->stream;
    for (int b = 0; b < nb && !rc; b++) {
        const int lo = first + b * batch_frames, n = (b + 1) * batch_frames <= count ? batch_frames : count - b * batch_frames;
        rc = enqueue_project(h, lo, n);
        if (!rc) rc = enqueue_features(h, lo, n);
        if (!rc) rc = enqueue_downsample(h, lo, n);
        if (rc) break;
        cudaError_t e = cudaEventRecord(h->pipeEvents[2 * b], front);
        h->stream = mapStream;                                   // enqueue_* launch on h->stream
        if (e == cudaSuccess) rc = enqueue_map_index(h, lo, n);
        if (e == cudaSuccess && !rc) e = cudaEventRecord(h->pipeEvents[2 * b + 1], mapStream);
        h->stream = h->lmStream;
        if (e == cudaSuccess && !rc) e = cudaStreamWaitEvent(h->lmStream, h->pipeEvents[2 * b], 0);
        if (e == cudaSuccess && !rc) e = cudaStreamWaitEvent(h->lmStream, h->pipeEvents[2 * b + 1], 0);
        if (e == cudaSuccess && !rc) rc = enqueue_lm(h, lo, n);
        h->stream = front;
        if (e != cudaSuccess) { h->timing = timing; return fbpr_fail(e, "event record / wait", __FILE__, __LINE__); }
    }
    h->timing = timing;
    if (rc) { cudaStreamSynchronize(h->stream); cudaStreamSynchronize(h->lmStream); cudaStreamSynchronize(mapStream); return rc; }
    FBPR_CUDA_OK(cudaEventRecord(evDone, h->lmStream));          // later operators on the handle's stream see the finished slots
    FBPR_CUDA_OK(cudaStreamWaitEvent(h->stream, evDone, 0));
    return 0;
}

// chunk schedule of the pipelined call: uniform chunks of `chunk_frames` (0 = 32).  A tapered tail (.., 16, 8, 8) was measured
// and is slower (20.6 vs 19.3 ms per 128 frames for a lone call: small LM batches cost more than the shorter tail saves) and
// makes no difference once two batches are in flight (16.8 ms, 97 % of the PCIe floor).
static void chunk_schedule(int count, int chunk_frames, std::vector<int>& bounds) {
    const int chunk = chunk_frames > 0 ? chunk_frames : 32;
    bounds.assign(1, 0);
    for (int lo = 0; lo < count; lo += chunk) bounds.push_back(lo + chunk < count ? lo + chunk : count);
}

// Upload a group of host buffers.  A pinned H2D copy costs ~4 us on top of its bytes (measured: 384 copies of 0.6-3 MB reach
// 50 GB/s, one copy of the same bytes 55.6 GB/s), so when the buffers of the group are packed densely in host memory (a caller
// that fills one pinned arena) their whole span crosses PCIe as ONE copy into a landing area and a small kernel scatters the
// pieces to their slots at HBM speed; otherwise one copy per buffer.
struct UploadPiece { const void* src; void* dst; size_t bytes; int kind; };      // bytes = SOURCE bytes; kind = FBPR_PIECE_*
// true when [lo, hi) lies inside ONE pinned allocation known to the driver (cuMemGetAddressRange through the runtime's driver
// entry point: no link-time dependency on libcuda).  Reading the gaps between separately allocated buffers would be a bug.
static bool host_span_is_one_allocation(uintptr_t lo, uintptr_t hi) {
    typedef int (*RangeFn)(unsigned long long*, size_t*, unsigned long long);
    static RangeFn fn = nullptr; static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr; cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess) fn = (RangeFn)f;
        else cudaGetLastError();
    }
    if (!fn) return false;
    unsigned long long base = 0; size_t size = 0;
    if (fn(&base, &size, (unsigned long long)lo) != 0) return false;
    return (unsigned long long)lo >= base && (unsigned long long)hi <= base + size;
}
// `ready` is recorded when the group is in place: on the copy stream, or (merged) on the scatter stream so that the next copy
// does not wait for the scatter kernel
static int upload_group(fbpr_handle* h, fbpr_handle::Ticket& tk, const std::vector<UploadPiece>& pcs, cudaStream_t st, cudaStream_t scatterSt, cudaEvent_t tmp, cudaEvent_t ready) {
    size_t sum = 0; uintptr_t lo = ~(uintptr_t)0, hi = 0;
    bool repack = false;                                         // wire-format pieces must pass through the landing area
    for (const auto& p : pcs) {
        sum += p.bytes;
        const uintptr_t a = (uintptr_t)p.src;
        if (a < lo) lo = a;
        if (a + p.bytes > hi) hi = a + p.bytes;
        repack = repack || p.kind != FBPR_PIECE_COPY;
    }
    const size_t span = pcs.empty() ? 0 : (size_t)(hi - lo);
    const size_t landing = (tk.stageUsed + 255) & ~(size_t)255;
    bool merged = pcs.size() >= 2 && pcs.size() <= FBPR_SCATTER_MAX && span <= sum + sum / 16 + 4096 && tk.stage && landing + span + 32 <= tk.stageBytes;
    if (merged) for (const auto& p : pcs) if (p.kind == FBPR_PIECE_COPY && ((lo & 3) || (p.bytes & 3) || (((uintptr_t)p.src - lo) & 3))) { merged = false; break; }
    if (merged) merged = host_span_is_one_allocation(lo, hi);
    if (!merged && !repack) {
        for (const auto& p : pcs) FBPR_CUDA_OK(cudaMemcpyAsync(p.dst, p.src, p.bytes, cudaMemcpyHostToDevice, st));
        FBPR_CUDA_OK(cudaEventRecord(ready, st));
        return 0;
    }
    cudaEvent_t last = ready;
    if (merged) {
        // keep the landing address congruent to the host address modulo 16 so that 16-byte-aligned pieces stay aligned
        unsigned char* land = tk.stage + landing + (lo & 15);
        FBPR_CUDA_OK(cudaMemcpyAsync(land, (const void*)lo, span, cudaMemcpyHostToDevice, st));
        ScatterTable t; t.stage = land; t.n = (int)pcs.size(); t.pad = 0;
        for (size_t i = 0; i < pcs.size(); i++)
            t.p[i] = ScatterPiece{ (unsigned long long)((uintptr_t)pcs[i].src - lo), pcs[i].dst, (unsigned long long)pcs[i].bytes | ((unsigned long long)pcs[i].kind << 60) };
        FBPR_CUDA_OK(cudaEventRecord(tmp, st));
        FBPR_CUDA_OK(cudaStreamWaitEvent(scatterSt, tmp, 0));
        { int rc = fbpr_launch_stage_scatter(t, scatterSt, &h->launches); if (rc) return rc; }
        tk.stageUsed = landing + (lo & 15) + span;
    } else {
        // scattered host buffers in a wire format: one copy per piece into the landing area (congruent modulo 16), repacked in
        // groups of FBPR_SCATTER_MAX; plain pieces go straight to their slots
        if (!tk.stage) return fbpr_fail_msg("no landing area for wire-format uploads (out of device memory)");
        size_t used = tk.stageUsed;
        ScatterTable t; t.stage = tk.stage; t.n = 0; t.pad = 0;
        auto flush = [&]() -> int {
            if (t.n == 0) return 0;
            FBPR_CUDA_OK(cudaEventRecord(tmp, st));
            FBPR_CUDA_OK(cudaStreamWaitEvent(scatterSt, tmp, 0));
            int rc = fbpr_launch_stage_scatter(t, scatterSt, &h->launches);
            t.n = 0;
            return rc;
        };
        for (const auto& p : pcs) {
            if (p.kind == FBPR_PIECE_COPY) { FBPR_CUDA_OK(cudaMemcpyAsync(p.dst, p.src, p.bytes, cudaMemcpyHostToDevice, st)); continue; }
            const size_t at = ((used + 255) & ~(size_t)255) + ((uintptr_t)p.src & 15);
            if (at + p.bytes + 32 > tk.stageBytes) return fbpr_fail_msg("landing area too small for the wire-format uploads");
            FBPR_CUDA_OK(cudaMemcpyAsync(tk.stage + at, p.src, p.bytes, cudaMemcpyHostToDevice, st));
            t.p[t.n++] = ScatterPiece{ (unsigned long long)at, p.dst, (unsigned long long)p.bytes | ((unsigned long long)p.kind << 60) };
            used = at + p.bytes;
            if (t.n == FBPR_SCATTER_MAX) { int rc = flush(); if (rc) return rc; }
        }
        { int rc = flush(); if (rc) return rc; }
        tk.stageUsed = used;
        FBPR_CUDA_OK(cudaEventRecord(tmp, st));                   // the plain pieces of the group
        FBPR_CUDA_OK(cudaStreamWaitEvent(scatterSt, tmp, 0));
    }
    FBPR_CUDA_OK(cudaEventRecord(last, scatterSt));
    return 0;
}

// everything of _begin that enqueues work; a failure midway leaves copies / kernels queued, which the caller below drains
static int register_frames_enqueue(fbpr_handle* h, int ticket, int first, int count, const fbpr_frame_input* fr, int chunk_frames) {
    int rc = 0;
    fbpr_handle::Ticket& tk = h->tickets[ticket];
    int inflight = 0;
    for (int t = 0; t < FBPR_MAX_TICKETS; t++) inflight += h->tickets[t].busy ? 1 : 0;
    if (count == 0) { FBPR_CUDA_OK(cudaEventRecord(tk.done, h->stream)); return 0; }
    std::vector<int> bounds; chunk_schedule(count, chunk_frames, bounds);
    const int nchunks = (int)bounds.size() - 1;
    while ((int)h->pipeEvents.size() < 5 * nchunks + 3) {
        cudaEvent_t e; FBPR_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); h->pipeEvents.push_back(e);
    }
    bool anyImu = false;
    rc = stage_frames(h, count, fr, &anyImu); if (rc) return rc;
    cudaEvent_t evStart = h->pipeEvents[3 * nchunks], evMeta = h->pipeEvents[3 * nchunks + 1];
    if (inflight == 0 || h->streamTouched) {
        // the upload and registration streams start after everything already queued on the handle's stream: earlier operators
        // (fbpr_run_frames, fbpr_registration, setters, asynchronous getters) may still read or write these slots.  Only when
        // nothing but pipelined batches has used the handle since the last such wait (their slot ranges are disjoint and their
        // previous occupants were drained by fbpr_register_frames_end) may the copies start at once, under that batch's kernels.
        FBPR_CUDA_OK(cudaEventRecord(evStart, h->stream));
        FBPR_CUDA_OK(cudaStreamWaitEvent(h->copyStream, evStart, 0));
        FBPR_CUDA_OK(cudaStreamWaitEvent(h->lmStream, evStart, 0));
        h->streamTouched = false;
    }
    rc = upload_staged(h, first, count, anyImu, h->copyStream); if (rc) return rc;
    FBPR_CUDA_OK(cudaEventRecord(evMeta, h->copyStream));
    FBPR_CUDA_OK(cudaEventRecord(h->stageDone, h->copyStream));
    h->stagePending = true;
    // three streams.  Upload: raw sweeps of chunk k, then its local maps.  Front-end (the handle's stream): projection, features
    // and downsample of chunk k as soon as its sweeps are in HBM.  Registration (second compute stream): map index + LM of chunk k
    // as soon as its maps are in HBM and its front-end is done -- so the PCIe copies of chunk k+1 and the front-end of chunk k+1
    // run under the latency-bound LM kernel of chunk k.
    {   // landing area for merged uploads: as large as everything this batch uploads (plus slack), kept with the ticket
        size_t need = 0;
        for (int i = 0; i < count; i++) {
            const FrameMeta& m = h->h_metaStage[i];
            need += (size_t)m.n_raw * sizeof(fbpr_raw_point) + ((size_t)m.n_map_corner + (size_t)m.n_map_surf) * sizeof(float4);
        }
        need += need / 8 + (size_t)nchunks * 2 * 8192 + 65536 + (size_t)count * 3 * 512;      // upper bound for the wire formats too, with per-piece alignment slack
        if (tk.stageBytes < need) {
            if (tk.stage) { FBPR_CUDA_OK(cudaStreamSynchronize(h->copyStream)); cudaFree(tk.stage); tk.stage = nullptr; tk.stageBytes = 0; }
            if (cudaMalloc((void**)&tk.stage, need) == cudaSuccess) tk.stageBytes = need; else { cudaGetLastError(); tk.stage = nullptr; }   // no landing area: plain copies
        }
        tk.stageUsed = 0;
    }
    std::vector<UploadPiece> pcs;
    for (int c = 0; c < nchunks; c++) {
        const int lo = bounds[c], hi = bounds[c + 1];
        pcs.clear();
        for (int i = lo; i < hi; i++) {
            const FrameMeta& m = h->h_metaStage[i];
            if (m.n_raw) pcs.push_back(UploadPiece{ fr[i].raw, h->raw + (size_t)(first + i) * h->rawCap, raw_src_bytes(fr[i].raw_format, m.n_raw), raw_kind(fr[i].raw_format) });
        }
        rc = upload_group(h, tk, pcs, h->copyStream, h->scatterStream, h->pipeEvents[3 * nchunks + 3 + 2 * c], h->pipeEvents[3 * c]); if (rc) return rc;
        pcs.clear();
        for (int i = lo; i < hi; i++) {
            const FrameMeta& m = h->h_metaStage[i];
            if (m.n_map_corner) pcs.push_back(UploadPiece{ fr[i].map_corner_xyzi, h->mapCorner + (size_t)(first + i) * h->mapCornerCap, map_src_bytes(fr[i].map_format, m.n_map_corner), map_kind(fr[i].map_format) });
            if (m.n_map_surf) pcs.push_back(UploadPiece{ fr[i].map_surf_xyzi, h->mapSurf + (size_t)(first + i) * h->mapSurfCap, map_src_bytes(fr[i].map_format, m.n_map_surf), map_kind(fr[i].map_format) });
        }
        rc = upload_group(h, tk, pcs, h->copyStream, h->scatterStream, h->pipeEvents[3 * nchunks + 4 + 2 * c], h->pipeEvents[3 * c + 1]); if (rc) return rc;
    }
    FBPR_CUDA_OK(cudaStreamWaitEvent(h->stream, evMeta, 0));
    FBPR_CUDA_OK(cudaStreamWaitEvent(h->lmStream, evMeta, 0));
    for (int c = 0; c < nchunks; c++) {
        const int lo = first + bounds[c], n = bounds[c + 1] - bounds[c];
        FBPR_CUDA_OK(cudaStreamWaitEvent(h->stream, h->pipeEvents[3 * c], 0));
        rc = enqueue_project(h, lo, n); if (rc) return rc;
        rc = enqueue_features(h, lo, n); if (rc) return rc;
        rc = enqueue_downsample(h, lo, n); if (rc) return rc;
        FBPR_CUDA_OK(cudaEventRecord(h->pipeEvents[3 * c + 2], h->stream));
        FBPR_CUDA_OK(cudaStreamWaitEvent(h->lmStream, h->pipeEvents[3 * c + 2], 0));
        FBPR_CUDA_OK(cudaStreamWaitEvent(h->lmStream, h->pipeEvents[3 * c + 1], 0));
        std::swap(h->stream, h->lmStream);                       // enqueue_* launch on h->stream
        rc = fr[0].map_format == FBPR_MAP_FROM_GLOBAL ? enqueue_crop_local_maps(h, lo, n) : 0;      // local maps cut from the resident global maps
        if (!rc) rc = enqueue_scan2map(h, lo, n);
        std::swap(h->stream, h->lmStream);
        if (rc) return rc;
    }
    // results: packed D2H on the registration stream, right behind the last LM kernel (the front-end stream stays free for the next batch)
    FBPR_CUDA_OK(cudaMemcpy2DAsync(tk.h_res, sizeof(fbpr_result), reinterpret_cast<char*>(h->meta + first) + offsetof(FrameMeta, pose), sizeof(FrameMeta),
                                   sizeof(fbpr_result), count, cudaMemcpyDeviceToHost, h->lmStream));
    FBPR_CUDA_OK(cudaEventRecord(tk.done, h->lmStream));
    return 0;
}

int fbpr_register_frames_begin(fbpr_handle* h, int first, int count, const fbpr_frame_input* fr, int chunk_frames) {
    const bool touched = h ? h->streamTouched : true;
    int rc = check_range(h, first, count); if (rc) return rc;
    h->streamTouched = touched;
    if (!fr && count > 0) return fbpr_fail_msg("null frames");
    if (count > 0 && fbpr_feat_ring_smem(feat_args(h, first)) > 200 * 1024) return fbpr_fail_msg("Horizon_SCAN too large for the per-ring shared-memory kernel");
    cudaSetDevice(h->device);
    int ticket = -1;
    for (int t = 0; t < FBPR_MAX_TICKETS; t++) if (!h->tickets[t].busy) { ticket = t; break; }
    if (ticket < 0) return fbpr_fail_msg("too many fbpr_register_frames_begin calls in flight (call fbpr_register_frames_end first)");
    for (int t = 0; t < FBPR_MAX_TICKETS; t++) {
        const auto& o = h->tickets[t];
        if (o.busy && first < o.first + o.count && o.first < first + count) return fbpr_fail_msg("slot range overlaps a batch that is still in flight");
    }
    fbpr_handle::Ticket& tk = h->tickets[ticket];
    if (!tk.done) FBPR_CUDA_OK(cudaEventCreateWithFlags(&tk.done, cudaEventDisableTiming));
    if (tk.cap < count) {
        if (tk.h_res) cudaFreeHost(tk.h_res);
        tk.h_res = nullptr; tk.cap = 0;
        FBPR_CUDA_OK(cudaHostAlloc((void**)&tk.h_res, sizeof(fbpr_result) * (size_t)(count > 0 ? count : 1), cudaHostAllocDefault));
        tk.cap = count;
    }
    tk.first = first; tk.count = count;
    if (!h->copyStream) FBPR_CUDA_OK(cudaStreamCreateWithFlags(&h->copyStream, cudaStreamNonBlocking));
    if (!h->lmStream) FBPR_CUDA_OK(create_lm_stream(h));
    if (!h->scatterStream) FBPR_CUDA_OK(cudaStreamCreateWithFlags(&h->scatterStream, cudaStreamNonBlocking));
    tk.busy = true;                                              // from here on work may be queued for these slots
    rc = register_frames_enqueue(h, ticket, first, count, fr, chunk_frames);
    if (rc) {
        // a failure midway (a launch or copy was refused) leaves earlier copies / kernels of this batch queued: wait for them so
        // that the caller's host buffers and the slots are free again, then give the ticket back.  The error text is kept.
        const std::string why = g_err;
        cudaStreamSynchronize(h->copyStream); cudaStreamSynchronize(h->scatterStream); cudaStreamSynchronize(h->stream); cudaStreamSynchronize(h->lmStream);
        cudaGetLastError();
        tk.busy = false;
        g_err = why;
        return rc;
    }
    return ticket;
}


int fbpr_register_frames_end(fbpr_handle* h, int ticket, fbpr_result* out) {
    if (!h) return fbpr_fail_msg("null handle");
    if (ticket < 0 || ticket >= FBPR_MAX_TICKETS || !h->tickets[ticket].busy) return fbpr_fail_msg("no such batch in flight");
    cudaSetDevice(h->device);
    fbpr_handle::Ticket& tk = h->tickets[ticket];
    FBPR_CUDA_OK(cudaEventSynchronize(tk.done));
    tk.busy = false;
    FBPR_CUDA_OK(cudaGetLastError());
    if (out && tk.count > 0) memcpy(out, tk.h_res, sizeof(fbpr_result) * (size_t)tk.count);
    bool any = false;
    for (int t = 0; t < FBPR_MAX_TICKETS; t++) any = any || h->tickets[t].busy;
    if (!any) {                                                  // later operators on the handle's stream see the finished slots
        FBPR_CUDA_OK(cudaStreamWaitEvent(h->stream, tk.done, 0));
    }
    return tk.count;
}

int fbpr_register_frames(fbpr_handle* h, int first, int count, const fbpr_frame_input* fr, int chunk_frames, fbpr_result* out) {
    if (count > 0 && !out) return fbpr_fail_msg("null frames / results");
    const int ticket = fbpr_register_frames_begin(h, first, count, fr, chunk_frames);
    if (ticket < 0) return ticket;
    const int rc = fbpr_register_frames_end(h, ticket, out);
    return rc < 0 ? rc : 0;
}

int fbpr_extract_surrounding_keyframes(fbpr_handle* h, int slot, int K, const float* key_poses6,
                                       const float* corner_xyzi, const int32_t* corner_off,
                                       const float* surf_xyzi, const int32_t* surf_off,
                                       const float last_key_xyz[3], int mem) {
    return fbpr_extract_cloud(h, slot, K, key_poses6, nullptr, corner_xyzi, corner_off, surf_xyzi, surf_off, last_key_xyz, mem);
}

int fbpr_extract_cloud(fbpr_handle* h, int slot, int K, const float* key_poses6, const float* check_xyz,
                       const float* corner_xyzi, const int32_t* corner_off,
                       const float* surf_xyzi, const int32_t* surf_off,
                       const float last_key_xyz[3], int mem) {
    int rc = check_range(h, slot, 1); if (rc) return rc;
    if (h->kfCap <= 0) return fbpr_fail_msg("handle created with max_keyframe_points = 0");
    if (mem != FBPR_MEM_HOST) return fbpr_fail_msg("extract_surrounding_keyframes takes host buffers");
    if (K < 0) return fbpr_fail_msg("bad K");
    const int nc = K ? corner_off[K] : 0, ns = K ? surf_off[K] : 0;
    if (nc > h->kfCap || ns > h->kfCap) return fbpr_fail_msg("keyframe clouds exceed max_keyframe_points");
    cudaSetDevice(h->device);
    // staging: poses, offsets, raw keyframe clouds (freed after the launches complete)
    float* d_poses = nullptr; int* d_coff = nullptr; int* d_soff = nullptr; float4* d_cin = nullptr; float4* d_sin = nullptr; float* d_last = nullptr;
    FBPR_CUDA_OK(cudaMallocAsync(&d_poses, sizeof(float) * 6 * (K + 1), h->stream));
    FBPR_CUDA_OK(cudaMallocAsync(&d_coff, sizeof(int) * (K + 1), h->stream));
    FBPR_CUDA_OK(cudaMallocAsync(&d_soff, sizeof(int) * (K + 1), h->stream));
    FBPR_CUDA_OK(cudaMallocAsync(&d_cin, sizeof(float4) * (nc + 1), h->stream));
    FBPR_CUDA_OK(cudaMallocAsync(&d_sin, sizeof(float4) * (ns + 1), h->stream));
    FBPR_CUDA_OK(cudaMallocAsync(&d_last, sizeof(float) * 4, h->stream));
    float* d_check = nullptr;
    if (check_xyz && K) {
        FBPR_CUDA_OK(cudaMallocAsync(&d_check, sizeof(float) * 3 * K, h->stream));
        FBPR_CUDA_OK(cudaMemcpyAsync(d_check, check_xyz, sizeof(float) * 3 * K, cudaMemcpyHostToDevice, h->stream));
    }
    int* d_outoff = nullptr; float* d_T = nullptr;
    FBPR_CUDA_OK(cudaMallocAsync(&d_outoff, sizeof(int) * (K + 2), h->stream));
    FBPR_CUDA_OK(cudaMallocAsync(&d_T, sizeof(float) * 12 * (K + 1), h->stream));
    if (K) {
        FBPR_CUDA_OK(cudaMemcpyAsync(d_poses, key_poses6, sizeof(float) * 6 * K, cudaMemcpyHostToDevice, h->stream));
        FBPR_CUDA_OK(cudaMemcpyAsync(d_coff, corner_off, sizeof(int) * (K + 1), cudaMemcpyHostToDevice, h->stream));
        FBPR_CUDA_OK(cudaMemcpyAsync(d_soff, surf_off, sizeof(int) * (K + 1), cudaMemcpyHostToDevice, h->stream));
        if (nc) FBPR_CUDA_OK(cudaMemcpyAsync(d_cin, corner_xyzi, sizeof(float4) * nc, cudaMemcpyHostToDevice, h->stream));
        if (ns) FBPR_CUDA_OK(cudaMemcpyAsync(d_sin, surf_xyzi, sizeof(float4) * ns, cudaMemcpyHostToDevice, h->stream));
    }
    FBPR_CUDA_OK(cudaMemcpyAsync(d_last, last_key_xyz, sizeof(float) * 3, cudaMemcpyHostToDevice, h->stream));
    FBPR_CUDA_OK(cudaMemsetAsync(&h->meta[slot].mapTruncated, 0, sizeof(int), h->stream));
    rc = fbpr_launch_keyframe_transform(d_poses, K, d_cin, d_coff, h->kfCorner + (size_t)slot * h->kfCap, h->kfCount + 2 * slot,
                                        d_last, h->p.surroundingKeyframeSearchRadius, d_check, nc, d_outoff, d_T, h->stream, &h->launches);
    if (!rc) rc = fbpr_launch_keyframe_transform(d_poses, K, d_sin, d_soff, h->kfSurf + (size_t)slot * h->kfCap, h->kfCount + 2 * slot + 1,
                                                 d_last, h->p.surroundingKeyframeSearchRadius, d_check, ns, d_outoff, d_T, h->stream, &h->launches);
    if (!rc) rc = fbpr_launch_voxel(h->d_kfSegs + 2 * (size_t)slot, 2, h->kfCap, h->tilesCap, h->stream, &h->launches);
    cudaFreeAsync(d_poses, h->stream); cudaFreeAsync(d_coff, h->stream); cudaFreeAsync(d_soff, h->stream);
    cudaFreeAsync(d_cin, h->stream); cudaFreeAsync(d_sin, h->stream); cudaFreeAsync(d_last, h->stream);
    cudaFreeAsync(d_outoff, h->stream); cudaFreeAsync(d_T, h->stream);
    if (d_check) cudaFreeAsync(d_check, h->stream);
    return rc;
}

// ---- resident keyframe store --------------------------------------------------------------------
static int kfs_grow_bytes(fbpr_handle* h, void** p, size_t usedBytes, size_t newBytes) {
    void* q = nullptr;
    FBPR_CUDA_OK(cudaMalloc(&q, newBytes));
    FBPR_CUDA_OK(cudaMemsetAsync(q, 0, newBytes, h->stream));
    if (*p && usedBytes) FBPR_CUDA_OK(cudaMemcpyAsync(q, *p, usedBytes, cudaMemcpyDeviceToDevice, h->stream));
    FBPR_CUDA_OK(cudaStreamSynchronize(h->stream));
    if (*p) cudaFree(*p);
    *p = q;
    return 0;
}
#define kfs_grow(h, pp, used, cnt) kfs_grow_bytes((h), reinterpret_cast<void**>(pp), (used) * sizeof(**(pp)), (cnt) * sizeof(**(pp)))

// scratch of a selection over up to `cap` key poses
static int kfs_build_select(fbpr_handle* h, int cap) {
    auto& K = h->kfs;
    FBPR_CUDA_OK(cudaStreamSynchronize(h->stream));
    for (void* p : K.selAllocs) cudaFree(p);
    K.selAllocs.clear();
    auto A = [&](void** p, size_t bytes) { cudaError_t e = cudaMalloc(p, bytes); if (e == cudaSuccess) { cudaMemset(*p, 0, bytes); K.selAllocs.push_back(*p); } return e; };
    KfSelect q = {};
    q.cap = cap;
    FBPR_CUDA_OK(A((void**)&q.keys, sizeof(unsigned long long) * (size_t)next_pow2(cap)));
    FBPR_CUDA_OK(A((void**)&q.counters, 4 * sizeof(int)));
    FBPR_CUDA_OK(A((void**)&q.hitPts, sizeof(float4) * (size_t)cap));
    FBPR_CUDA_OK(A((void**)&q.list, sizeof(float4) * 2 * (size_t)cap));
    FBPR_CUDA_OK(A((void**)&q.selIdx, sizeof(int) * 2 * (size_t)cap));
    for (int k = 0; k < 2; k++) FBPR_CUDA_OK(A((void**)&q.outoff[k], sizeof(int) * (2 * (size_t)cap + 1)));
    FBPR_CUDA_OK(A((void**)&q.T, sizeof(float) * 12 * 2 * (size_t)cap));
    // downSizeFilterSurroundingKeyPoses (mapOptmization.h:887-888): hit poses -> the head of cloudToExtract
    VoxSeg s = {};
    const int tiles = (cap + fbpr_voxel_tile() - 1) / fbpr_voxel_tile() + 1;
    for (int b = 0; b < 2; b++) { FBPR_CUDA_OK(A((void**)&s.key[b], 4 * (size_t)cap)); FBPR_CUDA_OK(A((void**)&s.val[b], 4 * (size_t)cap)); }
    FBPR_CUDA_OK(A((void**)&s.tile_hist, 4 * (size_t)256 * tiles)); FBPR_CUDA_OK(A((void**)&s.bbox, 32));
    FBPR_CUDA_OK(A((void**)&s.run_tile, 4 * (size_t)(tiles + 1))); FBPR_CUDA_OK(A((void**)&s.desc, sizeof(VoxDesc)));
    FBPR_CUDA_OK(A((void**)&K.d_poseSeg, sizeof(VoxSeg)));
    s.in = q.hitPts; s.n_in = q.counters; s.out = q.list; s.n_out = q.counters + 1; s.cap = cap; s.out_cap = cap; s.leaf = K.leaf > 0.f ? K.leaf : 1.0f;
    FBPR_CUDA_OK(cudaMemcpy(K.d_poseSeg, &s, sizeof(s), cudaMemcpyHostToDevice));
    K.sel = q; K.poseTilesCap = tiles; K.leaf = s.leaf;
    return 0;
}

int fbpr_keyframes_clear(fbpr_handle* h) {
    if (!h) return fbpr_fail_msg("null handle");
    h->kfs.n = 0; h->kfs.poolUsed[0] = h->kfs.poolUsed[1] = 0;      // capacity is kept; offsets restart at 0 (off[k][0] is always 0)
    return 0;
}
int fbpr_keyframes_count(fbpr_handle* h) { return h ? h->kfs.n : fbpr_fail_msg("null handle"); }

int fbpr_keyframe_push(fbpr_handle* h, const float pose6[6], double time, const float* corner_xyzi, int n_corner,
                       const float* surf_xyzi, int n_surf, int mem) {
    if (!h) return fbpr_fail_msg("null handle");
    if (!pose6 || n_corner < 0 || n_surf < 0 || (n_corner && !corner_xyzi) || (n_surf && !surf_xyzi)) return fbpr_fail_msg("bad keyframe");
    cudaSetDevice(h->device);
    auto& K = h->kfs;
    if (K.n + 1 > K.poseCap) {
        const int cap = K.poseCap ? 2 * K.poseCap : 1024;
        int rc = kfs_grow(h, &K.pose6, 6 * (size_t)K.n, 6 * (size_t)cap); if (rc) return rc;
        rc = kfs_grow(h, &K.time, (size_t)K.n, (size_t)cap); if (rc) return rc;
        for (int k = 0; k < 2; k++) { rc = kfs_grow(h, &K.off[k], (size_t)K.n + 1, (size_t)cap + 1); if (rc) return rc; }
        rc = kfs_build_select(h, cap); if (rc) return rc;
        K.poseCap = cap;
    }
    const float* src[2] = { corner_xyzi, surf_xyzi }; const int len[2] = { n_corner, n_surf };
    for (int k = 0; k < 2; k++) {
        if (K.poolUsed[k] + len[k] > 0x7fffffffLL) return fbpr_fail_msg("keyframe store: more than 2^31 points of one kind");
        if (K.poolUsed[k] + len[k] > K.poolCap[k]) {
            long long cap = K.poolCap[k] ? 2 * K.poolCap[k] : (1LL << 20);
            while (cap < K.poolUsed[k] + len[k]) cap *= 2;
            int rc = kfs_grow(h, &K.pool[k], (size_t)K.poolUsed[k], (size_t)cap); if (rc) return rc;
            K.poolCap[k] = cap;
        }
        if (len[k]) FBPR_CUDA_OK(cudaMemcpyAsync(K.pool[k] + K.poolUsed[k], src[k], sizeof(float4) * (size_t)len[k], kind_in(mem), h->stream));
        K.poolUsed[k] += len[k];
        const int end = (int)K.poolUsed[k];
        FBPR_CUDA_OK(cudaMemcpyAsync(K.off[k] + K.n + 1, &end, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    }
    FBPR_CUDA_OK(cudaMemcpyAsync(K.pose6 + 6 * (size_t)K.n, pose6, 6 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    FBPR_CUDA_OK(cudaMemcpyAsync(K.time + K.n, &time, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    return K.n++;
}

int fbpr_keyframes_set_poses(fbpr_handle* h, int first, int count, const float* pose6) {
    if (!h) return fbpr_fail_msg("null handle");
    if (first < 0 || count < 0 || first + count > h->kfs.n || (count && !pose6)) return fbpr_fail_msg("keyframe range out of bounds");
    cudaSetDevice(h->device);
    if (count) FBPR_CUDA_OK(cudaMemcpyAsync(h->kfs.pose6 + 6 * (size_t)first, pose6, 6 * sizeof(float) * (size_t)count, cudaMemcpyHostToDevice, h->stream));
    return 0;
}

int fbpr_extract_surrounding_keyframes_resident(fbpr_handle* h, int slot, double timeLaserCloudInfoLast, float surroundingKeyframeDensity,
                                                int loopClosureEnableFlag, int surroundingKeyframeSize) {
    int rc = check_range(h, slot, 1); if (rc) return rc;
    if (h->kfCap <= 0) return fbpr_fail_msg("handle created with max_keyframe_points = 0");
    auto& K = h->kfs;
    if (K.n == 0) return 0;                                       // mapOptmization.h:966-967: nothing to extract, the local map is left alone
    if (!loopClosureEnableFlag && !(surroundingKeyframeDensity > 0.f)) return fbpr_fail_msg("surroundingKeyframeDensity must be positive");
    cudaSetDevice(h->device);
    if (!loopClosureEnableFlag && surroundingKeyframeDensity != K.leaf) {
        FBPR_CUDA_OK(cudaMemcpyAsync(reinterpret_cast<char*>(K.d_poseSeg) + offsetof(VoxSeg, leaf), &surroundingKeyframeDensity, sizeof(float), cudaMemcpyHostToDevice, h->stream));
        K.leaf = surroundingKeyframeDensity;
    }
    KfStoreView v = {};
    v.pose6 = K.pose6; v.time = K.time; v.n = K.n;
    for (int k = 0; k < 2; k++) { v.off[k] = K.off[k]; v.pool[k] = K.pool[k]; }
    FBPR_CUDA_OK(cudaMemsetAsync(&h->meta[slot].mapTruncated, 0, sizeof(int), h->stream));
    rc = fbpr_launch_keyframe_select(v, K.sel, K.d_poseSeg, K.poseTilesCap, timeLaserCloudInfoLast, h->p.surroundingKeyframeSearchRadius,
                                     surroundingKeyframeDensity, loopClosureEnableFlag != 0, surroundingKeyframeSize,
                                     h->kfCorner + (size_t)slot * h->kfCap, h->kfSurf + (size_t)slot * h->kfCap, h->kfCap,
                                     h->kfCount + 2 * slot, &h->meta[slot].mapTruncated, h->stream, &h->launches);
    if (!rc) rc = fbpr_launch_voxel(h->d_kfSegs + 2 * (size_t)slot, 2, h->kfCap, h->tilesCap, h->stream, &h->launches);
    return rc;
}

int fbpr_get_keyframe_selection(fbpr_handle* h, float* list_xyzi, int32_t* key_index, int cap) {
    if (!h) return fbpr_fail_msg("null handle");
    if (!h->kfs.sel.counters) return 0;
    cudaSetDevice(h->device);
    int c[4] = { 0, 0, 0, 0 };
    FBPR_CUDA_OK(cudaMemcpyAsync(c, h->kfs.sel.counters, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
    FBPR_CUDA_OK(cudaStreamSynchronize(h->stream));
    const int Kn = c[2], m = Kn < cap ? Kn : cap;
    if (m > 0 && list_xyzi) FBPR_CUDA_OK(cudaMemcpy(list_xyzi, h->kfs.sel.list, sizeof(float4) * (size_t)m, cudaMemcpyDeviceToHost));
    if (m > 0 && key_index) FBPR_CUDA_OK(cudaMemcpy(key_index, h->kfs.sel.selIdx, sizeof(int) * (size_t)m, cudaMemcpyDeviceToHost));
    return Kn;
}

static int upload_global(fbpr_handle* h, const float* corner_global, int nCg, const float* surf_global, int nSg, int mem) {
    if (nCg + nSg > h->regGlobalCap) {
        int cap = nCg + nSg + 1024;
        FBPR_CUDA_OK(cudaStreamSynchronize(h->stream));
        ALLOC(h->regGlobal, cap); h->regGlobalCap = cap;
    }
    if (nCg) FBPR_CUDA_OK(cudaMemcpyAsync(h->regGlobal, corner_global, sizeof(float4) * nCg, kind_in(mem), h->stream));
    if (nSg) FBPR_CUDA_OK(cudaMemcpyAsync(h->regGlobal + nCg, surf_global, sizeof(float4) * nSg, kind_in(mem), h->stream));
    return 0;
}

int fbpr_set_global_map(fbpr_handle* h, const float* corner, int nC, const float* surf, int nS, int mem) {
    if (!h) return fbpr_fail_msg("null handle");
    if (nC < 0 || nS < 0) return fbpr_fail_msg("bad sizes");
    cudaSetDevice(h->device);
    int rc = upload_global(h, corner, nC, surf, nS, mem); if (rc) return rc;
    h->globalCornerN = nC; h->globalSurfN = nS;
    return 0;
}

int fbpr_crop_local_maps(fbpr_handle* h, int first, int count) {
    int rc = check_range(h, first, count); if (rc) return rc;
    cudaSetDevice(h->device);
    return enqueue_crop_local_maps(h, first, count);
}

int fbpr_registration(fbpr_handle* h, int slot, const float* corner_global, int nCg, const float* surf_global, int nSg, int mem, float pose12[12]) {
    int rc = check_range(h, slot, 1); if (rc) return rc;
    if (nCg < 0 || nSg < 0 || !pose12) return fbpr_fail_msg("bad registration arguments");
    cudaSetDevice(h->device);
    const float4* d_c = reinterpret_cast<const float4*>(corner_global);
    const float4* d_s = reinterpret_cast<const float4*>(surf_global);
    if (!corner_global && !surf_global) {
        if (h->globalCornerN < 0) return fbpr_fail_msg("no global map: pass the maps or call fbpr_set_global_map first");
        nCg = h->globalCornerN; nSg = h->globalSurfN;
        d_c = h->regGlobal; d_s = h->regGlobal + nCg;
    } else if (mem == FBPR_MEM_HOST) {
        rc = upload_global(h, corner_global, nCg, surf_global, nSg, mem); if (rc) return rc;
        h->globalCornerN = -1;
        d_c = h->regGlobal; d_s = h->regGlobal + nCg;
    }
    const int need = nCg > nSg ? nCg : nSg;
    if (!h->regTile) ALLOC(h->regTile, 1 << 16);
    if (need / 2048 + 2 > (1 << 16)) return fbpr_fail_msg("global map too large for the CropBox scratch");
    FBPR_CUDA_OK(cudaMemcpyAsync(h->regPose, pose12, sizeof(float) * 12, cudaMemcpyHostToDevice, h->stream));
    // CropBox +-30/+-30/+-10 m around the guess (mapOptmization.h:284-304), order preserving
    // a cropped map larger than max_map_corner / max_map_surf is cut (the reference keeps every point): FBPR_FLAG_MAP_TRUNCATED
    FBPR_CUDA_OK(cudaMemsetAsync(&h->meta[slot].mapTruncated, 0, sizeof(int), h->stream));
    rc = fbpr_launch_crop_box(d_c, nCg, h->regPose, h->mapCorner + (size_t)slot * h->mapCornerCap, h->mapCornerCap, &h->meta[slot].n_map_corner,
                              &h->meta[slot].mapTruncated, h->regTile, h->stream, &h->launches); if (rc) return rc;
    rc = fbpr_launch_crop_box(d_s, nSg, h->regPose, h->mapSurf + (size_t)slot * h->mapSurfCap, h->mapSurfCap, &h->meta[slot].n_map_surf,
                              &h->meta[slot].mapTruncated, h->regTile, h->stream, &h->launches); if (rc) return rc;
    rc = fbpr_launch_pose_decompose(h->regPose, h->meta, slot, h->stream, &h->launches); if (rc) return rc;            // :309-310
    rc = enqueue_downsample(h, slot, 1); if (rc) return rc;                                    // :313
    rc = enqueue_scan2map(h, slot, 1); if (rc) return rc;                                      // :317
    rc = fbpr_launch_pose_compose(h->meta, slot, h->regPose, h->stream, &h->launches); if (rc) return rc;              // :326
    FBPR_CUDA_OK(cudaMemcpyAsync(pose12, h->regPose, sizeof(float) * 12, cudaMemcpyDeviceToHost, h->stream));
    FBPR_CUDA_OK(cudaStreamSynchronize(h->stream));
    FBPR_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- results ---------------------------------------------------------------------------------
int fbpr_get_results(fbpr_handle* h, int first, int count, fbpr_result* out, int mem) {
    int rc = check_range(h, first, count); if (rc) return rc;
    if (count == 0) return 0;
    cudaSetDevice(h->device);
    static_assert(offsetof(FrameMeta, iters) == offsetof(FrameMeta, pose) + 24 && offsetof(FrameMeta, flags) == offsetof(FrameMeta, pose) + 28, "result fields must be contiguous");
    FBPR_CUDA_OK(cudaMemcpy2DAsync(out, sizeof(fbpr_result), reinterpret_cast<char*>(h->meta + first) + offsetof(FrameMeta, pose), sizeof(FrameMeta),
                                   sizeof(fbpr_result), count, mem == FBPR_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
    if (mem == FBPR_MEM_HOST) { FBPR_CUDA_OK(cudaStreamSynchronize(h->stream)); FBPR_CUDA_OK(cudaGetLastError()); }
    return 0;
}
int fbpr_get_pose(fbpr_handle* h, int slot, float pose6[6], int32_t* iters, uint32_t* flags) {
    fbpr_result r;
    int rc = fbpr_get_results(h, slot, 1, &r, FBPR_MEM_HOST); if (rc) return rc;
    if (pose6) memcpy(pose6, r.pose, sizeof(r.pose));
    if (iters) *iters = r.iters;
    if (flags) *flags = r.flags;
    return 0;
}
int fbpr_get_counts(fbpr_handle* h, int slot, int32_t counts[8]) {
    int rc = check_range(h, slot, 1); if (rc) return rc;
    cudaSetDevice(h->device);
    FBPR_CUDA_OK(cudaMemcpyAsync(counts, h->meta + slot, 8 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    FBPR_CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}

int fbpr_enable_stage_timing(fbpr_handle* h, int on) { if (!h) return fbpr_fail_msg("null handle"); h->timing = on != 0; return 0; }

int fbpr_get_stage_ms(fbpr_handle* h, float ms[FBPR_STAGE_COUNT], int32_t calls[FBPR_STAGE_COUNT], int reset) {
    if (!h) return fbpr_fail_msg("null handle");
    cudaSetDevice(h->device);
    FBPR_CUDA_OK(cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < h->spansUsed; i++) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, h->spans[i].a, h->spans[i].b) == cudaSuccess) { h->stageMs[h->spans[i].stage] += t; h->stageCalls[h->spans[i].stage]++; }
    }
    h->spansUsed = 0;
    for (int s = 0; s < FBPR_STAGE_COUNT; s++) { if (ms) ms[s] = h->stageMs[s]; if (calls) calls[s] = h->stageCalls[s]; }
    if (reset) for (int s = 0; s < FBPR_STAGE_COUNT; s++) { h->stageMs[s] = 0.f; h->stageCalls[s] = 0; }
    return 0;
}

int fbpr_set_debug_iteration(fbpr_handle* h, int iter) {
    if (!h) return fbpr_fail_msg("null handle");
    cudaSetDevice(h->device);
    if (iter >= 0 && h->dbgSlots == 0) {
        const int S = h->F < 2 ? h->F : 2;
        ALLOC(h->knnC, (size_t)S * h->cornerCap * 5); ALLOC(h->d2C, (size_t)S * h->cornerCap * 5);
        ALLOC(h->coeffC, (size_t)S * h->cornerCap); ALLOC(h->flagC, (size_t)S * h->cornerCap);
        ALLOC(h->knnS, (size_t)S * h->P * 5); ALLOC(h->d2S, (size_t)S * h->P * 5);
        ALLOC(h->coeffS, (size_t)S * h->P); ALLOC(h->flagS, (size_t)S * h->P);
        ALLOC(h->dbgAtA, (size_t)S * 36); ALLOC(h->dbgAtB, (size_t)S * 6); ALLOC(h->dbgX, (size_t)S * 6);
        h->dbgSlots = S;
    }
    h->debugIter = iter;
    return 0;
}

int64_t fbpr_get_buffer(fbpr_handle* h, int slot, int which, void* dst, int64_t cap_bytes) {
    int rc = check_range(h, slot, 1); if (rc) return rc;
    cudaSetDevice(h->device);
    FrameMeta m;
    FBPR_CUDA_OK(cudaStreamSynchronize(h->stream));
    FBPR_CUDA_OK(cudaMemcpy(&m, h->meta + slot, sizeof(m), cudaMemcpyDeviceToHost));
    const size_t P = h->P, N = h->p.N_SCAN, CC = h->cornerCap;
    const void* src = nullptr; size_t bytes = 0;
    const bool dbg = slot < h->dbgSlots;
    switch (which) {
    case FBPR_BUF_START_RING: src = h->startRing + slot * N; bytes = N * 4; break;
    case FBPR_BUF_END_RING: src = h->endRing + slot * N; bytes = N * 4; break;
    case FBPR_BUF_COL_IND: src = h->colInd + slot * P; bytes = (size_t)m.n_valid * 4; break;
    case FBPR_BUF_RANGE: src = h->range + slot * P; bytes = (size_t)m.n_valid * 4; break;
    case FBPR_BUF_CLOUD: src = h->cloud + slot * P; bytes = (size_t)m.n_valid * 16; break;
    case FBPR_BUF_WINNER_RAW: src = h->winner + slot * P; bytes = (size_t)m.n_valid * 4; break;
    case FBPR_BUF_CURVATURE: src = h->curv + slot * P; bytes = (size_t)m.n_valid * 4; break;
    case FBPR_BUF_PICKED: src = h->picked + slot * P; bytes = (size_t)m.n_valid * 4; break;
    case FBPR_BUF_LABEL: src = h->label + slot * P; bytes = (size_t)m.n_valid * 4; break;
    case FBPR_BUF_CORNER: src = h->corner + slot * CC; bytes = (size_t)m.n_corner * 16; break;
    case FBPR_BUF_CORNER_INDEX: src = h->cornerIndex + slot * CC; bytes = (size_t)m.n_corner * 4; break;
    case FBPR_BUF_SURF: src = h->surf + slot * P; bytes = (size_t)m.n_surf * 16; break;
    case FBPR_BUF_RING_SURF_COUNT: src = h->ringSurf + slot * N; bytes = N * 4; break;
    case FBPR_BUF_RING_SURF_COUNT_DS: src = h->ringSurfDS + slot * N; bytes = N * 4; break;
    case FBPR_BUF_CORNER_DS: src = h->cornerDS + slot * CC; bytes = (size_t)m.n_corner_ds * 16; break;
    case FBPR_BUF_SURF_DS: src = h->surfDS + slot * P; bytes = (size_t)m.n_surf_ds * 16; break;
    case FBPR_BUF_MAP_CORNER: src = h->mapCorner + (size_t)slot * h->mapCornerCap; bytes = (size_t)m.n_map_corner * 16; break;
    case FBPR_BUF_MAP_SURF: src = h->mapSurf + (size_t)slot * h->mapSurfCap; bytes = (size_t)m.n_map_surf * 16; break;
    case FBPR_BUF_POSE_TRACE: src = h->poseTrace + (size_t)slot * FBPR_MAX_ITERS * 6; bytes = (size_t)FBPR_MAX_ITERS * 24; break;
    case FBPR_BUF_KNN_CORNER: if (dbg) { src = h->knnC + slot * CC * 5; bytes = (size_t)m.n_corner_ds * 20; } break;
    case FBPR_BUF_KNN_D2_CORNER: if (dbg) { src = h->d2C + slot * CC * 5; bytes = (size_t)m.n_corner_ds * 20; } break;
    case FBPR_BUF_COEFF_CORNER: if (dbg) { src = h->coeffC + slot * CC; bytes = (size_t)m.n_corner_ds * 16; } break;
    case FBPR_BUF_FLAG_CORNER: if (dbg) { src = h->flagC + slot * CC; bytes = (size_t)m.n_corner_ds; } break;
    case FBPR_BUF_KNN_SURF: if (dbg) { src = h->knnS + slot * P * 5; bytes = (size_t)m.n_surf_ds * 20; } break;
    case FBPR_BUF_KNN_D2_SURF: if (dbg) { src = h->d2S + slot * P * 5; bytes = (size_t)m.n_surf_ds * 20; } break;
    case FBPR_BUF_COEFF_SURF: if (dbg) { src = h->coeffS + slot * P; bytes = (size_t)m.n_surf_ds * 16; } break;
    case FBPR_BUF_FLAG_SURF: if (dbg) { src = h->flagS + slot * P; bytes = (size_t)m.n_surf_ds; } break;
    case FBPR_BUF_ATA: if (dbg) { src = h->dbgAtA + slot * 36; bytes = 144; } break;
    case FBPR_BUF_ATB: if (dbg) { src = h->dbgAtB + slot * 6; bytes = 24; } break;
    case FBPR_BUF_X: if (dbg) { src = h->dbgX + slot * 6; bytes = 24; } break;
    default: break;
    }
    if (!src) return fbpr_fail_msg("buffer not available (unknown id, or debug capture not enabled for this slot)");
    if ((int64_t)bytes > cap_bytes) return fbpr_fail_msg("destination too small");
    if (bytes) FBPR_CUDA_OK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return (int64_t)bytes;
}

int64_t fbpr_get_buffer_xyzi32(fbpr_handle* h, int slot, int which, void* dst, int64_t cap_bytes) {
    int rc = check_range(h, slot, 1); if (rc) return rc;
    cudaSetDevice(h->device);
    FrameMeta m;
    FBPR_CUDA_OK(cudaStreamSynchronize(h->stream));
    FBPR_CUDA_OK(cudaMemcpy(&m, h->meta + slot, sizeof(m), cudaMemcpyDeviceToHost));
    const size_t P = h->P, CC = h->cornerCap;
    const float4* src = nullptr; int n = 0;
    switch (which) {
    case FBPR_BUF_CLOUD: src = h->cloud + slot * P; n = m.n_valid; break;
    case FBPR_BUF_CORNER: src = h->corner + slot * CC; n = m.n_corner; break;
    case FBPR_BUF_SURF: src = h->surf + slot * P; n = m.n_surf; break;
    case FBPR_BUF_CORNER_DS: src = h->cornerDS + slot * CC; n = m.n_corner_ds; break;
    case FBPR_BUF_SURF_DS: src = h->surfDS + slot * P; n = m.n_surf_ds; break;
    case FBPR_BUF_MAP_CORNER: src = h->mapCorner + (size_t)slot * h->mapCornerCap; n = m.n_map_corner; break;
    case FBPR_BUF_MAP_SURF: src = h->mapSurf + (size_t)slot * h->mapSurfCap; n = m.n_map_surf; break;
    default: return fbpr_fail_msg("not a point-cloud buffer");
    }
    const size_t bytes = (size_t)n * 32;
    if ((int64_t)bytes > cap_bytes) return fbpr_fail_msg("destination too small");
    if (n == 0) return 0;
    rc = wire_stage(h, bytes); if (rc) return rc;
    rc = fbpr_launch_xyzi_repack(src, n, reinterpret_cast<float4*>(h->wireStage), 1, h->stream, &h->launches); if (rc) return rc;
    FBPR_CUDA_OK(cudaMemcpyAsync(dst, h->wireStage, bytes, cudaMemcpyDeviceToHost, h->stream));
    FBPR_CUDA_OK(cudaStreamSynchronize(h->stream));
    return (int64_t)bytes;
}

// ---- stand-alone VoxelGrid / k-NN ---------------------------------------------------------------
int fbpr_voxel_grid(fbpr_handle* h, const float* xyzi, int n, float leaf, float* out_xyzi, int32_t* point_keys, int32_t* out_keys, int mem) {
    if (!h) return fbpr_fail_msg("null handle");
    if (n < 0) return fbpr_fail_msg("bad n");
    cudaSetDevice(h->device);
    if (n + 1 > h->soloVoxCap) {
        FBPR_CUDA_OK(cudaStreamSynchronize(h->stream));
        for (void* p : h->soloVoxAllocs) cudaFree(p);
        h->soloVoxAllocs.clear();
        int cap = n + 1024;
        auto A = [&](void** p, size_t bytes) { cudaError_t e = cudaMalloc(p, bytes); if (e == cudaSuccess) { cudaMemset(*p, 0, bytes); h->soloVoxAllocs.push_back(*p); } return e; };
        VoxSeg s = {};
        int tiles = (cap + fbpr_voxel_tile() - 1) / fbpr_voxel_tile() + 1;
        FBPR_CUDA_OK(A((void**)&h->soloIn, sizeof(float4) * cap)); FBPR_CUDA_OK(A((void**)&h->soloOut, sizeof(float4) * cap));
        FBPR_CUDA_OK(A((void**)&h->soloN, 16)); FBPR_CUDA_OK(A((void**)&h->soloNout, 16));
        FBPR_CUDA_OK(A((void**)&h->soloPK, sizeof(int) * cap)); FBPR_CUDA_OK(A((void**)&h->soloOK, sizeof(int) * cap));
        for (int b = 0; b < 2; b++) { FBPR_CUDA_OK(A((void**)&s.key[b], 4 * (size_t)cap)); FBPR_CUDA_OK(A((void**)&s.val[b], 4 * (size_t)cap)); }
        FBPR_CUDA_OK(A((void**)&s.tile_hist, 4 * (size_t)256 * tiles)); FBPR_CUDA_OK(A((void**)&s.bbox, 32));
        FBPR_CUDA_OK(A((void**)&s.run_tile, 4 * (size_t)(tiles + 1))); FBPR_CUDA_OK(A((void**)&s.desc, sizeof(VoxDesc)));
        FBPR_CUDA_OK(A((void**)&h->d_soloVox, sizeof(VoxSeg)));
        s.in = h->soloIn; s.n_in = h->soloN; s.out = h->soloOut; s.n_out = h->soloNout; s.cap = cap; s.out_cap = cap;
        s.point_keys = h->soloPK; s.out_keys = h->soloOK; s.leaf = leaf;
        FBPR_CUDA_OK(cudaMemcpy(h->d_soloVox, &s, sizeof(s), cudaMemcpyHostToDevice));
        h->soloVoxCap = cap;
    }
    // leaf may change between calls: patch it in the descriptor
    FBPR_CUDA_OK(cudaMemcpyAsync(reinterpret_cast<char*>(h->d_soloVox) + offsetof(VoxSeg, leaf), &leaf, sizeof(float), cudaMemcpyHostToDevice, h->stream));
    FBPR_CUDA_OK(cudaMemcpyAsync(h->soloN, &n, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (n) FBPR_CUDA_OK(cudaMemcpyAsync(h->soloIn, xyzi, sizeof(float4) * (size_t)n, kind_in(mem), h->stream));
    int tiles_cap = (h->soloVoxCap + fbpr_voxel_tile() - 1) / fbpr_voxel_tile() + 1;
    { int rc = fbpr_launch_voxel(h->d_soloVox, 1, n > 0 ? n : 1, tiles_cap, h->stream, &h->launches); if (rc) return rc; }
    int m = 0;
    FBPR_CUDA_OK(cudaMemcpyAsync(&m, h->soloNout, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    FBPR_CUDA_OK(cudaStreamSynchronize(h->stream));
    FBPR_CUDA_OK(cudaGetLastError());
    cudaMemcpyKind ko = mem == FBPR_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (m) FBPR_CUDA_OK(cudaMemcpy(out_xyzi, h->soloOut, sizeof(float4) * (size_t)m, ko));
    if (point_keys && n) FBPR_CUDA_OK(cudaMemcpy(point_keys, h->soloPK, sizeof(int) * (size_t)n, ko));
    if (out_keys && m) FBPR_CUDA_OK(cudaMemcpy(out_keys, h->soloOK, sizeof(int) * (size_t)m, ko));
    return m;
}

int fbpr_knn5_first_radius(fbpr_handle* h, int cells) {
    if (!h) return fbpr_fail_msg("null handle");
    if (cells < 1 || cells > 64) return fbpr_fail_msg("first radius must be 1..64 cells");
    h->knnRad0 = cells;
    return 0;
}

int fbpr_knn5(fbpr_handle* h, const float* map_xyzi, int n_map, float cell, const float* q_xyz, int nq, int32_t* idx, float* d2, int mem) {
    if (!h) return fbpr_fail_msg("null handle");
    if (n_map < 0 || nq < 0) return fbpr_fail_msg("bad sizes");
    if (mem != FBPR_MEM_HOST) return fbpr_fail_msg("fbpr_knn5 takes host buffers");
    cudaSetDevice(h->device);
    const int cells = 1 << 21;
    if (n_map + 1 > h->soloGridCap || cell != h->soloGridCell) {
        FBPR_CUDA_OK(cudaStreamSynchronize(h->stream));
        for (void* p : h->soloGridAllocs) cudaFree(p);
        h->soloGridAllocs.clear();
        int cap = n_map + 1024;
        auto A = [&](void** p, size_t bytes) { cudaError_t e = cudaMalloc(p, bytes); if (e == cudaSuccess) { cudaMemset(*p, 0, bytes); h->soloGridAllocs.push_back(*p); } return e; };
        GridSeg g = {};
        FBPR_CUDA_OK(A((void**)&h->soloMap, sizeof(float4) * cap)); FBPR_CUDA_OK(A((void**)&h->soloMapN, 16));
        FBPR_CUDA_OK(A((void**)&g.sorted, sizeof(float4) * cap)); FBPR_CUDA_OK(A((void**)&g.cell_start, 4 * (size_t)(cells + 2)));
        FBPR_CUDA_OK(A((void**)&g.cell_of, 4 * (size_t)cap));
        FBPR_CUDA_OK(A((void**)&g.tile_sum, 4 * (size_t)(cells / 4096 + 4))); FBPR_CUDA_OK(A((void**)&g.bbox, 32)); FBPR_CUDA_OK(A((void**)&g.desc, sizeof(GridDesc)));
        FBPR_CUDA_OK(A((void**)&h->d_soloGrid, sizeof(GridSeg)));
        g.pts = h->soloMap; g.n = h->soloMapN; g.cap = cap; g.cells_cap = cells; g.h0 = cell;
        FBPR_CUDA_OK(cudaMemcpy(h->d_soloGrid, &g, sizeof(g), cudaMemcpyHostToDevice));
        h->soloGridCap = cap; h->soloGridCell = cell;
    }
    FBPR_CUDA_OK(cudaMemcpyAsync(h->soloMapN, &n_map, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (n_map) FBPR_CUDA_OK(cudaMemcpyAsync(h->soloMap, map_xyzi, sizeof(float4) * (size_t)n_map, cudaMemcpyHostToDevice, h->stream));
    { int rc = fbpr_launch_grid_build(h->d_soloGrid, 1, n_map > 0 ? n_map : 1, cells, h->stream, &h->launches); if (rc) return rc; }
    float* d_q = nullptr; int* d_idx = nullptr; float* d_d2 = nullptr;
    FBPR_CUDA_OK(cudaMallocAsync(&d_q, sizeof(float) * 3 * (size_t)(nq + 1), h->stream));
    FBPR_CUDA_OK(cudaMallocAsync(&d_idx, sizeof(int) * 5 * (size_t)(nq + 1), h->stream));
    FBPR_CUDA_OK(cudaMallocAsync(&d_d2, sizeof(float) * 5 * (size_t)(nq + 1), h->stream));
    if (nq) FBPR_CUDA_OK(cudaMemcpyAsync(d_q, q_xyz, sizeof(float) * 3 * (size_t)nq, cudaMemcpyHostToDevice, h->stream));
    { int rc = fbpr_launch_knn5(h->d_soloGrid, d_q, nq, d_idx, d_d2, h->knnRad0, h->stream, &h->launches); if (rc) return rc; }
    if (nq) {
        FBPR_CUDA_OK(cudaMemcpyAsync(idx, d_idx, sizeof(int) * 5 * (size_t)nq, cudaMemcpyDeviceToHost, h->stream));
        FBPR_CUDA_OK(cudaMemcpyAsync(d2, d_d2, sizeof(float) * 5 * (size_t)nq, cudaMemcpyDeviceToHost, h->stream));
    }
    cudaFreeAsync(d_q, h->stream); cudaFreeAsync(d_idx, h->stream); cudaFreeAsync(d_d2, h->stream);
    FBPR_CUDA_OK(cudaStreamSynchronize(h->stream));
    FBPR_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
