// projection.cu -- range-image projection, deskew and ring-major compaction.
//
// Replaces ImageProjection::projectPointCloud (imageProjection.cpp:583-640) with deskewPoint
// (:545-580) / findRotation (:494-526), and cloudExtraction (:642-670), producing the
// cloud_info record (msg/cloud_info.msg) the feature front-end consumes.
//
// The reference walks the raw cloud serially and lets the FIRST point that lands in a pixel
// win (:623).  Here: one atomicMin of the raw index per pixel (pass 1), then one CTA per ring
// resolves the winners, deskews them and stream-compacts the ring in column order (pass 2/3).
// transStartInverse comes from the first raw point that passes all gates (firstPointFlag,
// :565-569) = the minimum winning raw index, found with one more atomicMin.
// HBM-bound: 24 B read per raw point, 24 B written per valid point (SURVEY.md 8(d) B_proj).
#include "internal.cuh"


namespace {

constexpr int TPB = 256;
constexpr int EMPTY = 0x7fffffff;

__global__ void proj_clear(ProjArgs a) {
    const int slot = a.first + blockIdx.y;
    int* pix = a.pix + (size_t)slot * a.P;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.P; i += gridDim.x * blockDim.x) pix[i] = EMPTY;
    if (blockIdx.x == 0 && threadIdx.x == 0) a.meta[slot].first_valid_raw = EMPTY;
}

// imageProjection.cpp:605-616 with the f32/f64 promotions of the original expression
__device__ inline int column_of(float x, float y, int H) {
    float at = (float)atan2((double)x, (double)y);
    float horizonAngle = (float)((double)(at * 180) / 3.14159265358979323846);
    float ang_res_x = (float)(360.0 / (double)(float)H);
    int col = (int)(-round(((double)horizonAngle - 90.0) / (double)ang_res_x) + (double)(H / 2));
    if (col >= H) col -= H;
    return col;
}

__global__ void __launch_bounds__(TPB) proj_scatter(ProjArgs a) {
    const int slot = a.first + blockIdx.y;
    FrameMeta& M = a.meta[slot];
    const int n = min(M.n_raw, a.rawCap);
    const fbpr_raw_point* raw = a.raw + (size_t)slot * a.rawCap;
    int* pix = a.pix + (size_t)slot * a.P;
    int firstv = EMPTY;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const fbpr_raw_point p = raw[i];
        if (p.ring < 0 || p.ring >= a.N_SCAN) continue;
        int col = column_of(p.x, p.y, a.H);
        if (col < 0 || col >= a.H) continue;
        float rg = sqrtf(p.x * p.x + p.y * p.y + p.z * p.z);
        if ((double)rg < 1.0) continue;
        if (!isfinite(rg)) continue;            // the reference refuses non-dense clouds (imageProjection.cpp:250-254); here their NaN / Inf points are dropped
        atomicMin(&pix[p.ring * a.H + col], i);
        firstv = min(firstv, i);
    }
    for (int o = 16; o; o >>= 1) firstv = min(firstv, __shfl_xor_sync(0xffffffffu, firstv, o));
    if ((threadIdx.x & 31) == 0 && firstv != EMPTY) atomicMin(&M.first_valid_raw, firstv);
}

__global__ void __launch_bounds__(TPB) proj_ring_count(ProjArgs a) {
    const int slot = a.first + blockIdx.y, ring = blockIdx.x;
    const int* pix = a.pix + (size_t)slot * a.P + (size_t)ring * a.H;
    int c = 0;
    for (int j = threadIdx.x; j < a.H; j += TPB) c += pix[j] != EMPTY;
    __shared__ int ws[TPB / 32];
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int k = 0; k < TPB / 32; k++) t += ws[k]; a.ringCount[slot * a.N_SCAN + ring] = t; }
}

// findRotation (imageProjection.cpp:494-526)
__device__ inline void find_rotation(const double* imuTime, const double* rX, const double* rY, const double* rZ,
                                     int imuPointerCur, double pointTime, float& rx, float& ry, float& rz, bool sorted = false) {
    int front = 0;
    if (sorted) {                                            // ascending stamps: the first entry later than pointTime, by bisection
        int lo = 0, hi = imuPointerCur;                      // answer in [0, imuPointerCur]
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (pointTime < imuTime[mid]) hi = mid; else lo = mid + 1; }
        front = lo;
    } else {
        while (front < imuPointerCur) { if (pointTime < imuTime[front]) break; ++front; }
    }
    if (pointTime > imuTime[front] || front == 0) {
        rx = (float)rX[front]; ry = (float)rY[front]; rz = (float)rZ[front];
    } else {
        int back = front - 1;
        double ratioFront = (pointTime - imuTime[back]) / (imuTime[front] - imuTime[back]);
        double ratioBack = (imuTime[front] - pointTime) / (imuTime[front] - imuTime[back]);
        rx = (float)(rX[front] * ratioFront + rX[back] * ratioBack);
        ry = (float)(rY[front] * ratioFront + rY[back] * ratioBack);
        rz = (float)(rZ[front] * ratioFront + rZ[back] * ratioBack);
    }
}

__device__ inline void affine_inverse(const float T[12], float Ti[12]) {
    const float a = T[0], b = T[1], c = T[2], d = T[4], e = T[5], f = T[6], g = T[8], h = T[9], i = T[10];
    float c00 = e * i - f * h, c01 = f * g - d * i, c02 = d * h - e * g;
    float det = a * c00 + b * c01 + c * c02;
    float inv = 1.0f / det;
    float M[9];
    M[0] = c00 * inv; M[1] = (c * h - b * i) * inv; M[2] = (b * f - c * e) * inv;
    M[3] = c01 * inv; M[4] = (a * i - c * g) * inv; M[5] = (c * d - a * f) * inv;
    M[6] = c02 * inv; M[7] = (b * g - a * h) * inv; M[8] = (a * e - b * d) * inv;
    for (int r = 0; r < 3; r++) {
        Ti[4 * r] = M[3 * r]; Ti[4 * r + 1] = M[3 * r + 1]; Ti[4 * r + 2] = M[3 * r + 2];
        Ti[4 * r + 3] = -(M[3 * r] * T[3] + M[3 * r + 1] * T[7] + M[3 * r + 2] * T[11]);
    }
}
__device__ inline void affine_mul(const float A[12], const float B[12], float C[12]) {
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) C[4 * r + c] = A[4 * r] * B[c] + A[4 * r + 1] * B[4 + c] + A[4 * r + 2] * B[8 + c];
        C[4 * r + 3] = A[4 * r] * B[3] + A[4 * r + 1] * B[7] + A[4 * r + 2] * B[11] + A[4 * r + 3];
    }
}

__global__ void __launch_bounds__(TPB) proj_compact(ProjArgs a) {
    const int slot = a.first + blockIdx.y, ring = blockIdx.x;
    FrameMeta& M = a.meta[slot];
    const fbpr_raw_point* raw = a.raw + (size_t)slot * a.rawCap;
    const int* pix = a.pix + (size_t)slot * a.P + (size_t)ring * a.H;
    const double* imuTime = a.imuTime + (size_t)slot * FBPR_IMU_CAP;
    const double* rX = a.imuRotX + (size_t)slot * FBPR_IMU_CAP;
    const double* rY = a.imuRotY + (size_t)slot * FBPR_IMU_CAP;
    const double* rZ = a.imuRotZ + (size_t)slot * FBPR_IMU_CAP;
    __shared__ int s_base, s_total;
    __shared__ float s_Tinv[12];
    __shared__ int ws[TPB / 32];
    const bool deskew = !(M.deskewFlag == -1 || M.imuAvailable == 0);
    // the frame's IMU ramp, staged once per CTA; when its stamps ascend (they do for a real IMU queue) findRotation's linear
    // scan is replaced by a bisection with the same answer
    __shared__ double s_imu[4][FBPR_IMU_CAP];
    bool sortedRamp = false;
    if (deskew) {
        const int np = min(max(M.imuPointerCur, 0), FBPR_IMU_CAP - 1) + 1;
        int bad = 0;
        for (int i = threadIdx.x; i < np; i += TPB) {
            const double t = imuTime[i];
            s_imu[0][i] = t; s_imu[1][i] = rX[i]; s_imu[2][i] = rY[i]; s_imu[3][i] = rZ[i];
            if (i + 1 < np && !(t <= imuTime[i + 1])) bad = 1;
        }
        sortedRamp = !__syncthreads_or(bad);
        imuTime = s_imu[0]; rX = s_imu[1]; rY = s_imu[2]; rZ = s_imu[3];
    }
    if (threadIdx.x < 32) {
        // ring base = counts of all previous rings; total = all rings
        int b = 0, t = 0;
        for (int r = threadIdx.x; r < a.N_SCAN; r += 32) { int c = a.ringCount[slot * a.N_SCAN + r]; t += c; if (r < ring) b += c; }
        for (int o = 16; o; o >>= 1) { b += __shfl_xor_sync(0xffffffffu, b, o); t += __shfl_xor_sync(0xffffffffu, t, o); }
        if (threadIdx.x == 0) {
            s_base = b; s_total = t;
            if (deskew && M.first_valid_raw != EMPTY) {
                float rx, ry, rz, T[12];
                find_rotation(imuTime, rX, rY, rZ, M.imuPointerCur, M.timeScanCur + (double)raw[M.first_valid_raw].time, rx, ry, rz);
                get_transformation(0.f, 0.f, 0.f, rx, ry, rz, T);
                affine_inverse(T, s_Tinv);
            }
        }
    }
    __syncthreads();
    int run = s_base;
    int* colInd = a.colInd + (size_t)slot * a.P;
    float* range = a.range + (size_t)slot * a.P;
    float4* cloud = a.cloud + (size_t)slot * a.P;
    int* winner = a.winner + (size_t)slot * a.P;
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int j0 = 0; j0 < a.H; j0 += TPB) {
        int j = j0 + threadIdx.x;
        int wi = j < a.H ? pix[j] : EMPTY;
        bool v = wi != EMPTY;
        unsigned bal = __ballot_sync(0xffffffffu, v);
        int wr = __popc(bal & ((1u << l) - 1));
        if (l == 0) ws[w] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int q = 0; q < TPB / 32; q++) { int c = ws[q]; if (q < w) woff += c; tot += c; }
        if (v) {
            const fbpr_raw_point p = raw[wi];
            float4 out = make_float4(p.x, p.y, p.z, p.intensity);
            if (deskew) {
                float rx, ry, rz, T[12], Bt[12];
                find_rotation(imuTime, rX, rY, rZ, M.imuPointerCur, M.timeScanCur + (double)p.time, rx, ry, rz, sortedRamp);
                get_transformation(0.f, 0.f, 0.f, rx, ry, rz, T);
                affine_mul(s_Tinv, T, Bt);
                out.x = Bt[0] * p.x + Bt[1] * p.y + Bt[2] * p.z + Bt[3];
                out.y = Bt[4] * p.x + Bt[5] * p.y + Bt[6] * p.z + Bt[7];
                out.z = Bt[8] * p.x + Bt[9] * p.y + Bt[10] * p.z + Bt[11];
            }
            int o = run + woff + wr;
            colInd[o] = j;
            range[o] = sqrtf(p.x * p.x + p.y * p.y + p.z * p.z);
            cloud[o] = out;
            winner[o] = wi;
        }
        run += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        a.startRing[slot * a.N_SCAN + ring] = s_base - 1 + 5;        // imageProjection.cpp:650
        a.endRing[slot * a.N_SCAN + ring] = run - 1 - 5;             // :668
        if (ring == 0) M.n_valid = s_total;
    }
}

}  // namespace

int fbpr_launch_projection(const ProjArgs& a, int count, cudaStream_t st, long long* launches) {
    if (count <= 0) return 0;
    int pb = (a.P + TPB * 4 - 1) / (TPB * 4);
    proj_clear<<<dim3(pb, count), TPB, 0, st>>>(a);
    int rb = (a.rawCap + TPB * 4 - 1) / (TPB * 4);
    proj_scatter<<<dim3(rb, count), TPB, 0, st>>>(a);
    proj_ring_count<<<dim3(a.N_SCAN, count), TPB, 0, st>>>(a);
    proj_compact<<<dim3(a.N_SCAN, count), TPB, 0, st>>>(a);
    if (launches) *launches += 4;
    return fbpr_launch_ok("projection (proj_clear / proj_scatter / proj_ring_count / proj_compact)");
}
