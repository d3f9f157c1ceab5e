// selftest.cu -- runs the device small-matrix routines (smallmat.cuh) on caller-supplied matrices, one thread per problem.
// It exists so that the parity tests can compare the DEVICE code of cv::eigen / cv::solve(DECOMP_QR) / cv::Mat::inv /
// Eigen::ColPivHouseholderQR<5x3> directly with the committed cv2 vectors (tests/golden/smallmat_cv2.npz) instead of only
// through the whole LM loop.  Reference call sites: mapOptmization.h:1060, :1169, :1343, :1353, :1370.
#include "internal.cuh"
#include "smallmat.cuh"

namespace {

__global__ void selftest_smallmat(int which, const float* __restrict__ in, int n, float* __restrict__ out) {
    if (which == FBPR_SELFTEST_QR6_WARP) {           // one WARP per problem (64 threads per CTA = 2 problems); in / out as QR6
        __shared__ float sA[2][36], sb[2][6];
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int p = blockIdx.x * 2 + w; p < n + (n & 1); p += gridDim.x * 2) {       // both warps of a CTA run the same trip count
            const bool live = p < n;
            for (int k = lane; k < 36; k += 32) sA[w][k] = live ? in[42 * p + k] : (k % 7 == 0 ? 1.f : 0.f);
            if (lane < 6) sb[w][lane] = live ? in[42 * p + 36 + lane] : 0.f;
            __syncwarp();
            float x[6];
            dev_qr_solve6_warp(sA[w], sb[w], x);
            if (live && lane == 0) for (int k = 0; k < 6; k++) out[6 * p + k] = x[k];
            __syncwarp();
        }
        return;
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (which == FBPR_SELFTEST_JACOBI3) {            // in: full 3x3 (row-major), out: W[3] then V[9] (rows = eigenvectors)
        const float* A = in + 9 * i;
        float W[3], V[9];
        dev_jacobi3(A[0], A[1], A[2], A[4], A[5], A[8], W, V);
        for (int k = 0; k < 3; k++) out[12 * i + k] = W[k];
        for (int k = 0; k < 9; k++) out[12 * i + 3 + k] = V[k];
    } else if (which == FBPR_SELFTEST_JACOBI6) {     // in: 6x6, out: W[6] then V[36]
        float A[36], W[6], V[36];
        for (int k = 0; k < 36; k++) A[k] = in[36 * i + k];
        dev_jacobi<6>(A, W, V);
        for (int k = 0; k < 6; k++) out[42 * i + k] = W[k];
        for (int k = 0; k < 36; k++) out[42 * i + 6 + k] = V[k];
    } else if (which == FBPR_SELFTEST_QR6) {         // in: 6x6 then b[6], out: x[6]
        float A[36], b[6], x[6];
        for (int k = 0; k < 36; k++) A[k] = in[42 * i + k];
        for (int k = 0; k < 6; k++) b[k] = in[42 * i + 36 + k];
        dev_qr_solve6(A, b, x);
        for (int k = 0; k < 6; k++) out[6 * i + k] = x[k];
    } else if (which == FBPR_SELFTEST_LU6) {         // in: 6x6, out: inverse 6x6
        float A[36], B[36];
        for (int k = 0; k < 36; k++) A[k] = in[36 * i + k];
        dev_lu_invert6(A, B);
        for (int k = 0; k < 36; k++) out[36 * i + k] = B[k];
    } else if (which == FBPR_SELFTEST_PLANE5X3) {    // in: 5x3 (row-major), out: x[3] of A x = -1
        float A[15], x[3];
        for (int k = 0; k < 15; k++) A[k] = in[15 * i + k];
        dev_plane_solve(A, x);
        for (int k = 0; k < 3; k++) out[3 * i + k] = x[k];
    } else if (which == FBPR_SELFTEST_SINCOS) {      // in: angle, out: sin, cos (sincosf_c: the polynomial kernels inside |x| <= pi/4, sincos(double) outside)
        float sv, cv;
        sincosf_c(in[i], sv, cv);
        out[2 * i] = sv; out[2 * i + 1] = cv;
    } else if (which == FBPR_SELFTEST_NOT_DEGENERATE) {   // in: 6x6, out: 1.0 when the LDL^T certificate holds
        out[i] = dev_surely_not_degenerate(in + 36 * i) ? 1.f : 0.f;
    }
}

}  // namespace

extern "C" int fbpr_selftest_smallmat(fbpr_handle* h, int which, const float* in, int n, float* out) {
    static const int in_w[] = { 9, 36, 42, 36, 15, 36, 42, 1 }, out_w[] = { 12, 42, 6, 36, 3, 1, 6, 2 };
    if (!h) return fbpr_fail_msg("null handle");
    if (which < 0 || which > FBPR_SELFTEST_SINCOS || n < 0 || (n > 0 && (!in || !out))) return fbpr_fail_msg("bad selftest arguments");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)fbpr_stream(h);
    float *d_in = nullptr, *d_out = nullptr;
    FBPR_CUDA_OK(cudaMallocAsync(&d_in, sizeof(float) * (size_t)n * in_w[which], st));
    FBPR_CUDA_OK(cudaMallocAsync(&d_out, sizeof(float) * (size_t)n * out_w[which], st));
    FBPR_CUDA_OK(cudaMemcpyAsync(d_in, in, sizeof(float) * (size_t)n * in_w[which], cudaMemcpyHostToDevice, st));
    selftest_smallmat<<<(n + 63) / 64, 64, 0, st>>>(which, d_in, n, d_out);
    FBPR_CUDA_OK(cudaGetLastError());
    FBPR_CUDA_OK(cudaMemcpyAsync(out, d_out, sizeof(float) * (size_t)n * out_w[which], cudaMemcpyDeviceToHost, st));
    cudaFreeAsync(d_in, st); cudaFreeAsync(d_out, st);
    FBPR_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}
