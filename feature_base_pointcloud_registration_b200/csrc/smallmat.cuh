// smallmat.cuh -- register/local-memory small-matrix routines used by the association and
// LM kernels.  They follow, operation for operation (f32, one rounding per op, no FMA), the
// third-party routines the reference calls:
//   dev_jacobi<N>      cv::eigen            (OpenCV JacobiImpl_)   mapOptmization.h:1060 (N=3), :1353 (N=6)
//   dev_qr_solve6      cv::solve(DECOMP_QR) (OpenCV hal::QR32f)    mapOptmization.h:1343
//   dev_lu_invert6     cv::Mat::inv()       (OpenCV hal::LU32f)    mapOptmization.h:1370
//   dev_plane_solve    Eigen::ColPivHouseholderQR<5x3>::solve      mapOptmization.h:1169
// SURVEY.md Appendix A / B-3 give the algorithms; tests compare against the CPU oracle, whose
// copies are pinned bit-for-bit to cv2.
#pragma once
#include <cfloat>

__device__ __forceinline__ float dev_hypot(float a, float b) {
    a = fabsf(a); b = fabsf(b);
    if (a > b) { b /= a; return a * sqrtf(1 + b * b); }
    if (b > 0) { a /= b; return b * sqrtf(1 + a * a); }
    return 0;
}

template <int N>
__device__ inline void dev_jacobi(float* A, float* W, float* V) {
    int indR[N], indC[N];
    for (int i = 0; i < N; i++) { for (int j = 0; j < N; j++) V[i * N + j] = 0.f; V[i * N + i] = 1.f; }
    float mv = 0.f;
    for (int k = 0; k < N; k++) {
        W[k] = A[(N + 1) * k];
        if (k < N - 1) {
            int m = k + 1; mv = fabsf(A[N * k + m]);
            for (int i = k + 2; i < N; i++) { float val = fabsf(A[N * k + i]); if (mv < val) { mv = val; m = i; } }
            indR[k] = m;
        }
        if (k > 0) {
            int m = 0; mv = fabsf(A[k]);
            for (int i = 1; i < k; i++) { float val = fabsf(A[N * i + k]); if (mv < val) { mv = val; m = i; } }
            indC[k] = m;
        }
    }
    for (int iters = 0; iters < N * N * 30; iters++) {
        int k = 0; mv = fabsf(A[indR[0]]);
        for (int i = 1; i < N - 1; i++) { float val = fabsf(A[N * i + indR[i]]); if (mv < val) { mv = val; k = i; } }
        int l = indR[k];
        for (int i = 1; i < N; i++) { float val = fabsf(A[N * indC[i] + i]); if (mv < val) { mv = val; k = indC[i]; l = i; } }
        float p = A[N * k + l];
        if (fabsf(p) <= FLT_EPSILON) break;
        float y = (W[l] - W[k]) * 0.5f;
        float t = fabsf(y) + dev_hypot(p, y);
        float s = dev_hypot(p, t);
        float c = t / s;
        s = p / s; t = (p / t) * p;
        if (y < 0) { s = -s; t = -t; }
        A[N * k + l] = 0;
        W[k] -= t; W[l] += t;
        float a0, b0;
#define FBPR_ROT(v0, v1) { a0 = v0; b0 = v1; v0 = a0 * c - b0 * s; v1 = a0 * s + b0 * c; }
        for (int i = 0; i < k; i++) FBPR_ROT(A[N * i + k], A[N * i + l]);
        for (int i = k + 1; i < l; i++) FBPR_ROT(A[N * k + i], A[N * i + l]);
        for (int i = l + 1; i < N; i++) FBPR_ROT(A[N * k + i], A[N * l + i]);
        for (int i = 0; i < N; i++) FBPR_ROT(V[N * k + i], V[N * l + i]);
#undef FBPR_ROT
        for (int j = 0; j < 2; j++) {
            int idx = j == 0 ? k : l;
            if (idx < N - 1) {
                int m = idx + 1; mv = fabsf(A[N * idx + m]);
                for (int i = idx + 2; i < N; i++) { float val = fabsf(A[N * idx + i]); if (mv < val) { mv = val; m = i; } }
                indR[idx] = m;
            }
            if (idx > 0) {
                int m = 0; mv = fabsf(A[idx]);
                for (int i = 1; i < idx; i++) { float val = fabsf(A[N * i + idx]); if (mv < val) { mv = val; m = i; } }
                indC[idx] = m;
            }
        }
    }
    for (int k = 0; k < N - 1; k++) {
        int m = k;
        for (int i = k + 1; i < N; i++) if (W[m] < W[i]) m = i;
        if (k != m) {
            float tw = W[m]; W[m] = W[k]; W[k] = tw;
            for (int i = 0; i < N; i++) { float tv = V[N * m + i]; V[N * m + i] = V[N * k + i]; V[N * k + i] = tv; }
        }
    }
}

// A (6x6 row-major) and b are destroyed.  Returns 0 and x = 0 when OpenCV reports singular.
__device__ inline int dev_qr_solve6(float* A, float* b, float* x) {
    const int n = 6;
    const float eps = FLT_EPSILON * 10;
    float vl[6], h[6];
    #pragma unroll
    for (int l = 0; l < n; l++) {
        float vlNorm = 0.f;
        #pragma unroll
        for (int i = 0; i < n - l; i++) { vl[i] = A[(l + i) * n + l]; vlNorm += vl[i] * vl[i]; }
        float tmpV = vl[0];
        vl[0] = vl[0] + ((vl[0] >= 0) ? 1.f : -1.f) * sqrtf(vlNorm);
        vlNorm = sqrtf(vlNorm + vl[0] * vl[0] - tmpV * tmpV);
        #pragma unroll
        for (int i = 0; i < n - l; i++) vl[i] /= vlNorm;
        #pragma unroll
        for (int j = l; j < n; j++) {
            float v_lA = 0.f;
            #pragma unroll
            for (int i = l; i < n; i++) v_lA += vl[i - l] * A[i * n + j];
            #pragma unroll
            for (int i = l; i < n; i++) A[i * n + j] -= 2 * vl[i - l] * v_lA;
        }
        h[l] = vl[0] * vl[0];
        #pragma unroll
        for (int i = 1; i < n - l; i++) A[(l + i) * n + l] = vl[i] / vl[0];
    }
    #pragma unroll
    for (int l = 0; l < n; l++) {
        vl[0] = 1.f;
        #pragma unroll
        for (int j = 1; j < n - l; j++) vl[j] = A[(j + l) * n + l];
        float v_lB = 0.f;
        #pragma unroll
        for (int i = l; i < n; i++) v_lB += vl[i - l] * b[i];
        #pragma unroll
        for (int i = l; i < n; i++) b[i] -= 2 * vl[i - l] * v_lB * h[l];
    }
    #pragma unroll
    for (int i = n - 1; i >= 0; i--) {
        #pragma unroll
        for (int j = n - 1; j > i; j--) b[i] -= b[j] * A[i * n + j];
        if (fabsf(A[i * n + i]) < eps) { for (int q = 0; q < n; q++) x[q] = 0.f; return 0; }
        b[i] /= A[i * n + i];
    }
    #pragma unroll
    for (int i = 0; i < n; i++) x[i] = b[i];
    return 1;
}

// The same solve by ONE WARP (all 32 lanes must call it), registers only: lane j < 6 owns column j of A, the Householder vector
// of step l is rebuilt by every lane from a broadcast of column l, every lane updates its own column (the columns of a step are
// independent; inside a column the operations keep hal::QR32f's order, so every element sees exactly the roundings of the
// one-thread routine), b and the back substitution are carried redundantly by all lanes.  A, b: 6x6 row-major and 6 values
// readable by every lane (shared memory).  Every lane returns the same x and the same status.
__device__ __forceinline__ int dev_qr_solve6_warp(const float* A, const float* b_in, float* x) {
    const unsigned FULL = 0xffffffffu;
    const int n = 6;
    const float eps = FLT_EPSILON * 10;
    const int lane = threadIdx.x & 31;
    const int col = lane < n ? lane : 0;                      // lanes >= 6 shadow column 0 (their results are never read)
    float c[6], b[6], h[6], vl[6];
    #pragma unroll
    for (int i = 0; i < n; i++) { c[i] = A[i * n + col]; b[i] = b_in[i]; }
    #pragma unroll
    for (int l = 0; l < n; l++) {
        float vlNorm = 0.f;
        #pragma unroll
        for (int i = 0; i < n - l; i++) { vl[i] = __shfl_sync(FULL, c[l + i], l); vlNorm += vl[i] * vl[i]; }
        const float tmpV = vl[0];
        vl[0] = vl[0] + ((vl[0] >= 0) ? 1.f : -1.f) * sqrtf(vlNorm);
        vlNorm = sqrtf(vlNorm + vl[0] * vl[0] - tmpV * tmpV);
        #pragma unroll
        for (int i = 0; i < n - l; i++) vl[i] /= vlNorm;
        if (lane >= l) {                                      // columns l .. 5 (and the shadows)
            float v_lA = 0.f;
            #pragma unroll
            for (int i = l; i < n; i++) v_lA += vl[i - l] * c[i];
            #pragma unroll
            for (int i = l; i < n; i++) c[i] -= 2 * vl[i - l] * v_lA;
        }
        h[l] = vl[0] * vl[0];
        if (lane == l) {
            #pragma unroll
            for (int i = 1; i < n - l; i++) c[l + i] = vl[i] / vl[0];
        }
    }
    #pragma unroll
    for (int l = 0; l < n; l++) {
        vl[0] = 1.f;
        #pragma unroll
        for (int j = 1; j < n - l; j++) vl[j] = __shfl_sync(FULL, c[j + l], l);
        float v_lB = 0.f;
        #pragma unroll
        for (int i = l; i < n; i++) v_lB += vl[i - l] * b[i];
        #pragma unroll
        for (int i = l; i < n; i++) b[i] -= 2 * vl[i - l] * v_lB * h[l];
    }
    int ok = 1;
    #pragma unroll
    for (int i = n - 1; i >= 0; i--) {
        #pragma unroll
        for (int j = n - 1; j > i; j--) b[i] -= b[j] * __shfl_sync(FULL, c[i], j);
        const float d = __shfl_sync(FULL, c[i], i);
        if (fabsf(d) < eps) ok = 0;                           // uniform: every lane sees the same d
        b[i] /= d;
    }
    #pragma unroll
    for (int i = 0; i < n; i++) x[i] = ok ? b[i] : 0.f;
    return ok;
}

// A destroyed; B = inverse (all zeros if singular by OpenCV's test).
__device__ inline int dev_lu_invert6(float* A, float* B) {
    const int n = 6;
    const float eps = FLT_EPSILON * 10;
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) B[i * n + j] = (i == j) ? 1.f : 0.f;
    for (int i = 0; i < n; i++) {
        int k = i;
        for (int j = i + 1; j < n; j++) if (fabsf(A[j * n + i]) > fabsf(A[k * n + i])) k = j;
        if (fabsf(A[k * n + i]) < eps) { for (int q = 0; q < n * n; q++) B[q] = 0.f; return 0; }
        if (k != i) {
            for (int j = i; j < n; j++) { float t = A[i * n + j]; A[i * n + j] = A[k * n + j]; A[k * n + j] = t; }
            for (int j = 0; j < n; j++) { float t = B[i * n + j]; B[i * n + j] = B[k * n + j]; B[k * n + j] = t; }
        }
        float d = -1 / A[i * n + i];
        for (int j = i + 1; j < n; j++) {
            float alpha = A[j * n + i] * d;
            for (int q = i + 1; q < n; q++) A[j * n + q] += alpha * A[i * n + q];
            for (int q = 0; q < n; q++) B[j * n + q] += alpha * B[i * n + q];
        }
    }
    for (int i = n - 1; i >= 0; i--)
        for (int j = 0; j < n; j++) {
            float s = B[i * n + j];
            for (int q = i + 1; q < n; q++) s -= A[i * n + q] * B[q * n + j];
            B[i * n + j] = s / A[i * n + i];
        }
    return 1;
}

// cv::eigen for 3x3 with every index resolved at compile time (registers only).  Follows JacobiImpl_
// exactly, including its incrementally maintained (and therefore sometimes stale) indR/indC pivot hints:
// for n = 3 only indR[0] (argmax of |a01|,|a02|) and indC[2] (argmax of |a02|,|a12|) are variable.
// Input: upper triangle a00..a22.  Output: W descending, V rows = eigenvectors (v[row][col]).
__device__ __forceinline__ void dev_rot(float& v0, float& v1, float c, float s) { float a0 = v0, b0 = v1; v0 = a0 * c - b0 * s; v1 = a0 * s + b0 * c; }

__device__ inline void dev_jacobi3(float a00, float a01, float a02, float a11, float a12, float a22, float* W, float* V) {
    float w0 = a00, w1 = a11, w2 = a22;
    float v00 = 1.f, v01 = 0.f, v02 = 0.f, v10 = 0.f, v11 = 1.f, v12 = 0.f, v20 = 0.f, v21 = 0.f, v22 = 1.f;
    int indR0 = (fabsf(a01) < fabsf(a02)) ? 2 : 1;
    int indC2 = (fabsf(a02) < fabsf(a12)) ? 1 : 0;
    for (int iters = 0; iters < 270; iters++) {
        // pivot search (same comparison order as the reference implementation)
        int k = 0;
        float mv = indR0 == 1 ? fabsf(a01) : fabsf(a02);
        { float val = fabsf(a12); if (mv < val) { mv = val; k = 1; } }
        int l = k == 0 ? indR0 : 2;
        { float val = fabsf(a01); if (mv < val) { mv = val; k = 0; l = 1; } }
        { float val = indC2 == 0 ? fabsf(a02) : fabsf(a12); if (mv < val) { mv = val; k = indC2; l = 2; } }
        const int cs = (k == 0) ? (l == 1 ? 0 : 1) : 2;       // (0,1) (0,2) (1,2)
        float p = cs == 0 ? a01 : (cs == 1 ? a02 : a12);
        if (fabsf(p) <= FLT_EPSILON) break;
        float wk = cs == 2 ? w1 : w0, wl = cs == 0 ? w1 : w2;
        float y = (wl - wk) * 0.5f;
        float t = fabsf(y) + dev_hypot(p, y);
        float s = dev_hypot(p, t);
        float c = t / s;
        s = p / s; t = (p / t) * p;
        if (y < 0) { s = -s; t = -t; }
        if (cs == 0) {
            a01 = 0; w0 -= t; w1 += t;
            dev_rot(a02, a12, c, s);
            dev_rot(v00, v10, c, s); dev_rot(v01, v11, c, s); dev_rot(v02, v12, c, s);
        } else if (cs == 1) {
            a02 = 0; w0 -= t; w2 += t;
            dev_rot(a01, a12, c, s);
            dev_rot(v00, v20, c, s); dev_rot(v01, v21, c, s); dev_rot(v02, v22, c, s);
        } else {
            a12 = 0; w1 -= t; w2 += t;
            dev_rot(a01, a02, c, s);
            dev_rot(v10, v20, c, s); dev_rot(v11, v21, c, s); dev_rot(v12, v22, c, s);
        }
        if (k == 0) indR0 = (fabsf(a01) < fabsf(a02)) ? 2 : 1;      // idx = k = 0 < n-1
        if (l == 2) indC2 = (fabsf(a02) < fabsf(a12)) ? 1 : 0;      // idx = l = 2 > 0
    }
    // selection sort, descending, first maximum wins
    {
        int m = 0; float wm = w0;
        if (wm < w1) { m = 1; wm = w1; }
        if (wm < w2) { m = 2; wm = w2; }
        if (m == 1) { float tw = w1; w1 = w0; w0 = tw; float t0 = v10, t1 = v11, t2 = v12; v10 = v00; v11 = v01; v12 = v02; v00 = t0; v01 = t1; v02 = t2; }
        else if (m == 2) { float tw = w2; w2 = w0; w0 = tw; float t0 = v20, t1 = v21, t2 = v22; v20 = v00; v21 = v01; v22 = v02; v00 = t0; v01 = t1; v02 = t2; }
        if (w1 < w2) { float tw = w2; w2 = w1; w1 = tw; float t0 = v20, t1 = v21, t2 = v22; v20 = v10; v21 = v11; v22 = v12; v10 = t0; v11 = t1; v12 = t2; }
    }
    W[0] = w0; W[1] = w1; W[2] = w2;
    V[0] = v00; V[1] = v01; V[2] = v02; V[3] = v10; V[4] = v11; V[5] = v12; V[6] = v20; V[7] = v21; V[8] = v22;
}

// Cheap certificate that cv::eigen(AtA) would report every eigenvalue >= 100, i.e. "not degenerate"
// (mapOptmization.h:1356-1366), so the 6x6 Jacobi + LU inverse of iteration 0 can be skipped: A - c*I is
// positive definite (f64 LDL^T, all pivots > 0) for c = 100 + a margin far above the f32 Jacobi's error.
// Returns false when in doubt; the caller then runs the exact decomposition.
__device__ inline bool dev_surely_not_degenerate(const float* A) {
    double fro = 0.0;
    #pragma unroll
    for (int i = 0; i < 36; i++) fro += (double)A[i] * (double)A[i];
    const double c = 100.0 + 1.0 + 1e-4 * sqrt(fro);
    double L[36];
    #pragma unroll
    for (int i = 0; i < 36; i++) L[i] = (double)A[i];
    #pragma unroll
    for (int i = 0; i < 6; i++) L[i * 6 + i] -= c;
    bool ok = true;
    #pragma unroll
    for (int j = 0; j < 6; j++) {
        double d = L[j * 6 + j];
        #pragma unroll
        for (int q = 0; q < j; q++) d -= L[j * 6 + q] * L[j * 6 + q] * L[q * 6 + q];
        ok = ok && (d > 1e-3 * c);
        L[j * 6 + j] = d;
        const double inv = 1.0 / (ok ? d : 1.0);
        #pragma unroll
        for (int i = j + 1; i < 6; i++) {
            double v = L[i * 6 + j];
            #pragma unroll
            for (int q = 0; q < j; q++) v -= L[i * 6 + q] * L[j * 6 + q] * L[q * 6 + q];
            L[i * 6 + j] = v * inv;
        }
    }
    return ok;
}

// x = argmin |A x - b| for the 5 x 3 row-major A, b = (-1,...,-1): the plane through 5 map points.
// Same operation sequence as Eigen's ColPivHouseholderQR (see the oracle); written so that every array
// index is a compile-time constant (column swaps are explicit branches) and everything stays in registers.
__device__ inline void dev_plane_solve(const float* Ain, float* x) {
    const int rows = 5, cols = 3;
    float q[3][5];                                   // q[col][row]
    #pragma unroll
    for (int i = 0; i < rows; i++) { q[0][i] = Ain[i * 3]; q[1][i] = Ain[i * 3 + 1]; q[2][i] = Ain[i * 3 + 2]; }
    float hC[3]; int perm[3] = { 0, 1, 2 };
    float nU[3], nD[3];
    #pragma unroll
    for (int k = 0; k < cols; k++) {
        float s = 0.f;
        #pragma unroll
        for (int i = 0; i < rows; i++) s += q[k][i] * q[k][i];
        nD[k] = sqrtf(s); nU[k] = nD[k];
    }
    float maxNorm = nU[0];
    if (nU[1] > maxNorm) maxNorm = nU[1];
    if (nU[2] > maxNorm) maxNorm = nU[2];
    float th = maxNorm * FLT_EPSILON;
    const float threshold_helper = (th * th) / (float)rows;
    const float norm_downdate_threshold = sqrtf(FLT_EPSILON);
    int nzp = cols;
#define FBPR_SWAPCOL(A_, B_) { \
        _Pragma("unroll") for (int i_ = 0; i_ < rows; i_++) { float t_ = q[A_][i_]; q[A_][i_] = q[B_][i_]; q[B_][i_] = t_; } \
        float t_ = nU[A_]; nU[A_] = nU[B_]; nU[B_] = t_; t_ = nD[A_]; nD[A_] = nD[B_]; nD[B_] = t_; \
        int ti_ = perm[A_]; perm[A_] = perm[B_]; perm[B_] = ti_; }
    #pragma unroll
    for (int k = 0; k < cols; k++) {
        int big = k; float bigv = nU[k];
        #pragma unroll
        for (int j = k + 1; j < cols; j++) if (nU[j] > bigv) { bigv = nU[j]; big = j; }
        if (nzp == cols && bigv * bigv < threshold_helper * (float)(rows - k)) nzp = k;
        if (k == 0) { if (big == 1) FBPR_SWAPCOL(0, 1) else if (big == 2) FBPR_SWAPCOL(0, 2) }
        else if (k == 1) { if (big == 2) FBPR_SWAPCOL(1, 2) }
        float tailSq = 0.f;
        #pragma unroll
        for (int i = k + 1; i < rows; i++) tailSq += q[k][i] * q[k][i];
        float c0 = q[k][k], tau, beta;
        if (tailSq <= FLT_MIN) {
            tau = 0.f; beta = c0;
            #pragma unroll
            for (int i = k + 1; i < rows; i++) q[k][i] = 0.f;
        } else {
            beta = sqrtf(c0 * c0 + tailSq);
            if (c0 >= 0.f) beta = -beta;
            float den = c0 - beta;
            #pragma unroll
            for (int i = k + 1; i < rows; i++) q[k][i] = q[k][i] / den;
            tau = (beta - c0) / beta;
        }
        hC[k] = tau; q[k][k] = beta;
        if (tau != 0.f) {
            #pragma unroll
            for (int j = k + 1; j < cols; j++) {
                float tmp = 0.f;
                #pragma unroll
                for (int i = k + 1; i < rows; i++) tmp += q[k][i] * q[j][i];
                tmp += q[j][k];
                q[j][k] -= tau * tmp;
                #pragma unroll
                for (int i = k + 1; i < rows; i++) q[j][i] -= (tau * q[k][i]) * tmp;
            }
        }
        #pragma unroll
        for (int j = k + 1; j < cols; j++) {
            if (nU[j] != 0.f) {
                float temp = fabsf(q[j][k]) / nU[j];
                temp = (1.f + temp) * (1.f - temp);
                temp = temp < 0.f ? 0.f : temp;
                float ratio = nU[j] / nD[j];
                float temp2 = temp * (ratio * ratio);
                if (temp2 <= norm_downdate_threshold) {
                    float s = 0.f;
                    #pragma unroll
                    for (int i = k + 1; i < rows; i++) s += q[j][i] * q[j][i];
                    nD[j] = sqrtf(s); nU[j] = nD[j];
                } else {
                    nU[j] *= sqrtf(temp);
                }
            }
        }
    }
#undef FBPR_SWAPCOL
    float x0 = 0.f, x1 = 0.f, x2 = 0.f;
    if (nzp != 0) {
        float c[5] = { -1.f, -1.f, -1.f, -1.f, -1.f };
        #pragma unroll
        for (int k = 0; k < cols; k++) {
            if (k < nzp) {
                float tau = hC[k];
                if (tau != 0.f) {
                    float tmp = 0.f;
                    #pragma unroll
                    for (int i = k + 1; i < rows; i++) tmp += q[k][i] * c[i];
                    tmp += c[k];
                    c[k] -= tau * tmp;
                    #pragma unroll
                    for (int i = k + 1; i < rows; i++) c[i] -= (tau * q[k][i]) * tmp;
                }
            }
        }
        #pragma unroll
        for (int i = cols - 1; i >= 0; i--) {
            if (i < nzp) {
                c[i] /= q[i][i];
                #pragma unroll
                for (int r = 0; r < i; r++) c[r] -= c[i] * q[i][r];
            }
        }
        #pragma unroll
        for (int i = 0; i < cols; i++) {
            if (i < nzp) { if (perm[i] == 0) x0 = c[i]; else if (perm[i] == 1) x1 = c[i]; else x2 = c[i]; }
        }
    }
    x[0] = x0; x[1] = x1; x[2] = x2;
}
