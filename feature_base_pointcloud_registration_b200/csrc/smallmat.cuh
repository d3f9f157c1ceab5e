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
    for (int l = 0; l < n; l++) {
        int vlSize = n - l;
        float vlNorm = 0.f;
        for (int i = 0; i < vlSize; i++) { vl[i] = A[(l + i) * n + l]; vlNorm += vl[i] * vl[i]; }
        float tmpV = vl[0];
        vl[0] = vl[0] + ((vl[0] >= 0) ? 1.f : -1.f) * sqrtf(vlNorm);
        vlNorm = sqrtf(vlNorm + vl[0] * vl[0] - tmpV * tmpV);
        for (int i = 0; i < vlSize; i++) vl[i] /= vlNorm;
        for (int j = l; j < n; j++) {
            float v_lA = 0.f;
            for (int i = l; i < n; i++) v_lA += vl[i - l] * A[i * n + j];
            for (int i = l; i < n; i++) A[i * n + j] -= 2 * vl[i - l] * v_lA;
        }
        h[l] = vl[0] * vl[0];
        for (int i = 1; i < vlSize; i++) A[(l + i) * n + l] = vl[i] / vl[0];
    }
    for (int l = 0; l < n; l++) {
        vl[0] = 1.f;
        for (int j = 1; j < n - l; j++) vl[j] = A[(j + l) * n + l];
        float v_lB = 0.f;
        for (int i = l; i < n; i++) v_lB += vl[i - l] * b[i];
        for (int i = l; i < n; i++) b[i] -= 2 * vl[i - l] * v_lB * h[l];
    }
    for (int i = n - 1; i >= 0; i--) {
        for (int j = n - 1; j > i; j--) b[i] -= b[j] * A[i * n + j];
        if (fabsf(A[i * n + i]) < eps) { for (int q = 0; q < n; q++) x[q] = 0.f; return 0; }
        b[i] /= A[i * n + i];
    }
    for (int i = 0; i < n; i++) x[i] = b[i];
    return 1;
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

// x = argmin |A x - b| for the 5 x 3 row-major A, b = (-1,...,-1): the plane through 5 map points.
__device__ inline void dev_plane_solve(const float* Ain, float* x) {
    const int rows = 5, cols = 3;
    float qr[5][3];
    for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) qr[i][j] = Ain[i * 3 + j];
    float hC[3]; int perm[3] = { 0, 1, 2 };
    float nU[3], nD[3];
    for (int k = 0; k < cols; k++) {
        float s = 0.f;
        for (int i = 0; i < rows; i++) s += qr[i][k] * qr[i][k];
        nD[k] = sqrtf(s); nU[k] = nD[k];
    }
    float maxNorm = nU[0];
    for (int k = 1; k < cols; k++) if (nU[k] > maxNorm) maxNorm = nU[k];
    float th = maxNorm * FLT_EPSILON;
    const float threshold_helper = (th * th) / (float)rows;
    const float norm_downdate_threshold = sqrtf(FLT_EPSILON);
    int nzp = cols;
    for (int k = 0; k < cols; k++) {
        int big = k; float bigv = nU[k];
        for (int j = k + 1; j < cols; j++) if (nU[j] > bigv) { bigv = nU[j]; big = j; }
        if (nzp == cols && bigv * bigv < threshold_helper * (float)(rows - k)) nzp = k;
        if (k != big) {
            for (int i = 0; i < rows; i++) { float t = qr[i][k]; qr[i][k] = qr[i][big]; qr[i][big] = t; }
            float t = nU[k]; nU[k] = nU[big]; nU[big] = t;
            t = nD[k]; nD[k] = nD[big]; nD[big] = t;
            int ti = perm[k]; perm[k] = perm[big]; perm[big] = ti;
        }
        float tailSq = 0.f;
        for (int i = k + 1; i < rows; i++) tailSq += qr[i][k] * qr[i][k];
        float c0 = qr[k][k], tau, beta;
        if (tailSq <= FLT_MIN) {
            tau = 0.f; beta = c0;
            for (int i = k + 1; i < rows; i++) qr[i][k] = 0.f;
        } else {
            beta = sqrtf(c0 * c0 + tailSq);
            if (c0 >= 0.f) beta = -beta;
            float den = c0 - beta;
            for (int i = k + 1; i < rows; i++) qr[i][k] = qr[i][k] / den;
            tau = (beta - c0) / beta;
        }
        hC[k] = tau; qr[k][k] = beta;
        if (tau != 0.f) {
            for (int j = k + 1; j < cols; j++) {
                float tmp = 0.f;
                for (int i = k + 1; i < rows; i++) tmp += qr[i][k] * qr[i][j];
                tmp += qr[k][j];
                qr[k][j] -= tau * tmp;
                for (int i = k + 1; i < rows; i++) qr[i][j] -= (tau * qr[i][k]) * tmp;
            }
        }
        for (int j = k + 1; j < cols; j++) {
            if (nU[j] != 0.f) {
                float temp = fabsf(qr[k][j]) / nU[j];
                temp = (1.f + temp) * (1.f - temp);
                temp = temp < 0.f ? 0.f : temp;
                float ratio = nU[j] / nD[j];
                float temp2 = temp * (ratio * ratio);
                if (temp2 <= norm_downdate_threshold) {
                    float s = 0.f;
                    for (int i = k + 1; i < rows; i++) s += qr[i][j] * qr[i][j];
                    nD[j] = sqrtf(s); nU[j] = nD[j];
                } else {
                    nU[j] *= sqrtf(temp);
                }
            }
        }
    }
    x[0] = x[1] = x[2] = 0.f;
    if (nzp == 0) return;
    float c[5] = { -1.f, -1.f, -1.f, -1.f, -1.f };
    for (int k = 0; k < nzp; k++) {
        float tau = hC[k];
        if (tau != 0.f) {
            float tmp = 0.f;
            for (int i = k + 1; i < rows; i++) tmp += qr[i][k] * c[i];
            tmp += c[k];
            c[k] -= tau * tmp;
            for (int i = k + 1; i < rows; i++) c[i] -= (tau * qr[i][k]) * tmp;
        }
    }
    for (int i = nzp - 1; i >= 0; i--) {
        c[i] /= qr[i][i];
        for (int r = 0; r < i; r++) c[r] -= c[i] * qr[r][i];
    }
    for (int i = 0; i < nzp; i++) x[perm[i]] = c[i];
}
