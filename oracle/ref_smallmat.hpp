// TEST INFRASTRUCTURE -- CPU oracle.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may build, load or call anything under oracle/.
//
// ref_smallmat.hpp -- restatements of the third-party small-matrix arithmetic the reference's
// scan-to-map path calls (none of it is vendored under /root/reference; versions unpinned,
// SURVEY.md section 8(c)).  All f32, one rounding per operation, no FMA (build with
// -ffp-contract=off, no -march=native, no -ffast-math).
//
//   jacobi_eigen_sym   = cv::eigen          (OpenCV JacobiImpl_)      mapOptmization.h:1060, :1353
//   qr_solve           = cv::solve(QR)      (OpenCV hal::QR32f)       mapOptmization.h:1343
//   lu_invert          = cv::Mat::inv()     (OpenCV hal::LU32f)       mapOptmization.h:1370
//   matmul_f64acc      = small cv::Mat product (f64 accumulate)       mapOptmization.h:1370, :1376
//   colpiv_householder_solve_5x3 = Eigen::ColPivHouseholderQR<5x3>::solve  mapOptmization.h:1169
//   get_transformation / get_translation_and_euler = pcl/common/eigen.hpp  mapOptmization.h:309,:326,:414,:447
//
// Pinning: jacobi_eigen_sym, qr_solve and lu_invert are checked bit-for-bit against the cv2 4.13
// wheel in tests/test_oracle_smallmat.py.  The Eigen and PCL pieces cannot be pinned here (no
// Eigen/PCL in the container): "parity unpinned" for those; the scalar sequential order below
// is the definition.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstring>
#include <utility>

namespace orc {

// ---- trig contract (hard part 1): both sides evaluate f32 trig as the
// correctly-rounded-in-double value rounded once to float.
static inline float sinf_c(float x) { return (float)std::sin((double)x); }
static inline float cosf_c(float x) { return (float)std::cos((double)x); }
static inline float atan2f_c(float y, float x) { return (float)std::atan2((double)y, (double)x); }
static inline float asinf_c(float x) { return (float)std::asin((double)x); }

static inline float cv_hypot(float a, float b) {
    a = std::fabs(a); b = std::fabs(b);
    if (a > b) { b /= a; return a * std::sqrt(1 + b * b); }
    if (b > 0) { a /= b; return b * std::sqrt(1 + a * a); }
    return 0;
}

// cv::eigen for a symmetric n x n f32 matrix (n <= 6).  A is row-major and is destroyed.
// W: eigenvalues descending.  V: eigenvectors as ROWS, row-major.
static inline void jacobi_eigen_sym(int n, float* A, float* W, float* V) {
    const float eps = FLT_EPSILON;
    int indR[8], indC[8];
    int i, j, k, m;
    for (i = 0; i < n; i++) { for (j = 0; j < n; j++) V[i * n + j] = 0.f; V[i * n + i] = 1.f; }
    float mv = 0.f;
    for (k = 0; k < n; k++) {
        W[k] = A[(n + 1) * k];
        if (k < n - 1) {
            for (m = k + 1, mv = std::fabs(A[n * k + m]), i = k + 2; i < n; i++) {
                float val = std::fabs(A[n * k + i]);
                if (mv < val) mv = val, m = i;
            }
            indR[k] = m;
        }
        if (k > 0) {
            for (m = 0, mv = std::fabs(A[k]), i = 1; i < k; i++) {
                float val = std::fabs(A[n * i + k]);
                if (mv < val) mv = val, m = i;
            }
            indC[k] = m;
        }
    }
    if (n > 1) for (int iters = 0, maxIters = n * n * 30; iters < maxIters; iters++) {
        for (k = 0, mv = std::fabs(A[indR[0]]), i = 1; i < n - 1; i++) {
            float val = std::fabs(A[n * i + indR[i]]);
            if (mv < val) mv = val, k = i;
        }
        int l = indR[k];
        for (i = 1; i < n; i++) {
            float val = std::fabs(A[n * indC[i] + i]);
            if (mv < val) mv = val, k = indC[i], l = i;
        }
        float p = A[n * k + l];
        if (std::fabs(p) <= eps) break;
        float y = (W[l] - W[k]) * 0.5f;
        float t = std::fabs(y) + cv_hypot(p, y);
        float s = cv_hypot(p, t);
        float c = t / s;
        s = p / s; t = (p / t) * p;
        if (y < 0) s = -s, t = -t;
        A[n * k + l] = 0;
        W[k] -= t;
        W[l] += t;
        float a0, b0;
#define ORC_ROT(v0, v1) a0 = v0, b0 = v1, v0 = a0 * c - b0 * s, v1 = a0 * s + b0 * c
        for (i = 0; i < k; i++) ORC_ROT(A[n * i + k], A[n * i + l]);
        for (i = k + 1; i < l; i++) ORC_ROT(A[n * k + i], A[n * i + l]);
        for (i = l + 1; i < n; i++) ORC_ROT(A[n * k + i], A[n * l + i]);
        for (i = 0; i < n; i++) ORC_ROT(V[n * k + i], V[n * l + i]);
#undef ORC_ROT
        for (j = 0; j < 2; j++) {
            int idx = j == 0 ? k : l;
            if (idx < n - 1) {
                for (m = idx + 1, mv = std::fabs(A[n * idx + m]), i = idx + 2; i < n; i++) {
                    float val = std::fabs(A[n * idx + i]);
                    if (mv < val) mv = val, m = i;
                }
                indR[idx] = m;
            }
            if (idx > 0) {
                for (m = 0, mv = std::fabs(A[idx]), i = 1; i < idx; i++) {
                    float val = std::fabs(A[n * i + idx]);
                    if (mv < val) mv = val, m = i;
                }
                indC[idx] = m;
            }
        }
    }
    for (k = 0; k < n - 1; k++) {
        m = k;
        for (i = k + 1; i < n; i++) if (W[m] < W[i]) m = i;
        if (k != m) {
            std::swap(W[m], W[k]);
            for (i = 0; i < n; i++) std::swap(V[n * m + i], V[n * k + i]);
        }
    }
}

// cv::solve(A, b, x, DECOMP_QR) for square n x n f32 (n <= 6), one right-hand side.
// A (row-major) and b are destroyed; returns 0 (and x = 0) when OpenCV would report singular.
static inline int qr_solve(int n, float* A, float* b, float* x) {
    const float eps = FLT_EPSILON * 10;
    const int m = n;
    float vl[8], h[8];
    for (int l = 0; l < n; l++) {
        int vlSize = m - l;
        float vlNorm = 0.f;
        for (int i = 0; i < vlSize; i++) { vl[i] = A[(l + i) * n + l]; vlNorm += vl[i] * vl[i]; }
        float tmpV = vl[0];
        vl[0] = vl[0] + ((vl[0] >= 0) ? 1 : -1) * std::sqrt(vlNorm);
        vlNorm = std::sqrt(vlNorm + vl[0] * vl[0] - tmpV * tmpV);
        for (int i = 0; i < vlSize; i++) vl[i] /= vlNorm;
        for (int j = l; j < n; j++) {
            float v_lA = 0.f;
            for (int i = l; i < m; i++) v_lA += vl[i - l] * A[i * n + j];
            for (int i = l; i < m; i++) A[i * n + j] -= 2 * vl[i - l] * v_lA;
        }
        h[l] = vl[0] * vl[0];
        for (int i = 1; i < vlSize; i++) A[(l + i) * n + l] = vl[i] / vl[0];
    }
    for (int l = 0; l < n; l++) {
        vl[0] = 1.f;
        for (int j = 1; j < m - l; j++) vl[j] = A[(j + l) * n + l];
        float v_lB = 0.f;
        for (int i = l; i < m; i++) v_lB += vl[i - l] * b[i];
        for (int i = l; i < m; i++) b[i] -= 2 * vl[i - l] * v_lB * h[l];
    }
    for (int i = n - 1; i >= 0; i--) {
        for (int j = n - 1; j > i; j--) b[i] -= b[j] * A[i * n + j];
        if (std::fabs(A[i * n + i]) < eps) { for (int q = 0; q < n; q++) x[q] = 0.f; return 0; }
        b[i] /= A[i * n + i];
    }
    for (int i = 0; i < n; i++) x[i] = b[i];
    return 1;
}

// cv::Mat::inv(DECOMP_LU) for n x n f32 (n <= 6).  A destroyed; B receives the inverse
// (all zeros when OpenCV would report singular).
static inline int lu_invert(int n, float* A, float* B) {
    const float eps = FLT_EPSILON * 10;
    const int m = n;
    int i, j, k, p = 1;
    for (i = 0; i < n; i++) for (j = 0; j < n; j++) B[i * n + j] = (i == j) ? 1.f : 0.f;
    for (i = 0; i < m; i++) {
        k = i;
        for (j = i + 1; j < m; j++) if (std::fabs(A[j * n + i]) > std::fabs(A[k * n + i])) k = j;
        if (std::fabs(A[k * n + i]) < eps) { for (int q = 0; q < n * n; q++) B[q] = 0.f; return 0; }
        if (k != i) {
            for (j = i; j < m; j++) std::swap(A[i * n + j], A[k * n + j]);
            for (j = 0; j < n; j++) std::swap(B[i * n + j], B[k * n + j]);
            p = -p;
        }
        float d = -1 / A[i * n + i];
        for (j = i + 1; j < m; j++) {
            float alpha = A[j * n + i] * d;
            for (k = i + 1; k < m; k++) A[j * n + k] += alpha * A[i * n + k];
            for (k = 0; k < n; k++) B[j * n + k] += alpha * B[i * n + k];
        }
    }
    for (i = m - 1; i >= 0; i--)
        for (j = 0; j < n; j++) {
            float s = B[i * n + j];
            for (k = i + 1; k < m; k++) s -= A[i * n + k] * B[k * n + j];
            B[i * n + j] = s / A[i * n + i];
        }
    return p;
}

// C(r x c) = A(r x k) * B(k x c), f32 in/out, f64 accumulation rounded once.
static inline void matmul_f64acc(int r, int k, int c, const float* A, const float* B, float* C) {
    for (int i = 0; i < r; i++)
        for (int j = 0; j < c; j++) {
            double s = 0.0;
            for (int q = 0; q < k; q++) s += (double)A[i * k + q] * (double)B[q * c + j];
            C[i * c + j] = (float)s;
        }
}

// Eigen::ColPivHouseholderQR<Matrix<float,5,3>>(A).solve(b), Eigen 3.3 algorithm, scalar
// sequential reductions.  A is 5x3 row-major.
static inline void colpiv_householder_solve_5x3(const float* Ain, const float* bin, float* x) {
    const int rows = 5, cols = 3, size = 3;
    float qr[5][3];
    for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) qr[i][j] = Ain[i * 3 + j];
    float hCoeffs[3];
    int perm[3] = {0, 1, 2};
    float normsUpdated[3], normsDirect[3];
    for (int k = 0; k < cols; k++) {
        float s = 0.f;
        for (int i = 0; i < rows; i++) s += qr[i][k] * qr[i][k];
        normsDirect[k] = std::sqrt(s);
        normsUpdated[k] = normsDirect[k];
    }
    float maxNorm = normsUpdated[0];
    for (int k = 1; k < cols; k++) if (normsUpdated[k] > maxNorm) maxNorm = normsUpdated[k];
    float th = maxNorm * FLT_EPSILON;
    const float threshold_helper = (th * th) / (float)rows;
    const float norm_downdate_threshold = std::sqrt(FLT_EPSILON);
    int nonzero_pivots = size;
    for (int k = 0; k < size; k++) {
        int big = k; float bigv = normsUpdated[k];
        for (int j = k + 1; j < cols; j++) if (normsUpdated[j] > bigv) { bigv = normsUpdated[j]; big = j; }
        float biggest_col_sq_norm = bigv * bigv;
        if (nonzero_pivots == size && biggest_col_sq_norm < threshold_helper * (float)(rows - k)) nonzero_pivots = k;
        if (k != big) {
            for (int i = 0; i < rows; i++) std::swap(qr[i][k], qr[i][big]);
            std::swap(normsUpdated[k], normsUpdated[big]);
            std::swap(normsDirect[k], normsDirect[big]);
            std::swap(perm[k], perm[big]);
        }
        // makeHouseholderInPlace on qr[k..rows-1][k]
        float tailSqNorm = 0.f;
        for (int i = k + 1; i < rows; i++) tailSqNorm += qr[i][k] * qr[i][k];
        float c0 = qr[k][k];
        float tau, beta;
        if (tailSqNorm <= FLT_MIN) {
            tau = 0.f; beta = c0;
            for (int i = k + 1; i < rows; i++) qr[i][k] = 0.f;
        } else {
            beta = std::sqrt(c0 * c0 + tailSqNorm);
            if (c0 >= 0.f) beta = -beta;
            float den = c0 - beta;
            for (int i = k + 1; i < rows; i++) qr[i][k] = qr[i][k] / den;
            tau = (beta - c0) / beta;
        }
        hCoeffs[k] = tau;
        qr[k][k] = beta;
        // apply H_k to the trailing columns
        if (tau != 0.f) {
            for (int j = k + 1; j < cols; j++) {
                float tmp = 0.f;
                for (int i = k + 1; i < rows; i++) tmp += qr[i][k] * qr[i][j];
                tmp += qr[k][j];
                qr[k][j] -= tau * tmp;
                for (int i = k + 1; i < rows; i++) qr[i][j] -= (tau * qr[i][k]) * tmp;
            }
        }
        // norm downdate
        for (int j = k + 1; j < cols; j++) {
            if (normsUpdated[j] != 0.f) {
                float temp = std::fabs(qr[k][j]) / normsUpdated[j];
                temp = (1.f + temp) * (1.f - temp);
                temp = temp < 0.f ? 0.f : temp;
                float ratio = normsUpdated[j] / normsDirect[j];
                float temp2 = temp * (ratio * ratio);
                if (temp2 <= norm_downdate_threshold) {
                    float s = 0.f;
                    for (int i = k + 1; i < rows; i++) s += qr[i][j] * qr[i][j];
                    normsDirect[j] = std::sqrt(s);
                    normsUpdated[j] = normsDirect[j];
                } else {
                    normsUpdated[j] *= std::sqrt(temp);
                }
            }
        }
    }
    x[0] = x[1] = x[2] = 0.f;
    if (nonzero_pivots == 0) return;
    float c[5];
    for (int i = 0; i < rows; i++) c[i] = bin[i];
    for (int k = 0; k < nonzero_pivots; k++) {
        float tau = hCoeffs[k];
        if (tau != 0.f) {
            float tmp = 0.f;
            for (int i = k + 1; i < rows; i++) tmp += qr[i][k] * c[i];
            tmp += c[k];
            c[k] -= tau * tmp;
            for (int i = k + 1; i < rows; i++) c[i] -= (tau * qr[i][k]) * tmp;
        }
    }
    // back substitution, column-oriented (as Eigen's triangular solver does for col-major)
    for (int i = nonzero_pivots - 1; i >= 0; i--) {
        c[i] /= qr[i][i];
        for (int r = 0; r < i; r++) c[r] -= c[i] * qr[r][i];
    }
    for (int i = 0; i < nonzero_pivots; i++) x[perm[i]] = c[i];
}

// pcl::getTransformation(x,y,z,roll,pitch,yaw) -> 3x4 row-major (R = Rz*Ry*Rx), f32.
static inline void get_transformation(float x, float y, float z, float roll, float pitch, float yaw, float T[12]) {
    float A = cosf_c(yaw), B = sinf_c(yaw), C = cosf_c(pitch), D = sinf_c(pitch), E = cosf_c(roll), F = sinf_c(roll);
    float DE = D * E, DF = D * F;
    T[0] = A * C; T[1] = A * DF - B * E; T[2]  = B * F + A * DE; T[3]  = x;
    T[4] = B * C; T[5] = A * E + B * DF; T[6]  = B * DE - A * F; T[7]  = y;
    T[8] = -D;    T[9] = C * F;          T[10] = C * E;          T[11] = z;
}

// pcl::getTranslationAndEulerAngles: out = (x,y,z,roll,pitch,yaw)
static inline void get_translation_and_euler(const float T[12], float& x, float& y, float& z, float& roll, float& pitch, float& yaw) {
    x = T[3]; y = T[7]; z = T[11];
    roll = atan2f_c(T[9], T[10]);
    pitch = asinf_c(-T[8]);
    yaw = atan2f_c(T[4], T[0]);
}

}  // namespace orc
