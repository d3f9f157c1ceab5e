// TEST INFRASTRUCTURE -- CPU oracle (see ref_smallmat.hpp header).
//
// ref_cloud.hpp -- point-cloud primitives the reference takes from PCL / FLANN, restated:
//   voxel_grid        = pcl::VoxelGrid<PointXYZI>::filter   
//                       call sites featureExtraction.h:289-290, mapOptmization.h:251-257,:948-953,:985-991
//   KdTree5           = pcl::KdTreeFLANN<PointXYZI> setInputCloud + nearestKSearch(k=5)
//                       (FLANN KDTreeSingleIndex, leaf 15, L2_Simple, exact; Appendix B-2)
//                       call sites mapOptmization.h:1413-1414, :1020, :1143
//   crop_box          = pcl::CropBox (no transform): inclusive AABB, order preserving  mapOptmization.h:288-303
// Deliberate deviation from the literal libraries (hard part 6): every
// unstable order is made total by the point index -- (voxel key, idx), (d^2, idx).
#pragma once
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <vector>

namespace orc {

struct P4 { float x, y, z, i; };

// Oracle contract vs the literal reference: by default every unstable sort is made a total order by the
// point index.  With this flag set the two std::sort calls of the path compare exactly what the reference compares
// (featureExtraction.h:13-17 by_value on the curvature only; pcl::VoxelGrid on the voxel index only), so ties fall wherever
// libstdc++'s introsort leaves them.  tests / scripts use it to COUNT how many frames the tie-break rule changes.
inline int& oracle_literal_sort() { static int v = 0; return v; }


// Returns the number of output points.  keys_out (optional): per INPUT point voxel key.
// overflow (optional) is set when PCL's "leaf size too small" path copies input to output.
static inline int voxel_grid(const P4* in, int n, float leaf, std::vector<P4>& out,
                             std::vector<int>* keys_out = nullptr, std::vector<int>* out_keys = nullptr,
                             int* overflow = nullptr) {
    out.clear();
    if (keys_out) keys_out->assign(n, 0);
    if (out_keys) out_keys->clear();
    if (overflow) *overflow = 0;
    if (n <= 0) return 0;
    const float inv = 1.0f / leaf;
    float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (int k = 0; k < n; k++) {
        const float p[3] = { in[k].x, in[k].y, in[k].z };
        for (int c = 0; c < 3; c++) { if (p[c] < mn[c]) mn[c] = p[c]; if (p[c] > mx[c]) mx[c] = p[c]; }
    }
    int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1;
    int64_t dy = (int64_t)((mx[1] - mn[1]) * inv) + 1;
    int64_t dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
    // PCL tests dx*dy*dz in int64; the partial products are checked first so that the triple product itself cannot wrap
    // (found by the UBSan build, scripts/oracle_sanitize.sh) -- the outcome is the same wherever PCL's own test is defined
    const bool too_many = dx > (int64_t)INT_MAX || dy > (int64_t)INT_MAX || dz > (int64_t)INT_MAX || dx * dy > (int64_t)INT_MAX
                          || dx * dy * dz > (int64_t)INT_MAX;
    if (too_many) {
        out.assign(in, in + n);
        if (overflow) *overflow = 1;
        return n;
    }
    int min_b[3], max_b[3], div_b[3], mul[3];
    for (int c = 0; c < 3; c++) {
        min_b[c] = (int)std::floor(mn[c] * inv);
        max_b[c] = (int)std::floor(mx[c] * inv);
        div_b[c] = max_b[c] - min_b[c] + 1;
    }
    mul[0] = 1; mul[1] = div_b[0]; mul[2] = div_b[0] * div_b[1];
    std::vector<std::pair<int, int>> ki(n);
    for (int k = 0; k < n; k++) {
        int i0 = (int)(std::floor(in[k].x * inv) - (float)min_b[0]);
        int i1 = (int)(std::floor(in[k].y * inv) - (float)min_b[1]);
        int i2 = (int)(std::floor(in[k].z * inv) - (float)min_b[2]);
        int idx = i0 * mul[0] + i1 * mul[1] + i2 * mul[2];
        ki[k] = { idx, k };
        if (keys_out) (*keys_out)[k] = idx;
    }
    // (key, point index): total order.  Literal mode: pcl::VoxelGrid's own call, std::sort on the voxel index alone
    // (cloud_point_index_idx::operator<, voxel_grid.h) -- libstdc++ introsort, members of a voxel in unspecified order
    if (oracle_literal_sort()) std::sort(ki.begin(), ki.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first < b.first; });
    else std::sort(ki.begin(), ki.end());
    int first = 0;
    while (first < n) {
        int last = first + 1;
        while (last < n && ki[last].first == ki[first].first) last++;
        float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
        for (int q = first; q < last; q++) {
            const P4& p = in[ki[q].second];
            sx += p.x; sy += p.y; sz += p.z; si += p.i;
        }
        float cnt = (float)(last - first);
        out.push_back(P4{ sx / cnt, sy / cnt, sz / cnt, si / cnt });
        if (out_keys) out_keys->push_back(ki[first].first);
        first = last;
    }
    return (int)out.size();
}

static inline void crop_box(const P4* in, int n, const float mn[3], const float mx[3], std::vector<P4>& out) {
    out.clear();
    for (int k = 0; k < n; k++) {
        const P4& p = in[k];
        if (p.x < mn[0] || p.y < mn[1] || p.z < mn[2]) continue;
        if (p.x > mx[0] || p.y > mx[1] || p.z > mx[2]) continue;
        out.push_back(p);
    }
}

// ---------------------------------------------------------------- exact 5-NN
struct Knn5 {
    float d[5]; int id[5]; int cnt;
    Knn5() : cnt(0) { for (int k = 0; k < 5; k++) { d[k] = FLT_MAX; id[k] = INT_MAX; } }
    inline bool better(float dd, int ii, int slot) const { return dd < d[slot] || (dd == d[slot] && ii < id[slot]); }
    inline void offer(float dd, int ii) {
        if (!better(dd, ii, 4)) return;
        int k = 4;
        while (k > 0 && better(dd, ii, k - 1)) { d[k] = d[k - 1]; id[k] = id[k - 1]; k--; }
        d[k] = dd; id[k] = ii;
        if (cnt < 5) cnt++;
    }
    inline float worst() const { return d[4]; }
};

static inline float l2_simple(const float* a, const float* b) {
    float d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
    float r = d0 * d0; r += d1 * d1; r += d2 * d2;     // FLANN L2_Simple accumulation order
    return r;
}

// Split rule of the kd-tree: 1 (default) = FLANN's KDTreeSingleIndex::middleSplit_ (cut the widest box side at its middle, clipped
// to the points' extent; planeSplit partition, index = lim1 / lim2 / count/2) so that tree depth and leaf occupancy -- i.e. the CPU
// time of build and search -- follow the library the reference calls; 0 = median split (balanced).  The search is exact with
// either rule, so results are identical by construction (tests assert it).  Restated from memory of FLANN 1.8 / 1.9
// (flann/algorithms/kdtree_single_index.h).  The search RESULTS (index order, squared distances, the strict radius rule) are
// pinned against a real FLANN KDTreeSingleIndex -- cv2.flann, OpenCV's vendored copy of the library: tests/golden/flann_cv2.npz,
// tests/test_oracle_cloud.py -- bit for bit wherever distances are distinct; among exactly equal distances FLANN's order follows
// its tree traversal and the oracle's (d^2, index) rule is the documented deviation.
inline int& oracle_kdtree_flann_split() { static int v = 1; return v; }

class KdTree5 {
public:
    void build(const P4* pts, int n) {
        n_ = n;
        xyz_.resize((size_t)n * 3); ids_.resize(n);
        for (int k = 0; k < n; k++) ids_[k] = k;
        src_ = pts;
        nodes_.clear(); nodes_.reserve(n / 6 + 16);
        for (int c = 0; c < 3; c++) { lo_[c] = FLT_MAX; hi_[c] = -FLT_MAX; }
        for (int k = 0; k < n; k++) {
            const float p[3] = { pts[k].x, pts[k].y, pts[k].z };
            for (int c = 0; c < 3; c++) { lo_[c] = std::min(lo_[c], p[c]); hi_[c] = std::max(hi_[c], p[c]); }
        }
        if (n > 0) {
            float lo[3] = { lo_[0], lo_[1], lo_[2] }, hi[3] = { hi_[0], hi_[1], hi_[2] };
            if (oracle_kdtree_flann_split()) divide_flann(0, n, lo, hi); else divide(0, n, lo, hi);
        }
        for (int k = 0; k < n; k++) { const P4& p = pts[ids_[k]]; xyz_[3 * k] = p.x; xyz_[3 * k + 1] = p.y; xyz_[3 * k + 2] = p.z; }
        src_ = nullptr;
    }
    int size() const { return n_; }
    // fills idx[5], d2[5] ascending by (d2, idx); slots beyond the map size keep (FLT_MAX, INT_MAX)
    void knn5(const float q[3], int* idx, float* d2) const {
        Knn5 rs;
        if (n_ > 0) {
            float dists[3] = { 0, 0, 0 }; float mind = 0.f;
            for (int c = 0; c < 3; c++) {
                if (q[c] < lo_[c]) { dists[c] = (q[c] - lo_[c]) * (q[c] - lo_[c]); mind += dists[c]; }
                if (q[c] > hi_[c]) { dists[c] = (q[c] - hi_[c]) * (q[c] - hi_[c]); mind += dists[c]; }
            }
            search(0, q, mind, dists, rs);
        }
        for (int k = 0; k < 5; k++) { idx[k] = rs.id[k]; d2[k] = rs.d[k]; }
    }
private:
    struct Node { int left, right; int feat; float divlow, divhigh; int lo, hi; };   // leaf: left = -1, points [lo,hi)
    int divide(int lo, int hi, float* bl, float* bh) {
        int me = (int)nodes_.size();
        nodes_.push_back(Node{ -1, -1, 0, 0.f, 0.f, lo, hi });
        if (hi - lo <= 15) return me;
        // exact bbox of this subset, split the widest axis at the median
        float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
        for (int k = lo; k < hi; k++) {
            const P4& p = src_[ids_[k]]; const float v[3] = { p.x, p.y, p.z };
            for (int c = 0; c < 3; c++) { mn[c] = std::min(mn[c], v[c]); mx[c] = std::max(mx[c], v[c]); }
        }
        int feat = 0; float span = mx[0] - mn[0];
        for (int c = 1; c < 3; c++) if (mx[c] - mn[c] > span) { span = mx[c] - mn[c]; feat = c; }
        if (!(span > 0.f)) return me;   // all points identical: keep as one (large) leaf
        int mid = (lo + hi) / 2;
        auto key = [&](int id) { const P4& p = src_[id]; return feat == 0 ? p.x : (feat == 1 ? p.y : p.z); };
        std::nth_element(ids_.begin() + lo, ids_.begin() + mid, ids_.begin() + hi,
                         [&](int a, int b) { float ka = key(a), kb = key(b); return ka < kb || (ka == kb && a < b); });
        float divlow = -FLT_MAX, divhigh = FLT_MAX;
        for (int k = lo; k < mid; k++) divlow = std::max(divlow, key(ids_[k]));
        for (int k = mid; k < hi; k++) divhigh = std::min(divhigh, key(ids_[k]));
        (void)bl; (void)bh;
        int l = divide(lo, mid, bl, bh);
        int r = divide(mid, hi, bl, bh);
        nodes_[me].left = l; nodes_[me].right = r; nodes_[me].feat = feat;
        nodes_[me].divlow = divlow; nodes_[me].divhigh = divhigh;
        return me;
    }
    inline float coord(int id, int c) const { const P4& p = src_[id]; return c == 0 ? p.x : (c == 1 ? p.y : p.z); }
    // FLANN KDTreeSingleIndex::divideTree + middleSplit_ + planeSplit.  bl / bh = the node's bounding box: on entry the box
    // handed down by the parent, on exit the tight box of the points below (as FLANN updates it).
    int divide_flann(int lo, int hi, float* bl, float* bh) {
        const int me = (int)nodes_.size();
        nodes_.push_back(Node{ -1, -1, 0, 0.f, 0.f, lo, hi });
        const int count = hi - lo;
        int* ind = ids_.data() + lo;
        if (count <= 15) {                                   // leaf: its box is the extent of its points
            for (int c = 0; c < 3; c++) { bl[c] = bh[c] = coord(ind[0], c); }
            for (int k = 1; k < count; k++) for (int c = 0; c < 3; c++) { const float v = coord(ind[k], c); if (v < bl[c]) bl[c] = v; if (v > bh[c]) bh[c] = v; }
            return me;
        }
        const float EPS = 0.00001f;
        float max_span = bh[0] - bl[0];
        for (int c = 1; c < 3; c++) { const float span = bh[c] - bl[c]; if (span > max_span) max_span = span; }
        float max_spread = -1.f; int cutfeat = 0;
        for (int c = 0; c < 3; c++) {
            const float span = bh[c] - bl[c];
            if (span > (float)((1 - EPS) * max_span)) {
                float mn = coord(ind[0], c), mx = mn;
                for (int k = 1; k < count; k++) { const float v = coord(ind[k], c); if (v < mn) mn = v; if (v > mx) mx = v; }
                const float spread = mx - mn;
                if (spread > max_spread) { cutfeat = c; max_spread = spread; }
            }
        }
        const float split_val = (bl[cutfeat] + bh[cutfeat]) / 2;
        float mn = coord(ind[0], cutfeat), mx = mn;
        for (int k = 1; k < count; k++) { const float v = coord(ind[k], cutfeat); if (v < mn) mn = v; if (v > mx) mx = v; }
        const float cutval = split_val < mn ? mn : (split_val > mx ? mx : split_val);
        // planeSplit: [0, lim1) < cutval, [lim1, lim2) == cutval, [lim2, count) > cutval
        int left = 0, right = count - 1;
        for (;;) {
            while (left <= right && coord(ind[left], cutfeat) < cutval) ++left;
            while (left <= right && coord(ind[right], cutfeat) >= cutval) --right;
            if (left > right) break;
            std::swap(ind[left], ind[right]); ++left; --right;
        }
        const int lim1 = left;
        right = count - 1;
        for (;;) {
            while (left <= right && coord(ind[left], cutfeat) <= cutval) ++left;
            while (left <= right && coord(ind[right], cutfeat) > cutval) --right;
            if (left > right) break;
            std::swap(ind[left], ind[right]); ++left; --right;
        }
        const int lim2 = left;
        int index = lim1 > count / 2 ? lim1 : (lim2 < count / 2 ? lim2 : count / 2);
        if (index <= 0 || index >= count) return me;         // every point identical in every wide dimension: keep one (large) leaf
        float lbl[3] = { bl[0], bl[1], bl[2] }, lbh[3] = { bh[0], bh[1], bh[2] }, rbl[3] = { bl[0], bl[1], bl[2] }, rbh[3] = { bh[0], bh[1], bh[2] };
        lbh[cutfeat] = cutval; rbl[cutfeat] = cutval;
        const int l = divide_flann(lo, lo + index, lbl, lbh);
        const int r = divide_flann(lo + index, hi, rbl, rbh);
        nodes_[me].left = l; nodes_[me].right = r; nodes_[me].feat = cutfeat;
        nodes_[me].divlow = lbh[cutfeat]; nodes_[me].divhigh = rbl[cutfeat];
        for (int c = 0; c < 3; c++) { bl[c] = std::min(lbl[c], rbl[c]); bh[c] = std::max(lbh[c], rbh[c]); }
        return me;
    }
    void search(int ni, const float q[3], float mindistsq, float* dists, Knn5& rs) const {
        const Node& nd = nodes_[ni];
        if (nd.left < 0) {
            for (int k = nd.lo; k < nd.hi; k++) rs.offer(l2_simple(q, &xyz_[3 * (size_t)k]), ids_[k]);
            return;
        }
        int f = nd.feat; float val = q[f];
        float diff1 = val - nd.divlow, diff2 = val - nd.divhigh;
        int best, other; float cut;
        if (diff1 + diff2 < 0) { best = nd.left; other = nd.right; cut = diff2 * diff2; }
        else { best = nd.right; other = nd.left; cut = diff1 * diff1; }
        search(best, q, mindistsq, dists, rs);
        float dst = dists[f];
        float md = mindistsq + cut - dst;
        dists[f] = cut;
        // conservative pruning: rounding slack so no candidate that could tie or win is skipped
        if (md * 0.9999f <= rs.worst()) search(other, q, md, dists, rs);
        dists[f] = dst;
    }
    int n_ = 0;
    const P4* src_ = nullptr;
    std::vector<float> xyz_; std::vector<int> ids_; std::vector<Node> nodes_;
    float lo_[3], hi_[3];
};

static inline void brute_knn5(const P4* map, int n, const float q[3], int* idx, float* d2) {
    Knn5 rs;
    for (int k = 0; k < n; k++) { const float p[3] = { map[k].x, map[k].y, map[k].z }; rs.offer(l2_simple(q, p), k); }
    for (int k = 0; k < 5; k++) { idx[k] = rs.id[k]; d2[k] = rs.d[k]; }
}

}  // namespace orc
