// bvh.cpp — binned-SAH BVH builder (host, C++, multi-threaded over subtrees): a binary tree by binned SAH, collapsed to the
// 4-wide nodes of bvh.h when it is flattened.
#include "bvh.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <future>
#include <limits>
#include <memory>

namespace ptb {
namespace {

struct Box {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    void grow(const float* p) { for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); } }
    void grow(const Box& b) { for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); } }
    double area() const {
        double dx = (double)hi[0] - lo[0], dy = (double)hi[1] - lo[1], dz = (double)hi[2] - lo[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0;
        return 2.0 * (dx * dy + dy * dz + dz * dx);
    }
};

struct BuildNode {
    Box box;
    int64_t first = 0, count = 0;            // leaf: range in the permuted index array
    std::unique_ptr<BuildNode> child[2];
};

struct Builder {
    const float* verts;
    std::vector<Box> tbox;                   // per triangle
    std::vector<float> cent;                 // 3 per triangle
    std::vector<int64_t> idx;                // permutation
    std::atomic<int> tasks{0};
    int max_tasks = 1;

    std::unique_ptr<BuildNode> build(int64_t first, int64_t count, int depth) {
        auto node = std::make_unique<BuildNode>();
        Box cb;
        for (int64_t i = first; i < first + count; i++) { node->box.grow(tbox[idx[i]]); cb.grow(&cent[3 * idx[i]]); }
        node->first = first; node->count = count;
        if (count <= kMaxLeafTris) return node;
        // binned SAH over the three axes of the centroid box
        constexpr int kBins = 16;
        int best_axis = -1, best_split = -1;
        double best_cost = std::numeric_limits<double>::infinity();
        for (int ax = 0; ax < 3; ax++) {
            const float lo = cb.lo[ax], ext = cb.hi[ax] - cb.lo[ax];
            if (!(ext > 0)) continue;
            Box bb[kBins]; int64_t bn[kBins] = {0};
            const float scale = kBins / ext;
            for (int64_t i = first; i < first + count; i++) {
                int b = (int)((cent[3 * idx[i] + ax] - lo) * scale);
                b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                bb[b].grow(tbox[idx[i]]); bn[b]++;
            }
            double right_area[kBins]; int64_t right_n[kBins];
            Box acc; int64_t n = 0;
            for (int b = kBins - 1; b > 0; b--) { acc.grow(bb[b]); n += bn[b]; right_area[b] = acc.area(); right_n[b] = n; }
            Box accl; int64_t nl = 0;
            for (int b = 0; b < kBins - 1; b++) {
                accl.grow(bb[b]); nl += bn[b];
                if (nl == 0 || right_n[b + 1] == 0) continue;
                double cost = accl.area() * (double)nl + right_area[b + 1] * (double)right_n[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = ax; best_split = b; }
            }
        }
        int64_t mid;
        int lg = 0;
        while ((1ll << lg) < count) lg++;
        if (best_axis >= 0 && depth + lg >= 35) {
            // SAH on overlapping triangle soup can peel off a few triangles per level; the traversal stack is finite
            // (api.cu rejects depth > 38), so near the limit split at the centroid median of the widest axis instead:
            // from here on depth + log2(count) no longer grows
            int ax = 0;
            for (int k = 1; k < 3; k++) if (cb.hi[k] - cb.lo[k] > cb.hi[ax] - cb.lo[ax]) ax = k;
            std::nth_element(idx.begin() + first, idx.begin() + first + count / 2, idx.begin() + first + count,
                             [&](int64_t a, int64_t b) { return cent[3 * a + ax] < cent[3 * b + ax]; });
            mid = first + count / 2;
        } else if (best_axis < 0) {          // all centroids coincide: split in the middle
            mid = first + count / 2;
        } else {
            const float lo = cb.lo[best_axis], scale = kBins / (cb.hi[best_axis] - cb.lo[best_axis]);
            auto it = std::partition(idx.begin() + first, idx.begin() + first + count, [&](int64_t t) {
                int b = (int)((cent[3 * t + best_axis] - lo) * scale);
                b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                return b <= best_split;
            });
            mid = it - idx.begin();
            if (mid == first || mid == first + count) mid = first + count / 2;
        }
        const int64_t nl = mid - first, nr = count - nl;
        if (count > 100000 && tasks.load() < max_tasks) {       // big subtree: build the left half on another thread
            tasks++;
            auto fut = std::async(std::launch::async, [&, first, nl, depth] { return build(first, nl, depth + 1); });
            node->child[1] = build(mid, nr, depth + 1);
            node->child[0] = fut.get();
            tasks--;
        } else {
            node->child[0] = build(first, nl, depth + 1);
            node->child[1] = build(mid, nr, depth + 1);
        }
        return node;
    }
};

inline int32_t f2i(float f) { int32_t i; std::memcpy(&i, &f, 4); return i; }
inline float i2f(int32_t i) { float f; std::memcpy(&f, &i, 4); return f; }

}  // namespace

void build_bvh(const BvhBuildInput& in, BvhBuildOutput& out, int threads) {
    auto t0 = std::chrono::steady_clock::now();
    out.nodes.clear(); out.tris.clear(); out.max_depth = 0; out.sah_cost = 0;
    const int64_t n = in.n_tri;
    if (n <= 0) return;
    Builder B;
    B.verts = in.tri_vertices;
    B.max_tasks = threads > 1 ? threads - 1 : 0;
    B.tbox.resize(n); B.cent.resize(3 * n); B.idx.resize(n);
    float amax = 1.0f;
    for (int64_t t = 0; t < n; t++) {
        const float* v = in.tri_vertices + 9 * t;
        for (int k = 0; k < 3; k++) B.tbox[t].grow(v + 3 * k);
        for (int k = 0; k < 3; k++) B.cent[3 * t + k] = (v[k] + v[3 + k] + v[6 + k]) * (1.0f / 3.0f);
        for (int k = 0; k < 9; k++) amax = std::max(amax, std::fabs(v[k]));
        B.idx[t] = t;
    }
    const float pad = 1e-5f * amax;
    std::unique_ptr<BuildNode> root = B.build(0, n, 0);

    // ---- flatten: every BuildNode with children becomes one 64-byte node; leaves become ranges of out.tris
    out.tris.resize(n);
    for (int64_t i = 0; i < n; i++) {
        const int64_t t = B.idx[i];
        const float* v = in.tri_vertices + 9 * t;
        BvhTri& o = out.tris[i];
        o.q[0] = v[0]; o.q[1] = v[1]; o.q[2] = v[2]; o.q[3] = i2f((int32_t)t);
        o.q[4] = v[3] - v[0]; o.q[5] = v[4] - v[1]; o.q[6] = v[5] - v[2]; o.q[7] = i2f(in.tri_meta[t]);
        o.q[8] = v[6] - v[0]; o.q[9] = v[7] - v[1]; o.q[10] = v[8] - v[2]; o.q[11] = i2f(in.tri_world[t]);
        o.q[12] = o.q[13] = o.q[14] = o.q[15] = 0.0f;
    }
    struct Item { const BuildNode* n; int32_t slot; int depth; int pending; };
    std::vector<Item> stack;
    out.nodes.reserve((size_t)(n / 4 + 16));
    const double root_area = std::max(root->box.area(), 1e-30);
    out.max_stack = 0;
    auto child_link = [&](const BuildNode* c, int depth, int pending) -> int32_t {
        if (c->child[0]) {                                   // inner: allocate its node now, fill it later
            out.nodes.emplace_back();
            int32_t id = (int32_t)out.nodes.size() - 1;
            stack.push_back({c, id, depth, pending});
            return id;
        }
        out.sah_cost += c->box.area() / root_area * (double)c->count;
        return ~(int32_t)((c->first << 2) | (c->count - 1));
    };
    auto put_box = [&](BvhNode& nd, int which, const Box* b) {
        // (centre, half extent) in binary32 such that [c - h, c + h] contains the padded box: the half extent absorbs the
        // rounding of the centre and is rounded up.  An empty child gets h = -1 (far < near on every axis: never hit).
        float c[3] = {0.0f, 0.0f, 0.0f}, h[3] = {-1.0f, -1.0f, -1.0f};
        if (b) for (int k = 0; k < 3; k++) {
            const double lo = (double)b->lo[k] - pad, hi = (double)b->hi[k] + pad;
            c[k] = (float)(0.5 * (lo + hi));
            const double need = std::max(hi - (double)c[k], (double)c[k] - lo);
            h[k] = std::nextafter((float)need, INFINITY);
        }
        float* q = nd.q + 6 * which;
        q[0] = c[0]; q[1] = c[1]; q[2] = c[2]; q[3] = h[0]; q[4] = h[1]; q[5] = h[2];
    };
    // the (up to) kBvhWidth children of the wide node made from binary node b: its two children, then repeatedly the inner
    // one with the largest box replaced by ITS two children
    auto wide_children = [&](const BuildNode* b, const BuildNode* kids[kBvhWidth]) -> int {
        int nk = 2;
        kids[0] = b->child[0].get(); kids[1] = b->child[1].get();
        while (nk < kBvhWidth) {
            int pick = -1;
            double best = -1.0;
            for (int k = 0; k < nk; k++) if (kids[k]->child[0] && kids[k]->box.area() > best) { best = kids[k]->box.area(); pick = k; }
            if (pick < 0) break;
            const BuildNode* p = kids[pick];
            kids[pick] = p->child[0].get();
            kids[nk++] = p->child[1].get();
        }
        return nk;
    };
    out.nodes.emplace_back();
    if (!root->child[0]) {                                   // <= kMaxLeafTris triangles: root with one leaf child
        BvhNode& nd = out.nodes[0];
        std::memset(&nd, 0, sizeof nd);
        put_box(nd, 0, &root->box);
        nd.q[24] = i2f(~(int32_t)((root->first << 2) | (root->count - 1)));
        for (int k = 1; k < kBvhWidth; k++) { put_box(nd, k, nullptr); nd.q[24 + k] = i2f(kEmptyLeaf); }
        out.max_depth = 1;
    } else {
        stack.push_back({root.get(), 0, 1, 0});
        while (!stack.empty()) {
            Item it = stack.back(); stack.pop_back();
            out.max_depth = std::max(out.max_depth, it.depth);
            out.sah_cost += it.n->box.area() / root_area;
            const BuildNode* kids[kBvhWidth];
            const int nk = wide_children(it.n, kids);
            out.max_stack = std::max(out.max_stack, it.pending + nk - 1);
            int32_t link[kBvhWidth];
            for (int k = 0; k < nk; k++) link[k] = child_link(kids[k], it.depth + 1, it.pending + nk - 1);
            BvhNode& nd = out.nodes[it.slot];                // (re-fetch: emplace_back may have moved the array)
            std::memset(&nd, 0, sizeof nd);
            for (int k = 0; k < kBvhWidth; k++) {
                put_box(nd, k, k < nk ? &kids[k]->box : nullptr);
                nd.q[24 + k] = i2f(k < nk ? link[k] : kEmptyLeaf);
            }
        }
    }
    out.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    (void)f2i;
}

}  // namespace ptb
