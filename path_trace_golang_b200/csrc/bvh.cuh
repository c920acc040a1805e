// bvh.cuh — binary32 BVH traversal for the integrators (EXTENSION: triangle meshes; layout in bvh.h).
// Stack-based while-while traversal of the 4-wide tree: one 128-byte node fetch (7 x 16-byte read-only loads) decides four
// children; they are sorted by entry distance, the nearest is followed, the others pushed farthest first.  Triangles: Moeller-Trumbore, two-sided,
// t in [tmin, best); among triangles of equal t the lowest triangle id wins (so the result does not depend on the
// traversal order and equals a brute-force scan).
#pragma once

namespace ptb {

// One 256-bit read-only load (LDG.E.256.CONSTANT, sm_100): two adjacent float4 at a 32-byte aligned address.
__device__ __forceinline__ void ld256(const float4* __restrict__ p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
constexpr int kTriQuads = 4;               // float4 per triangle record (bvh.h)

constexpr int kTriBit = 0x40000000;        // hit id = kTriBit | slot in the leaf-ordered triangle array

constexpr int kTravStack = 40;
constexpr int kTravStride = kTravStack + 4;     // ints of global scratch per path slot: sp, cur, best_tri, -, stack[]
#ifndef PTB_BVH_STEP_BUDGET
#define PTB_BVH_STEP_BUDGET 16                 // inner-node visits per ray and wavefront iteration
#endif
#ifndef PTB_BVH_ROUND_NODES
#define PTB_BVH_ROUND_NODES 4                  // inner-node visits between two refills of a warp's idle lanes
#endif

// a*b - c*d and a 3-term dot product with a fixed rounding sequence.
__device__ __forceinline__ float xmy(float a, float b, float c, float d) { return fmaf(a, b, -__fmul_rn(c, d)); }
__device__ __forceinline__ float dot3x(float ax, float ay, float az, float bx, float by, float bz) { return fmaf(az, bz, fmaf(ay, by, __fmul_rn(ax, bx))); }

// Traversal state of one ray.  A traversal can be stopped after any round and continued later, by any thread: the state
// goes to kTravStride ints of global scratch (trav_save / trav_begin with resume).  A ray that grazes the mesh needs
// hundreds of dependent node fetches; the wavefront kernel therefore gives every ray a step budget per iteration and lets
// the lanes of a warp pick up new rays as soon as theirs finish (wavefront.cuh).
struct TravState {
    int sp, cur, best_tri, budget;
    int stack[kTravStack];
};
constexpr int kTravDone = (int)0x80000000;     // `cur` value of a finished traversal (not a valid leaf link)

__device__ __forceinline__ void trav_begin(TravState& T, const int* __restrict__ save, bool resume, int budget) {
    T.sp = 0; T.cur = 0; T.best_tri = -1; T.budget = budget;     // root is always an inner node; best_tri = -1: best hit is analytic
    if (resume) {
        T.sp = save[0]; T.cur = save[1]; T.best_tri = save[2];
        for (int i = 0; i < T.sp; ++i) T.stack[i] = save[4 + i];
    }
}
__device__ __forceinline__ void trav_save(const TravState& T, int* __restrict__ save) {
    save[0] = T.sp; save[1] = T.cur; save[2] = T.best_tri;
    for (int i = 0; i < T.sp; ++i) save[4 + i] = T.stack[i];
}

// One round: up to max_nodes inner-node visits (while-while: the warp reconverges behind the node loop, so the triangle
// tests run with all the lanes that reached a leaf), then the triangles of the leaf the ray set aside on the way, if any.
// Returns 0 = in progress, 1 = finished, 2 = out of budget (T.cur is an inner node; trav_save it to continue later).
template <bool STATS>
__device__ __forceinline__ int trav_round(const float4* __restrict__ nodes, const float4* __restrict__ tris, const RayK& r, float tmin,
                                          float& best, int& bid, unsigned long long* st, TravState& T, int max_nodes) {
    int cur = T.cur, sp = T.sp;
    // Postponed leaf: the first leaf a ray reaches in a round is set aside and the ray goes on with its next node, so the lanes
    // of a warp stay together in the node loop instead of idling from their first leaf to the end of the round (ncu r02g: the
    // traversal ran with 9.8 of 32 lanes); the triangle tests of all the lanes' postponed leaves then run together.
    int pend = 0;                              // 0 = none (node 0 is the root, never a leaf link)
    for (;;) {
        if (cur < 0 && cur != kTravDone && pend == 0) { pend = cur; cur = sp ? T.stack[--sp] : kTravDone; }
        if (cur < 0 || max_nodes <= 0 || T.budget <= 0) break;
        --max_nodes; --T.budget;
        const float4* nd = nodes + 8 * cur;                 // 128-byte node: four 256-bit read-only loads
        float4 q0, q1, q2, q3, q4, q5, q6, q7;
        ld256(nd, q0, q1); ld256(nd + 2, q2, q3); ld256(nd + 4, q4, q5); ld256(nd + 6, q6, q7);
        if (STATS) st[ST_BVH_NODES]++;
        // children as (centre, half extent): near/far of an axis are (c - o)/d -+ h/|d| (see hit_box); a child that is missed
        // (or unused: h = -1) gets the distance +inf
        float d[4];
        int l[4] = {__float_as_int(q6.x), __float_as_int(q6.y), __float_as_int(q6.z), __float_as_int(q6.w)};
        {
            float cx = fmaf(q0.x, r.inv.x, -r.oi.x), cy = fmaf(q0.y, r.inv.y, -r.oi.y), cz = fmaf(q0.z, r.inv.z, -r.oi.z);
            float n = fmaxf(fmaxf(fmaxf(fmaf(-q0.w, r.ainv.x, cx), fmaf(-q1.x, r.ainv.y, cy)), fmaf(-q1.y, r.ainv.z, cz)), tmin);
            float f = fminf(fminf(fminf(fmaf(q0.w, r.ainv.x, cx), fmaf(q1.x, r.ainv.y, cy)), fmaf(q1.y, r.ainv.z, cz)), best);
            d[0] = f >= n ? n : __int_as_float(0x7f800000);
            cx = fmaf(q1.z, r.inv.x, -r.oi.x); cy = fmaf(q1.w, r.inv.y, -r.oi.y); cz = fmaf(q2.x, r.inv.z, -r.oi.z);
            n = fmaxf(fmaxf(fmaxf(fmaf(-q2.y, r.ainv.x, cx), fmaf(-q2.z, r.ainv.y, cy)), fmaf(-q2.w, r.ainv.z, cz)), tmin);
            f = fminf(fminf(fminf(fmaf(q2.y, r.ainv.x, cx), fmaf(q2.z, r.ainv.y, cy)), fmaf(q2.w, r.ainv.z, cz)), best);
            d[1] = f >= n ? n : __int_as_float(0x7f800000);
            cx = fmaf(q3.x, r.inv.x, -r.oi.x); cy = fmaf(q3.y, r.inv.y, -r.oi.y); cz = fmaf(q3.z, r.inv.z, -r.oi.z);
            n = fmaxf(fmaxf(fmaxf(fmaf(-q3.w, r.ainv.x, cx), fmaf(-q4.x, r.ainv.y, cy)), fmaf(-q4.y, r.ainv.z, cz)), tmin);
            f = fminf(fminf(fminf(fmaf(q3.w, r.ainv.x, cx), fmaf(q4.x, r.ainv.y, cy)), fmaf(q4.y, r.ainv.z, cz)), best);
            d[2] = f >= n ? n : __int_as_float(0x7f800000);
            cx = fmaf(q4.z, r.inv.x, -r.oi.x); cy = fmaf(q4.w, r.inv.y, -r.oi.y); cz = fmaf(q5.x, r.inv.z, -r.oi.z);
            n = fmaxf(fmaxf(fmaxf(fmaf(-q5.y, r.ainv.x, cx), fmaf(-q5.z, r.ainv.y, cy)), fmaf(-q5.w, r.ainv.z, cz)), tmin);
            f = fminf(fminf(fminf(fmaf(q5.y, r.ainv.x, cx), fmaf(q5.z, r.ainv.y, cy)), fmaf(q5.w, r.ainv.z, cz)), best);
            d[3] = f >= n ? n : __int_as_float(0x7f800000);
        }
        // sort the four (distance, link) pairs by distance (5 compare-exchanges): the nearest child is visited next, the others are
        // pushed farthest first, so the nearer ones are popped — and shrink `best` — before the farther ones are looked at
#define PTB_CE(a, b) { const bool sw = d[b] < d[a]; const float da = sw ? d[b] : d[a], db = sw ? d[a] : d[b]; \
                       const int la = sw ? l[b] : l[a], lb = sw ? l[a] : l[b]; d[a] = da; d[b] = db; l[a] = la; l[b] = lb; }
        PTB_CE(0, 1) PTB_CE(2, 3) PTB_CE(0, 2) PTB_CE(1, 3) PTB_CE(1, 2)
#undef PTB_CE
        const float inf = __int_as_float(0x7f800000);
        // the builder bounds the pending entries (BvhBuildOutput::max_stack < kTravStack, checked at upload), so the stack
        // cannot overflow; should that bound ever regress, dropped subtrees are counted (ptb_stats.bvh_stack_overflows, STATS builds)
        if (d[3] < inf) { if (sp < kTravStack) T.stack[sp++] = l[3]; else if (STATS) st[ST_BVH_STACK_OVERFLOW]++; }
        if (d[2] < inf) { if (sp < kTravStack) T.stack[sp++] = l[2]; else if (STATS) st[ST_BVH_STACK_OVERFLOW]++; }
        if (d[1] < inf) { if (sp < kTravStack) T.stack[sp++] = l[1]; else if (STATS) st[ST_BVH_STACK_OVERFLOW]++; }
        cur = d[0] < inf ? l[0] : (sp ? T.stack[--sp] : kTravDone);
    }
    if (pend != 0) {
        const int link = ~pend, first = link >> 2, cnt = (link & 3) + 1;
        for (int k = 0; k < cnt; ++k) {
            float4 a, b, c, pad_;
            ld256(tris + kTriQuads * (first + k), a, b); ld256(tris + kTriQuads * (first + k) + 2, c, pad_);
            if (STATS) st[ST_BVH_TRIS]++;
            // (products and sums spelled out: the contraction ptxas would pick for a*b - c*d is not the same in every
            // instantiation of the kernel, and the STATS build must trace exactly the paths of the plain one)
            const float px = xmy(r.d.y, c.z, r.d.z, c.y), py = xmy(r.d.z, c.x, r.d.x, c.z), pz = xmy(r.d.x, c.y, r.d.y, c.x);   // d x e2
            const float det = dot3x(b.x, b.y, b.z, px, py, pz);
            if (det == 0.0f) continue;
            const float idet = rcp_(det);
            const float tx = r.o.x - a.x, ty = r.o.y - a.y, tz = r.o.z - a.z;
            const float u = __fmul_rn(dot3x(tx, ty, tz, px, py, pz), idet);
            if (u < 0.0f || u > 1.0f) continue;
            const float qx = xmy(ty, b.z, tz, b.y), qy = xmy(tz, b.x, tx, b.z), qz = xmy(tx, b.y, ty, b.x);                       // tv x e1
            const float v = __fmul_rn(dot3x(r.d.x, r.d.y, r.d.z, qx, qy, qz), idet);
            if (v < 0.0f || __fadd_rn(u, v) > 1.0f) continue;
            const float t = __fmul_rn(dot3x(c.x, c.y, c.z, qx, qy, qz), idet);
            if (t < tmin || t > best) continue;
            const int id = __float_as_int(a.w);
            if (t < best || (T.best_tri >= 0 && id < T.best_tri)) { best = t; bid = kTriBit | (first + k); T.best_tri = id; }
        }
    }
    // (cur may be a second leaf reached in this round: it is postponed first thing in the next one)
    T.cur = cur; T.sp = sp;
    if (cur == kTravDone) return 1;
    if (cur >= 0 && T.budget <= 0) return 2;
    return 0;
}

// Whole traversal in one call (megakernel-style callers): returns when the closest triangle hit, if any, is in best / bid.
template <bool STATS>
__device__ __forceinline__ void bvh_closest(const float4* __restrict__ nodes, const float4* __restrict__ tris, const RayK& r,
                                            float tmin, float& best, int& bid, unsigned long long* st) {
    TravState T;
    trav_begin(T, nullptr, false, 0x7fffffff);
    while (trav_round<STATS>(nodes, tris, r, tmin, best, bid, st, T, 0x7fffffff) == 0) {}
}

// Surface of a triangle hit: point, geometric normal flipped against the ray (setFaceNormal, objects.go:17-24), frontFace.
__device__ __forceinline__ void tri_surface(const float4* __restrict__ tris, int slot, F3 o, F3 d, float t, F3& p, F3& n, bool& front, int& meta) {
    const float4 b = __ldg(tris + kTriQuads * slot + 1), c = __ldg(tris + kTriQuads * slot + 2);
    p = f3(fmaf(d.x, t, o.x), fmaf(d.y, t, o.y), fmaf(d.z, t, o.z));
    F3 g = f3(xmy(b.y, c.z, b.z, c.y), xmy(b.z, c.x, b.x, c.z), xmy(b.x, c.y, b.y, c.x));    // e1 x e2
    {   // unit3 and dot3 with the rounding sequence fixed (see trav_round)
        const float l = sqrt_(dot3x(g.x, g.y, g.z, g.x, g.y, g.z));
        if (l != 0.0f) { const float il = rcp_(l); g = f3(__fmul_rn(g.x, il), __fmul_rn(g.y, il), __fmul_rn(g.z, il)); }
    }
    front = dot3x(d.x, d.y, d.z, g.x, g.y, g.z) < 0.0f;
    n = front ? g : f3(-g.x, -g.y, -g.z);
    meta = __float_as_int(b.w);
}

}  // namespace ptb
