// bvh.cuh — binary32 BVH traversal for the integrators (EXTENSION: triangle meshes; layout in bvh.h).
// Stack-based while-while traversal: one 64-byte node fetch (4 x 16-byte read-only loads) decides both children,
// the nearer child is followed first and the farther one pushed.  Triangles: Moeller-Trumbore, two-sided,
// t in [tmin, best); among triangles of equal t the lowest triangle id wins (so the result does not depend on the
// traversal order and equals a brute-force scan).
#pragma once

namespace ptb {

constexpr int kTriBit = 0x40000000;        // hit id = kTriBit | slot in the leaf-ordered triangle array

constexpr int kTravStack = 40;
constexpr int kTravStride = kTravStack + 4;     // ints of global scratch per path slot: sp, cur, best_tri, -, stack[]
#ifndef PTB_BVH_STEP_BUDGET
#define PTB_BVH_STEP_BUDGET 24                 // node / leaf visits per ray and wavefront iteration
#endif

// Returns true when the traversal is complete.  With `save` != nullptr the traversal stops after `budget` node / leaf
// visits, writes its state (stack, current node, best triangle id) to save[] and returns false; called again with
// resume = true (and best / bid as they were left) it continues where it stopped.  A ray that grazes the mesh needs
// hundreds of dependent node fetches; without the budget every other slot of the CTA waits for it at the phase barrier.
template <bool STATS>
__device__ __forceinline__ bool bvh_closest(const float4* __restrict__ nodes, const float4* __restrict__ tris, const RayK& r,
                                            float tmin, float& best, int& bid, unsigned long long* st,
                                            int* __restrict__ save = nullptr, bool resume = false, int budget = 0x7fffffff) {
    int stack[kTravStack];
    int sp = 0;
    int cur = 0;                            // root is always an inner node
    int best_tri = -1;                      // triangle id of the current best hit, -1 while it is an analytic object
    if (resume) {
        sp = save[0]; cur = save[1]; best_tri = save[2];
        for (int i = 0; i < sp; ++i) stack[i] = save[4 + i];
    }
    constexpr int kDone = (int)0x80000000;  // not a valid leaf link
    for (;;) {
        // while-while: every lane first walks inner nodes until it stands on a leaf (or is finished); the warp
        // reconverges behind this loop, so the triangle tests below run with all the lanes that found a leaf
        // instead of one lane at a time interleaved with the others' node tests.
        while (cur >= 0) {
            if (budget-- <= 0) {            // only reachable with save != nullptr (the default budget never runs out)
                save[0] = sp; save[1] = cur; save[2] = best_tri;
                for (int i = 0; i < sp; ++i) save[4 + i] = stack[i];
                return false;
            }
            const float4 q0 = __ldg(nodes + 4 * cur), q1 = __ldg(nodes + 4 * cur + 1), q2 = __ldg(nodes + 4 * cur + 2),
                         q3 = __ldg(nodes + 4 * cur + 3);
            if (STATS) st[ST_BVH_NODES]++;
            float ax = fmaf(q0.x, r.inv.x, -r.oi.x), bx = fmaf(q0.w, r.inv.x, -r.oi.x);
            float ay = fmaf(q0.y, r.inv.y, -r.oi.y), by = fmaf(q1.x, r.inv.y, -r.oi.y);
            float az = fmaf(q0.z, r.inv.z, -r.oi.z), bz = fmaf(q1.y, r.inv.z, -r.oi.z);
            const float n0 = fmaxf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)), tmin);
            const float f0 = fminf(fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)), best);
            ax = fmaf(q1.z, r.inv.x, -r.oi.x); bx = fmaf(q2.y, r.inv.x, -r.oi.x);
            ay = fmaf(q1.w, r.inv.y, -r.oi.y); by = fmaf(q2.z, r.inv.y, -r.oi.y);
            az = fmaf(q2.x, r.inv.z, -r.oi.z); bz = fmaf(q2.w, r.inv.z, -r.oi.z);
            const float n1 = fmaxf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)), tmin);
            const float f1 = fminf(fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)), best);
            const bool h0 = f0 >= n0, h1 = f1 >= n1;
            const int c0 = __float_as_int(q3.x), c1 = __float_as_int(q3.y);
            if (h0 && h1) {
                const bool first0 = n0 <= n1;
                if (sp < kTravStack) stack[sp++] = first0 ? c1 : c0;
                cur = first0 ? c0 : c1;
            } else if (h0) cur = c0;
            else if (h1) cur = c1;
            else cur = sp ? stack[--sp] : kDone;
        }
        if (cur == kDone) return true;
        {
            const int link = ~cur, first = link >> 2, cnt = (link & 3) + 1;
            for (int k = 0; k < cnt; ++k) {
                const float4 a = __ldg(tris + 3 * (first + k)), b = __ldg(tris + 3 * (first + k) + 1), c = __ldg(tris + 3 * (first + k) + 2);
                if (STATS) st[ST_BVH_TRIS]++;
                const float px = r.d.y * c.z - r.d.z * c.y, py = r.d.z * c.x - r.d.x * c.z, pz = r.d.x * c.y - r.d.y * c.x;   // d x e2
                const float det = b.x * px + b.y * py + b.z * pz;
                if (det == 0.0f) continue;
                const float idet = rcp_(det);
                const float tx = r.o.x - a.x, ty = r.o.y - a.y, tz = r.o.z - a.z;
                const float u = (tx * px + ty * py + tz * pz) * idet;
                if (u < 0.0f || u > 1.0f) continue;
                const float qx = ty * b.z - tz * b.y, qy = tz * b.x - tx * b.z, qz = tx * b.y - ty * b.x;                       // tv x e1
                const float v = (r.d.x * qx + r.d.y * qy + r.d.z * qz) * idet;
                if (v < 0.0f || u + v > 1.0f) continue;
                const float t = (c.x * qx + c.y * qy + c.z * qz) * idet;
                if (t < tmin || t > best) continue;
                const int id = __float_as_int(a.w);
                if (t < best || (best_tri >= 0 && id < best_tri)) { best = t; bid = kTriBit | (first + k); best_tri = id; }
            }
            cur = sp ? stack[--sp] : kDone;
        }
    }
}

// Surface of a triangle hit: point, geometric normal flipped against the ray (setFaceNormal, objects.go:17-24), frontFace.
__device__ __forceinline__ void tri_surface(const float4* __restrict__ tris, int slot, F3 o, F3 d, float t, F3& p, F3& n, bool& front, int& meta) {
    const float4 b = __ldg(tris + 3 * slot + 1), c = __ldg(tris + 3 * slot + 2);
    p = f3(o.x + d.x * t, o.y + d.y * t, o.z + d.z * t);
    F3 g = f3(b.y * c.z - b.z * c.y, b.z * c.x - b.x * c.z, b.x * c.y - b.y * c.x);    // e1 x e2
    g = unit3(g);
    front = dot3(d, g) < 0.0f;
    n = front ? g : f3(-g.x, -g.y, -g.z);
    meta = __float_as_int(b.w);
}

}  // namespace ptb
