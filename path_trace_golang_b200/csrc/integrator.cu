// integrator.cu — the per-pixel render loop of internal/engine (renderer.go:163-238 + 286-404) as one
// persistent-lane sm_100a megakernel.
//
// Mapping (DESIGN.md "Integrator kernel"):
//   * one lane = one pixel; it walks that pixel's samples [s_begin, s_end) in order and keeps the fp32
//     radiance sum in registers, so the per-pixel sum order is the reference's (renderer.go:181-187) and
//     no atomics / HBM accumulation traffic exist;
//   * one warp = an 8x4 pixel tile, one CTA = 16x8 pixels (coherent primary rays per warp);
//   * rayColorOpt's recursion (renderer.go:286-404) is run iteratively (throughput beta, radiance L);
//     when a lane's path ends it immediately regenerates the next camera sample of its pixel, so lanes of
//     a warp are at different bounces of different samples but all of them execute the same closest-hit
//     scan over the world, which is where the time goes (renderer.go:297-302);
//   * the warp leaves the loop on a ballot (__any_sync) when every lane has finished its samples;
//   * the scan indexes the world in __constant__ memory with a warp-uniform index (broadcast); the winning
//     object and its material are fetched from a shared-memory copy (divergent index).
//
// Semantics are the reference's, quirks included: un-normalised primary rays,
// tMin=0.001 in ray-parameter units, inclusive tMax for spheres/planes (later object wins ties) and
// exclusive for boxes, box hit from inside returns t=tMin with the nearest-face normal, dielectric exit
// search without refraction on exit, Russian roulette on the REMAINING depth <= 3.
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>
#include <cstdlib>

#include "scene_dev.h"

namespace ptb {

#ifndef PTB_BLOCK_THREADS
#define PTB_BLOCK_THREADS 128
#endif
#ifndef PTB_MIN_BLOCKS
#define PTB_MIN_BLOCKS 8
#endif
#ifndef PTB_WF_MIN_BLOCKS
#define PTB_WF_MIN_BLOCKS 4
#endif

constexpr uint32_t kGolden = 0x9E3779B9u;

// ---- arithmetic policy.  PTB_FAST_MATH=1 (default): reciprocal / square root / sin / cos use the SFU
// approximations (rcp.approx, sqrt.approx, sin/cos.approx: <= 2 ulp, resp. ~1e-6 absolute) and every
// division of the reference becomes a multiplication by such a reciprocal.  The radiance estimate is
// statistical and the parity tests (tests/test_gpu_parity.py) run against this build.  PTB_FAST_MATH=0 keeps
// IEEE division / sqrt and libdevice sincosf/expf.
#ifndef PTB_FAST_MATH
#define PTB_FAST_MATH 1
#endif
__device__ __forceinline__ float rcp_(float x) {
#if PTB_FAST_MATH
    float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
    return 1.0f / x;
#endif
}
__device__ __forceinline__ float sqrt_(float x) {
#if PTB_FAST_MATH
    float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
    return sqrtf(x);
#endif
}
__device__ __forceinline__ void sincos_(float x, float* s, float* c) {
#if PTB_FAST_MATH
    __sincosf(x, s, c);
#else
    sincosf(x, s, c);
#endif
}
__device__ __forceinline__ float exp_(float x) {
#if PTB_FAST_MATH
    return __expf(x);
#else
    return expf(x);
#endif
}

// ---- counter RNG (DESIGN.md "RNG"); stands in for randSource.Float64 (random.go:27-34).
// Draw i of a path is a pure function of (key, i), so the draws a bounce MIGHT need are evaluated up front
// at full warp width (peek) and the counter is advanced afterwards by the number actually consumed — the
// consumed sequence is exactly the sequential one the oracle uses.
__device__ __forceinline__ uint32_t fmix(uint32_t x) {
    x ^= x >> 16; x *= 0x21f0aaadu; x ^= x >> 15; x *= 0x735a2d97u; x ^= x >> 15;
    return x;
}
struct Rng {
    uint32_t key, ctr;
    __device__ __forceinline__ float peek(uint32_t k) const {
        return (float)(fmix(key + (ctr + k) * kGolden) >> 8) * (1.0f / 16777216.0f);
    }
    __device__ __forceinline__ float next() { float v = peek(0); ++ctr; return v; }
};

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 f3(float x, float y, float z) { return F3{x, y, z}; }
__device__ __forceinline__ float dot3(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }   // math.go:19
// unit vector by one reciprocal square root (callers whose input cannot be the zero vector)
__device__ __forceinline__ F3 unit3_nz(F3 a) {
#if PTB_FAST_MATH
    float inv; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(dot3(a, a)));
    return f3(a.x * inv, a.y * inv, a.z * inv);
#else
    const float inv = 1.0f / sqrtf(dot3(a, a));
    return f3(a.x * inv, a.y * inv, a.z * inv);
#endif
}
__device__ __forceinline__ F3 unit3(F3 a) {                                                       // math.go:31-37, 14-17
    float l = sqrt_(dot3(a, a));
    if (l == 0.0f) return a;
    float inv = rcp_(l);
    return f3(a.x * inv, a.y * inv, a.z * inv);
}
__device__ __forceinline__ F3 in_unit_sphere(Rng& rng) {                                          // math.go:66-85
    for (;;) {
        float x = rng.peek(0) * 2.0f - 1.0f;
        float y = rng.peek(1) * 2.0f - 1.0f;
        float z = rng.peek(2) * 2.0f - 1.0f;
        rng.ctr += 3u;
        if (x * x + y * y + z * z >= 1.0f) continue;
        return f3(x, y, z);
    }
}
// randomCosineDirection (math.go:94-131) about axis n with the two uniforms already drawn.
__device__ __forceinline__ F3 cosine_direction(F3 n, float r1, float r2) {
    float phi = 6.28318530717958647692f * r1;
    float ct = sqrt_(r2);
    float st = sqrt_(1.0f - r2);
    // v = unit(n x helper), u = v x n with helper = (0,1,0) if |n.x| > 0.9 else (1,0,0)
    F3 v = (fabsf(n.x) > 0.9f) ? f3(-n.z, 0.0f, n.x) : f3(0.0f, n.z, -n.y);
    v = unit3_nz(v);                   // n x helper is never zero: the helper axis is chosen away from the unit vector n
    F3 u = f3(v.y * n.z - v.z * n.y, v.z * n.x - v.x * n.z, v.x * n.y - v.y * n.x);
    float sp, cp;
    sincos_(phi, &sp, &cp);
    float lx = st * cp, ly = st * sp, lz = ct;
    return f3(lx * u.x + ly * v.x + lz * n.x, lx * u.y + ly * v.y + lz * n.y, lx * u.z + ly * v.z + lz * n.z);
}

// Per-ray constants of the closest-hit tests: a = d.d (objects.go:43), inv = 1/d (objects.go:149-161),
// oi = o*inv (so a slab distance (b - o)*inv is one FFMA: b*inv - oi), inv_a = 1/a (the divisions of
// objects.go:55,57 become multiplications).
// A direction component that is exactly 0 would make the (centre, half extent) slab form of hit_box compute inf - inf = NaN and
// silently drop that axis from the box test (the reference misses such a box when the origin is outside the slab: tNear, tFar
// = -+inf, objects.go:149-165).  Such components (about 2^-24 of all) become +-1e-30 WHERE A DIRECTION IS MADE (camera ray,
// scatter): 1/d = +-1e30 stays finite, every slab distance keeps its sign, hit points and normals are unchanged in binary32 —
// and the closest-hit scan itself stays free of the check.
__device__ __forceinline__ void sanitize_dir(F3& d) {
    if (fminf(fminf(fabsf(d.x), fabsf(d.y)), fabsf(d.z)) == 0.0f) {
        if (d.x == 0.0f) d.x = copysignf(1e-30f, d.x);
        if (d.y == 0.0f) d.y = copysignf(1e-30f, d.y);
        if (d.z == 0.0f) d.z = copysignf(1e-30f, d.z);
    }
}
struct RayK { F3 o, d, inv, oi, ainv; float a, inv_a; };
__device__ __forceinline__ RayK make_ray(F3 o, F3 d) {
    RayK r;
    r.o = o; r.d = d;
    r.a = d.x * d.x + d.y * d.y + d.z * d.z;
    r.inv = f3(rcp_(d.x), rcp_(d.y), rcp_(d.z));
    r.oi = f3(o.x * r.inv.x, o.y * r.inv.y, o.z * r.inv.z);
    r.ainv = f3(fabsf(r.inv.x), fabsf(r.inv.y), fabsf(r.inv.z));
    r.inv_a = rcp_(r.a);
    return r;
}

// box.hit (objects.go:141-183), branch-free: t0 = max(tmin, near_x, near_y, near_z), t1 = min(tmax, far_x, far_y,
// far_z), hit iff t1 > t0 (the per-axis early exits of the reference are equivalent: t0 only grows, t1 only
// shrinks).  near/far of an axis are min/max of the two slab distances, which is what the reference's swap on
// invD < 0 produces.  make_ray keeps 1/d finite, so no slab distance is NaN (the reference's 0 * inf = NaN case —
// origin exactly on a slab plane and that direction component exactly 0 — resolves here as "on the boundary = inside";
// the binary64 parity kernel keeps the reference's exact form).
__device__ __forceinline__ bool hit_box(float4 lo, float4 hi, const RayK& r, float tmin, float tmax, float& t_out) {
#if PTB_FAST_MATH
    // record = (centre c, half extent h >= 0): the slab interval of an axis is (c - o)/d -+ h/|d|, so near and far come
    // out ordered and the per-axis min/max pair (half-rate ALU pipe) disappears: 9 FFMA (|1/d| is an operand modifier)
    // + 2 three-input and 2 two-input min/max.
    const float cx = fmaf(lo.x, r.inv.x, -r.oi.x), cy = fmaf(lo.y, r.inv.y, -r.oi.y), cz = fmaf(lo.z, r.inv.z, -r.oi.z);
    const float nx = fmaf(-hi.x, r.ainv.x, cx), ny = fmaf(-hi.y, r.ainv.y, cy), nz = fmaf(-hi.z, r.ainv.z, cz);
    const float fx = fmaf(hi.x, r.ainv.x, cx), fy = fmaf(hi.y, r.ainv.y, cy), fz = fmaf(hi.z, r.ainv.z, cz);
    const float t0 = fmaxf(fmaxf(fmaxf(nx, ny), nz), tmin);
    const float t1 = fminf(fminf(fminf(fx, fy), fz), tmax);
#else
    const float ax = (lo.x - hi.x - r.o.x) * r.inv.x, bx = (lo.x + hi.x - r.o.x) * r.inv.x;
    const float ay = (lo.y - hi.y - r.o.y) * r.inv.y, by = (lo.y + hi.y - r.o.y) * r.inv.y;
    const float az = (lo.z - hi.z - r.o.z) * r.inv.z, bz = (lo.z + hi.z - r.o.z) * r.inv.z;
    const float t0 = fmaxf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)), tmin);
    const float t1 = fminf(fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)), tmax);
#endif
    t_out = t0;
    return t1 > t0;
}
// sphere.hit (objects.go:37-59), branch-free.  lo = (centre, -), hi = (radius, radius^2, 1/radius, -).
__device__ __forceinline__ bool hit_sphere(float4 lo, float4 hi, const RayK& r, float tmin, float tmax, float& t_out) {
    const float ocx = r.o.x - lo.x, ocy = r.o.y - lo.y, ocz = r.o.z - lo.z;
    const float hb = ocx * r.d.x + ocy * r.d.y + ocz * r.d.z;
    const float c = (ocx * ocx + ocy * ocy + ocz * ocz) - hi.y;
    const float disc = hb * hb - r.a * c;
    const float sq = sqrt_(fmaxf(disc, 0.0f));
    const float r1 = (-hb - sq) * r.inv_a;
    const float r2 = (-hb + sq) * r.inv_a;
    const float root = (r1 < tmin || r1 > tmax) ? r2 : r1;
    t_out = root;
    return !(disc < 0.0f) && !(root < tmin || root > tmax);
}
// Same decisions as hit_sphere from the scan table record (centre, radius^2), fewer ALU-pipe ops: r1 <= r2 always (sq >= 0,
// inv_a > 0), so "r1 if it is in [tmin, tmax] else r2, then range-check" == "root = r1 >= tmin ? r1 : r2; root in range";
// a negative discriminant makes sq, r1, r2 and root NaN and every ordered comparison false.
__device__ __forceinline__ bool hit_sphere4(float cx, float cy, float cz, float r2, const RayK& r, float tmin, float tmax, float& t_out) {
    const float ocx = r.o.x - cx, ocy = r.o.y - cy, ocz = r.o.z - cz;
    const float hb = ocx * r.d.x + ocy * r.d.y + ocz * r.d.z;
    const float c = fmaf(ocz, ocz, fmaf(ocy, ocy, fmaf(ocx, ocx, -r2)));
    const float disc = fmaf(hb, hb, -(r.a * c));
    const float sq = sqrt_(disc);
    const float q1 = (-hb - sq) * r.inv_a;
    const float q2 = (sq - hb) * r.inv_a;
    const float root = (q1 >= tmin) ? q1 : q2;
    t_out = root;
    return root >= tmin && root <= tmax;
}
// plane.hit from the scan table (p.y only; normal (0,1,0)): t = p.y/d.y - o.y/d.y.
__device__ __forceinline__ bool hit_plane1(float py, const RayK& r, float tmin, float tmax, float& t_out) {
    const float t = fmaf(py, r.inv.y, -r.oi.y);
    t_out = t;
    return !(fabsf(r.d.y) < 1e-6f) && t >= tmin && t <= tmax;
}
// plane.hit with normal (0,1,0): denom = d.y, t = (p.y - o.y)/d.y (objects.go:100-110, 251-257).
__device__ __forceinline__ bool hit_plane(float4 lo, const RayK& r, float tmin, float tmax, float& t_out) {
    float t = (lo.y - r.o.y) * r.inv.y;
    t_out = t;
    return !(fabsf(r.d.y) < 1e-6f) && !(t < tmin || t > tmax);
}
// ---- two rays per instruction: sm_100a's packed binary32 arithmetic (PTX fma/add/mul.rn.f32x2 -> SASS FFMA2 / FADD2 / FMUL2 on
// 64-bit register pairs).  The wavefront scan tests the TWO rays of a thread against every record, and the kernel is bound by
// instruction issue (profiles/r02z: issue slots 80.7 % busy, FMA pipe 30 %): with half 0 of a pair = ray 0 and half 1 = ray 1 the
// arithmetic of both rays issues once.  A record field is a scalar from the constant bank; `dup2` makes the pair (x, x), which
// ptxas folds into the instruction as a broadcast uniform operand (URn.F32, negation included) — no extra moves.  Same roundings
// as the scalar forms (each half is an ordinary round-to-nearest FFMA/FADD/FMUL).
// A packed instruction runs at HALF the scalar rate (tools/probes/ffma2_probe.cu), so it frees issue slots, not FMA-pipe time.
// Measured (profiles/r02t_packed.log, r02u_packed.log):
//   * the sphere test (3-register FFMA/FADD/FMUL, then two square roots and the root selection): 141 instead of 192 instructions
//     per 4 spheres x 2 rays; C5 (17 spheres) +5.3 %, C2 (11) +2.4 %, C4 +1.4 % — but C3 (2 spheres) -2.9 %: building the pairs
//     costs more than two sphere tests save.  Hence a kernel instantiation of its own (PK), chosen per scene (kPackedSphereMin);
//   * the slab test, whose scalar FFMAs take a uniform operand and already overlap with its min/max chain, loses when packed
//     (116 instead of 149 instructions per 4 boxes x 2 rays, yet C3 -3 %): boxes stay scalar.
typedef unsigned long long P2;
__device__ __forceinline__ P2 pk2(float a, float b) { P2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ P2 dup2(float a) { return pk2(a, a); }
#pragma nv_diag_suppress 550           // (the other half of an unpacked pair is "set but never used")
__device__ __forceinline__ float lo2(P2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi2(P2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
#pragma nv_diag_default 550
__device__ __forceinline__ P2 fma2(P2 a, P2 b, P2 c) { P2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ P2 mul2(P2 a, P2 b) { P2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ P2 add2(P2 a, P2 b) { P2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ P2 sub2(P2 a, P2 b) { P2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// The per-ray constants the sphere test consumes, for a pair of rays: -o, d, -a, 1/a.
struct RayP { P2 no[3], d[3], na, inv_a; };
__device__ __forceinline__ RayP pair_rays(const RayK& r0, const RayK& r1) {
    RayP p;
    // Origin and direction arrive as scalars (one LDS.128 per ray).  Passing the pairs through one packed instruction (x + -0:
    // exact) makes the PAIR the value ptxas keeps; otherwise it re-packs them from the scalars inside the sphere loop
    // (two moves per operand per trip).
    const P2 z = dup2(-0.0f);
    p.no[0] = add2(pk2(-r0.o.x, -r1.o.x), z); p.no[1] = add2(pk2(-r0.o.y, -r1.o.y), z); p.no[2] = add2(pk2(-r0.o.z, -r1.o.z), z);
    p.d[0] = add2(pk2(r0.d.x, r1.d.x), z); p.d[1] = add2(pk2(r0.d.y, r1.d.y), z); p.d[2] = add2(pk2(r0.d.z, r1.d.z), z);
    p.na = pk2(-r0.a, -r1.a); p.inv_a = pk2(r0.inv_a, r1.inv_a);
    return p;
}
// hit_sphere4 for two rays.  Works with c - o and -(o - c).d (exact negations of the scalar form's oc and hb, so the roots are
// the same binary32 values): 13 packed instructions instead of 30 scalar ones; the two square roots and the root selection
// stay per ray.
__device__ __forceinline__ void hit_sphere_pair(float cx, float cy, float cz, float r2, const RayP& r, float tmin, const float (&tmax)[2],
                                                float (&t_out)[2], bool (&hit)[2]) {
    const P2 ox = add2(dup2(cx), r.no[0]), oy = add2(dup2(cy), r.no[1]), oz = add2(dup2(cz), r.no[2]);     // c - o
    // -hb, contracted as nvcc contracts the scalar `ocx * dx + ocy * dy + ocz * dz`: fma(z, fma(x, (y product)))
    const P2 nhb = fma2(oz, r.d[2], fma2(ox, r.d[0], mul2(oy, r.d[1])));
    const P2 c = fma2(oz, oz, fma2(oy, oy, fma2(ox, ox, dup2(-r2))));
    const P2 disc = fma2(nhb, nhb, mul2(r.na, c));
    const P2 sq = pk2(sqrt_(lo2(disc)), sqrt_(hi2(disc)));
    const P2 q1 = mul2(sub2(nhb, sq), r.inv_a), q2 = mul2(add2(nhb, sq), r.inv_a);
    const float a0 = lo2(q1), a1 = hi2(q1);
    const float root0 = (a0 >= tmin) ? a0 : lo2(q2), root1 = (a1 >= tmin) ? a1 : hi2(q2);
    t_out[0] = root0; t_out[1] = root1;
    hit[0] = root0 >= tmin && root0 <= tmax[0]; hit[1] = root1 >= tmin && root1 <= tmax[1];
}

// Object record as two 16-byte vectors: lo = (a.xyz, meta), hi = (b.xyz, world_idx).
__device__ __forceinline__ float4 obj_lo(const DevObj* objs, int i) { return reinterpret_cast<const float4*>(objs + i)[0]; }
__device__ __forceinline__ float4 obj_hi(const DevObj* objs, int i) { return reinterpret_cast<const float4*>(objs + i)[1]; }
__device__ __forceinline__ bool hit_any(float4 lo, float4 hi, int type, const RayK& r, float tmin, float tmax, float& t) {
    if (type == PTB_OBJ_BOX) return hit_box(lo, hi, r, tmin, tmax, t);
    if (type == PTB_OBJ_SPHERE) return hit_sphere(lo, hi, r, tmin, tmax, t);
    return hit_plane(lo, r, tmin, tmax, t);
}

// Hit point, face normal and frontFace of object `ob` at parameter t (objects.go:61-88, 114-132, 181-221).
__device__ __forceinline__ void surface(const DevObj& ob, int type, F3 o, F3 d, float t, F3& p, F3& n, bool& front) {
    p = f3(o.x + d.x * t, o.y + d.y * t, o.z + d.z * t);
    F3 on;
    if (type == PTB_OBJ_SPHERE) {
        float ir = ob.bz;   // 1/radius
        on = f3((p.x - ob.ax) * ir, (p.y - ob.ay) * ir, (p.z - ob.az) * ir);
    } else if (type == PTB_OBJ_PLANE) {
        on = f3(0.0f, 1.0f, 0.0f);
    } else {
        // Nearest face in the reference's order -x,+x,-y,+y,-z,+z with strict '<' (objects.go:188-217).  With the record
        // (centre c, half extent h) and q = p - c the two face distances of an axis are h + q and h - q: their minimum is
        // h - |q| and it is the + face iff q > 0; across axes the earlier axis keeps a tie.  (Differs from the sequential
        // form only if |q| < ulp(h) on the chosen axis, i.e. never for a point on a face.)
        const float qx = p.x - ob.ax, qy = p.y - ob.ay, qz = p.z - ob.az;
        const float mx = ob.bx - fabsf(qx), my = ob.by - fabsf(qy), mz = ob.bz - fabsf(qz);
        const bool py = my < mx;
        const float mxy = py ? my : mx;
        const bool pz = mz < mxy;
        const float qa = pz ? qz : (py ? qy : qx), da = pz ? d.z : (py ? d.y : d.x);
        const float sgn = qa > 0.0f ? 1.0f : -1.0f;
        front = da * sgn < 0.0f;
        const float v = front ? sgn : -sgn;
        n = f3((py || pz) ? 0.0f : v, (py && !pz) ? v : 0.0f, pz ? v : 0.0f);
        return;
    }
    front = dot3(d, on) < 0.0f;
    n = front ? on : f3(-on.x, -on.y, -on.z);
}

// frontFace (and the hit point) of object `ob` at parameter t WITHOUT building the normal — all the dielectric exit
// search needs (renderer.go:335: it only keeps back-face hits).  Same decisions as surface(): the sphere's outward
// normal is (p - c) / r with r > 0, so its sign test can skip the scaling.
__device__ __forceinline__ bool front_face_only(const DevObj& ob, int type, F3 o, F3 d, float t, F3& p) {
    p = f3(o.x + d.x * t, o.y + d.y * t, o.z + d.z * t);
    if (type == PTB_OBJ_SPHERE) return d.x * (p.x - ob.ax) + d.y * (p.y - ob.ay) + d.z * (p.z - ob.az) < 0.0f;
    if (type == PTB_OBJ_PLANE) return d.y < 0.0f;
    // box: nearest face in the reference's order -x,+x,-y,+y,-z,+z with strict '<' (objects.go:188-217); the face
    // normal is +-e_axis, so d.n = +-d[axis]
    const float qx = p.x - ob.ax, qy = p.y - ob.ay, qz = p.z - ob.az;      // see surface()
    const float mx = ob.bx - fabsf(qx), my = ob.by - fabsf(qy), mz = ob.bz - fabsf(qz);
    const bool py = my < mx;
    const bool pz = mz < (py ? my : mx);
    const float qa = pz ? qz : (py ? qy : qx), da = pz ? d.z : (py ? d.y : d.x);
    return (qa > 0.0f ? da : -da) < 0.0f;
}

__device__ __forceinline__ F3 sky_color(const DevSky& sky, F3 d) {                                // renderer.go:56-92
    if (sky.kind == PTB_SKY_GRADIENT) {
        float len = sqrt_(d.x * d.x + d.y * d.y + d.z * d.z);
        if (len == 0.0f) return f3(sky.horizon[0], sky.horizon[1], sky.horizon[2]);
        float t = (d.y * rcp_(len) + 1.0f) * 0.5f;
        t = t < 0.0f ? 0.0f : t;
        t = t > 1.0f ? 1.0f : t;
        return f3(sky.horizon[0] * (1.0f - t) + sky.zenith[0] * t,
                  sky.horizon[1] * (1.0f - t) + sky.zenith[1] * t,
                  sky.horizon[2] * (1.0f - t) + sky.zenith[2] * t);
    }
    return f3(sky.color[0], sky.color[1], sky.color[2]);
}

// Pixel epilogue of renderer.go:189-221 in binary64 (one value per channel per pixel; cost is nil).
__device__ __forceinline__ uint8_t to_u8(float sum, double inv_spp) {
    double v = sqrt((double)sum * inv_spp) * 255.999;
    if (v < 0.0) v = 0.0; else if (v > 255.999) v = 255.999;
    return (v == v) ? (uint8_t)(int)v : (uint8_t)0;
}

template <bool STATS>
__global__ void __launch_bounds__(PTB_BLOCK_THREADS, PTB_MIN_BLOCKS)
integrate_kernel(const __grid_constant__ KernelArgs ka) {
    const FrameParams& fp = ka.fp;
    const SceneK& c_scene = ka.sc;
    extern __shared__ uint4 s_blob[];
    const int n_obj = c_scene.n_obj;
    {
        const int n_words = n_obj * 2 + c_scene.n_mat * 3;
        for (int i = threadIdx.x; i < n_words; i += blockDim.x) s_blob[i] = fp.scene_blob[i];
        __syncthreads();
    }
    const DevObj* __restrict__ s_obj = reinterpret_cast<const DevObj*>(s_blob);
    const DevMat* __restrict__ s_mat = reinterpret_cast<const DevMat*>(s_blob + 2 * n_obj);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int px = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int py = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    const bool in_frame = px < fp.width && py < fp.height;

    const float xf = (float)px;                               // renderer.go:179
    const float flip_y = fp.h_minus_1 - (float)py;            // renderer.go:174
    const uint32_t pixel_key = fmix(fp.seed_key ^ (uint32_t)(py * fp.width + px));

    unsigned long long st[STATS ? kStatsWords : 1] = {0};

    float acc_x = 0.0f, acc_y = 0.0f, acc_z = 0.0f;
    if (fp.accum_resume && in_frame) {
        const float* a = fp.accum + ((size_t)py * fp.width + px) * 3;
        acc_x = a[0]; acc_y = a[1]; acc_z = a[2];
    }
    int s = fp.s_begin;
    bool alive = in_frame && s < fp.s_end && fp.max_depth > 0;   // max_depth <= 0: black frame (renderer.go:287-289)

    // path state
    Rng rng{0u, 0u};
    F3 o = f3(0, 0, 0), d = f3(0, 0, 1), beta = f3(1, 1, 1), L = f3(0, 0, 0);
    int depth = 0;

    auto start_path = [&]() {                                 // renderer.go:182-185 + camera.go:60-74
        rng.key = fmix(pixel_key + (uint32_t)s * kGolden);
        rng.ctr = 0u;
        const float u = (xf + rng.peek(0)) * fp.inv_w;
        const float v = (flip_y + rng.peek(1)) * fp.inv_h;
        rng.ctr = 2u;
        const DevCamera& cam = c_scene.cam;
        F3 dir = f3(cam.llc[0] + cam.horizontal[0] * u + cam.vertical[0] * v - cam.origin[0],
                    cam.llc[1] + cam.horizontal[1] * u + cam.vertical[1] * v - cam.origin[1],
                    cam.llc[2] + cam.horizontal[2] * u + cam.vertical[2] * v - cam.origin[2]);
        F3 org = f3(cam.origin[0], cam.origin[1], cam.origin[2]);
        if (cam.lens_radius > 0.0f) {
            F3 rd = in_unit_sphere(rng);
            float rx = rd.x * cam.lens_radius, ry = rd.y * cam.lens_radius;
            F3 off = f3(cam.u[0] * rx + cam.v[0] * ry, cam.u[1] * rx + cam.v[1] * ry, cam.u[2] * rx + cam.v[2] * ry);
            org = f3(org.x + off.x, org.y + off.y, org.z + off.z);
            dir = f3(dir.x - off.x, dir.y - off.y, dir.z - off.z);
        }
        sanitize_dir(dir);
        o = org; d = dir;
        beta = f3(1.0f, 1.0f, 1.0f);
        L = f3(0.0f, 0.0f, 0.0f);
        depth = fp.max_depth;
        if (STATS) st[ST_SAMPLES]++;
    };
    if (alive) start_path();

    const int n_box = c_scene.n_box;
    for (;;) {
        if (__ballot_sync(0xffffffffu, alive) == 0u) break;   // warp-uniform exit: every lane has finished its samples
        if (STATS) { st[ST_LANE_TOTAL]++; if (alive) st[ST_LANE_ACTIVE]++; }

        // ---------------- closest hit over the whole world (renderer.go:292-302).
        // Executed by ALL lanes of the warp (finished lanes trace a dummy ray): the loops are warp-uniform, so the
        // object records come straight from the constant bank through the uniform datapath.
        // Device order = boxes first, then spheres/planes, each group in world order.  Reference tie rule kept:
        // a box wins only with t < closest (first box wins ties), a sphere/plane with t <= closest (a later one
        // wins ties, and beats a box at equal t whatever their order) — so grouping boxes first changes nothing.
        const RayK ray = make_ray(o, d);
        float best = FLT_MAX;
        int bid = -1;
#pragma unroll 2
        for (int i = 0; i < n_box; ++i) {
            float t;
            if (hit_box(obj_lo(s_obj, i), obj_hi(s_obj, i), ray, 0.001f, best, t)) { best = t; bid = i; }
        }
        for (int i = n_box; i < n_obj; ++i) {
            const float4 lo = obj_lo(s_obj, i), hi = obj_hi(s_obj, i);
            float t;
            const bool h = (__float_as_int(lo.w) & 3) == PTB_OBJ_SPHERE ? hit_sphere(lo, hi, ray, 0.001f, best, t) : hit_plane(lo, ray, 0.001f, best, t);
            if (h) { best = t; bid = i; }
        }

        if (alive) {
            if (STATS) st[ST_SEGMENTS]++;
            bool done = false;
            if (bid < 0) {                                    // renderer.go:304-306
                F3 sk = sky_color(c_scene.sky, d);
                L.x += beta.x * sk.x; L.y += beta.y * sk.y; L.z += beta.z * sk.z;
                done = true;
                if (STATS) st[ST_END_SKY]++;
            } else {
                const DevObj ob = s_obj[bid];
                const int type = ob.meta & 3;
                if (STATS) st[ST_ACC_SPHERE + type]++;
                F3 p, n; bool front;
                surface(ob, type, o, d, best, p, n, front);
                const DevMat m = s_mat[ob.meta >> 6];

                // the draws this bounce can consume, evaluated together (scatter: <= 2, Russian roulette: 1)
                const float u0 = rng.peek(0), u1 = rng.peek(1), u2 = rng.peek(2);
                uint32_t used = 0u;
                bool repeek = false;

                // |rIn.dir|, unit direction and mirror direction (materials.go:102-112, 177-183, 207-215); lambert ignores them
                const float len = sqrt_(ray.a);
                const float il = rcp_(len);
                const F3 ud = f3(d.x * il, d.y * il, d.z * il);
                const float udn = ud.x * n.x + ud.y * n.y + ud.z * n.z;
                const F3 refl = f3(ud.x - n.x * 2.0f * udn, ud.y - n.y * 2.0f * udn, ud.z - n.z * 2.0f * udn);   // math.go:39-46

                F3 att = f3(m.albedo[0], m.albedo[1], m.albedo[2]), sd = refl, so = p;
                bool ok = true;
                const bool rough_metal = m.type == PTB_MAT_METAL && m.rough > 1e-6f;
                if (m.type == PTB_MAT_EMISSIVE) {             // materials.go:67-72, 202-203; renderer.go:308-312
                    L.x += beta.x * m.emit[0]; L.y += beta.y * m.emit[1]; L.z += beta.z * m.emit[2];
                    ok = false;
                    if (STATS) st[ST_END_EMISSIVE]++;
                } else if (m.type != PTB_MAT_LAMBERT && len == 0.0f) {   // materials.go:103-105, 178-180, 208-210
                    ok = false;
                    if (STATS) st[ST_END_NOSCATTER]++;
                } else if (m.type == PTB_MAT_LAMBERT || rough_metal) {
                    // one cosine-weighted sample, about the normal (lambert, materials.go:76-97) or about the mirror
                    // direction (rough metal, materials.go:114-147): the two materials share this code
                    const F3 cd = cosine_direction(m.type == PTB_MAT_LAMBERT ? n : refl, u0, u1);
                    used = 2u;
                    if (m.type == PTB_MAT_LAMBERT) {
                        sd = cd;
                        if (m.rough > 1e-6f) {                // rejection loop: consumes its draws itself (rare path)
                            rng.ctr += 2u;
                            F3 off = in_unit_sphere(rng);
                            used = 0u; repeek = true;
                            sd.x += off.x * m.rough * 0.1f; sd.y += off.y * m.rough * 0.1f; sd.z += off.z * m.rough * 0.1f;
                            sd = unit3(sd);
                        }
                    } else {
                        const float alpha = m.rough * m.rough;
                        float sx = refl.x * (1.0f - alpha) + cd.x * alpha;
                        float sy = refl.y * (1.0f - alpha) + cd.y * alpha;
                        float sz = refl.z * (1.0f - alpha) + cd.z * alpha;
                        const float l2 = sx * sx + sy * sy + sz * sz;
                        if (l2 < 1e-8f) { sx = refl.x; sy = refl.y; sz = refl.z; }
                        else { const float i2 = rcp_(sqrt_(l2)); sx *= i2; sy *= i2; sz *= i2; }
                        if (sx * n.x + sy * n.y + sz * n.z <= 0.0f) { sx = refl.x; sy = refl.y; sz = refl.z; }
                        sd = f3(sx, sy, sz);
                    }
                } else if (m.type == PTB_MAT_DIELECTRIC) {    // materials.go:162-200
                    att = f3(1.0f, 1.0f, 1.0f);
                    const float ratio = front ? rcp_(m.ior) : m.ior;
                    const float cos_t = fminf(-udn, 1.0f);
                    const float sin_t = sqrt_(1.0f - cos_t * cos_t);
                    const bool cannot = ratio * sin_t > 1.0f;
                    float r0 = (1.0f - ratio) * rcp_(1.0f + ratio);
                    r0 = r0 * r0;
                    const float om = 1.0f - cos_t;
                    const float om2 = om * om;
                    const float refl_prob = r0 + (1.0f - r0) * (om2 * om2 * om);   // Schlick, materials.go:226-231
                    bool reflect = cannot;
                    if (!cannot) { reflect = refl_prob > u0; used = 1u; }           // `||` short-circuit: no draw when cannot
                    if (!reflect) {                           // refractVec, math.go:48-64
                        const float c2 = fminf(-ud.x * n.x - ud.y * n.y - ud.z * n.z, 1.0f);
                        float qx = (ud.x + n.x * c2) * ratio, qy = (ud.y + n.y * c2) * ratio, qz = (ud.z + n.z * c2) * ratio;
                        const float par = -sqrt_(fabsf(1.0f - (qx * qx + qy * qy + qz * qz)));
                        sd = f3(qx + n.x * par, qy + n.y * par, qz + n.z * par);
                    }
                    sanitize_dir(sd);
                    if (front) {                              // exit search, renderer.go:316-371
                        if (STATS) st[ST_EXIT_SCANS]++;
                        const RayK er = make_ray(p, sd);
                        float exit_t = FLT_MAX;
                        bool hit_exit = false;
                        F3 ep = p;
                        const int n_diel = c_scene.n_diel;
                        for (int k = 0; k < n_diel; ++k) {    // only dielectric objects can be accepted (:335)
                            const int ei = __ldg(c_scene.diel_idx + k);
                            const DevObj& eo = s_obj[ei];
                            const int et = eo.meta & 3;
                            float t;
                            if (!hit_any(obj_lo(s_obj, ei), obj_hi(s_obj, ei), et, er, 0.0001f, exit_t, t)) continue;
                            F3 q;
                            const bool qf = front_face_only(eo, et, p, sd, t, q);
                            if (!qf && t < exit_t) {
                                float ex = q.x - p.x, ey = q.y - p.y, ez = q.z - p.z;
                                float d2 = ex * ex + ey * ey + ez * ez;
                                if (d2 > 1e-8f && d2 < 1000.0f) { hit_exit = true; exit_t = t; ep = q; }
                            }
                        }
                        if (hit_exit) {                       // renderer.go:352-369
                            float ex = ep.x - p.x, ey = ep.y - p.y, ez = ep.z - p.z;
                            float dist = sqrt_(ex * ex + ey * ey + ez * ez);
                            if (m.absorption[0] > 0.0f || m.absorption[1] > 0.0f || m.absorption[2] > 0.0f) {
                                att = f3(exp_(-m.absorption[0] * dist), exp_(-m.absorption[1] * dist), exp_(-m.absorption[2] * dist));
                            }
                            so = ep;
                        }
                    }
                }
                // (mirror and smooth metal: sd = refl, att = albedo — the defaults; materials.go:148-158, 205-221)

                if (!ok) {
                    done = true;
                } else {
                    if (STATS) st[ST_SCATTERS]++;
                    if (depth <= 3) {                         // Russian roulette, renderer.go:374-393
                        const float mx = fmaxf(att.x, fmaxf(att.y, att.z));
                        if (mx < 1e-6f) {
                            done = true;
                        } else {
                            const float pr = fminf(mx, 0.95f);
                            float ur = used == 0u ? u0 : (used == 1u ? u1 : u2);
                            if (repeek) ur = rng.peek(0);
                            used += 1u;
                            if (ur > pr) done = true;
                            else { const float ip = rcp_(pr); att.x *= ip; att.y *= ip; att.z *= ip; }
                        }
                        if (STATS && done) st[ST_END_RR]++;
                    }
                    rng.ctr += used;
                    if (!done) {                              // renderer.go:398-403
                        beta.x *= att.x; beta.y *= att.y; beta.z *= att.z;
                        sanitize_dir(sd);
                        o = so; d = sd;
                        if (--depth <= 0) {                   // renderer.go:287-289
                            done = true;
                            if (STATS) st[ST_END_DEPTH]++;
                        }
                    }
                }
            }

            if (done) {                                       // renderer.go:186: col = col.add(...)
                acc_x += L.x; acc_y += L.y; acc_z += L.z;
                if (++s < fp.s_end) start_path(); else alive = false;
            }
        }
    }

    if (in_frame) {
        const size_t pix = (size_t)py * fp.width + px;
        if (fp.accum) {
            float* a = fp.accum + pix * 3;
            a[0] = acc_x; a[1] = acc_y; a[2] = acc_z;
        }
        if (fp.rgba) {
            const double inv_spp = 1.0 / (double)fp.spp_total;
            uchar4 c = make_uchar4(to_u8(acc_x, inv_spp), to_u8(acc_y, inv_spp), to_u8(acc_z, inv_spp), 255);
            reinterpret_cast<uchar4*>(fp.rgba)[pix] = c;
        }
    }
    if (STATS) {
        for (int k = 0; k < kStatsWords; ++k) {
            unsigned long long v = st[k];
            for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
            if (lane == 0 && v) atomicAdd(fp.stats + k, v);
        }
    }
}

}  // namespace ptb
#include "bvh.cuh"
#include "wavefront.cuh"
#include "mesh_pipeline.cuh"
namespace ptb {

__global__ void finalize_kernel(const float* __restrict__ accum, int n_pix, double inv_spp, uchar4* __restrict__ rgba) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    const float* a = accum + (size_t)i * 3;
    rgba[i] = make_uchar4(to_u8(a[0], inv_spp), to_u8(a[1], inv_spp), to_u8(a[2], inv_spp), 255);
}

// max_depth <= 0: rayColorOpt returns black at once (renderer.go:287-289) — every pixel is (0, 0, 0, 255) and the sums
// do not change.
__global__ void clear_frame_kernel(float* __restrict__ accum, int accum_resume, uchar4* __restrict__ rgba, int n_pix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    if (accum && !accum_resume) { accum[3 * (size_t)i] = 0.0f; accum[3 * (size_t)i + 1] = 0.0f; accum[3 * (size_t)i + 2] = 0.0f; }
    if (rgba) rgba[i] = make_uchar4(0, 0, 0, 255);
}

// Split frames (FrameParams::split_k > 1): add the k partial-sum planes of every pixel in plane order, then accumulation
// buffer and / or pixel epilogue.
__global__ void finalize_planes_kernel(const float* __restrict__ planes, int k, int n_pix, float* __restrict__ accum, int accum_resume,
                                       double inv_spp, uchar4* __restrict__ rgba) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    float r = 0.0f, g = 0.0f, b = 0.0f;
    if (accum && accum_resume) { r = accum[3 * (size_t)i]; g = accum[3 * (size_t)i + 1]; b = accum[3 * (size_t)i + 2]; }
    for (int p = 0; p < k; ++p) {
        const float* s = planes + ((size_t)p * n_pix + i) * 3;
        r += s[0]; g += s[1]; b += s[2];
    }
    if (accum) { accum[3 * (size_t)i] = r; accum[3 * (size_t)i + 1] = g; accum[3 * (size_t)i + 2] = b; }
    if (rgba) rgba[i] = make_uchar4(to_u8(r, inv_spp), to_u8(g, inv_spp), to_u8(b, inv_spp), 255);
}

// Multi-GPU reduce fused with the pixel epilogue: bufs[k] is device k's fp32 sum buffer; bufs[1..] are PEER pointers, read
// over NVLink with 16-byte loads.  Sum order is device 0, 1, 2, ... (deterministic).
__global__ void finalize_peers_kernel(const float* const* __restrict__ bufs, int n_bufs, int n_pix, double inv_spp, uchar4* __restrict__ rgba) {
    // 4 pixels = 12 floats = 3 float4 per thread
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int p0 = g * 4;
    if (p0 >= n_pix) return;
    if (p0 + 4 <= n_pix) {
        float4 a = make_float4(0, 0, 0, 0), b = a, c = a;
        for (int k = 0; k < n_bufs; ++k) {
            const float4* src = reinterpret_cast<const float4*>(bufs[k]) + (size_t)g * 3;
            const float4 x = src[0], y = src[1], z = src[2];
            a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
            b.x += y.x; b.y += y.y; b.z += y.z; b.w += y.w;
            c.x += z.x; c.y += z.y; c.z += z.z; c.w += z.w;
        }
        rgba[p0 + 0] = make_uchar4(to_u8(a.x, inv_spp), to_u8(a.y, inv_spp), to_u8(a.z, inv_spp), 255);
        rgba[p0 + 1] = make_uchar4(to_u8(a.w, inv_spp), to_u8(b.x, inv_spp), to_u8(b.y, inv_spp), 255);
        rgba[p0 + 2] = make_uchar4(to_u8(b.z, inv_spp), to_u8(b.w, inv_spp), to_u8(c.x, inv_spp), 255);
        rgba[p0 + 3] = make_uchar4(to_u8(c.y, inv_spp), to_u8(c.z, inv_spp), to_u8(c.w, inv_spp), 255);
    } else {
        for (int p = p0; p < n_pix; ++p) {
            float r = 0, gg = 0, bb = 0;
            for (int k = 0; k < n_bufs; ++k) { const float* s = bufs[k] + (size_t)p * 3; r += s[0]; gg += s[1]; bb += s[2]; }
            rgba[p] = make_uchar4(to_u8(r, inv_spp), to_u8(gg, inv_spp), to_u8(bb, inv_spp), 255);
        }
    }
}

// The same reduce + epilogue for ONE SLICE of the frame, the multi-process arrangement (one rank per GPU): rank r sums pixels
// [p_begin, p_end) of all ranks' buffers — its own from HBM, the others through NVLink peer loads (CUDA IPC mappings) — and
// stores the finalised RGBA8 pixels straight into rank 0's image, again over NVLink: reduce-scatter, epilogue and gather in
// one kernel, every link of the switch carrying 1/N of the traffic instead of everything converging on rank 0.
// The cross-rank ordering is part of the kernel as well (no collective library call per frame): PeerSync::flags_of[k] is rank
// k's flag block (2 x world words, mapped on every rank); a rank announces "my sums of frame `seq` are complete" by storing seq
// into word [rank] of every rank's block, waits for all world announcements in its OWN block, reduces, and — its last CTA —
// stores seq into word [world + rank] of every block: "my slice of the image is written and I no longer read anybody's sums".
// p_begin is a multiple of 4 (3 x 16-byte loads per 4 pixels and buffer).  Sum order: rank 0, 1, 2, ... on every rank.
__device__ __forceinline__ bool peer_wait_all(const volatile unsigned* flags, int world, unsigned seq) {
    const long long t0 = clock64();
    for (int k = 0; k < world; ++k)
        while ((int)(flags[k] - seq) < 0)                                  // (sequence numbers only grow)
            if (clock64() - t0 > 8000000000ll) return false;                // ~4 s: a rank died; do not hang the device
    return true;
}

__global__ void reduce_finalize_slice_kernel(const float* const* __restrict__ bufs, size_t buf_off, int n_bufs, long long p_begin, long long p_end, double inv_spp,
                                             uchar4* __restrict__ rgba_root, PeerSync sync) {
    if (sync.world > 1) {
        if (blockIdx.x == 0 && (int)threadIdx.x < sync.world) {
            __threadfence_system();                                         // (the integrator's sums: written by the previous kernel of this stream)
            *(volatile unsigned*)(sync.flags_of[threadIdx.x] + sync.rank) = sync.seq;
        }
        __shared__ int ok;
        if (threadIdx.x == 0) ok = peer_wait_all(sync.flags_of[sync.rank], sync.world, sync.seq) ? 1 : 0;
        __syncthreads();
        if (!ok) { if (threadIdx.x == 0) *sync.error = 1u; return; }
        __threadfence_system();
    }
    const long long p0 = p_begin + 4ll * ((long long)blockIdx.x * blockDim.x + threadIdx.x);
    if (p0 < p_end) {
        if (p0 + 4 <= p_end) {
            float4 a = make_float4(0, 0, 0, 0), b = a, c = a;
            // (plain loads: this kernel has not touched the peers' sums before their "complete" flags arrived, and L1 does not
            // outlive a kernel, so no stale line can be read; the loop is unrolled to keep many NVLink reads in flight per thread)
#pragma unroll 4
            for (int k = 0; k < n_bufs; ++k) {
                const float4* src = reinterpret_cast<const float4*>(bufs[k] + buf_off + 3 * p0);
                const float4 x = src[0], y = src[1], z = src[2];
                a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
                b.x += y.x; b.y += y.y; b.z += y.z; b.w += y.w;
                c.x += z.x; c.y += z.y; c.z += z.z; c.w += z.w;
            }
            uint4 o;   // 4 pixels = one 16-byte store
            o.x = (uint32_t)to_u8(a.x, inv_spp) | (uint32_t)to_u8(a.y, inv_spp) << 8 | (uint32_t)to_u8(a.z, inv_spp) << 16 | 0xFF000000u;
            o.y = (uint32_t)to_u8(a.w, inv_spp) | (uint32_t)to_u8(b.x, inv_spp) << 8 | (uint32_t)to_u8(b.y, inv_spp) << 16 | 0xFF000000u;
            o.z = (uint32_t)to_u8(b.z, inv_spp) | (uint32_t)to_u8(b.w, inv_spp) << 8 | (uint32_t)to_u8(c.x, inv_spp) << 16 | 0xFF000000u;
            o.w = (uint32_t)to_u8(c.y, inv_spp) | (uint32_t)to_u8(c.z, inv_spp) << 8 | (uint32_t)to_u8(c.w, inv_spp) << 16 | 0xFF000000u;
            *reinterpret_cast<uint4*>(rgba_root + p0) = o;
        } else {
            for (long long p = p0; p < p_end; ++p) {
                float r = 0, g = 0, bb = 0;
                for (int k = 0; k < n_bufs; ++k) { const float* s = bufs[k] + buf_off + 3 * p; r += s[0]; g += s[1]; bb += s[2]; }
                rgba_root[p] = make_uchar4(to_u8(r, inv_spp), to_u8(g, inv_spp), to_u8(bb, inv_spp), 255);
            }
        }
    }
    if (sync.world > 1) {                       // last CTA of the grid: this rank's slice is in rank 0's image
        __threadfence_system();
        __syncthreads();
        __shared__ unsigned last;
        if (threadIdx.x == 0) last = atomicAdd(sync.block_counter, 1u) == gridDim.x - 1 ? 1u : 0u;
        __syncthreads();
        if (last) {
            if (threadIdx.x == 0) *sync.block_counter = 0u;
            if ((int)threadIdx.x < sync.world) *(volatile unsigned*)(sync.flags_of[threadIdx.x] + sync.world + sync.rank) = sync.seq;
        }
    }
}

// Stream-ordered wait for "every rank has finished frame `seq`".  The sums are double-buffered by frame parity, so a rank only
// has to know that frame seq - 1 is finished everywhere before it renders frame seq + 1 into the same buffer — the ranks may
// drift apart by a frame instead of marching in lock-step — while rank 0 waits for frame seq itself before it reads the image.
__global__ void peer_wait_done_kernel(PeerSync sync, unsigned seq) {
    if (threadIdx.x == 0 && !peer_wait_all(sync.flags_of[sync.rank] + sync.world, sync.world, seq)) *sync.error = 1u;
}

// FP32 FMA throughput probe: 8 independent chains per thread, 2 flop per FMA.
__global__ void fma_peak_kernel(float* out, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f, c = 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ---------------------------------------------------------------- launchers
int launch_integrator(const KernelArgs& ka, bool stats, void* stream) {
    const FrameParams& fp = ka.fp;
    dim3 grid((fp.width + 15) / 16, (fp.height + 7) / 8);
    size_t smem = (size_t)(ka.sc.n_obj * 2 + ka.sc.n_mat * 3) * sizeof(uint4);     // <= kSmallBlobBytes (checked by the caller)
    if (stats) integrate_kernel<true><<<grid, PTB_BLOCK_THREADS, smem, (cudaStream_t)stream>>>(ka);
    else integrate_kernel<false><<<grid, PTB_BLOCK_THREADS, smem, (cudaStream_t)stream>>>(ka);
    return (int)cudaGetLastError();
}

int launch_clear_frame(float* accum, int accum_resume, uint8_t* rgba, int n_pix, void* stream) {
    clear_frame_kernel<<<(n_pix + 255) / 256, 256, 0, (cudaStream_t)stream>>>(accum, accum_resume, reinterpret_cast<uchar4*>(rgba), n_pix);
    return (int)cudaGetLastError();
}

constexpr int kMeshBlocksPerSm = PTB_WF_MIN_BLOCKS;
size_t wf_trav_scratch_bytes(int sm_count) { return (size_t)sm_count * kMeshBlocksPerSm * WF_SLOTS * kTravStride * sizeof(int); }

// The opt-in to more than 48 KB of dynamic shared memory is a per-DEVICE attribute of the kernel and the occupancy depends on
// the shared-memory size of the scene at hand, so both are remembered per context (= per device) and per size: a small
// scene rendered first, a large one later, or several devices in one process all get the right attribute and grid.
template <bool STATS, bool MESH, bool BIG, bool PK>
static int launch_wf_variant(const KernelArgs& ka, size_t smem, int sm_count, LaunchCache::Entry& lc, cudaStream_t stream) {
    const FrameParams& fp = ka.fp;
    if (smem > 48 * 1024 && smem > lc.smem_optin) {
        cudaError_t e = cudaFuncSetAttribute(integrate_wf_kernel<STATS, MESH, BIG, PK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        lc.smem_optin = smem;
    }
    if (lc.smem_occ != smem) {
        int nb = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, integrate_wf_kernel<STATS, MESH, BIG, PK>, WF_THREADS, smem);
        if (e != cudaSuccess) return (int)e;
        lc.blocks_per_sm = nb > 0 ? nb : 1;
        lc.smem_occ = smem;
    }
    const long long n_pix = (long long)fp.width * fp.rows;
    long long grid = (long long)sm_count * (MESH && lc.blocks_per_sm > kMeshBlocksPerSm ? kMeshBlocksPerSm : lc.blocks_per_sm);   // trav_scratch is sized for kMeshBlocksPerSm
    const long long need = (n_pix * (fp.split_k > 1 ? fp.split_k : 1) + WF_SLOTS - 1) / WF_SLOTS;
    if (grid > need) grid = need;
    integrate_wf_kernel<STATS, MESH, BIG, PK><<<(unsigned)grid, WF_THREADS, smem, stream>>>(ka);
    return (int)cudaGetLastError();
}

int launch_integrator_wf(const KernelArgs& ka, bool stats, bool big, bool packed, int sm_count, LaunchCache* cache, void* stream) {
    // BIG worlds read the object / material records in place (global memory): no shared-memory copy
    const size_t smem = ((sizeof(WfState) + 15) / 16 + (big ? 0 : (size_t)(ka.sc.n_obj * 2 + ka.sc.n_mat * 3))) * sizeof(uint4);
    cudaStream_t st = (cudaStream_t)stream;
    const bool mesh = ka.fp.bvh_nodes != nullptr;      // the mesh-free instantiation carries no traversal code (it costs ~10 %)
    LaunchCache::Entry& lc = cache->wf[(packed ? 8 : 0) | (stats ? 4 : 0) | (mesh ? 2 : 0) | (big ? 1 : 0)];
#define PTB_WF_CASE(S, M, B) if (stats == S && mesh == M && big == B) \
        return packed ? launch_wf_variant<S, M, B, true>(ka, smem, sm_count, lc, st) : launch_wf_variant<S, M, B, false>(ka, smem, sm_count, lc, st);
    PTB_WF_CASE(false, false, false) PTB_WF_CASE(false, false, true) PTB_WF_CASE(false, true, false) PTB_WF_CASE(false, true, true)
    PTB_WF_CASE(true, false, false) PTB_WF_CASE(true, false, true) PTB_WF_CASE(true, true, false) PTB_WF_CASE(true, true, true)
#undef PTB_WF_CASE
    return (int)cudaErrorInvalidValue;
}

// ---- mesh pipeline (mesh_pipeline.cuh)
size_t mesh_pool_bytes(int n_slots) {
    return (size_t)n_slots * (4 * sizeof(float4) + sizeof(float) + sizeof(int) + sizeof(int) + sizeof(unsigned short)) + kMpCtlWords * sizeof(unsigned int) + 1024;
}
void mesh_pool_bind(MeshPool& pool, void* d_mem, int n_slots) {
    char* p = (char*)d_mem;
    auto take = [&](size_t bytes) { void* r = p; p += (bytes + 255) / 256 * 256; return r; };
    pool.n_slots = n_slots;
    pool.O = (float4*)take((size_t)n_slots * sizeof(float4)); pool.D = (float4*)take((size_t)n_slots * sizeof(float4));
    pool.B = (float4*)take((size_t)n_slots * sizeof(float4)); pool.A = (float4*)take((size_t)n_slots * sizeof(float4));
    pool.best = (float*)take((size_t)n_slots * sizeof(float)); pool.bid = (int*)take((size_t)n_slots * sizeof(int));
    pool.queue = (int*)take((size_t)n_slots * sizeof(int)); pool.dep = (unsigned short*)take((size_t)n_slots * sizeof(unsigned short));
    pool.ctl = (unsigned int*)take(kMpCtlWords * sizeof(unsigned int));
}
int mesh_pool_slots(int sm_count, long long n_items) {
    long long want = (long long)sm_count * PTB_MP_CTAS_PER_SM * WF_SLOTS;
    if (const char* e = std::getenv("PTB_MP_CTAS_PER_SM")) { int k = std::atoi(e); if (k >= 1 && k <= 64) want = (long long)sm_count * k * WF_SLOTS; }
    const long long need = (n_items + WF_SLOTS - 1) / WF_SLOTS * WF_SLOTS;
    return (int)(want < need ? want : need);
}

template <bool STATS, bool BIG>
static int mp_prepare(size_t smem, LaunchCache::Entry& lc) {       // (outside the stream capture)
    if (smem > 48 * 1024 && smem > lc.smem_optin) {
        cudaError_t e = cudaFuncSetAttribute(mp_shade_scan_kernel<STATS, BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        lc.smem_optin = smem;
    }
    return 0;
}
template <bool STATS, bool BIG>
static int launch_mp_iteration(const KernelArgs& ka, const MeshPool& pool, size_t smem, int traverse_blocks, cudaStream_t stream) {
    mp_shade_scan_kernel<STATS, BIG><<<pool.n_slots / WF_SLOTS, WF_THREADS, smem, stream>>>(ka, pool);
    mp_traverse_kernel<STATS><<<traverse_blocks, 256, 0, stream>>>(ka.fp.bvh_nodes, ka.fp.bvh_tris, pool, ka.fp.stats);
    return (int)cudaGetLastError();
}

int launch_mesh_pipeline(const KernelArgs& ka, bool stats, bool big, int sm_count, LaunchCache* cache, MeshPipe& mp, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const MeshPool& pool = mp.pool;
    const size_t smem = ((sizeof(WfState) + 15) / 16 + (big ? 0 : (size_t)(ka.sc.n_obj * 2 + ka.sc.n_mat * 3))) * sizeof(uint4);
    if (!mp.traverse_blocks) {
        int nb = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, mp_traverse_kernel<false>, 256, 0);
        if (e != cudaSuccess) return (int)e;
        mp.traverse_blocks = sm_count * (nb > 0 ? nb : 1);
    }
    LaunchCache::Entry& lc = cache->mp[(stats ? 2 : 0) | (big ? 1 : 0)];
    {
        const int pe = stats ? (big ? mp_prepare<true, true>(smem, lc) : mp_prepare<true, false>(smem, lc)) : (big ? mp_prepare<false, true>(smem, lc) : mp_prepare<false, false>(smem, lc));
        if (pe) return pe;
    }
    auto iteration = [&]() -> int {
        if (stats) return big ? launch_mp_iteration<true, true>(ka, pool, smem, mp.traverse_blocks, stream) : launch_mp_iteration<true, false>(ka, pool, smem, mp.traverse_blocks, stream);
        return big ? launch_mp_iteration<false, true>(ka, pool, smem, mp.traverse_blocks, stream) : launch_mp_iteration<false, false>(ka, pool, smem, mp.traverse_blocks, stream);
    };
    // One period = kCheck iterations + a copy of the newest "a path is alive" flag to pinned memory, captured ONCE as a CUDA
    // graph and replayed: the kernels take the iteration number from device memory, so every period is the same graph except
    // for the flag word copied, which is why the copy and the event stay outside the graph.
    constexpr int kCheck = 8;
    mp_init_kernel<<<(pool.n_slots + 255) / 256, 256, 0, stream>>>(pool);
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaError_t ce = cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal);
    if (ce != cudaSuccess) return (int)ce;
    int e = 0;
    for (int k = 0; k < kCheck && !e; ++k) e = iteration();
    ce = cudaStreamEndCapture(stream, &graph);
    if (e) { if (graph) cudaGraphDestroy(graph); return e; }
    if (ce == cudaSuccess) ce = cudaGraphInstantiate(&exec, graph, 0);
    if (ce != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return (int)ce; }
    int pending = -1;                    // event slot of the flag copy not yet looked at
    for (int period = 0; ce == cudaSuccess; ++period) {
        ce = cudaGraphLaunch(exec, stream);
        if (ce != cudaSuccess) break;
        if (pending >= 0) {              // the host runs one period ahead of the flag it reads: no bubble on the device
            ce = cudaEventSynchronize((cudaEvent_t)mp.events[pending]);
            if (ce != cudaSuccess || mp.h_flags[pending] == 0u) break;     // no path was alive at the end of the previous period
        }
        const int slot = period & 1, last_iter = period * kCheck + kCheck - 1;
        ce = cudaMemcpyAsync(&mp.h_flags[slot], pool.ctl + 8 + (last_iter & 15), sizeof(unsigned int), cudaMemcpyDeviceToHost, stream);
        if (ce == cudaSuccess) ce = cudaEventRecord((cudaEvent_t)mp.events[slot], stream);
        pending = slot;
    }
    cudaError_t se = cudaStreamSynchronize(stream);
    cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
    return (int)(ce != cudaSuccess ? ce : se);
}

int launch_finalize(const float* accum, int width, int height, int spp_total, uint8_t* rgba, void* stream) {
    int n = width * height;
    finalize_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(accum, n, 1.0 / (double)spp_total,
                                                                      reinterpret_cast<uchar4*>(rgba));
    return (int)cudaGetLastError();
}

int launch_finalize_planes(const float* planes, int split_k, int width, int height, int spp_total, float* accum, int accum_resume, uint8_t* rgba, void* stream) {
    const int n = width * height, threads = 256;
    finalize_planes_kernel<<<(n + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(planes, split_k, n, accum, accum_resume, 1.0 / (double)spp_total,
                                                                                            reinterpret_cast<uchar4*>(rgba));
    return (int)cudaGetLastError();
}

// Whole pixels as work items leave path slots idle on small frames, and on mid-size frames the kernel ends on a long tail: the
// slots finish their last items at different times while every CTA still pays for full-width scans.  The tail lasts about one
// item, so it costs ~1 / (items per slot); pixels of very different cost (sky next to glass) make it worse.  Measured on a B200
// (profiles/r02s_split.log): C2 (1920x1080, 64 spp, 6.8 pixels per slot) 22.6 ms as whole pixels, 20.7 ms in 8 planes; C3 (4K, 27
// pixels per slot) gains nothing from planes and C5 (8K, 109 per slot) loses 1 % to the plane traffic.  Policy: about 40 work
// items per resident slot (rounded to the nearest plane count), items of at least one sample, at most 64 planes — so the plane
// buffer never exceeds 40 x slots x 12 bytes = 146 MB.
int wf_split_factor(int sm_count, long long n_pix, int n_samples) {
    const long long slots = (long long)sm_count * PTB_WF_MIN_BLOCKS * WF_SLOTS;
    if (n_samples < 2 || n_pix < 1) return 1;
    long long k = (40 * slots + n_pix / 2) / n_pix;
    if (k > n_samples) k = n_samples;
    if (k > 64) k = 64;
    if (k < 1) k = 1;
    return (int)k;
}

int launch_finalize_peers(const float* const* d_bufs_on_dev0, int n_bufs, int width, int height, int spp_total, uint8_t* rgba, void* stream) {
    const int n = width * height, threads = 256, groups = (n + 3) / 4;
    finalize_peers_kernel<<<(groups + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(d_bufs_on_dev0, n_bufs, n, 1.0 / (double)spp_total,
                                                                                              reinterpret_cast<uchar4*>(rgba));
    return (int)cudaGetLastError();
}

int launch_reduce_finalize_slice(const float* const* d_bufs, size_t buf_off, int n_bufs, long long p_begin, long long p_end, int spp_total, uint8_t* rgba_root,
                                 const PeerSync& sync, void* stream) {
    // (an empty slice still takes part in the flag protocol: one CTA)
    const long long groups = p_end > p_begin ? (p_end - p_begin + 3) / 4 : 1;
    const int threads = 256;
    reduce_finalize_slice_kernel<<<(unsigned)((groups + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(d_bufs, buf_off, n_bufs, p_begin, p_end > p_begin ? p_end : p_begin,
                                                                                                                   1.0 / (double)spp_total, reinterpret_cast<uchar4*>(rgba_root), sync);
    if (sync.world > 1) {
        const unsigned need = sync.rank == 0 ? sync.seq : sync.seq - 1u;        // (seq >= 1; "frame 0" is finished by definition)
        peer_wait_done_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sync, need);
    }
    return (int)cudaGetLastError();
}

int launch_fma_peak(float* d_out, int blocks, int threads, int iters, void* stream) {
    fma_peak_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(d_out, iters);
    return (int)cudaGetLastError();
}

}  // namespace ptb
