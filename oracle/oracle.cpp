/*
 * oracle.cpp — CPU oracle: a C++ restatement of the Go CPU path tracer's hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  PARITY UNPINNED: the reference has no
 * tests/golden vectors and cannot be executed here (no Go toolchain), so this file
 * is checked only against hand-derived known-answer vectors.
 *
 * Every function cites the reference lines (relative to /root/reference) it follows.
 * Arithmetic is written operation-for-operation in the reference's evaluation order
 * and must be compiled with -ffp-contract=off (Go's gc compiler does not fuse
 * multiply-add on amd64 at the default GOAMD64=v1).
 *
 * Everything is templated on Real: double = Go-faithful binary64; float = the same
 * algorithm in binary32, which is what the CUDA integrator computes in.
 *
 * One deliberate difference from the reference: random.go:14-16 seeds math/rand from
 * the wall clock per worker, so reference output is not reproducible and only the
 * DISTRIBUTION (uniform on [0,1)) is contractual.  The oracle draws from the same
 * counter hash as the CUDA path (spec in DESIGN.md "RNG"), keyed by
 * (seed, pixel, sample) with a per-path draw counter, so both sides can be compared
 * path for path.
 */
#include "oracle.h"

#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>
#include <algorithm>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------- RNG (shared spec)
inline uint32_t fmix(uint32_t x) {
    x ^= x >> 16; x *= 0x21f0aaadu; x ^= x >> 15; x *= 0x735a2d97u; x ^= x >> 15;
    return x;
}
inline uint32_t path_key(uint32_t seed, uint32_t pixel, uint32_t sample) {
    uint32_t k = fmix(seed ^ 0x9E3779B9u);
    k = fmix(k ^ pixel);
    k = fmix(k + sample * 0x9E3779B9u);
    return k;
}
struct Rng {                         // stands in for randSource (random.go:10-34)
    uint32_t key, ctr;
    Rng(uint32_t seed, uint32_t pixel, uint32_t sample) : key(path_key(seed, pixel, sample)), ctr(0) {}
    inline uint32_t next_u32() { uint32_t x = fmix(key + ctr * 0x9E3779B9u); ++ctr; return x; }
    template <class R> inline R uniform() {      // Float64(): uniform on [0,1) (random.go:27-34)
        return (R)(next_u32() >> 8) * (R)(1.0 / 16777216.0);
    }
};

// ---------------------------------------------------------------- vec3 (math.go:5-37)
template <class R> struct V3 { R x, y, z; };
template <class R> inline V3<R> mk(R x, R y, R z) { return V3<R>{x, y, z}; }
template <class R> inline V3<R> add(V3<R> a, V3<R> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }   // math.go:11
template <class R> inline V3<R> sub(V3<R> a, V3<R> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }   // math.go:12
template <class R> inline V3<R> mul(V3<R> a, R t) { return {a.x * t, a.y * t, a.z * t}; }             // math.go:13
template <class R> inline V3<R> divv(V3<R> a, R t) { R inv = (R)1.0 / t; return {a.x * inv, a.y * inv, a.z * inv}; } // math.go:14-17
template <class R> inline R dot(V3<R> a, V3<R> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }       // math.go:19
template <class R> inline V3<R> cross(V3<R> a, V3<R> b) {                                              // math.go:21-27
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <class R> inline R length(V3<R> a) { return std::sqrt(dot(a, a)); }                          // math.go:29
template <class R> inline V3<R> unit(V3<R> a) {                                                       // math.go:31-37
    R l = length(a);
    if (l == 0) return a;
    return divv(a, l);
}
template <class R> struct Ray { V3<R> orig, dir; };                                                    // math.go:133-140

template <class R> inline V3<R> reflectVec(V3<R> v, V3<R> n) {                                        // math.go:39-46
    R d = dot(v, n);
    return {v.x - n.x * 2 * d, v.y - n.y * 2 * d, v.z - n.z * 2 * d};
}
template <class R> inline V3<R> refractVec(V3<R> uv, V3<R> n, R eta) {                                // math.go:48-64
    R cosTheta = std::fmin(-uv.x * n.x - uv.y * n.y - uv.z * n.z, (R)1.0);
    R px = uv.x + n.x * cosTheta, py = uv.y + n.y * cosTheta, pz = uv.z + n.z * cosTheta;
    px *= eta; py *= eta; pz *= eta;
    R perpLenSq = px * px + py * py + pz * pz;
    R par = -std::sqrt(std::fabs((R)1.0 - perpLenSq));
    return {px + n.x * par, py + n.y * par, pz + n.z * par};
}
template <class R> inline V3<R> randomInUnitSphere(Rng& rng) {                                        // math.go:66-85
    for (;;) {
        R x = rng.uniform<R>() * 2 - 1;
        R y = rng.uniform<R>() * 2 - 1;
        R z = rng.uniform<R>() * 2 - 1;
        R lenSq = x * x + y * y + z * z;
        if (lenSq >= (R)1.0) continue;
        return {x, y, z};
    }
}
template <class R> inline V3<R> randomCosineDirection(V3<R> normal, Rng& rng) {                       // math.go:94-131
    R r1 = rng.uniform<R>();
    R r2 = rng.uniform<R>();
    R phi = (R)(2.0 * M_PI) * r1;
    R cosTheta = std::sqrt(r2);
    R sinTheta = std::sqrt((R)1.0 - r2);
    V3<R> u = (std::fabs(normal.x) > (R)0.9) ? mk<R>(0, 1, 0) : mk<R>(1, 0, 0);
    V3<R> w = normal;
    V3<R> vVec = unit(cross(w, u));
    V3<R> uVec = cross(vVec, w);
    V3<R> l = {sinTheta * std::cos(phi), sinTheta * std::sin(phi), cosTheta};
    return {l.x * uVec.x + l.y * vVec.x + l.z * w.x,
            l.x * uVec.y + l.y * vVec.y + l.z * w.y,
            l.x * uVec.z + l.y * vVec.z + l.z * w.z};
}

// ---------------------------------------------------------------- materials (materials.go:9-66)
enum { matLambert = 0, matMetal, matDielectric, matEmissive, matMirror };
template <class R> struct Material {
    int typ = matLambert;
    V3<R> albedo{0, 0, 0};
    R rough = 0, ior = 0;
    V3<R> emit{0, 0, 0};
    V3<R> absorption{0, 0, 0};
};
inline double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }      // materials.go:57-65

Material<double> convertMaterial(const orc_raw_material& m) {                                          // materials.go:28-55
    Material<double> out;
    V3<double> al = {m.albedo[0], m.albedo[1], m.albedo[2]};
    V3<double> em = {m.emit[0] * m.power, m.emit[1] * m.power, m.emit[2] * m.power};
    V3<double> ab = {m.absorption[0], m.absorption[1], m.absorption[2]};
    std::string t = m.type ? m.type : "";
    if (t == "metal") {
        double rough = m.rough;
        if (m.smoothness > 0) rough = 1.0 - clampd(m.smoothness, 0, 1);
        out.typ = matMetal; out.albedo = al; out.rough = clampd(rough, 0, 1);
    } else if (t == "dielectric") {
        double ior = m.ior;
        if (ior == 0) ior = 1.5;
        out.typ = matDielectric; out.albedo = al; out.ior = ior; out.absorption = ab;
    } else if (t == "emissive") {
        out.typ = matEmissive; out.emit = em;
    } else if (t == "mirror") {
        out.typ = matMirror; out.albedo = al;
    } else {
        out.typ = matLambert; out.albedo = al; out.rough = clampd(m.rough, 0, 1);
    }
    return out;
}

template <class R> Material<R> castMat(const Material<double>& m) {
    Material<R> o;
    o.typ = m.typ;
    o.albedo = {(R)m.albedo.x, (R)m.albedo.y, (R)m.albedo.z};
    o.rough = (R)m.rough; o.ior = (R)m.ior;
    o.emit = {(R)m.emit.x, (R)m.emit.y, (R)m.emit.z};
    o.absorption = {(R)m.absorption.x, (R)m.absorption.y, (R)m.absorption.z};
    return o;
}

// ---------------------------------------------------------------- primitives (objects.go:9-222)
enum { objSphere = 0, objPlane = 1, objBox = 2, objMesh = 3 };

// ---------------------------------------------------------------- EXTENSION: triangle meshes (not in the reference)
struct OTri { float v0[3], e1[3], e2[3]; int32_t id; };      // e1 = v1 - v0, e2 = v2 - v0 in binary32, like the device records
struct OBvhNode { double lo[3], hi[3]; int32_t left, right, first, count; };   // leaf: count > 0
struct OMesh {
    std::vector<OTri> tris;
    std::vector<OBvhNode> nodes;      // median-split BVH, the oracle's own (independent of the product's SAH builder)
    std::vector<int32_t> order;       // triangle permutation of the BVH leaves
    bool accel = true;
    int build(int first, int count, double pad) {
        OBvhNode n{};
        for (int k = 0; k < 3; k++) { n.lo[k] = 1e300; n.hi[k] = -1e300; }
        double clo[3] = {1e300, 1e300, 1e300}, chi[3] = {-1e300, -1e300, -1e300};
        for (int i = first; i < first + count; i++) {
            const OTri& t = tris[order[i]];
            for (int k = 0; k < 3; k++) {
                const double a = t.v0[k], b = (double)t.v0[k] + t.e1[k], c = (double)t.v0[k] + t.e2[k];
                n.lo[k] = std::min(n.lo[k], std::min(a, std::min(b, c)) - pad);
                n.hi[k] = std::max(n.hi[k], std::max(a, std::max(b, c)) + pad);
                const double ce = (a + b + c) / 3;
                clo[k] = std::min(clo[k], ce); chi[k] = std::max(chi[k], ce);
            }
        }
        n.first = first; n.count = count; n.left = n.right = -1;
        const int me = (int)nodes.size();
        nodes.push_back(n);
        if (count > 8) {
            int ax = 0;
            if (chi[1] - clo[1] > chi[ax] - clo[ax]) ax = 1;
            if (chi[2] - clo[2] > chi[ax] - clo[ax]) ax = 2;
            const int mid = first + count / 2;
            std::nth_element(order.begin() + first, order.begin() + mid, order.begin() + first + count, [&](int32_t a, int32_t b) {
                const OTri& A = tris[a]; const OTri& B = tris[b];
                return 3.0 * A.v0[ax] + A.e1[ax] + A.e2[ax] < 3.0 * B.v0[ax] + B.e1[ax] + B.e2[ax];
            });
            const int l = build(first, mid - first, pad), r = build(mid, first + count - mid, pad);
            nodes[me].left = l; nodes[me].right = r; nodes[me].count = 0;
        }
        return me;
    }
    void finish() {
        order.resize(tris.size());
        for (size_t i = 0; i < tris.size(); i++) order[i] = (int32_t)i;
        double amax = 1;
        for (auto& t : tris) for (int k = 0; k < 3; k++) amax = std::max(amax, std::fabs((double)t.v0[k]) + std::fabs((double)t.e1[k]) + std::fabs((double)t.e2[k]));
        nodes.clear();
        if (!tris.empty()) build(0, (int)tris.size(), 1e-5 * amax);
    }
};

// Moeller-Trumbore in Real, operation order shared with the CUDA kernels (primary_fp64.cu / bvh.cuh).
template <class R> inline bool hitTri(const OTri& tr, const R o[3], const R d[3], R tMin, R tMax, R& tOut) {
    const R e1x = tr.e1[0], e1y = tr.e1[1], e1z = tr.e1[2], e2x = tr.e2[0], e2y = tr.e2[1], e2z = tr.e2[2];
    const R px = d[1] * e2z - d[2] * e2y, py = d[2] * e2x - d[0] * e2z, pz = d[0] * e2y - d[1] * e2x;
    const R det = e1x * px + e1y * py + e1z * pz;
    if (det == 0) return false;
    const R idet = (R)1.0 / det;
    const R tx = o[0] - (R)tr.v0[0], ty = o[1] - (R)tr.v0[1], tz = o[2] - (R)tr.v0[2];
    const R u = (tx * px + ty * py + tz * pz) * idet;
    if (u < 0 || u > 1) return false;
    const R qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;
    const R v = (d[0] * qx + d[1] * qy + d[2] * qz) * idet;
    if (v < 0 || u + v > 1) return false;
    const R t = (e2x * qx + e2y * qy + e2z * qz) * idet;
    if (t < tMin || t > tMax) return false;
    tOut = t;
    return true;
}
template <class R> struct HitRecord {                                                                  // objects.go:9-15
    V3<R> p{0, 0, 0}, normal{0, 0, 0};
    R t = 0;
    bool frontFace = false;
    const Material<R>* mat = nullptr;   // the reference copies the material by value; a pointer is equivalent
    int index = -1;                     // world index (bookkeeping only)
};
template <class R> struct Object {
    int type;
    V3<R> a, b;     // sphere: a=centre, b.x=radius | plane: a=point, b=normal | box: a=min, b=max | mesh: bounding box
    Material<R> mat;
    const OMesh* mesh = nullptr;   // EXTENSION
};

// Closest triangle of a mesh for t in [tMin, closest): updates closest / best_tri (lowest id wins exact ties).
template <class R> inline bool hitMesh(const OMesh& m, const Ray<R>& r, R tMin, R& closest, int& best_tri, const OTri*& hit) {
    const R o[3] = {r.orig.x, r.orig.y, r.orig.z}, d[3] = {r.dir.x, r.dir.y, r.dir.z};
    bool any = false;
    auto test = [&](const OTri& tr) {
        R t;
        if (!hitTri<R>(tr, o, d, tMin, closest, t)) return;
        if (t < closest || (best_tri >= 0 && tr.id < best_tri)) { closest = t; best_tri = tr.id; hit = &tr; any = true; }
    };
    if (!m.accel || m.tris.size() <= 64) { for (auto& tr : m.tris) test(tr); return any; }
    int stack[128]; int sp = 0; stack[sp++] = 0;
    while (sp) {
        const OBvhNode& n = m.nodes[stack[--sp]];
        double t0 = (double)tMin, t1 = (double)closest;
        bool miss = false;
        for (int k = 0; k < 3 && !miss; k++) {
            const double invD = 1.0 / (double)d[k];
            double tn = (n.lo[k] - (double)o[k]) * invD, tf = (n.hi[k] - (double)o[k]) * invD;
            if (invD < 0) std::swap(tn, tf);
            if (tn > t0) t0 = tn;
            if (tf < t1) t1 = tf;
            if (t1 < t0) miss = true;
        }
        if (miss) continue;
        if (n.count > 0) { for (int i = n.first; i < n.first + n.count; i++) test(m.tris[m.order[i]]); }
        else { stack[sp++] = n.left; stack[sp++] = n.right; }
    }
    return any;
}

template <class R> inline bool hitSphere(const Object<R>& s, const Ray<R>& r, R tMin, R tMax, HitRecord<R>& rec) {   // objects.go:37-89
    R ocX = r.orig.x - s.a.x, ocY = r.orig.y - s.a.y, ocZ = r.orig.z - s.a.z;
    R a = r.dir.x * r.dir.x + r.dir.y * r.dir.y + r.dir.z * r.dir.z;
    R halfB = ocX * r.dir.x + ocY * r.dir.y + ocZ * r.dir.z;
    R ocLenSq = ocX * ocX + ocY * ocY + ocZ * ocZ;
    R radius = s.b.x;
    R radiusSq = radius * radius;
    R c = ocLenSq - radiusSq;
    R disc = halfB * halfB - a * c;
    if (disc < 0) return false;
    R sqrtD = std::sqrt(disc);
    R root = (-halfB - sqrtD) / a;
    if (root < tMin || root > tMax) {
        root = (-halfB + sqrtD) / a;
        if (root < tMin || root > tMax) return false;
    }
    rec.t = root;
    rec.p.x = r.orig.x + r.dir.x * root;
    rec.p.y = r.orig.y + r.dir.y * root;
    rec.p.z = r.orig.z + r.dir.z * root;
    R invRadius = (R)1.0 / radius;
    R nx = (rec.p.x - s.a.x) * invRadius, ny = (rec.p.y - s.a.y) * invRadius, nz = (rec.p.z - s.a.z) * invRadius;
    R d = r.dir.x * nx + r.dir.y * ny + r.dir.z * nz;
    rec.frontFace = d < 0;
    if (rec.frontFace) rec.normal = {nx, ny, nz}; else rec.normal = {-nx, -ny, -nz};
    rec.mat = &s.mat;
    return true;
}
template <class R> inline bool hitPlane(const Object<R>& p, const Ray<R>& r, R tMin, R tMax, HitRecord<R>& rec) {    // objects.go:98-133
    const V3<R>& n = p.b;
    R denom = n.x * r.dir.x + n.y * r.dir.y + n.z * r.dir.z;
    if (std::fabs(denom) < (R)1e-6) return false;
    R mx = p.a.x - r.orig.x, my = p.a.y - r.orig.y, mz = p.a.z - r.orig.z;
    R t = (mx * n.x + my * n.y + mz * n.z) / denom;
    if (t < tMin || t > tMax) return false;
    rec.t = t;
    rec.p.x = r.orig.x + r.dir.x * t;
    rec.p.y = r.orig.y + r.dir.y * t;
    rec.p.z = r.orig.z + r.dir.z * t;
    rec.frontFace = denom < 0;
    if (rec.frontFace) rec.normal = n; else rec.normal = {-n.x, -n.y, -n.z};
    rec.mat = &p.mat;
    return true;
}
template <class R> inline bool hitBox(const Object<R>& b, const Ray<R>& r, R tMin, R tMax, HitRecord<R>& rec) {      // objects.go:141-222
    R t0 = tMin, t1 = tMax;
    for (int i = 0; i < 3; i++) {
        R invD, orig, minV, maxV;
        switch (i) {
        case 0: invD = 1 / r.dir.x; orig = r.orig.x; minV = b.a.x; maxV = b.b.x; break;
        case 1: invD = 1 / r.dir.y; orig = r.orig.y; minV = b.a.y; maxV = b.b.y; break;
        default: invD = 1 / r.dir.z; orig = r.orig.z; minV = b.a.z; maxV = b.b.z; break;
        }
        R tNear = (minV - orig) * invD;
        R tFar = (maxV - orig) * invD;
        if (invD < 0) { R tmp = tNear; tNear = tFar; tFar = tmp; }
        if (tNear > t0) t0 = tNear;
        if (tFar < t1) t1 = tFar;
        if (t1 <= t0) return false;
    }
    rec.t = t0;
    rec.p = add(r.orig, mul(r.dir, t0));                                     // r.at(t0), math.go:138
    R dxMin = rec.p.x - b.a.x, dxMax = b.b.x - rec.p.x;
    R dyMin = rec.p.y - b.a.y, dyMax = b.b.y - rec.p.y;
    R dzMin = rec.p.z - b.a.z, dzMax = b.b.z - rec.p.z;
    R minDist = dxMin;
    V3<R> n = {-1, 0, 0};
    if (dxMax < minDist) { minDist = dxMax; n = {1, 0, 0}; }
    if (dyMin < minDist) { minDist = dyMin; n = {0, -1, 0}; }
    if (dyMax < minDist) { minDist = dyMax; n = {0, 1, 0}; }
    if (dzMin < minDist) { minDist = dzMin; n = {0, 0, -1}; }
    if (dzMax < minDist) { n = {0, 0, 1}; }
    rec.frontFace = dot(r.dir, n) < 0;                                       // setFaceNormal, objects.go:17-24
    if (rec.frontFace) rec.normal = n; else rec.normal = mul(n, (R)-1);
    rec.mat = &b.mat;
    return true;
}
template <class R> inline bool hitObject(const Object<R>& o, const Ray<R>& r, R tMin, R tMax, HitRecord<R>& rec) {
    switch (o.type) {
    case objSphere: return hitSphere(o, r, tMin, tMax, rec);
    case objPlane: return hitPlane(o, r, tMin, tMax, rec);
    case objMesh: return false;      // meshes are scanned by closestHit() after the analytic objects
    default: return hitBox(o, r, tMin, tMax, rec);
    }
}

// ---------------------------------------------------------------- scatter (materials.go:67-231)
template <class R> inline R reflectance(R cosine, R refIdx) {                                          // materials.go:226-231
    R r0 = (1 - refIdx) / (1 + refIdx);
    r0 = r0 * r0;
    return r0 + (1 - r0) * std::pow(1 - cosine, (R)5);
}
template <class R> inline V3<R> emitted(const Material<R>& m) {                                        // materials.go:67-72
    if (m.typ == matEmissive) return m.emit;
    return {0, 0, 0};
}
template <class R>
inline bool scatter(const Material<R>& m, Rng& rng, const Ray<R>& rIn, const HitRecord<R>& rec,
                    V3<R>& attenuation, Ray<R>& scattered) {                                           // materials.go:74-224
    switch (m.typ) {
    case matLambert: {                                                                                 // :76-97
        V3<R> d = randomCosineDirection(rec.normal, rng);
        if (m.rough > (R)1e-6) {
            V3<R> off = randomInUnitSphere<R>(rng);
            d.x += off.x * m.rough * (R)0.1;
            d.y += off.y * m.rough * (R)0.1;
            d.z += off.z * m.rough * (R)0.1;
            d = unit(d);
        }
        scattered = {rec.p, d};
        attenuation = m.albedo;
        return true;
    }
    case matMetal: {                                                                                   // :99-160
        R dirLen = std::sqrt(rIn.dir.x * rIn.dir.x + rIn.dir.y * rIn.dir.y + rIn.dir.z * rIn.dir.z);
        if (dirLen == 0) { attenuation = {0, 0, 0}; scattered = {rec.p, rIn.dir}; return false; }
        R invLen = (R)1.0 / dirLen;
        V3<R> ud = {rIn.dir.x * invLen, rIn.dir.y * invLen, rIn.dir.z * invLen};
        V3<R> refl = reflectVec(ud, rec.normal);
        attenuation = m.albedo;
        if (m.rough > (R)1e-6) {
            V3<R> s = randomCosineDirection(refl, rng);
            R alpha = m.rough * m.rough;
            R sx = refl.x * ((R)1.0 - alpha) + s.x * alpha;
            R sy = refl.y * ((R)1.0 - alpha) + s.y * alpha;
            R sz = refl.z * ((R)1.0 - alpha) + s.z * alpha;
            R lenSq = sx * sx + sy * sy + sz * sz;
            if (lenSq < (R)1e-8) { sx = refl.x; sy = refl.y; sz = refl.z; }
            else { R l = std::sqrt(lenSq); R il = (R)1.0 / l; sx *= il; sy *= il; sz *= il; }
            R dn = sx * rec.normal.x + sy * rec.normal.y + sz * rec.normal.z;
            if (dn <= 0) { sx = refl.x; sy = refl.y; sz = refl.z; }
            scattered = {rec.p, {sx, sy, sz}};
            return true;
        }
        scattered = {rec.p, refl};
        return true;
    }
    case matDielectric: {                                                                              // :162-200
        attenuation = {1, 1, 1};
        R ratio = rec.frontFace ? (R)1.0 / m.ior : m.ior;
        R dirLen = std::sqrt(rIn.dir.x * rIn.dir.x + rIn.dir.y * rIn.dir.y + rIn.dir.z * rIn.dir.z);
        if (dirLen == 0) { scattered = {rec.p, rIn.dir}; return false; }
        R invLen = (R)1.0 / dirLen;
        V3<R> ud = {rIn.dir.x * invLen, rIn.dir.y * invLen, rIn.dir.z * invLen};
        R cosTheta = std::fmin(-(ud.x * rec.normal.x + ud.y * rec.normal.y + ud.z * rec.normal.z), (R)1.0);
        R sinTheta = std::sqrt((R)1.0 - cosTheta * cosTheta);
        bool cannotRefract = ratio * sinTheta > (R)1.0;
        R reflectProb = reflectance(cosTheta, ratio);
        V3<R> direction;
        if (cannotRefract || reflectProb > rng.uniform<R>())      // '||' short-circuits: no draw when cannotRefract
            direction = reflectVec(ud, rec.normal);
        else
            direction = refractVec(ud, rec.normal, ratio);
        scattered = {rec.p, direction};
        return true;
    }
    case matEmissive:                                                                                  // :202-203
        attenuation = {0, 0, 0}; scattered = {{0, 0, 0}, {0, 0, 0}};
        return false;
    case matMirror: {                                                                                  // :205-221
        R dirLen = std::sqrt(rIn.dir.x * rIn.dir.x + rIn.dir.y * rIn.dir.y + rIn.dir.z * rIn.dir.z);
        if (dirLen == 0) { attenuation = {0, 0, 0}; scattered = {rec.p, rIn.dir}; return false; }
        R invLen = (R)1.0 / dirLen;
        V3<R> ud = {rIn.dir.x * invLen, rIn.dir.y * invLen, rIn.dir.z * invLen};
        scattered = {rec.p, reflectVec(ud, rec.normal)};
        attenuation = m.albedo;
        return true;
    }
    }
    attenuation = {0, 0, 0}; scattered = {{0, 0, 0}, {0, 0, 0}};
    return false;
}

// ---------------------------------------------------------------- sky (renderer.go:56-92)
enum { skyConst = 0, skyGradient = 1 };
template <class R> struct Sky { int kind; V3<R> color, horizon, zenith; };
template <class R> inline V3<R> background(const Sky<R>& s, const Ray<R>& r) {
    if (s.kind == skyGradient) {                                                                       // :59-79
        R dirLen = std::sqrt(r.dir.x * r.dir.x + r.dir.y * r.dir.y + r.dir.z * r.dir.z);
        if (dirLen == 0) return s.horizon;
        R t = (r.dir.y / dirLen + (R)1.0) * (R)0.5;
        if (t < 0) t = 0;
        if (t > 1) t = 1;
        return {s.horizon.x * (1 - t) + s.zenith.x * t,
                s.horizon.y * (1 - t) + s.zenith.y * t,
                s.horizon.z * (1 - t) + s.zenith.z * t};
    }
    return s.color;                                                                                    // :80-92
}

// ---------------------------------------------------------------- camera (camera.go:9-74)
template <class R> struct Camera {
    V3<R> origin, lowerLeftCorner, horizontal, vertical, u, v, w;
    R lensRadius;
};
Camera<double> newCamera(const orc_raw_camera& c, int width, int height) {                            // camera.go:19-58
    double aspect = (double)width / (double)height;
    if (c.aspect_ratio != 0) aspect = c.aspect_ratio;
    double theta = c.fov * M_PI / 180;
    double h = std::tan(theta / 2);     // Go: math.Tan (pure-Go Cephes port); glibc tan may differ by <=1 ulp
    double viewportHeight = 2.0 * h;
    double viewportWidth = aspect * viewportHeight;
    V3<double> origin = {c.position[0], c.position[1], c.position[2]};
    V3<double> target = {c.target[0], c.target[1], c.target[2]};
    V3<double> up = {c.up[0], c.up[1], c.up[2]};
    V3<double> w = unit(sub(origin, target));
    V3<double> u = unit(cross(up, w));
    V3<double> vVec = cross(w, u);
    double focusDist = c.focus_dist;
    if (focusDist == 0) focusDist = length(sub(origin, target));
    V3<double> horizontal = mul(u, viewportWidth * focusDist);
    V3<double> vertical = mul(vVec, viewportHeight * focusDist);
    V3<double> llc = sub(sub(sub(origin, divv(horizontal, 2.0)), divv(vertical, 2.0)), mul(w, focusDist));
    return {origin, llc, horizontal, vertical, u, vVec, w, c.aperture / 2};
}
template <class R> Camera<R> castCam(const Camera<double>& c) {
    auto cv = [](V3<double> a) { return V3<R>{(R)a.x, (R)a.y, (R)a.z}; };
    return {cv(c.origin), cv(c.lowerLeftCorner), cv(c.horizontal), cv(c.vertical), cv(c.u), cv(c.v), cv(c.w), (R)c.lensRadius};
}
template <class R> inline Ray<R> getRay(const Camera<R>& c, R s, R t, Rng* rng) {                     // camera.go:60-74
    if (c.lensRadius > 0 && rng != nullptr) {
        V3<R> rd = mul(randomInUnitSphere<R>(*rng), c.lensRadius);
        V3<R> offset = add(mul(c.u, rd.x), mul(c.v, rd.y));
        return {add(c.origin, offset),
                sub(sub(add(add(c.lowerLeftCorner, mul(c.horizontal, s)), mul(c.vertical, t)), c.origin), offset)};
    }
    return {c.origin, sub(add(add(c.lowerLeftCorner, mul(c.horizontal, s)), mul(c.vertical, t)), c.origin)};
}

// Closest-hit scan of renderer.go:292-302 over the analytic objects, then (EXTENSION) over the meshes.
template <class R>
inline bool closestHit(const std::vector<Object<R>>& world, const Ray<R>& r, R tMin, HitRecord<R>& rec, orc_stats* st) {
    bool hitAnything = false;
    R closest = std::numeric_limits<R>::max();
    bool any_mesh = false;
    for (size_t i = 0; i < world.size(); i++) {
        if (world[i].type == objMesh) { any_mesh = true; continue; }
        if (st) st->prim_tests++;
        if (hitObject(world[i], r, tMin, closest, rec)) {
            hitAnything = true; closest = rec.t; rec.index = (int)i;
            if (st) st->accepts[world[i].type]++;
        }
    }
    if (any_mesh) {
        int best_tri = -1; const OTri* tri = nullptr; int mesh_obj = -1;
        for (size_t i = 0; i < world.size(); i++) {
            if (world[i].type != objMesh) continue;
            if (hitMesh<R>(*world[i].mesh, r, tMin, closest, best_tri, tri)) mesh_obj = (int)i;
        }
        if (mesh_obj >= 0) {
            hitAnything = true;
            rec.t = closest; rec.index = mesh_obj; rec.mat = &world[mesh_obj].mat;
            rec.p = {r.orig.x + r.dir.x * closest, r.orig.y + r.dir.y * closest, r.orig.z + r.dir.z * closest};
            V3<R> e1 = {(R)tri->e1[0], (R)tri->e1[1], (R)tri->e1[2]}, e2 = {(R)tri->e2[0], (R)tri->e2[1], (R)tri->e2[2]};
            V3<R> g = unit(cross(e1, e2));
            rec.frontFace = dot(r.dir, g) < 0;                 // setFaceNormal, objects.go:17-24
            rec.normal = rec.frontFace ? g : mul(g, (R)-1);
        }
    }
    return hitAnything;
}

// ---------------------------------------------------------------- integrator (renderer.go:286-404)
struct PathLog { int cap = 0, n = 0; int32_t* ids = nullptr; double* t = nullptr; int32_t* ff = nullptr; };

template <class R>
V3<R> rayColorOpt(const Ray<R>& r, const std::vector<Object<R>>& world, const Sky<R>& sky, int depth,
                  Rng& rng, orc_stats& st, PathLog* log) {
    if (depth <= 0) { st.end_depth++; return {0, 0, 0}; }                                              // :287-289
    const R tMin = (R)0.001;                                                                           // :292
    HitRecord<R> rec;
    st.segments++;
    const bool hitAnything = closestHit(world, r, tMin, rec, &st);                                     // :294-302
    if (log && log->n < log->cap) {
        log->ids[log->n] = hitAnything ? rec.index : -1;
        log->t[log->n] = hitAnything ? (double)rec.t : 0.0;
        log->ff[log->n] = hitAnything ? (rec.frontFace ? 1 : 0) : 0;
        log->n++;
    }
    if (!hitAnything) { st.end_sky++; return background(sky, r); }                                    // :304-306

    const Material<R>& mat = *rec.mat;
    V3<R> em = emitted(mat);                                                                           // :308
    V3<R> attenuation; Ray<R> scattered;
    bool ok = scatter(mat, rng, r, rec, attenuation, scattered);                                      // :309
    if (!ok) { if (mat.typ == matEmissive) st.end_emissive++; else st.end_noscatter++; return em; }   // :310-312
    st.scatters++;

    if (mat.typ == matDielectric) {                                                                    // :316
        if (rec.frontFace) {                                                                           // :319
            const R exitTMin = (R)0.0001;                                                              // :322
            HitRecord<R> exitRec;
            bool hitExit = false;
            R exitT = std::numeric_limits<R>::max();
            st.exit_scans++;
            for (size_t i = 0; i < world.size(); i++) {                                                // :329-349
                HitRecord<R> tempRec;
                st.prim_tests++;
                if (hitObject(world[i], scattered, exitTMin, exitT, tempRec)) {
                    if (tempRec.mat->typ == matDielectric && !tempRec.frontFace && tempRec.t < exitT) {
                        R dx = tempRec.p.x - rec.p.x, dy = tempRec.p.y - rec.p.y, dz = tempRec.p.z - rec.p.z;
                        R distSq = dx * dx + dy * dy + dz * dz;
                        if (distSq > (R)1e-8 && distSq < (R)1000.0) {
                            hitExit = true;
                            exitT = tempRec.t;
                            exitRec = tempRec;
                        }
                    }
                }
            }
            if (hitExit) {                                                                             // :352-369
                R dx = exitRec.p.x - rec.p.x, dy = exitRec.p.y - rec.p.y, dz = exitRec.p.z - rec.p.z;
                R distance = std::sqrt(dx * dx + dy * dy + dz * dz);
                if (mat.absorption.x > 0 || mat.absorption.y > 0 || mat.absorption.z > 0) {
                    attenuation.x = std::exp(-mat.absorption.x * distance);
                    attenuation.y = std::exp(-mat.absorption.y * distance);
                    attenuation.z = std::exp(-mat.absorption.z * distance);
                }
                scattered.orig = exitRec.p;
            }
        }
    }

    const int rrThreshold = 3;                                                                         // :374
    if (depth <= rrThreshold) {                                                                        // :375-393
        R maxAtt = std::fmax(attenuation.x, std::fmax(attenuation.y, attenuation.z));
        if (maxAtt < (R)1e-6) { st.end_rr++; return em; }
        R rrProb = std::fmin(maxAtt, (R)0.95);
        if (rng.uniform<R>() > rrProb) { st.end_rr++; return em; }
        attenuation.x /= rrProb; attenuation.y /= rrProb; attenuation.z /= rrProb;
    }

    V3<R> next = rayColorOpt(scattered, world, sky, depth - 1, rng, st, log);                         // :398
    return {em.x + attenuation.x * next.x, em.y + attenuation.y * next.y, em.z + attenuation.z * next.z};   // :399-403
}

}  // namespace

// ---------------------------------------------------------------- scene handle
struct orc_scene {
    std::vector<std::unique_ptr<OMesh>> meshes;   // EXTENSION
    std::vector<Object<double>> world;       // sceneToWorld result (objects.go:225-269)
    orc_raw_camera cam;
    Sky<double> sky;
};

namespace {
template <class R> std::vector<Object<R>> castWorld(const std::vector<Object<double>>& w) {
    std::vector<Object<R>> out;
    out.reserve(w.size());
    auto cv = [](V3<double> a) { return V3<R>{(R)a.x, (R)a.y, (R)a.z}; };
    for (auto& o : w) out.push_back({o.type, cv(o.a), cv(o.b), castMat<R>(o.mat), o.mesh});
    return out;
}
template <class R> Sky<R> castSky(const Sky<double>& s) {
    auto cv = [](V3<double> a) { return V3<R>{(R)a.x, (R)a.y, (R)a.z}; };
    return {s.kind, cv(s.color), cv(s.horizon), cv(s.zenith)};
}
void addStats(orc_stats& a, const orc_stats& b) {
    a.samples += b.samples; a.segments += b.segments; a.exit_scans += b.exit_scans; a.prim_tests += b.prim_tests;
    for (int i = 0; i < 3; i++) a.accepts[i] += b.accepts[i];
    a.scatters += b.scatters; a.end_sky += b.end_sky; a.end_emissive += b.end_emissive; a.end_rr += b.end_rr;
    a.end_depth += b.end_depth; a.end_noscatter += b.end_noscatter;
}

// Pixel loop over a tile queue (renderer.go:114-238) for samples [s0,s1); sums are left un-normalised.
template <class R>
void renderSum(const orc_scene* sc, int W, int H, int s0, int s1, int maxDepth, uint32_t seed, int threads,
               double* rgbSum, orc_stats* statsOut) {
    const std::vector<Object<R>> world = castWorld<R>(sc->world);
    const Sky<R> sky = castSky<R>(sc->sky);
    const Camera<R> cam = castCam<R>(newCamera(sc->cam, W, H));
    const R invWidth = (R)1.0 / (R)(W - 1);                                                            // :95
    const R invHeight = (R)1.0 / (R)(H - 1);                                                           // :96
    const R heightMinus1 = (R)(H - 1);                                                                 // :98
    const int tileSize = 32;                                                                           // :132
    const int ntx = (W + tileSize - 1) / tileSize, nty = (H + tileSize - 1) / tileSize;
    std::atomic<int> nextTile{0};
    if (threads < 1) threads = 1;
    std::vector<orc_stats> perThread(threads);
    for (auto& s : perThread) std::memset(&s, 0, sizeof s);
    auto worker = [&](int tid) {
        orc_stats& st = perThread[tid];
        for (;;) {
            int t = nextTile.fetch_add(1);
            if (t >= ntx * nty) break;
            int tx = (t % ntx) * tileSize, ty = (t / ntx) * tileSize;                                  // row-major tiles (:146-160)
            int x1 = std::min(tx + tileSize, W), y1 = std::min(ty + tileSize, H);
            for (int y = ty; y < y1; y++) {
                R flipY = heightMinus1 - (R)y;                                                         // :174
                for (int x = tx; x < x1; x++) {
                    V3<R> col = {0, 0, 0};
                    R xFloat = (R)x;
                    for (int s = s0; s < s1; s++) {                                                    // :181-187
                        Rng rng(seed, (uint32_t)(y * W + x), (uint32_t)s);
                        R u = (xFloat + rng.uniform<R>()) * invWidth;
                        R vv = (flipY + rng.uniform<R>()) * invHeight;
                        Ray<R> r = getRay(cam, u, vv, &rng);
                        st.samples++;
                        col = add(col, rayColorOpt(r, world, sky, maxDepth, rng, st, nullptr));
                    }
                    double* o = rgbSum + ((size_t)y * W + x) * 3;
                    o[0] = (double)col.x; o[1] = (double)col.y; o[2] = (double)col.z;
                }
            }
        }
    };
    if (threads == 1) worker(0);
    else {
        std::vector<std::thread> pool;
        for (int i = 0; i < threads; i++) pool.emplace_back(worker, i);
        for (auto& th : pool) th.join();
    }
    if (statsOut) { std::memset(statsOut, 0, sizeof *statsOut); for (auto& s : perThread) addStats(*statsOut, s); }
}
}  // namespace

extern "C" {

orc_scene* orc_scene_create(const orc_raw_object* objs, int n_objs, const orc_raw_material* mats, int n_mats,
                            const orc_raw_camera* cam, const orc_raw_sky* sky) {
    orc_scene* sc = new orc_scene();
    std::map<std::string, Material<double>> materials;                                                 // objects.go:226-229 (later duplicate wins)
    for (int i = 0; i < n_mats; i++) materials[mats[i].id ? mats[i].id : ""] = convertMaterial(mats[i]);
    int32_t n_tri_total = 0;
    for (int i = 0; i < n_objs; i++) {                                                                 // objects.go:232-267
        const orc_raw_object& o = objs[i];
        Material<double> mat;                                                                          // missing id -> zero material
        auto it = materials.find(o.material_id ? o.material_id : "");
        if (it != materials.end()) mat = it->second;
        V3<double> pos = {o.position[0], o.position[1], o.position[2]};
        V3<double> size = {o.size[0], o.size[1], o.size[2]};
        std::string t = o.type ? o.type : "";
        if (t == "sphere" || t == "sphere_light") sc->world.push_back({objSphere, pos, {size.x, 0, 0}, mat});
        else if (t == "plane") sc->world.push_back({objPlane, pos, {0, 1, 0}, mat});
        else if (t == "box") sc->world.push_back({objBox, sub(pos, mul(size, 0.5)), add(pos, mul(size, 0.5)), mat});
        else if (t == "mesh" && o.tri_vertices && o.n_tri > 0) {          // EXTENSION
            auto m = std::make_unique<OMesh>();
            V3<double> lo = {1e300, 1e300, 1e300}, hi = {-1e300, -1e300, -1e300};
            for (int64_t q = 0; q < o.n_tri; q++) {
                const float* v = o.tri_vertices + 9 * q;
                OTri tr;
                for (int k = 0; k < 3; k++) { tr.v0[k] = v[k]; tr.e1[k] = v[3 + k] - v[k]; tr.e2[k] = v[6 + k] - v[k]; }
                tr.id = n_tri_total++;
                m->tris.push_back(tr);
                for (int c = 0; c < 3; c++) {
                    lo.x = std::min(lo.x, (double)v[3 * c]); lo.y = std::min(lo.y, (double)v[3 * c + 1]); lo.z = std::min(lo.z, (double)v[3 * c + 2]);
                    hi.x = std::max(hi.x, (double)v[3 * c]); hi.y = std::max(hi.y, (double)v[3 * c + 1]); hi.z = std::max(hi.z, (double)v[3 * c + 2]);
                }
            }
            m->finish();
            sc->world.push_back({objMesh, lo, hi, mat, m.get()});
            sc->meshes.push_back(std::move(m));
        }
        /* unknown object types are dropped */
    }
    sc->cam = *cam;
    std::string st = (sky->has_sky && sky->sky_type) ? sky->sky_type : "";
    auto c3 = [](const double* p) { return V3<double>{p[0], p[1], p[2]}; };
    if (sky->has_sky && st == "gradient") sc->sky = {skyGradient, {0, 0, 0}, c3(sky->horizon), c3(sky->zenith)};   // renderer.go:56-79
    else if (sky->has_sky && st == "solid") sc->sky = {skyConst, c3(sky->color), {0, 0, 0}, {0, 0, 0}};             // :83-84
    else sc->sky = {skyConst, c3(sky->background), {0, 0, 0}, {0, 0, 0}};                                           // :85-87
    return sc;
}
void orc_scene_destroy(orc_scene* sc) { delete sc; }
int orc_world_size(const orc_scene* sc) { return (int)sc->world.size(); }
void orc_world_get(const orc_scene* sc, int i, orc_world_entry* out) {
    const Object<double>& o = sc->world[i];
    out->type = o.type; out->mat_type = o.mat.typ;
    out->a[0] = o.a.x; out->a[1] = o.a.y; out->a[2] = o.a.z;
    out->b[0] = o.b.x; out->b[1] = o.b.y; out->b[2] = o.b.z;
    out->albedo[0] = o.mat.albedo.x; out->albedo[1] = o.mat.albedo.y; out->albedo[2] = o.mat.albedo.z;
    out->rough = o.mat.rough; out->ior = o.mat.ior;
    out->emit[0] = o.mat.emit.x; out->emit[1] = o.mat.emit.y; out->emit[2] = o.mat.emit.z;
    out->absorption[0] = o.mat.absorption.x; out->absorption[1] = o.mat.absorption.y; out->absorption[2] = o.mat.absorption.z;
}

void orc_camera(const orc_scene* sc, int width, int height, double out[22]) {
    Camera<double> c = newCamera(sc->cam, width, height);
    const V3<double>* v[7] = {&c.origin, &c.lowerLeftCorner, &c.horizontal, &c.vertical, &c.u, &c.v, &c.w};
    for (int i = 0; i < 7; i++) { out[3 * i] = v[i]->x; out[3 * i + 1] = v[i]->y; out[3 * i + 2] = v[i]->z; }
    out[21] = c.lensRadius;
}

void orc_primary_hits(const orc_scene* sc, int W, int H, double xi_u, double xi_v, int32_t* ids, double* tOut) {
    const Camera<double> cam = newCamera(sc->cam, W, H);
    const double invWidth = 1.0 / (double)(W - 1), invHeight = 1.0 / (double)(H - 1), heightMinus1 = (double)(H - 1);
    for (int y = 0; y < H; y++) {
        double flipY = heightMinus1 - (double)y;
        for (int x = 0; x < W; x++) {
            double u = ((double)x + xi_u) * invWidth;
            double vv = (flipY + xi_v) * invHeight;
            Ray<double> r = getRay<double>(cam, u, vv, nullptr);                                       // camera.go:70-73
            HitRecord<double> rec;
            const bool hitAnything = closestHit<double>(sc->world, r, 0.001, rec, nullptr);            // renderer.go:292-302
            ids[(size_t)y * W + x] = hitAnything ? rec.index : -1;
            tOut[(size_t)y * W + x] = hitAnything ? rec.t : 0.0;
        }
    }
}

void orc_render_sum(const orc_scene* sc, int W, int H, int s_begin, int s_end, int max_depth, uint32_t seed,
                    int precision, int threads, double* rgb_sum, orc_stats* stats) {
    if (precision == 32) renderSum<float>(sc, W, H, s_begin, s_end, max_depth, seed, threads, rgb_sum, stats);
    else renderSum<double>(sc, W, H, s_begin, s_end, max_depth, seed, threads, rgb_sum, stats);
}

void orc_finalize(const double* rgb_sum, int W, int H, int spp, uint8_t* rgba) {                      // renderer.go:189-221
    const double invSamples = 1.0 / (double)spp;                                                       // :97
    for (size_t i = 0; i < (size_t)W * H; i++) {
        for (int c = 0; c < 3; c++) {
            double v = rgb_sum[i * 3 + c] * invSamples;
            v = std::sqrt(v);
            v = v * 255.999;
            if (v < 0) v = 0; else if (v > 255.999) v = 255.999;
            rgba[i * 4 + c] = (v == v) ? (uint8_t)v : 0;    /* uint8(NaN) on amd64 truncates to 0 */
        }
        rgba[i * 4 + 3] = 255;
    }
}

void orc_render_rgba(const orc_scene* sc, int W, int H, int spp, int max_depth, uint32_t seed, int threads,
                     uint8_t* rgba, orc_stats* stats) {
    std::vector<double> sum((size_t)W * H * 3);
    renderSum<double>(sc, W, H, 0, spp, max_depth, seed, threads, sum.data(), stats);
    orc_finalize(sum.data(), W, H, spp, rgba);
}

int orc_trace_path(const orc_scene* sc, const double orig[3], const double dir[3], int max_depth, uint32_t seed,
                   int cap, int32_t* hit_ids, double* hit_t, int32_t* front_face, double rgb[3]) {
    const Sky<double> sky = sc->sky;
    Rng rng(seed, 0, 0);
    orc_stats st; std::memset(&st, 0, sizeof st);
    PathLog log; log.cap = cap; log.ids = hit_ids; log.t = hit_t; log.ff = front_face;
    Ray<double> r = {{orig[0], orig[1], orig[2]}, {dir[0], dir[1], dir[2]}};
    V3<double> c = rayColorOpt<double>(r, sc->world, sky, max_depth, rng, st, &log);
    rgb[0] = c.x; rgb[1] = c.y; rgb[2] = c.z;
    return log.n;
}

void orc_set_mesh_accel(orc_scene* sc, int enabled) { for (auto& m : sc->meshes) m->accel = enabled != 0; }

double orc_rng_uniform(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t i) {
    Rng rng(seed, pixel, sample);
    rng.ctr = i;
    return rng.uniform<double>();
}

}  // extern "C"
