// primary_fp64.cu — parity hook: closest hit of every pixel's primary ray in binary64.
//
// Restates, operation for operation, camera.getRay's lens-free branch (camera.go:70-73), the sample
// position of the pixel loop (renderer.go:95-98, 174, 182-183) and the closest-hit scan
// (renderer.go:292-302) over sphere.hit / plane.hit / box.hit (objects.go:37-59, 98-112, 141-183).
// THIS FILE MUST BE COMPILED WITH -fmad=false: Go's gc compiler does not contract a*b+c into an FMA on
// amd64, so neither may we, otherwise hit ids on silhouettes can differ.  Division and sqrt are IEEE
// (nvcc defaults -prec-div=true -prec-sqrt=true; binary64 has no approximate variants anyway).
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "scene_dev.h"

namespace ptb {

__device__ __forceinline__ bool hit64(const Obj64& ob, const double o[3], const double d[3], double tMin, double tMax,
                                      double& tOut) {
    if (ob.type == PTB_OBJ_SPHERE) {                                  // objects.go:37-59
        double ocX = o[0] - ob.a[0], ocY = o[1] - ob.a[1], ocZ = o[2] - ob.a[2];
        double a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        double halfB = ocX * d[0] + ocY * d[1] + ocZ * d[2];
        double ocLenSq = ocX * ocX + ocY * ocY + ocZ * ocZ;
        double radiusSq = ob.b[0] * ob.b[0];
        double c = ocLenSq - radiusSq;
        double disc = halfB * halfB - a * c;
        if (disc < 0) return false;
        double sqrtD = sqrt(disc);
        double root = (-halfB - sqrtD) / a;
        if (root < tMin || root > tMax) {
            root = (-halfB + sqrtD) / a;
            if (root < tMin || root > tMax) return false;
        }
        tOut = root;
        return true;
    } else if (ob.type == PTB_OBJ_PLANE) {                            // objects.go:98-112
        double denom = ob.b[0] * d[0] + ob.b[1] * d[1] + ob.b[2] * d[2];
        if (fabs(denom) < 1e-6) return false;
        double mx = ob.a[0] - o[0], my = ob.a[1] - o[1], mz = ob.a[2] - o[2];
        double t = (mx * ob.b[0] + my * ob.b[1] + mz * ob.b[2]) / denom;
        if (t < tMin || t > tMax) return false;
        tOut = t;
        return true;
    } else {                                                          // objects.go:141-183
        double t0 = tMin, t1 = tMax;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            double invD = 1 / d[i];
            double tNear = (ob.a[i] - o[i]) * invD;
            double tFar = (ob.b[i] - o[i]) * invD;
            if (invD < 0) { double s = tNear; tNear = tFar; tFar = s; }
            if (tNear > t0) t0 = tNear;
            if (tFar < t1) t1 = tFar;
            if (t1 <= t0) return false;
        }
        tOut = t0;
        return true;
    }
}

// EXTENSION (triangle meshes): binary64 traversal of the BVH of bvh.h.  Boxes are binary32 padded boxes widened to
// binary64 (conservative), triangles are the binary32 (v0, e1, e2) records widened to binary64; Moeller-Trumbore
// without FMA.  Result = the triangle of smallest t (lowest triangle id on ties), independent of the traversal order,
// so it equals the oracle's brute-force scan bit for bit.
__device__ __forceinline__ bool slab64(const double lo[3], const double hi[3], const double o[3], const double d[3], double tMin, double tMax) {
    double t0 = tMin, t1 = tMax;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double invD = 1 / d[i];
        double tNear = (lo[i] - o[i]) * invD;
        double tFar = (hi[i] - o[i]) * invD;
        if (invD < 0) { double s = tNear; tNear = tFar; tFar = s; }
        if (tNear > t0) t0 = tNear;
        if (tFar < t1) t1 = tFar;
    }
    return t1 >= t0;
}
__device__ void bvh_closest64(const float4* __restrict__ nodes, const float4* __restrict__ tris, const double o[3], const double d[3],
                              double tMin, double& closest, int& id) {
    int stack[48];
    int sp = 0, cur = 0, best_tri = -1;
    for (;;) {
        if (cur >= 0) {
            // 4-wide node (bvh.h): floats 6k..6k+5 = child k's (centre, half extent) in binary32 — c -+ h is exact in binary64 —
            // floats 24..27 = links.  The result does not depend on the visiting order: every hit child is pushed.
            float q[28];
#pragma unroll
            for (int v = 0; v < 7; ++v) { const float4 w = __ldg(nodes + 8 * cur + v); q[4 * v] = w.x; q[4 * v + 1] = w.y; q[4 * v + 2] = w.z; q[4 * v + 3] = w.w; }
            int next = 0x7fffffff;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float* b = q + 6 * k;
                const double lo[3] = {(double)b[0] - b[3], (double)b[1] - b[4], (double)b[2] - b[5]}, hi[3] = {(double)b[0] + b[3], (double)b[1] + b[4], (double)b[2] + b[5]};
                if (b[3] < 0.0f || !slab64(lo, hi, o, d, tMin, closest)) continue;
                const int link = __float_as_int(q[24 + k]);
                if (next == 0x7fffffff) next = link;
                else if (sp < 48) stack[sp++] = link;
            }
            if (next != 0x7fffffff) cur = next;
            else { if (sp == 0) break; cur = stack[--sp]; }
        } else {
            const int link = ~cur, first = link >> 2, cnt = (link & 3) + 1;
            for (int k = 0; k < cnt; ++k) {
                const float4 a = __ldg(tris + 4 * (first + k)), b = __ldg(tris + 4 * (first + k) + 1), c = __ldg(tris + 4 * (first + k) + 2);   // 64-byte records
                const double e1x = b.x, e1y = b.y, e1z = b.z, e2x = c.x, e2y = c.y, e2z = c.z;
                const double px = d[1] * e2z - d[2] * e2y, py = d[2] * e2x - d[0] * e2z, pz = d[0] * e2y - d[1] * e2x;
                const double det = e1x * px + e1y * py + e1z * pz;
                if (det == 0) continue;
                const double idet = 1.0 / det;
                const double tx = o[0] - (double)a.x, ty = o[1] - (double)a.y, tz = o[2] - (double)a.z;
                const double u = (tx * px + ty * py + tz * pz) * idet;
                if (u < 0 || u > 1) continue;
                const double qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;
                const double v = (d[0] * qx + d[1] * qy + d[2] * qz) * idet;
                if (v < 0 || u + v > 1) continue;
                const double t = (e2x * qx + e2y * qy + e2z * qz) * idet;
                if (t < tMin || t > closest) continue;
                const int tid = __float_as_int(a.w);
                if (t < closest || (best_tri >= 0 && tid < best_tri)) { closest = t; id = __float_as_int(c.w); best_tri = tid; }
            }
            if (sp == 0) break;
            cur = stack[--sp];
        }
    }
}

__global__ void primary_hits_kernel(const Obj64* __restrict__ world, int n_obj, const float4* __restrict__ bvh_nodes,
                                    const float4* __restrict__ bvh_tris, Camera64 cam, int W, int H,
                                    double xi_u, double xi_v, int32_t* __restrict__ ids, double* __restrict__ tt) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const double invWidth = 1.0 / (double)(W - 1);                    // renderer.go:95
    const double invHeight = 1.0 / (double)(H - 1);                   // renderer.go:96
    const double flipY = (double)(H - 1) - (double)y;                 // renderer.go:98,174
    const double u = ((double)x + xi_u) * invWidth;                   // renderer.go:182
    const double v = (flipY + xi_v) * invHeight;                      // renderer.go:183
    double o[3], d[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {                                     // camera.go:70-73
        o[k] = cam.origin[k];
        d[k] = cam.llc[k] + cam.horizontal[k] * u + cam.vertical[k] * v - cam.origin[k];
    }
    double closest = DBL_MAX;                                         // math.MaxFloat64, renderer.go:294
    int id = -1;
    for (int i = 0; i < n_obj; i++) {                                 // renderer.go:297-302
        double t;
        if (hit64(world[i], o, d, 0.001, closest, t)) { closest = t; id = world[i].pad; }   // pad = world index
    }
    if (bvh_nodes) bvh_closest64(bvh_nodes, bvh_tris, o, d, 0.001, closest, id);
    ids[(size_t)y * W + x] = id;
    tt[(size_t)y * W + x] = id >= 0 ? closest : 0.0;
}

int launch_primary_hits(const Obj64* d_world, int n_obj, const float4* bvh_nodes, const float4* bvh_tris, const Camera64& cam,
                        int width, int height, double xi_u, double xi_v, int32_t* d_ids, double* d_t, void* stream) {
    dim3 block(32, 8), grid((width + 31) / 32, (height + 7) / 8);
    primary_hits_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(d_world, n_obj, bvh_nodes, bvh_tris, cam, width, height, xi_u, xi_v, d_ids, d_t);
    return (int)cudaGetLastError();
}

}  // namespace ptb
