// bvh.h — BVH over the triangles of the scene's meshes (EXTENSION; the reference has no triangles and no BVH —
// north-star: "internal/scene gains a BVH builder that emits a flattened, cache-line-aligned node array").
//
// Layout (device, global memory, fetched with 16-byte loads):
//   node  = 64 bytes = 4 x float4, cache-line aligned:  both children's boxes + both child links, so ONE node fetch
//           decides both children ("Aila-Laine" BVH2 layout)
//             boxes as (centre c, half extent h), the form whose slab test needs no per-axis min/max (integrator.cu hit_box):
//             q0 = (c0.x, c0.y, c0.z, h0.x)   q1 = (h0.y, h0.z, c1.x, c1.y)   q2 = (c1.z, h1.x, h1.y, h1.z)
//             q3 = (bits l0, bits l1, 0, 0)   link l >= 0: inner node index;  l < 0: leaf, ~l = first_tri << 2 | (count - 1)
//           an empty child has h = -1 (never hit) and link kEmptyLeaf
//   tri   = 48 bytes = 3 x float4, in leaf order:  (v0.xyz, bits tri_id)  (e1.xyz, bits meta)  (e2.xyz, bits world_idx)
//           e1 = v1 - v0, e2 = v2 - v0 (binary32);  meta = the DevObj::meta of the mesh's material
// Boxes are padded by 1e-5 x (largest absolute coordinate of the scene, at least 1) so that the fp32 slab test with
// approximate reciprocals stays conservative, also for axis-aligned (zero-thickness) triangles.
#pragma once
#include <cstdint>
#include <vector>

namespace ptb {

struct alignas(64) BvhNode { float q[16]; };
struct alignas(16) BvhTri { float q[12]; };
constexpr int32_t kEmptyLeaf = 0x7fffffff;   // never followed (its box has a negative half extent)
constexpr int kMaxLeafTris = 4;

struct BvhBuildInput {
    const float* tri_vertices;   // 9 floats per triangle, world space
    int64_t n_tri;
    const int32_t* tri_meta;     // per triangle: DevObj::meta of its mesh
    const int32_t* tri_world;    // per triangle: world index of its mesh object
};
struct BvhBuildOutput {
    std::vector<BvhNode> nodes;  // nodes[0] = root (always an inner node when n_tri > 0)
    std::vector<BvhTri> tris;    // leaf order
    int max_depth = 0;
    double sah_cost = 0;         // sum over inner nodes of area(node)/area(root) (+ leaves weighted by count)
    double build_ms = 0;
};
void build_bvh(const BvhBuildInput& in, BvhBuildOutput& out, int threads);

}  // namespace ptb
