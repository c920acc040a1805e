// bvh.h — BVH over the triangles of the scene's meshes (EXTENSION; the reference has no triangles and no BVH —
// north-star: "internal/scene gains a BVH builder that emits a flattened, cache-line-aligned node array").
//
// Layout (device, global memory, fetched with 16-byte loads):
// Everything is sized for 256-bit loads (LDG.E.256, new on sm_100): a lane's scattered fetch costs the L1 one tag look-up per
// load INSTRUCTION whatever its width, and the traversal is bound by exactly that (ncu r02d: L1TEX 47 % busy at 45 % issue
// utilisation) — so a node is 4 loads, a triangle 2.
//   node  = 128 bytes = 8 x float4, aligned to the 128-byte L2 line: a 4-WIDE node — the boxes of up to four children and
//           their links, so ONE node fetch decides four subtrees and a ray needs about half the DEPENDENT fetches of a binary
//           tree (the traversal is bound by the latency of those fetches).  Built by collapsing the binned-SAH
//           binary tree: a node adopts its grandchildren, largest box first, until it has four children.
//             floats 6k .. 6k+5  (k = 0..3): child k's box as (centre c.xyz, half extent h.xyz) — the form whose slab test
//                                            needs no per-axis min/max (integrator.cu hit_box)
//             floats 24 .. 27              : bits of link k;  link >= 0: inner node index;  link < 0: leaf,
//                                            ~link = first_tri << 2 | (count - 1)
//             floats 28 .. 31              : 0 (pad)
//           an unused child has h = -1 (never hit) and link kEmptyLeaf
//   tri   = 64 bytes = 4 x float4, in leaf order:  (v0.xyz, bits tri_id)  (e1.xyz, bits meta)  (e2.xyz, bits world_idx)  (0, 0, 0, 0)
//           e1 = v1 - v0, e2 = v2 - v0 (binary32);  meta = the DevObj::meta of the mesh's material
// Boxes are padded by 1e-5 x (largest absolute coordinate of the scene, at least 1) so that the fp32 slab test with
// approximate reciprocals stays conservative, also for axis-aligned (zero-thickness) triangles.
#pragma once
#include <cstdint>
#include <vector>

namespace ptb {

constexpr int kBvhWidth = 4;
struct alignas(128) BvhNode { float q[32]; };
struct alignas(64) BvhTri { float q[16]; };    // 64 bytes: two 32-byte (256-bit) loads, naturally aligned
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
    int max_depth = 0;           // depth of the 4-wide tree (root = 1)
    int max_stack = 0;           // entries a depth-first traversal can have pending: sum over a root-to-leaf path of (children - 1)
    double sah_cost = 0;         // sum over inner nodes of area(node)/area(root) (+ leaves weighted by count)
    double build_ms = 0;
};
void build_bvh(const BvhBuildInput& in, BvhBuildOutput& out, int threads);

}  // namespace ptb
