// scene_dev.h — device-side scene layout shared by the host API (api.cu) and the kernels.
//
// The analytic world of the reference (objects.go:26-222: spheres, one-normal planes, boxes; at
// most a few dozen objects) is tiny, so it is not streamed from HBM at all: the fp32 copy lives in
// __constant__ memory.  The closest-hit scan reads it with a warp-uniform index (every lane tests
// object i at the same time), which the constant cache serves as a broadcast; the few divergent
// look-ups (the winning object / its material) go through a shared-memory copy made per CTA.
#pragma once
#include <stdint.h>
#include <vector_types.h>

#include "../../include/ptb200.h"

namespace ptb {

// 32-byte records, two 16-byte vector loads each.
struct alignas(16) DevObj {
    float ax, ay, az;   // sphere centre | plane point | box CENTRE         (objects.go:31-35, 92-96, 136-139)
    int32_t meta;       // bits 0..1 type, bit 2 dielectric material, bits 3..5 shading class (wavefront.cuh), bits 6.. material index
    float bx, by, bz;   // sphere (radius, radius^2, 1/radius) | plane unused | box HALF EXTENT (see hit_box)
    int32_t world_idx;  // index in the reference's world order (device order groups boxes first)
};

// 48-byte records. materials.go:19-26.
struct alignas(16) DevMat {
    float albedo[3];
    int32_t type;
    float emit[3];
    float rough;
    float absorption[3];
    float ior;
};

struct DevCamera {   // camera.go:9-17, binary32 copy of the binary64 newCamera result
    float origin[3], llc[3], horizontal[3], vertical[3], u[3], v[3];
    float lens_radius;
};

struct DevSky {      // renderer.go:56-92
    int32_t kind;
    float color[3], horizon[3], zenith[3];
};

constexpr int kMaxExitTyped = 64;
struct DevScene {                        // 57.3 KB of the 64 KB constant bank
    int32_t n_obj, n_mat, n_diel, n_box;   // n_mat counts the appended zero material (index n_mat-1); boxes are obj[0..n_box)
    // ---- the closest-hit scan's own tables (wavefront kernel).  Device order of the analytic objects:
    //   [0, n_box) boxes | [n_box, n_box + n_plane_run) planes | [.., + n_sphere_run) spheres | rest (generic loop),
    // where the plane run / sphere run / rest split the non-box objects WITHOUT reordering them (world order kept, so the
    // reference's "later object wins a tie" rule for spheres and planes survives).  scan_tab holds, 16-byte aligned:
    //   boxes   6 floats each (centre, half extent), padded to whole groups of kBoxGroup with never-hit boxes (h = -1);
    //   planes  1 float each (p.y), padded to a multiple of 4;
    //   spheres 4 floats each (centre, radius^2), padded to whole groups of kSphereGroup with never-hit spheres (r^2 = -1).
    int32_t n_box_groups, n_plane_run, n_sphere_run, n_sphere_groups;
    int32_t plane_off4, sphere_off4, n_typed, pad_;                       // offsets in float4 units; n_typed = first "rest" index
    float mesh_c[4], mesh_h[4];          // EXTENSION: bounds of all mesh triangles as (centre, half extent); h = -1 without meshes
    // dielectric exit search (renderer.go:316-371): when every dielectric object is a box or a sphere (and there are at
    // most kMaxExitTyped of each) they are listed again behind the scan records — boxes as 2 x float4 (centre | half extent),
    // spheres as 1 x float4 (centre, radius^2) — so the search runs branch-free loops without dependent index loads
    int32_t exit_typed, n_dbox, n_dsph, dbox_off4;
    int32_t dsph_off4, pad2_[3];
    alignas(16) float scan_tab[PTB_MAX_OBJECTS * 6 + 16 + kMaxExitTyped * 12];
    DevSky sky;
    DevCamera cam;
    int32_t diel_idx[PTB_MAX_OBJECTS];   // DEVICE indices of objects with a dielectric material, in ascending world order
    DevObj obj[PTB_MAX_OBJECTS];
    DevMat mat[PTB_MAX_MATERIALS + 1];   // slot n_mat-1 = the zero material (missing material_id, objects.go:234)
};
#ifndef PTB_BOX_GROUP
#define PTB_BOX_GROUP 4
#endif
constexpr int kBoxGroup = PTB_BOX_GROUP;
#ifndef PTB_SPHERE_GROUP
#define PTB_SPHERE_GROUP 2
#endif
constexpr int kSphereGroup = PTB_SPHERE_GROUP;

static_assert(sizeof(DevScene) <= 64 * 1024, "DevScene must fit the 64 KB constant bank");

// fp64 world for the primary-hit parity kernel (global memory; N is tiny).
struct Obj64 {
    int32_t type, pad;   // pad = world index (meshes interleave with analytic objects in world order)
    double a[3], b[3];   // sphere: a centre, b.x radius | plane: a point, b normal | box: a min, b max
};
struct Camera64 {
    double origin[3], llc[3], horizontal[3], vertical[3];
};

struct FrameParams {
    int32_t width, height;
    int32_t rows, row_offset, row_step;   // wavefront kernel: the launch renders `rows` rows (row_offset + k * row_step) into compact outputs
    int32_t s_begin, s_end;      // sample range of every pixel traced by this launch
    int32_t spp_total;           // divisor of the pixel epilogue
    int32_t max_depth;
    uint32_t seed_key;           // fmix(seed ^ GOLDEN), hoisted out of the kernel
    float inv_w, inv_h, h_minus_1;   // renderer.go:95-98
    const uint4* scene_blob;     // global copy of obj[0..n_obj) then mat[0..n_mat), 16-byte words (for the smem fill)
    float* accum;                // W*H*3 sums, or nullptr
    int32_t accum_resume;        // 1: start each pixel's sum from accum[] (progressive batches), 0: from zero
    uint8_t* rgba;               // W*H*4 finalised pixels, or nullptr
    unsigned long long* stats;   // kStatsWords counters, or nullptr
    unsigned int* work_counter;  // wavefront kernel: next unassigned work item (zeroed before the launch)
    // Small frames (fewer pixels than a few times the resident path slots): a work item is (pixel, 1/split_k of its sample
    // range) instead of a whole pixel, item w = plane * n_pix + pixel; partial sums go to planes[plane][pixel][3] and
    // finalize_planes_kernel adds the planes in order.  split_k == 1: one item per pixel, fused epilogue (large frames).
    int32_t split_k;
    float* planes;
    const float4* bvh_nodes;     // EXTENSION: BVH over the mesh triangles (bvh.h), nullptr when the scene has no mesh
    const float4* bvh_tris;
    int* trav_scratch;           // kTravStride ints per path slot of the launch (suspended traversals), mesh scenes only
};

enum StatWord {
    ST_SAMPLES = 0, ST_SEGMENTS, ST_EXIT_SCANS, ST_ACC_SPHERE, ST_ACC_PLANE, ST_ACC_BOX, ST_SCATTERS,
    ST_END_SKY, ST_END_EMISSIVE, ST_END_RR, ST_END_DEPTH, ST_END_NOSCATTER, ST_LANE_ACTIVE, ST_LANE_TOTAL,
    ST_ACC_MESH, ST_BVH_NODES, ST_BVH_TRIS,
    kStatsWords
};

// launchers (integrator.cu / primary_fp64.cu)
int upload_scene_constants(const DevScene& host_scene, void* stream);
int launch_integrator(const FrameParams& fp, bool stats, int n_obj, int n_mat, void* stream);
int launch_integrator_wf(const FrameParams& fp, bool stats, int n_obj, int n_mat, int sm_count, void* stream);
size_t wf_trav_scratch_bytes(int sm_count);      // size of FrameParams::trav_scratch the mesh instantiation needs
int launch_integrator_wq(const FrameParams& fp, bool stats, int n_obj, int n_mat, int sm_count, void* stream);
int launch_finalize(const float* accum, int width, int height, int spp_total, uint8_t* rgba, void* stream);
int launch_primary_hits(const Obj64* d_world, int n_obj, const float4* bvh_nodes, const float4* bvh_tris, const Camera64& cam,
                        int width, int height, double xi_u, double xi_v, int32_t* d_ids, double* d_t, void* stream);
int launch_finalize_peers(const float* const* d_bufs_on_dev0, int n_bufs, int width, int height, int spp_total, uint8_t* rgba, void* stream);
int launch_finalize_planes(const float* planes, int split_k, int width, int height, int spp_total, float* accum, int accum_resume, uint8_t* rgba, void* stream);
int wf_split_factor(int sm_count, long long n_pix, int n_samples);   // split_k the wavefront launcher wants for this frame
int launch_fma_peak(float* d_out, int blocks, int threads, int iters, void* stream);

}  // namespace ptb
