// scene_dev.h — device-side scene layout shared by the host API (api.cu) and the kernels.
//
// The analytic world of the reference (objects.go:26-222: spheres, one-normal planes, boxes; a few dozen objects in the
// shipped scenes) is tiny, so it is not streamed from HBM at all.  The tables the closest-hit scan reads travel with
// every launch as a `__grid_constant__` KERNEL PARAMETER (SceneK, constant bank 0): the scan indexes them with a
// warp-uniform index, which the uniform datapath serves as LDCU/ULDC broadcasts — and because the copy belongs to the
// launch, two contexts (or two queued frames) of one device can never see each other's scene.  The few divergent
// look-ups (the winning object, its material) go through a shared-memory copy made per CTA.
// Worlds too large for the parameter table (more than ~300 objects) take the BIG instantiation: same code, tables and
// object / material records read from per-context global memory.
#pragma once
#include <stdint.h>
#include <vector_types.h>

#include <vector>

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

#ifndef PTB_BOX_GROUP
#define PTB_BOX_GROUP 4
#endif
constexpr int kBoxGroup = PTB_BOX_GROUP;
#ifndef PTB_SPHERE_GROUP
#define PTB_SPHERE_GROUP 2
#endif
constexpr int kSphereGroup = PTB_SPHERE_GROUP;
constexpr int kMaxExitTyped = 64;
// Scenes with at least this many spheres in the typed sphere run get the kernel instantiation that tests a thread's two rays
// against a sphere in packed arithmetic (integrator.cu "two rays per instruction"): below it, building the pairs costs more than it saves.
constexpr int kPackedSphereMin = 6;

// Scan tables that fit travel as kernel parameters; larger worlds use the BIG instantiation (tables in global memory).
constexpr int kTabFloats = 1920;                 // 7.5 KB: e.g. 256 boxes + 64 spheres + the typed exit-search copies
constexpr int kSmallBlobBytes = 12 * 1024;       // object + material records a CTA copies to shared memory (else: BIG)

// The scene as one launch sees it.  Device order of the analytic objects:
//   [0, n_box) boxes | [n_box, n_box + n_plane_run) planes | [.., + n_sphere_run) spheres | rest (generic loop),
// where the plane run / sphere run / rest split the non-box objects WITHOUT reordering them (world order kept, so the
// reference's "later object wins a tie" rule for spheres and planes survives).  The scan table holds, 16-byte aligned:
//   boxes   6 floats each (centre, half extent), padded to whole groups of kBoxGroup with never-hit boxes (h = -1);
//   planes  1 float each (p.y), padded to a multiple of 4;
//   spheres 4 floats each (centre, radius^2), padded to whole groups of kSphereGroup with never-hit spheres (r^2 = -1);
//   dielectric exit search (renderer.go:316-371): when every dielectric object is a box or a sphere (at most
//   kMaxExitTyped of each) they are listed again — boxes as 2 x float4 (centre | half extent), spheres as 1 x float4
//   (centre, radius^2) — so the search runs branch-free loops without dependent index loads.
struct SceneHdr {
    int32_t n_obj, n_mat, n_diel, n_box;   // n_mat counts the appended zero material (index n_mat-1); boxes are obj[0..n_box)
    int32_t n_box_groups, n_plane_run, n_sphere_run, n_sphere_groups;
    int32_t plane_off4, sphere_off4, n_typed, exit_typed;                 // offsets in float4 units; n_typed = first "rest" index
    int32_t n_dbox, n_dsph, dbox_off4, dsph_off4;
    float mesh_c[4], mesh_h[4];          // EXTENSION: bounds of all mesh triangles as (centre, half extent); h = -1 without meshes
    DevSky sky;
    DevCamera cam;
    int32_t tab_floats, pad_;            // used length of the scan table
    const float4* tab_global;            // BIG: the scan table in global memory (same layout)
    const int32_t* diel_idx;             // global: DEVICE indices of the dielectric objects in ascending world order (untyped exit search)
};
struct SceneK : SceneHdr {
    alignas(16) float scan_tab[kTabFloats];
};

// Host-side staging of one uploaded scene (pageable memory; launches copy what they need BY VALUE).
struct HostScene {
    SceneHdr hdr{};
    std::vector<float> scan_tab;         // hdr.tab_floats entries
    std::vector<int32_t> diel_idx;
    std::vector<DevObj> obj;
    std::vector<DevMat> mat;
    bool big = false;                    // tables / records do not fit the parameter + shared-memory budget
};

// fp64 world for the primary-hit parity kernel (global memory).
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
    const uint4* scene_blob;     // global copy of obj[0..n_obj) then mat[0..n_mat), 16-byte words (smem fill; read in place when BIG)
    float* accum;                // W*H*3 sums, or nullptr
    int32_t accum_resume;        // 1: start each pixel's sum from accum[] (progressive batches), 0: from zero
    uint8_t* rgba;               // W*H*4 finalised pixels, or nullptr
    unsigned long long* stats;   // kStatsWords counters, or nullptr
    unsigned int* work_counter;  // wavefront kernel: next unassigned work item (zeroed before the launch)
    // Small frames (fewer pixels than a few times the resident path slots): a work item is (pixel, one sample sub-range)
    // instead of a whole pixel, item w = plane * n_pix + pixel; partial sums go to planes[plane][pixel][3] and
    // finalize_planes_kernel adds the planes in order.  split_k == 1: one item per pixel, fused epilogue (large frames).
    // The sub-ranges are those of the WHOLE render: plane p of this launch is global plane split_base + p of split_total,
    // samples [full_begin + n * (split_base + p) / split_total, full_begin + n * (split_base + p + 1) / split_total) with
    // n = full_end - full_begin — so a progressive render (several launches) adds exactly the sums one launch adds.
    int32_t split_k, split_base, split_total, full_begin, full_end;
    float* planes;
    const float4* bvh_nodes;     // EXTENSION: BVH over the mesh triangles (bvh.h), nullptr when the scene has no mesh
    const float4* bvh_tris;
    int* trav_scratch;           // kTravStride ints per path slot of the launch (suspended traversals), mesh scenes only
};

struct KernelArgs {              // the one __grid_constant__ parameter of the integrator kernels
    FrameParams fp;
    SceneK sc;
};
static_assert(sizeof(KernelArgs) <= 16 * 1024, "kernel parameters: keep well below the 32,764-byte limit");

// Mesh scenes (mesh_pipeline.cuh): path state in global memory between the kernels of the pipeline, the ray queue of the
// traversal kernel and its control words.  One pool per context, sized for SMs x PTB_MP_CTAS_PER_SM x 512 path slots.
struct MeshPool {
    float4 *O, *D, *B, *A;       // per slot, as SlotState (wavefront.cuh)
    float* best; int* bid; unsigned short* dep;
    int* queue;                  // slots whose ray must be traversed this iteration
    unsigned int* ctl;           // control words (mesh_pipeline.cuh)
    int n_slots;
};
struct MeshPipe {                // host side of the pool
    MeshPool pool{};
    void* d_mem = nullptr; size_t d_bytes = 0;
    unsigned int* h_flags = nullptr;     // pinned: copies of the "alive" flags
    void* events[4] = {nullptr, nullptr, nullptr, nullptr};
    int traverse_blocks = 0;             // persistent grid of the traversal kernel (per device)
};
size_t mesh_pool_bytes(int n_slots);
void mesh_pool_bind(MeshPool& pool, void* d_mem, int n_slots);
int mesh_pool_slots(int sm_count, long long n_items);

enum StatWord {
    ST_SAMPLES = 0, ST_SEGMENTS, ST_EXIT_SCANS, ST_ACC_SPHERE, ST_ACC_PLANE, ST_ACC_BOX, ST_SCATTERS,
    ST_END_SKY, ST_END_EMISSIVE, ST_END_RR, ST_END_DEPTH, ST_END_NOSCATTER, ST_LANE_ACTIVE, ST_LANE_TOTAL,
    ST_ACC_MESH, ST_BVH_NODES, ST_BVH_TRIS, ST_BVH_STACK_OVERFLOW,
    kStatsWords
};

// Per-context cache of what a launcher learns about one kernel instantiation ON THIS DEVICE: the dynamic shared memory
// it has opted into and the resident CTAs per SM for the last shared-memory size asked for.
struct LaunchCache {
    struct Entry { size_t smem_optin = 0, smem_occ = ~(size_t)0; int blocks_per_sm = 0; };
    Entry wf[16];                // integrate_wf_kernel<STATS, MESH, BIG, PK>
    Entry mp[4];                 // mp_shade_scan_kernel<STATS, BIG>
};

// launchers (integrator.cu / primary_fp64.cu)
int launch_integrator(const KernelArgs& ka, bool stats, void* stream);          // pixel-per-lane megakernel (small worlds only)
int launch_integrator_wf(const KernelArgs& ka, bool stats, bool big, bool packed, int sm_count, LaunchCache* cache, void* stream);
// mesh scenes: the wavefront across kernels.  Launches iteration after iteration on `stream` and WAITS for completion in
// steps (the host reads an "alive" flag a few iterations behind): returns when the frame is complete.
int launch_mesh_pipeline(const KernelArgs& ka, bool stats, bool big, int sm_count, LaunchCache* cache, MeshPipe& mp, void* stream);
int launch_clear_frame(float* accum, int accum_resume, uint8_t* rgba, int n_pix, void* stream);   // max_depth <= 0: black frame
size_t wf_trav_scratch_bytes(int sm_count);      // size of FrameParams::trav_scratch the mesh instantiation needs
int launch_finalize(const float* accum, int width, int height, int spp_total, uint8_t* rgba, void* stream);
int launch_primary_hits(const Obj64* d_world, int n_obj, const float4* bvh_nodes, const float4* bvh_tris, const Camera64& cam,
                        int width, int height, double xi_u, double xi_v, int32_t* d_ids, double* d_t, void* stream);
int launch_finalize_peers(const float* const* d_bufs_on_dev0, int n_bufs, int width, int height, int spp_total, uint8_t* rgba, void* stream);
// Cross-rank flags of the multi-process exchange (ptb_peer_*): flags_of[k] = this process's mapping of rank k's flag block,
// 2 x world words: [j] "rank j's sums of frame seq are complete", [world + j] "rank j has written its slice of frame seq".
constexpr int kMaxPeers = 16;
struct PeerSync {
    unsigned* flags_of[kMaxPeers];
    unsigned* block_counter;     // local: CTAs of the slice kernel that have finished
    unsigned* error;             // local: set when a wait timed out
    int rank, world;
    unsigned seq;                // frame sequence number (1, 2, ...)
};
int launch_reduce_finalize_slice(const float* const* d_bufs, size_t buf_off, int n_bufs, long long p_begin, long long p_end, int spp_total, uint8_t* rgba_root,
                                 const PeerSync& sync, void* stream);
int launch_finalize_planes(const float* planes, int split_k, int width, int height, int spp_total, float* accum, int accum_resume, uint8_t* rgba, void* stream);
int wf_split_factor(int sm_count, long long n_pix, int n_samples);   // split_k the wavefront launcher wants for this frame
int launch_fma_peak(float* d_out, int blocks, int threads, int iters, void* stream);

}  // namespace ptb
