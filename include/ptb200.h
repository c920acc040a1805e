/*
 * ptb200.h — C ABI of libptb200.so, the B200 (sm_100a) CUDA backend for the per-pixel
 * render loop of MarkJulian19/path_trace_golang.
 *
 * This is the boundary the reference's Go code binds through cgo (INTEGRATION.md shows
 * the binding).  It replaces, for the CPU hot path
 *     engine.RenderInto(sc, cfg, img, progress)          internal/engine/renderer.go:34-41
 *       -> renderIntoCPU                                  internal/engine/renderer.go:44-246
 * exactly what the existing OpenGL plug-in replaces today
 *     gpu.Render(sc, cfg, img, progress) error            internal/engine/gpu/gpu.go:2534-2546
 * Plain pointers and sizes only; no C++/torch types.  All functions return PTB_OK (0) or a
 * negative PTB_ERR_* code; ptb_last_error() gives the message.  There is NO CPU fallback:
 * if CUDA is unavailable every call fails (the reference's GL path falls back to the CPU at
 * renderer.go:257-262; this backend deliberately does not).
 *
 * Threading: a context renders one frame at a time (calls on one context are serialised by
 * an internal mutex); any OS thread may call (cudaSetDevice is done per call), which is what
 * a goroutine-per-render caller such as internal/ui/app.go:135 needs.
 * Ownership: the library never keeps a host pointer after a call returns.
 * Determinism: a render is a pure function of (scene, cfg): same bytes run to run, on any number of streams, with or
 * without a progress callback.  Per pixel the samples are added in index order; frames with fewer than ~1.2 M pixels
 * add them in up to 64 sub-range sums that are then added in order (so does a multi-device render), i.e. the same
 * samples with fp32 sums re-associated.  ptb_scene_upload waits for the device before it replaces the scene.
 * Isolation: a launch carries its scene tables and camera BY VALUE in its kernel parameters, so any number of
 * contexts and streams may render on the same device concurrently, and a scene uploaded (or a frame of another size
 * queued) afterwards cannot disturb a frame already in flight.  The host-buffer calls (ptb_render, ptb_render_accum,
 * ptb_render_resume, ptb_primary_hits) synchronise before they return.
 */
#ifndef PTB200_H
#define PTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTB_ABI_VERSION 5   /* 2: meshes (ptb_scene.n_mesh ...), 3: ptb_multi_*, 4: ptb_cfg.row_offset/row_step,
                               5: ptb_scene.mesh_generation, ptb_render_resume, ptb_host_buffer_pin, ptb_stats.bvh_stack_overflows,
                                  no constant-bank limit on the number of objects */

enum {
    PTB_OK = 0,
    PTB_ERR_INVALID = -1,  /* bad argument (NULL, non-positive size, unknown code, ...) */
    PTB_ERR_CUDA = -2,     /* CUDA runtime/driver error, message has the cudaError string */
    PTB_ERR_NO_SCENE = -3, /* render called before ptb_scene_upload */
    PTB_ERR_LIMIT = -4     /* scene exceeds a compiled-in limit (PTB_MAX_OBJECTS, ...) */
};

/* Object type codes — same values as the GL plug-in's OBJ_* (gpu.go:244-248).
 * "sphere_light" is a sphere (objects.go:246-252).  Any other code: the object is dropped,
 * like unknown types in sceneToWorld (objects.go:237-266). */
enum { PTB_OBJ_SPHERE = 0, PTB_OBJ_PLANE = 1, PTB_OBJ_BOX = 2,
       PTB_OBJ_MESH = 3 /* EXTENSION (not in the reference): triangle mesh, see ptb_scene.n_mesh */ };

/* Material type codes — materialType iota (materials.go:11-17) == MAT_* (gpu.go:236-242).
 * Any scene type string other than metal/dielectric/emissive/mirror maps to LAMBERT
 * (default branch, materials.go:50-53). */
enum { PTB_MAT_LAMBERT = 0, PTB_MAT_METAL = 1, PTB_MAT_DIELECTRIC = 2, PTB_MAT_EMISSIVE = 3, PTB_MAT_MIRROR = 4 };

/* Limits of the analytic world (sceneToWorld itself has none, objects.go:225-269; the shipped scenes have <= 44 objects
 * and 27 materials).  Worlds whose scan tables fit the kernel-parameter block (about 300 objects) are scanned from the
 * constant bank; larger ones from global memory by the same code — a linear scan like the reference's, every object
 * tested on every ray (renderer.go:297-302). */
#define PTB_MAX_OBJECTS 1000000
#define PTB_MAX_MATERIALS 1000000

/* scene.Camera (internal/scene/scene.go:23-32), raw fields; newCamera (camera.go:19-58)
 * is evaluated inside the library in binary64 for the frame size of each render. */
typedef struct {
    double position[3], target[3], up[3];
    double fov;          /* degrees, vertical */
    double aperture;
    double focus_dist;   /* 0 = |position - target| */
    double aspect_ratio; /* 0 = width/height */
} ptb_camera;

/* Sky selection of renderIntoCPU (renderer.go:56-92), resolved by the caller:
 * kind 1 (gradient) uses horizon/zenith; kind 0 uses color (= sky.color for a "solid" sky,
 * else scene.background). */
enum { PTB_SKY_CONST = 0, PTB_SKY_GRADIENT = 1 };
typedef struct {
    int32_t kind;
    double color[3], horizon[3], zenith[3];
} ptb_sky;

/* Flattened scene, SoA, RAW scene.go fields: the library applies convertMaterial
 * (materials.go:28-55) and the sceneToWorld geometry (objects.go:225-269: sphere radius =
 * size.x, plane normal (0,1,0), box min/max = position -/+ size*0.5) itself, in binary64.
 * obj_mat[i] is the index of the material whose id equals the object's material_id, where
 * a later duplicate id wins (map semantics, objects.go:226-229); -1 = id not found = the
 * zero material (black lambert). World index = index among kept objects, in array order. */
typedef struct {
    int32_t n_obj;
    const int32_t* obj_type; /* [n_obj] PTB_OBJ_* (other values: dropped) */
    const int32_t* obj_mat;  /* [n_obj] index into mat_* or -1 */
    const double* obj_pos;   /* [n_obj*3] scene.Object.Position */
    const double* obj_size;  /* [n_obj*3] scene.Object.Size */
    int32_t n_mat;
    const int32_t* mat_type;       /* [n_mat] PTB_MAT_* */
    const double* mat_albedo;      /* [n_mat*3] */
    const double* mat_rough;       /* [n_mat] */
    const double* mat_ior;         /* [n_mat] */
    const double* mat_emit;        /* [n_mat*3] (before * power) */
    const double* mat_power;       /* [n_mat] */
    const double* mat_absorption;  /* [n_mat*3] */
    const double* mat_smoothness;  /* [n_mat] */
    ptb_camera camera;
    ptb_sky sky;
    /* EXTENSION — triangle meshes (the reference has none; north-star: "internal/scene gains a BVH builder that emits
     * a flattened, cache-line-aligned node array").  An object of type PTB_OBJ_MESH is ONE world entry; obj_mesh[i]
     * names its mesh; triangles are given in world space, binary32, 9 floats each (v0, v1, v2).  The library builds
     * the BVH (binned SAH collapsed to 4-wide 128-byte nodes).  Hit rule: Moeller-Trumbore, two-sided, t in [tMin, closest); among
     * triangles of equal t the lowest triangle index wins; geometric normal.  n_mesh == 0: all three may be NULL. */
    int32_t n_mesh;
    const int32_t* obj_mesh;         /* [n_obj] mesh index for PTB_OBJ_MESH objects, -1 otherwise */
    const int64_t* mesh_tri_begin;   /* [n_mesh+1] triangle range of mesh k = [begin[k], begin[k+1]) */
    const float* tri_vertices;       /* [9 * mesh_tri_begin[n_mesh]] */
    /* Caller-held identity of the triangle data: RenderInto hands the scene over on every call (renderer.go:34), and a
     * 10 M-triangle mesh is 360 MB.  Non-zero: the caller's promise that two uploads with the same mesh_generation carry
     * identical mesh_tri_begin / tri_vertices contents — the library then reuses the BVH it built for that generation
     * without reading the triangles at all.  0: no promise; the triangles are hashed (all host cores) to recognise a
     * BVH built before.  A host bumps the generation whenever it edits a mesh. */
    uint64_t mesh_generation;
} ptb_scene;

/* engine.RenderConfig (renderer.go:17-22) + the extra knobs a GPU backend needs. */
typedef struct {
    int32_t width, height, samples_per_px, max_depth;
    uint32_t seed;          /* key of the counter RNG; the reference seeds from the clock (random.go:14) */
    int32_t sample_begin;   /* this context traces samples [sample_begin, sample_begin+sample_count) of    */
    int32_t sample_count;   /* every pixel; sample_count <= 0 means all of [0, samples_per_px)            */
    uint32_t flags;         /* PTB_FLAG_* */
    /* Row partition (multi-GPU by image tiles instead of sample ranges): row_step > 1 renders only the rows
     * row_offset, row_offset + row_step, ... of the frame.  Every output of such a call is COMPACT: it holds just
     * those rows, ptb_rows_of(cfg) of them, in order.  Pixels are computed exactly as in a whole-frame render (same
     * camera, same RNG keys), so interleaving the ranks' rows gives the single-device image bit for bit.
     * row_step <= 1: all rows (row_offset must then be 0). */
    int32_t row_offset, row_step;
} ptb_cfg;

#define PTB_FLAG_STATS 1u       /* run the counting variant of the integrator (slower); fills ptb_stats.  Ignored by ptb_render
                                   when a progress callback is given (the batches are not counted) */
#define PTB_FLAG_MEGAKERNEL 2u  /* use the pixel-per-lane megakernel instead of the default wavefront-in-shared-memory
                                   kernel (same results up to fp32 rounding; kept for A/B profiling) */

typedef struct ptb_ctx ptb_ctx;

/* progress callback, same role as `progress func()` of RenderInto (renderer.go:34):
 * invoked on the calling thread after rgba has been refreshed with the samples so far.  The context is locked
 * while it runs: it must not call back into the library with the same context. */
typedef void (*ptb_progress_fn)(void* user);

typedef struct {
    uint64_t samples;     /* camera samples traced */
    uint64_t segments;    /* closest-hit scans (renderer.go:297) */
    uint64_t exit_scans;  /* dielectric exit searches (renderer.go:329) */
    uint64_t accepts[3];  /* winning hits per primitive type */
    uint64_t scatters;    /* scatter() == ok */
    uint64_t end_sky, end_emissive, end_rr, end_depth, end_noscatter;
    uint64_t lane_iters_active; /* integrator loop iterations with the lane alive           */
    uint64_t lane_iters_total;  /* 32 x warp loop iterations (SIMT utilisation denominator) */
    double last_render_ms;      /* device time of the last render call (CUDA events) */
    uint64_t accepts_mesh;      /* EXTENSION: winning hits on mesh triangles */
    uint64_t bvh_nodes_visited; /* 128-byte (4-wide) node fetches */
    uint64_t bvh_tris_tested;   /* 64-byte triangle fetches */
    uint64_t bvh_stack_overflows; /* subtrees dropped because the traversal stack was full: must be 0 (the builder bounds the depth) */
} ptb_stats;

/* EXTENSION: the BVH built by ptb_scene_upload over the mesh triangles (all zero when the scene has no mesh). */
typedef struct {
    int64_t n_triangles, n_nodes;
    int32_t max_depth, node_bytes, triangle_bytes, reserved;
    double sah_cost, build_ms;
} ptb_bvh_info;

typedef struct {
    char name[128];
    int32_t sm_count, cc_major, cc_minor, clock_khz;
    uint64_t global_mem_bytes;
} ptb_device_info;

/* Create a context on one CUDA device (one context per GPU; multi-GPU = one context per
 * process/rank).  Fails with PTB_ERR_CUDA when no usable GPU exists. */
int ptb_create(int device, ptb_ctx** out);
void ptb_destroy(ptb_ctx* ctx);
/* Message of the last failure on ctx (ctx may be NULL: last ptb_create failure of this thread). */
const char* ptb_last_error(const ptb_ctx* ctx);
int ptb_abi_version(void);
int ptb_get_device_info(ptb_ctx* ctx, ptb_device_info* out);

/* Copies and converts the scene (sceneToWorld + convertMaterial).  May be called again to
 * replace the scene. */
int ptb_scene_upload(ptb_ctx* ctx, const ptb_scene* scene);
/* Converted world, for flatten parity tests: number of kept objects, and entry i as 19 doubles
 * {type, mat_type, a[3], b[3], albedo[3], rough, ior, emit[3], absorption[3]}
 * (a/b: sphere centre/(radius,0,0); plane point/normal; box min/max). */
int ptb_world_size(ptb_ctx* ctx);
int ptb_world_get(ptb_ctx* ctx, int i, double out[19]);

/* The whole of renderIntoCPU (renderer.go:44-246) into a HOST image: rgba is height rows of
 * `stride` bytes (stride >= 4*width), top row first, R,G,B,A=255 per pixel — the layout of
 * image.RGBA.Pix/.Stride the caller of RenderInto owns.  progress may be NULL. */
int ptb_rows_of(const ptb_cfg* cfg);   /* rows a call with this cfg outputs: height, or ceil((height - row_offset) / row_step) */
int ptb_render(ptb_ctx* ctx, const ptb_cfg* cfg, uint8_t* rgba, size_t stride,
               ptb_progress_fn progress, void* user);

/* Page-lock a caller-owned host buffer (cudaHostRegister) so that ptb_render / ptb_multi_render / ptb_render_resume copy
 * the image straight into it instead of through the context's staging buffer (saves one 33 MB memcpy per 4K frame).
 * Entirely optional and entirely the caller's: the library keeps no record of the buffer; unpin it before freeing it.
 * A host that re-renders into one long-lived image (internal/ui/app.go:157-168 reallocates only on a size change) pins
 * it once. */
int ptb_host_buffer_pin(ptb_ctx* ctx, void* p, size_t bytes);
int ptb_host_buffer_unpin(ptb_ctx* ctx, void* p);

/* Linear fp32 radiance sums (NOT divided by the sample count) of this context's sample range,
 * width*height*3 floats, top row first — to a host buffer ... */
int ptb_render_accum(ptb_ctx* ctx, const ptb_cfg* cfg, float* rgb_sum);
/* ... or asynchronously into DEVICE memory on a caller-supplied CUDA stream (cudaStream_t passed as
 * void*, used exactly as given: NULL is the CUDA default stream, which is also PyTorch's default
 * stream).  This is what a multi-GPU caller hands to the exchange (ptb_peer_* below, or NCCL).  (With the environment
 * variable PTB_MESH_PIPELINE set, scenes with meshes are rendered by a multi-kernel pipeline driven from the host: the call
 * then returns when the frame is complete.) */
int ptb_render_accum_device(ptb_ctx* ctx, const ptb_cfg* cfg, void* d_rgb_sum, void* stream);
/* Checkpointable rendering (the reference persists only scenes and PNGs, util.go:45-55; a long render on a GPU backend
 * wants to survive a restart): renders samples [sample_begin, sample_begin + sample_count) of cfg CONTINUING the sums
 * in rgb_sum (host, width*height*3 floats): on entry the sums of samples [0, sample_begin) (ignored when sample_begin
 * is 0), on return those of [0, sample_begin + sample_count).  rgba (may be NULL) receives the image of the mean over
 * the samples so far.  Per pixel the samples are added strictly in index order, so ANY partition of [0, spp) into
 * consecutive calls — with the sums written to disk and read back in between — gives the same bits as one call. */
int ptb_render_resume(ptb_ctx* ctx, const ptb_cfg* cfg, float* rgb_sum, uint8_t* rgba, size_t stride);

/* Pixel epilogue (renderer.go:189-221) on device buffers: mean over spp_total, sqrt, *255.999,
 * clamp, truncate, A=255.  d_rgba: height*width*4 bytes, tightly packed. */
int ptb_finalize_device(ptb_ctx* ctx, const void* d_rgb_sum, int32_t width, int32_t height,
                        int32_t spp_total, void* d_rgba, void* stream);
/* The same epilogue for HOST buffers: rgb_sum (width*height*3 floats) is sent to the device, finalised there, and the image
 * comes back into rgba (stride >= 4*width).  Used to turn stored sums (a finished checkpoint) into an image. */
int ptb_finalize_host(ptb_ctx* ctx, const float* rgb_sum, int32_t width, int32_t height, int32_t spp_total, uint8_t* rgba, size_t stride);
/* Fused render + epilogue into a DEVICE RGBA8 image (no host copies; stream as above). */
int ptb_render_device(ptb_ctx* ctx, const ptb_cfg* cfg, void* d_rgba, void* stream);

/* Parity hook: closest hit of the primary ray through (x+xi_u, y+xi_v) of every pixel with the
 * lens disabled (camera.go:70-73), computed in binary64 with the reference's operation order
 * and no FMA contraction.  ids: world index or -1; t: ray parameter (0 on miss). Host buffers. */
int ptb_primary_hits(ptb_ctx* ctx, const ptb_cfg* cfg, double xi_u, double xi_v,
                     int32_t* ids, double* t);

/* Counters of the last render that ran with PTB_FLAG_STATS (last_render_ms is always valid). */
int ptb_get_stats(ptb_ctx* ctx, ptb_stats* out);
int ptb_get_bvh_info(ptb_ctx* ctx, ptb_bvh_info* out);
/* Symbols of the kernels the last render of ctx launched for the frame, as ncu prints them, joined by " + ": the integrator
 * (e.g. "integrate_wf_kernel<0, 0, 0, 1>": counting build?, mesh traversal?, global-memory tables?, packed two-ray sphere
 * test?) and, when the frame was rendered as (pixel, sample sub-range) work items, "finalize_planes_kernel".  Valid until the
 * next render on ctx. */
const char* ptb_last_kernel(ptb_ctx* ctx);

/* ---- single-process multi-GPU (what a Go host that owns every GPU of the box calls; one process per GPU + NCCL is the
 * other supported arrangement).  Device k traces samples [k*spp/n, (k+1)*spp/n) of every pixel
 * into its own fp32 buffer; device 0 then runs ONE kernel that reads all n buffers — its own and, through NVLink
 * peer access, the others' — sums them and applies the pixel epilogue (reduce + finalise fused, no staging copy).
 * Same image as one device up to fp32 re-association of the n partial sums. */
typedef struct ptb_multi ptb_multi;
int ptb_multi_create(const int* devices, int n_devices, ptb_multi** out);   /* devices == NULL: devices 0..n-1 */
void ptb_multi_destroy(ptb_multi* m);
const char* ptb_multi_last_error(const ptb_multi* m);                       /* m may be NULL (failed create) */
int ptb_multi_scene_upload(ptb_multi* m, const ptb_scene* scene);
int ptb_multi_render(ptb_multi* m, const ptb_cfg* cfg, uint8_t* rgba, size_t stride);
/* device time of the last ptb_multi_render: slowest device's integrator, and the fused reduce+epilogue on device 0 */
int ptb_multi_last_timing(ptb_multi* m, double* render_ms, double* reduce_ms);

/* ---- multi-process multi-GPU: one process (rank) per GPU, as under torchrun / MPI.  The ranks' fp32 sum buffers and rank 0's
 * image are shared through CUDA IPC, and ONE kernel per rank does reduce-scatter + pixel epilogue + gather: rank r sums its
 * 1/world slice of the pixels over all ranks' buffers (NVLink peer loads) and stores the finalised RGBA8 pixels directly into
 * rank 0's image (NVLink peer stores).  Against "NCCL reduce to rank 0, then finalise" this moves 12 B/pixel * (world-1)/world
 * spread over all links plus 4 B/pixel into rank 0, instead of 12 B/pixel * (world-1) into rank 0 alone.
 * The cross-rank ordering is part of the same kernel — flag words in peer memory: "my sums of frame seq are complete" /
 * "my slice of frame seq is written" — so no collective-library call is made per frame at all.
 * Protocol: every rank ptb_peer_create -> ptb_peer_handles -> exchange the handles with the host's own communicator
 * (all-gather of 3 x 64 bytes, once) -> ptb_peer_connect.  Per frame, on every rank, on one stream: render into
 * ptb_peer_accum() (ptb_render_accum_device), then ptb_peer_reduce_finalize — when it has run on the stream, every rank has
 * finished the frame: rank 0 owns the image at ptb_peer_image() and every rank may render the next frame into its buffer.
 * Every rank must call ptb_peer_reduce_finalize the same number of times (frames are matched by a sequence number); a rank
 * that never arrives makes the others give up after a few seconds (ptb_peer_status reports it) instead of hanging the device.
 * Sum order is rank 0, 1, 2, ... on every rank: the image does not depend on which rank finalises which slice. */
#define PTB_IPC_HANDLE_BYTES 64
typedef struct ptb_peer ptb_peer;
int ptb_peer_create(ptb_ctx* ctx, int rank, int world, int32_t max_width, int32_t max_height, ptb_peer** out);
void ptb_peer_destroy(ptb_peer* p);
int ptb_peer_handles(ptb_peer* p, unsigned char accum_handle[PTB_IPC_HANDLE_BYTES], unsigned char image_handle[PTB_IPC_HANDLE_BYTES],
                     unsigned char flags_handle[PTB_IPC_HANDLE_BYTES]);
int ptb_peer_connect(ptb_peer* p, const unsigned char* accum_handles /* world x 64 bytes, in rank order */,
                     const unsigned char* root_image_handle /* rank 0's image handle */,
                     const unsigned char* flags_handles /* world x 64 bytes, in rank order */);
int ptb_peer_status(ptb_peer* p);                  /* PTB_OK, or PTB_ERR_CUDA when a wait of the exchange timed out */
void* ptb_peer_accum(ptb_peer* p);                 /* device pointer: this rank's width*height*3 float sums of the NEXT frame (two buffers
                                                      alternate, so ranks may run a frame apart): ask again before every frame */
void* ptb_peer_image(ptb_peer* p);                 /* device pointer on rank 0 (NULL elsewhere): width*height*4 bytes */
int ptb_peer_slice(const ptb_peer* p, int32_t width, int32_t height, int64_t* begin, int64_t* end);   /* pixel range this rank finalises */
int ptb_peer_reduce_finalize(ptb_peer* p, int32_t width, int32_t height, int32_t spp_total, void* stream);

/* Test hook, host only (no CUDA call): the device layout ptb_scene_upload would give the scene's analytic objects.
 * order[k] = world index of device object k (at most cap entries are written); counts = {boxes, planes in the typed
 * plane run, spheres in the typed sphere run, objects left to the generic loop, dielectric boxes and dielectric
 * spheres of the typed exit search (both 0 when the search is not typed)}.  Returns the number of analytic objects
 * or a negative PTB_ERR_* (message: ptb_last_error(NULL)). */
int ptb_scene_device_order(const ptb_scene* scene, int32_t* order, int32_t cap, int32_t counts[6]);

/* Test hook, host only (no CUDA call): the launch plan the wavefront integrator would use on a device with sm_count SMs.
 * Returns k = the number of sample planes a frame of n_pixels pixels and n_samples samples per pixel is rendered in (work items
 * are (pixel, 1/k of its sample range); 1 = whole pixels), chosen for about 40 work items per resident path slot; if packed is
 * not NULL, *packed = 1 when a scene with n_spheres spheres in its typed sphere run gets the kernel instantiation with the packed
 * two-ray sphere test, else 0.  Negative PTB_ERR_* on bad arguments. */
int ptb_launch_plan(int32_t sm_count, int64_t n_pixels, int32_t n_samples, int32_t n_spheres, int32_t* packed);

/* Test hook, host only (no CUDA call): builds the BVH over n_tri world-space triangles (9 floats each) exactly as
 * ptb_scene_upload does and checks the emitted node array: every child box (centre/half extent in binary32) contains all
 * the triangles below it, every triangle sits in exactly one leaf, links and counts are consistent.
 * Returns the number of violations (0 = sound) or a negative PTB_ERR_*; n_nodes / max_depth may be NULL. */
int64_t ptb_bvh_selfcheck(const float* tri_vertices, int64_t n_tri, int64_t* n_nodes, int32_t* max_depth);

/* Host only (no CUDA call): the BVH builder on its own — binned SAH over n_tri world-space triangles (9 floats each), output
 * as the flattened arrays the device traverses: 128-byte cache-line-aligned 4-wide nodes (the children's boxes as centre / half
 * extent + links; layout in path_trace_golang_b200/csrc/bvh.h) and 64-byte triangles (v0, e1, e2, pad) in leaf order carrying the
 * original triangle index.  ptb_scene_upload calls the same builder; this entry point lets the host inspect, cache or
 * serialise the structure (north-star: "internal/scene gains a BVH builder that emits a flattened, cache-line-aligned node
 * array").  Free with ptb_bvh_free. */
typedef struct {
    const float* nodes;       /* info.n_nodes x 32 floats (128-byte nodes), 64-byte aligned */
    const float* triangles;   /* info.n_triangles x 16 floats */
    ptb_bvh_info info;
} ptb_bvh;
int ptb_bvh_build(const float* tri_vertices, int64_t n_tri, ptb_bvh* out);
void ptb_bvh_free(ptb_bvh* b);

/* Measurement helper: FP32 FMA throughput of the device (2 flop per FMA), the roofline
 * denominator MEASURED_PEAKS.json does not carry. */
int ptb_measure_fp32_peak(ptb_ctx* ctx, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* PTB200_H */
