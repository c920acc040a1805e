/*
 * oracle.h — C ABI of the CPU oracle (TEST INFRASTRUCTURE, NOT THE PRODUCT).
 *
 * The oracle restates the CPU hot path of MarkJulian19/path_trace_golang
 * (internal/engine/renderer.go:44-246, 286-404; camera.go; objects.go;
 * materials.go; math.go) in C++ so that the CUDA path can be checked against
 * it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product (libptb200.so) never does.
 *
 * PARITY UNPINNED: the reference ships no tests, no golden vectors and no
 * stored renders, and its Go toolchain is absent here, so this restatement
 * cannot be pinned against reference outputs.  It is pinned against
 * hand-derived known-answer vectors (tests/golden/, tests/test_oracle_kat.py)
 * and against a second, independent restatement of the same Go sources in
 * plain Python (oracle/goref.py): tests/test_oracle_crosscheck.py demands
 * bit-for-bit agreement of worlds, cameras, primary hits, per-pixel radiance
 * sums and event counters on the five shipped scenes and on random scenes.
 * go/cmd/gengolden produces the reference-side vectors on a box with Go.
 */
#ifndef PTB_ORACLE_H
#define PTB_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Raw scene fields, i.e. internal/scene/scene.go BEFORE sceneToWorld /
 * convertMaterial are applied (the oracle does that conversion itself so the
 * product's flattener can be checked against it). */
typedef struct {
    const char* id;       /* scene.go:47 */
    const char* type;     /* scene.go:48  "lambert"|"metal"|"dielectric"|"emissive"|"mirror"|other */
    double albedo[3];     /* scene.go:50 */
    double rough;         /* scene.go:51 */
    double ior;           /* scene.go:52 */
    double emit[3];       /* scene.go:54 */
    double power;         /* scene.go:55 */
    double absorption[3]; /* scene.go:59 */
    double smoothness;    /* scene.go:62 */
} orc_raw_material;

typedef struct {
    const char* type;        /* scene.go:83 "sphere"|"plane"|"box"|"sphere_light"|other (dropped) */
    double position[3];      /* scene.go:85 */
    double size[3];          /* scene.go:86 */
    const char* material_id; /* scene.go:88 */
    /* EXTENSION (type "mesh", not in the reference): world-space triangles of this object, binary32, 9 floats each.
     * Hit rule (DESIGN.md "Meshes"): Moeller-Trumbore, two-sided, geometric normal; meshes are tested after all
     * analytic objects and win only with t < closest; among triangles of equal t the lowest triangle id wins
     * (ids count over all meshes in world order). */
    const float* tri_vertices;
    int64_t n_tri;
} orc_raw_object;

typedef struct {
    double position[3], target[3], up[3]; /* scene.go:24-26 */
    double fov, aperture, focus_dist, aspect_ratio; /* scene.go:27-31 */
} orc_raw_camera;

typedef struct {
    int has_sky;           /* sc.Sky != nil (scene.go:154) */
    const char* sky_type;  /* scene.go:139 */
    double color[3], horizon[3], zenith[3]; /* scene.go:140-142 */
    double background[3];  /* scene.go:153 */
} orc_raw_sky;

typedef struct orc_scene orc_scene;

/* Converted world entry (objects.go:225-269 + materials.go:28-55), for flatten parity. */
typedef struct {
    int32_t type;      /* 0 sphere, 1 plane, 2 box, 3 mesh (a/b = bounding box) */
    int32_t mat_type;  /* materials.go:11-17: 0 lambert 1 metal 2 dielectric 3 emissive 4 mirror */
    double a[3];       /* sphere centre | plane point | box min */
    double b[3];       /* sphere (radius,0,0) | plane normal | box max */
    double albedo[3], rough, ior, emit[3], absorption[3];
} orc_world_entry;

typedef struct {
    uint64_t samples;        /* camera samples traced */
    uint64_t segments;       /* entries into the closest-hit scan (renderer.go:297) */
    uint64_t exit_scans;     /* dielectric exit searches (renderer.go:329) */
    uint64_t prim_tests;     /* hittable.hit calls */
    uint64_t accepts[3];     /* accepted hits per primitive type (main scan) */
    uint64_t scatters;       /* scatter() calls that returned ok */
    uint64_t end_sky, end_emissive, end_rr, end_depth, end_noscatter;
} orc_stats;

orc_scene* orc_scene_create(const orc_raw_object* objs, int n_objs,
                            const orc_raw_material* mats, int n_mats,
                            const orc_raw_camera* cam, const orc_raw_sky* sky);
void orc_scene_destroy(orc_scene*);
int  orc_world_size(const orc_scene*);
void orc_world_get(const orc_scene*, int i, orc_world_entry* out);

/* Camera constants of newCamera (camera.go:19-58) for a W x H frame, fp64:
 * out[0..2] origin, [3..5] lowerLeftCorner, [6..8] horizontal, [9..11] vertical,
 * [12..14] u, [15..17] v, [18..20] w, [21] lensRadius. */
void orc_camera(const orc_scene*, int width, int height, double out[22]);

/* Primary-ray closest hit (renderer.go:292-302 on the ray of camera.go:70-73,
 * i.e. the lens-free branch) at sample offset (xi_u, xi_v) inside every pixel.
 * ids[y*W+x] = world index or -1; t[y*W+x] = ray parameter (0 on miss). fp64. */
void orc_primary_hits(const orc_scene*, int width, int height, double xi_u, double xi_v,
                      int32_t* ids, double* t);

/* Pixel loop (renderer.go:171-187) for samples [s_begin, s_end) of every pixel,
 * RNG = counter hash keyed (seed, pixel, sample) — see DESIGN.md "RNG".
 * precision: 64 = Go-faithful binary64, 32 = same algorithm in binary32.
 * rgb_sum: W*H*3 doubles, un-normalised sum over the samples (row 0 = top).
 * threads: worker count (tile queue of 32x32 tiles, renderer.go:132-161). */
void orc_render_sum(const orc_scene*, int width, int height, int s_begin, int s_end,
                    int max_depth, uint32_t seed, int precision, int threads,
                    double* rgb_sum, orc_stats* stats);

/* Pixel epilogue (renderer.go:189-221): mean, sqrt, *255.999, clamp, truncate; A=255. */
void orc_finalize(const double* rgb_sum, int width, int height, int spp, uint8_t* rgba);

/* renderIntoCPU equivalent (renderer.go:44-246): RGBA8 out, stride 4W. Used as the CPU baseline. */
void orc_render_rgba(const orc_scene*, int width, int height, int spp, int max_depth,
                     uint32_t seed, int threads, uint8_t* rgba, orc_stats* stats);

/* Single-path trace for known-answer tests: follows rayColorOpt from the given
 * ray and records, per segment, the hit world index (-1 = sky), t and frontFace.
 * Returns number of segments recorded (<= cap). RNG keyed (seed, pixel=0, sample=0). */
int orc_trace_path(const orc_scene*, const double orig[3], const double dir[3], int max_depth,
                   uint32_t seed, int cap, int32_t* hit_ids, double* hit_t, int32_t* front_face,
                   double rgb[3]);

/* EXTENSION: 1 (default) = meshes with more than 64 triangles use the oracle's own median-split BVH, 0 = brute force. */
void orc_set_mesh_accel(orc_scene*, int enabled);

/* The shared RNG spec, exposed for tests: i-th uniform of path (seed,pixel,sample). */
double orc_rng_uniform(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t i);

#ifdef __cplusplus
}
#endif
#endif
