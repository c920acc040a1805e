// api.cu — the C ABI of include/ptb200.h: context, scene conversion (sceneToWorld + convertMaterial +
// newCamera, all in binary64 on the host exactly as the reference does them), and the render entry points.
//
// No CPU fallback exists anywhere in this file: every path either launches the CUDA kernels or returns an
// error (the reference's GL plug-in falls back to the CPU, renderer.go:257-262 — this backend must not).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <array>
#include <atomic>
#include <cfloat>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include <thread>

#include "bvh.h"
#include "scene_dev.h"

using namespace ptb;

namespace {

thread_local std::string g_create_error;

struct World64Entry {   // result of sceneToWorld + convertMaterial, binary64 (objects.go:225-269, materials.go:28-55)
    int type, mat_type;
    double a[3], b[3];
    double albedo[3], rough, ior, emit[3], absorption[3];
    int mat_slot;       // index into the device material table
};

}  // namespace

struct ptb_ctx {
    int device = 0;
    std::mutex mu;
    std::string err;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaDeviceProp prop{};

    bool has_scene = false;
    ptb_camera cam{};
    std::vector<World64Entry> world;
    HostScene hs;                       // host staging of the uploaded scene (launches copy header + scan table by value)
    LaunchCache launch_cache;           // per device: shared-memory opt-in and occupancy of each kernel instantiation
    uint4* d_blob = nullptr;            // obj[] then mat[] (global copy for the smem fill; read in place by BIG worlds)
    size_t blob_words = 0, blob_cap = 0, world64_cap = 0;
    float4* d_tab = nullptr; size_t tab_cap = 0;        // BIG worlds: the scan table in global memory
    int32_t* d_diel = nullptr; size_t diel_cap = 0;     // dielectric object indices (untyped exit search)
    Obj64* d_world64 = nullptr;
    int n_world64 = 0;
    BvhNode* d_bvh_nodes = nullptr;      // EXTENSION: mesh BVH
    BvhTri* d_bvh_tris = nullptr;
    float* d_planes = nullptr; size_t planes_cap = 0;   // partial-sum planes of split (small) frames
    int* d_trav = nullptr; size_t trav_cap = 0;   // suspended-traversal scratch of the in-kernel traversal (the default for mesh scenes)
    MeshPipe mesh_pipe;                           // mesh scenes: global path state + ray queue of the wavefront across kernels
    BvhNode* d_bvh_nodes_keep = nullptr; // last built BVH, reused when the same triangles are uploaded again
    BvhTri* d_bvh_tris_keep = nullptr;
    ptb_bvh_info bvh_keep{};
    uint64_t bvh_key = 0;                // content hash of the kept BVH's triangles and tags
    uint64_t bvh_generation = 0;         // ptb_scene.mesh_generation the kept BVH was built for (0 = none given)
    uint64_t bvh_tag_key = 0;            // hash of the per-mesh tags alone (checked with the generation)
    std::vector<std::array<double, 6>> mesh_bounds_keep;   // per mesh object of the kept BVH: its vertex bounds
    ptb_bvh_info bvh{};

    // scratch device buffers, grown on demand
    float* d_accum = nullptr; size_t accum_cap = 0;
    uint8_t* d_rgba = nullptr; size_t rgba_cap = 0;
    uint8_t* d_rgba2 = nullptr; size_t rgba2_cap = 0;   // second image of the double-buffered progressive path
    uint8_t* h_rgba = nullptr; size_t h_rgba_cap = 0;   // pinned staging for ptb_render
    uint8_t* h_rgba2 = nullptr; size_t h_rgba2_cap = 0;
    cudaStream_t copy_stream = nullptr;                 // progressive path: read-back of batch b overlaps batch b+1
    cudaEvent_t ev_batch[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
    unsigned long long* d_stats = nullptr;
    unsigned int* d_work = nullptr;
    ptb_stats stats{};
    std::string last_kernel;            // symbol of the integrator instantiation the last render launched
};

namespace {

int fail(ptb_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}
#define CK(c, call)                                                                                     \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) return fail((c), PTB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }   // materials.go:57-65

// convertMaterial, materials.go:28-55
void convert_material(const ptb_scene* s, int i, World64Entry& w) {
    const double al[3] = {s->mat_albedo[3 * i], s->mat_albedo[3 * i + 1], s->mat_albedo[3 * i + 2]};
    const double power = s->mat_power[i];
    const double em[3] = {s->mat_emit[3 * i] * power, s->mat_emit[3 * i + 1] * power, s->mat_emit[3 * i + 2] * power};
    w.rough = 0; w.ior = 0;
    for (int k = 0; k < 3; k++) { w.albedo[k] = 0; w.emit[k] = 0; w.absorption[k] = 0; }
    switch (s->mat_type[i]) {
    case PTB_MAT_METAL: {
        double rough = s->mat_rough[i];
        if (s->mat_smoothness[i] > 0) rough = 1.0 - clampd(s->mat_smoothness[i], 0, 1);
        w.mat_type = PTB_MAT_METAL;
        for (int k = 0; k < 3; k++) w.albedo[k] = al[k];
        w.rough = clampd(rough, 0, 1);
        break;
    }
    case PTB_MAT_DIELECTRIC: {
        double ior = s->mat_ior[i];
        if (ior == 0) ior = 1.5;
        w.mat_type = PTB_MAT_DIELECTRIC;
        for (int k = 0; k < 3; k++) { w.albedo[k] = al[k]; w.absorption[k] = s->mat_absorption[3 * i + k]; }
        w.ior = ior;
        break;
    }
    case PTB_MAT_EMISSIVE:
        w.mat_type = PTB_MAT_EMISSIVE;
        for (int k = 0; k < 3; k++) w.emit[k] = em[k];
        break;
    case PTB_MAT_MIRROR:
        w.mat_type = PTB_MAT_MIRROR;
        for (int k = 0; k < 3; k++) w.albedo[k] = al[k];
        break;
    default:
        w.mat_type = PTB_MAT_LAMBERT;
        for (int k = 0; k < 3; k++) w.albedo[k] = al[k];
        w.rough = clampd(s->mat_rough[i], 0, 1);
        break;
    }
}

struct Cam64 { double origin[3], llc[3], horizontal[3], vertical[3], u[3], v[3], w[3], lens_radius; };

// newCamera, camera.go:19-58 (vec3 helpers of math.go:11-37 written out; div multiplies by the reciprocal)
Cam64 new_camera(const ptb_camera& c, int width, int height) {
    auto sub = [](const double* a, const double* b, double* o) { for (int k = 0; k < 3; k++) o[k] = a[k] - b[k]; };
    auto dot = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    auto cross = [](const double* a, const double* b, double* o) {
        o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
    };
    auto unit = [&](double* a) {
        double l = std::sqrt(dot(a, a));
        if (l == 0) return;
        double inv = 1.0 / l;
        for (int k = 0; k < 3; k++) a[k] = a[k] * inv;
    };
    Cam64 r{};
    double aspect = (double)width / (double)height;
    if (c.aspect_ratio != 0) aspect = c.aspect_ratio;
    double theta = c.fov * M_PI / 180;
    double h = std::tan(theta / 2);
    double vh = 2.0 * h, vw = aspect * vh;
    for (int k = 0; k < 3; k++) r.origin[k] = c.position[k];
    double ot[3];
    sub(c.position, c.target, ot);
    for (int k = 0; k < 3; k++) r.w[k] = ot[k];
    unit(r.w);
    cross(c.up, r.w, r.u);
    unit(r.u);
    cross(r.w, r.u, r.v);
    double focus = c.focus_dist;
    if (focus == 0) focus = std::sqrt(dot(ot, ot));
    const double inv2 = 1.0 / 2.0;
    for (int k = 0; k < 3; k++) {
        r.horizontal[k] = r.u[k] * (vw * focus);
        r.vertical[k] = r.v[k] * (vh * focus);
    }
    for (int k = 0; k < 3; k++)
        r.llc[k] = r.origin[k] - r.horizontal[k] * inv2 - r.vertical[k] * inv2 - r.w[k] * focus;
    r.lens_radius = c.aperture / 2;
    return r;
}

int ensure(ptb_ctx* c, void** p, size_t* cap, size_t need, bool pinned_host = false) {
    if (*cap >= need && *p) return PTB_OK;
    if (*p) { if (pinned_host) cudaFreeHost(*p); else cudaFree(*p); *p = nullptr; *cap = 0; }
    if (pinned_host) CK(c, cudaMallocHost(p, need)); else CK(c, cudaMalloc(p, need));
    *cap = need;
    return PTB_OK;
}

int rows_of(const ptb_cfg* cfg) {
    if (cfg->row_step <= 1) return cfg->height;
    return (cfg->height - cfg->row_offset + cfg->row_step - 1) / cfg->row_step;
}

int check_cfg(ptb_ctx* c, const ptb_cfg* cfg, int& s0, int& s1) {
    if (!cfg) return fail(c, PTB_ERR_INVALID, "cfg is NULL");
    if (cfg->width < 2 || cfg->height < 2)   // 1/(W-1), 1/(H-1) (renderer.go:95-96) need W,H >= 2
        return fail(c, PTB_ERR_INVALID, "width and height must be >= 2 (got %dx%d)", cfg->width, cfg->height);
    if ((long long)cfg->width * cfg->height > (1ll << 30)) return fail(c, PTB_ERR_LIMIT, "frame too large");
    if (cfg->samples_per_px < 1) return fail(c, PTB_ERR_INVALID, "samples_per_px must be >= 1");
    if (cfg->max_depth > 65534) return fail(c, PTB_ERR_LIMIT, "max_depth %d > 65534 (the remaining depth of a path is kept in 16 bits)", cfg->max_depth);
    s0 = 0; s1 = cfg->samples_per_px;
    if (cfg->sample_count > 0) {
        s0 = cfg->sample_begin; s1 = cfg->sample_begin + cfg->sample_count;
        if (s0 < 0 || s1 > cfg->samples_per_px)
            return fail(c, PTB_ERR_INVALID, "sample range [%d,%d) outside [0,%d)", s0, s1, cfg->samples_per_px);
    }
    if (cfg->row_step <= 1 ? cfg->row_offset != 0 : (cfg->row_offset < 0 || cfg->row_offset >= cfg->row_step || cfg->row_offset >= cfg->height))
        return fail(c, PTB_ERR_INVALID, "row partition (offset %d, step %d) is not valid for %d rows", cfg->row_offset, cfg->row_step, cfg->height);
    if (!c->has_scene) return fail(c, PTB_ERR_NO_SCENE, "no scene uploaded");
    return PTB_OK;
}

uint32_t fmix_host(uint32_t x) { x ^= x >> 16; x *= 0x21f0aaadu; x ^= x >> 15; x *= 0x735a2d97u; x ^= x >> 15; return x; }

// How the samples of a render are grouped into partial-sum planes (small frames, scene_dev.h FrameParams::split_*).
// A launch covers planes [base, base + count) of `total`, whose boundaries are those of the whole range [full_begin, full_end).
struct SplitPlan { int total = 1, base = 0, count = 1, full_begin = 0, full_end = 0; };

int plan_bound(const SplitPlan& p, int g) { return p.full_begin + (int)((unsigned)(p.full_end - p.full_begin) * (unsigned)g / (unsigned)p.total); }

// split factor of a render of samples [s0, s1): chosen from the WHOLE frame and the WHOLE sample range, so that a row
// partition and a progressive render group every pixel's samples exactly like one launch does
SplitPlan make_plan(ptb_ctx* c, const ptb_cfg* cfg, int s0, int s1) {
    SplitPlan p;
    int k = std::getenv("PTB_NO_SPLIT") ? 1 : wf_split_factor(c->prop.multiProcessorCount, (long long)cfg->width * rows_of(cfg), s1 - s0);
    if (const char* force = std::getenv("PTB_SPLIT")) { k = std::atoi(force); if (k < 1) k = 1; if (k > s1 - s0) k = s1 - s0; if (k > 64) k = 64; }   // tuning
    if ((cfg->flags & PTB_FLAG_MEGAKERNEL) || cfg->max_depth <= 0) k = 1;
    p.total = k; p.base = 0; p.count = k; p.full_begin = s0; p.full_end = s1;
    return p;
}

// Launch the integrator for planes [plan.base, plan.base + plan.count) of the plan on `stream`.  accum/rgba are device
// pointers (either may be NULL).  Everything the kernel needs of the scene travels BY VALUE in its parameter block, so
// frames queued on any streams and scenes uploaded later cannot disturb each other.
int render_launch(ptb_ctx* c, const ptb_cfg* cfg, const SplitPlan& plan, float* d_accum, uint8_t* d_rgba, cudaStream_t stream,
                  bool resume = false) {
    const int W = cfg->width, H = cfg->height;
    const int s0 = plan_bound(plan, plan.base), s1 = plan_bound(plan, plan.base + plan.count);
    const HostScene& hs = c->hs;
    KernelArgs ka;
    static_cast<SceneHdr&>(ka.sc) = hs.hdr;
    // camera for this frame size (newCamera is per render in the reference too, renderer.go:168)
    Cam64 cam = new_camera(c->cam, W, H);
    DevCamera& dc = ka.sc.cam;
    for (int k = 0; k < 3; k++) {
        dc.origin[k] = (float)cam.origin[k]; dc.llc[k] = (float)cam.llc[k];
        dc.horizontal[k] = (float)cam.horizontal[k]; dc.vertical[k] = (float)cam.vertical[k];
        dc.u[k] = (float)cam.u[k]; dc.v[k] = (float)cam.v[k];
    }
    dc.lens_radius = (float)cam.lens_radius;
    ka.sc.tab_global = c->d_tab; ka.sc.diel_idx = c->d_diel;
    if (!hs.big && hs.hdr.tab_floats > 0) std::memcpy(ka.sc.scan_tab, hs.scan_tab.data(), sizeof(float) * (size_t)hs.hdr.tab_floats);

    const bool stats = (cfg->flags & PTB_FLAG_STATS) != 0;
    const bool dbg_timing = std::getenv("PTB_DEBUG_TIMING") != nullptr;   // only meaningful in -DPTB_WF_TIMING builds
    if (stats || dbg_timing) CK(c, cudaMemsetAsync(c->d_stats, 0, sizeof(unsigned long long) * (kStatsWords + 24), stream));

    const int R = rows_of(cfg);                       // rows this launch outputs (row partition: a subset, compact)
    FrameParams& fp = ka.fp;
    fp = FrameParams{};
    fp.width = W; fp.height = H; fp.s_begin = s0; fp.s_end = s1;
    fp.rows = R; fp.row_offset = cfg->row_step > 1 ? cfg->row_offset : 0; fp.row_step = cfg->row_step > 1 ? cfg->row_step : 1;
    fp.spp_total = cfg->samples_per_px; fp.max_depth = cfg->max_depth;
    fp.seed_key = fmix_host(cfg->seed ^ 0x9E3779B9u);
    fp.inv_w = 1.0f / (float)(W - 1); fp.inv_h = 1.0f / (float)(H - 1); fp.h_minus_1 = (float)(H - 1);
    fp.scene_blob = c->d_blob; fp.accum = d_accum; fp.accum_resume = resume ? 1 : 0; fp.rgba = d_rgba; fp.stats = (stats || dbg_timing) ? c->d_stats : nullptr;
    fp.bvh_nodes = (const float4*)c->d_bvh_nodes; fp.bvh_tris = (const float4*)c->d_bvh_tris;
    fp.split_k = 1; fp.split_base = plan.base; fp.split_total = plan.total; fp.full_begin = plan.full_begin; fp.full_end = plan.full_end;
    int e;
    if (cfg->max_depth <= 0) {                        // rayColorOpt returns black at depth <= 0 (renderer.go:287-289)
        c->last_kernel = "clear_frame_kernel";
        e = launch_clear_frame(d_accum, resume ? 1 : 0, d_rgba, W * R, stream);
        if (e) return fail(c, PTB_ERR_CUDA, "clear launch: %s", cudaGetErrorString((cudaError_t)e));
        return PTB_OK;
    }
    // mesh scenes: in-kernel traversal (wavefront.cuh), or — experiment, PTB_MESH_PIPELINE=1 — the wavefront across kernels (mesh_pipeline.cuh)
    const bool mesh_pipeline = c->d_bvh_nodes && !(cfg->flags & PTB_FLAG_MEGAKERNEL) && std::getenv("PTB_MESH_PIPELINE") != nullptr;
    if (c->d_bvh_nodes && !mesh_pipeline) {
        int rc = ensure(c, (void**)&c->d_trav, &c->trav_cap, wf_trav_scratch_bytes(c->prop.multiProcessorCount));
        if (rc) return rc;
        fp.trav_scratch = c->d_trav;
    }
    if (cfg->flags & PTB_FLAG_MEGAKERNEL) {
        if (c->d_bvh_nodes) return fail(c, PTB_ERR_INVALID, "the megakernel integrator does not support meshes");
        if (R != H) return fail(c, PTB_ERR_INVALID, "the megakernel integrator does not support row partitions");
        if (hs.big) return fail(c, PTB_ERR_LIMIT, "the megakernel integrator supports only worlds that fit the kernel-parameter table (%d objects here)", hs.hdr.n_obj);
        c->last_kernel = stats ? "integrate_kernel<1>" : "integrate_kernel<0>";
        e = launch_integrator(ka, stats, stream);
    } else {
        fp.work_counter = c->d_work;
        CK(c, cudaMemsetAsync(c->d_work, 0, sizeof(unsigned int), stream));
        const int k = plan.count;
        if (plan.total > 1) {                          // small frame: (pixel, sample sub-range) work items, summed afterwards
            int rc = ensure(c, (void**)&c->d_planes, &c->planes_cap, (size_t)k * W * R * 3 * sizeof(float));
            if (rc) return rc;
            fp.split_k = k; fp.planes = c->d_planes;
        }
        // sphere-rich scenes: the packed two-ray sphere test (PTB_PACKED=0/1 forces the choice: tests, tuning)
        bool packed = hs.hdr.n_sphere_run >= kPackedSphereMin;
        if (const char* force = std::getenv("PTB_PACKED")) packed = std::atoi(force) != 0;
        c->last_kernel = std::string("integrate_wf_kernel<") + (stats ? "1, " : "0, ") + (c->d_bvh_nodes ? "1, " : "0, ") + (hs.big ? "1, " : "0, ") + (packed ? "1>" : "0>");
        if (mesh_pipeline) {
            MeshPipe& mp = c->mesh_pipe;
            const int n_slots = mesh_pool_slots(c->prop.multiProcessorCount, (long long)W * R * k);
            const size_t need = mesh_pool_bytes(n_slots);
            if (mp.d_bytes < need) {
                cudaFree(mp.d_mem); mp.d_mem = nullptr; mp.d_bytes = 0;
                CK(c, cudaMalloc(&mp.d_mem, need));
                mp.d_bytes = need;
            }
            mesh_pool_bind(mp.pool, mp.d_mem, n_slots);
            if (!mp.h_flags) CK(c, cudaMallocHost((void**)&mp.h_flags, 4 * sizeof(unsigned int)));
            for (void*& ev : mp.events) if (!ev) { cudaEvent_t x; CK(c, cudaEventCreateWithFlags(&x, cudaEventDisableTiming)); ev = x; }
            c->last_kernel = std::string("mp_shade_scan_kernel<") + (stats ? "1, " : "0, ") + (hs.big ? "1>" : "0>") + " + mp_traverse_kernel<" + (stats ? "1>" : "0>");
            // the pipeline replays a CUDA graph, which the legacy default stream cannot capture: it runs on the context's own
            // stream, ordered behind the caller's stream, and has finished when the call returns
            CK(c, cudaEventRecord(c->ev_batch[0], stream));
            CK(c, cudaStreamWaitEvent(c->stream, c->ev_batch[0], 0));
            e = launch_mesh_pipeline(ka, stats, hs.big, c->prop.multiProcessorCount, &c->launch_cache, mp, c->stream);
        } else
        e = launch_integrator_wf(ka, stats, hs.big, packed, c->prop.multiProcessorCount, &c->launch_cache, stream);
        if (plan.total > 1) c->last_kernel += " + finalize_planes_kernel";
        if (!e && plan.total > 1) e = launch_finalize_planes(c->d_planes, k, W, R, cfg->samples_per_px, d_accum, resume ? 1 : 0, d_rgba, stream);
    }
    if (e) return fail(c, PTB_ERR_CUDA, "integrator launch: %s", cudaGetErrorString((cudaError_t)e));
    return PTB_OK;
}

int copy_image_out(ptb_ctx* c, const uint8_t* d_img, uint8_t* staging, uint8_t* rgba, size_t stride, int W, int H, cudaStream_t stream);

int fetch_stats(ptb_ctx* c, bool stats, float ms) {
    c->stats.last_render_ms = ms;
    if (std::getenv("PTB_DEBUG_TIMING")) {
        unsigned long long t[6];
        cudaMemcpy(t, c->d_stats + kStatsWords, sizeof t, cudaMemcpyDeviceToHost);
        double tot = 0; for (int k = 0; k < 6; k++) tot += (double)t[k];
        if (tot > 0 && t[4] > 0) {
            const double iters = (double)t[4];                 // CTA-iterations (thread 0 of every CTA)
            const double warps_iters = iters * 8.0;            // lane 0 of every warp contributes to t[0..2]
            std::fprintf(stderr, "wf cycles per CTA-iteration: scan %.0f  sort %.0f  shade(mean over warps) %.0f  shade(max over warps) %.0f\n",
                         t[0] / warps_iters, t[1] / warps_iters, t[2] / warps_iters, t[3] / iters);
            unsigned long long cc[12];
            cudaMemcpy(cc, c->d_stats + kStatsWords + 8, sizeof cc, cudaMemcpyDeviceToHost);
            const char* names[6] = {"DIEL", "TERM", "REGEN", "DIFFUSE", "SPEC", "CONT"};
            for (int k = 0; k < 6; k++) if (cc[2 * k + 1]) std::fprintf(stderr, "   chunk class %-8s mean %.0f cycles  (%llu chunks)\n", names[k], (double)cc[2 * k] / cc[2 * k + 1], cc[2 * k + 1]);
        }
    }
    if (!stats) return PTB_OK;
    unsigned long long w[kStatsWords];
    CK(c, cudaMemcpy(w, c->d_stats, sizeof w, cudaMemcpyDeviceToHost));
    ptb_stats& s = c->stats;
    s.samples = w[ST_SAMPLES]; s.segments = w[ST_SEGMENTS]; s.exit_scans = w[ST_EXIT_SCANS];
    s.accepts[0] = w[ST_ACC_SPHERE]; s.accepts[1] = w[ST_ACC_PLANE]; s.accepts[2] = w[ST_ACC_BOX];
    s.scatters = w[ST_SCATTERS]; s.end_sky = w[ST_END_SKY]; s.end_emissive = w[ST_END_EMISSIVE];
    s.end_rr = w[ST_END_RR]; s.end_depth = w[ST_END_DEPTH]; s.end_noscatter = w[ST_END_NOSCATTER];
    s.lane_iters_active = w[ST_LANE_ACTIVE]; s.lane_iters_total = w[ST_LANE_TOTAL];
    s.accepts_mesh = w[ST_ACC_MESH]; s.bvh_nodes_visited = w[ST_BVH_NODES]; s.bvh_tris_tested = w[ST_BVH_TRIS];
    s.bvh_stack_overflows = w[ST_BVH_STACK_OVERFLOW];
    return PTB_OK;
}

}  // namespace

extern "C" {

int ptb_abi_version(void) { return PTB_ABI_VERSION; }
int ptb_rows_of(const ptb_cfg* cfg) { return cfg ? rows_of(cfg) : PTB_ERR_INVALID; }

int ptb_create(int device, ptb_ctx** out) {
    if (!out) return fail(nullptr, PTB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(nullptr, PTB_ERR_CUDA, "cudaGetDeviceCount: %s (no CPU fallback exists)", cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(nullptr, PTB_ERR_INVALID, "device %d out of range (%d CUDA devices)", device, n);
    ptb_ctx* c = new ptb_ctx();
    c->device = device;
    auto bail = [&](const char* what, cudaError_t err) {
        int rc = fail(nullptr, PTB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(err));
        ptb_destroy(c);
        return rc;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    if ((e = cudaGetDeviceProperties(&c->prop, device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
    if (c->prop.major < 10) {
        int rc = fail(nullptr, PTB_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, c->prop.major, c->prop.minor);
        delete c;
        return rc;
    }
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaEventCreate(&c->ev0)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&c->ev1)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    for (int k = 0; k < 2; k++) {
        if ((e = cudaEventCreateWithFlags(&c->ev_batch[k], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
        if ((e = cudaEventCreateWithFlags(&c->ev_copy[k], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    }
    if ((e = cudaMalloc((void**)&c->d_stats, sizeof(unsigned long long) * (kStatsWords + 24))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc((void**)&c->d_work, sizeof(unsigned int))) != cudaSuccess) return bail("cudaMalloc", e);
    *out = c;
    return PTB_OK;
}

void ptb_destroy(ptb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();            // frames queued on caller streams may still read this context's buffers
    cudaFree(c->mesh_pipe.d_mem);
    if (c->mesh_pipe.h_flags) cudaFreeHost(c->mesh_pipe.h_flags);
    for (void* e : c->mesh_pipe.events) if (e) cudaEventDestroy((cudaEvent_t)e);
    cudaFree(c->d_planes); cudaFree(c->d_trav); cudaFree(c->d_blob); cudaFree(c->d_tab); cudaFree(c->d_diel); cudaFree(c->d_world64);
    cudaFree(c->d_bvh_nodes_keep); cudaFree(c->d_bvh_tris_keep); cudaFree(c->d_accum); cudaFree(c->d_rgba); cudaFree(c->d_rgba2);
    cudaFree(c->d_stats); cudaFree(c->d_work);
    if (c->h_rgba) cudaFreeHost(c->h_rgba);
    if (c->h_rgba2) cudaFreeHost(c->h_rgba2);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (int k = 0; k < 2; k++) { if (c->ev_batch[k]) cudaEventDestroy(c->ev_batch[k]); if (c->ev_copy[k]) cudaEventDestroy(c->ev_copy[k]); }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* ptb_last_error(const ptb_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int ptb_get_device_info(ptb_ctx* c, ptb_device_info* out) {
    if (!c || !out) return fail(c, PTB_ERR_INVALID, "NULL argument");
    std::memset(out, 0, sizeof *out);
    std::snprintf(out->name, sizeof out->name, "%.127s", c->prop.name);
    out->sm_count = c->prop.multiProcessorCount; out->cc_major = c->prop.major; out->cc_minor = c->prop.minor;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device);
    out->clock_khz = khz;
    out->global_mem_bytes = c->prop.totalGlobalMem;
    return PTB_OK;
}

}  // extern "C"

namespace {
// DevObj::meta / triangle tag: type, dielectric bit, shading class of the wavefront kernel (enum in wavefront.cuh: 0 dielectric,
// 1 terminate (emissive), 3 diffuse (lambert, rough metal), 4 specular (mirror, smooth metal)), material slot.
int meta_of(const World64Entry& w) {
    int cls = 3;
    if (w.mat_type == PTB_MAT_METAL) cls = ((float)w.rough > 1e-6f) ? 3 : 4;
    else if (w.mat_type == PTB_MAT_MIRROR) cls = 4;
    else if (w.mat_type == PTB_MAT_DIELECTRIC) cls = 0;
    else if (w.mat_type == PTB_MAT_EMISSIVE) cls = 1;
    return w.type | ((w.mat_type == PTB_MAT_DIELECTRIC) << 2) | (cls << 3) | (w.mat_slot << 6);
}

// convertMaterial + sceneToWorld in binary64 (materials.go:28-55, objects.go:225-269), plus the triangle soup of the mesh objects.
struct WorldBuild {
    std::vector<World64Entry> mats, world;
    struct MeshRef { int64_t t0, t1; int32_t world_idx; };   // triangles [t0, t1) of ptb_scene::tri_vertices belong to world[world_idx]
    std::vector<MeshRef> meshes;               // the non-empty mesh objects in world order
    int64_t n_tris = 0;
    int n_analytic = 0;
};
// trusted_bounds: per mesh object (in world order) the vertex bounds found by an earlier upload of the SAME triangles
// (ptb_scene.mesh_generation matched): the pass over the vertices is skipped.
int build_world(ptb_ctx* c, const ptb_scene* s, WorldBuild& wb, const std::vector<std::array<double, 6>>* trusted_bounds = nullptr) {
    std::vector<World64Entry>& mats = wb.mats;
    std::vector<World64Entry>& world = wb.world;
    int& n_analytic = wb.n_analytic;
    // materials (convertMaterial) + the zero material in slot n_mat (objects.go:234: missing map key)
    mats.assign(s->n_mat + 1, World64Entry{});
    for (int i = 0; i < s->n_mat; i++) convert_material(s, i, mats[i]);
    {
        World64Entry& z = mats[s->n_mat];
        std::memset(&z, 0, sizeof z);
        z.mat_type = PTB_MAT_LAMBERT;
    }
    // objects (sceneToWorld); EXTENSION: mesh objects are world entries too (type PTB_OBJ_MESH, a/b = bounding box)
    if (s->n_mesh < 0 || (s->n_mesh > 0 && (!s->obj_mesh || !s->mesh_tri_begin || !s->tri_vertices)))
        return fail(c, PTB_ERR_INVALID, "mesh arrays are NULL");
    std::vector<char> mesh_used(s->n_mesh > 0 ? s->n_mesh : 0, 0);
    for (int i = 0; i < s->n_obj; i++) {
        const int t = s->obj_type[i];
        if (t != PTB_OBJ_SPHERE && t != PTB_OBJ_PLANE && t != PTB_OBJ_BOX && t != PTB_OBJ_MESH) continue;   // dropped (objects.go:237-266)
        int mi = s->obj_mat[i];
        if (mi >= s->n_mat) return fail(c, PTB_ERR_INVALID, "obj_mat[%d]=%d out of range", i, mi);
        if (mi < 0) mi = s->n_mat;
        World64Entry w = mats[mi];
        w.mat_slot = mi;
        w.type = t;
        const double* pos = s->obj_pos + 3 * i;
        const double* size = s->obj_size + 3 * i;
        if (t == PTB_OBJ_SPHERE) { for (int k = 0; k < 3; k++) { w.a[k] = pos[k]; w.b[k] = 0; } w.b[0] = size[0]; }
        else if (t == PTB_OBJ_PLANE) { for (int k = 0; k < 3; k++) w.a[k] = pos[k]; w.b[0] = 0; w.b[1] = 1; w.b[2] = 0; }
        else if (t == PTB_OBJ_BOX) { for (int k = 0; k < 3; k++) { w.a[k] = pos[k] - size[k] * 0.5; w.b[k] = pos[k] + size[k] * 0.5; } }
        else {
            const int m = s->n_mesh > 0 ? s->obj_mesh[i] : -1;
            if (m < 0 || m >= s->n_mesh) return fail(c, PTB_ERR_INVALID, "obj_mesh[%d]=%d out of range", i, m);
            if (mesh_used[m]) return fail(c, PTB_ERR_INVALID, "mesh %d is referenced by more than one object", m);
            mesh_used[m] = 1;
            const int64_t t0 = s->mesh_tri_begin[m], t1 = s->mesh_tri_begin[m + 1];
            if (t0 < 0 || t1 < t0) return fail(c, PTB_ERR_INVALID, "mesh_tri_begin is not monotone");
            if (t1 == t0) continue;            // empty mesh: dropped
            if (wb.n_tris + (t1 - t0) > (1ll << 28)) return fail(c, PTB_ERR_LIMIT, "more than 2^28 triangles");
            if (trusted_bounds && wb.meshes.size() < trusted_bounds->size()) {
                const std::array<double, 6>& bb = (*trusted_bounds)[wb.meshes.size()];
                for (int k = 0; k < 3; k++) { w.a[k] = bb[k]; w.b[k] = bb[3 + k]; }
            } else {
                float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
                bool finite = true;
                const float* v = s->tri_vertices + t0 * 9;
                for (int64_t q = 0; q < (t1 - t0) * 3; q++, v += 3) {          // one vertex per trip; NaN fails the range test
                    finite &= (std::fabs(v[0]) <= 1e18f) & (std::fabs(v[1]) <= 1e18f) & (std::fabs(v[2]) <= 1e18f);
                    lo[0] = std::min(lo[0], v[0]); lo[1] = std::min(lo[1], v[1]); lo[2] = std::min(lo[2], v[2]);
                    hi[0] = std::max(hi[0], v[0]); hi[1] = std::max(hi[1], v[1]); hi[2] = std::max(hi[2], v[2]);
                }
                if (!finite) return fail(c, PTB_ERR_INVALID, "non-finite mesh vertex");
                for (int k = 0; k < 3; k++) { w.a[k] = lo[k]; w.b[k] = hi[k]; }
            }
            wb.meshes.push_back({t0, t1, (int32_t)world.size()});
            wb.n_tris += t1 - t0;
        }
        if (t != PTB_OBJ_MESH) n_analytic++;
        world.push_back(w);
    }
    if (n_analytic > PTB_MAX_OBJECTS) return fail(c, PTB_ERR_LIMIT, "%d analytic objects > PTB_MAX_OBJECTS=%d", n_analytic, PTB_MAX_OBJECTS);
    return PTB_OK;
}

// binary32 device tables: materials, analytic objects in device order, the scan tables and the exit-search tables
// (scene_dev.h).  w64 = the analytic objects in world order for the binary64 parity kernel.
void build_device_tables(const WorldBuild& wb, int n_mat_in, HostScene& out, std::vector<Obj64>& w64) {
    const std::vector<World64Entry>& mats = wb.mats;
    const std::vector<World64Entry>& world = wb.world;
    const int n_analytic = wb.n_analytic;
    SceneHdr& hs = out.hdr;
    hs = SceneHdr{};
    hs.n_obj = n_analytic; hs.n_mat = n_mat_in + 1; hs.n_diel = 0;
    out.mat.assign((size_t)hs.n_mat, DevMat{});
    out.obj.assign((size_t)hs.n_obj, DevObj{});
    out.diel_idx.clear();
    for (int i = 0; i < hs.n_mat; i++) {
        DevMat& m = out.mat[i];
        const World64Entry& w = mats[i];
        m.type = w.mat_type; m.rough = (float)w.rough; m.ior = (float)w.ior;
        for (int k = 0; k < 3; k++) { m.albedo[k] = (float)w.albedo[k]; m.emit[k] = (float)w.emit[k]; m.absorption[k] = (float)w.absorption[k]; }
    }
    // device order of the analytic objects: boxes first (world order), then spheres/planes (world order); see
    // integrator.cu for why the reference's tie rule survives this grouping
    std::vector<int> order, dev_of(world.size(), -1);
    for (int i = 0; i < (int)world.size(); i++) if (world[i].type == PTB_OBJ_BOX) order.push_back(i);
    hs.n_box = (int)order.size();
    {   // non-box objects keep their world order; the leading planes and the spheres right after them get typed tables
        std::vector<int> rest;
        for (int i = 0; i < (int)world.size(); i++) if (world[i].type == PTB_OBJ_SPHERE || world[i].type == PTB_OBJ_PLANE) rest.push_back(i);
        size_t k = 0;
        while (k < rest.size() && world[rest[k]].type == PTB_OBJ_PLANE) k++;
        hs.n_plane_run = (int)k;
        while (k < rest.size() && world[rest[k]].type == PTB_OBJ_SPHERE) k++;
        hs.n_sphere_run = (int)k - hs.n_plane_run;
        order.insert(order.end(), rest.begin(), rest.end());
        hs.n_typed = hs.n_box + hs.n_plane_run + hs.n_sphere_run;
    }
    w64.clear();
    for (int i = 0; i < (int)world.size(); i++) {
        if (world[i].type == PTB_OBJ_MESH) continue;
        Obj64 o64;
        o64.type = world[i].type; o64.pad = i;
        for (int j = 0; j < 3; j++) { o64.a[j] = world[i].a[j]; o64.b[j] = world[i].b[j]; }
        w64.push_back(o64);
    }
    for (int k = 0; k < hs.n_obj; k++) {
        const int i = order[k];
        dev_of[i] = k;
        const World64Entry& w = world[i];
        DevObj& o = out.obj[k];
        o.ax = (float)w.a[0]; o.ay = (float)w.a[1]; o.az = (float)w.a[2];
        if (w.type == PTB_OBJ_SPHERE) {
            float r = (float)w.b[0];
            o.bx = r; o.by = r * r; o.bz = 1.0f / r;    // radiusSq (objects.go:46), invRadius (objects.go:68)
        } else if (w.type == PTB_OBJ_BOX) {      // (centre, half extent), see hit_box
            o.ax = (float)((w.a[0] + w.b[0]) * 0.5); o.ay = (float)((w.a[1] + w.b[1]) * 0.5); o.az = (float)((w.a[2] + w.b[2]) * 0.5);
            o.bx = (float)((w.b[0] - w.a[0]) * 0.5); o.by = (float)((w.b[1] - w.a[1]) * 0.5); o.bz = (float)((w.b[2] - w.a[2]) * 0.5);
        } else { o.bx = (float)w.b[0]; o.by = (float)w.b[1]; o.bz = (float)w.b[2]; }
        o.meta = meta_of(w);
        o.world_idx = i;
    }
    std::vector<float>& tab = out.scan_tab;
    tab.clear();
    {   // scan tables (scene_dev.h)
        hs.n_box_groups = (hs.n_box + kBoxGroup - 1) / kBoxGroup;
        for (int k = 0; k < hs.n_box_groups * kBoxGroup; k++) {
            if (k < hs.n_box) { const DevObj& o = out.obj[k]; tab.insert(tab.end(), {o.ax, o.ay, o.az, o.bx, o.by, o.bz}); }
            else tab.insert(tab.end(), {0.0f, 0.0f, 0.0f, -1.0f, -1.0f, -1.0f});      // far < near on every axis: never hit
        }
        tab.resize((tab.size() + 3) / 4 * 4, 0.0f);
        hs.plane_off4 = (int)(tab.size() / 4);
        for (int k = 0; k < hs.n_plane_run; k++) tab.push_back(out.obj[hs.n_box + k].ay);
        tab.resize((tab.size() + 3) / 4 * 4, 0.0f);
        hs.sphere_off4 = (int)(tab.size() / 4);
        hs.n_sphere_groups = (hs.n_sphere_run + kSphereGroup - 1) / kSphereGroup;
        for (int k = 0; k < hs.n_sphere_groups * kSphereGroup; k++) {
            if (k < hs.n_sphere_run) { const DevObj& o = out.obj[hs.n_box + hs.n_plane_run + k]; tab.insert(tab.end(), {o.ax, o.ay, o.az, o.by}); }
            else tab.insert(tab.end(), {0.0f, 0.0f, 0.0f, -1.0f});                    // radius^2 = -1: discriminant < 0 for every ray
        }
    }
    for (int i = 0; i < (int)world.size(); i++)      // exit-search candidates: analytic dielectric objects (meshes are not searched)
        if (world[i].type != PTB_OBJ_MESH && world[i].mat_type == PTB_MAT_DIELECTRIC) out.diel_idx.push_back(dev_of[i]);
    hs.n_diel = (int)out.diel_idx.size();
    {   // typed copies for the exit search (scene_dev.h)
        int nb = 0, ns = 0, other = 0;
        for (int k = 0; k < hs.n_diel; k++) {
            const int t = out.obj[out.diel_idx[k]].meta & 3;
            if (t == PTB_OBJ_BOX) nb++; else if (t == PTB_OBJ_SPHERE) ns++; else other++;
        }
        hs.exit_typed = (other == 0 && nb <= kMaxExitTyped && ns <= kMaxExitTyped) ? 1 : 0;
        hs.n_dbox = hs.n_dsph = 0; hs.dbox_off4 = hs.dsph_off4 = 0;
        if (hs.exit_typed) {
            hs.dbox_off4 = (int)(tab.size() / 4);
            for (int k = 0; k < hs.n_diel; k++) {
                const DevObj& o = out.obj[out.diel_idx[k]];
                if ((o.meta & 3) != PTB_OBJ_BOX) continue;
                hs.n_dbox++;
                tab.insert(tab.end(), {o.ax, o.ay, o.az, 0.0f, o.bx, o.by, o.bz, 0.0f});
            }
            hs.dsph_off4 = (int)(tab.size() / 4);
            for (int k = 0; k < hs.n_diel; k++) {
                const DevObj& o = out.obj[out.diel_idx[k]];
                if ((o.meta & 3) != PTB_OBJ_SPHERE) continue;
                hs.n_dsph++;
                tab.insert(tab.end(), {o.ax, o.ay, o.az, o.by});      // radius^2
            }
        }
    }
    hs.tab_floats = (int)tab.size();
    for (int k = 0; k < 4; k++) { hs.mesh_c[k] = 0.0f; hs.mesh_h[k] = -1.0f; }
    const size_t blob_bytes = (size_t)hs.n_obj * sizeof(DevObj) + (size_t)hs.n_mat * sizeof(DevMat);
    out.big = hs.tab_floats > kTabFloats || blob_bytes > (size_t)kSmallBlobBytes || std::getenv("PTB_FORCE_BIG") != nullptr;
}

// EXTENSION: bounds of all mesh triangles and the BVH over them (built on the host, bvh.cpp).  Re-uploading the same meshes
// (RenderInto takes the scene on every call, renderer.go:34) does not rebuild.  Two ways to recognise "the same":
//   * ptb_scene.mesh_generation != 0: the caller's promise that equal generations mean equal triangle arrays — O(n_mesh)
//     per call, the 360 MB of a 10 M-triangle mesh are not touched at all;
//   * otherwise a content hash of the triangles, computed by all host cores (memory-bandwidth bound).
// The result is STAGED in `st` and installed by the caller only when the whole upload succeeded.
struct MeshStage {
    BvhNode* nodes = nullptr; BvhTri* tris = nullptr;      // what the context will render with (may alias the kept BVH)
    BvhNode* fresh_nodes = nullptr; BvhTri* fresh_tris = nullptr;   // newly allocated: owned by the stage until committed
    ptb_bvh_info info{};
    uint64_t key = 0, tag_key = 0, generation = 0;
    std::vector<std::array<double, 6>> bounds;
    float mesh_c[4] = {0, 0, 0, 0}, mesh_h[4] = {-1, -1, -1, -1};
    void drop() { cudaFree(fresh_nodes); cudaFree(fresh_tris); fresh_nodes = nullptr; fresh_tris = nullptr; }
};

uint64_t fnv64(uint64_t h, const void* p, size_t n) {
    const uint64_t* w = (const uint64_t*)p;
    for (size_t i = 0; i < n / 8; i++) { uint64_t x; std::memcpy(&x, w + i, 8); h ^= x; h *= 0x100000001b3ull; }
    const unsigned char* b = (const unsigned char*)p + (n / 8) * 8;
    for (size_t i = 0; i < n % 8; i++) { h ^= b[i]; h *= 0x100000001b3ull; }
    return h;
}
// hash of a large array: 4 MB blocks hashed independently by `threads` workers, block hashes combined in order
uint64_t fnv64_parallel(const void* p, size_t n, int threads) {
    constexpr size_t kBlock = 4u << 20;
    const size_t n_blocks = (n + kBlock - 1) / kBlock;
    if (n_blocks <= 1 || threads <= 1) return fnv64(0xcbf29ce484222325ull, p, n);
    std::vector<uint64_t> part(n_blocks);
    std::atomic<size_t> next{0};
    auto work = [&]() {
        for (size_t b; (b = next.fetch_add(1)) < n_blocks;) {
            const size_t off = b * kBlock;
            part[b] = fnv64(0xcbf29ce484222325ull, (const char*)p + off, std::min(kBlock, n - off));
        }
    };
    std::vector<std::thread> pool;
    const int nt = (int)std::min<size_t>((size_t)threads, n_blocks);
    for (int t = 1; t < nt; t++) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    return fnv64(0xcbf29ce484222325ull, part.data(), part.size() * sizeof(uint64_t));
}

bool mesh_generation_matches(const ptb_ctx* c, const ptb_scene* s) {
    return s->n_mesh > 0 && s->mesh_generation != 0 && c->d_bvh_nodes_keep && c->bvh_generation == s->mesh_generation;
}

int build_mesh_accel(ptb_ctx* c, const ptb_scene* s, const WorldBuild& wb, MeshStage& st) {
    const std::vector<World64Entry>& world = wb.world;
    if (wb.n_tris == 0) return PTB_OK;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (const WorldBuild::MeshRef& m : wb.meshes) {     // union of the mesh objects' boxes (computed from their vertices in build_world)
        const World64Entry& w = world[m.world_idx];
        st.bounds.push_back({w.a[0], w.a[1], w.a[2], w.b[0], w.b[1], w.b[2]});
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], (float)w.a[k]); hi[k] = std::max(hi[k], (float)w.b[k]); }
    }
    for (int k = 0; k < 3; k++) {        // padded like the BVH boxes so that the prefilter never rejects what the root would accept
        const float pad = 1e-5f * std::max(1.0f, std::max(std::fabs(lo[k]), std::fabs(hi[k])));
        st.mesh_c[k] = 0.5f * (lo[k] + hi[k]); st.mesh_h[k] = 0.5f * (hi[k] - lo[k]) + 2.0f * pad;
    }
    unsigned hw = std::thread::hardware_concurrency();
    const int threads = hw ? (int)hw : 4;
    uint64_t tag_key = 0xcbf29ce484222325ull;            // the tags every triangle of a mesh carries (cheap: per mesh)
    for (const WorldBuild::MeshRef& m : wb.meshes) {
        const int64_t tag[3] = {m.t1 - m.t0, (int64_t)m.world_idx, (int64_t)meta_of(world[m.world_idx])};
        tag_key = fnv64(tag_key, tag, sizeof tag);
    }
    st.tag_key = tag_key; st.generation = s->mesh_generation;
    if (mesh_generation_matches(c, s) && c->bvh_tag_key == tag_key && c->bvh_keep.n_triangles == wb.n_tris) {
        st.nodes = c->d_bvh_nodes_keep; st.tris = c->d_bvh_tris_keep; st.info = c->bvh_keep; st.key = c->bvh_key;
        return PTB_OK;
    }
    uint64_t key = tag_key;
    for (const WorldBuild::MeshRef& m : wb.meshes) {     // vertices in place (no copy)
        const uint64_t h = fnv64_parallel(s->tri_vertices + m.t0 * 9, (size_t)(m.t1 - m.t0) * 9 * sizeof(float), threads);
        key = fnv64(key, &h, sizeof h);
    }
    st.key = key;
    if (c->d_bvh_nodes_keep && c->bvh_key == key && c->bvh_keep.n_triangles == wb.n_tris) {
        st.nodes = c->d_bvh_nodes_keep; st.tris = c->d_bvh_tris_keep; st.info = c->bvh_keep;
        return PTB_OK;
    }
    std::vector<float> tri_v;                  // triangles of the mesh objects, concatenated in world order
    std::vector<int32_t> tri_world, tri_meta;  // per triangle: world index of its mesh object, DevObj-style tag
    tri_v.reserve((size_t)wb.n_tris * 9); tri_world.reserve((size_t)wb.n_tris); tri_meta.reserve((size_t)wb.n_tris);
    for (const WorldBuild::MeshRef& m : wb.meshes) {
        tri_v.insert(tri_v.end(), s->tri_vertices + m.t0 * 9, s->tri_vertices + m.t1 * 9);
        tri_world.insert(tri_world.end(), (size_t)(m.t1 - m.t0), m.world_idx);
        tri_meta.insert(tri_meta.end(), (size_t)(m.t1 - m.t0), (int32_t)meta_of(world[m.world_idx]));
    }
    BvhBuildInput in{tri_v.data(), (int64_t)tri_world.size(), tri_meta.data(), tri_world.data()};
    BvhBuildOutput out;
    build_bvh(in, out, threads);
    if (out.max_stack >= 40) return fail(c, PTB_ERR_LIMIT, "BVH needs %d pending traversal entries: more than the traversal stack holds", out.max_stack);
    cudaError_t e = cudaMalloc((void**)&st.fresh_nodes, out.nodes.size() * sizeof(BvhNode));
    if (e == cudaSuccess) e = cudaMalloc((void**)&st.fresh_tris, out.tris.size() * sizeof(BvhTri));
    if (e == cudaSuccess) e = cudaMemcpy(st.fresh_nodes, out.nodes.data(), out.nodes.size() * sizeof(BvhNode), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(st.fresh_tris, out.tris.data(), out.tris.size() * sizeof(BvhTri), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { st.drop(); return fail(c, PTB_ERR_CUDA, "BVH upload: %s", cudaGetErrorString(e)); }
    st.nodes = st.fresh_nodes; st.tris = st.fresh_tris;
    st.info.n_triangles = (int64_t)out.tris.size(); st.info.n_nodes = (int64_t)out.nodes.size();
    st.info.max_depth = out.max_depth; st.info.sah_cost = out.sah_cost; st.info.build_ms = out.build_ms;
    st.info.node_bytes = sizeof(BvhNode); st.info.triangle_bytes = sizeof(BvhTri);
    return PTB_OK;
}
}  // namespace

extern "C" {

// The upload is staged: every table is built in temporaries first and the context changes only when nothing can fail any
// more, so a rejected scene (limit, bad index, allocation failure) leaves the previous scene fully usable.
int ptb_scene_upload(ptb_ctx* c, const ptb_scene* s) {
    if (!c) return PTB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!s) return fail(c, PTB_ERR_INVALID, "scene is NULL");
    if (s->n_obj < 0 || s->n_mat < 0) return fail(c, PTB_ERR_INVALID, "negative counts");
    if (s->n_obj > 0 && (!s->obj_type || !s->obj_mat || !s->obj_pos || !s->obj_size)) return fail(c, PTB_ERR_INVALID, "object arrays are NULL");
    if (s->n_mat > 0 && (!s->mat_type || !s->mat_albedo || !s->mat_rough || !s->mat_ior || !s->mat_emit || !s->mat_power ||
                         !s->mat_absorption || !s->mat_smoothness)) return fail(c, PTB_ERR_INVALID, "material arrays are NULL");
    if (s->n_mat > PTB_MAX_MATERIALS) return fail(c, PTB_ERR_LIMIT, "%d materials > PTB_MAX_MATERIALS=%d", s->n_mat, PTB_MAX_MATERIALS);
    CK(c, cudaSetDevice(c->device));

    // ---- host side, into temporaries
    WorldBuild wb;
    int rc = build_world(c, s, wb, mesh_generation_matches(c, s) ? &c->mesh_bounds_keep : nullptr);
    if (rc) return rc;
    HostScene nhs;
    std::vector<Obj64> w64;
    build_device_tables(wb, s->n_mat, nhs, w64);
    MeshStage ms;
    if ((rc = build_mesh_accel(c, s, wb, ms))) return rc;
    for (int k = 0; k < 4; k++) { nhs.hdr.mesh_c[k] = ms.mesh_c[k]; nhs.hdr.mesh_h[k] = ms.mesh_h[k]; }
    nhs.hdr.sky.kind = s->sky.kind == PTB_SKY_GRADIENT ? PTB_SKY_GRADIENT : PTB_SKY_CONST;
    for (int k = 0; k < 3; k++) { nhs.hdr.sky.color[k] = (float)s->sky.color[k]; nhs.hdr.sky.horizon[k] = (float)s->sky.horizon[k]; nhs.hdr.sky.zenith[k] = (float)s->sky.zenith[k]; }

    // ---- device side.  The global copies are grow-only buffers (an interactive host uploads before every frame;
    // cudaMalloc/cudaFree per upload cost more than a preview frame).  A frame still in flight on some stream reads
    // them: wait for the device before the first byte changes.  From here on a failure leaves the context WITHOUT a
    // scene (never with a mixture of two).
    const size_t words = (size_t)nhs.hdr.n_obj * 2 + (size_t)nhs.hdr.n_mat * 3;
    cudaError_t se = cudaDeviceSynchronize();
    if (se != cudaSuccess) { ms.drop(); return fail(c, PTB_ERR_CUDA, "cudaDeviceSynchronize: %s", cudaGetErrorString(se)); }
    c->has_scene = false;
    auto bail = [&](int code) { ms.drop(); return code; };
    auto put = [&](void* dst, const void* src, size_t bytes) -> int {
        if (!bytes) return PTB_OK;
        cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
        return e == cudaSuccess ? PTB_OK : fail(c, PTB_ERR_CUDA, "scene upload: %s", cudaGetErrorString(e));
    };
    if ((rc = ensure(c, (void**)&c->d_blob, &c->blob_cap, std::max<size_t>(words, 1) * sizeof(uint4)))) return bail(rc);
    if ((rc = put(c->d_blob, nhs.obj.data(), nhs.obj.size() * sizeof(DevObj)))) return bail(rc);
    if ((rc = put(c->d_blob + (size_t)nhs.hdr.n_obj * 2, nhs.mat.data(), nhs.mat.size() * sizeof(DevMat)))) return bail(rc);
    if ((rc = ensure(c, (void**)&c->d_diel, &c->diel_cap, std::max<size_t>(nhs.diel_idx.size(), 1) * sizeof(int32_t)))) return bail(rc);
    if ((rc = put(c->d_diel, nhs.diel_idx.data(), nhs.diel_idx.size() * sizeof(int32_t)))) return bail(rc);
    if (nhs.big) {
        if ((rc = ensure(c, (void**)&c->d_tab, &c->tab_cap, std::max<size_t>(nhs.scan_tab.size(), 4) * sizeof(float)))) return bail(rc);
        if ((rc = put(c->d_tab, nhs.scan_tab.data(), nhs.scan_tab.size() * sizeof(float)))) return bail(rc);
    }
    if ((rc = ensure(c, (void**)&c->d_world64, &c->world64_cap, sizeof(Obj64) * (w64.size() + 1)))) return bail(rc);
    if ((rc = put(c->d_world64, w64.data(), sizeof(Obj64) * w64.size()))) return bail(rc);

    // ---- commit
    c->blob_words = words;
    c->n_world64 = (int)w64.size();
    c->hs = std::move(nhs);
    if (ms.fresh_nodes) {                    // a new BVH replaces the kept one (nothing reads the old one: the device is idle)
        cudaFree(c->d_bvh_nodes_keep); cudaFree(c->d_bvh_tris_keep);
        c->d_bvh_nodes_keep = ms.fresh_nodes; c->d_bvh_tris_keep = ms.fresh_tris; c->bvh_keep = ms.info;
        ms.fresh_nodes = nullptr; ms.fresh_tris = nullptr;
    }
    if (ms.nodes) { c->bvh_key = ms.key; c->bvh_tag_key = ms.tag_key; c->bvh_generation = ms.generation; c->mesh_bounds_keep = ms.bounds; }
    c->d_bvh_nodes = ms.nodes; c->d_bvh_tris = ms.tris; c->bvh = ms.info;
    c->world = std::move(wb.world);
    c->cam = s->camera;
    c->has_scene = true;
    return PTB_OK;
}

int ptb_world_size(ptb_ctx* c) { return c && c->has_scene ? (int)c->world.size() : PTB_ERR_NO_SCENE; }

int ptb_world_get(ptb_ctx* c, int i, double out[19]) {
    if (!c || !out) return PTB_ERR_INVALID;
    if (!c->has_scene) return fail(c, PTB_ERR_NO_SCENE, "no scene uploaded");
    if (i < 0 || i >= (int)c->world.size()) return fail(c, PTB_ERR_INVALID, "index out of range");
    const World64Entry& w = c->world[i];
    int k = 0;
    out[k++] = w.type; out[k++] = w.mat_type;
    for (int j = 0; j < 3; j++) out[k++] = w.a[j];
    for (int j = 0; j < 3; j++) out[k++] = w.b[j];
    for (int j = 0; j < 3; j++) out[k++] = w.albedo[j];
    out[k++] = w.rough; out[k++] = w.ior;
    for (int j = 0; j < 3; j++) out[k++] = w.emit[j];
    for (int j = 0; j < 3; j++) out[k++] = w.absorption[j];
    return PTB_OK;
}

int ptb_render_accum_device(ptb_ctx* c, const ptb_cfg* cfg, void* d_rgb_sum, void* stream) {
    if (!c) return PTB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    int s0 = 0, s1 = 0, rc;
    if ((rc = check_cfg(c, cfg, s0, s1))) return rc;
    if (!d_rgb_sum) return fail(c, PTB_ERR_INVALID, "d_rgb_sum is NULL");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = render_launch(c, cfg, make_plan(c, cfg, s0, s1), (float*)d_rgb_sum, nullptr, st))) return rc;
    if (cfg->flags & PTB_FLAG_STATS) { CK(c, cudaStreamSynchronize(st)); return fetch_stats(c, true, 0.f); }
    return PTB_OK;
}

int ptb_render_device(ptb_ctx* c, const ptb_cfg* cfg, void* d_rgba, void* stream) {
    if (!c) return PTB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    int s0 = 0, s1 = 0, rc;
    if ((rc = check_cfg(c, cfg, s0, s1))) return rc;
    if (!d_rgba) return fail(c, PTB_ERR_INVALID, "d_rgba is NULL");
    if (s0 != 0 || s1 != cfg->samples_per_px) return fail(c, PTB_ERR_INVALID, "ptb_render_device needs the full sample range");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = render_launch(c, cfg, make_plan(c, cfg, s0, s1), nullptr, (uint8_t*)d_rgba, st))) return rc;
    if (cfg->flags & PTB_FLAG_STATS) { CK(c, cudaStreamSynchronize(st)); return fetch_stats(c, true, 0.f); }
    return PTB_OK;
}

int ptb_finalize_device(ptb_ctx* c, const void* d_rgb_sum, int32_t width, int32_t height, int32_t spp_total, void* d_rgba, void* stream) {
    if (!c) return PTB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!d_rgb_sum || !d_rgba || width < 1 || height < 1 || spp_total < 1) return fail(c, PTB_ERR_INVALID, "bad argument");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    int e = launch_finalize((const float*)d_rgb_sum, width, height, spp_total, (uint8_t*)d_rgba, st);
    if (e) return fail(c, PTB_ERR_CUDA, "finalize launch: %s", cudaGetErrorString((cudaError_t)e));
    return PTB_OK;
}

int ptb_finalize_host(ptb_ctx* c, const float* rgb_sum, int32_t width, int32_t height, int32_t spp_total, uint8_t* rgba, size_t stride) {
    if (!c) return PTB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!rgb_sum || !rgba || width < 1 || height < 1 || spp_total < 1 || stride < (size_t)width * 4) return fail(c, PTB_ERR_INVALID, "bad argument");
    CK(c, cudaSetDevice(c->device));
    const size_t acc_bytes = (size_t)width * height * 3 * sizeof(float), img_bytes = (size_t)width * height * 4;
    int rc;
    if ((rc = ensure(c, (void**)&c->d_accum, &c->accum_cap, acc_bytes))) return rc;
    if ((rc = ensure(c, (void**)&c->d_rgba, &c->rgba_cap, img_bytes))) return rc;
    if ((rc = ensure(c, (void**)&c->h_rgba, &c->h_rgba_cap, img_bytes, true))) return rc;
    CK(c, cudaMemcpyAsync(c->d_accum, rgb_sum, acc_bytes, cudaMemcpyHostToDevice, c->stream));
    int e = launch_finalize(c->d_accum, width, height, spp_total, c->d_rgba, c->stream);
    if (e) return fail(c, PTB_ERR_CUDA, "finalize launch: %s", cudaGetErrorString((cudaError_t)e));
    return copy_image_out(c, c->d_rgba, c->h_rgba, rgba, stride, width, height, c->stream);
}

int ptb_render_accum(ptb_ctx* c, const ptb_cfg* cfg, float* rgb_sum) {
    if (!c) return PTB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    int s0 = 0, s1 = 0, rc;
    if ((rc = check_cfg(c, cfg, s0, s1))) return rc;
    if (!rgb_sum) return fail(c, PTB_ERR_INVALID, "rgb_sum is NULL");
    CK(c, cudaSetDevice(c->device));
    const size_t bytes = (size_t)cfg->width * rows_of(cfg) * 3 * sizeof(float);
    if ((rc = ensure(c, (void**)&c->d_accum, &c->accum_cap, bytes))) return rc;
    CK(c, cudaEventRecord(c->ev0, c->stream));
    if ((rc = render_launch(c, cfg, make_plan(c, cfg, s0, s1), c->d_accum, nullptr, c->stream))) return rc;
    CK(c, cudaEventRecord(c->ev1, c->stream));
    CK(c, cudaMemcpyAsync(rgb_sum, c->d_accum, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    return fetch_stats(c, (cfg->flags & PTB_FLAG_STATS) != 0, ms);
}

}  // extern "C"

namespace {
// Is the caller's image page-locked (cudaHostRegister / cudaHostAlloc, e.g. through ptb_host_buffer_pin)?  Then the
// device-to-host copy lands in it directly; otherwise it goes through the context's pinned staging buffer.
bool host_range_pinned(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// Image of W x H RGBA8 pixels from device memory into the caller's (strided) host image, on `stream`; `staging` is
// pinned host memory of W*H*4 bytes.  Synchronises `stream` before it returns (the caller's image is complete).
int copy_image_out(ptb_ctx* c, const uint8_t* d_img, uint8_t* staging, uint8_t* rgba, size_t stride, int W, int H, cudaStream_t stream) {
    const size_t row = (size_t)W * 4, img_bytes = row * H;
    if (host_range_pinned(rgba) && host_range_pinned(rgba + (size_t)(H - 1) * stride + row - 1)) {
        if (stride == row) CK(c, cudaMemcpyAsync(rgba, d_img, img_bytes, cudaMemcpyDeviceToHost, stream));
        else CK(c, cudaMemcpy2DAsync(rgba, stride, d_img, row, row, (size_t)H, cudaMemcpyDeviceToHost, stream));
        CK(c, cudaStreamSynchronize(stream));
        return PTB_OK;
    }
    CK(c, cudaMemcpyAsync(staging, d_img, img_bytes, cudaMemcpyDeviceToHost, stream));
    CK(c, cudaStreamSynchronize(stream));
    if (stride == row) std::memcpy(rgba, staging, img_bytes);
    else for (int y = 0; y < H; y++) std::memcpy(rgba + (size_t)y * stride, staging + (size_t)y * row, row);
    return PTB_OK;
}
}  // namespace

extern "C" {

int ptb_host_buffer_pin(ptb_ctx* c, void* p, size_t bytes) {
    if (!c || !p || !bytes) return fail(c, PTB_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return PTB_OK;
}
int ptb_host_buffer_unpin(ptb_ctx* c, void* p) {
    if (!c || !p) return fail(c, PTB_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaHostUnregister(p));
    return PTB_OK;
}

int ptb_render(ptb_ctx* c, const ptb_cfg* cfg, uint8_t* rgba, size_t stride, ptb_progress_fn progress, void* user) {
    if (!c) return PTB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    int s0 = 0, s1 = 0, rc;
    if ((rc = check_cfg(c, cfg, s0, s1))) return rc;
    if (!rgba) return fail(c, PTB_ERR_INVALID, "rgba is NULL");
    const int W = cfg->width, H = rows_of(cfg);       // (row partition: H compact rows)
    if (stride < (size_t)W * 4) return fail(c, PTB_ERR_INVALID, "stride %zu < 4*width", stride);
    if (s0 != 0 || s1 != cfg->samples_per_px) return fail(c, PTB_ERR_INVALID, "ptb_render needs the full sample range (use ptb_render_accum*)");
    CK(c, cudaSetDevice(c->device));
    const size_t img_bytes = (size_t)W * H * 4;
    if ((rc = ensure(c, (void**)&c->d_rgba, &c->rgba_cap, img_bytes))) return rc;
    if ((rc = ensure(c, (void**)&c->h_rgba, &c->h_rgba_cap, img_bytes, true))) return rc;
    const SplitPlan plan = make_plan(c, cfg, s0, s1);
    CK(c, cudaEventRecord(c->ev0, c->stream));
    if (!progress || cfg->max_depth <= 0) {
        // one fused launch: integrate + pixel epilogue, 4 bytes per pixel written once
        if ((rc = render_launch(c, cfg, plan, nullptr, c->d_rgba, c->stream))) return rc;
        CK(c, cudaEventRecord(c->ev1, c->stream));
        if ((rc = copy_image_out(c, c->d_rgba, c->h_rgba, rgba, stride, W, H, c->stream))) return rc;
        if (progress) progress(user);                      // (a black frame: max_depth <= 0)
    } else {
        // Progressive: ~10 sample batches, image refreshed and progress() called after each — the cadence of the
        // reference back-ends (every ~5 % of tiles renderer.go:226-235; every spp/10 passes gpu.go:2209-2229).  Partial
        // images show the mean of the samples so far; the fp32 sums are carried in d_accum, and every batch adds exactly
        // the partial sums a single launch adds (whole planes of the same plan), so the final image equals the
        // non-progressive one byte for byte.
        // Asynchronous read-back (gpu.go:2229-2293 reads back through a PBO): batch b + 1 is queued before the host waits
        // for image b, whose copy runs on a second stream into the other of two image buffers — the GPU integrates while
        // the host copies the image out and runs the callback.
        const size_t acc_bytes = (size_t)W * H * 3 * sizeof(float);
        if ((rc = ensure(c, (void**)&c->d_accum, &c->accum_cap, acc_bytes))) return rc;
        if ((rc = ensure(c, (void**)&c->d_rgba2, &c->rgba2_cap, img_bytes))) return rc;
        if ((rc = ensure(c, (void**)&c->h_rgba2, &c->h_rgba2_cap, img_bytes, true))) return rc;
        uint8_t* d_img[2] = {c->d_rgba, c->d_rgba2};
        uint8_t* h_img[2] = {c->h_rgba, c->h_rgba2};
        const int spp = cfg->samples_per_px;
        // batch boundaries: in planes when the frame is split, in samples otherwise
        const int units = plan.total > 1 ? plan.total : spp;
        const int per = units >= 10 ? (units + 9) / 10 : 1;
        const int n_batches = (units + per - 1) / per;
        ptb_cfg sub = *cfg;
        sub.flags &= ~PTB_FLAG_STATS;                      // (counters are not collected in progressive renders; ptb200.h)
        auto launch_batch = [&](int b) -> int {
            const int u0 = b * per, u1 = std::min(units, u0 + per);
            SplitPlan bp = plan;
            if (plan.total > 1) { bp.base = u0; bp.count = u1 - u0; }
            else { bp.total = 1; bp.base = 0; bp.count = 1; bp.full_begin = u0; bp.full_end = u1; }
            sub.samples_per_px = plan.total > 1 ? plan_bound(plan, u1) : u1;     // epilogue divisor: samples so far
            if (b >= 2) CK(c, cudaStreamWaitEvent(c->stream, c->ev_copy[b & 1], 0));   // image buffer b & 1 has been read back
            int r = render_launch(c, &sub, bp, c->d_accum, d_img[b & 1], c->stream, /*resume=*/b > 0);
            if (r) return r;
            CK(c, cudaEventRecord(c->ev_batch[b & 1], c->stream));
            if (b == n_batches - 1) CK(c, cudaEventRecord(c->ev1, c->stream));
            return PTB_OK;
        };
        if ((rc = launch_batch(0))) return rc;
        for (int b = 0; b < n_batches; b++) {
            if (b + 1 < n_batches && (rc = launch_batch(b + 1))) return rc;          // keep the GPU busy while image b is handled
            CK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_batch[b & 1], 0));      // image b is complete; batch b + 1 runs on
            if ((rc = copy_image_out(c, d_img[b & 1], h_img[b & 1], rgba, stride, W, H, c->copy_stream))) return rc;
            CK(c, cudaEventRecord(c->ev_copy[b & 1], c->copy_stream));
            progress(user);
        }
        CK(c, cudaStreamSynchronize(c->stream));
        progress(user);   // final refresh, renderer.go:243-245
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    return fetch_stats(c, (cfg->flags & PTB_FLAG_STATS) != 0 && !progress, ms);
}

// Checkpointable rendering: continue a render from fp32 sums held by the caller.
int ptb_render_resume(ptb_ctx* c, const ptb_cfg* cfg, float* rgb_sum, uint8_t* rgba, size_t stride) {
    if (!c) return PTB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    int s0 = 0, s1 = 0, rc;
    if ((rc = check_cfg(c, cfg, s0, s1))) return rc;
    if (!rgb_sum) return fail(c, PTB_ERR_INVALID, "rgb_sum is NULL");
    if (cfg->row_step > 1) return fail(c, PTB_ERR_INVALID, "ptb_render_resume does not take a row partition");
    const int W = cfg->width, H = cfg->height;
    if (rgba && stride < (size_t)W * 4) return fail(c, PTB_ERR_INVALID, "stride %zu < 4*width", stride);
    CK(c, cudaSetDevice(c->device));
    const size_t acc_bytes = (size_t)W * H * 3 * sizeof(float), img_bytes = (size_t)W * H * 4;
    if ((rc = ensure(c, (void**)&c->d_accum, &c->accum_cap, acc_bytes))) return rc;
    if (rgba) {
        if ((rc = ensure(c, (void**)&c->d_rgba, &c->rgba_cap, img_bytes))) return rc;
        if ((rc = ensure(c, (void**)&c->h_rgba, &c->h_rgba_cap, img_bytes, true))) return rc;
    }
    const bool resume = s0 > 0;
    if (resume) CK(c, cudaMemcpyAsync(c->d_accum, rgb_sum, acc_bytes, cudaMemcpyHostToDevice, c->stream));
    // whole pixels, strictly sequential sums: any partition of [0, spp) into consecutive calls gives the same bits
    SplitPlan plan; plan.full_begin = s0; plan.full_end = s1;
    ptb_cfg sub = *cfg;
    sub.samples_per_px = s1;                               // epilogue divisor: the samples accumulated so far
    CK(c, cudaEventRecord(c->ev0, c->stream));
    if ((rc = render_launch(c, &sub, plan, c->d_accum, rgba ? c->d_rgba : nullptr, c->stream, resume))) return rc;
    CK(c, cudaEventRecord(c->ev1, c->stream));
    CK(c, cudaMemcpyAsync(rgb_sum, c->d_accum, acc_bytes, cudaMemcpyDeviceToHost, c->stream));
    if (rgba) { if ((rc = copy_image_out(c, c->d_rgba, c->h_rgba, rgba, stride, W, H, c->stream))) return rc; }
    else CK(c, cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    return fetch_stats(c, (cfg->flags & PTB_FLAG_STATS) != 0, ms);
}

int ptb_primary_hits(ptb_ctx* c, const ptb_cfg* cfg, double xi_u, double xi_v, int32_t* ids, double* t) {
    if (!c) return PTB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!cfg || cfg->width < 2 || cfg->height < 2) return fail(c, PTB_ERR_INVALID, "width and height must be >= 2");
    if (!c->has_scene) return fail(c, PTB_ERR_NO_SCENE, "no scene uploaded");
    if (!ids || !t) return fail(c, PTB_ERR_INVALID, "ids/t is NULL");
    CK(c, cudaSetDevice(c->device));
    const int W = cfg->width, H = cfg->height;
    const size_t n = (size_t)W * H;
    int32_t* d_ids = nullptr; double* d_t = nullptr;
    CK(c, cudaMalloc((void**)&d_ids, n * sizeof(int32_t)));
    if (cudaMalloc((void**)&d_t, n * sizeof(double)) != cudaSuccess) { cudaFree(d_ids); return fail(c, PTB_ERR_CUDA, "cudaMalloc failed"); }
    Cam64 cam = new_camera(c->cam, W, H);
    Camera64 dc;
    for (int k = 0; k < 3; k++) { dc.origin[k] = cam.origin[k]; dc.llc[k] = cam.llc[k]; dc.horizontal[k] = cam.horizontal[k]; dc.vertical[k] = cam.vertical[k]; }
    int e = launch_primary_hits(c->d_world64, c->n_world64, (const float4*)c->d_bvh_nodes, (const float4*)c->d_bvh_tris, dc, W, H, xi_u, xi_v,
                                d_ids, d_t, c->stream);
    cudaError_t ce = (cudaError_t)e;
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(ids, d_ids, n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(t, d_t, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(c->stream);
    cudaFree(d_ids); cudaFree(d_t);
    if (ce != cudaSuccess) return fail(c, PTB_ERR_CUDA, "primary hits: %s", cudaGetErrorString(ce));
    return PTB_OK;
}

int ptb_get_bvh_info(ptb_ctx* c, ptb_bvh_info* out) {
    if (!c || !out) return PTB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->has_scene) return fail(c, PTB_ERR_NO_SCENE, "no scene uploaded");
    *out = c->bvh;
    return PTB_OK;
}

const char* ptb_last_kernel(ptb_ctx* c) {
    if (!c) return "";
    std::lock_guard<std::mutex> lk(c->mu);
    return c->last_kernel.c_str();
}

int ptb_get_stats(ptb_ctx* c, ptb_stats* out) {
    if (!c || !out) return PTB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    *out = c->stats;
    return PTB_OK;
}

// ------------------------------------------------------------------ single-process multi-GPU
}  // extern "C"

struct ptb_multi {
    std::vector<ptb_ctx*> ctx;
    std::string err;
    std::mutex mu;
    std::vector<float*> d_accum;         // per device, W*H*3 floats
    size_t accum_cap = 0;
    const float** d_ptrs = nullptr;      // device-0 array of the n buffer pointers
    std::vector<cudaEvent_t> done, begin;
    double render_ms = 0, reduce_ms = 0;
};
namespace {
thread_local std::string g_multi_error;
int mfail(ptb_multi* m, int code, const std::string& msg) { if (m) m->err = msg; else g_multi_error = msg; return code; }
}  // namespace

extern "C" {

int ptb_multi_create(const int* devices, int n, ptb_multi** out) {
    if (!out || n < 1) return mfail(nullptr, PTB_ERR_INVALID, "bad argument");
    *out = nullptr;
    ptb_multi* m = new ptb_multi();
    for (int k = 0; k < n; k++) {
        ptb_ctx* c = nullptr;
        int rc = ptb_create(devices ? devices[k] : k, &c);
        if (rc != PTB_OK) { std::string e = ptb_last_error(nullptr); ptb_multi_destroy(m); return mfail(nullptr, rc, e); }
        m->ctx.push_back(c);
    }
    // peer access: device 0 reads every other device's accumulation buffer
    cudaSetDevice(m->ctx[0]->device);
    for (int k = 1; k < n; k++) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, m->ctx[0]->device, m->ctx[k]->device);
        if (!can) { ptb_multi_destroy(m); return mfail(nullptr, PTB_ERR_CUDA, "device " + std::to_string(m->ctx[0]->device) + " cannot access device " + std::to_string(m->ctx[k]->device) + " as a peer"); }
        cudaError_t e = cudaDeviceEnablePeerAccess(m->ctx[k]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { ptb_multi_destroy(m); return mfail(nullptr, PTB_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); }
        cudaGetLastError();
    }
    m->d_accum.assign(n, nullptr);
    m->done.assign(n, nullptr); m->begin.assign(n, nullptr);
    for (int k = 0; k < n; k++) { cudaSetDevice(m->ctx[k]->device); cudaEventCreate(&m->done[k]); cudaEventCreate(&m->begin[k]); }
    *out = m;
    return PTB_OK;
}

void ptb_multi_destroy(ptb_multi* m) {
    if (!m) return;
    for (size_t k = 0; k < m->ctx.size(); k++) {
        cudaSetDevice(m->ctx[k]->device);
        if (k < m->d_accum.size()) cudaFree(m->d_accum[k]);
        if (k < m->done.size() && m->done[k]) cudaEventDestroy(m->done[k]);
        if (k < m->begin.size() && m->begin[k]) cudaEventDestroy(m->begin[k]);
    }
    if (!m->ctx.empty()) { cudaSetDevice(m->ctx[0]->device); cudaFree((void*)m->d_ptrs); }
    for (ptb_ctx* c : m->ctx) ptb_destroy(c);
    delete m;
}

const char* ptb_multi_last_error(const ptb_multi* m) { return m ? m->err.c_str() : g_multi_error.c_str(); }

int ptb_multi_scene_upload(ptb_multi* m, const ptb_scene* s) {
    if (!m) return PTB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(m->mu);
    for (ptb_ctx* c : m->ctx) { int rc = ptb_scene_upload(c, s); if (rc != PTB_OK) return mfail(m, rc, ptb_last_error(c)); }
    return PTB_OK;
}

int ptb_multi_render(ptb_multi* m, const ptb_cfg* cfg, uint8_t* rgba, size_t stride) {
    if (!m || !cfg || !rgba) return mfail(m, PTB_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(m->mu);
    const int n = (int)m->ctx.size(), W = cfg->width, H = cfg->height, spp = cfg->samples_per_px;
    if (W < 2 || H < 2 || spp < 1) return mfail(m, PTB_ERR_INVALID, "bad frame configuration");
    if (cfg->row_step > 1 || cfg->row_offset != 0) return mfail(m, PTB_ERR_INVALID, "ptb_multi_render partitions the samples itself: no row partition");
    if (stride < (size_t)W * 4) return mfail(m, PTB_ERR_INVALID, "stride < 4*width");
    const size_t bytes = (size_t)W * H * 3 * sizeof(float);
    if (m->accum_cap < bytes) {
        // any failure below leaves the object with NO buffers (accum_cap = 0), never with stale or freed ones
        auto drop_all = [&]() {
            for (int k = 0; k < n; k++) { cudaSetDevice(m->ctx[k]->device); cudaFree(m->d_accum[k]); m->d_accum[k] = nullptr; }
            m->accum_cap = 0;
        };
        drop_all();
        for (int k = 0; k < n; k++) {
            cudaSetDevice(m->ctx[k]->device);
            cudaError_t e = cudaMalloc((void**)&m->d_accum[k], bytes);
            if (e != cudaSuccess) { m->d_accum[k] = nullptr; drop_all(); return mfail(m, PTB_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
        }
        cudaSetDevice(m->ctx[0]->device);
        if (!m->d_ptrs) { cudaError_t e = cudaMalloc((void**)&m->d_ptrs, sizeof(float*) * n); if (e != cudaSuccess) { m->d_ptrs = nullptr; drop_all(); return mfail(m, PTB_ERR_CUDA, "cudaMalloc failed"); } }
        cudaError_t e = cudaMemcpy((void*)m->d_ptrs, m->d_accum.data(), sizeof(float*) * n, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { drop_all(); return mfail(m, PTB_ERR_CUDA, std::string("cudaMemcpy: ") + cudaGetErrorString(e)); }
        m->accum_cap = bytes;
    }
    // every device traces its sample range.  The analytic integrator is one asynchronous launch per device, so issuing them
    // one after the other already runs the devices concurrently; the mesh pipeline drives its device from the host until the
    // frame is done, so mesh scenes get one host thread per device.
    std::vector<int> rcs(n, PTB_OK);
    auto work = [&](int k) {
        const int base = spp / n, rem = spp % n;
        const int b = k * base + (k < rem ? k : rem), cnt = base + (k < rem ? 1 : 0);
        ptb_ctx* c = m->ctx[k];
        cudaSetDevice(c->device);
        cudaEventRecord(m->begin[k], c->stream);
        if (cnt > 0) {
            ptb_cfg sub = *cfg;
            sub.sample_begin = b; sub.sample_count = cnt;
            sub.flags &= ~PTB_FLAG_STATS;
            rcs[k] = ptb_render_accum_device(c, &sub, m->d_accum[k], c->stream);
        } else {
            cudaMemsetAsync(m->d_accum[k], 0, bytes, c->stream);
        }
        cudaEventRecord(m->done[k], c->stream);
    };
    if (n > 1 && m->ctx[0]->d_bvh_nodes) {
        std::vector<std::thread> pool;
        for (int k = 0; k < n; k++) pool.emplace_back(work, k);
        for (auto& t : pool) t.join();
    } else {
        for (int k = 0; k < n; k++) work(k);
    }
    for (int k = 0; k < n; k++) if (rcs[k] != PTB_OK) return mfail(m, rcs[k], ptb_last_error(m->ctx[k]));
    // device 0: wait for everybody, then the fused reduce + epilogue over peer memory, then read back
    ptb_ctx* c0 = m->ctx[0];
    cudaSetDevice(c0->device);
    const size_t img_bytes = (size_t)W * H * 4;
    int rc;
    if ((rc = ensure(c0, (void**)&c0->d_rgba, &c0->rgba_cap, img_bytes))) return mfail(m, rc, ptb_last_error(c0));
    if ((rc = ensure(c0, (void**)&c0->h_rgba, &c0->h_rgba_cap, img_bytes, true))) return mfail(m, rc, ptb_last_error(c0));
    for (int k = 1; k < n; k++) cudaStreamWaitEvent(c0->stream, m->done[k], 0);
    cudaEventRecord(c0->ev0, c0->stream);
    int e = launch_finalize_peers((const float* const*)m->d_ptrs, n, W, H, spp, c0->d_rgba, c0->stream);
    if (e) return mfail(m, PTB_ERR_CUDA, std::string("finalize_peers launch: ") + cudaGetErrorString((cudaError_t)e));
    cudaEventRecord(c0->ev1, c0->stream);
    if ((rc = copy_image_out(c0, c0->d_rgba, c0->h_rgba, rgba, stride, W, H, c0->stream))) return mfail(m, rc, ptb_last_error(c0));
    float ms = 0;
    cudaEventElapsedTime(&ms, c0->ev0, c0->ev1);
    m->reduce_ms = ms;
    m->render_ms = 0;
    for (int k = 0; k < n; k++) {
        cudaSetDevice(m->ctx[k]->device);
        float t = 0;
        if (cudaEventElapsedTime(&t, m->begin[k], m->done[k]) == cudaSuccess && t > m->render_ms) m->render_ms = t;
    }
    return PTB_OK;
}

int ptb_multi_last_timing(ptb_multi* m, double* render_ms, double* reduce_ms) {
    if (!m) return PTB_ERR_INVALID;
    if (render_ms) *render_ms = m->render_ms;
    if (reduce_ms) *reduce_ms = m->reduce_ms;
    return PTB_OK;
}

// ------------------------------------------------------------------ multi-process peer group (one rank per GPU)
}  // extern "C"

struct ptb_peer {
    ptb_ctx* ctx = nullptr;
    int rank = 0, world = 1;
    size_t max_pix = 0;
    float* d_accum = nullptr;                 // this rank's sums (exported)
    uint8_t* d_image = nullptr;               // rank 0: the assembled RGBA8 image (exported)
    std::vector<const float*> accum_of;       // per rank: this process's mapping of that rank's buffer
    uint8_t* root_image = nullptr;            // this process's mapping of rank 0's image
    std::vector<void*> opened;                // cudaIpcOpenMemHandle mappings to close
    const float** d_ptrs = nullptr;           // device copy of accum_of
    unsigned* d_flags = nullptr;              // this rank's flag block (exported): 2 x world words + [2 world] CTA counter + [2 world + 1] error
    PeerSync sync{};                          // flags_of[k]: mapping of rank k's flag block
    bool connected = false;
};

extern "C" {

int ptb_peer_create(ptb_ctx* c, int rank, int world, int32_t max_width, int32_t max_height, ptb_peer** out) {
    if (!c || !out || world < 1 || rank < 0 || rank >= world || max_width < 2 || max_height < 2) return fail(c, PTB_ERR_INVALID, "bad argument");
    *out = nullptr;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(c, cudaSetDevice(c->device));
    auto* p = new ptb_peer();
    p->ctx = c; p->rank = rank; p->world = world;
    p->max_pix = ((size_t)max_width * max_height + 3) / 4 * 4;
    cudaError_t e = cudaMalloc((void**)&p->d_accum, 2 * p->max_pix * 3 * sizeof(float));      // two buffers: frame parity
    if (e == cudaSuccess && rank == 0) e = cudaMalloc((void**)&p->d_image, p->max_pix * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_ptrs, sizeof(float*) * world);
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_flags, sizeof(unsigned) * (2 * kMaxPeers + 8));
    if (e == cudaSuccess) e = cudaMemset(p->d_flags, 0, sizeof(unsigned) * (2 * kMaxPeers + 8));
    if (e != cudaSuccess || world > kMaxPeers) {
        cudaFree(p->d_accum); cudaFree(p->d_image); cudaFree((void*)p->d_ptrs); cudaFree(p->d_flags); delete p;
        return world > kMaxPeers ? fail(c, PTB_ERR_LIMIT, "at most %d ranks per peer group", kMaxPeers) : fail(c, PTB_ERR_CUDA, "ptb_peer_create: %s", cudaGetErrorString(e));
    }
    p->sync.rank = rank; p->sync.world = world; p->sync.seq = 0;
    p->sync.flags_of[rank] = p->d_flags;
    p->sync.block_counter = p->d_flags + 2 * kMaxPeers; p->sync.error = p->d_flags + 2 * kMaxPeers + 1;
    p->accum_of.assign(world, nullptr);
    p->accum_of[rank] = p->d_accum;
    if (rank == 0) p->root_image = p->d_image;
    if (world == 1) {
        cudaMemcpy((void*)p->d_ptrs, p->accum_of.data(), sizeof(float*), cudaMemcpyHostToDevice);
        p->connected = true;
    }
    *out = p;
    return PTB_OK;
}

void ptb_peer_destroy(ptb_peer* p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    cudaDeviceSynchronize();
    for (void* m : p->opened) cudaIpcCloseMemHandle(m);
    cudaFree(p->d_accum); cudaFree(p->d_image); cudaFree((void*)p->d_ptrs); cudaFree(p->d_flags);
    delete p;
}

int ptb_peer_handles(ptb_peer* p, unsigned char accum_handle[PTB_IPC_HANDLE_BYTES], unsigned char image_handle[PTB_IPC_HANDLE_BYTES],
                     unsigned char flags_handle[PTB_IPC_HANDLE_BYTES]) {
    if (!p || !accum_handle || !image_handle || !flags_handle) return PTB_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == PTB_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
    ptb_ctx* c = p->ctx;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    CK(c, cudaIpcGetMemHandle(&h, p->d_accum));
    std::memcpy(accum_handle, &h, sizeof h);
    std::memset(image_handle, 0, PTB_IPC_HANDLE_BYTES);
    if (p->rank == 0) { CK(c, cudaIpcGetMemHandle(&h, p->d_image)); std::memcpy(image_handle, &h, sizeof h); }
    CK(c, cudaIpcGetMemHandle(&h, p->d_flags));
    std::memcpy(flags_handle, &h, sizeof h);
    return PTB_OK;
}

int ptb_peer_connect(ptb_peer* p, const unsigned char* accum_handles, const unsigned char* root_image_handle, const unsigned char* flags_handles) {
    if (!p || !accum_handles || !root_image_handle || !flags_handles) return PTB_ERR_INVALID;
    ptb_ctx* c = p->ctx;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(c, cudaSetDevice(c->device));
    if (p->connected) return fail(c, PTB_ERR_INVALID, "peer group is already connected");
    for (int k = 0; k < p->world; k++) {
        if (k == p->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, accum_handles + (size_t)k * PTB_IPC_HANDLE_BYTES, sizeof h);
        void* m = nullptr;
        CK(c, cudaIpcOpenMemHandle(&m, h, cudaIpcMemLazyEnablePeerAccess));
        p->opened.push_back(m);
        p->accum_of[k] = (const float*)m;
        std::memcpy(&h, flags_handles + (size_t)k * PTB_IPC_HANDLE_BYTES, sizeof h);
        m = nullptr;
        CK(c, cudaIpcOpenMemHandle(&m, h, cudaIpcMemLazyEnablePeerAccess));
        p->opened.push_back(m);
        p->sync.flags_of[k] = (unsigned*)m;
    }
    if (p->rank != 0) {
        cudaIpcMemHandle_t h;
        std::memcpy(&h, root_image_handle, sizeof h);
        void* m = nullptr;
        CK(c, cudaIpcOpenMemHandle(&m, h, cudaIpcMemLazyEnablePeerAccess));
        p->opened.push_back(m);
        p->root_image = (uint8_t*)m;
    }
    CK(c, cudaMemcpy((void*)p->d_ptrs, p->accum_of.data(), sizeof(float*) * p->world, cudaMemcpyHostToDevice));
    p->connected = true;
    return PTB_OK;
}

void* ptb_peer_accum(ptb_peer* p) {     // the buffer of the NEXT frame (frames alternate between the two halves of the allocation)
    return p ? p->d_accum + (size_t)((p->sync.seq + 1u) & 1u) * p->max_pix * 3 : nullptr;
}
void* ptb_peer_image(ptb_peer* p) { return p ? p->d_image : nullptr; }

int ptb_peer_slice(const ptb_peer* p, int32_t width, int32_t height, int64_t* begin, int64_t* end) {
    if (!p || width < 1 || height < 1) return PTB_ERR_INVALID;
    const long long n_pix = (long long)width * height;
    const long long chunk = ((n_pix + p->world - 1) / p->world + 3) / 4 * 4;
    const long long b = std::min(n_pix, chunk * p->rank), e = std::min(n_pix, chunk * (p->rank + 1));
    if (begin) *begin = b;
    if (end) *end = e;
    return PTB_OK;
}

int ptb_peer_reduce_finalize(ptb_peer* p, int32_t width, int32_t height, int32_t spp_total, void* stream) {
    if (!p) return PTB_ERR_INVALID;
    ptb_ctx* c = p->ctx;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!p->connected) return fail(c, PTB_ERR_INVALID, "peer group is not connected (ptb_peer_connect)");
    if (width < 1 || height < 1 || spp_total < 1 || (size_t)width * height > p->max_pix) return fail(c, PTB_ERR_INVALID, "frame does not fit the peer buffers");
    CK(c, cudaSetDevice(c->device));
    int64_t b = 0, e = 0;
    ptb_peer_slice(p, width, height, &b, &e);
    p->sync.seq += 1;                      // every rank calls this once per frame, in the same order
    const size_t buf_off = (size_t)(p->sync.seq & 1u) * p->max_pix * 3;
    int rc = launch_reduce_finalize_slice(p->d_ptrs, buf_off, p->world, b, e, spp_total, p->root_image, p->sync, stream);
    if (rc) return fail(c, PTB_ERR_CUDA, "reduce_finalize_slice launch: %s", cudaGetErrorString((cudaError_t)rc));
    return PTB_OK;
}

int ptb_peer_status(ptb_peer* p) {
    if (!p) return PTB_ERR_INVALID;
    ptb_ctx* c = p->ctx;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(c, cudaSetDevice(c->device));
    unsigned err = 0;
    CK(c, cudaMemcpy(&err, p->sync.error, sizeof err, cudaMemcpyDeviceToHost));
    if (err) return fail(c, PTB_ERR_CUDA, "peer exchange: a rank did not signal within the time-out");
    return PTB_OK;
}

int ptb_scene_device_order(const ptb_scene* s, int32_t* order, int32_t cap, int32_t counts[6]) {
    if (!s || (!order && cap > 0) || !counts) return fail(nullptr, PTB_ERR_INVALID, "NULL argument");
    if (s->n_obj < 0 || s->n_mat < 0 || s->n_mat > PTB_MAX_MATERIALS) return fail(nullptr, PTB_ERR_INVALID, "bad counts");
    if (s->n_obj > 0 && (!s->obj_type || !s->obj_mat || !s->obj_pos || !s->obj_size)) return fail(nullptr, PTB_ERR_INVALID, "object arrays are NULL");
    if (s->n_mat > 0 && (!s->mat_type || !s->mat_albedo || !s->mat_rough || !s->mat_ior || !s->mat_emit || !s->mat_power ||
                         !s->mat_absorption || !s->mat_smoothness)) return fail(nullptr, PTB_ERR_INVALID, "material arrays are NULL");
    WorldBuild wb;
    int rc = build_world(nullptr, s, wb);
    if (rc) return rc;
    HostScene host;
    std::vector<Obj64> w64;
    build_device_tables(wb, s->n_mat, host, w64);
    const SceneHdr& hs = host.hdr;
    for (int k = 0; k < hs.n_obj && k < cap; k++) order[k] = host.obj[k].world_idx;
    counts[0] = hs.n_box; counts[1] = hs.n_plane_run; counts[2] = hs.n_sphere_run; counts[3] = hs.n_obj - hs.n_typed;
    counts[4] = hs.n_dbox; counts[5] = hs.n_dsph;
    return hs.n_obj;
}

int ptb_launch_plan(int32_t sm_count, int64_t n_pixels, int32_t n_samples, int32_t n_spheres, int32_t* packed) {
    if (sm_count < 1 || n_pixels < 1 || n_samples < 1 || n_spheres < 0) return fail(nullptr, PTB_ERR_INVALID, "bad argument");
    if (packed) *packed = n_spheres >= kPackedSphereMin ? 1 : 0;
    return wf_split_factor(sm_count, (long long)n_pixels, n_samples);
}

int64_t ptb_bvh_selfcheck(const float* tri_vertices, int64_t n_tri, int64_t* n_nodes, int32_t* max_depth) {
    if (!tri_vertices || n_tri < 1) return PTB_ERR_INVALID;
    std::vector<int32_t> zeros((size_t)n_tri, 0);
    BvhBuildInput in{tri_vertices, n_tri, zeros.data(), zeros.data()};
    BvhBuildOutput out;
    unsigned hw = std::thread::hardware_concurrency();
    build_bvh(in, out, hw ? (int)hw : 4);
    if (n_nodes) *n_nodes = (int64_t)out.nodes.size();
    if (max_depth) *max_depth = out.max_depth;
    int64_t bad = 0;
    std::vector<char> seen((size_t)n_tri, 0);
    auto f2i = [](float f) { int32_t i; std::memcpy(&i, &f, 4); return i; };
    // walk the tree; `box` = (c, h) of the child we came through (nullptr for the root)
    struct Item { int32_t link; float c[3], h[3]; bool has_box; };
    std::vector<Item> stack;
    stack.push_back({0, {0, 0, 0}, {0, 0, 0}, false});
    auto inside = [](const float* c, const float* h, const float* p) {
        for (int k = 0; k < 3; k++) if (!((double)p[k] >= (double)c[k] - h[k] && (double)p[k] <= (double)c[k] + h[k])) return false;
        return true;
    };
    // containment is checked triangle by triangle against EVERY ancestor box: carry the chain of boxes
    std::vector<std::vector<std::pair<std::array<float, 3>, std::array<float, 3>>>> chains;
    chains.push_back({});
    std::vector<int> chain_of{0};
    while (!stack.empty()) {
        Item it = stack.back(); stack.pop_back();
        const int my_chain = chain_of.back(); chain_of.pop_back();
        if (it.link >= 0) {
            if (it.link >= (int32_t)out.nodes.size()) { bad++; continue; }
            const BvhNode& nd = out.nodes[it.link];
            for (int ch = 0; ch < kBvhWidth; ch++) {
                const int32_t l = f2i(nd.q[24 + ch]);
                if (l == kEmptyLeaf) { if (!(nd.q[6 * ch + 3] < 0.0f)) bad++; continue; }
                Item nx{l, {nd.q[6 * ch], nd.q[6 * ch + 1], nd.q[6 * ch + 2]}, {nd.q[6 * ch + 3], nd.q[6 * ch + 4], nd.q[6 * ch + 5]}, true};
                auto chain = chains[my_chain];
                chain.push_back({{nx.c[0], nx.c[1], nx.c[2]}, {nx.h[0], nx.h[1], nx.h[2]}});
                chains.push_back(std::move(chain));
                stack.push_back(nx); chain_of.push_back((int)chains.size() - 1);
            }
        } else {
            const int32_t link = ~it.link, first = link >> 2, cnt = (link & 3) + 1;
            for (int k = 0; k < cnt; k++) {
                if (first + k >= (int32_t)out.tris.size()) { bad++; continue; }
                const BvhTri& t = out.tris[first + k];
                const int32_t id = f2i(t.q[3]);
                if (id < 0 || id >= n_tri || seen[id]) { bad++; continue; }
                seen[id] = 1;
                const float* v = tri_vertices + 9 * (size_t)id;
                for (const auto& bx : chains[my_chain])
                    for (int q = 0; q < 3; q++) if (!inside(bx.first.data(), bx.second.data(), v + 3 * q)) bad++;
            }
        }
        chains[my_chain].clear(); chains[my_chain].shrink_to_fit();
    }
    for (int64_t i = 0; i < n_tri; i++) if (!seen[i]) bad++;
    return bad;
}

// Host only: the BVH builder as a service (north-star: "internal/scene gains a BVH builder that emits a flattened,
// cache-line-aligned node array" — the Go side binds this, go/internal/engine/cuda/bvh.go).
int ptb_bvh_build(const float* tri_vertices, int64_t n_tri, ptb_bvh* out) {
    if (!tri_vertices || n_tri < 1 || !out) return fail(nullptr, PTB_ERR_INVALID, "bad argument");
    if (n_tri > (1ll << 28)) return fail(nullptr, PTB_ERR_LIMIT, "more than 2^28 triangles");
    std::memset(out, 0, sizeof *out);
    std::vector<int32_t> zeros((size_t)n_tri, 0);
    BvhBuildInput in{tri_vertices, n_tri, zeros.data(), zeros.data()};
    BvhBuildOutput bo;
    unsigned hw = std::thread::hardware_concurrency();
    build_bvh(in, bo, hw ? (int)hw : 4);
    const size_t nb = bo.nodes.size() * sizeof(BvhNode), tb = bo.tris.size() * sizeof(BvhTri);
    void* nodes = std::aligned_alloc(64, (nb + 63) / 64 * 64);
    void* tris = std::aligned_alloc(64, (tb + 63) / 64 * 64);
    if (!nodes || !tris) { std::free(nodes); std::free(tris); return fail(nullptr, PTB_ERR_LIMIT, "out of host memory"); }
    std::memcpy(nodes, bo.nodes.data(), nb);
    std::memcpy(tris, bo.tris.data(), tb);
    out->nodes = (const float*)nodes; out->triangles = (const float*)tris;
    out->info.n_triangles = (int64_t)bo.tris.size(); out->info.n_nodes = (int64_t)bo.nodes.size();
    out->info.max_depth = bo.max_depth; out->info.node_bytes = sizeof(BvhNode); out->info.triangle_bytes = sizeof(BvhTri);
    out->info.sah_cost = bo.sah_cost; out->info.build_ms = bo.build_ms;
    return PTB_OK;
}
void ptb_bvh_free(ptb_bvh* b) {
    if (!b) return;
    std::free((void*)b->nodes); std::free((void*)b->triangles);
    std::memset(b, 0, sizeof *b);
}

int ptb_measure_fp32_peak(ptb_ctx* c, double* tflops) {
    if (!c || !tflops) return PTB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(c, cudaSetDevice(c->device));
    const int threads = 256, blocks = c->prop.multiProcessorCount * 8, iters = 4096;
    float* d = nullptr;
    CK(c, cudaMalloc((void**)&d, sizeof(float) * threads * blocks));
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(c->ev0, c->stream);
        int e = launch_fma_peak(d, blocks, threads, iters, c->stream);
        cudaEventRecord(c->ev1, c->stream);
        cudaError_t ce = e ? (cudaError_t)e : cudaStreamSynchronize(c->stream);
        if (ce != cudaSuccess) { cudaFree(d); return fail(c, PTB_ERR_CUDA, "fma probe: %s", cudaGetErrorString(ce)); }
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        double flops = 2.0 * 64.0 * (double)iters * threads * blocks;   // 8 chains x 8 unrolled FMAs per iteration
        double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaFree(d);
    *tflops = best;
    return PTB_OK;
}

}  // extern "C"
