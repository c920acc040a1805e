// engine.cpp — host mirror of internal/engine's public API over the C ABI, plus the ptb200_host.h exports.
#include "engine.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <mutex>
#include <stdexcept>

#include "../../../include/ptb200_host.h"

namespace engine {
namespace {
Backend g_backend = BackendCUDA;
uint32_t g_seed = 1;
int g_device = 0;
ptb_ctx* g_ctx = nullptr;
int g_ctx_device = -1;
int g_devices = 1;
ptb_multi* g_multi = nullptr;
int g_multi_n = 0;
std::mutex g_mu;
std::string g_err;

ptb_ctx* default_ctx() {           // like gpu.ensureWorker (gpu.go:266-297): created once, init failure is reported
    if (g_ctx && g_ctx_device == g_device) return g_ctx;
    if (g_ctx) { ptb_destroy(g_ctx); g_ctx = nullptr; }
    if (ptb_create(g_device, &g_ctx) != PTB_OK) { g_err = std::string("CUDA initialization failed: ") + ptb_last_error(nullptr); g_ctx = nullptr; return nullptr; }
    g_ctx_device = g_device;
    return g_ctx;
}
}  // namespace

void SetBackend(Backend b) { g_backend = (b == BackendCPU || b == BackendGPU || b == BackendCUDA) ? b : BackendCPU; }   // backend.go:16-23
Backend GetBackend() { return g_backend; }
void SetSeed(uint32_t s) { g_seed = s; }
void SetDevice(int d) { g_device = d; }
void SetDevices(int n) { g_devices = n < 1 ? 1 : n; }
const std::string& LastError() { return g_err; }

int RenderIntoCtx(ptb_ctx* ctx, const scene::Scene& sc, RenderConfig cfg, uint32_t seed, uint8_t* pix, size_t stride,
                  int img_w, int img_h, ptb_progress_fn progress, void* user) {
    if (img_w != cfg.Width || img_h != cfg.Height) return PTB_OK;     // renderer.go:46-49: silent return
    scene::Flat flat = scene::Flatten(sc);                            // per call: the caller may have edited the scene (the UI does)
    return RenderFlatCtx(ctx, flat.view(), cfg, seed, pix, stride, img_w, img_h, progress, user);
}

int RenderFlatCtx(ptb_ctx* ctx, const ptb_scene& view, RenderConfig cfg, uint32_t seed, uint8_t* pix, size_t stride,
                  int img_w, int img_h, ptb_progress_fn progress, void* user) {
    if (img_w != cfg.Width || img_h != cfg.Height) return PTB_OK;
    int rc = ptb_scene_upload(ctx, &view);
    if (rc != PTB_OK) return rc;
    ptb_cfg c{};
    c.width = cfg.Width; c.height = cfg.Height; c.samples_per_px = cfg.SamplesPerPx; c.max_depth = cfg.MaxDepth;
    c.seed = seed;
    return ptb_render(ctx, &c, pix, stride, progress, user);
}

int RenderInto(const scene::Scene& sc, RenderConfig cfg, RGBA& img, const std::function<void()>& progress) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_backend != BackendCUDA) {
        g_err = "backend not available in this build: only BackendCUDA exists (no CPU fallback)";
        std::fprintf(stderr, "render error: %s\n", g_err.c_str());
        return PTB_ERR_INVALID;
    }
    if (g_devices > 1) {
        if (img.W != cfg.Width || img.H != cfg.Height) return PTB_OK;
        if (g_multi && g_multi_n != g_devices) { ptb_multi_destroy(g_multi); g_multi = nullptr; }
        if (!g_multi) {
            if (ptb_multi_create(nullptr, g_devices, &g_multi) != PTB_OK) {
                g_err = std::string("CUDA initialization failed: ") + ptb_multi_last_error(nullptr);
                g_multi = nullptr;
                std::fprintf(stderr, "%s\n", g_err.c_str());
                return PTB_ERR_CUDA;
            }
            g_multi_n = g_devices;
        }
        scene::Flat flat = scene::Flatten(sc);
        ptb_scene view = flat.view();
        ptb_cfg c{};
        c.width = cfg.Width; c.height = cfg.Height; c.samples_per_px = cfg.SamplesPerPx; c.max_depth = cfg.MaxDepth;
        c.seed = g_seed;
        int rc = ptb_multi_scene_upload(g_multi, &view);
        if (rc == PTB_OK) rc = ptb_multi_render(g_multi, &c, img.Pix.data(), (size_t)img.Stride);
        if (rc != PTB_OK) {
            g_err = ptb_multi_last_error(g_multi);
            std::fprintf(stderr, "CUDA render error: %s\n", g_err.c_str());
        } else if (progress) {
            progress();
        }
        return rc;
    }
    ptb_ctx* ctx = default_ctx();
    if (!ctx) { std::fprintf(stderr, "%s\n", g_err.c_str()); return PTB_ERR_CUDA; }
    struct Thunk { const std::function<void()>* f; } th{&progress};
    ptb_progress_fn cb = nullptr;
    if (progress) cb = [](void* u) { (*static_cast<Thunk*>(u)->f)(); };
    int rc = RenderIntoCtx(ctx, sc, cfg, g_seed, img.Pix.data(), (size_t)img.Stride, img.W, img.H, cb, &th);
    if (rc != PTB_OK) {               // the GL path logs and falls back to the CPU (renderer.go:257-262); we only log
        g_err = ptb_last_error(ctx);
        std::fprintf(stderr, "CUDA render error: %s\n", g_err.c_str());
    }
    return rc;
}

// ---- accumulation checkpoints.  File = CheckpointHeader + width*height*3 binary32 sums of samples [0, spp_done).
namespace {
struct CheckpointHeader {
    char magic[8];                       // "PTBACC1\0"
    int32_t width, height, spp_total, spp_done, max_depth;
    uint32_t seed;
    uint64_t scene_key;                  // FNV-1a of the flattened scene (analytic arrays, camera, sky, mesh generation-free triangle count)
};
uint64_t fnv(uint64_t h, const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 0x100000001b3ull; } return h; }
uint64_t scene_key_of(const ptb_scene& v) {
    uint64_t h = 0xcbf29ce484222325ull;
    h = fnv(h, &v.n_obj, 4); h = fnv(h, &v.n_mat, 4);
    if (v.n_obj > 0) { h = fnv(h, v.obj_type, 4 * (size_t)v.n_obj); h = fnv(h, v.obj_mat, 4 * (size_t)v.n_obj); h = fnv(h, v.obj_pos, 24 * (size_t)v.n_obj); h = fnv(h, v.obj_size, 24 * (size_t)v.n_obj); }
    if (v.n_mat > 0) {
        h = fnv(h, v.mat_type, 4 * (size_t)v.n_mat); h = fnv(h, v.mat_albedo, 24 * (size_t)v.n_mat); h = fnv(h, v.mat_rough, 8 * (size_t)v.n_mat);
        h = fnv(h, v.mat_ior, 8 * (size_t)v.n_mat); h = fnv(h, v.mat_emit, 24 * (size_t)v.n_mat); h = fnv(h, v.mat_power, 8 * (size_t)v.n_mat);
        h = fnv(h, v.mat_absorption, 24 * (size_t)v.n_mat); h = fnv(h, v.mat_smoothness, 8 * (size_t)v.n_mat);
    }
    h = fnv(h, &v.camera, sizeof v.camera); h = fnv(h, &v.sky.kind, 4); h = fnv(h, v.sky.color, 72);
    if (v.n_mesh > 0) { h = fnv(h, v.mesh_tri_begin, 8 * (size_t)(v.n_mesh + 1)); h = fnv(h, v.tri_vertices, 36 * (size_t)v.mesh_tri_begin[v.n_mesh]); }
    return h;
}
}  // namespace

int RenderCheckpointed(ptb_ctx* ctx, const ptb_scene& view, RenderConfig cfg, uint32_t seed, uint8_t* pix, size_t stride, int img_w, int img_h,
                       const std::string& path, int samples_per_call, int max_calls, int* spp_done_out, ptb_progress_fn progress, void* user) {
    if (img_w != cfg.Width || img_h != cfg.Height) return PTB_OK;     // renderer.go:46-49
    if (samples_per_call < 1) samples_per_call = cfg.SamplesPerPx;
    int rc = ptb_scene_upload(ctx, &view);
    if (rc != PTB_OK) return rc;
    CheckpointHeader want{};
    std::memcpy(want.magic, "PTBACC1", 8);
    want.width = cfg.Width; want.height = cfg.Height; want.spp_total = cfg.SamplesPerPx; want.max_depth = cfg.MaxDepth; want.seed = seed;
    want.scene_key = scene_key_of(view);
    const size_t n = (size_t)cfg.Width * cfg.Height * 3;
    std::vector<float> sums(n, 0.0f);
    int done = 0;
    {   // resume from the file when it belongs to this very render; anything else is ignored (and overwritten)
        std::ifstream f(path, std::ios::binary);
        CheckpointHeader h{};
        if (f && f.read((char*)&h, sizeof h) && !std::memcmp(h.magic, want.magic, 8) && h.width == want.width && h.height == want.height &&
            h.spp_total == want.spp_total && h.max_depth == want.max_depth && h.seed == want.seed && h.scene_key == want.scene_key &&
            h.spp_done > 0 && h.spp_done <= h.spp_total && f.read((char*)sums.data(), (std::streamsize)(n * sizeof(float))))
            done = h.spp_done;
        else std::fill(sums.begin(), sums.end(), 0.0f);
    }
    ptb_cfg c{};
    c.width = cfg.Width; c.height = cfg.Height; c.samples_per_px = cfg.SamplesPerPx; c.max_depth = cfg.MaxDepth; c.seed = seed;
    int calls = 0;
    if (done >= cfg.SamplesPerPx) {       // complete checkpoint: nothing left to trace, the device only runs the pixel epilogue
        if ((rc = ptb_finalize_host(ctx, sums.data(), cfg.Width, cfg.Height, cfg.SamplesPerPx, pix, stride)) != PTB_OK) return rc;
    }
    while (done < cfg.SamplesPerPx && (max_calls <= 0 || calls < max_calls)) {
        const int cnt = std::min(samples_per_call, cfg.SamplesPerPx - done);
        c.sample_begin = done; c.sample_count = cnt;
        if ((rc = ptb_render_resume(ctx, &c, sums.data(), pix, stride)) != PTB_OK) return rc;
        done += cnt; calls++;
        CheckpointHeader h = want;
        h.spp_done = done;
        const std::string tmp = path + ".tmp";
        {
            std::ofstream f(tmp, std::ios::binary | std::ios::trunc);
            f.write((const char*)&h, sizeof h);
            f.write((const char*)sums.data(), (std::streamsize)(n * sizeof(float)));
            if (!f) { g_err = "checkpoint: write failed: " + tmp; return PTB_ERR_INVALID; }
        }
        if (std::rename(tmp.c_str(), path.c_str()) != 0) { g_err = "checkpoint: rename failed: " + path; return PTB_ERR_INVALID; }
        if (progress) progress(user);
    }
    if (spp_done_out) *spp_done_out = done;
    return PTB_OK;
}

RGBA Render(const scene::Scene& sc, RenderConfig cfg) {
    RGBA img = RGBA::New(cfg.Width, cfg.Height);
    RenderInto(sc, cfg, img, nullptr);
    return img;
}

RGBA RenderScene(const scene::Scene& sc, scene::RenderSettings s) {
    return Render(sc, RenderConfig{s.Width, s.Height, s.SamplesPerPx, s.MaxDepth});
}

scene::RenderSettings RenderSettingsForMode(const std::string& mode) {
    if (mode == "final") return scene::RenderSettings{1920, 1080, 1000, 80};
    return scene::RenderSettings{400, 225, 20, 20};
}

// ---- PNG (util.go:45-55 uses image/png).  8-bit RGBA, zlib stream of stored (uncompressed) deflate blocks.
namespace {
uint32_t crc_table[256];
bool crc_ready = false;
uint32_t crc32(uint32_t crc, const uint8_t* p, size_t n) {
    if (!crc_ready) {
        for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; crc_table[i] = c; }
        crc_ready = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; i++) crc = crc_table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
    return ~crc;
}
void put32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
void chunk(std::ofstream& f, const char* type, const std::vector<uint8_t>& data) {
    std::vector<uint8_t> hdr; put32(hdr, (uint32_t)data.size());
    f.write((const char*)hdr.data(), 4);
    std::vector<uint8_t> body(type, type + 4);
    body.insert(body.end(), data.begin(), data.end());
    f.write((const char*)body.data(), (std::streamsize)body.size());
    std::vector<uint8_t> c; put32(c, crc32(0, body.data(), body.size()));
    f.write((const char*)c.data(), 4);
}
}  // namespace

void SavePNG(const std::string& path, const uint8_t* pix, size_t stride, int w, int h) {
    std::ofstream f(path, std::ios::binary | std::ios::trunc);
    if (!f) throw std::runtime_error("create png: open " + path + ": " + std::strerror(errno));
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    f.write((const char*)sig, 8);
    std::vector<uint8_t> ihdr; put32(ihdr, (uint32_t)w); put32(ihdr, (uint32_t)h);
    ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(f, "IHDR", ihdr);
    std::vector<uint8_t> raw;
    raw.reserve((size_t)h * (1 + (size_t)w * 4));
    for (int y = 0; y < h; y++) { raw.push_back(0); raw.insert(raw.end(), pix + (size_t)y * stride, pix + (size_t)y * stride + (size_t)w * 4); }
    std::vector<uint8_t> z; z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1, b = 0;
    for (uint8_t c : raw) { a = (a + c) % 65521; b = (b + a) % 65521; }
    for (size_t off = 0; off < raw.size() || off == 0; off += 65535) {
        size_t n = raw.size() - off < 65535 ? raw.size() - off : 65535;
        z.push_back(off + n >= raw.size() ? 1 : 0);
        z.push_back(n & 0xFF); z.push_back(n >> 8); z.push_back(~n & 0xFF); z.push_back((~n >> 8) & 0xFF);
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
        if (raw.empty()) break;
    }
    put32(z, (b << 16) | a);
    chunk(f, "IDAT", z);
    chunk(f, "IEND", {});
    if (!f) throw std::runtime_error("encode png: write failed");
}

}  // namespace engine

// ------------------------------------------------------------------ ptb200_host.h exports
struct ptb_host_scene {
    std::unique_ptr<scene::Scene> sc;
    scene::Flat flat;
};
namespace {
thread_local std::string h_err;
int hfail(int code, const std::string& m) { h_err = m; return code; }
}  // namespace

extern "C" {

const char* ptb_host_last_error(void) { return h_err.c_str(); }

static int wrap(std::unique_ptr<scene::Scene> sc, ptb_host_scene** out) {
    auto* h = new ptb_host_scene();
    h->sc = std::move(sc);
    h->flat = scene::Flatten(*h->sc);
    *out = h;
    return PTB_OK;
}
int ptb_host_scene_load(const char* path, ptb_host_scene** out) {
    if (!path || !out) return hfail(PTB_ERR_INVALID, "NULL argument");
    try { return wrap(scene::Load(path), out); } catch (const std::exception& e) { *out = nullptr; return hfail(PTB_ERR_INVALID, e.what()); }
}
int ptb_host_scene_parse(const char* json, size_t len, ptb_host_scene** out) {
    if (!json || !out) return hfail(PTB_ERR_INVALID, "NULL argument");
    try { return wrap(scene::Parse(std::string(json, len)), out); } catch (const std::exception& e) { *out = nullptr; return hfail(PTB_ERR_INVALID, e.what()); }
}
int ptb_host_scene_save(const ptb_host_scene* sc, const char* path) {
    if (!sc || !path) return hfail(PTB_ERR_INVALID, "NULL argument");
    try { scene::Save(path, *sc->sc); return PTB_OK; } catch (const std::exception& e) { return hfail(PTB_ERR_INVALID, e.what()); }
}
size_t ptb_host_scene_marshal(const ptb_host_scene* sc, char* buf, size_t cap) {
    if (!sc) return 0;
    std::string s = scene::Marshal(*sc->sc);
    if (buf && cap) std::memcpy(buf, s.data(), s.size() < cap ? s.size() : cap);
    return s.size();
}
void ptb_host_scene_free(ptb_host_scene* sc) { delete sc; }
int ptb_host_scene_flat(const ptb_host_scene* sc, ptb_scene* out) {
    if (!sc || !out) return hfail(PTB_ERR_INVALID, "NULL argument");
    *out = sc->flat.view();
    return PTB_OK;
}
int ptb_host_scene_settings(const ptb_host_scene* sc, int32_t out[4]) {
    if (!sc || !out) return hfail(PTB_ERR_INVALID, "NULL argument");
    const scene::RenderSettings& s = sc->sc->Settings;
    out[0] = s.Width; out[1] = s.Height; out[2] = s.SamplesPerPx; out[3] = s.MaxDepth;
    return PTB_OK;
}
int ptb_host_scene_counts(const ptb_host_scene* sc, int32_t* n_objects, int32_t* n_materials) {
    if (!sc) return hfail(PTB_ERR_INVALID, "NULL argument");
    if (n_objects) *n_objects = (int32_t)sc->sc->Objects.size();
    if (n_materials) *n_materials = (int32_t)sc->sc->Materials.size();
    return PTB_OK;
}
void ptb_engine_settings_for_mode(const char* mode, int32_t out[4]) {
    scene::RenderSettings s = engine::RenderSettingsForMode(mode ? mode : "");
    out[0] = s.Width; out[1] = s.Height; out[2] = s.SamplesPerPx; out[3] = s.MaxDepth;
}
int ptb_engine_render_into(ptb_ctx* ctx, const ptb_host_scene* sc, int32_t width, int32_t height, int32_t spp, int32_t max_depth,
                           uint32_t seed, uint8_t* pix, size_t stride, int32_t img_w, int32_t img_h, ptb_progress_fn progress, void* user) {
    if (!ctx || !sc) return hfail(PTB_ERR_INVALID, "NULL argument");
    // a ptb_host_scene is immutable after load/parse, so its flattening (done once, ptb_host_scene_flat returns the same view)
    // is reused; a 1 M-triangle mesh takes 35 ms to flatten
    int rc = engine::RenderFlatCtx(ctx, sc->flat.view(), engine::RenderConfig{width, height, spp, max_depth}, seed, pix, stride, img_w, img_h, progress, user);
    if (rc != PTB_OK) h_err = ptb_last_error(ctx);
    return rc;
}
int ptb_engine_render_checkpointed(ptb_ctx* ctx, const ptb_host_scene* sc, int32_t width, int32_t height, int32_t spp, int32_t max_depth,
                                   uint32_t seed, uint8_t* pix, size_t stride, int32_t img_w, int32_t img_h, const char* path,
                                   int32_t samples_per_call, int32_t max_calls, int32_t* spp_done, ptb_progress_fn progress, void* user) {
    if (!ctx || !sc || !pix || !path) return hfail(PTB_ERR_INVALID, "NULL argument");
    int done = 0;
    int rc = engine::RenderCheckpointed(ctx, sc->flat.view(), engine::RenderConfig{width, height, spp, max_depth}, seed, pix, stride, img_w, img_h,
                                        path, samples_per_call, max_calls, &done, progress, user);
    if (spp_done) *spp_done = done;
    if (rc != PTB_OK) h_err = engine::LastError().empty() ? ptb_last_error(ctx) : engine::LastError();
    return rc;
}
int ptb_engine_save_png(const char* path, const uint8_t* pix, size_t stride, int32_t width, int32_t height) {
    if (!path || !pix || width < 1 || height < 1 || stride < (size_t)width * 4) return hfail(PTB_ERR_INVALID, "bad argument");
    try { engine::SavePNG(path, pix, stride, width, height); return PTB_OK; } catch (const std::exception& e) { return hfail(PTB_ERR_INVALID, e.what()); }
}

}  // extern "C"
