// engine.h — C++ mirror of the reference's internal/engine public API (renderer.go:17-41, backend.go,
// util.go) with the CUDA backend plugged in where gpu.Render sits today (renderer.go:35-40, 250-263).
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "scene.h"

namespace engine {

struct RenderConfig { int Width = 0, Height = 0, SamplesPerPx = 0, MaxDepth = 0; };   // renderer.go:17-22

// image.RGBA: Pix holds R,G,B,A bytes, row y starts at y*Stride, Rect = (0,0)-(W,H).
struct RGBA {
    std::vector<uint8_t> Pix;
    int Stride = 0, W = 0, H = 0;
    static RGBA New(int w, int h) { RGBA i; i.W = w; i.H = h; i.Stride = 4 * w; i.Pix.assign((size_t)4 * w * h, 0); return i; }
};

// backend.go:5-28, extended with the CUDA backend.  BackendCPU / BackendGPU (OpenGL) are not part of this
// build: selecting them makes RenderInto report an error — there is no CPU fallback.
enum Backend { BackendCPU = 0, BackendGPU = 1, BackendCUDA = 2 };
void SetBackend(Backend b);
Backend GetBackend();

// Counter-RNG key and device for subsequent renders (the reference seeds from the clock, random.go:14-16).
void SetSeed(uint32_t seed);
void SetDevice(int device);
void SetDevices(int n);          // n > 1: render every frame on devices 0..n-1 of the box (ptb_multi_*); progress fires once at the end
const std::string& LastError();

RGBA Render(const scene::Scene& sc, RenderConfig cfg);                                               // renderer.go:25-29
int RenderInto(const scene::Scene& sc, RenderConfig cfg, RGBA& img, const std::function<void()>& progress);   // renderer.go:34-41
RGBA RenderScene(const scene::Scene& sc, scene::RenderSettings settings);                            // util.go:13-22
scene::RenderSettings RenderSettingsForMode(const std::string& mode);                                // util.go:25-42
void SavePNG(const std::string& path, const uint8_t* pix, size_t stride, int w, int h);              // util.go:45-55

// Same as RenderInto but on an explicit context and raw image memory (what the C ABI wrapper calls).
int RenderIntoCtx(ptb_ctx* ctx, const scene::Scene& sc, RenderConfig cfg, uint32_t seed, uint8_t* pix, size_t stride,
                  int img_w, int img_h, ptb_progress_fn progress, void* user);
int RenderFlatCtx(ptb_ctx* ctx, const ptb_scene& view, RenderConfig cfg, uint32_t seed, uint8_t* pix, size_t stride,
                  int img_w, int img_h, ptb_progress_fn progress, void* user);


// Accumulation checkpoints (the reference persists only scenes and PNGs, util.go:45-55): renders cfg in calls of
// samples_per_call samples through ptb_render_resume and rewrites `path` (header + binary32 sums, atomically via rename) after
// every call; a later call with the same scene / cfg / seed and an existing file continues where that one stopped and ends
// with the same bits as an uninterrupted render.  max_calls > 0 stops after that many calls (an "interrupted" render);
// *spp_done = samples accumulated so far.
int RenderCheckpointed(ptb_ctx* ctx, const ptb_scene& view, RenderConfig cfg, uint32_t seed, uint8_t* pix, size_t stride, int img_w, int img_h,
                       const std::string& path, int samples_per_call, int max_calls, int* spp_done, ptb_progress_fn progress, void* user);

}  // namespace engine
