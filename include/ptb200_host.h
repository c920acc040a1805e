/*
 * ptb200_host.h — C ABI of the C++ host mirror that sits ABOVE include/ptb200.h.
 *
 * The reference's host code is Go (internal/scene, internal/engine); there is no Go toolchain in
 * this image, so the same host logic — scene.Load/Save (internal/scene/io.go:10-38), the scene
 * flattening a cgo shim would do, and engine.RenderInto / RenderSettingsForMode / SavePNG
 * (internal/engine/renderer.go:34-41, util.go:25-55) — is written in C++ inside libptb200.so
 * (path_trace_golang_b200/csrc/host/) and exported here so that tests and bench.py can drive it
 * through ctypes.  A Go maintainer does NOT need this header: the cgo binding in go/ talks to
 * ptb200.h directly.
 */
#ifndef PTB200_HOST_H
#define PTB200_HOST_H

#include "ptb200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ptb_host_scene ptb_host_scene;

/* Message of the last failed ptb_host_* / ptb_engine_* call on this thread. */
const char* ptb_host_last_error(void);

/* scene.Load (io.go:10-22) / json decode of an in-memory document.  Errors: "open scene: ...",
 * "decode scene: ..." like the reference's wrapped errors. */
int ptb_host_scene_load(const char* path, ptb_host_scene** out);
int ptb_host_scene_parse(const char* json, size_t len, ptb_host_scene** out);
/* scene.Save (io.go:25-38): two-space indented JSON in Go's field order. */
int ptb_host_scene_save(const ptb_host_scene* sc, const char* path);
/* Marshalled JSON; returns the byte length, copies at most cap bytes into buf (buf may be NULL). */
size_t ptb_host_scene_marshal(const ptb_host_scene* sc, char* buf, size_t cap);
void ptb_host_scene_free(ptb_host_scene* sc);

/* The SoA view handed to ptb_scene_upload; pointers stay valid until ptb_host_scene_free. */
int ptb_host_scene_flat(const ptb_host_scene* sc, ptb_scene* out);
/* scene.RenderSettings of the file: out = {width, height, samples_per_px, max_depth} (scene.go:92-97). */
int ptb_host_scene_settings(const ptb_host_scene* sc, int32_t out[4]);
int ptb_host_scene_counts(const ptb_host_scene* sc, int32_t* n_objects, int32_t* n_materials);

/* engine.RenderSettingsForMode (util.go:25-42): "final" -> 1920x1080/1000/80, else 400x225/20/20. */
void ptb_engine_settings_for_mode(const char* mode, int32_t out[4]);

/* engine.RenderInto (renderer.go:34-41) on the CUDA backend: uploads the scene to ctx and renders
 * into the caller's image (pix/stride/img_w/img_h = image.RGBA Pix/Stride/Bounds).  If the image
 * size differs from cfg the call returns PTB_OK without touching pix — the reference's silent
 * "basic safety" return (renderer.go:46-49).  seed: key of the counter RNG.  No CPU fallback. */
int ptb_engine_render_into(ptb_ctx* ctx, const ptb_host_scene* sc, int32_t width, int32_t height,
                           int32_t samples_per_px, int32_t max_depth, uint32_t seed, uint8_t* pix,
                           size_t stride, int32_t img_w, int32_t img_h, ptb_progress_fn progress,
                           void* user);

/* Accumulation checkpoints (no reference equivalent: the reference persists only scenes and PNGs, util.go:45-55).  Renders
 * like ptb_engine_render_into, in calls of samples_per_call samples (ptb_render_resume), rewriting the checkpoint file `path`
 * after every call; if `path` already holds a checkpoint of this very render (same scene, size, spp, depth, seed) the render
 * continues from it and finishes with the same bytes as an uninterrupted one.  max_calls > 0: stop after that many calls
 * (what an interrupted run leaves behind).  *spp_done (may be NULL) = samples per pixel accumulated so far. */
int ptb_engine_render_checkpointed(ptb_ctx* ctx, const ptb_host_scene* sc, int32_t width, int32_t height, int32_t samples_per_px,
                                   int32_t max_depth, uint32_t seed, uint8_t* pix, size_t stride, int32_t img_w, int32_t img_h,
                                   const char* path, int32_t samples_per_call, int32_t max_calls, int32_t* spp_done,
                                   ptb_progress_fn progress, void* user);

/* engine.SavePNG (util.go:45-55) for an RGBA8 image. */
int ptb_engine_save_png(const char* path, const uint8_t* pix, size_t stride, int32_t width, int32_t height);

#ifdef __cplusplus
}
#endif
#endif
