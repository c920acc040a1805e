/*
 * render_c.c — the C ABI from plain C99: what a cgo / JNI / FFI binding does, without the binding.
 *
 *   gcc -std=c99 -Iinclude -o render_c examples/render_c.c -Lpath_trace_golang_b200 -lptb200 -Wl,-rpath,$PWD/path_trace_golang_b200
 *   ./render_c scenes/example_simple.json out.png 640 360 16 8
 *
 * Mirrors cmd/render (main.go:47-59): scene.Load -> engine.RenderInto -> SavePNG.  There is no CPU fallback: without
 * a CUDA device ptb_create fails and the program exits 1 with the library's message.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ptb200_host.h"

static void tick(void* user) { ++*(int*)user; }

int main(int argc, char** argv) {
    const char* scene_path = argc > 1 ? argv[1] : "scenes/example_simple.json";
    const char* out_path = argc > 2 ? argv[2] : "output.png";
    const int w = argc > 3 ? atoi(argv[3]) : 400, h = argc > 4 ? atoi(argv[4]) : 225;
    const int spp = argc > 5 ? atoi(argv[5]) : 20, depth = argc > 6 ? atoi(argv[6]) : 20;

    ptb_host_scene* sc = NULL;
    if (ptb_host_scene_load(scene_path, &sc) != PTB_OK) {
        fprintf(stderr, "load scene: %s\n", ptb_host_last_error());
        return 1;
    }
    ptb_ctx* ctx = NULL;
    if (ptb_create(0, &ctx) != PTB_OK) {
        fprintf(stderr, "CUDA initialization failed: %s\n", ptb_last_error(NULL));
        ptb_host_scene_free(sc);
        return 1;
    }
    const size_t stride = (size_t)w * 4;
    uint8_t* pix = (uint8_t*)calloc((size_t)h, stride);      /* image.RGBA.Pix */
    int ticks = 0, rc;
    rc = ptb_engine_render_into(ctx, sc, w, h, spp, depth, /*seed=*/1u, pix, stride, w, h, tick, &ticks);
    if (rc != PTB_OK) fprintf(stderr, "render: %s\n", ptb_host_last_error());
    else if ((rc = ptb_engine_save_png(out_path, pix, stride, w, h)) != PTB_OK) fprintf(stderr, "save: %s\n", ptb_host_last_error());
    else {
        ptb_stats st;
        memset(&st, 0, sizeof st);
        ptb_get_stats(ctx, &st);
        printf("%s: %dx%d, %d spp, depth %d -> %s (%d progress callbacks, last kernel %.2f ms)\n", scene_path, w, h, spp, depth,
               out_path, ticks, st.last_render_ms);
    }
    free(pix);
    ptb_destroy(ctx);
    ptb_host_scene_free(sc);
    return rc == PTB_OK ? 0 : 1;
}
