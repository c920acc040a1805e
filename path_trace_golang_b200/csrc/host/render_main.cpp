// render_main.cpp — `ptb_render`: the reference's cmd/render (main.go:14-63) in C++ on the CUDA backend.
//
// Same flags and defaults as the reference (-scene, -mode, -gpu, -headless, -out; Go-style single-dash flags, `-flag value`
// or `-flag=value`), plus what SURVEY §8f rank 2 asks for: -width -height -spp -depth -seed -settings -device -devices.
// The UI path (internal/ui) is out of scope: without -headless this driver still renders headless and says so.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "engine.h"

namespace {
struct Flags {
    std::string scene = "scenes/example_simple.json", mode = "preview", out = "output.png";   // main.go:17-21
    bool gpu = false, headless = false, settings = false;
    int width = 0, height = 0, spp = 0, depth = -1, device = 0, devices = 1;
    unsigned seed = 1;
};
bool parse(int argc, char** argv, Flags& f) {
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a.size() > 1 && a[0] == '-' && a[1] == '-') a = a.substr(1);      // Go's flag package accepts --flag too
        std::string val;
        bool has_val = false;
        size_t eq = a.find('=');
        if (eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); has_val = true; }
        auto next = [&]() -> const char* {
            if (has_val) return val.c_str();
            if (i + 1 < argc) return argv[++i];
            std::fprintf(stderr, "flag needs an argument: %s\n", a.c_str());
            return nullptr;
        };
        const char* v = nullptr;
        // Go's flag package: a boolean flag is `-flag` or `-flag=value` (strconv.ParseBool values); `-flag value` is NOT its argument
        auto boolean = [&](bool& dst) -> bool {
            if (!has_val) { dst = true; return true; }
            if (val == "1" || val == "t" || val == "T" || val == "true" || val == "TRUE" || val == "True") { dst = true; return true; }
            if (val == "0" || val == "f" || val == "F" || val == "false" || val == "FALSE" || val == "False") { dst = false; return true; }
            std::fprintf(stderr, "invalid boolean value \"%s\" for %s: parse error\n", val.c_str(), a.c_str());
            return false;
        };
        if (a == "-gpu") { if (!boolean(f.gpu)) return false; }
        else if (a == "-headless") { if (!boolean(f.headless)) return false; }
        else if (a == "-settings") { if (!boolean(f.settings)) return false; }
        else if (a == "-scene") { if (!(v = next())) return false; f.scene = v; }
        else if (a == "-mode") { if (!(v = next())) return false; f.mode = v; }
        else if (a == "-out") { if (!(v = next())) return false; f.out = v; }
        else if (a == "-width") { if (!(v = next())) return false; f.width = std::atoi(v); }
        else if (a == "-height") { if (!(v = next())) return false; f.height = std::atoi(v); }
        else if (a == "-spp") { if (!(v = next())) return false; f.spp = std::atoi(v); }
        else if (a == "-depth") { if (!(v = next())) return false; f.depth = std::atoi(v); }
        else if (a == "-seed") { if (!(v = next())) return false; f.seed = (unsigned)std::strtoul(v, nullptr, 10); }
        else if (a == "-devices") { if (!(v = next())) return false; f.devices = std::atoi(v); }
        else if (a == "-device") { if (!(v = next())) return false; f.device = std::atoi(v); }
        else { std::fprintf(stderr, "flag provided but not defined: %s\n", a.c_str()); return false; }
    }
    return true;
}
}  // namespace

int main(int argc, char** argv) {
    std::fprintf(stderr, "pathtracer: starting main()\n");                    // main.go:15
    Flags f;
    if (!parse(argc, argv, f)) return 2;
    std::fprintf(stderr, "flags: scene=%s mode=%s headless=%d out=%s\n", f.scene.c_str(), f.mode.c_str(), (int)f.headless, f.out.c_str());
    if (!f.headless) std::fprintf(stderr, "note: the desktop UI is out of scope of this backend; rendering headless\n");
    engine::SetBackend(engine::BackendCUDA);                                  // main.go:26-30 selects CPU/GPU; here: CUDA only
    engine::SetDevice(f.device);
    engine::SetDevices(f.devices);
    engine::SetSeed(f.seed);
    try {
        std::unique_ptr<scene::Scene> sc = scene::Load(f.scene);              // main.go:47
        scene::RenderSettings s = engine::RenderSettingsForMode(f.mode);     // main.go:52 (scene.settings is ignored, like the reference)
        if (f.settings) {                                                     // what the UI does instead, app.go:61-70
            if (sc->Settings.Width) s.Width = sc->Settings.Width;
            if (sc->Settings.Height) s.Height = sc->Settings.Height;
            if (sc->Settings.SamplesPerPx) s.SamplesPerPx = sc->Settings.SamplesPerPx;
            if (sc->Settings.MaxDepth) s.MaxDepth = sc->Settings.MaxDepth;
        }
        if (f.width > 0) s.Width = f.width;
        if (f.height > 0) s.Height = f.height;
        if (f.spp > 0) s.SamplesPerPx = f.spp;
        if (f.depth >= 0) s.MaxDepth = f.depth;
        engine::RGBA img = engine::RGBA::New(s.Width, s.Height);
        auto t0 = std::chrono::steady_clock::now();
        int rc = engine::RenderInto(*sc, engine::RenderConfig{s.Width, s.Height, s.SamplesPerPx, s.MaxDepth}, img, nullptr);   // main.go:54
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (rc != 0) { std::fprintf(stderr, "headless render error: render scene: %s\n", engine::LastError().c_str()); return 1; }
        engine::SavePNG(f.out, img.Pix.data(), (size_t)img.Stride, img.W, img.H);                                              // main.go:59
        std::fprintf(stderr, "%s: %dx%d, %d spp, depth %d -> %s in %.1f ms (%.0f Msamples/s incl. context creation, upload, read-back)\n",
                     f.scene.c_str(), s.Width, s.Height, s.SamplesPerPx, s.MaxDepth, f.out.c_str(), ms,
                     (double)s.Width * s.Height * s.SamplesPerPx / ms / 1e3);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "headless render error: %s\n", e.what());        // main.go:38-41
        return 1;
    }
    return 0;
}
