#!/usr/bin/env python3
"""Headless render driver — the twin of the reference's `cmd/render -headless` (cmd/render/main.go:14-63) on the CUDA
backend, with the explicit knobs the reference lacks (its headless mode ignores scene.settings and
has only the two presets of util.go:25-42, so e.g. "640x360, 16 spp, depth 8" is unreachable there).

    python tools/render.py -scene scenes/example_simple.json -mode preview -out out.png
    python tools/render.py -scene scenes/metal_glass_room.json -width 3840 -height 2160 -spp 256 -depth 16 -seed 7 -out c3.png
    python tools/render.py -scene s.json -settings          # take width/height/spp/depth from the scene file (what the UI does, app.go:61-70)

Flags -scene/-mode/-out keep the reference's names and defaults (main.go:17-21); -gpu/-headless are accepted and
ignored (this driver is always headless and always CUDA).
"""
from __future__ import annotations

import argparse
import pathlib
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def parse(argv=None):
    ap = argparse.ArgumentParser(prefix_chars="-", description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-scene", default="scenes/example_simple.json", help="path to scene JSON file")
    ap.add_argument("-mode", default="preview", help="render mode: preview or final (presets of util.go:25-42)")
    ap.add_argument("-out", default="output.png", help="output PNG file")
    ap.add_argument("-gpu", action="store_true", help="accepted for compatibility (always CUDA)")
    ap.add_argument("-headless", action="store_true", help="accepted for compatibility (always headless)")
    ap.add_argument("-settings", action="store_true", help="use the scene file's settings block instead of the mode preset")
    ap.add_argument("-width", type=int, default=0)
    ap.add_argument("-height", type=int, default=0)
    ap.add_argument("-spp", type=int, default=0)
    ap.add_argument("-depth", type=int, default=-1)
    ap.add_argument("-seed", type=int, default=1, help="key of the counter RNG (the reference seeds from the clock)")
    ap.add_argument("-device", type=int, default=0)
    ap.add_argument("-devices", type=int, default=1, help="render on devices 0..N-1 of the box from this one process")
    return ap.parse_args(argv)


def resolve_settings(args, scene_settings, mode_settings):
    """mode preset -> (optionally) scene settings with non-zero fields -> explicit flags."""
    w, h, spp, depth = mode_settings.Width, mode_settings.Height, mode_settings.SamplesPerPx, mode_settings.MaxDepth
    if args.settings:
        s = scene_settings
        w, h = (s.Width or w), (s.Height or h)
        spp, depth = (s.SamplesPerPx or spp), (s.MaxDepth or depth)
    if args.width > 0:
        w = args.width
    if args.height > 0:
        h = args.height
    if args.spp > 0:
        spp = args.spp
    if args.depth >= 0:
        depth = args.depth
    return w, h, spp, depth


def main(argv=None) -> int:
    args = parse(argv)
    from path_trace_golang_b200 import engine, scene, PtbError
    try:
        sc = scene.Load(args.scene)                                       # main.go:47-50
    except PtbError as e:
        print(f"headless render error: load scene: {e.message}", file=sys.stderr)
        return 1
    w, h, spp, depth = resolve_settings(args, sc.Settings, engine.RenderSettingsForMode(args.mode))
    try:
        t0 = time.perf_counter()
        if args.devices > 1:
            m = engine.MultiContext(args.devices)
            m.upload(sc)
            img = m.render(engine.Context.cfg(w, h, spp, depth, seed=args.seed))
        else:
            ctx = engine.Context(args.device)
            t0 = time.perf_counter()
            img = engine.Render(sc, engine.RenderConfig(w, h, spp, depth), ctx=ctx, seed=args.seed)   # main.go:54
        dt = time.perf_counter() - t0
        engine.SavePNG(args.out, img)                                     # main.go:59
    except PtbError as e:
        print(f"headless render error: {e.message}", file=sys.stderr)
        return 1
    print(f"{args.scene}: {w}x{h}, {spp} spp, depth {depth} -> {args.out} in {dt * 1e3:.1f} ms "
          f"({w * h * spp / dt / 1e6:.0f} Msamples/s incl. upload and read-back)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
