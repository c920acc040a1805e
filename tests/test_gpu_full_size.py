"""BASELINE.json full-size configurations on the GPU, checked through size-independent properties:
determinism, sample-range linearity, counters vs the oracle's per-sample statistics at low resolution,
block-mean radiance vs the oracle at low resolution."""
import numpy as np
import pytest

from conftest import CONFIGS, SCENE_DEPTH

pytestmark = pytest.mark.gpu


def test_c3_full_size_properties(ctx, host_scenes, oracle_scenes):
    """C3 = metal_glass_room 3840x2160, 256 spp, depth 16 (the headline workload)."""
    name, W, H, spp, depth = CONFIGS["C3"]
    ctx.upload(host_scenes[name])
    img = ctx.render(ctx.cfg(W, H, spp, depth, seed=1))
    assert img.shape == (H, W, 4) and (img[..., 3] == 255).all()
    ms = ctx.stats()["last_render_ms"]
    print(f"C3 render {ms:.1f} ms -> {W * H * spp / ms / 1e3:.1f} Msamples/s")
    # image statistics vs a converged low-resolution oracle render (same camera => same field of view)
    small_spp = 256
    ora, st = oracle_scenes[name].render_sum(240, 135, small_spp, depth, seed=123, precision=64)
    ora /= small_spp
    lin = (img[..., :3].astype(np.float64) / 255.999) ** 2                 # invert the gamma-2 epilogue (approx.)
    dev_small = lin.reshape(135, 16, 240, 16, 3).mean(axis=(1, 3))
    ora_g = np.sqrt(np.clip(ora, 0, None))                                  # compare in gamma space, block means
    dev_g = np.sqrt(dev_small)
    rmse = np.sqrt(((dev_g - ora_g) ** 2).mean())
    print(f"C3 gamma-space block RMSE vs oracle {rmse:.4f}")
    assert rmse < 0.03
    # linearity at full size: two half sample ranges sum to the full range
    a = ctx.render_accum(ctx.cfg(W, H, 8, depth, seed=1, sample_begin=0, sample_count=4))
    b = ctx.render_accum(ctx.cfg(W, H, 8, depth, seed=1, sample_begin=4, sample_count=4))
    f = ctx.render_accum(ctx.cfg(W, H, 8, depth, seed=1))
    assert np.allclose(a.astype(np.float64) + b, f, rtol=1e-5, atol=1e-6)


def test_c2_full_size_counters(ctx, host_scenes, oracle_scenes):
    """C2 = test_scene 1920x1080, 64 spp, depth 10: device counters per sample agree with the oracle's
    per-sample statistics measured at 1/8 resolution (different pixels, same distribution): within 3 %."""
    name, W, H, spp, depth = CONFIGS["C2"]
    ctx.upload(host_scenes[name])
    ctx.render_accum(ctx.cfg(W, H, spp, depth, seed=1, stats=True))
    d = ctx.stats()
    _, o = oracle_scenes[name].render_sum(W // 8, H // 8, spp, depth, seed=1, precision=64)
    for k in ["segments", "exit_scans", "end_sky", "end_emissive", "end_depth"]:
        dv, ov = d[k] / d["samples"], o[k] / o["samples"]
        print(k, dv, ov)
        assert abs(dv - ov) <= 0.03 * max(ov, 0.05), (k, dv, ov)
