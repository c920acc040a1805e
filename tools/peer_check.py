#!/usr/bin/env python3
"""Multi-process exchange check, run under torchrun on N >= 2 GPUs of one box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/peer_check.py

Every rank renders its sample range; the frame is assembled three ways — the fused peer-memory kernel (ptb_peer_*), NCCL
reduce_scatter + epilogue + gather, NCCL reduce to rank 0 + epilogue — and rank 0 compares them with each other and with the
single-GPU image of the same frame (same samples; only the association of the per-rank partial sums differs from one GPU).
Also exercises ptb_multi_* (one process, all devices) on rank 0 when it can see more than one device.  Prints one JSON line."""
import json
import os
import pathlib
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from path_trace_golang_b200 import dist as pdist  # noqa: E402
from path_trace_golang_b200 import engine, scene  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = engine.Context(local)
    out = {"world": world, "cases": []}
    peers = pdist.PeerGroup(ctx, 3840, 2160)
    for name, W, H, spp, depth in [("metal_glass_room", 3840, 2160, 8, 16), ("metal_glass_room", 1920, 1080, 32, 16), ("test_scene", 333, 211, 7, 10), ("example_simple", 64, 64, 1, 8)]:
        sc = scene.Load(ROOT / "scenes" / f"{name}.json")
        ctx.upload(sc)
        cfg = ctx.cfg(W, H, spp, depth, seed=9)
        img_peer = pdist.render_distributed_peer(ctx, cfg, peers)
        img_peer = img_peer.cpu().numpy().copy() if rank == 0 else None
        chunk = pdist.slice_pixels(W * H, world)
        accum_padded = torch.zeros((world * chunk, 3), dtype=torch.float32, device=dev)
        rgba_padded = torch.zeros((world * chunk, 4), dtype=torch.uint8, device=dev) if rank == 0 else None
        img_sc = pdist.render_distributed_scatter(ctx, cfg, accum_padded, rgba_padded)
        img_sc = img_sc.cpu().numpy().copy() if rank == 0 else None
        _, img_red = pdist.render_distributed(ctx, cfg)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            pdist.render_distributed_peer(ctx, cfg, peers)
        torch.cuda.synchronize()
        t_peer = (time.perf_counter() - t0) / 5
        # the exchange alone (ranks aligned by a barrier first): fused peer kernel vs NCCL reduce + epilogue on rank 0
        stream = torch.cuda.current_stream(dev).cuda_stream
        acc = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
        rgba = torch.empty((H, W, 4), dtype=torch.uint8, device=dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        t_ex_peer = t_ex_reduce = 0.0
        for _ in range(5):
            torch.cuda.synchronize(); dist.barrier()
            ev[0].record(); peers.reduce_finalize(W, H, spp, stream); ev[1].record()
            torch.cuda.synchronize(); dist.barrier()
            ev[2].record(); pdist.reduce_to_root(acc)
            if rank == 0:
                ctx.finalize_device(acc.data_ptr(), W, H, spp, rgba.data_ptr(), stream)
            ev[3].record()
            torch.cuda.synchronize()
            t_ex_peer, t_ex_reduce = ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])
        tt = torch.tensor([t_ex_peer, t_ex_reduce], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_ex_peer, t_ex_reduce = tt.tolist()
        if rank == 0:
            img_red = img_red.cpu().numpy()
            one = ctx.render(cfg)
            d = lambda a, b: int(np.abs(a.astype(np.int16) - b.astype(np.int16)).max())
            frac = lambda a, b: float((a != b).any(axis=2).mean())
            out["cases"].append({"scene": name, "size": [W, H, spp], "peer_vs_one_max": d(img_peer, one), "peer_vs_one_frac": frac(img_peer, one),
                                 "peer_vs_scatter_max": d(img_peer, img_sc), "peer_vs_reduce_max": d(img_peer, img_red), "peer_ms": t_peer * 1e3, "exchange_only_peer_ms": t_ex_peer, "exchange_only_nccl_reduce_ms": t_ex_reduce,
                                 "alpha_ok": bool((img_peer[..., 3] == 255).all())})
    peers.check()
    peers.close()
    dist.barrier()
    if rank == 0 and torch.cuda.device_count() >= 2:
        # one process driving every device (what a Go host owning the box does): after the ranks are done with their GPUs
        sc = scene.Load(ROOT / "scenes" / "metal_glass_room.json")
        ctx.upload(sc)
        cfg = ctx.cfg(1920, 1080, 32, 16, seed=9)
        one = ctx.render(cfg)
        m = engine.MultiContext(torch.cuda.device_count())
        m.upload(sc)
        img = m.render(cfg)
        t0 = time.perf_counter()
        for _ in range(3):
            m.render(cfg)
        dt = (time.perf_counter() - t0) / 3
        out["multi"] = {"devices": torch.cuda.device_count(), "vs_one_max": int(np.abs(img.astype(np.int16) - one.astype(np.int16)).max()),
                        "vs_one_frac": float((img != one).any(axis=2).mean()), "wall_ms": dt * 1e3, **m.last_timing()}
        m.close()
    if rank == 0:
        ok = all(c["peer_vs_one_max"] <= 1 and c["peer_vs_one_frac"] < 1e-3 and c["peer_vs_scatter_max"] <= 1 and c["peer_vs_reduce_max"] <= 1 and c["alpha_ok"]
                 for c in out["cases"]) and out.get("multi", {"vs_one_max": 0})["vs_one_max"] <= 1
        out["ok"] = bool(ok)
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0)


if __name__ == "__main__":
    main()
