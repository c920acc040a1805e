"""N > 1 host-side logic on CPU: sample-range partition, reduce-to-root (gloo, world_size 2 and 3), epilogue.
The per-rank "renderer" here is the CPU oracle (test infrastructure) — the plumbing under test is
path_trace_golang_b200.dist, which the GPU path uses unchanged with NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, scene_path


def test_sample_range_partition():
    from path_trace_golang_b200.dist import sample_range
    for spp in (1, 2, 7, 16, 256, 1000):
        for world in (1, 2, 3, 4, 8):
            ranges = [sample_range(spp, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == spp
            for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
                assert a1 == b0 and a1 >= a0
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sample_range(8, 2, 2)


def finalize_host(rgb_sum: np.ndarray, spp: int) -> np.ndarray:
    """Pixel epilogue of renderer.go:189-221 in numpy — a TEST stand-in for ptb_finalize_device, so that the
    reduce plumbing can be exercised on a CPU-only box."""
    v = np.sqrt(rgb_sum.astype(np.float64) * (1.0 / spp)) * 255.999
    v = np.where(v < 0, 0.0, np.where(v > 255.999, 255.999, v))
    v = np.where(np.isnan(v), 0.0, v)
    out = np.empty(rgb_sum.shape[:2] + (4,), dtype=np.uint8)
    out[..., :3] = v.astype(np.uint8)
    out[..., 3] = 255
    return out


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, name, W, H, spp, depth, out_path):
    import sys
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle
    from path_trace_golang_b200 import dist as pdist
    ora = pyoracle.OracleScene.load(scene_path(name))
    b, e = pdist.sample_range(spp, rank, world)
    if e > b:
        part, _ = ora.render_sum(W, H, e - b, depth, seed=5, precision=32, threads=2, s_begin=b)
    else:
        part = np.zeros((H, W, 3))
    accum = torch.from_numpy(part.astype(np.float32))
    pdist.reduce_to_root(accum)
    if rank == 0:
        np.save(out_path, finalize_host(accum.numpy(), spp))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,spp", [(2, 8), (3, 2)])
def test_partition_reduce_epilogue_gloo(world, spp, tmp_path, oracle_mod):
    """world ranks each trace their sample range; the reduced buffer + epilogue on rank 0 equals the single-process
    image (fp32 partial sums re-associated: at most 1 LSB on a handful of channel values)."""
    name, W, H, depth = "example_simple", 96, 54, 8
    out = tmp_path / "img.npy"
    port = _free_port()
    mp.spawn(_worker, args=(world, port, name, W, H, spp, depth, str(out)), nprocs=world, join=True)
    img = np.load(out)
    ora = oracle_mod.OracleScene.load(scene_path(name))
    full, _ = ora.render_sum(W, H, spp, depth, seed=5, precision=32, threads=2)
    ref = oracle_mod.finalize(full, spp)
    diff = np.abs(img.astype(int) - ref.astype(int))
    assert diff.max() <= 1 and (diff > 0).mean() < 0.01
    assert (img[..., 3] == 255).all()


def test_finalize_host_matches_oracle_epilogue(oracle_mod):
    rng = np.random.default_rng(1)
    sums = (rng.random((31, 17, 3)) ** 3 * 40).astype(np.float32)
    sums[0, 0] = [0.0, 1e-9, 1e6]
    assert (finalize_host(sums, 7) == oracle_mod.finalize(sums.astype(np.float64), 7)).all()


def test_row_partition_assembles_the_frame():
    """Interleaved-row partition (dist.rows_of_rank / assemble_rows): the ranks' compact row sets tile the frame."""
    from path_trace_golang_b200.dist import assemble_rows, rows_of_rank
    for H in (2, 3, 7, 90, 1080):
        for world in (1, 2, 3, 8):
            assert sum(rows_of_rank(H, r, world) for r in range(world)) == H
            full = torch.arange(H * 5 * 4, dtype=torch.int64).reshape(H, 5, 4)
            rows_max = rows_of_rank(H, 0, world)
            parts = []
            for r in range(world):
                p = torch.full((rows_max, 5, 4), -1, dtype=torch.int64)      # padded like the all_gather buffer
                mine = full[r::world]
                p[:mine.shape[0]] = mine
                parts.append(p)
            assert torch.equal(assemble_rows(parts, H), full)


def _rows_worker(rank, world, port, H, W, out_path):
    import sys
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from path_trace_golang_b200 import dist as pdist
    full = (torch.arange(H * W * 4, dtype=torch.int64) % 251).to(torch.uint8).reshape(H, W, 4)    # what a renderer would produce
    rows_max = pdist.rows_of_rank(H, 0, world)
    mine = torch.zeros((rows_max, W, 4), dtype=torch.uint8)
    part = full[rank::world]
    mine[:part.shape[0]] = part
    gathered = torch.empty((world * rows_max, W, 4), dtype=torch.uint8)
    dist.all_gather_into_tensor(gathered, mine)
    img = pdist.assemble_rows(list(gathered.view(world, rows_max, W, 4)), H)
    ok = torch.equal(img, full)
    if rank == 0:
        np.save(out_path, np.array([int(ok)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_partition_gather_gloo(world, tmp_path):
    out = tmp_path / "ok.npy"
    mp.spawn(_rows_worker, args=(world, _free_port(), 37, 16, str(out)), nprocs=world, join=True)
    assert np.load(out)[0] == 1


def test_slice_pixels_tiles_the_frame():
    from path_trace_golang_b200.dist import slice_pixels
    for n_pix in (4, 37 * 23, 1920 * 1080, 3840 * 2160, 7680 * 4320 + 1):
        for world in (1, 2, 3, 8):
            c = slice_pixels(n_pix, world)
            assert c % 4 == 0 and world * c >= n_pix and (world == 1 or (world - 1) * c < n_pix + 4 * world)


def _scatter_worker(rank, world, port, H, W, spp, out_path):
    import sys
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from path_trace_golang_b200 import dist as pdist
    n_pix = H * W
    chunk = pdist.slice_pixels(n_pix, world)
    rng = np.random.default_rng(100 + rank)
    part = (rng.random((n_pix, 3)) ** 2 * spp / world).astype(np.float32)          # this rank's partial sums
    padded = torch.zeros((world * chunk, 3), dtype=torch.float32)
    padded[:n_pix] = torch.from_numpy(part)
    rgba = torch.zeros((world * chunk, 4), dtype=torch.uint8) if rank == 0 else None
    fin = lambda sums: torch.from_numpy(finalize_host(sums.numpy().reshape(1, -1, 3), spp).reshape(-1, 4))
    img = pdist.exchange_scatter(padded, rgba, n_pix, fin)
    np.save(f"{out_path}.part{rank}.npy", part)
    if rank == 0:
        np.save(out_path, img.numpy().reshape(H, W, 4))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_reduce_scatter_epilogue_gather_gloo(world, tmp_path):
    """dist.exchange_scatter (reduce_scatter -> per-rank epilogue of its slice -> gather to rank 0) equals "sum everything on one
    rank, then the epilogue", for a frame whose pixel count does not divide by the world size."""
    H, W, spp = 23, 37, 8
    out = str(tmp_path / "img.npy")
    mp.spawn(_scatter_worker, args=(world, _free_port(), H, W, spp, out), nprocs=world, join=True)
    total = np.zeros((H * W, 3), dtype=np.float32)
    for r in range(world):                                  # gloo's ring order is not ours: compare within 1 LSB
        total = total + np.load(f"{out}.part{r}.npy")
    ref = finalize_host(total.reshape(H, W, 3), spp)
    img = np.load(out)
    diff = np.abs(img.astype(int) - ref.astype(int))
    assert diff.max() <= 1 and (diff > 0).mean() < 0.01 and (img[..., 3] == 255).all()
