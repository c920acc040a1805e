"""Multi-GPU: one process per GPU (torch.distributed), the frame's samples partitioned across ranks.

Every (pixel, sample) path is independent and the counter RNG is keyed by (seed, pixel, sample), so rank r
simply traces samples [r*spp/N, (r+1)*spp/N) of every pixel into its own fp32 accumulation buffer; the only
exchange is the additive reduction of those buffers to rank 0 (NCCL over NVLink on GPUs, gloo in CPU tests),
followed by the pixel epilogue on rank 0.  The image is identical (up to fp32 re-association of the N
partial sums) for every N.  SURVEY.md §8(e).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def sample_range(spp: int, rank: int, world: int) -> tuple[int, int]:
    """Samples [begin, end) of every pixel traced by `rank`; ranges tile [0, spp) and differ by <= 1."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(int(spp), world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def reduce_to_root(accum: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the per-rank accumulation buffers onto rank 0, in place (ncclReduce / gloo reduce)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM, group=group)
    return accum


def render_partition(ctx, cfg_full, rank: int, world: int, accum: torch.Tensor, stream: int = 0) -> bool:
    """Launch this rank's share of the frame into `accum` (CUDA fp32 tensor, H*W*3).  Returns False when the
    rank has no samples (spp < world): its buffer is zero-filled instead."""
    from ._lib import PtbCfg
    b, e = sample_range(cfg_full.samples_per_px, rank, world)
    if e <= b:
        accum.zero_()
        return False
    cfg = PtbCfg(cfg_full.width, cfg_full.height, cfg_full.samples_per_px, cfg_full.max_depth, cfg_full.seed, b, e - b,
                 cfg_full.flags, 0, 0)
    ctx.render_accum_device(cfg, accum.data_ptr(), stream)
    return True


def render_distributed(ctx, cfg_full, out_rgba: torch.Tensor | None = None, group=None):
    """Whole multi-GPU frame: partition -> render -> reduce -> epilogue on rank 0.

    Returns (accum, rgba): rgba is a CUDA uint8 (H, W, 4) tensor on rank 0, None elsewhere.
    Everything is enqueued on torch's current stream; nothing synchronises with the host.
    """
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    H, W = cfg_full.height, cfg_full.width
    dev = torch.device("cuda", ctx.device)
    accum = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    render_partition(ctx, cfg_full, rank, world, accum, stream)
    reduce_to_root(accum, group)
    rgba = None
    if rank == 0:
        rgba = out_rgba if out_rgba is not None else torch.empty((H, W, 4), dtype=torch.uint8, device=dev)
        ctx.finalize_device(accum.data_ptr(), W, H, cfg_full.samples_per_px, rgba.data_ptr(), stream)
    return accum, rgba


# ---- the other partition SURVEY §8(e) allows: image tiles.  Rank r renders rows r, r+N, r+2N, ... of the frame at
# the full sample count (interleaved rows balance the load: neighbouring rows cost the same), with the pixel epilogue
# fused into the integrator; the only exchange is a gather of the ranks' compact RGBA8 row sets (4 bytes per pixel in
# total, against 12 bytes per pixel and rank for the fp32 reduce), and because every pixel is computed exactly as in a
# whole-frame render the assembled image equals the single-GPU image bit for bit.

def rows_of_rank(height: int, rank: int, world: int) -> int:
    """Number of rows rank, rank+world, ... below `height`."""
    return max(0, (int(height) - rank + world - 1) // world)


def assemble_rows(parts, height: int) -> torch.Tensor:
    """parts[r]: (>= rows_of_rank(height, r, N), W, C) tensor of rank r's rows -> the (height, W, C) frame."""
    world = len(parts)
    out = parts[0].new_empty((height,) + tuple(parts[0].shape[1:]))
    for r, p in enumerate(parts):
        n = rows_of_rank(height, r, world)
        if n:
            out[r::world] = p[:n]
    return out


def render_rows_distributed(ctx, cfg_full, group=None):
    """Whole multi-GPU frame by interleaved rows: render (fused epilogue) -> all_gather -> assemble on every rank.

    Returns the CUDA uint8 (H, W, 4) image.  Everything is enqueued on torch's current stream."""
    from ._lib import PtbCfg
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    H, W = cfg_full.height, cfg_full.width
    dev = torch.device("cuda", ctx.device)
    stream = torch.cuda.current_stream(dev).cuda_stream
    rows_max = rows_of_rank(H, 0, world)
    mine = torch.zeros((rows_max, W, 4), dtype=torch.uint8, device=dev)
    if rows_of_rank(H, rank, world) > 0:
        cfg = PtbCfg(cfg_full.width, cfg_full.height, cfg_full.samples_per_px, cfg_full.max_depth, cfg_full.seed, 0, 0,
                     cfg_full.flags, rank if world > 1 else 0, world if world > 1 else 0)
        ctx.render_device(cfg, mine.data_ptr(), stream)
    if world == 1:
        return mine
    gathered = torch.empty((world * rows_max, W, 4), dtype=torch.uint8, device=dev)      # rank-major concatenation
    dist.all_gather_into_tensor(gathered, mine, group=group)
    return assemble_rows(list(gathered.view(world, rows_max, W, 4)), H)
