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


def render_rows_distributed(ctx, cfg_full, group=None, to_all: bool = False):
    """Whole multi-GPU frame by interleaved rows (image tiles): render (fused epilogue) -> gather of the ranks' compact RGBA8 row
    sets to rank 0 -> assemble there.  4 bytes per pixel cross the links in total.  Returns the CUDA uint8 (H, W, 4) image on rank
    0 and None elsewhere (to_all=True: all_gather, every rank assembles the frame).  Everything is enqueued on torch's current stream."""
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
    if to_all:
        gathered = torch.empty((world * rows_max, W, 4), dtype=torch.uint8, device=dev)      # rank-major concatenation
        dist.all_gather_into_tensor(gathered, mine, group=group)
        return assemble_rows(list(gathered.view(world, rows_max, W, 4)), H)
    parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, parts, dst=0, group=group)
    return assemble_rows(parts, H) if rank == 0 else None


# ---- the fused exchange (default on GPUs): ONE kernel per rank does reduce-scatter + pixel epilogue + gather over NVLink peer
# memory, cross-rank ordering included (flag words in peer memory; ptb_peer_*, include/ptb200.h).  torch.distributed is the
# plumbing only: it carries the 64-byte CUDA IPC handles ONCE; no collective runs per frame.

class _DevPtr:
    """A raw device pointer dressed up for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr, "version": 3}


class PeerGroup:
    """One rank's end of the multi-process peer group: its fp32 sum buffer, rank 0's image, and the fused kernel."""

    def __init__(self, ctx, max_width: int, max_height: int, group=None):
        import ctypes as C

        from . import _lib
        self._L, self._ctx, self._group = _lib.lib(), ctx, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.max_width, self.max_height = int(max_width), int(max_height)
        h = C.c_void_p()
        ctx._check(self._L.ptb_peer_create(ctx._h, self.rank, self.world, self.max_width, self.max_height, C.byref(h)))
        self._h = h
        dev = torch.device("cuda", ctx.device)
        if self.world > 1:
            mine = (C.c_ubyte * 192)()
            ctx._check(self._L.ptb_peer_handles(self._h, mine, C.byref(mine, 64), C.byref(mine, 128)))
            t = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
            allh = torch.empty(self.world * 192, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allh, t, group=group)
            allh = allh.cpu().numpy().reshape(self.world, 192)
            accum = (C.c_ubyte * (64 * self.world))(*allh[:, :64].reshape(-1).tolist())
            image = (C.c_ubyte * 64)(*allh[0, 64:128].tolist())
            flags = (C.c_ubyte * (64 * self.world))(*allh[:, 128:].reshape(-1).tolist())
            ctx._check(self._L.ptb_peer_connect(self._h, accum, image, flags))
            dist.barrier(group=group)                      # nobody renders before every mapping exists

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            if self.world > 1 and dist.is_initialized():
                torch.cuda.synchronize()
                dist.barrier(group=self._group)            # no rank unmaps while another still reads its buffer
            self._L.ptb_peer_destroy(h)

    @property
    def accum_ptr(self) -> int:
        return int(self._L.ptb_peer_accum(self._h))

    def accum_tensor(self, height: int, width: int) -> torch.Tensor:
        return torch.as_tensor(_DevPtr(self.accum_ptr, (height, width, 3), "<f4"), device=torch.device("cuda", self._ctx.device))

    def image_tensor(self, height: int, width: int):
        """Rank 0: the assembled RGBA8 frame (a view of the peer image buffer); None on the other ranks."""
        p = self._L.ptb_peer_image(self._h)
        if not p:
            return None
        return torch.as_tensor(_DevPtr(int(p), (height, width, 4), "|u1"), device=torch.device("cuda", self._ctx.device))

    def reduce_finalize(self, width: int, height: int, spp_total: int, stream: int = 0):
        """Stream-ordered behind this rank's render: announce it, wait for every rank's announcement, reduce + finalise + store this
        rank's slice, wait until every rank has done so (all of it inside the library's kernels: flag words in peer memory)."""
        self._ctx._check(self._L.ptb_peer_reduce_finalize(self._h, int(width), int(height), int(spp_total), stream))

    def check(self):
        """Raises if a wait of the exchange timed out (a rank never arrived)."""
        self._ctx._check(self._L.ptb_peer_status(self._h))


def render_distributed_peer(ctx, cfg_full, peers: PeerGroup):
    """Whole multi-GPU frame with the fused exchange: partition -> render into the peer buffer -> reduce_finalize.
    Returns the CUDA uint8 (H, W, 4) image on rank 0 (a view of the peer image buffer), None elsewhere."""
    from ._lib import PtbCfg
    H, W = cfg_full.height, cfg_full.width
    stream = torch.cuda.current_stream(torch.device("cuda", ctx.device)).cuda_stream
    b, e = sample_range(cfg_full.samples_per_px, peers.rank, peers.world)
    if e > b:
        cfg = PtbCfg(W, H, cfg_full.samples_per_px, cfg_full.max_depth, cfg_full.seed, b, e - b, cfg_full.flags, 0, 0)
        ctx.render_accum_device(cfg, peers.accum_ptr, stream)
    else:
        peers.accum_tensor(H, W).zero_()
    peers.reduce_finalize(W, H, cfg_full.samples_per_px, stream)
    return peers.image_tensor(H, W)


# ---- the same data movement with NCCL collectives only (reduce_scatter -> per-rank epilogue of its slice -> gather of RGBA8):
# the library-call baseline of the fused kernel above, and the path for groups without CUDA IPC (e.g. ranks on several hosts).

def slice_pixels(n_pix: int, world: int) -> int:
    """Pixels per rank slice: ceil(n_pix / world) rounded up to a multiple of 4 (the slices tile [0, world * chunk) >= n_pix)."""
    return ((n_pix + world - 1) // world + 3) // 4 * 4


def exchange_scatter(accum_padded: torch.Tensor, rgba_padded, n_pix: int, finalize, group=None):
    """reduce_scatter of the padded sums -> finalize(slice sums (chunk, 3)) -> (chunk, 4) uint8 on every rank -> gather into rank 0's
    rgba_padded (world * chunk, 4).  Returns rank 0's first n_pix pixels (a view), None elsewhere."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    chunk = slice_pixels(n_pix, world)
    assert accum_padded.shape == (world * chunk, 3)
    if world == 1:
        return finalize(accum_padded)[:n_pix]
    mine = torch.empty((chunk, 3), dtype=accum_padded.dtype, device=accum_padded.device)
    dist.reduce_scatter_tensor(mine, accum_padded, op=dist.ReduceOp.SUM, group=group)
    my_rgba = finalize(mine)
    dist.gather(my_rgba, list(rgba_padded.view(world, chunk, 4)) if rank == 0 else None, dst=0, group=group)
    return rgba_padded[:n_pix] if rank == 0 else None


def render_distributed_scatter(ctx, cfg_full, accum_padded: torch.Tensor, rgba_padded: torch.Tensor | None, group=None):
    """accum_padded: CUDA fp32 (world * chunk, 3) with chunk = slice_pixels(H * W, world), zero beyond H * W; rgba_padded: rank 0's
    CUDA uint8 (world * chunk, 4).  Returns rank 0's (H, W, 4) view of rgba_padded, None elsewhere."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    H, W = cfg_full.height, cfg_full.width
    n_pix = H * W
    dev = torch.device("cuda", ctx.device)
    stream = torch.cuda.current_stream(dev).cuda_stream
    render_partition(ctx, cfg_full, rank, world, accum_padded[:n_pix].view(H, W, 3), stream)

    def finalize(sums):
        out = torch.empty((sums.shape[0], 4), dtype=torch.uint8, device=dev)
        ctx.finalize_device(sums.data_ptr(), sums.shape[0], 1, cfg_full.samples_per_px, out.data_ptr(), stream)
        return out

    img = exchange_scatter(accum_padded, rgba_padded, n_pix, finalize, group)
    return img.view(H, W, 4) if img is not None else None
