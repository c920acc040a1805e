"""path_trace_golang_b200 — B200 (sm_100a) CUDA backend for the per-pixel render loop of
MarkJulian19/path_trace_golang, behind the reference's scene / engine API names.

  scene   Load / Save / Scene / RenderSettings          (internal/scene)
  engine  RenderConfig / Render / RenderInto / ...      (internal/engine), Context = one ptb_ctx
  dist    sample-range partition + NCCL reduce          (multi-GPU, one process per GPU)
"""
from . import _lib, engine, scene  # noqa: F401
from ._lib import PtbError  # noqa: F401

__all__ = ["scene", "engine", "PtbError"]
