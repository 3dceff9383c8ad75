"""Host-side logic of the one-process-per-GPU mode: rank r owns sample r -- the reference's m_batch
index (src/onnxstream.cpp:3040-3050, 3847) -- there is no activation exchange, and the only collective is the NCCL broadcast
of each streamed weight block, whose communicator is bootstrapped with a unique id carried over torch.distributed."""
from __future__ import annotations

from typing import Callable, Dict

import numpy as np

from . import emit


def exchange_unique_id(dist, rank: int, make_id: Callable[[], bytes]) -> bytes:
    """Rank 0 creates the 128-byte ncclUniqueId; everyone receives it through the already-initialised process group."""
    obj = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    assert isinstance(obj[0], (bytes, bytearray)) and len(obj[0]) == 128
    return bytes(obj[0])


def rank_inputs(cfg: "emit.UNetConfig", rank: int) -> Dict[str, np.ndarray]:
    """Independent diffusion samples: seed + rank (src/sd.cpp:2671)."""
    return emit.unet_inputs(cfg, seed=rank)


def aggregate_steps_per_sec(dist, world: int, steps: int, seconds: float) -> float:
    """Whole-job throughput under weak scaling: every rank did `steps` steps; time = max over ranks."""
    import torch
    t = torch.tensor([seconds], dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return world * steps / float(t.item())
