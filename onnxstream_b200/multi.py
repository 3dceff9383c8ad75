"""Host-side logic of the one-process-per-GPU mode: rank r owns sample r -- the reference's m_batch
index (src/onnxstream.cpp:3040-3050, 3847) -- there is no activation exchange, and the only collectives are those of the streamed
weight blocks (all-gather of per-rank slices, or a broadcast), whose NCCL communicator lives inside the engine library and is
bootstrapped with a unique id carried over the launcher's torch.distributed group.  bench.py (make_comm) calls exchange_unique_id;
tests/test_cpu.py runs bench.make_comm / dist_max / dist_bcast_array over gloo with world_size 2."""
from __future__ import annotations

from typing import Callable


def exchange_unique_id(dist, rank: int, make_id: Callable[[], bytes]) -> bytes:
    """Rank 0 creates the 128-byte ncclUniqueId; everyone receives it through the already-initialised process group."""
    obj = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    assert isinstance(obj[0], (bytes, bytearray)) and len(obj[0]) == 128
    return bytes(obj[0])
