"""Row-sharded multi-GPU runs: one process per GPU, NCCL over NVLink 5 / NVSwitch.

Replaces MultiDeviceKernel (gpytorch/kernels/multi_device_kernel.py:14-95): rank r owns a contiguous row
block of K and of every CG vector; X is replicated.  Per CG iteration the engine all-gathers the [n/g, 16]
direction block and all-reduces the packed fp64 dot-product messages (csrc/cg.cu, csrc/comm.cu).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib


def shard_rows(n: int, world: int, rank: int):
    """Equal row blocks (the all-gather needs equal counts): returns (row_begin, row_count, rows_per_rank).
    n must be divisible by world for the native path; `padded_size` gives the next valid n."""
    if n % world:
        raise ValueError(f"N={n} is not divisible by the world size {world}: the engine's all-gather needs equal row "
                         f"shards (pad the problem to padded_size(n, world) = {padded_size(n, world)} rows)")
    per = n // world
    return rank * per, per, per


def padded_size(n: int, world: int) -> int:
    return ((n + world - 1) // world) * world


class Comm:
    """gp_comm handle: rank 0 draws the NCCL unique id, torch.distributed broadcasts it, every rank inits."""

    def __init__(self, rank: int | None = None, world: int | None = None):
        import torch.distributed as dist

        self.lib = _lib.load()
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        buf = (C.c_uint8 * 128)()
        if self.rank == 0:
            _lib.check(self.lib.gp_comm_unique_id(buf))
        t = torch.tensor(list(buf), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=0)
        ids = (C.c_uint8 * 128)(*t.cpu().tolist())
        self.handle = C.c_void_p()
        _lib.check(self.lib.gp_comm_init(C.byref(self.handle), ids, self.rank, self.world))

    def close(self):
        if self.handle:
            self.lib.gp_comm_destroy(self.handle)
            self.handle = C.c_void_p()


def init_from_env():
    """torchrun entry: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment."""
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local
