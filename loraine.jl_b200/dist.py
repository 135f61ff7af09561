"""Multi-GPU plumbing on the host: one process per GPU (torchrun), NCCL communicator inside the library.

`torch.distributed` is only used to hand the 128-byte NCCL unique id from rank 0 to the other ranks (any backend: nccl on
the GPU box, gloo in the CPU tests); the data path (panel broadcasts of the distributed Cholesky) runs inside the library.
"""
from __future__ import annotations

import ctypes as C


def panel_owner(col: int, pw: int, world: int) -> int:
    """1-D block-cyclic ownership of Schur-matrix columns (mirrors lrn::ColOwner::owns in csrc/ops.cuh)."""
    return (col // pw) % world if world > 1 else 0


def panel_width(n_var: int) -> int:
    """mirrors lrn_dist_init (csrc/dist.cu)"""
    return 512 if n_var >= 16384 else (256 if n_var >= 4096 else 128)


def exchange_unique_id(make_id, rank: int, world: int) -> bytes:
    """rank 0 calls make_id() -> 128 bytes; every rank returns the same bytes."""
    import torch.distributed as dist
    obj = [make_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(obj, src=0)
    if not isinstance(obj[0], (bytes, bytearray)) or len(obj[0]) != 128:
        raise RuntimeError("NCCL unique id exchange failed")
    return bytes(obj[0])


def init_distributed(solver) -> bool:
    """Attach the solver's device handle to the job's NCCL communicator (call after setup_solver).  Returns True when the
    Schur assembly / factorisation is sharded over more than one rank."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() <= 1:
        return False
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = solver.lib

    def make_id():
        buf = C.create_string_buffer(128)
        rc = lib.lrn_dist_unique_id(buf)
        if rc != 0:
            raise RuntimeError(f"lrn_dist_unique_id failed ({rc})")
        return buf.raw
    uid = exchange_unique_id(make_id, rank, world)
    buf = C.create_string_buffer(uid, 128)
    solver._call("lrn_dist_init", rank, world, buf)
    solver.dist_rank, solver.dist_world = rank, world
    return True
