"""Sharded Monte-Carlo BLER sweep: the multi-GPU form of the reference's BLER loop
(PolarCode.cpp:658-785) without its sequential early-stop shortcuts.

Codewords are independent, so the global batch is block-partitioned over ranks with no data-path
collective; only the integer (num_err, num_run) counters per (list size, Eb/N0) cell cross ranks,
in one all-reduce (NCCL on GPUs, gloo in the CPU tests). Integer sums are order independent, so
the reduced table is identical for every world size.
"""
import numpy as np

from . import synth


def shard_range(total, rank, world):
    """contiguous split of `total` codewords (a multiple of synth.BLOCK) in BLOCK units"""
    blocks = total // synth.BLOCK
    lo = blocks * rank // world
    hi = blocks * (rank + 1) // world
    return lo * synth.BLOCK, (hi - lo) * synth.BLOCK


def sweep_counts(code, decode_fn, list_sizes, ebno_db, total, seed, rank=0, world=1):
    """Local counters int64 [len(list_sizes)][len(ebno_db)][2] = (num_err, num_run) for this rank's
    shard. decode_fn(llr [B][N] f32, L) -> [B][K] uint8 bits (the GPU decoder in production)."""
    counts = np.zeros((len(list_sizes), len(ebno_db), 2), np.int64)
    first, count = shard_range(total, rank, world)
    if count == 0:
        return counts
    for ie, eb in enumerate(ebno_db):
        info, llr = synth.make_shard(code, seed + ie, first, count, ebno_db=eb)
        for il, L in enumerate(list_sizes):
            dec = decode_fn(llr, int(L))
            counts[il, ie, 0] = int((dec != info).any(axis=1).sum())
            counts[il, ie, 1] = count
    return counts


def sweep_counts_device(code, list_sizes, ebno_db, total, seed, rank=0, world=1, mode=None):
    """Same counters with everything on the GPU (a binding of polar_b200_bler_sweep): `total` codewords PER Eb/N0 point;
    the interleaved index space [0, total * len(ebno_db)) is split over ranks in contiguous blocks, every rank synthesises,
    decodes and compares its block on its device, only the counters come back."""
    n_pts = len(ebno_db)
    all_cw = total * n_pts
    lo = all_cw * rank // world
    hi = all_cw * (rank + 1) // world
    return code.bler_sweep_device(ebno_db, list_sizes, hi - lo, seed, first_index=lo, mode=mode)


class Comm:
    """The C++ side's NCCL communicator (polar_b200_comm_*), one per process and GPU: the collective of a sharded sweep is
    ncclAllReduce over the int64 counters. Bootstrap: rank 0's unique id reaches the other ranks through `exchange`,
    a callable bytes -> bytes that broadcasts rank 0's value (e.g. over torch.distributed)."""

    def __init__(self, device, world, rank, exchange):
        import ctypes as C
        from . import _lib
        lib = _lib.dev()
        buf = (C.c_ubyte * 128)()
        if rank == 0:
            _lib.check(lib.polar_b200_comm_unique_id(buf))
        ident = exchange(bytes(buf))
        self._h = C.c_void_p()
        _lib.check(lib.polar_b200_comm_init_rank(C.byref(self._h), int(device), int(world), int(rank),
                                                 (C.c_ubyte * 128).from_buffer_copy(ident)))

    def all_reduce(self, counts):
        from . import _lib
        out = np.ascontiguousarray(counts, np.int64).copy()
        _lib.check(_lib.dev().polar_b200_comm_allreduce_i64(self._h, out.ctypes.data, out.size))
        return out

    def close(self):
        from . import _lib
        if self._h:
            _lib.dev().polar_b200_comm_destroy(self._h)
            self._h = None


def all_reduce_counts(counts, device=None):
    """Sum the counters over ranks (no-op without an initialised process group)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return counts
    t = torch.from_numpy(counts.copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def bler_table(counts):
    with np.errstate(divide="ignore", invalid="ignore"):
        return counts[..., 0] / counts[..., 1]
