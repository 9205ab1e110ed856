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


def sweep_counts_device(code, list_sizes, ebno_db, total, seed, rank=0, world=1, chunk=65536):
    """Same counters with everything on the GPU: codewords are synthesised on the device
    (polar_b200_synthesize), decoded and compared there; only the counters come back.
    `total` codewords PER Eb/N0 point, split over ranks; point ie uses seed + ie."""
    import torch
    counts = np.zeros((len(list_sizes), len(ebno_db), 2), np.int64)
    lo = total * rank // world
    hi = total * (rank + 1) // world
    dev = torch.device("cuda", code.device)
    nerr = torch.zeros(1, dtype=torch.int64, device=dev)
    for ie, eb in enumerate(ebno_db):
        for first in range(lo, hi, chunk):
            nb = min(chunk, hi - first)
            llr, truth = code.synthesize(nb, [eb], seed + ie, first_index=first)
            for il, L in enumerate(list_sizes):
                out = code.decode_device(llr, int(L))
                nerr.zero_()
                code.count_errors(out, truth, None, nerr)
                counts[il, ie, 0] += int(nerr.item())
                counts[il, ie, 1] += nb
    return counts


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
