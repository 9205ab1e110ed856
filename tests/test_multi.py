"""Multi-rank logic on CPU (gloo, world_size 2): the batch shards over ranks with no data-path
collective and the all-reduced (num_err, num_run) counters equal the single-rank counters exactly.
The decoder plugged in here is the CPU oracle -- a stand-in so the sharding / reduction code can
run without GPUs; on the GPU box bench.py runs the same code with the CUDA decoder."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle_lib import Port


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from polar_b200 import bler
    code = Port(7, 64, 0.32, 8)
    local = bler.sweep_counts(code, lambda llr, L: code.decode_batch(llr, L), [1, 4], [1.0, 3.0], total, 99, rank, world)
    red = bler.all_reduce_counts(local)
    if rank == 0:
        ret["counts"] = red
        ret["local0"] = local
    dist.destroy_process_group()


def test_shard_ranges_tile_the_batch():
    from polar_b200 import bler, synth
    for world in (1, 2, 3, 4, 8):
        total = 16 * synth.BLOCK
        spans = [bler.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        for (a, ca), (b, _) in zip(spans, spans[1:]):
            assert a + ca == b


def test_synth_is_independent_of_sharding():
    from polar_b200 import synth
    code = Port(6, 32, 0.32, 4)
    i_all, l_all = synth.make_shard(code, 5, 0, 4 * synth.BLOCK)
    i_hi, l_hi = synth.make_shard(code, 5, 2 * synth.BLOCK, 2 * synth.BLOCK)
    assert np.array_equal(i_all[2 * synth.BLOCK:], i_hi) and np.array_equal(l_all[2 * synth.BLOCK:], l_hi)


@pytest.mark.timeout(300)
def test_two_rank_counters_equal_single_rank():
    from polar_b200 import bler, synth
    total = 4 * synth.BLOCK
    code = Port(7, 64, 0.32, 8)
    single = bler.sweep_counts(code, lambda llr, L: code.decode_batch(llr, L), [1, 4], [1.0, 3.0], total, 99, 0, 1)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), total, ret), nprocs=2, join=True)
    assert np.array_equal(ret["counts"], single)
    assert ret["local0"][0, 0, 1] == total // 2
    assert np.all(ret["counts"][..., 1] == total)
    t = bler.bler_table(ret["counts"])
    assert t[1, 1] <= t[0, 0]
