"""Synthetic BPSK/AWGN workloads: info bits -> reference encoder -> channel ->
fp32 LLRs, the reference's own recipe (PolarCode.cpp:703-716, 744-753) with a counter-based RNG.

Codewords are generated in blocks of BLOCK; block i of a stream always uses Philox(key=(seed, i)), so
what a codeword looks like depends only on (seed, global codeword index) and not on how the batch is
sharded across ranks.
"""
import numpy as np

BLOCK = 256
SWEEP_DB = (1.0, 1.25, 1.5, 1.75, 2.0, 2.25, 2.5)   # BASELINE.json configs[3]


def ebno_of_codeword(idx, sweep=SWEEP_DB):
    """Eb/N0 (dB) of global codeword idx: the sweep points are interleaved over the batch."""
    return np.asarray(sweep, np.float64)[np.asarray(idx) % len(sweep)]


def make_block(code, seed, block_index, ebno_db=None, sweep=SWEEP_DB):
    """-> (info [BLOCK][K] u8, llr [BLOCK][N] f32) for global codewords block_index*BLOCK ..."""
    rng = np.random.Generator(np.random.Philox(key=[int(seed) & (2**64 - 1), int(block_index)]))
    info = rng.integers(0, 2, size=(BLOCK, code.K), dtype=np.uint8)
    coded = code.encode(info)
    idx = np.arange(block_index * BLOCK, (block_index + 1) * BLOCK)
    eb = np.full(BLOCK, float(ebno_db)) if ebno_db is not None else ebno_of_codeword(idx, sweep)
    a = (10.0 ** (eb / 20.0) * np.sqrt(code.K / code.N))[:, None]          # PolarCode.cpp:744-745
    r = a * (2.0 * coded.astype(np.float64) - 1.0) + np.sqrt(0.5) * rng.standard_normal((BLOCK, code.N))
    llr = (-4.0 * r * a).astype(np.float32)                                 # PolarCode.cpp:752, N0 = 1
    return info, llr


def make_shard(code, seed, first, count, ebno_db=None, sweep=SWEEP_DB, out_llr=None):
    """Global codewords [first, first+count) -> (info [count][K], llr [count][N]); first and count
    must be multiples of BLOCK. out_llr: optional preallocated float32 array (e.g. pinned) to fill."""
    assert first % BLOCK == 0 and count % BLOCK == 0
    info = np.empty((count, code.K), np.uint8)
    llr = out_llr if out_llr is not None else np.empty((count, code.N), np.float32)
    for i in range(count // BLOCK):
        bi, bl = make_block(code, seed, first // BLOCK + i, ebno_db, sweep)
        info[i * BLOCK:(i + 1) * BLOCK] = bi
        llr[i * BLOCK:(i + 1) * BLOCK] = bl
    return info, llr
