"""CPU tests of the checkers themselves: the port oracle against the golden fixtures generated
from the unmodified reference, against the compiled reference where it is available, and the
numpy lane-level model of the CUDA kernel's organisation against the port."""
import numpy as np
import pytest

from conftest import (CONSTRUCTIONS, decode_fixture_paths, load_construction, load_decode, load_edge, load_p1,
                      p1_fixture_paths)
from oracle_lib import Port, Ref, awgn_llrs, awgn_probs, edge_probs, have_ref


@pytest.mark.parametrize("n,K,crc", CONSTRUCTIONS)
def test_port_construction_matches_golden(n, K, crc):
    g = load_construction(n, K, crc)
    c = Port(n, K, 0.32, crc).construction()
    for key in ("frozen", "order", "crc_matrix", "bitrev"):
        assert np.array_equal(c[key], g[key]), key


@pytest.mark.parametrize("path", decode_fixture_paths(), ids=lambda p: p.split("decode_")[-1][:-4])
def test_port_decode_matches_golden(path):
    d = load_decode(path)
    port = Port(d["n"], d["K"], 0.32, d["crc"])
    assert np.array_equal(port.encode(d["info"])[:2], Port(d["n"], d["K"], 0.32, d["crc"]).encode(d["info"][:2]))
    got = port.decode_batch(d["llr"], d["L"], nthreads=8)
    assert np.array_equal(got, d["decoded"])


@pytest.mark.parametrize("n,K,crc", [(9, 256, 16), (7, 64, 8)])
@pytest.mark.parametrize("L", [1, 2, 4, 32])
def test_port_edge_cases_match_golden(n, K, crc, L):
    e = load_edge(n, K, crc)
    got = Port(n, K, 0.32, crc).decode_batch(e["llr"], L, nthreads=8)
    assert np.array_equal(got, e[L])


def test_port_from_tables_equals_port_from_construction():
    g = load_construction(9, 256, 16)
    a, b = Port(9, 256, 0.32, 16), Port(9, 256, 0.32, 16, tables=g)
    _, llr = awgn_llrs(a, 16, 1.5, 3)
    assert np.array_equal(a.decode_batch(llr, 8), b.decode_batch(llr, 8))


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("n,K,crc,L,B,eb", [(9, 256, 0, 1, 64, 2.0), (11, 1024, 16, 4, 12, 1.5), (11, 1024, 16, 32, 6, 1.25),
                                            (9, 256, 16, 32, 24, 2.0), (9, 256, 0, 3, 32, 1.0), (6, 20, 3, 5, 64, -1.0),
                                            (8, 100, 7, 127, 8, 1.0)])
def test_port_equals_compiled_reference(n, K, crc, L, B, eb):
    ref, port = Ref(n, K, 0.32, crc), Port(n, K, 0.32, crc)
    cr, cp = ref.construction(), port.construction()
    for key in cr:
        assert np.array_equal(cr[key], cp[key]), key
    info, llr = awgn_llrs(port, B, eb, 4242 + n + L)
    assert np.array_equal(ref.encode(info), port.encode(info))
    assert np.array_equal(ref.decode_batch(llr, L, 8), port.decode_batch(llr, L, 8))


@pytest.mark.parametrize("path", p1_fixture_paths(), ids=lambda p: p.split("p1_")[-1][:-4])
def test_port_probability_domain_matches_golden(path):
    """decode_scl_p1 (PolarCode.cpp:110-128, 375-420): the restatement against fixtures generated from the
    unmodified reference, edge rows (ties, zeros, sigma == 0) included."""
    d = load_p1(path)
    got = Port(d["n"], d["K"], 0.32, d["crc"]).decode_p1_batch(d["p1"], d["p0"], d["L"], nthreads=8)
    assert np.array_equal(got, d["decoded"])


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("n,K,crc,L,B,eb", [(5, 16, 4, 4, 40, 1.0), (9, 256, 16, 32, 8, 1.5), (8, 100, 7, 127, 4, 1.0),
                                            (7, 64, 0, 3, 40, 0.0), (3, 4, 0, 5, 20, 0.0), (10, 512, 8, 2, 8, 2.0)])
def test_port_probability_domain_equals_compiled_reference(n, K, crc, L, B, eb):
    ref, port = Ref(n, K, 0.32, crc), Port(n, K, 0.32, crc)
    _, p1, p0 = awgn_probs(port, B, eb, seed=31 + n + L)
    e1, e0 = edge_probs(1 << n, seed=n)
    p1, p0 = np.concatenate([e1, p1]), np.concatenate([e0, p0])
    want = np.stack([ref.decode_p1_one(p1[b], p0[b], L) for b in range(len(p1))])
    assert np.array_equal(port.decode_p1_batch(p1, p0, L, nthreads=8), want)
    assert np.array_equal(port.decode_p1_one(p1[-1], p0[-1], L), want[-1])


def test_port_bler_harness_counts():
    # tiny deterministic run of the restated harness: shape, monotone counters, bler = err / run
    port = Port(6, 32, 0.32, 0)
    bler, counts = port.get_bler_quick([0.0, 2.0, 4.0], [1, 4], max_err=100, max_runs=50)
    assert bler.shape == (2, 3)
    assert np.all(counts[..., 1] == 50)
    assert np.allclose(bler, counts[..., 0] / counts[..., 1])
    assert bler[0, 0] >= bler[0, 2]


@pytest.mark.parametrize("n,K,crc,L,B,eb,tweak", [(3, 4, 0, 1, 32, 1.0, None), (4, 8, 2, 2, 32, 0.0, "zeros"),
                                                  (5, 16, 4, 4, 32, 0.0, "round"), (6, 32, 8, 8, 16, 0.0, None),
                                                  (7, 64, 8, 32, 3, 1.0, None), (7, 64, 0, 3, 16, 0.0, None),
                                                  (6, 20, 3, 5, 12, -1.0, "round")])
def test_kernel_model_equals_port(n, K, crc, L, B, eb, tweak):
    """The data organisation the CUDA kernel uses (butterfly order, column pointers, packed partial
    sums, stack across lanes, u-hat by polar transform) is exactly the reference algorithm."""
    from kernel_model import decode_group
    port = Port(n, K, 0.32, crc)
    c = port.construction()
    _, llr = awgn_llrs(port, B, eb, 77 + n)
    if tweak == "zeros":
        llr[:4] = 0
    if tweak == "round":
        llr[:8] = np.round(llr[:8])
    want = port.decode_batch(llr, L)
    W = 1
    while W < L:
        W <<= 1
    G = 32 // W
    for s in range(0, B, G):
        got = decode_group(llr[s:s + G], n, K, crc, c["frozen"], c["order"], c["crc_matrix"], L)
        assert np.array_equal(got, want[s:s + G])
