"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header
declares, the C++ drop-in class reproduces the reference's construction and encoder, and the
product path refuses to run (loudly) without a GPU instead of falling back to anything."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import CONSTRUCTIONS, ROOT, decode_fixture_paths, load_construction, load_decode
from oracle_lib import Port


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "polar_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(polar_b200_[a-z_0-9]+)\s*\(", txt)))


def test_abi_exports_every_declared_symbol(native_libs):
    dev_so, _ = native_libs
    lib = C.CDLL(dev_so)
    syms = _declared_symbols()
    assert len(syms) >= 9
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s
    lib.polar_b200_abi_version.restype = C.c_int
    hdr = open(os.path.join(ROOT, "include", "polar_b200.h")).read()
    assert lib.polar_b200_abi_version() == int(re.search(r"POLAR_B200_ABI_VERSION (\d+)", hdr).group(1))
    lib.polar_b200_strerror.restype = C.c_char_p
    assert b"no CPU fallback" in lib.polar_b200_strerror(-3)
    assert lib.polar_b200_info_words(1024) == 32 and lib.polar_b200_info_words(33) == 2


def test_abi_argument_validation_needs_no_gpu(native_libs):
    dev_so, _ = native_libs
    lib = C.CDLL(dev_so)
    ctx = C.c_void_p()
    frozen = (C.c_uint8 * 8)(1, 1, 1, 0, 1, 0, 0, 0)
    order = (C.c_uint16 * 4)(7, 6, 5, 3)
    assert lib.polar_b200_create(None, 0, 3, 4, 0, frozen, order, None, 1, 1) == -1
    assert lib.polar_b200_create(C.byref(ctx), 0, 3, 4, 0, None, order, None, 1, 1) == -1
    assert lib.polar_b200_create(C.byref(ctx), 0, 16, 4, 0, frozen, order, None, 1, 1) == -2   # n too large (the reference stops at 15)
    assert lib.polar_b200_create(C.byref(ctx), 0, 3, 4, 0, frozen, order, None, 128, 1) == -2  # list too large (the reference's own limit is 127)
    assert lib.polar_b200_create(C.byref(ctx), 0, 3, 4, 2, frozen, order, None, 1, 1) == -1    # crc without matrix
    assert lib.polar_b200_decode_scl_llr(None, None, 1, 1, None, None) == -1
    assert lib.polar_b200_destroy(None) == -1


@pytest.mark.parametrize("n,K,crc", CONSTRUCTIONS)
def test_host_class_construction_matches_golden(native_libs, n, K, crc):
    from polar_b200 import PolarCode
    g = load_construction(n, K, crc)
    c = PolarCode(n, K, 0.32, crc).construction()
    for key in ("frozen", "order", "crc_matrix", "bitrev"):
        assert np.array_equal(c[key], g[key]), key


def test_host_class_encode_matches_oracle(native_libs):
    from polar_b200 import PolarCode
    for path in decode_fixture_paths()[:4]:
        d = load_decode(path)
        pc, port = PolarCode(d["n"], d["K"], 0.32, d["crc"]), Port(d["n"], d["K"], 0.32, d["crc"])
        assert np.array_equal(pc.encode(d["info"]), port.encode(d["info"]))
        assert np.array_equal(pc.encode(d["info"][0]), port.encode(d["info"][0]))


def test_pack_unpack_roundtrip():
    from polar_b200 import pack_bits, unpack_bits
    rng = np.random.default_rng(1)
    for K in (1, 31, 32, 33, 256, 1000):
        bits = rng.integers(0, 2, (5, K), dtype=np.uint8)
        assert np.array_equal(unpack_bits(pack_bits(bits), K), bits)
    assert pack_bits(np.array([[1, 0, 1]], np.uint8))[0, 0] == 5


def test_no_cpu_fallback_without_gpu(native_libs):
    """On a machine without a CUDA device every decode must raise; nothing may route through the oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from polar_b200 import PolarB200Error, PolarCode
    pc = PolarCode(5, 16, 0.32, 4)
    with pytest.raises(PolarB200Error, match="no CPU fallback"):
        pc.decode_batch(np.zeros((2, 32), np.float32), 4)
    with pytest.raises(PolarB200Error):
        pc.decode_scl_llr(np.zeros(32), 1)
    with pytest.raises(PolarB200Error):
        pc.get_bler_quick([1.0], [1], max_runs=10)


def test_product_sources_do_not_touch_the_oracle():
    """oracle/ is test infrastructure: nothing under polar_b200/ or include/ may mention it."""
    bad = []
    for base in ("polar_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                    p = os.path.join(dp, f)
                    txt = open(p).read()
                    if re.search(r"oracle_lib|libpolar_oracle|libpolar_ref|oracle/_build|kernel_model", txt):
                        bad.append(p)
    assert bad == []


SSC_CODES = [(8, 128, 0), (9, 256, 0), (9, 256, 16), (10, 300, 8), (11, 1024, 0), (11, 1024, 16), (11, 1536, 16), (12, 1000, 0),
             (9, 3, 0), (8, 1, 0), (10, 1000, 0), (11, 2000, 16), (9, 500, 8), (12, 4000, 0)]      # all-frozen / all-unfrozen halves


@pytest.mark.parametrize("n,K,crc", SSC_CODES)
def test_pruned_tree_schedule_decodes_like_the_oracle(n, K, crc):
    """sc_ssc.cuh's host side: the schedule and the output map the library builds for a code, interpreted by the numpy
    model of the kernel (tests/ssc_model.py), decode exactly what the oracle decodes with list size 1."""
    import ssc_model
    from oracle_lib import Port, awgn_llrs
    from polar_b200 import _lib
    port = Port(n, K, 0.32, crc)
    con = port.construction()
    lib = _lib.dev()
    ops = ssc_model.schedule(lib, n, con["frozen"])
    assert ops is not None and not ops[-3:].any()
    pos = ssc_model.positions(lib, n, con["order"], K)
    B = 192 if n <= 10 else 64
    for eb in (1.0, 3.0) if K * 2 <= (1 << n) else (7.0, 9.0):      # (rates near 1 need the SNR: below it the leaf LLRs vanish)
        _, llr = awgn_llrs(port, B, eb, seed=77 + n)
        got, margin = ssc_model.decode(ops, pos, n, K, llr)
        want = port.decode_batch(llr, 1)
        ok = margin > 1e-9          # (a zero / vanishing deciding LLR is what strict mode's second pass is for)
        assert ok.sum() >= B - 2
        assert np.array_equal(got[ok], want[ok])


def test_pruned_tree_schedule_refuses_what_the_kernel_does_not_serve():
    from polar_b200 import _lib
    lib = _lib.dev()
    fr = np.zeros(1 << 13, np.uint8); fr[:100] = 1
    assert lib.polar_b200_ssc_schedule(13, fr.ctypes.data, None, 0) == 0       # block length out of range
    assert lib.polar_b200_ssc_schedule(7, fr.ctypes.data, None, 0) == 0
    assert lib.polar_b200_ssc_schedule(9, np.zeros(512, np.uint8).ctypes.data, None, 0) == 0   # nothing frozen
    assert lib.polar_b200_ssc_schedule(9, np.ones(512, np.uint8).ctypes.data, None, 0) == 0    # everything frozen
