"""ctypes bindings for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

* ``Port``  -- oracle/_build/libpolar_oracle.so, the CPU restatement (oracle/polar_oracle.cpp).
* ``Ref``   -- oracle/_ref/libpolar_ref.so, the UNMODIFIED reference compiled in place from
  /root/reference/PolarC (oracle/ref_shim.cpp); only exists where it was prebuilt.

Both expose the same small surface so tests can swap one for the other.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PORT_SO = os.path.join(ROOT, "oracle", "_build", "libpolar_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libpolar_ref.so")
REF_O3_SO = os.path.join(ROOT, "oracle", "_ref", "libpolar_ref_o3.so")      # same sources, -O3 -march=native of the build host
REF_MAIN = os.path.join(ROOT, "oracle", "_ref", "polar_ref_main")

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build_port():
    """(Re)build the port oracle if its .so is missing or stale."""
    src = os.path.join(ROOT, "oracle", "polar_oracle.cpp")
    if not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
    return PORT_SO


def have_ref():
    return os.path.exists(REF_SO)


def have_ref_o3():
    """the -O3 -march=native build exists AND this host has every ISA extension it was compiled for"""
    flags_file = os.path.join(os.path.dirname(REF_O3_SO), "o3_flags.txt")
    if not (os.path.exists(REF_O3_SO) and os.path.exists(flags_file)):
        return False
    try:
        need = set(open(flags_file).read().split())
        have = set()
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                have = set(line.split(":", 1)[1].split())
                break
        return need <= have
    except OSError:
        return False


class _Base:
    prefix = None

    def __init__(self, lib, n, K, epsilon=0.32, crc=0, reseed=True, handle=None):
        self.lib = lib
        self.n, self.N, self.K, self.crc = n, 1 << n, K, crc
        p = self.prefix
        f = lambda name: getattr(lib, p + name)
        f("create").restype = C.c_void_p
        f("create").argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]
        f("destroy").argtypes = [C.c_void_p]
        f("get_construction").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        f("encode").argtypes = [C.c_void_p, _u8p, _u8p]
        f("decode_scl_llr").argtypes = [C.c_void_p, _f64p, C.c_int, _u8p]
        f("decode_batch").restype = C.c_double
        self._f = f
        self.h = handle if handle is not None else f("create")(n, K, float(epsilon), crc, 1 if reseed else 0)
        self.last_seconds = None

    def close(self):
        if self.h:
            self._f("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def construction(self):
        frozen = np.zeros(self.N, np.uint8)
        order = np.zeros(self.N, np.uint16)
        bitrev = np.zeros(self.N, np.uint16)
        crcm = np.zeros((max(self.crc, 1), self.K), np.uint8)
        self._f("get_construction")(self.h, frozen.ctypes.data, order.ctypes.data, crcm.ctypes.data, bitrev.ctypes.data)
        return dict(frozen=frozen, order=order, crc_matrix=crcm[: self.crc], bitrev=bitrev)

    def encode(self, info):
        info = np.ascontiguousarray(info, np.uint8)
        if info.ndim == 1:
            out = np.zeros(self.N, np.uint8)
            self._f("encode")(self.h, info, out)
            return out
        out = np.zeros((info.shape[0], self.N), np.uint8)
        for b in range(info.shape[0]):
            self._f("encode")(self.h, info[b], out[b])
        return out

    def decode_one(self, llr, L):
        llr = np.ascontiguousarray(llr, np.float64)
        out = np.zeros(self.K, np.uint8)
        self._f("decode_scl_llr")(self.h, llr, int(L), out)
        return out

    def decode_p1_one(self, p1, p0, L):
        """PolarCode.h:31 decode_scl_p1(p1, p0, list_size): probability-domain decoder, one codeword."""
        fn = self._f("decode_scl_p1")
        fn.argtypes = [C.c_void_p, _f64p, _f64p, C.c_int, _u8p]
        out = np.zeros(self.K, np.uint8)
        fn(self.h, np.ascontiguousarray(p1, np.float64), np.ascontiguousarray(p0, np.float64), int(L), out)
        return out


class Port(_Base):
    prefix = "oracle_"

    def __init__(self, n, K, epsilon=0.32, crc=0, reseed=True, tables=None):
        lib = C.CDLL(build_port())
        handle = None
        if tables is not None:
            lib.oracle_create_from_tables.restype = C.c_void_p
            lib.oracle_create_from_tables.argtypes = [C.c_int, C.c_int, C.c_int, _u8p, _u16p, C.c_void_p]
            crcm = np.ascontiguousarray(tables["crc_matrix"], np.uint8)
            handle = lib.oracle_create_from_tables(
                n, K, crc, np.ascontiguousarray(tables["frozen"], np.uint8),
                np.ascontiguousarray(tables["order"], np.uint16), crcm.ctypes.data if crc else None)
        super().__init__(lib, n, K, epsilon, crc, reseed, handle)
        lib.oracle_decode_batch.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, _u8p, C.c_int, C.c_int, C.c_int]
        lib.oracle_get_bler_quick.argtypes = [C.c_void_p, _f64p, C.c_int, _u8p, C.c_int, C.c_int, C.c_int, _f64p, _f64p]

    def decode_batch(self, llr, L, nthreads=1, precision=0, minsum_only=0):
        llr = np.ascontiguousarray(llr, np.float32).reshape(-1, self.N)
        out = np.zeros((llr.shape[0], self.K), np.uint8)
        self.last_seconds = self.lib.oracle_decode_batch(self.h, llr, llr.shape[0], int(L), out, int(nthreads),
                                                         int(precision), int(minsum_only))
        return out

    def decode_p1_batch(self, p1, p0, L, nthreads=1):
        p1 = np.ascontiguousarray(p1, np.float64).reshape(-1, self.N)
        p0 = np.ascontiguousarray(p0, np.float64).reshape(-1, self.N)
        out = np.zeros((p1.shape[0], self.K), np.uint8)
        fn = self.lib.oracle_decode_p1_batch
        fn.restype = C.c_double
        fn.argtypes = [C.c_void_p, _f64p, _f64p, C.c_int, C.c_int, _u8p, C.c_int]
        self.last_seconds = fn(self.h, p1, p0, p1.shape[0], int(L), out, int(nthreads))
        return out

    def get_bler_quick(self, ebno, lists, max_err=100, max_runs=1000):
        ebno = np.ascontiguousarray(ebno, np.float64)
        lists = np.ascontiguousarray(lists, np.uint8)
        bler = np.zeros((len(lists), len(ebno)), np.float64)
        counts = np.zeros((len(lists), len(ebno), 2), np.float64)
        self.lib.oracle_get_bler_quick(self.h, ebno, len(ebno), lists, len(lists), max_err, max_runs, bler, counts)
        return bler, counts


class Ref(_Base):
    prefix = "ref_"

    def __init__(self, n, K, epsilon=0.32, crc=0, reseed=True, so=None):
        lib = C.CDLL(so or REF_SO)
        super().__init__(lib, n, K, epsilon, crc, reseed)
        lib.ref_decode_batch.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, _u8p, C.c_int]

    def decode_batch(self, llr, L, nthreads=1):
        llr = np.ascontiguousarray(llr, np.float32).reshape(-1, self.N)
        out = np.zeros((llr.shape[0], self.K), np.uint8)
        self.last_seconds = self.lib.ref_decode_batch(self.h, llr, llr.shape[0], int(L), out, int(nthreads))
        return out


def awgn_llrs(code, B, ebno_db, seed):
    """Synthetic BPSK/AWGN LLR batch following PolarCode.cpp:744-753 (float32, [B][N]),
    with fresh random info bits per codeword. Returns (info [B][K] u8, llr [B][N] f32)."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    info = rng.integers(0, 2, size=(B, code.K), dtype=np.uint8)
    coded = code.encode(info)
    a = 10.0 ** (ebno_db / 20.0) * np.sqrt(code.K / code.N)
    r = a * (2.0 * coded.astype(np.float64) - 1.0) + np.sqrt(0.5) * rng.standard_normal((B, code.N))
    llr = (-4.0 * r * a).astype(np.float32)
    return info, llr


def awgn_probs(code, B, ebno_db, seed):
    """Channel likelihoods for the probability-domain decoder, as in the reference's commented-out harness
    lines PolarCode.cpp:749-751: p0 = exp(-(r + a)^2 / N0) / sqrt(pi N0), p1 = exp(-(r - a)^2 / N0) / sqrt(pi N0),
    N0 = 1 (BPSK 0 -> -a, 1 -> +a). Returns (info [B][K] u8, p1 [B][N] f64, p0 [B][N] f64)."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    info = rng.integers(0, 2, size=(B, code.K), dtype=np.uint8)
    coded = code.encode(info)
    a = 10.0 ** (ebno_db / 20.0) * np.sqrt(code.K / code.N)
    r = a * (2.0 * coded.astype(np.float64) - 1.0) + np.sqrt(0.5) * rng.standard_normal((B, code.N))
    p0 = np.exp(-(r + a) ** 2) / np.sqrt(np.pi)
    p1 = np.exp(-(r - a) ** 2) / np.sqrt(np.pi)
    return info, np.ascontiguousarray(p1), np.ascontiguousarray(p0)


def edge_probs(N, seed=3):
    """Edge inputs of the probability-domain decoder: exact ties (p0 == p1 everywhere), certain and impossible
    symbols (exact 0 / 1), everything zero (the sigma == 0 underflow branch, PolarCode.cpp:410-411), denormal
    magnitudes, and ties mixed with noise. Returns (p1 [R][N], p0 [R][N])."""
    rng = np.random.default_rng(seed)
    rows = []
    rows.append((np.full(N, 0.5), np.full(N, 0.5)))
    rows.append((np.zeros(N), np.zeros(N)))
    bits = rng.integers(0, 2, N).astype(np.float64)
    rows.append((bits, 1.0 - bits))
    rows.append((np.full(N, 0.3), np.full(N, 0.7)))
    rows.append((np.full(N, 1e-300), np.full(N, 3e-300)))
    p = rng.random(N)
    mix = np.where(rng.random(N) < 0.5, 0.5, p)
    rows.append((mix, 1.0 - mix))
    p = rng.random(N)
    rows.append((np.where(rng.random(N) < 0.3, 0.0, p), np.where(rng.random(N) < 0.3, 0.0, 1.0 - p)))
    rows.append((np.full(N, 0.25), np.full(N, 0.25)))
    p1 = np.stack([r[0] for r in rows]).astype(np.float64)
    p0 = np.stack([r[1] for r in rows]).astype(np.float64)
    return np.ascontiguousarray(p1), np.ascontiguousarray(p0)
