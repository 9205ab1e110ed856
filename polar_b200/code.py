"""Python mirror of the drop-in `PolarCode` class (polar_b200/csrc/PolarCode.h), which itself
mirrors the reference's PolarC/PolarCode.h:19-34: constructor(n, K, epsilon, crc), encode,
decode_scl_llr, get_bler_quick -- plus the batched entry points.

Everything here is a thin ctypes veneer: construction, encoding and the BLER harness run in the
C++ host class, decoding runs in the CUDA library. Nothing falls back to Python or CPU decoding.
"""
import ctypes as C

import numpy as np

from . import _lib


def _ptr(x):
    """raw address of a numpy array or torch tensor"""
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()


def unpack_bits(packed, K):
    """[B][KW] uint32 (numpy) -> [B][K] uint8"""
    packed = np.ascontiguousarray(packed, np.uint32)
    bits = np.unpackbits(packed.view(np.uint8), axis=-1, bitorder="little")
    return np.ascontiguousarray(bits[..., :K])


def pack_bits(bits):
    """[B][K] 0/1 -> [B][ceil(K/32)] uint32, bit j of a row at word j//32, bit j%32"""
    bits = np.ascontiguousarray(bits, np.uint8)
    B, K = bits.shape
    KW = (K + 31) // 32
    pad = np.zeros((B, KW * 32), np.uint8)
    pad[:, :K] = bits
    return np.packbits(pad, axis=-1, bitorder="little").view(np.uint32).reshape(B, KW)


MODES = {"fp32": 0, "strict": 1, "f64": 2, "minsum": 3}


class HostBuffer:
    """Page-locked host memory from the library (polar_b200_host_alloc), exposed as a numpy array `.array`.
    write_combined=True is meant for LLR input buffers: filled once by the CPU, read by the GPU over PCIe."""

    def __init__(self, shape, dtype=np.float32, write_combined=False):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = _lib.dev().polar_b200_host_alloc(self.nbytes, 1 if write_combined else 0)
        if not self.ptr:
            raise _lib.PolarB200Error("polar_b200_host_alloc(%d bytes) failed" % self.nbytes)
        buf = (C.c_ubyte * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def close(self):
        if self.ptr:
            self.array = None
            _lib.dev().polar_b200_host_free(self.ptr)
            self.ptr = None


class PolarCode:
    def __init__(self, n, K, epsilon=0.32, crc=0, reseed=True, device=0, mode=None):
        """n = log2(block length) as in the C++ reference (PolarCode.h:19). reseed=True calls
        srand(1) first so the random parity matrix (PolarCode.cpp:51-56) is the one a fresh
        reference process draws."""
        self._h = None
        lib = _lib.host()
        self.n, self.N, self.K, self.crc = int(n), 1 << int(n), int(K), int(crc)
        self.KW = (self.K + 31) // 32
        self.device = int(device)
        self._h = lib.polar_host_create(self.n, self.K, float(epsilon), self.crc, 1 if reseed else 0, self.device)
        if not self._h:
            raise _lib.PolarB200Error("PolarCode: " + lib.polar_host_last_error().decode())
        if mode is not None:
            self.mode = mode

    # arithmetic mode of every decode (include/polar_b200.h): "fp32" kernels alone, "strict" (default: fp32 kernels +
    # double re-decode of the codewords with a close decision), "f64" (everything in double)
    @property
    def mode(self):
        m = _lib.host().polar_host_get_mode(self._h)
        return [k for k, v in MODES.items() if v == m][0]

    @mode.setter
    def mode(self, m):
        _lib.host().polar_host_set_mode(self._h, MODES[m] if isinstance(m, str) else int(m))

    class _Mode:
        def __init__(self, code, mode):
            self.code, self.new = code, mode

        def __enter__(self):
            self.old = self.code.mode
            if self.new is not None:
                self.code.mode = self.new

        def __exit__(self, *a):
            self.code.mode = self.old

    def close(self):
        if self._h:
            _lib.host().polar_host_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- construction tables (PolarCode.h:43-46) ----
    def construction(self):
        frozen = np.zeros(self.N, np.uint8)
        order = np.zeros(self.N, np.uint16)
        bitrev = np.zeros(self.N, np.uint16)
        crcm = np.zeros((max(self.crc, 1), self.K), np.uint8)
        _lib.host().polar_host_get_construction(self._h, frozen.ctypes.data, order.ctypes.data, crcm.ctypes.data,
                                                bitrev.ctypes.data)
        return dict(frozen=frozen, order=order, crc_matrix=crcm[: self.crc], bitrev=bitrev)

    # ---- PolarCode.h:30 ----
    def encode(self, info):
        info = np.ascontiguousarray(info, np.uint8)
        single = info.ndim == 1
        info2 = info.reshape(-1, self.K)
        out = np.zeros((info2.shape[0], self.N), np.uint8)
        _lib.check_host(_lib.host().polar_host_encode(self._h, info2.ctypes.data, info2.shape[0], out.ctypes.data))
        return out[0] if single else out

    # ---- PolarCode.h:32: one codeword, double LLRs in, K bytes out ----
    def decode_scl_llr(self, llr, list_size):
        llr = np.ascontiguousarray(llr, np.float64)
        if llr.shape != (self.N,):
            raise ValueError("llr must have shape (N,)")
        out = np.zeros(self.K, np.uint8)
        _lib.check_host(_lib.host().polar_host_decode_scl_llr(self._h, llr.ctypes.data, int(list_size), out.ctypes.data))
        return out

    # ---- batched, host memory (numpy or pinned torch CPU tensors) ----
    def decode_batch(self, llr, list_size, packed=False, out=None, mode=None):
        """llr: [B][N] float32 in host memory. Returns [B][K] uint8 bits, or the packed
        [B][KW] uint32 words if packed=True. H2D, decode and D2H all happen inside the call."""
        if isinstance(llr, np.ndarray):
            llr = np.ascontiguousarray(llr, np.float32).reshape(-1, self.N)
        B = int(llr.shape[0])
        if out is None:
            out = np.zeros((B, self.KW), np.uint32)
        with self._Mode(self, mode):
            _lib.check_host(_lib.host().polar_host_decode_batch_packed(self._h, _ptr(llr), B, int(list_size), _ptr(out)))
        if packed or not isinstance(out, np.ndarray):
            return out
        return unpack_bits(out, self.K)

    def decode_batch_double(self, llr, list_size, packed=False, mode=None):
        """[B][N] float64 host LLRs in the current arithmetic mode; in "strict" the fp32 kernels decode the rounded
        LLRs and the flagged codewords are decoded again in double on these doubles (what get_bler_quick does)."""
        llr = np.ascontiguousarray(llr, np.float64).reshape(-1, self.N)
        out = np.zeros((llr.shape[0], self.KW), np.uint32)
        with self._Mode(self, mode):
            _lib.check_host(_lib.host().polar_host_decode_batch_packed_double(self._h, llr.ctypes.data, llr.shape[0],
                                                                              int(list_size), out.ctypes.data))
        return out if packed else unpack_bits(out, self.K)

    def decode_batch_f64(self, llr, list_size, packed=False):
        """Reference-precision mode: [B][N] float64 host LLRs, evaluated in double on the GPU with the
        reference's literal formulas (include/polar_b200.h: polar_b200_decode_scl_llr_f64_host)."""
        llr = np.ascontiguousarray(llr, np.float64).reshape(-1, self.N)
        out = np.zeros((llr.shape[0], self.KW), np.uint32)
        _lib.check_host(_lib.host().polar_host_decode_batch_packed_f64(self._h, llr.ctypes.data, llr.shape[0],
                                                                       int(list_size), out.ctypes.data))
        return out if packed else unpack_bits(out, self.K)

    # ---- PolarCode.h:31: probability-domain decoder (p1 first, like the reference) ----
    def decode_scl_p1(self, p1, p0, list_size):
        p1 = np.ascontiguousarray(p1, np.float64)
        p0 = np.ascontiguousarray(p0, np.float64)
        out = np.zeros(self.K, np.uint8)
        _lib.check_host(_lib.host().polar_host_decode_scl_p1(self._h, p1.ctypes.data, p0.ctypes.data, int(list_size),
                                                             out.ctypes.data))
        return out

    def decode_p1_batch(self, p1, p0, list_size, packed=False):
        """[B][N] float64 likelihoods P(y|1), P(y|0) in host memory -> [B][K] bits (include/polar_b200.h:
        polar_b200_decode_scl_p1_host)."""
        p1 = np.ascontiguousarray(p1, np.float64).reshape(-1, self.N)
        p0 = np.ascontiguousarray(p0, np.float64).reshape(-1, self.N)
        out = np.zeros((p1.shape[0], self.KW), np.uint32)
        _lib.check_host(_lib.host().polar_host_decode_p1_batch_packed(self._h, p1.ctypes.data, p0.ctypes.data, p1.shape[0],
                                                                      int(list_size), out.ctypes.data))
        return out if packed else unpack_bits(out, self.K)

    def set_exact(self, exact=True):
        """decode_scl_llr / get_bler_quick in reference precision (double) from now on"""
        _lib.host().polar_host_set_exact(self._h, 1 if exact else 0)

    # ---- batched, device memory (torch CUDA tensors), asynchronous on the current stream ----
    def decode_device(self, llr, list_size, out=None, stream=None, mode=None, margin=None):
        """margin: optional float32 cuda tensor [B], receives every codeword's smallest decision margin."""
        import torch
        assert llr.is_cuda and llr.dtype == torch.float32 and llr.is_contiguous()
        B = llr.numel() // self.N
        if out is None:
            out = torch.empty((B, self.KW), dtype=torch.int32, device=llr.device)
        if stream is None:
            stream = torch.cuda.current_stream(llr.device).cuda_stream
        with self._Mode(self, mode):
            _lib.check_host(_lib.host().polar_host_decode_device(self._h, llr.data_ptr(), B, int(list_size), out.data_ptr(),
                                                                 C.c_void_p(stream),
                                                                 margin.data_ptr() if margin is not None else None))
        return out

    @property
    def last_flagged(self):
        """codewords the last strict-mode call decoded again in double (synchronises)"""
        return self.info(8)

    def set_strict_tau(self, tau):
        _lib.check(_lib.dev().polar_b200_set_strict_tau(self.ctx(1), float(tau)))

    def count_errors(self, dec, truth, block_err=None, n_err=None, stream=None):
        """device tensors [B][KW] int32; fills block_err (uint8 [B]) / adds to n_err (int64 [1])."""
        import torch
        B = dec.shape[0]
        if stream is None:
            stream = torch.cuda.current_stream(dec.device).cuda_stream
        ctx = self.ctx(1)
        _lib.check(_lib.dev().polar_b200_count_errors(
            ctx, dec.data_ptr(), truth.data_ptr(), B, block_err.data_ptr() if block_err is not None else None,
            n_err.data_ptr() if n_err is not None else None, C.c_void_p(stream)))

    def synthesize(self, B, ebno_db, seed, first_index=0, llr=None, truth=None, stream=None, device=None):
        """Device-side synthetic workload (include/polar_b200.h: polar_b200_synthesize): B codewords with
        global indices first_index.., Eb/N0 of codeword i = ebno_db[i % len(ebno_db)].
        Returns (llr [B][N] float32 cuda, truth [B][KW] int32 cuda = packed info bits)."""
        import torch
        dev = torch.device("cuda", self.device) if device is None else device
        if llr is None:
            llr = torch.empty((B, self.N), dtype=torch.float32, device=dev)
        if truth is None:
            truth = torch.empty((B, self.KW), dtype=torch.int32, device=dev)
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        eb = np.ascontiguousarray(np.atleast_1d(ebno_db), np.float64)
        _lib.check(_lib.dev().polar_b200_synthesize(self.ctx(1), int(seed) & (2**64 - 1), int(first_index), int(B),
                                                    eb.ctypes.data, len(eb), llr.data_ptr(), truth.data_ptr(),
                                                    C.c_void_p(stream)))
        return llr, truth

    def bler_sweep_device(self, ebno_db, list_sizes, count, seed, first_index=0, mode=None):
        """polar_b200_bler_sweep on this object's device: `count` codewords from global index first_index (codeword g at
        point g % len(ebno_db)) synthesised, decoded and compared on the GPU. Returns int64 [lists][ebno][2] = (num_err,
        num_run) of this shard."""
        eb = np.ascontiguousarray(np.atleast_1d(ebno_db), np.float64)
        ls = np.ascontiguousarray(np.atleast_1d(list_sizes), np.int32)
        counts = np.zeros((len(ls), len(eb), 2), np.int64)
        m = MODES[self.mode if mode is None else mode] if not isinstance(mode, int) else mode
        _lib.check(_lib.dev().polar_b200_bler_sweep(self.ctx(1), int(seed) & (2**64 - 1), int(first_index), int(count),
                                                    eb.ctypes.data, len(eb), ls.ctypes.data, len(ls), m, counts.ctypes.data, None))
        return counts

    def bler_sweep(self, ebno_db, list_sizes, total, seed=1, devices=None):
        """PolarCode::bler_sweep (C++): the index range [0, total) sharded over `devices` (default all visible), one host
        thread per device, counters summed with ncclAllReduce. Returns (bler [lists][ebno], counts [lists][ebno][2])."""
        eb = np.ascontiguousarray(np.atleast_1d(ebno_db), np.float64)
        ls = np.ascontiguousarray(np.atleast_1d(list_sizes), np.uint8)
        counts = np.zeros((len(ls), len(eb), 2), np.int64)
        dv = np.ascontiguousarray(devices, np.int32) if devices is not None else None
        _lib.check_host(_lib.host().polar_host_bler_sweep(self._h, eb.ctypes.data, len(eb), ls.ctypes.data, len(ls), int(total),
                                                          int(seed) & (2**64 - 1), dv.ctypes.data if dv is not None else None,
                                                          len(dv) if dv is not None else 0, counts.ctypes.data))
        with np.errstate(divide="ignore", invalid="ignore"):
            return counts[..., 0] / counts[..., 1], counts

    def ctx(self, min_batch=1):
        c = _lib.host().polar_host_ctx(self._h, int(min_batch))
        if not c:
            raise _lib.PolarB200Error("PolarCode.ctx: " + _lib.host().polar_host_last_error().decode())
        return c

    def info(self, key):
        return int(_lib.dev().polar_b200_get_info(self.ctx(1), int(key)))

    @property
    def kernel_launches(self):
        return self.info(0)

    # ---- PolarCode.h:34 ----
    def get_bler_quick(self, ebno_vec, list_sizes, max_err=100, max_runs=1000, verbose=False):
        ebno = np.ascontiguousarray(ebno_vec, np.float64)
        lists = np.ascontiguousarray(list_sizes, np.uint8)
        bler = np.zeros((len(lists), len(ebno)), np.float64)
        _lib.check_host(_lib.host().polar_host_get_bler_quick(self._h, ebno.ctypes.data, len(ebno), lists.ctypes.data,
                                                              len(lists), int(max_err), int(max_runs), 1 if verbose else 0,
                                                              bler.ctypes.data))
        return bler
