"""ctypes loader for the native libraries. Fails loudly: there is no Python or CPU fallback."""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
# POLAR_B200_LIB_DIR points the loader at another build of the same two libraries (A/B runs of kernel variants)
LIB_DIR = os.environ.get("POLAR_B200_LIB_DIR") or os.path.join(PKG, "lib")
DEV_SO = os.path.join(LIB_DIR, "libpolar_b200.so")
HOST_SO = os.path.join(LIB_DIR, "libpolar_host.so")

_dev = None
_host = None


class PolarB200Error(RuntimeError):
    pass


def _need(path):
    if not os.path.exists(path):
        raise PolarB200Error(
            "%s is missing: build it with `python -m polar_b200.build` (needs nvcc). "
            "polar_b200 has no CPU fallback." % path)
    return path


def _preload_nccl():
    """polar_b200_comm_* dlopens "libnccl.so.2" at first use. In a process that also runs torch the copy torch was
    built against (the nvidia-nccl wheel next to it) must be the one in the process: if the system's older libnccl got
    loaded first, torch's own import would later bind to it and fail on missing symbols. So the wheel's library, when
    there is one, is loaded here before anything else can ask for that SONAME."""
    try:
        import glob
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (spec.submodule_search_locations if spec else []):
            for p in sorted(glob.glob(os.path.join(d, "lib", "libnccl.so*"))):
                C.CDLL(p, mode=C.RTLD_GLOBAL)
                return p
    except Exception:
        pass
    return None


def dev():
    """libpolar_b200.so with argtypes set for every symbol of include/polar_b200.h."""
    global _dev
    if _dev is None:
        _preload_nccl()
        lib = C.CDLL(_need(DEV_SO), mode=C.RTLD_GLOBAL)
        vp, ip = C.c_void_p, C.c_int
        lib.polar_b200_abi_version.restype = ip
        lib.polar_b200_strerror.restype = C.c_char_p
        lib.polar_b200_strerror.argtypes = [ip]
        lib.polar_b200_info_words.argtypes = [ip]
        lib.polar_b200_create.argtypes = [C.POINTER(vp), ip, ip, ip, ip, vp, vp, vp, ip, ip]
        lib.polar_b200_destroy.argtypes = [vp]
        lib.polar_b200_decode_scl_llr.argtypes = [vp, vp, ip, ip, vp, vp]
        lib.polar_b200_decode_scl_llr_host.argtypes = [vp, vp, ip, ip, vp, vp]
        lib.polar_b200_decode_scl_llr_ex.argtypes = [vp, vp, ip, ip, vp, ip, vp, vp]
        lib.polar_b200_decode_scl_llr_host_ex.argtypes = [vp, vp, ip, ip, vp, ip, vp]
        lib.polar_b200_decode_scl_llr_f64_strict_host.argtypes = [vp, vp, ip, ip, vp, vp]
        lib.polar_b200_set_strict_tau.argtypes = [vp, C.c_float]
        lib.polar_b200_reserve.argtypes = [vp, ip]
        lib.polar_b200_count_errors.argtypes = [vp, vp, vp, ip, vp, vp, vp]
        lib.polar_b200_decode_scl_llr_f64.argtypes = [vp, vp, ip, ip, vp, vp]
        lib.polar_b200_decode_scl_llr_f64_host.argtypes = [vp, vp, ip, ip, vp, vp]
        lib.polar_b200_decode_scl_p1.argtypes = [vp, vp, vp, ip, ip, vp, vp]
        lib.polar_b200_decode_scl_p1_host.argtypes = [vp, vp, vp, ip, ip, vp, vp]
        lib.polar_b200_synthesize.argtypes = [vp, C.c_ulonglong, C.c_longlong, ip, vp, ip, vp, vp, vp]
        lib.polar_b200_bler_sweep.argtypes = [vp, C.c_ulonglong, C.c_longlong, C.c_longlong, vp, ip, vp, ip, ip, vp, vp]
        lib.polar_b200_device_count.restype = ip
        lib.polar_b200_fast_variant_count.restype = ip
        lib.polar_b200_fast_variant_desc.argtypes = [ip, C.POINTER(ip), C.POINTER(ip), C.POINTER(ip)]
        lib.polar_b200_ssc_schedule.argtypes = [ip, vp, vp, ip]
        lib.polar_b200_ssc_positions.argtypes = [ip, vp, ip, vp]
        lib.polar_b200_host_alloc.restype = vp
        lib.polar_b200_host_alloc.argtypes = [C.c_size_t, ip]
        lib.polar_b200_host_free.argtypes = [vp]
        lib.polar_b200_comm_unique_id.argtypes = [vp]
        lib.polar_b200_comm_init_rank.argtypes = [C.POINTER(vp), ip, ip, ip, vp]
        lib.polar_b200_comm_init_all.argtypes = [vp, ip, vp]
        lib.polar_b200_comm_allreduce_i64.argtypes = [vp, vp, ip]
        lib.polar_b200_comm_allreduce_i64_group.argtypes = [vp, ip, vp, ip]
        lib.polar_b200_comm_destroy.argtypes = [vp]
        lib.polar_b200_get_info.restype = C.c_longlong
        lib.polar_b200_get_info.argtypes = [vp, ip]
        _dev = lib
    return _dev


def host():
    """libpolar_host.so (the C++ PolarCode class behind C wrappers)."""
    global _host
    if _host is None:
        dev()
        lib = C.CDLL(_need(HOST_SO))
        vp, ip = C.c_void_p, C.c_int
        lib.polar_host_last_error.restype = C.c_char_p
        lib.polar_host_create.restype = vp
        lib.polar_host_create.argtypes = [ip, ip, C.c_double, ip, ip, ip]
        lib.polar_host_destroy.argtypes = [vp]
        lib.polar_host_get_construction.argtypes = [vp, vp, vp, vp, vp]
        lib.polar_host_encode.argtypes = [vp, vp, ip, vp]
        lib.polar_host_decode_scl_llr.argtypes = [vp, vp, ip, vp]
        lib.polar_host_decode_batch_packed.argtypes = [vp, vp, ip, ip, vp]
        lib.polar_host_decode_device.argtypes = [vp, vp, ip, ip, vp, vp, vp]
        lib.polar_host_set_mode.argtypes = [vp, ip]
        lib.polar_host_get_mode.argtypes = [vp]
        lib.polar_host_decode_batch_packed_double.argtypes = [vp, vp, ip, ip, vp]
        lib.polar_host_decode_batch_packed_f64.argtypes = [vp, vp, ip, ip, vp]
        lib.polar_host_decode_p1_batch_packed.argtypes = [vp, vp, vp, ip, ip, vp]
        lib.polar_host_decode_scl_p1.argtypes = [vp, vp, vp, ip, vp]
        lib.polar_host_set_exact.argtypes = [vp, ip]
        lib.polar_host_ctx.restype = vp
        lib.polar_host_ctx.argtypes = [vp, ip]
        lib.polar_host_get_bler_quick.argtypes = [vp, vp, ip, vp, ip, ip, ip, ip, vp]
        lib.polar_host_bler_sweep.argtypes = [vp, vp, ip, vp, ip, C.c_longlong, C.c_ulonglong, vp, ip, vp]
        _host = lib
    return _host


def check(rc, what="polar_b200"):
    if rc != 0:
        raise PolarB200Error("%s: %s" % (what, dev().polar_b200_strerror(rc).decode()))


def check_host(rc, what="polar_host"):
    if rc != 0:
        raise PolarB200Error("%s: %s" % (what, host().polar_host_last_error().decode()))
