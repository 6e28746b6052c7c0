"""Loads th_rl_b200/libthrl.so (the C ABI of include/thrl.h) with ctypes.

There is NO CPU fallback: if the library is missing or was not built, importing the compute path fails loudly.
"""
import ctypes as C
import os

from . import abi

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libthrl.so")
_lib = None


class ThrlError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("thrl error %d: %s" % (code, msg))
        self.code = code


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            "%s not found: build it with `python -m th_rl_b200.build` (needs nvcc; sm_100a only). "
            "th_rl_b200 has no CPU fallback." % SO_PATH)
    L = C.CDLL(SO_PATH)
    L.thrl_abi_version.restype = C.c_int
    if L.thrl_abi_version() != abi.THRL_ABI_VERSION:
        raise ImportError("libthrl.so ABI %d != python mirror %d: rebuild" % (L.thrl_abi_version(), abi.THRL_ABI_VERSION))
    L.thrl_last_error.restype = C.c_char_p
    L.thrl_game_layout.argtypes = [C.POINTER(abi.ThrlGame)]
    L.thrl_game_layout.restype = C.c_int
    L.thrl_ring_bytes.argtypes = [C.POINTER(abi.ThrlGame)]
    L.thrl_ring_bytes.restype = C.c_int64
    L.thrl_qtable_scan.argtypes = [C.POINTER(abi.ThrlScanArgs), C.c_void_p]
    L.thrl_qtable_scan.restype = C.c_int
    L.thrl_qtable_scan_host.argtypes = [C.POINTER(abi.ThrlScanArgs), C.c_int]
    L.thrl_qtable_scan_host.restype = C.c_int
    L.thrl_qtable_init.argtypes = [C.POINTER(abi.ThrlGame), C.c_int64, C.c_int64, C.c_uint64, C.c_int32, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.thrl_qtable_init.restype = C.c_int
    L.thrl_game_init.argtypes = L.thrl_qtable_init.argtypes[:-1] + [C.c_void_p, C.c_void_p]
    L.thrl_game_init.restype = C.c_int
    L.thrl_greedy_eval.argtypes = [C.POINTER(abi.ThrlGame), C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]
    L.thrl_greedy_eval.restype = C.c_int
    L.thrl_greedy_eval_mlp.argtypes = [C.POINTER(abi.ThrlGame), C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.thrl_greedy_eval_mlp.restype = C.c_int
    L.thrl_greedy_eval_noise.argtypes = [C.POINTER(abi.ThrlGame), C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.thrl_greedy_eval_noise.restype = C.c_int
    L.thrl_release_device_memory.restype = C.c_int
    L.thrl_curve_hist.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_double,
                                  C.c_double, C.c_int32, C.c_void_p, C.c_void_p]
    L.thrl_curve_hist.restype = C.c_int
    L.thrl_launch_count.restype = C.c_int64
    L.thrl_last_kernel.restype = C.c_char_p
    L.thrl_last_wave_runs.restype = C.c_int64
    _lib = L
    return L


def check(rc):
    if rc != 0:
        msg = lib().thrl_last_error().decode("utf-8", "replace")
        if rc == abi.THRL_ERR_BAD_CONFIG:
            # the reference raises AssertionError / IndexError for these (trainer.py:21-23, agents.py:88)
            raise ValueError("bad config: " + msg)
        raise ThrlError(rc, msg)


def game_layout(config):
    """config dict (reference JSON schema) -> validated ThrlGame with offsets filled (host only, no device needed)."""
    g = abi.game_from_config(config)
    check(lib().thrl_game_layout(C.byref(g)))
    return g


def launch_count():
    return int(lib().thrl_launch_count())


def last_kernel():
    """Name of the scan kernel this thread's latest scan call ran (include/thrl.h thrl_last_kernel)."""
    return lib().thrl_last_kernel().decode()


def last_wave_runs():
    """Runs the latest scan launch of this thread kept resident at once (include/thrl.h thrl_last_wave_runs)."""
    return int(lib().thrl_last_wave_runs())
