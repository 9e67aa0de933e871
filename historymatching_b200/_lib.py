"""ctypes binding of libhm_b200.so (the C ABI of include/hm_b200.h).

There is no CPU fallback: if the library is missing, or no sm_100 device is
visible when a context is requested, the error is raised to the caller.
"""

from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhm_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
i64 = C.c_int64


class HmError(RuntimeError):
    pass


class SimDesc(C.Structure):
    """Mirror of ``hm_sim_desc`` (include/hm_b200.h)."""

    _fields_ = [
        ("n_members", C.c_int32), ("Nx", C.c_int32), ("Ny", C.c_int32),
        ("Lx", C.c_double), ("Ly", C.c_double),
        ("vw", C.c_double), ("vo", C.c_double), ("swc", C.c_double), ("sor", C.c_double),
        ("K", C.c_void_p), ("K_member_stride", i64), ("K_comp_stride", i64),
        ("por", C.c_void_p),
        ("n_wells", C.c_int32),
        ("well_cell", C.c_void_p), ("well_cell_member_stride", i64),
        ("well_rate", C.c_void_p), ("well_rate_member_stride", i64), ("well_rate_step_stride", i64),
        ("S0", C.c_void_p), ("S0_member_stride", i64),
        ("dt", C.c_double), ("n_steps", C.c_int32),
        ("n_obs", C.c_int32), ("obs_cell", C.c_void_p),
        ("S_last", C.c_void_p), ("S_hist", C.c_void_p), ("obs", C.c_void_p), ("P_last", C.c_void_p),
        ("status", C.c_void_p), ("substeps", C.c_void_p), ("cg_iters", C.c_void_p),
        ("cg_rtol", C.c_double), ("cg_max_iter", C.c_int32),
        ("chunk_members", C.c_int32), ("precond", C.c_int32), ("mg_switch_iters", C.c_int32), ("sat_block", C.c_int32), ("hist_stride", C.c_int32), ("warm_start", C.c_int32),
        ("tb_cluster_rows", C.c_int32), ("tb_halo", C.c_int32),
        ("K_transform", C.c_int32), ("K_a", C.c_double), ("K_b", C.c_double),
    ]


class SimStats(C.Structure):
    _fields_ = [
        ("cg_iterations", i64), ("sat_substeps", i64), ("kernel_launches", i64),
        ("cg_kernel_launches", i64), ("sat_kernel_launches", i64), ("mg_fp64_fallbacks", i64),
        ("cg_restarts", i64),
        ("sat_resident_ctas", i64),
        ("sat_tb_cluster", i64), ("sat_tb_strips", i64), ("sat_tb_halo", i64), ("sat_cell_updates", i64),
    ]


#: every symbol declared in include/hm_b200.h: name -> (restype, argtypes)
SYMBOLS = {
    "hm_version": (C.c_int, []),
    "hm_last_error": (C.c_char_p, []),
    "hm_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "hm_ctx_destroy": (C.c_int, [C.c_void_p]),
    "hm_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hm_synchronize": (C.c_int, [C.c_void_p]),
    "hm_launch_count": (C.c_int, [C.c_void_p, C.POINTER(i64)]),
    "hm_sim_batch": (C.c_int, [C.c_void_p, C.POINTER(SimDesc)]),
    "hm_sim_batch_host": (C.c_int, [C.c_void_p, C.POINTER(SimDesc)]),
    "hm_sim_get_stats": (C.c_int, [C.c_void_p, C.POINTER(SimStats)]),
    "hm_sim_get_phase_ms": (C.c_int, [C.c_void_p, c_double_p]),
    "hm_dgemm": (C.c_int, [C.c_void_p, C.c_int, C.c_int, i64, i64, i64, C.c_double, C.c_void_p, i64,
                           C.c_void_p, i64, C.c_double, C.c_void_p, i64]),
    "hm_es_update": (C.c_int, [C.c_void_p, i64, i64, i64, C.c_void_p, i64, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p]),
    "hm_es_update_host": (C.c_int, [C.c_void_p, i64, i64, i64, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "hm_les_update": (C.c_int, [C.c_void_p, i64, i64, i64, C.c_void_p, i64, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p]),
    "hm_taper_bump": (C.c_int, [C.c_void_p, i64, i64, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                C.c_void_p]),
    "hm_center": (C.c_int, [C.c_void_p, i64, i64, C.c_void_p, i64, C.c_void_p, i64, C.c_void_p, C.c_int]),
    "hm_ies_step": (C.c_int, [C.c_void_p, i64, i64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_double]),
    "hm_iles_step": (C.c_int, [C.c_void_p, i64, i64, i64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_double]),
    "hm_iles_recompose": (C.c_int, [C.c_void_p, i64, i64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hm_copy2d": (C.c_int, [C.c_void_p, i64, i64, C.c_void_p, i64, C.c_void_p, i64]),
    "hm_swap01": (C.c_int, [C.c_void_p, i64, i64, i64, C.c_void_p, C.c_void_p]),
    "hm_corr": (C.c_int, [C.c_void_p, i64, i64, i64, C.c_void_p, i64, C.c_void_p, i64, C.c_void_p, C.c_int]),
}

_lib = None
_lock = threading.Lock()


def load(build_if_missing: bool = True):
    """Load the shared library (building it with nvcc first if it is absent)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        from . import build as _build

        if _build.needs_build():  # missing, or built from other sources than the tree's (content hash)
            if not build_if_missing:
                raise HmError(f"{LIB_PATH} is missing or stale (run python -m historymatching_b200.build)")
            _build.build()
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int):
    if rc != 0:
        msg = load().hm_last_error().decode(errors="replace")
        raise HmError(f"libhm_b200 error {rc}: {msg}")


class Context:
    """One ``hm_ctx`` per process / GPU; owns the library workspace."""

    _by_device: dict = {}

    def __init__(self, device: int = 0):
        self.lib = load()
        self.device = int(device)
        h = C.c_void_p()
        check(self.lib.hm_ctx_create(self.device, C.byref(h)))
        self.handle = h

    @classmethod
    def get(cls, device: int | None = None) -> "Context":
        if device is None:
            import sys

            torch = sys.modules.get("torch")
            device = torch.cuda.current_device() if torch is not None and torch.cuda.is_available() else 0
        if device not in cls._by_device:
            cls._by_device[device] = Context(device)
        return cls._by_device[device]

    def use_torch_stream(self):
        import torch

        s = torch.cuda.current_stream(self.device).cuda_stream
        check(self.lib.hm_set_stream(self.handle, C.c_void_p(s)))

    def launch_count(self) -> int:
        n = i64()
        check(self.lib.hm_launch_count(self.handle, C.byref(n)))
        return int(n.value)

    def synchronize(self):
        check(self.lib.hm_synchronize(self.handle))

    def close(self):
        if self.handle:
            self.lib.hm_ctx_destroy(self.handle)
            self.handle = None
