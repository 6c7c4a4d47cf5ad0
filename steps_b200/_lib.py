"""ctypes binding of libstepsb200.so (the C ABI declared in include/steps_b200.h).

This is the reference-side stub a Python caller would write; INTEGRATION.md shows the C++ one.
The library is built in-tree by ``__graft_entry__.build()`` / ``steps_b200.build``; there is no
fallback if it is missing and no CPU path if there is no GPU.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstepsb200.so")

ABI_VERSION = 1
TOPO_R3, TOPO_T3, TOPO_S1R2_LOOKUP, TOPO_S1R2_NOLOOKUP = 0, 1, 2, 3


class StepsError(RuntimeError):
    """non-zero status from libstepsb200 (message = steps_b200_last_error())"""


class CParams(C.Structure):
    """struct steps_b200_params (include/steps_b200.h)"""

    _fields_ = [
        ("abi_version", C.c_int32),
        ("topology", C.c_int32),
        ("n", C.c_int32),
        ("cosmology", C.c_int32),
        ("comoving", C.c_int32),
        ("is_periodic", C.c_int32),
        ("s1r2_interp_order", C.c_int32),
        ("table_dim0", C.c_int32),
        ("table_dim1", C.c_int32),
        ("radial_table_size", C.c_int32),
        ("L", C.c_double),
        ("Rsim", C.c_double),
        ("mass_in_unit_sphere", C.c_double),
        ("H0", C.c_double),
        ("Omega_lambda", C.c_double),
        ("ewald_table", C.c_void_p),
        ("radial_table", C.c_void_p),
    ]


class CCosmo(C.Structure):
    """struct steps_b200_cosmo"""

    _fields_ = [("H0", C.c_double), ("Omega_m", C.c_double), ("Omega_r", C.c_double), ("Omega_lambda", C.c_double), ("Omega_k", C.c_double)]


# every symbol include/steps_b200.h declares: name -> (restype, argtypes)
_VP, _I, _D = C.c_void_p, C.c_int, C.c_double
_PP = C.POINTER(CParams)
_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int)
SYMBOLS = {
    "steps_b200_last_error": (C.c_char_p, []),
    "steps_b200_abi_version": (_I, []),
    "steps_b200_device_count": (_I, []),
    "steps_b200_forces_f64": (_I, [_PP, _VP, _VP, _VP, _VP, _I, _I]),
    "steps_b200_forces_f32": (_I, [_PP, _VP, _VP, _VP, _VP, _I, _I]),
    "steps_b200_forces_multi_f64": (_I, [_PP, _VP, _VP, _VP, _VP, _I, _I, _I]),
    "steps_b200_forces_multi_f32": (_I, [_PP, _VP, _VP, _VP, _VP, _I, _I, _I]),
    "steps_b200_release_cached": (None, []),
    "steps_b200_softening_f64": (_I, [_VP, _I, _D, _VP, _PD, _PD]),
    "steps_b200_softening_f32": (_I, [_VP, _I, C.c_float, _VP, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "steps_b200_t3_ewald_defaults": (_I, [_I, _D, _PI, _PD, _PD, _PD]),
    "steps_b200_ewald_space_count": (_I, [_D]),
    "steps_b200_t3_ewald_table_f64": (_I, [_I, _D, _D, _D, _D, _VP, _I]),
    "steps_b200_s1r2_ewald_defaults": (_I, [_I, _D, _D, _PI, _PI, _PD, _PD, _PI, _PI]),
    "steps_b200_s1r2_ewald_table_f64": (_I, [_I, _I, _D, _D, _D, _I, _I, _VP, _I]),
    "steps_b200_radial_force_table_f64": (_I, [_D, _D, _I, _I, _VP, _I]),
    "steps_b200_engine_create": (_I, [C.POINTER(_VP), _PP, _I, _I]),
    "steps_b200_engine_destroy": (None, [_VP]),
    "steps_b200_partition": (None, [_I, _I, _I, _PI, _PI]),
    "steps_b200_engine_set_symmetric": (_I, [_VP, _I]),
    "steps_b200_engine_is_symmetric": (_I, [_VP]),
    "steps_b200_engine_range": (_I, [_VP, _PI, _PI]),
    "steps_b200_sym_rules": (_I, [_I, _I, _I, _I, _PI, _PI, _PI, _I]),
    "steps_b200_sym_schedule_host": (_I, [_I, _I, _I, _I, _I, C.c_longlong, _PI, _PI, _I, _PI, _I, C.POINTER(C.c_ulonglong), C.c_longlong, _PI]),
    "steps_b200_sym_chunk_target": (_I, [_I, _I, C.c_longlong, _I]),
    "steps_b200_engine_debug_set_rank": (_I, [_VP, _I, _I, _I]),
    "steps_b200_engine_debug_fsym": (_I, [_VP, _VP, _VP, _PI]),
    "steps_b200_nccl_unique_id": (_I, [_VP]),
    "steps_b200_engine_comm_init": (_I, [_VP, _VP, _I, _I]),
    "steps_b200_engine_upload": (_I, [_VP, _VP, _VP, _VP, _VP]),
    "steps_b200_engine_upload_x": (_I, [_VP, _VP]),
    "steps_b200_engine_upload_forces": (_I, [_VP, _VP]),
    "steps_b200_group_create": (_I, [C.POINTER(_VP), _PP, _I, _I, _I]),
    "steps_b200_group_destroy": (None, [_VP]),
    "steps_b200_group_size": (_I, [_VP]),
    "steps_b200_group_engine": (_VP, [_VP, _I]),
    "steps_b200_group_cone_select": (_I, [_VP, _D, _I, _PI]),
    "steps_b200_group_cone_rows": (_I, [_VP, _VP, _PI]),
    "steps_b200_group_cone_reset": (_I, [_VP]),
    "steps_b200_redshift_cone_ascii_host": (_I, [C.c_char_p, _VP, _PI, _I, _I, _D, _I, _PD, _I, _PD, _I]),
    "steps_b200_group_cone_write_ascii": (_I, [_VP, C.c_char_p, _D, _I, _PD, _I, _PD, _I]),
    "steps_b200_order_incoherence": (_D, [_VP, _I, _I, _D]),
    "steps_b200_group_upload": (_I, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "steps_b200_group_forces": (_I, [_VP]),
    "steps_b200_group_init_errmax": (_I, [_VP, _D, _D, _PD]),
    "steps_b200_group_kdk_step": (_I, [_VP, _D, _D, _D, _D, _D, _PD]),
    "steps_b200_group_download": (_I, [_VP, _VP, _VP, _VP]),
    "steps_b200_engine_forces": (_I, [_VP, _I, _I]),
    "steps_b200_engine_download_forces": (_I, [_VP, _VP, _I, _I]),
    "steps_b200_engine_download": (_I, [_VP, _VP, _VP, _VP]),
    "steps_b200_engine_init_errmax": (_I, [_VP, _D, _D, _PD]),
    "steps_b200_engine_kdk_step": (_I, [_VP, _D, _D, _D, _D, _D, _PD]),
    "steps_b200_engine_set_glass_making": (_I, [_VP, _I]),
    "steps_b200_engine_glass_stats": (_I, [_VP, _PD]),
    "steps_b200_group_set_glass_making": (_I, [_VP, _I]),
    "steps_b200_snapshot_ascii_host": (_I, [C.c_char_p, _VP, _VP, _VP, _I, _I, _D, _D, _I, _I]),
    "steps_b200_group_snapshot_ascii_async": (_I, [_VP, C.c_char_p, _D, _D, _I]),
    "steps_b200_group_snapshot_wait": (_I, [_VP]),
    "steps_b200_spatial_order": (_I, [_VP, _I, _I, _I, _PI]),
    "steps_b200_permute": (_I, [_VP, _VP, _PI, _I, _I, _I, _I]),
    "steps_b200_group_set_spatial_order": (_I, [_VP, _I]),
    "steps_b200_group_permutation": (_I, [_VP, _PI]),
    "steps_b200_group_glass_stats": (_I, [_VP, _PD]),
    "steps_b200_engine_timings": (_I, [_VP, _PD, _PD]),
    "steps_b200_engine_pair_kernel_ms": (_I, [_VP, _PD]),
    "steps_b200_engine_mark": (_I, [_VP, _I]),
    "steps_b200_engine_elapsed_ms": (_I, [_VP, _I, _I, _PD]),
    "steps_b200_engine_launch_count": (C.c_longlong, [_VP]),
    "steps_b200_engine_sync": (_I, [_VP]),
    "steps_b200_engine_launch_shape": (_I, [_VP, _I, _I, _PI]),
    "steps_b200_friedmann_step": (_D, [C.POINTER(CCosmo), _D, _D]),
    "steps_b200_hubble": (_D, [C.POINTER(CCosmo), _D]),
    "steps_b200_next_timestep": (_D, [_D, _D, _D, _D]),
    "steps_b200_next_timestep_to_output": (_D, [_D, _D, _D, _D, _D, _D, _I]),
    "steps_b200_fma_peak": (_I, [_I, _I, _PD, _PD]),
    "steps_b200_fma_peak_sustained": (_I, [_I, _I, _D, _PD]),
}

_lib = None


def _point_at_bundled_nccl() -> None:
    """A Python process that may `import torch` later must not load an older system libnccl.so.2 first (one object per
    soname): tell the library where the NCCL that PyTorch bundles lives, without importing torch (engine.cu: nccl_load)."""
    if os.environ.get("STEPS_B200_NCCL_LIB"):
        return
    try:
        import importlib.util

        spec = importlib.util.find_spec("nvidia.nccl")
        for root in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(root, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["STEPS_B200_NCCL_LIB"] = cand
                return
    except Exception:  # noqa: BLE001
        pass


def load() -> C.CDLL:
    """dlopen the in-tree library; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise StepsError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  steps_b200 has no fallback path."
        )
    _point_at_bundled_nccl()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.steps_b200_abi_version() != ABI_VERSION:
        raise StepsError("libstepsb200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise StepsError(load().steps_b200_last_error().decode() or f"status {rc}")
