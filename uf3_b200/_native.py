"""ctypes binding of libuf3b.so (C ABI in `include/uf3b.h`).

The library is the product: if it is missing or fails to load, importing a compute
entry point raises — there is no CPU fallback anywhere in `uf3_b200`.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UF3B_LIB") or os.path.join(_HERE, "lib", "libuf3b.so")     # UF3B_LIB: another build (bisection)
CSRC_DIR = os.path.join(_HERE, "csrc")

OK = 0
ERR_INVALID, ERR_CUDA, ERR_ELEMENT, ERR_CAPACITY, ERR_STATE = -1, -2, -3, -4, -5
RETRY = 1


class UF3BError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libuf3b error {code}: {message}")
        self.code = code


class ElementError(UF3BError, ValueError):
    """A configuration holds an element that is not part of the basis."""


_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)


class BasisDesc(C.Structure):
    """`uf3b_basis_desc` (include/uf3b.h)."""
    _fields_ = [("n_elements", C.c_int32), ("atomic_numbers", _i32p),
                ("n_feats", C.c_int32),
                ("leading_trim_2b", C.c_int32), ("trailing_trim_2b", C.c_int32),
                ("leading_trim_3b", C.c_int32), ("trailing_trim_3b", C.c_int32),
                ("pair_n_knots", _i32p), ("pair_knots", _f64p),
                ("pair_r_min", _f64p), ("pair_r_max", _f64p), ("pair_col", _i32p),
                ("n_trios", C.c_int32), ("trio_n_knots", _i32p), ("trio_knots", _f64p),
                ("trio_col", _i32p), ("trio_n_cols", _i32p),
                ("bin_col", _i32p), ("bin_weight", _f64p), ("trio_symmetry", _i32p)]


# name -> (restype, argtypes); every symbol declared in include/uf3b.h
SIGNATURES = {
    "uf3b_last_error": (C.c_char_p, []),
    "uf3b_abi_version": (C.c_int, []),
    "uf3b_set_device": (C.c_int, [C.c_int]),
    "uf3b_device_count": (C.c_int, [_i32p]),
    "uf3b_set_blocking_sync": (C.c_int, [C.c_int]),
    "uf3b_basis_create": (C.c_int, [C.POINTER(BasisDesc), C.POINTER(C.c_void_p)]),
    "uf3b_basis_set_coefficients": (C.c_int, [C.c_void_p, _f64p, C.c_int32]),
    "uf3b_basis_set_frames_in_flight": (C.c_int, [C.c_void_p, C.c_int32]),
    "uf3b_basis_set_deferred_lists": (C.c_int, [C.c_void_p, C.c_int]),
    "uf3b_basis_destroy": (None, [C.c_void_p]),
    "uf3b_neighbors_build": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32,
                                       C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]),
    "uf3b_neighbors_build_range": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32,
                                             C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                             C.POINTER(C.c_void_p), C.c_void_p]),
    "uf3b_neighbors_count": (C.c_int, [C.c_void_p, C.c_int, _i64p]),
    "uf3b_neighbors_export": (C.c_int, [C.c_void_p, C.c_int, _i64p, _i64p]),
    "uf3b_nlist_destroy": (None, [C.c_void_p]),
    "uf3b_featurize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                 C.c_void_p]),
    "uf3b_energy_forces": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "uf3b_gram_create": (C.c_int, [C.c_int32, C.POINTER(C.c_void_p)]),
    "uf3b_gram_accumulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                       C.c_int, C.c_void_p]),
    "uf3b_gram_export": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "uf3b_gram_destroy": (None, [C.c_void_p]),
    "uf3b_pipeline_create": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]),
    "uf3b_pipeline_submit": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, _i64p]),
    "uf3b_pipeline_wait": (C.c_int, [C.c_void_p, C.c_int64]),
    "uf3b_pipeline_submit_fit": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, _i64p]),
    "uf3b_pipeline_export_gram": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "uf3b_pipeline_destroy": (None, [C.c_void_p]),
    "uf3b_pair_histogram": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "uf3b_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "uf3b_host_eval_basis": (C.c_int, [_f64p, C.c_int32, C.c_double, _f64p, _f64p]),
    "uf3b_launch_count": (C.c_int64, []),
    "uf3b_set_timing": (C.c_int, [C.c_int]),
    "uf3b_last_kernel_ms": (C.c_double, []),
    "uf3b_probe_fp64_tflops": (C.c_int, [_f64p]),
}

_lib = None


def build(verbose=False):
    """Compile the CUDA sources for sm_100a into uf3_b200/lib/libuf3b.so (in-tree)."""
    cmd = ["make", "-C", CSRC_DIR, "-j", str(os.cpu_count() or 4)]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.run(cmd, check=True)
    return LIB_PATH


def lib():
    """The loaded library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C uf3_b200/csrc`. uf3_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            if os.environ.get("UF3B_LIB") and not hasattr(handle, name):
                continue
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        if handle.uf3b_abi_version() != 1:
            raise ImportError("libuf3b.so ABI version mismatch")
        _lib = handle
    return _lib


def check(code):
    if code == OK:
        return
    message = lib().uf3b_last_error().decode("utf-8", "replace")
    if code == ERR_ELEMENT:
        raise ElementError(code, message)
    raise UF3BError(code, message)
