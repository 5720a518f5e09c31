"""ctypes binding of the C ABI in include/d3dp_b200.h (libd3dp_b200.so, built in-tree by d3dp_b200/csrc/build.sh).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# D3DP_LIB: run against another BUILD of the same sources (profiles/ab_lib.py variants, sanitizer builds); the default
# is the in-tree library.  Either way it is this C-ABI library or an exception — there is nothing else to fall back to.
LIB_PATH = os.environ.get("D3DP_LIB") or os.path.join(_HERE, "csrc", "libd3dp_b200.so")

# every symbol include/d3dp_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "d3dp_create", "d3dp_destroy", "d3dp_last_error", "d3dp_set_weight", "d3dp_weights_missing",
    "d3dp_set_schedule", "d3dp_schedule_host", "d3dp_get_alphas_cumprod", "d3dp_time_list", "d3dp_workspace_bytes", "d3dp_denoise",
    "d3dp_ddim_sample", "d3dp_q_sample", "d3dp_jpma", "d3dp_jpma_gt", "d3dp_pmpjpe",
    "d3dp_philox_normal", "d3dp_test_gemm", "d3dp_test_attn",
    "d3dp_version",
]


class D3dpConfig(C.Structure):
    _fields_ = [
        ("frames", C.c_int32), ("joints", C.c_int32), ("channels", C.c_int32), ("depth", C.c_int32),
        ("heads", C.c_int32), ("mlp_hidden", C.c_int32), ("num_timesteps", C.c_int32), ("scale", C.c_float),
        ("flip_perm", C.c_int32 * 17), ("output_scale", C.c_float),
    ]


class D3dpError(RuntimeError):
    pass


_lib = None


def load():
    """Load libd3dp_b200.so (once) and declare the prototypes. Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise D3dpError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or d3dp_b200/csrc/build.sh). d3dp_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, f32p, i32p, i64p, f64p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
    lib.d3dp_version.restype = C.c_char_p
    lib.d3dp_create.argtypes = [C.POINTER(D3dpConfig), C.POINTER(vp)]
    lib.d3dp_destroy.argtypes = [vp]
    lib.d3dp_destroy.restype = None
    lib.d3dp_last_error.argtypes = [vp]
    lib.d3dp_last_error.restype = C.c_char_p
    lib.d3dp_set_weight.argtypes = [vp, C.c_char_p, f32p, C.c_int64, vp]
    lib.d3dp_weights_missing.argtypes = [vp]
    lib.d3dp_set_schedule.argtypes = [vp, f64p, f64p, f64p, f64p, f64p, C.c_int32, vp]
    lib.d3dp_schedule_host.argtypes = [C.c_int32, f64p]
    lib.d3dp_get_alphas_cumprod.argtypes = [vp, f64p, C.c_int32]
    lib.d3dp_time_list.argtypes = [C.c_int32, C.c_int32, i32p]
    lib.d3dp_workspace_bytes.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]
    lib.d3dp_denoise.argtypes = [vp, f32p, f32p, i64p, f32p, f32p, C.c_int32, C.c_int32, vp, C.c_size_t, vp]
    lib.d3dp_ddim_sample.argtypes = [vp, f32p, f32p, f32p, f32p, C.c_uint64, C.c_int32, C.c_int32, i32p, f32p,
                                     C.c_int32, C.c_int32, C.c_int32, vp, C.c_size_t, vp]
    lib.d3dp_q_sample.argtypes = [vp, f32p, f32p, i64p, f32p, C.c_int32, C.c_int64, C.c_int32, vp]
    lib.d3dp_jpma.argtypes = [vp, f32p, f32p, f32p, f32p, f32p, i32p, f32p, f32p, C.c_int32, C.c_int32, C.c_int32,
                              C.c_int32, C.c_int32, C.c_int32, vp]
    lib.d3dp_jpma_gt.argtypes = [vp, f32p, f32p, f32p, f32p, f32p, f32p, i32p, f32p, f32p, f32p, f32p, C.c_int32,
                                 C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]
    lib.d3dp_pmpjpe.argtypes = [vp, f32p, f32p, f32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]
    lib.d3dp_philox_normal.argtypes = [vp, f32p, C.c_int32, C.c_int32, C.c_int64, C.c_uint64, C.c_int32, C.c_int32,
                                       C.c_uint32, vp]
    lib.d3dp_test_gemm.argtypes = [vp, C.c_int32, vp, vp, f32p, vp, f32p, f32p, f32p, C.c_float, f32p, f32p,
                                   C.c_float, f32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]
    lib.d3dp_test_attn.argtypes = [vp, C.c_int32, vp, vp, C.c_int32, vp]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("d3dp_version", "d3dp_last_error", "d3dp_destroy"):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(handle, rc, what):
    if rc != 0:
        msg = load().d3dp_last_error(handle).decode() if handle else ""
        raise D3dpError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())
