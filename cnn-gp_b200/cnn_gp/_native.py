"""ctypes binding of libcnngp.so (include/cnngp.h).  There is no fallback: if the library is
missing or a call fails, this raises."""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# CNNGP_LIB: another build of the same library (A/B measurements of kernel changes)
LIB_PATH = os.environ.get("CNNGP_LIB") or os.path.join(os.path.dirname(_PKG), "libcnngp.so")

F32, F64 = 0, 1
OP_CONV, OP_RELU, OP_COPY, OP_ADD, OP_SCALE = 1, 2, 3, 4, 5
PATH_AUTO, PATH_GENERIC, PATH_FUSED = 0, 1, 2


class Op(ctypes.Structure):
    """struct cnngp_op of include/cnngp.h."""
    _fields_ = [("opcode", ctypes.c_int32), ("src", ctypes.c_int32), ("dst", ctypes.c_int32),
                ("ke", ctypes.c_int32), ("zero_first", ctypes.c_int32), ("stride", ctypes.c_int32),
                ("pad", ctypes.c_int32), ("dil", ctypes.c_int32),
                ("scale", ctypes.c_double), ("bias", ctypes.c_double)]


_lib = None

_SIGNATURES = {
    # name: (restype, argtypes)
    "cnngp_abi_version": (ctypes.c_int, []),
    "cnngp_last_error": (ctypes.c_char_p, []),
    "cnngp_last_path": (ctypes.c_int, []),
    "cnngp_last_launches": (ctypes.c_int, []),
    "cnngp_plan_create": (ctypes.c_int, [ctypes.POINTER(Op), ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                         ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p)]),
    "cnngp_plan_destroy": (None, [ctypes.c_void_p]),
    "cnngp_plan_aux_elems": (ctypes.c_int64, [ctypes.c_void_p]),
    "cnngp_plan_flops_per_pair": (ctypes.c_double, [ctypes.c_void_p, ctypes.c_int32]),
    "cnngp_plan_has_fused": (ctypes.c_int, [ctypes.c_void_p]),
    "cnngp_plan_describe": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64]),
    "cnngp_plan_dump": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64]),
    "cnngp_variances": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                       ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p]),
    "cnngp_gram": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                  ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                  ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p]),
    "cnngp_gram_band": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                       ctypes.c_void_p]),
    "cnngp_gram_symmetric_to_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                                    ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                                    ctypes.c_void_p, ctypes.c_void_p]),
    "cnngp_conv_maps": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                       ctypes.POINTER(Op), ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]),
    "cnngp_relu_maps": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                       ctypes.c_int64, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                       ctypes.c_int32, ctypes.c_void_p]),
    "cnngp_potrf_upper_f64": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                             ctypes.c_void_p]),
    "cnngp_potrf_panel_f64": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "cnngp_trsm_fwd_panel_f64": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                                ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p]),
    "cnngp_trsm_bwd_diag_f64": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                               ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p]),
    "cnngp_rows_update_f64": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32,
                                             ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                             ctypes.c_int32, ctypes.c_void_p]),
    "cnngp_syrk_upper_f64": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p,
                                            ctypes.c_int64, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                            ctypes.c_void_p]),
    "cnngp_syrk_upper_strided_f64": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p,
                                                    ctypes.c_int64, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                                    ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]),
    "cnngp_potrs_upper_f64": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                             ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p]),
    "cnngp_predict_argmax": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                            ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p]),
    "cnngp_predict_argmax_f64": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                            ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p]),
}

EXPORTS = tuple(_SIGNATURES)


def lib():
    """Load libcnngp.so once.  Raises ImportError with build instructions if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python cnn-gp_b200/build.py` "
                "(or __graft_entry__.build()).  cnn_gp on B200 has no CPU or PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.cnngp_abi_version() != 1:
            raise ImportError("libcnngp.so ABI version mismatch; rebuild")
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().cnngp_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


class Plan:
    """Owns a cnngp_plan*."""

    def __init__(self, ops, n_slots, H, W, dtype_code):
        L = lib()
        arr = (Op * max(1, len(ops)))(*ops)
        handle = ctypes.c_void_p()
        rc = L.cnngp_plan_create(arr, len(ops), n_slots, H, W, dtype_code, ctypes.byref(handle))
        self.handle = None
        if rc != 0:
            msg = L.cnngp_last_error().decode("utf-8", "replace")
            # the reference fails with a RuntimeError (view error) when the final map is not 1x1
            raise RuntimeError(msg)
        self.handle = handle
        self.aux_elems = int(L.cnngp_plan_aux_elems(handle))
        self.fused_kind = int(L.cnngp_plan_has_fused(handle))  # 0 none, 2 straight-line, 3 net
        self.has_fused = self.fused_kind != 0
        self.H, self.W, self.dtype_code = H, W, dtype_code
        self.n_ops, self.n_slots = len(ops), n_slots

    def describe(self):
        """How the plan will run: kernel family and the fused kernels' register-level op list."""
        L = lib()
        n = int(L.cnngp_plan_describe(self.handle, None, 0))
        buf = ctypes.create_string_buffer(n)
        L.cnngp_plan_describe(self.handle, buf, n)
        return buf.value.decode()

    def dump(self):
        """The fused kernels' op lists with every numeric field, one op per line."""
        L = lib()
        n = int(L.cnngp_plan_dump(self.handle, None, 0))
        buf = ctypes.create_string_buffer(n)
        L.cnngp_plan_dump(self.handle, buf, n)
        return buf.value.decode()

    def flops_per_pair(self, C):
        return float(lib().cnngp_plan_flops_per_pair(self.handle, C))

    def __del__(self):
        if getattr(self, "handle", None) is not None and _lib is not None:
            _lib.cnngp_plan_destroy(self.handle)
            self.handle = None
