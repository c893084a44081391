"""CPU oracle for the cnn-gp Gram recursion -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product path (cnn-gp_b200/) never does and has no CPU route.

It restates, step for step, what the reference computes (paths relative to /root/reference):

  gram()                 cnn_gp/kernels.py:18-57     NNGPKernel.forward
  _propagate()           cnn_gp/kernels.py:184-187   Sequential.propagate
                         cnn_gp/kernels.py:252-254   Sum.propagate (0 + kp0 + kp1 ...)
                         cnn_gp/kernels.py:221-225   Mixture.propagate (softmax weights)
                         cnn_gp/kernel_patch.py:31-63 element-wise + and *
  arithmetic (C)         oracle/cnngp_oracle.c       conv / relu / init, see that header
  product_tiles(), worker_slice()   cnn_gp/data.py:11-60  tile enumeration and per-worker split
  save_k_blocks()        cnn_gp/kernel_save_tools.py:26-58  block layout written by save_K
  solve_system(), predict()         exp_mnist_resnet/classify_gp.py:17-42

Modules are recognised by duck typing (class name + attributes), so the same walker
evaluates the reference's own module tree (when /root/reference is importable) and this
repository's drop-in ``cnn_gp`` modules.

Pinning: tests/test_oracle.py checks this oracle against tests/golden/*.npz, which
tests/golden/make_golden.py produced by running the unmodified reference.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libcnngp_oracle.so")
_lib = None


def build(force=False):
    """Compile the C part with the recipe in oracle/Makefile."""
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
            for f in ("cnngp_oracle.c", "cnngp_oracle_impl.h")):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def build_ref(src="/root/reference"):
    """oracle/_ref/: the unmodified reference, copied by oracle/make_ref.sh when `src` exists (the
    authoring container); on the GPU box the copy that travelled with the snapshot is used as is.
    Only bench.py's CPU legs (through oracle/ref_cpu.py) execute it.  -> path or None"""
    ref = os.path.join(_HERE, "_ref")
    if os.path.isdir(os.path.join(src, "cnn_gp")):
        subprocess.run(["bash", os.path.join(_HERE, "make_ref.sh"), src], check=True, stdout=subprocess.DEVNULL)
    return ref if os.path.isdir(os.path.join(ref, "cnn_gp")) else None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        i64, ci, vp = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p
        for sfx, real in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
            getattr(L, f"oracle_init_{sfx}").argtypes = [vp, vp, i64, i64, i64, i64, ci, vp, vp, vp]
            getattr(L, f"oracle_conv_{sfx}").argtypes = [vp, i64, i64, i64, ci, ci, ci, ci, ci,
                                                         real, real, vp, i64, i64]
            getattr(L, f"oracle_relu_{sfx}").argtypes = [vp, vp, vp, i64, i64, i64, ci, ci]
            for n in ("init", "conv", "relu"):
                getattr(L, f"oracle_{n}_{sfx}").restype = None
        L.oracle_num_threads.restype = ci
        L.oracle_set_num_threads.argtypes = [ci]
        _lib = L
    return _lib


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))


def num_threads():
    return int(lib().oracle_num_threads())


def _sfx(dtype):
    return {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}[np.dtype(dtype)]


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Patch:
    """(same, diag, xy, xx, yy) of cnn_gp/kernel_patch.py:4-29, as flat [M, H, W] arrays."""

    def __init__(self, same, diag, xy, xx, yy, Nx, Ny):
        self.same, self.diag, self.xy, self.xx, self.yy, self.Nx, self.Ny = same, diag, xy, xx, yy, Nx, Ny

    def _elementwise(self, other, op):  # kernel_patch.py:43-63
        if isinstance(other, Patch):
            assert bool(self.same) == bool(other.same) and bool(self.diag) == bool(other.diag)
            return Patch(self.same, self.diag, op(self.xy, other.xy), op(self.xx, other.xx),
                         op(self.yy, other.yy), self.Nx, self.Ny)
        return Patch(self.same, self.diag, op(self.xy, other), op(self.xx, other), op(self.yy, other),
                     self.Nx, self.Ny)

    def __add__(self, o):
        return self._elementwise(o, lambda a, b: a + b)

    __radd__ = __add__

    def __mul__(self, o):
        return self._elementwise(o, lambda a, b: a * b)


def conv_out_size(n, ke, stride, pad, dil):
    return (n + 2 * pad - dil * (ke - 1) - 1) // stride + 1


def _conv_maps(a, mod, dtype):
    """One F.conv2d(...) + var_bias of kernels.py:94-97 on a stack of maps [M, H, W]."""
    M, Hi, Wi = a.shape
    zero_first = bool(getattr(mod, "kernel_has_row_of_zeros", False))
    ke = int(mod.kernel_size) + (1 if zero_first else 0)
    stride, pad, dil = int(mod.stride), int(mod.padding), int(mod.dilation)
    Ho, Wo = conv_out_size(Hi, ke, stride, pad, dil), conv_out_size(Wi, ke, stride, pad, dil)
    if Ho < 1 or Wo < 1:
        raise RuntimeError("conv output would be empty")
    # the buffer is float32(var_weight / k^2), widened for .double() models (kernels.py:87-88)
    tap = dtype.type(np.float32(float(mod.var_weight) / int(mod.kernel_size) ** 2))
    bias = dtype.type(float(mod.var_bias))
    out = np.empty((M, Ho, Wo), dtype=dtype)
    getattr(lib(), f"oracle_conv_{_sfx(dtype)}")(
        _ptr(np.ascontiguousarray(a)), M, Hi, Wi, ke, int(zero_first), stride, pad, dil,
        tap.item(), bias.item(), _ptr(out), Ho, Wo)
    return out


def _propagate(mod, kp):
    kind = type(mod).__name__
    dtype = kp.xy.dtype
    if kind == "Conv2d":  # kernels.py:92-98
        return Patch(kp.same, kp.diag, _conv_maps(kp.xy, mod, dtype), _conv_maps(kp.xx, mod, dtype),
                     _conv_maps(kp.yy, mod, dtype), kp.Nx, kp.Ny)
    if kind == "ReLU":  # kernels.py:134-165
        xy, xx, yy = (np.array(v, dtype=dtype, order="C", copy=True) for v in (kp.xy, kp.xx, kp.yy))
        if kp.same and not kp.diag and kp.Nx != kp.Ny:
            raise RuntimeError("same=True needs N1 == N2 (eye broadcast, kernels.py:161)")
        P = xy.shape[1] * xy.shape[2]
        getattr(lib(), f"oracle_relu_{_sfx(dtype)}")(_ptr(xy), _ptr(xx), _ptr(yy), kp.Nx, kp.Ny, P,
                                                      int(bool(kp.same)), int(bool(kp.diag)))
        return Patch(kp.same, kp.diag, xy, xx, yy, kp.Nx, kp.Ny)
    if kind == "Sequential":  # kernels.py:184-187
        for m in mod.mods:
            kp = _propagate(m, kp)
        return kp
    if kind == "Sum":  # kernels.py:252-254 : sum() starts from int 0
        total = 0
        for m in mod.mods:
            total = total + _propagate(m, kp)
        return total
    if kind == "Mixture":  # kernels.py:221-225
        logit = np.asarray(mod.logit.detach().cpu().numpy(), dtype=dtype)
        e = np.exp(logit - logit.max())
        prop = (e / e.sum()).astype(dtype)
        total = _propagate(mod.mods[0], kp) * prop[0]
        for i in range(1, len(mod.mods)):
            total = total + _propagate(mod.mods[i], kp) * prop[i]
        return total
    raise TypeError(f"oracle: unknown module kind {kind}")


def gram(model, X, Z=None, same=None, diag=False):
    """model(X, Z, same=, diag=) of kernels.py:18-57 on numpy arrays [N, C, H, W]."""
    X = np.ascontiguousarray(X)
    if Z is None:
        assert same is None
        Z, same = X, True
    Z = np.ascontiguousarray(Z)
    assert not diag or len(X) == len(Z), "diagonal kernels must operate with data of equal length"
    assert X.ndim == 4 and Z.ndim == 4 and X.shape[1:] == Z.shape[1:]
    assert X.dtype == Z.dtype
    dtype = X.dtype
    N1, N2 = X.shape[0], Z.shape[0]
    C, H, W = X.shape[1:]
    xy = np.empty(((N1 if diag else N1 * N2), H, W), dtype=dtype)
    xx = np.empty((N1, H, W), dtype=dtype)
    yy = np.empty((N2, H, W), dtype=dtype)
    getattr(lib(), f"oracle_init_{_sfx(dtype)}")(_ptr(X), _ptr(Z), N1, N2, C, H * W, int(bool(diag)),
                                                  _ptr(xy), _ptr(xx), _ptr(yy))
    kp = _propagate(model, Patch(same, diag, xy, xx, yy, N1, N2))
    if kp.xy.shape[1:] != (1, 1):
        raise RuntimeError(f"final map is {kp.xy.shape[1:]}, not 1x1 (view error at kernels.py:54-57)")
    return kp.xy.reshape(N1) if diag else kp.xy.reshape(N1, N2)


# ---------------------------------------------------------------------------------------
# tile enumeration, cnn_gp/data.py:11-60
def round_up_div(a, b):
    return (a + b - 1) // b


def product_tiles(n_batches_x, n_batches_x2, same):
    """data.py:22-29: row-major, (True,i,i) first then (False,i,j>i) when same."""
    out = []
    for i in range(n_batches_x):
        if same:
            out.append((True, i, i))
        for j in range(i + 1 if same else 0, n_batches_x2):
            out.append((False, i, j))
    return out


def worker_slice(n_batches, worker_rank, n_workers):
    """data.py:11-19: contiguous split, the first n_batches % n_workers workers get one more."""
    per = [n_batches // n_workers + (1 if w < n_batches % n_workers else 0) for w in range(n_workers)]
    return sum(per[:worker_rank]), per[worker_rank]


def worker_tiles(N, N2, batch_size, worker_rank=0, n_workers=1):
    """Tiles ProductIterator(batch_size, X, X2, rank, n) serves (data.py:42-60)."""
    nbx = round_up_div(N, batch_size)
    if N2 is None:
        same, nb2 = True, nbx
        total = max(1, nbx * (nbx + 1) // 2)
    else:
        same, nb2 = False, round_up_div(N2, batch_size)
        total = nbx * nb2
    start, count = worker_slice(total, worker_rank, n_workers)
    return product_tiles(nbx, nb2, same)[start:start + count]


def save_k_blocks(model, X, X2, diag, batch_size, worker_rank=0, n_workers=1):
    """The array save_K leaves in the file (kernel_save_tools.py:7-58): shape (1,N,N2) or (1,N),
    float32, NaN where this worker wrote nothing."""
    N = len(X)
    N2 = N if X2 is None else len(X2)
    out = np.full((1, N) if diag else (1, N, N2), np.nan, dtype=np.float32)
    Xb = X2 if X2 is not None else X
    if diag:  # DiagIterator, data.py:99-126
        for i in range(0, min(N, N2), batch_size):
            x = X[i:i + batch_size]
            x2 = Xb[i:i + batch_size]
            n = min(len(x), len(x2))
            out[0, i:i + n] = gram(model, x[:n], x2[:n], same=(X2 is None), diag=True)
        return out
    for same, bi, bj in worker_tiles(N, None if X2 is None else N2, batch_size, worker_rank, n_workers):
        i, j = bi * batch_size, bj * batch_size
        x, x2 = X[i:i + batch_size], Xb[j:j + batch_size]
        out[0, i:i + len(x), j:j + len(x2)] = gram(model, x, x2, same=same, diag=False)
    return out


# ---------------------------------------------------------------------------------------
# exp_mnist_resnet/classify_gp.py:17-42
def solve_system(Kxx, Y):
    """scipy.linalg.solve(assume_a='pos', lower=False): LAPACK posv reading the upper triangle."""
    import scipy.linalg
    assert Kxx.dtype == np.float64 and Y.dtype == np.float64
    return scipy.linalg.solve(np.array(Kxx, copy=True), Y, overwrite_a=True, overwrite_b=False,
                              check_finite=False, assume_a='pos', lower=False)


def diag_add(K, jitter):  # classify_gp.py:30-36
    K.flat[::K.shape[-1] + 1] += jitter


def predict(Kxvx, A):  # classify_gp.py:39-41
    return (Kxvx @ A).argmax(axis=1)
