"""Host side of the Gram recursion: argument handling, plan cache, launches.

Everything here enqueues work on torch's current CUDA stream through the C ABI of
include/cnngp.h; torch is only used for device memory and streams.  There is no CPU route:
CPU tensors raise.
"""
import ctypes
import weakref

import torch

from . import _native as nat
from . import program

_DTYPE_CODE = {torch.float32: nat.F32, torch.float64: nat.F64}

# "auto" | "generic" | "fused" -- tests and benchmarks pin a path through this knob
_force_path = "auto"
_PATH_CODE = {"auto": nat.PATH_AUTO, "generic": nat.PATH_GENERIC, "fused": nat.PATH_FUSED}


def set_path(name):
    """Select the kernel family: "auto" (fused when the program is covered), "generic", "fused"."""
    global _force_path
    if name not in _PATH_CODE:
        raise ValueError(name)
    prev, _force_path = _force_path, name
    return prev


def last_path():
    return {0: "none", 1: "generic", 2: "fused", 3: "fused_net"}[nat.lib().cnngp_last_path()]


def last_launches():
    """Kernels the last Gram call of this thread launched (two per chunk of super-tiles for programs with a
    folded phase, one otherwise)."""
    return nat.lib().cnngp_last_launches()


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(
            f"cnn_gp (B200): {what} is on {t.device}; this implementation has no CPU path. "
            "Move the model inputs to a CUDA device.")
    if t.dtype not in _DTYPE_CODE:
        raise TypeError(f"cnn_gp (B200): unsupported dtype {t.dtype}; use float32 or float64")


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


# compiled plans per model object.  Kept OUTSIDE the module (a plan holds a ctypes handle, which
# copy.deepcopy / pickle / torch.save of the model must not meet) and keyed by everything the
# program depends on, so that editing a layer's hyperparameters after the first call recompiles
# (the reference reads them on every propagate, kernels.py:92-98).
_PLANS = weakref.WeakKeyDictionary()


def plan_for(model, H, W, dtype):
    """Compiled plan for ``model`` on H x W maps, cached per model object and program signature."""
    cache = _PLANS.get(model)
    if cache is None:
        cache = _PLANS[model] = {}
    key = (H, W, dtype, model._program_signature())
    plan = cache.get(key)
    if plan is None:
        ops, n_slots = program.compile_model(model)
        plan = nat.Plan(ops, n_slots, H, W, _DTYPE_CODE[dtype])
        if len(cache) >= 8:  # a model whose hyperparameters keep changing must not pile plans up
            cache.clear()
        cache[key] = plan
    return plan


def variances(plan, x, z=None):
    """Per-image variance maps at every ReLU input and the diagonal kernel value.

    -> (aux_x [N, aux_elems], aux_z or None, kdiag [N]).  With ``z`` given, evaluates the
    literal same=True semantics for two different image sets (reference kernels.py:155-156).
    """
    N, C = x.shape[0], x.shape[1]
    # an even number of rows: the fused kernel's maps interleave images 2k and 2k+1 over both rows
    aux_x = torch.empty((N + (N & 1), max(1, plan.aux_elems)), dtype=x.dtype, device=x.device)
    aux_z = torch.empty_like(aux_x) if z is not None else None
    kdiag = torch.empty((N,), dtype=x.dtype, device=x.device)
    nat.check(nat.lib().cnngp_variances(
        plan.handle, x.data_ptr(), z.data_ptr() if z is not None else None, N, C,
        aux_x.data_ptr(), aux_z.data_ptr() if aux_z is not None else None, kdiag.data_ptr(),
        _stream()), "cnngp_variances")
    return aux_x, aux_z, kdiag


def gram_with_aux(plan, x, z, aux_x, aux_z, same, diag, symmetric, out=None, kdiag=None, path=None):
    """One cnngp_gram launch on prepared operands; ``out`` may be a (strided-row) view.
    ``path`` overrides the module-wide kernel choice for this call."""
    N1, N2, C = x.shape[0], z.shape[0], x.shape[1]
    if out is None:
        out = torch.empty((N1,) if diag else (N1, N2), dtype=x.dtype, device=x.device)
    # the kernels write elements of the plan's dtype: a buffer of another width would be overrun
    for name, t in (("z", z), ("aux_x", aux_x), ("aux_z", aux_z), ("out", out), ("kdiag", kdiag)):
        if t is not None and (t.dtype != x.dtype or t.device != x.device):
            raise TypeError(f"cnn_gp (B200): {name} is {t.dtype} on {t.device}, the images are {x.dtype} on {x.device}")
    if _DTYPE_CODE[x.dtype] != plan.dtype_code:
        raise TypeError(f"cnn_gp (B200): plan compiled for dtype code {plan.dtype_code}, images are {x.dtype}")
    ld = 1 if diag else out.stride(0)
    if not diag:
        assert out.stride(1) == 1 and out.shape == (N1, N2)
    nat.check(nat.lib().cnngp_gram(
        plan.handle, x.data_ptr(), N1, z.data_ptr(), N2, C, aux_x.data_ptr(), aux_z.data_ptr(),
        kdiag.data_ptr() if kdiag is not None else None,
        int(same), int(diag), int(symmetric), out.data_ptr(), ld, _PATH_CODE[path or _force_path], _stream()),
        "cnngp_gram")
    return out


def gram_band(plan, x, n_rows, aux, kdiag, block, out):
    """Rows [0, n_rows) x all columns of model(x) in one launch (cnngp_gram_band): entries j >= i, mirrored inside
    the diagonal blocks of ``block`` rows -- the reference's tiles (i, j >= i) of these block rows.  ``out``:
    [n_rows, len(x)] view with unit column stride; what lies below the diagonal blocks is left untouched."""
    N2, C = x.shape[0], x.shape[1]
    for name, t in (("aux", aux), ("out", out), ("kdiag", kdiag)):
        if t.dtype != x.dtype or t.device != x.device:
            raise TypeError(f"cnn_gp (B200): {name} is {t.dtype} on {t.device}, the images are {x.dtype} on {x.device}")
    assert out.stride(1) == 1 and out.shape == (n_rows, N2) and x.is_contiguous()
    nat.check(nat.lib().cnngp_gram_band(plan.handle, x.data_ptr(), n_rows, N2, C, aux.data_ptr(), kdiag.data_ptr(),
                                        int(block), out.data_ptr(), out.stride(0), _stream()), "cnngp_gram_band")
    return out


@torch.no_grad()
def gram(model, x, y, same, diag):
    """model(x, y, same, diag) -- reference kernels.py:18-57."""
    if not x.is_cuda and not y.is_cuda and _model_device(model) is not None:
        # host tensors, model on a GPU: host in, host out (the reference returns on the input device)
        return gram_host(model, x, None if y is x else y, same, diag)
    _require_cuda(x, "x")
    _require_cuda(y, "y")
    if x.dtype != y.dtype or x.device != y.device:
        raise RuntimeError("x and y must share dtype and device")
    identical = (y is x) or (y.data_ptr() == x.data_ptr() and y.shape == x.shape
                             and y.stride() == x.stride())
    x = x.detach().contiguous()
    if same and not identical and not diag and x.shape == y.shape:
        # the reference's tile driver uploads the diagonal tile's batch twice
        # (save_kernel.py:23-24): equal images take the symmetric route (j >= i, mirrored)
        identical = bool(torch.equal(x, y))
    y = x if identical else y.detach().contiguous()
    N1, N2 = x.shape[0], y.shape[0]
    if same and not diag and N1 != N2:
        # the reference fails broadcasting eye(N1) against [N1, N2, W, H] (kernels.py:161-162)
        raise RuntimeError(f"same=True needs equally many images, got {N1} and {N2}")
    with torch.cuda.device(x.device):
        plan = plan_for(model, x.shape[2], x.shape[3], x.dtype)
        if N1 == 0 or N2 == 0:
            return torch.empty((N1,) if diag else (N1, N2), dtype=x.dtype, device=x.device)
        symmetric = bool(same and identical)
        if same and not identical:
            aux_x, aux_z, kdiag = variances(plan, x, y)
        else:
            aux_x, _, kdiag = variances(plan, x)
            aux_z = aux_x if identical else variances(plan, y)[0]
        return gram_with_aux(plan, x, y, aux_x, aux_z, same, diag, symmetric,
                             kdiag=kdiag if symmetric else None)


def _model_device(model):
    for t in list(model.buffers()) + list(model.parameters()):
        if t.is_cuda:
            return t.device
    return None


@torch.no_grad()
def gram_host(model, x, y=None, same=None, diag=False, out=None):
    """``model(x, y, same, diag)`` for HOST tensors and a model that lives on a GPU: host in, host out.

    This is the round trip the reference's tile driver makes per tile (``model(x.cuda(), ...)
    .cpu()``, exp_mnist_resnet/save_kernel.py:21-24), as one call.  For ``model(X)`` on a program
    a fused kernel covers, bands of finished rows are copied to the (pinned) result
    while the kernel is still running (cnngp_gram_symmetric_to_host); anything else is upload,
    compute, copy.  ``out``: optional pinned float32 ``[N1, N2]`` result buffer to reuse."""
    dev = _model_device(model)
    if dev is None:
        raise RuntimeError("cnn_gp (B200): neither the inputs nor the model are on a CUDA device; "
                           "this implementation has no CPU path")
    if y is None:
        y, same = x, True if same is None else same
    identical = y is x
    with torch.cuda.device(dev):
        xd = x.to(dev, non_blocking=True)
        yd = xd if identical else y.to(dev, non_blocking=True)
        N = xd.shape[0]
        streamed = (identical and same and not diag and xd.dtype == torch.float32 and N > 0
                    and _force_path != "generic")
        if streamed:
            xd = xd.contiguous()
            plan = plan_for(model, xd.shape[2], xd.shape[3], xd.dtype)
            streamed = plan.fused_kind in (2, 3)
        if not streamed:
            res = gram(model, xd, yd, bool(same), bool(diag))
            if out is None:
                return res.cpu()
            out.copy_(res)
            return out
        aux, _, kdiag = variances(plan, xd)
        K = torch.empty((N, N), dtype=torch.float32, device=dev)
        if out is None:
            out = torch.empty((N, N), dtype=torch.float32, pin_memory=True)
        assert out.shape == (N, N) and out.dtype == torch.float32 and out.stride(1) == 1 and not out.is_cuda
        scratch = torch.empty(max(64, N // 32 + 8), dtype=torch.int32, device=dev)  # one counter per band of >= 48 rows
        side = torch.cuda.Stream(dev)
        rc = nat.lib().cnngp_gram_symmetric_to_host(
            plan.handle, xd.data_ptr(), N, xd.shape[1], aux.data_ptr(), kdiag.data_ptr(), K.data_ptr(), K.stride(0),
            out.data_ptr(), out.stride(0), scratch.data_ptr(), scratch.numel() * 4, _stream(),
            ctypes.c_void_p(side.cuda_stream))
        if rc == 7:  # no stream memory operations on this driver / setup: plain launch + copy
            gram_with_aux(plan, xd, xd, aux, aux, True, False, True, out=K, kdiag=kdiag)
            out.copy_(K)
            return out
        nat.check(rc, "cnngp_gram_symmetric_to_host")
        side.synchronize()
        return out


def _conv_op(mod):
    import numpy as np
    zf = bool(mod.kernel_has_row_of_zeros)
    return nat.Op(opcode=nat.OP_CONV, src=0, dst=0, ke=int(mod.kernel_size) + int(zf), zero_first=int(zf),
                  stride=int(mod.stride), pad=int(mod.padding), dil=int(mod.dilation),
                  scale=float(np.float32(float(mod.var_weight) / int(mod.kernel_size) ** 2)),
                  bias=float(mod.var_bias))


@torch.no_grad()
def conv_maps(mod, maps):
    """F.conv2d(patch, box) + var_bias of reference kernels.py:94-97 on [M, 1, W, H] maps."""
    _require_cuda(maps, "kernel patch")
    maps = maps.contiguous()
    M, _, Hi, Wi = maps.shape
    op = _conv_op(mod)

    def osz(n):
        return (n + 2 * op.pad - op.dil * (op.ke - 1) - 1) // op.stride + 1
    Ho, Wo = osz(Hi), osz(Wi)
    if Ho < 1 or Wo < 1:
        raise RuntimeError("conv output would be empty")
    out = torch.empty((M, 1, Ho, Wo), dtype=maps.dtype, device=maps.device)
    with torch.cuda.device(maps.device):
        nat.check(nat.lib().cnngp_conv_maps(maps.data_ptr(), M, Hi, Wi, ctypes.byref(op),
                                            _DTYPE_CODE[maps.dtype], out.data_ptr(), _stream()),
                  "cnngp_conv_maps")
    return out


@torch.no_grad()
def relu_maps(kp):
    """xy' of reference kernels.py:146-162 for a NonlinKP."""
    _require_cuda(kp.xy, "kernel patch")
    if kp.same and not kp.diag and kp.Nx != kp.Ny:
        raise RuntimeError("same=True needs Nx == Ny")
    xy = kp.xy.contiguous().clone()
    xx, yy = kp.xx.contiguous(), kp.yy.contiguous()
    with torch.cuda.device(xy.device):
        nat.check(nat.lib().cnngp_relu_maps(xy.data_ptr(), xx.data_ptr(), yy.data_ptr(), kp.Nx, kp.Ny,
                                            kp.W * kp.H, int(bool(kp.same)), int(bool(kp.diag)),
                                            _DTYPE_CODE[xy.dtype], _stream()), "cnngp_relu_maps")
    return xy
