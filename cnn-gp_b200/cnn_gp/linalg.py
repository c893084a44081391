"""The dense fp64 stage of the GP classifier on the GPU: SPD solve from the upper triangle and
prediction, through the C ABI (include/cnngp.h: cnngp_potrf_upper_f64, cnngp_potrs_upper_f64,
cnngp_predict_argmax).

Replaces, for CUDA tensors, the reference's
    scipy.linalg.solve(Kxx, Y, assume_a='pos', lower=False)   exp_mnist_resnet/classify_gp.py:24-26
    (Kxvx @ A).argmax(dim=1)                                   exp_mnist_resnet/classify_gp.py:40
Only the upper triangle of ``Kxx`` is read, so the NaN blocks ``save_K`` leaves below the block
diagonal (cnn_gp/kernel_save_tools.py:21-23) are harmless, exactly as with LAPACK ``uplo='U'``.
There is no CPU route.
"""
import ctypes

import torch

from . import _native as nat


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check_matrix(A, name, dtype):
    if not A.is_cuda:
        raise RuntimeError(f"cnn_gp.linalg: {name} is on {A.device}; this implementation has no CPU path")
    if A.dtype != dtype:
        raise TypeError(f"cnn_gp.linalg: {name} must be {dtype}, got {A.dtype}")
    if A.dim() != 2 or A.stride(1) != 1:
        raise ValueError(f"cnn_gp.linalg: {name} must be a row-major 2-D tensor")


class NotPositiveDefiniteError(RuntimeError):
    """The leading minor of order ``info`` is not positive definite (LAPACK dpotrf info > 0;
    scipy raises LinAlgError at the same point)."""

    def __init__(self, info):
        super().__init__(f"the leading minor of order {info} is not positive definite")
        self.info = info


@torch.no_grad()
def potrf_upper_(A, check=True):
    """In place: the upper triangle of the square float64 matrix ``A`` becomes U with A = U^T U.
    Returns LAPACK's ``info`` (0 = success); raises NotPositiveDefiniteError when ``check``."""
    _check_matrix(A, "A", torch.float64)
    n = A.shape[0]
    assert A.shape[1] == n, "A must be square"
    info = torch.zeros(1, dtype=torch.int32, device=A.device)
    with torch.cuda.device(A.device):
        nat.check(nat.lib().cnngp_potrf_upper_f64(A.data_ptr(), n, A.stride(0) if n > 1 else max(n, 1),
                                                  info.data_ptr(), _stream()), "cnngp_potrf_upper_f64")
    if not check:
        return info
    code = int(info.item())
    if code != 0:
        raise NotPositiveDefiniteError(code)
    return 0


@torch.no_grad()
def potrs_upper_(U, B):
    """In place on ``B`` [n, nrhs]: solve U^T U X = B with U from ``potrf_upper_``."""
    _check_matrix(U, "U", torch.float64)
    _check_matrix(B, "B", torch.float64)
    n = U.shape[0]
    assert B.shape[0] == n
    with torch.cuda.device(U.device):
        nat.check(nat.lib().cnngp_potrs_upper_f64(U.data_ptr(), n, U.stride(0) if n > 1 else max(n, 1), B.data_ptr(),
                                                  B.shape[1], B.stride(0) if n > 1 else max(B.shape[1], 1),
                                                  _stream()), "cnngp_potrs_upper_f64")
    return B


@torch.no_grad()
def solve_pos_upper(K, Y, overwrite_a=False):
    """X = K^{-1} Y for a symmetric positive definite ``K`` given by its upper triangle."""
    U = K if overwrite_a else K.clone()
    potrf_upper_(U)
    return potrs_upper_(U, Y.clone().contiguous())


@torch.no_grad()
def predict_argmax(K, A, return_scores=False):
    """``(K.double() @ A).argmax(1)`` for a kernel block ``K`` [R, n] -- float32 as stored by save_K, or
    float64 -- and float64 weights ``A`` [n, classes]; accumulates in float64 without materialising a
    widened ``K``."""
    if K.dtype not in (torch.float32, torch.float64):
        raise TypeError(f"cnn_gp.linalg: K must be float32 or float64, got {K.dtype}")
    _check_matrix(K, "K", K.dtype)
    _check_matrix(A, "A", torch.float64)
    A = A.contiguous()
    R, n = K.shape
    assert A.shape[0] == n
    pred = torch.empty(R, dtype=torch.int64, device=K.device)
    scores = torch.empty((R, A.shape[1]), dtype=torch.float64, device=K.device) if return_scores else None
    with torch.cuda.device(K.device):
        fn = nat.lib().cnngp_predict_argmax if K.dtype == torch.float32 else nat.lib().cnngp_predict_argmax_f64
        nat.check(fn(K.data_ptr(), R, n, K.stride(0) if R > 1 else max(n, 1),
                                                 A.data_ptr(), A.shape[1], pred.data_ptr(),
                                                 scores.data_ptr() if scores is not None else None, _stream()),
                  "cnngp_predict_argmax")
    return (pred, scores) if return_scores else pred
