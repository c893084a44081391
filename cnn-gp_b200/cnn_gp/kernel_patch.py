"""The (same, diag, xy, xx, yy) state of the recursion with the reference's two view
conventions (reference cnn_gp/kernel_patch.py:4-89):

  ConvKP    xy [Nx*Ny, 1, W, H] (diag: [Nx, 1, W, H]), xx [Nx, 1, W, H], yy [Ny, 1, W, H]
  NonlinKP  xy [Nx, Ny, W, H], xx [Nx, 1, W, H], yy [Ny, W, H] (broadcast-ready); the diag
            variant keeps all three as [N, 1, W, H]

Inside the fused kernels this state never exists in HBM; these classes are kept for code that
drives ``module.propagate(kp)`` by hand.
"""
__all__ = ('ConvKP', 'NonlinKP')


class KernelPatch:
    def __init__(self, same_or_kp, diag=False, xy=None, xx=None, yy=None):
        if isinstance(same_or_kp, KernelPatch):
            src = same_or_kp
            same, diag, xy, xx, yy = src.same, src.diag, src.xy, src.xx, src.yy
        else:
            same = same_or_kp
        self.Nx, self.Ny = xx.size(0), yy.size(0)
        self.W, self.H = xy.size(-2), xy.size(-1)
        self.same, self.diag = same, diag
        self.xy, self.xx, self.yy = self._views(diag, xy, xx, yy)

    def _views(self, diag, xy, xx, yy):
        raise NotImplementedError

    def _combine(self, other, fn):
        cls = type(self)
        if isinstance(other, KernelPatch):
            other = cls(other)
            assert self.same == other.same
            assert self.diag == other.diag
            parts = (fn(self.xy, other.xy), fn(self.xx, other.xx), fn(self.yy, other.yy))
        else:
            parts = (fn(self.xy, other), fn(self.xx, other), fn(self.yy, other))
        return cls(self.same, self.diag, *parts)

    def __add__(self, other):
        return self._combine(other, lambda a, b: a + b)

    def __mul__(self, other):
        return self._combine(other, lambda a, b: a * b)

    __radd__ = __add__
    __rmul__ = __mul__


class ConvKP(KernelPatch):
    def _views(self, diag, xy, xx, yy):
        W, H = self.W, self.H
        n_xy = self.Nx if diag else self.Nx * self.Ny
        return xy.view(n_xy, 1, W, H), xx.view(self.Nx, 1, W, H), yy.view(self.Ny, 1, W, H)


class NonlinKP(KernelPatch):
    def _views(self, diag, xy, xx, yy):
        W, H = self.W, self.H
        if diag:
            return xy.view(self.Nx, 1, W, H), xx.view(self.Nx, 1, W, H), yy.view(self.Ny, 1, W, H)
        return xy.view(self.Nx, self.Ny, W, H), xx.view(self.Nx, 1, W, H), yy.view(self.Ny, W, H)
