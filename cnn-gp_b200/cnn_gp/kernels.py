"""Drop-in ``cnn_gp.kernels``: same classes, constructor arguments and call protocol as the
reference (cnn_gp/kernels.py:9-10 ``__all__``), backed by the sm_100a kernels in libcnngp.so.

``model(x, y=None, same=None, diag=False)`` (reference kernels.py:18-57) compiles the module
tree into a layer program (program.py) and evaluates the whole Gram tile in one fused kernel
launch; no intermediate patch is materialised in HBM.  ``module.propagate(kp)`` keeps the
reference's per-module protocol for code that drives modules by hand; it runs the map-level
CUDA kernels.  ``nn()`` / ``layers()`` build the matching finite network exactly like the
reference does (plain PyTorch; not part of the accelerated path).

Inference only: outputs carry no autograd graph (the reference's save_kernel runs under
``no_grad``, exp_mnist_resnet/save_kernel.py:22).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import engine
from .kernel_patch import ConvKP, NonlinKP

__all__ = ("NNGPKernel", "Conv2d", "ReLU", "Sequential", "Mixture",
           "MixtureModule", "Sum", "SumModule", "resnet_block")


class NNGPKernel(nn.Module):
    """A kernel transformation [N1, N2, W, H] -> [N1, N2, W, H] (reference kernels.py:13-17)."""

    def forward(self, x, y=None, same=None, diag=False):
        """Kernel between minibatches ``x`` and ``y`` (or ``x`` with itself).

        Same contract as reference kernels.py:18-57: 4-D inputs with equal C, W, H; ``diag``
        needs equal lengths; returns ``[N1, N2]`` (``[N1]`` when ``diag``) on the input device
        in the input dtype.  The final map must be 1x1 (RuntimeError otherwise, as the
        reference's ``view`` raises).
        """
        if y is None:
            assert same is None
            y, same = x, True
        assert not diag or len(x) == len(y), (
            "diagonal kernels must operate with data of equal length")
        assert x.dim() == 4 and y.dim() == 4
        assert x.shape[1:] == y.shape[1:]
        return engine.gram(self, x, y, bool(same), bool(diag))

    # program emission, see program.py
    def _emit(self, builder, src, owned):
        raise NotImplementedError

    def propagate(self, kp):
        raise NotImplementedError

    def _program_signature(self):
        """Everything the compiled program depends on: the module tree with each layer's
        hyperparameters (read on every call, like the reference's propagate) and the Mixture weights."""
        sig = []
        for m in self.modules():
            if isinstance(m, Conv2d):
                sig.append(("C", m.kernel_size, m.stride, m.padding, m.dilation, float(m.var_weight), float(m.var_bias),
                            bool(m.kernel_has_row_of_zeros)))
            elif isinstance(m, Mixture):
                sig.append(("M", len(m.mods)) + tuple(F.softmax(m.logit.detach().double(), dim=0).tolist()))
            elif isinstance(m, Sum):
                sig.append(("S", len(m.mods)))
            elif isinstance(m, Sequential):
                sig.append(("Q", len(m.mods)))
            else:
                sig.append((type(m).__name__,))
        return tuple(sig)


class Conv2d(NNGPKernel):
    """Convolution layer; arguments as reference kernels.py:61-63.  ``padding="same"`` pads
    ``dilation*(kernel_size//2)``; for even kernel sizes the reference realises it with a
    (k+1)x(k+1) kernel whose first row and column are zero (kernels.py:71-84), which is kept
    as the ``kernel_has_row_of_zeros`` flag and the registered ``kernel`` buffer."""

    def __init__(self, kernel_size, stride=1, padding="same", dilation=1,
                 var_weight=1., var_bias=0., in_channel_multiplier=1,
                 out_channel_multiplier=1):
        super().__init__()
        self.kernel_size, self.stride, self.dilation = kernel_size, stride, dilation
        self.var_weight, self.var_bias = var_weight, var_bias
        self.in_channel_multiplier = in_channel_multiplier
        self.out_channel_multiplier = out_channel_multiplier
        same_pad = padding == "same"
        self.kernel_has_row_of_zeros = bool(same_pad and kernel_size % 2 == 0)
        self.padding = dilation * (kernel_size // 2) if same_pad else padding
        extent = kernel_size + int(self.kernel_has_row_of_zeros)
        box = torch.full((1, 1, extent, extent), var_weight / kernel_size ** 2)
        if self.kernel_has_row_of_zeros:
            box[..., 0, :] = 0.
            box[..., :, 0] = 0.
        self.register_buffer("kernel", box)

    def _emit(self, builder, src, owned):
        return builder.conv(src, owned, self.kernel_size, self.kernel_has_row_of_zeros, self.stride,
                            self.padding, self.dilation, self.var_weight, self.var_bias)

    def propagate(self, kp):
        kp = ConvKP(kp)
        maps = engine.conv_maps
        return ConvKP(kp.same, kp.diag, maps(self, kp.xy), maps(self, kp.xx), maps(self, kp.yy))

    def nn(self, channels, in_channels=None, out_channels=None):
        cin = (channels if in_channels is None else in_channels) * self.in_channel_multiplier
        cout = (channels if out_channels is None else out_channels) * self.out_channel_multiplier
        has_bias = self.var_bias > 0.
        layer = nn.Conv2d(cin, cout, self.kernel.shape[-1], stride=self.stride, padding=self.padding,
                          dilation=self.dilation, bias=has_bias)
        with torch.no_grad():
            layer.weight.normal_(0, math.sqrt(self.var_weight / cin) / self.kernel_size)
            if self.kernel_has_row_of_zeros:
                layer.weight[:, :, 0, :] = 0
                layer.weight[:, :, :, 0] = 0
            if has_bias:
                layer.bias.normal_(0, math.sqrt(self.var_bias))
        return layer

    def layers(self):
        return 1


class ReLU(NNGPKernel):
    """ReLU nonlinearity: the arccos expectation of reference kernels.py:134-165."""

    def _emit(self, builder, src, owned):
        return builder.relu(src, owned)

    def propagate(self, kp):
        kp = NonlinKP(kp)
        xy = engine.relu_maps(kp)
        xx = kp.xx / 2.
        yy = xx if kp.same else kp.yy / 2.
        return NonlinKP(kp.same, kp.diag, xy, xx, yy)

    def nn(self, channels, in_channels=None, out_channels=None):
        assert in_channels is None
        assert out_channels is None
        return nn.ReLU()

    def layers(self):
        return 0


class _Container(NNGPKernel):
    def _register(self, mods):
        self.mods = mods
        for idx, mod in enumerate(mods):
            self.add_module(str(idx), mod)


class Sequential(_Container):
    def __init__(self, *mods):
        super().__init__()
        self._register(mods)

    def _emit(self, builder, src, owned):
        return builder.sequential(self.mods, src, owned)

    def propagate(self, kp):
        for mod in self.mods:
            kp = mod.propagate(kp)
        return kp

    def nn(self, channels, in_channels=None, out_channels=None):
        n = len(self.mods)
        if n == 0:
            return nn.Sequential()
        if n == 1:
            return self.mods[0].nn(channels, in_channels=in_channels, out_channels=out_channels)
        first = self.mods[0].nn(channels, in_channels=in_channels)
        middle = [mod.nn(channels) for mod in self.mods[1:-1]]
        last = self.mods[-1].nn(channels, out_channels=out_channels)
        return nn.Sequential(first, *middle, last)

    def layers(self):
        return sum(mod.layers() for mod in self.mods)


class Sum(_Container):
    """Applies every module to the same input and adds the results (reference kernels.py:246-254)."""

    def __init__(self, mods):
        super().__init__()
        self._register(mods)

    def _emit(self, builder, src, owned):
        return builder.branches(self.mods, src, owned)

    def propagate(self, kp):
        return sum(m.propagate(kp) for m in self.mods)

    def nn(self, channels, in_channels=None, out_channels=None):
        return SumModule([m.nn(channels, in_channels=in_channels, out_channels=out_channels)
                          for m in self.mods])

    def layers(self):
        return max(mod.layers() for mod in self.mods)


class Mixture(_Container):
    """Softmax-weighted sum of modules (reference kernels.py:203-229)."""

    def __init__(self, mods, logit_proportions=None):
        super().__init__()
        self._register(mods)
        if logit_proportions is None:
            logit_proportions = torch.zeros(len(mods))
        self.logit = nn.Parameter(logit_proportions)

    def _proportions(self):
        return F.softmax(self.logit.detach(), dim=0)

    def _emit(self, builder, src, owned):
        return builder.branches(self.mods, src, owned, weights=self._proportions().tolist())

    def propagate(self, kp):
        prop = self._proportions()
        total = self.mods[0].propagate(kp) * prop[0]
        for i in range(1, len(self.mods)):
            total = total + self.mods[i].propagate(kp) * prop[i]
        return total

    def nn(self, channels, in_channels=None, out_channels=None):
        return MixtureModule([m.nn(channels, in_channels=in_channels, out_channels=out_channels)
                              for m in self.mods], self.logit)

    def layers(self):
        return max(mod.layers() for mod in self.mods)


class _ModuleList(nn.Module):
    def __init__(self, mods):
        super().__init__()
        self.mods = mods
        for idx, mod in enumerate(mods):
            self.add_module(str(idx), mod)


class SumModule(_ModuleList):
    def forward(self, input):
        return sum(m(input) for m in self.mods)


class MixtureModule(_ModuleList):
    def __init__(self, mods, logit_parameter):
        super().__init__(mods)
        self.logit = torch.as_tensor(logit_parameter).detach().clone()

    def forward(self, input):
        # the reference scales only the first branch (kernels.py:239-243)
        w = F.softmax(self.logit, dim=0).sqrt()
        total = self.mods[0](input) * w[0]
        for m in self.mods[1:]:
            total = total + m(input)
        return total


def resnet_block(stride=1, projection_shortcut=False, multiplier=1):
    """Pre-activation residual block (reference kernels.py:274-296): identity shortcut when the
    shape is unchanged, otherwise ReLU followed by a 1x1 projection in parallel with the
    two-conv branch."""
    def conv3(s, cin, cout):
        return Conv2d(3, stride=s, in_channel_multiplier=cin, out_channel_multiplier=cout)

    if stride == 1 and not projection_shortcut:
        body = Sequential(ReLU(), conv3(stride, multiplier, multiplier), ReLU(),
                          conv3(1, multiplier, multiplier))
        return Sum([Sequential(), body])
    narrow = multiplier // stride
    shortcut = Conv2d(1, stride=stride, in_channel_multiplier=narrow, out_channel_multiplier=multiplier)
    body = Sequential(conv3(stride, narrow, multiplier), ReLU(), conv3(1, multiplier, multiplier))
    return Sequential(ReLU(), Sum([shortcut, body]))
