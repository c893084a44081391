#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the authoring container only (the reference lives at /root/reference and
does not travel to the GPU box):

    python tests/golden/make_golden.py

The reference package is imported as-is (``numpy.int`` is shimmed because
cnn_gp/data.py:12 uses the alias that numpy 2 removed).  Nothing from this
repository is imported here, so the fixtures are independent of our code.

Outputs (all small, committed):
  gram_<case>.npz      inputs X, Z and the reference's Kxx / Kxz / Kdiag ... in f32 and f64
  tiles.json           ProductIterator tile enumeration per (N, N2, bs, n_workers, rank)
  save_k_layout.npz    which entries the reference's save_K loop writes (NaN elsewhere)
  solve.npz            classify_gp.solve_system-equivalent scipy solve + argmax decisions
"""
import importlib
import json
import os
import sys

import numpy as np

np.int = int  # cnn_gp/data.py:12
REF = "/root/reference"
sys.path.insert(0, REF)

import torch  # noqa: E402
import cnn_gp as ref  # noqa: E402

assert os.path.realpath(ref.__file__).startswith(REF), ref.__file__
HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def _np(t):
    return t.detach().cpu().numpy()


def gram_case(name, model, X, Z, extra=None):
    """Evaluate every call form of kernels.py:18-57 in f32 and f64."""
    out = {"X": _np(X), "Z": _np(Z)}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        m = model.to(dt)
        x, z = X.to(dt), Z.to(dt)
        with torch.no_grad():
            out[f"Kxx_{tag}"] = _np(m(x))
            out[f"Kxz_{tag}"] = _np(m(x, z))
            out[f"Kxx_diag_{tag}"] = _np(m(x, diag=True))
            n = min(len(x), len(z))
            out[f"Kxz_diag_{tag}"] = _np(m(x[:n], z[:n], diag=True))
            # same=True with different data (forces diagonal := xx path, kernels.py:155-162)
            out[f"Kxz_same_{tag}"] = _np(m(x[:n], z[:n], same=True))
            out[f"Kxz_same_diag_{tag}"] = _np(m(x[:n], z[:n], same=True, diag=True))
    if extra:
        out.update(extra)
    np.savez_compressed(os.path.join(HERE, f"gram_{name}.npz"), **out)
    print(name, {k: v.shape for k, v in out.items() if k.startswith("K") and k.endswith("f32")})


def main():
    from cnn_gp import Conv2d, ReLU, Sequential, Sum, Mixture, resnet_block

    # ---- config 1: README model, README inputs (README.md:22-46) -----------------
    torch.manual_seed(0)
    X = torch.randn(2, 3, 28, 28)
    Z = torch.randn(2, 3, 28, 28)
    readme = Sequential(Conv2d(kernel_size=3), ReLU(), Conv2d(kernel_size=3, stride=2),
                        ReLU(), Conv2d(kernel_size=14, padding=0))
    gram_case("readme", readme, X, Z)

    # ---- configs 2..5: the shipped config modules --------------------------------
    g = torch.Generator().manual_seed(1234)
    for cfg_name, C, S in (("mnist_paper_convnet_gp", 1, 28),
                           ("mnist_paper_residual_cnn_gp", 1, 28),
                           ("mnist_as_tf", 1, 28),
                           ("mnist", 1, 28),
                           ("cifar10", 3, 32)):
        cfg = importlib.import_module(f"configs.{cfg_name}")
        X = torch.rand(6, C, S, S, generator=g)
        Z = torch.rand(5, C, S, S, generator=g)
        gram_case(cfg_name, cfg.initial_model, X, Z)
        # randn variant (negative correlations, parity only)
        Xn = torch.randn(4, C, S, S, generator=g)
        Zn = torch.randn(4, C, S, S, generator=g)
        gram_case(cfg_name + "_randn", cfg.initial_model, Xn, Zn)

    # ---- edge programs -----------------------------------------------------------
    X = torch.rand(4, 2, 12, 12, generator=g)
    Z = torch.rand(3, 2, 12, 12, generator=g)
    X[1] = 0.0  # all-zero image: f32_tiny path (kernels.py:133,146)
    edge = {
        # even kernel "same" (zero first row/col, kernels.py:73-84), bias, Sum of unequal branches
        "evenk_sum": Sequential(
            Conv2d(4, var_weight=1.7, var_bias=0.3), ReLU(),
            Sum([Conv2d(1, var_weight=0.5), Sequential(Conv2d(3), ReLU(), Conv2d(2))]),
            ReLU(), Conv2d(12, padding=0)),
        # dilation, explicit padding, stride 3
        "dilated": Sequential(
            Conv2d(3, dilation=2), ReLU(), Conv2d(3, stride=3, padding=1, var_bias=0.1),
            ReLU(), Conv2d(2, padding=0, dilation=3)),
        # nested Sum, empty Sequential, Mixture with non-uniform logits
        "nested": Sequential(
            Conv2d(3),
            Sum([Sequential(), Sequential(ReLU(), Sum([Sequential(), Conv2d(3)]), ReLU(), Conv2d(3))]),
            Mixture([Conv2d(1), Sequential(ReLU(), Conv2d(5, var_bias=0.2))],
                    logit_proportions=torch.tensor([0.3, -0.7])),
            ReLU(), Conv2d(12, padding=0)),
        # no ReLU at all (pure linear kernel) and a single dense layer
        "linear": Sequential(Conv2d(3, var_bias=0.5), Conv2d(12, padding=0)),
    }
    for name, model in edge.items():
        gram_case("edge_" + name, model, X, Z)

    # non-square input (kernels.py:61,95-96 allow it): 10x14 -> k3 same -> k10 s4 p0 -> 1x2
    # -> k3 s3 p1 -> 1x1
    Xr = torch.rand(3, 2, 10, 14, generator=g)
    Zr = torch.rand(4, 2, 10, 14, generator=g)
    gram_case("edge_nonsquare",
              Sequential(Conv2d(3), ReLU(), Conv2d(10, stride=4, padding=0, var_bias=0.05), ReLU(),
                         Conv2d(3, stride=3, padding=1)), Xr, Zr)
    # Conv2d.propagate alone on a non-square patch, even k, stride 2
    from cnn_gp.kernel_patch import ConvKP
    xy = torch.rand(6, 1, 10, 14, generator=g, dtype=torch.float64)
    conv = Conv2d(4, stride=2, var_weight=0.9, var_bias=0.2).double()
    kp = conv.propagate(ConvKP(False, False, xy, xy[:2], xy[:3]))
    np.savez_compressed(os.path.join(HERE, "conv_nonsquare.npz"), xy=_np(xy), out=_np(kp.xy))

    # ---- tile enumeration (data.py:11-60) ---------------------------------------
    from cnn_gp.data import _product_generator, _this_worker_batch, _round_up_div
    tiles = []
    for (N, N2, bs) in ((10, None, 3), (7, None, 7), (1, None, 4), (9, 5, 2), (12, 12, 5),
                        (1000, None, 200), (1000, 600, 200)):
        nbx = _round_up_div(N, bs)
        if N2 is None:
            same, nb2 = True, nbx
            total = max(1, nbx * (nbx + 1) // 2)
        else:
            same, nb2 = False, _round_up_div(N2, bs)
            total = nbx * nb2
        full = [list(map(int, t)) for t in _product_generator(nbx, nb2, same)]
        for nw in (1, 2, 3, 8):
            for r in range(nw):
                start, cnt = _this_worker_batch(total, r, nw)
                tiles.append(dict(N=N, N2=N2, bs=bs, n_workers=nw, rank=r, start=start, count=cnt,
                                  tiles=full[start:start + cnt]))
    with open(os.path.join(HERE, "tiles.json"), "w") as f:
        json.dump(tiles, f)
    print("tiles cases", len(tiles))

    # ---- save_K layout through the reference loop with a fake h5py file ----------
    class FakeDS:
        def __init__(self, shape, dtype, fillvalue, chunks, maxshape):
            self.a = np.full(shape, fillvalue, dtype=dtype)
            self.chunks, self.maxshape = chunks, maxshape

        def __setitem__(self, k, v):
            self.a[k] = v

    class FakeFile(dict):
        def create_dataset(self, name, shape, dtype, fillvalue, chunks, maxshape):
            self[name] = FakeDS(shape, dtype, fillvalue, chunks, maxshape)
            return self[name]

    class DS(torch.utils.data.Dataset):
        def __init__(self, x):
            self.x = x

        def __len__(self):
            return len(self.x)

        def __getitem__(self, i):
            return self.x[i], 0

    Xs = torch.rand(11, 3, 28, 28, generator=g)
    Xt = torch.rand(5, 3, 28, 28, generator=g)

    def kern(x, x2, same, diag):
        with torch.no_grad():
            return readme.float()(x, x2, same, diag).numpy()

    lay = {"Xs": _np(Xs), "Xt": _np(Xt)}
    meta = {}
    for nw in (1, 3):
        for r in range(nw):
            f = FakeFile()
            ref.save_K(f, kern, "Kxx", DS(Xs), None, diag=False, batch_size=4, worker_rank=r, n_workers=nw)
            ref.save_K(f, kern, "Kxtx", DS(Xt), DS(Xs), diag=False, batch_size=4, worker_rank=r, n_workers=nw)
            if r == 0:
                ref.save_K(f, kern, "Kt_diag", DS(Xt), None, diag=True, batch_size=4)
            for k, ds in f.items():
                lay[f"{k}_nw{nw}_r{r}"] = ds.a
                meta[f"{k}_nw{nw}_r{r}"] = dict(chunks=list(ds.chunks), maxshape=[m for m in ds.maxshape])
    lay["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "save_k_layout.npz"), **lay)

    # ---- solve (classify_gp.py:17-42) -------------------------------------------
    import scipy.linalg
    cfg = importlib.import_module("configs.mnist_paper_convnet_gp")
    T = torch.rand(10, 1, 28, 28, generator=g)
    ytr = torch.randint(10, (300,), generator=g)
    yte = torch.randint(10, (100,), generator=g)
    Xtr = 0.6 * T[ytr] + 0.4 * torch.rand(300, 1, 28, 28, generator=g)
    Xte = 0.6 * T[yte] + 0.4 * torch.rand(100, 1, 28, 28, generator=g)
    with torch.no_grad():
        Kxx = cfg.initial_model.float()(Xtr).numpy()
        Kxtx = cfg.initial_model.float()(Xte, Xtr).numpy()
    Y = -np.ones((300, 10)); Y[np.arange(300), ytr.numpy()] = 1.0
    K64 = Kxx.astype(np.float64)
    Kup = np.triu(K64)  # only the upper triangle is ever written (save_K) / read (lower=False)
    A = scipy.linalg.solve(Kup.copy(), Y, overwrite_a=True, overwrite_b=False,
                           check_finite=False, assume_a='pos', lower=False)
    F = Kxtx.astype(np.float64) @ A
    np.savez_compressed(os.path.join(HERE, "solve.npz"), Xtr=_np(Xtr), Xte=_np(Xte), ytr=ytr.numpy(),
                        yte=yte.numpy(), Kxx=Kxx, Kxtx=Kxtx, Y=Y, A=A, F=F, pred=F.argmax(1))
    print("solve acc", (F.argmax(1) == yte.numpy()).mean(), "cond", np.linalg.cond(K64))


if __name__ == "__main__":
    main()
