"""Host-side logic that needs no GPU: the drop-in API surface, the program compiler, plan
validation through the C ABI, tile enumeration, the block store and the exported symbols."""
import ctypes
import inspect
import json
import os
import re

import numpy as np
import pytest
import torch

import cnn_gp
from cnn_gp import Conv2d, ReLU, Sequential, Sum, Mixture, resnet_block
from cnn_gp import _native as nat
from cnn_gp import program, data, block_store
from models import golden_models, readme_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_public_api_matches_reference():
    # reference cnn_gp/__init__.py:1-6 + kernels.py:9-10, data.py:7-8, kernel_save_tools.py:4
    assert set(cnn_gp.__all__) == {
        "NNGPKernel", "Conv2d", "ReLU", "Sequential", "Mixture", "MixtureModule", "Sum", "SumModule",
        "resnet_block", "DatasetFromConfig", "ProductIterator", "DiagIterator", "print_timings",
        "create_h5py_dataset", "save_K"}
    sig = inspect.signature
    assert list(sig(cnn_gp.NNGPKernel.forward).parameters) == ["self", "x", "y", "same", "diag"]
    assert list(sig(Conv2d.__init__).parameters) == [
        "self", "kernel_size", "stride", "padding", "dilation", "var_weight", "var_bias",
        "in_channel_multiplier", "out_channel_multiplier"]
    assert list(sig(cnn_gp.ProductIterator.__init__).parameters) == [
        "self", "batch_size", "X", "X2", "worker_rank", "n_workers"]
    assert list(sig(cnn_gp.save_K).parameters) == [
        "f", "kern", "name", "X", "X2", "diag", "batch_size", "worker_rank", "n_workers", "print_interval"]
    assert list(sig(resnet_block).parameters) == ["stride", "projection_shortcut", "multiplier"]


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "cnngp.h")).read()
    declared = set(re.findall(r"\b(cnngp_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    L = ctypes.CDLL(nat.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/cnngp.h but not exported"
    assert declared == set(nat.EXPORTS)
    assert nat.lib().cnngp_abi_version() == 1


def test_conv_module_attributes():
    c = Conv2d(4, var_weight=3.2)
    assert c.kernel_has_row_of_zeros and c.padding == 2 and tuple(c.kernel.shape) == (1, 1, 5, 5)
    assert float(c.kernel[0, 0, 0, 3]) == 0.0 and float(c.kernel[0, 0, 2, 0]) == 0.0
    assert float(c.kernel[0, 0, 1, 1]) == np.float32(3.2 / 16)
    c = Conv2d(3, dilation=2)
    assert c.padding == 2 and not c.kernel_has_row_of_zeros
    c = Conv2d(14, padding=0)
    assert c.padding == 0
    # .double() widens the float32-rounded tap (part of the reference's semantics)
    assert float(Conv2d(3).double().kernel[0, 0, 0, 0]) == float(np.float32(1 / 9))


def test_nn_builds_matching_network():
    m = readme_model()
    net = m.nn(channels=16, in_channels=3, out_channels=10)
    y = net(torch.randn(2, 3, 28, 28))
    assert tuple(y.shape) == (2, 10, 1, 1)
    assert m.layers() == 3
    r = Sequential(Conv2d(3), resnet_block(1, True), resnet_block(), resnet_block(2, True, 2))
    assert r.layers() == 1 + 2 + 2 + 2
    y = r.nn(channels=4, in_channels=1)(torch.randn(1, 1, 28, 28))
    assert tuple(y.shape) == (1, 8, 14, 14)
    e = Conv2d(4, var_bias=0.5).nn(3)
    assert e.kernel_size == (5, 5) and float(e.weight.detach()[0, 0, 0, 2]) == 0.0 and e.bias is not None


def _simulate(ops, n_slots, shape):
    """Interpret a program symbolically: every slot holds an expression string."""
    slots = {0: "x"}
    for o in ops:
        s = slots[o.src]
        if o.opcode == nat.OP_CONV:
            slots[o.dst] = f"conv{o.ke - o.zero_first}s{o.stride}({s})"
        elif o.opcode == nat.OP_RELU:
            slots[o.dst] = f"relu({s})"
        elif o.opcode == nat.OP_COPY:
            slots[o.dst] = s
        elif o.opcode == nat.OP_SCALE:
            slots[o.dst] = f"{o.scale:.3f}*({s})"
        elif o.opcode == nat.OP_ADD:
            slots[o.dst] = f"({slots[o.dst]}+{s})"
    return slots[ops[-1].dst] if ops else slots[0]


def test_program_compiler_semantics():
    m = Sequential(Conv2d(3), Sum([Sequential(), Sequential(ReLU(), Conv2d(3))]))
    ops, ns = program.compile_model(m)
    assert _simulate(ops, ns, None) in ("(conv3s1(relu(conv3s1(x)))+conv3s1(x))",)
    assert ns == 2
    # projection block: both branches read the post-ReLU map
    m = resnet_block(2, True, 2)
    ops, ns = program.compile_model(m)
    assert _simulate(ops, ns, None) == "(conv1s2(relu(x))+conv3s1(relu(conv3s2(relu(x)))))"
    # sum of two identities needs a copy
    ops, ns = program.compile_model(Sum([Sequential(), Sequential()]))
    assert _simulate(ops, ns, None) == "(x+x)"
    # three branches keep left-to-right association up to commutativity of the first add
    m = Sum([Conv2d(1), Sequential(), Conv2d(3)])
    ops, ns = program.compile_model(m)
    assert _simulate(ops, ns, None) == "((conv1s1(x)+x)+conv3s1(x))"
    # mixture weights
    m = Mixture([Sequential(), Conv2d(3)], logit_proportions=torch.tensor([0.0, 0.0]))
    ops, ns = program.compile_model(m)
    assert _simulate(ops, ns, None) == "(0.500*(x)+0.500*(conv3s1(x)))"


def test_shipped_programs_need_two_slots_and_match_survey_flops():
    # algorithmic flop per pair
    want = {"readme": (3, 25873), "mnist_paper_convnet_gp": (1, 161505),
            "mnist_paper_residual_cnn_gp": (1, 170913), "mnist_as_tf": (1, 241147), "mnist": (1, 241147),
            "cifar10": (3, 319060)}
    relu_px = {"readme": 980, "mnist_paper_convnet_gp": 5488, "mnist_paper_residual_cnn_gp": 7056,
               "mnist_as_tf": 11026, "cifar10": 14401}
    models = golden_models()
    for name, (C, flops) in want.items():
        ops, ns = program.compile_model(models[name])
        S = 32 if name == "cifar10" else 28
        plan = nat.Plan(ops, ns, S, S, nat.F32)
        assert ns <= 2, name
        assert plan.flops_per_pair(C) == flops, (name, plan.flops_per_pair(C))
        if name in relu_px:
            # rows also carry the fused kernel's (s, 1/s) pairs when it covers the program
            # (cnngp.h: xx maps, padded to 16 bytes, then 4 * ceil(pixels / 2) floats per ReLU)
            n_relu = sum(1 for o in ops if o.opcode == nat.OP_RELU)
            assert plan.has_fused and plan.aux_elems % 4 == 0
            assert 3 * relu_px[name] <= plan.aux_elems <= 3 * relu_px[name] + 3 + 2 * n_relu


def test_plan_rejects_bad_programs():
    ops, ns = program.compile_model(readme_model())
    with pytest.raises(RuntimeError, match="not 1x1"):
        nat.Plan(ops, ns, 30, 30, nat.F32)
    with pytest.raises(RuntimeError, match="empty"):
        nat.Plan(ops, ns, 4, 4, nat.F32)
    bad = [nat.Op(opcode=nat.OP_ADD, src=1, dst=0)]
    with pytest.raises(RuntimeError, match="before it is written"):
        nat.Plan(bad, 2, 1, 1, nat.F32)


def test_band_call_validates_its_arguments_before_touching_the_gpu():
    """cnngp_gram_band (include/cnngp.h): argument errors come back as codes with a message, no CUDA call made --
    more rows than columns, a leading dimension shorter than the band, a program no fused kernel covers."""
    import ctypes
    L = nat.lib()
    ops, ns = program.compile_model(golden_models()["mnist_paper_convnet_gp"])
    plan = nat.Plan(ops, ns, 28, 28, nat.F32)
    assert plan.fused_kind == 2
    one = ctypes.c_void_p(16)  # never dereferenced: validation fails first
    assert L.cnngp_gram_band(plan.handle, one, 8, 4, 1, one, one, 2, one, 8, None) == 1
    assert b"bad arguments" in L.cnngp_last_error()
    assert L.cnngp_gram_band(plan.handle, one, 4, 8, 1, one, one, 2, one, 7, None) == 1
    assert L.cnngp_gram_band(plan.handle, None, 4, 8, 1, one, one, 2, one, 8, None) == 1
    assert L.cnngp_gram_band(plan.handle, one, 0, 8, 1, one, one, 2, one, 8, None) == 0  # nothing to do
    ops64, ns64 = program.compile_model(golden_models()["mnist_paper_convnet_gp"])
    plan64 = nat.Plan(ops64, ns64, 28, 28, nat.F64)  # float64 runs on the generic kernel only
    assert plan64.fused_kind == 0
    assert L.cnngp_gram_band(plan64.handle, one, 4, 8, 1, one, one, 2, one, 8, None) == 4
    assert b"fused" in L.cnngp_last_error()


def test_cpu_tensors_are_refused():
    m = readme_model()
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.randn(2, 3, 28, 28))
    with pytest.raises(AssertionError):
        m(torch.randn(2, 3, 28, 28), torch.randn(3, 3, 28, 28), diag=True)
    with pytest.raises(AssertionError):
        m(torch.randn(2, 3, 28, 28), torch.randn(2, 1, 28, 28))


def test_tiles_match_reference_enumeration():
    with open(os.path.join(GOLD, "tiles.json")) as f:
        cases = json.load(f)
    for c in cases:
        got = data.worker_tiles(c["N"], c["N2"], c["bs"], c["rank"], c["n_workers"])
        assert [list(map(int, t)) for t in got] == c["tiles"], c
        start, count = data._this_worker_batch(data.tile_count(c["N"], c["N2"], c["bs"]), c["rank"], c["n_workers"])
        assert (start, count) == (c["start"], c["count"])


def test_product_iterator_serves_reference_batches():
    X = data.ResidentDataset(torch.arange(10.).view(10, 1, 1, 1), torch.arange(10))
    Z = data.ResidentDataset(torch.arange(7.).view(7, 1, 1, 1) + 100, torch.arange(7))
    seen = []
    it = cnn_gp.ProductIterator(4, X, None, worker_rank=1, n_workers=2)
    assert len(it) == 3
    for same, (i, (x, y)), (j, (x2, y2)) in it:
        seen.append((same, i, j, x.flatten().tolist(), x2.flatten().tolist()))
    assert seen == [(True, 4, 4, [4., 5., 6., 7.], [4., 5., 6., 7.]), (False, 4, 8, [4., 5., 6., 7.], [8., 9.]),
                    (True, 8, 8, [8., 9.], [8., 9.])]
    it = cnn_gp.ProductIterator(4, X, Z)
    tiles = [(s, i, j, len(a[0]), len(b[0])) for s, (i, a), (j, b) in it]
    assert tiles == [(False, 0, 0, 4, 4), (False, 0, 4, 4, 3), (False, 4, 0, 4, 4), (False, 4, 4, 4, 3),
                     (False, 8, 0, 2, 4), (False, 8, 4, 2, 3)]
    # generic map-style datasets go through default_collate
    ds = torch.utils.data.TensorDataset(torch.arange(5.).view(5, 1, 1, 1), torch.arange(5))
    (same, (i, (x, y)), (j, (x2, y2))), = list(cnn_gp.ProductIterator(8, ds))
    assert same and i == 0 and j == 0 and x.shape == (5, 1, 1, 1) and y.tolist() == [0, 1, 2, 3, 4]
    d = list(cnn_gp.DiagIterator(2, X, Z))
    assert [(s, i, len(a[0]), len(b[0])) for s, (i, a), (j, b) in d] == [
        (False, 0, 2, 2), (False, 2, 2, 2), (False, 4, 2, 2), (False, 6, 2, 1)]
    assert len(cnn_gp.DiagIterator(3, X)) == 4


def test_save_k_block_layout(tmp_path):
    """save_K with a fake kernel: which blocks each worker writes (NaN elsewhere), chunking and
    skip-if-present -- against the reference's own loop (tests/golden/save_k_layout.npz)."""
    g = np.load(os.path.join(GOLD, "save_k_layout.npz"))
    meta = json.loads(str(g["meta"]))
    Xs = data.ResidentDataset(torch.from_numpy(g["Xs"]))
    Xt = data.ResidentDataset(torch.from_numpy(g["Xt"]))

    def kern(x, x2, same, diag):  # value = 1000*i + j, enough to check placement
        if diag:
            return np.ones(len(x), np.float32)
        return np.ones((len(x), len(x2)), np.float32)

    for nw in (1, 3):
        for r in range(nw):
            with block_store.open_store(str(tmp_path / f"s{nw}_{r}"), "w") as f:
                cnn_gp.save_K(f, kern, "Kxx", Xs, None, diag=False, batch_size=4, worker_rank=r, n_workers=nw,
                              print_interval=1e9)
                cnn_gp.save_K(f, kern, "Kxtx", Xt, Xs, diag=False, batch_size=4, worker_rank=r, n_workers=nw,
                              print_interval=1e9)
                for name in ("Kxx", "Kxtx"):
                    want = g[f"{name}_nw{nw}_r{r}"]
                    got = f[name][...]
                    assert got.shape == want.shape and got.dtype == np.float32
                    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
                    assert list(f[name].chunks) == meta[f"{name}_nw{nw}_r{r}"]["chunks"]
                    assert list(f[name].maxshape) == meta[f"{name}_nw{nw}_r{r}"]["maxshape"]
                # an existing dataset is skipped, not recomputed
                cnn_gp.save_K(f, lambda *a: 1 / 0, "Kxx", Xs, None, diag=False, batch_size=4)
    with block_store.open_store(str(tmp_path / "s1_0"), "a") as f:
        cnn_gp.save_K(f, kern, "Kt_diag", Xt, None, diag=True, batch_size=4, print_interval=1e9)
        assert f["Kt_diag"].shape == (1, 5) and list(f["Kt_diag"].chunks) == [1, 4]
        A = np.empty((11, 11), np.float32)
        f["Kxx"].read_direct(A, source_sel=np.s_[0, :, :])
        assert np.isnan(A[4, 0]) and A[0, 4] == 1.0

    def bad(x, x2, same, diag):
        k = np.ones((len(x), len(x2)), np.float32)
        k[0, 0] = np.inf
        return k
    with block_store.open_store(str(tmp_path / "bad"), "w") as f:
        with pytest.raises(FloatingPointError):
            cnn_gp.save_K(f, bad, "Kxx", Xs, None, diag=False, batch_size=4, print_interval=1e9)


def test_merge_fills_only_nan(tmp_path):
    a = block_store.open_store(str(tmp_path / "a"), "w")
    b = block_store.open_store(str(tmp_path / "b"), "w")
    da = a.create_dataset("K", shape=(1, 2, 2), dtype=np.float32, fillvalue=np.nan)
    db = b.create_dataset("K", shape=(1, 2, 2), dtype=np.float32, fillvalue=np.nan)
    b.create_dataset("only_b", shape=(1, 2), dtype=np.float32, fillvalue=np.nan)
    da[0, 0, :] = [1, 2]
    db[0, :, :] = [[9, 9], [3, 4]]
    block_store.merge_into(a, b)
    np.testing.assert_array_equal(a["K"][0], [[1, 2], [3, 4]])
    assert "only_b" not in a


def test_fused_translation_of_shipped_programs():
    """Host-side translation into register-level ops (cnngp_plan_describe; no GPU needed): which
    kernel family covers which shipped program, and the structure of the fused-net op lists --
    one ReLU per program ReLU, every Sum an ADD fed by a STASH, strided maps half / quarter size."""
    import re
    models = golden_models()
    expect = {"readme": ("fused_net", 28), "mnist_paper_convnet_gp": ("fused", 28),
              "mnist_paper_residual_cnn_gp": ("fused_net", 28), "mnist_as_tf": ("fused_net", 28),
              "mnist": ("fused_net", 28), "cifar10": ("fused_net", 32)}
    for name, (family, S) in expect.items():
        ops, ns = program.compile_model(models[name])
        plan = nat.Plan(ops, ns, S, S, nat.F32)
        text = plan.describe()
        assert text.split()[0] == family, (name, text[:80])
        n_relu = sum(1 for o in ops if o.opcode == nat.OP_RELU)
        n_add = sum(1 for o in ops if o.opcode == nat.OP_ADD)
        toks = text.split(":", 1)[1].split()
        assert sum(t.startswith(("RELU", "T_RELU")) for t in toks) == n_relu, name
        if family == "fused":
            assert toks[-1] == "DENSE" and n_add == 0
            continue
        assert sum(t.startswith("ADD") for t in toks) == n_add, name
        assert sum(t.startswith("DENSE") for t in toks) == 1, name
        # every ADD reads a tensor-memory slot that an earlier STASH filled and nobody re-filled since
        live = {}
        for t in toks:
            m = re.match(r"(STASH|UNSTASH|ADD)\((\d+),t(\d)\)", t)
            if not m:
                continue
            kind, size, slot = m.group(1), int(m.group(2)), int(m.group(3))
            if kind == "STASH":
                live[slot] = size
            else:
                assert live.get(slot) == size, (name, t)
        sizes = {int(re.match(r"\w+\((\d+)", t).group(1)) for t in toks}
        assert sizes <= {S, S // 2, S // 4, 1}, (name, sizes)
    # float64 and odd input sizes stay on the generic kernel
    ops, ns = program.compile_model(models["mnist_as_tf"])
    assert nat.Plan(ops, ns, 28, 28, nat.F64).describe().startswith("generic")
    ops, ns = program.compile_model(models["edge_linear"])
    assert nat.Plan(ops, ns, 12, 12, nat.F32).describe().startswith("generic")


# ---- the program compiler against the recursive tree walk, on random module trees ----------------
class _ReLUTag:
    pass


_ReLUTag.__name__ = "ReLU"  # the oracle recognises modules by class name


def _run_program(ops, n_slots, X, Z, relu_inputs=None, first=None):
    """Interpret the three-address program (include/cnngp.h) with the oracle's C primitives.
    ``relu_inputs``: list that receives (xx, yy) at every RELU op, in program order; ``first``: list
    that receives the initial xy maps."""
    import ctypes
    from oracle import oracle
    L = oracle.lib()
    N1, N2, (C, H, W) = X.shape[0], Z.shape[0], X.shape[1:]
    xy, xx, yy = np.empty((N1 * N2, H, W)), np.empty((N1, H, W)), np.empty((N2, H, W))
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    L.oracle_init_f64(vp(X), vp(Z), N1, N2, C, H * W, 0, vp(xy), vp(xx), vp(yy))
    slots = [None] * n_slots
    slots[0] = oracle.Patch(False, False, xy, xx, yy, N1, N2)
    if first is not None:
        first.append(xy.copy())

    def conv(a, o):
        M, Hi, Wi = a.shape
        Ho = oracle.conv_out_size(Hi, o.ke, o.stride, o.pad, o.dil)
        Wo = oracle.conv_out_size(Wi, o.ke, o.stride, o.pad, o.dil)
        out = np.empty((M, Ho, Wo))
        L.oracle_conv_f64(vp(np.ascontiguousarray(a)), M, Hi, Wi, o.ke, o.zero_first, o.stride, o.pad, o.dil,
                          o.scale, o.bias, vp(out), Ho, Wo)
        return out
    for o in ops:
        s = slots[o.src]
        assert s is not None, "slot read before it is written"
        if o.opcode == nat.OP_CONV:
            slots[o.dst] = oracle.Patch(False, False, conv(s.xy, o), conv(s.xx, o), conv(s.yy, o), N1, N2)
        elif o.opcode == nat.OP_RELU:
            if relu_inputs is not None:
                relu_inputs.append((s.xx.copy(), s.yy.copy()))
            slots[o.dst] = oracle._propagate(_ReLUTag(), s)
        elif o.opcode == nat.OP_COPY:
            slots[o.dst] = oracle.Patch(False, False, s.xy.copy(), s.xx.copy(), s.yy.copy(), N1, N2)
        elif o.opcode == nat.OP_ADD:
            slots[o.dst] = slots[o.dst] + s
        elif o.opcode == nat.OP_SCALE:
            slots[o.dst] = s * o.scale
        else:
            raise AssertionError(o.opcode)
    return slots[ops[-1].dst].xy.reshape(N1, N2)


def _random_tree(rng, size):
    """A random Sequential over `size` x `size` maps that ends in a 1 x 1 map."""
    def same_conv():
        k = rng.choice([1, 2, 3, 4, 5])
        dil = rng.choice([1, 1, 2]) if size >= 2 * k else 1
        return Conv2d(k, dilation=dil, var_weight=rng.uniform(0.5, 3.0), var_bias=rng.choice([0.0, rng.uniform(0, 1)]))

    def preserving(depth):  # shape-preserving modules only: usable inside Sum / Mixture branches
        kind = rng.choice(["conv", "relu", "seq", "sum", "mix"] if depth < 3 else ["conv", "relu"])
        if kind == "conv":
            return same_conv()
        if kind == "relu":
            return ReLU()
        if kind == "seq":
            return Sequential(*[preserving(depth + 1) for _ in range(rng.randint(0, 3))])
        branches = [preserving(depth + 1) for _ in range(rng.randint(1, 3))]
        if kind == "sum":
            return Sum(branches)
        return Mixture(branches, logit_proportions=torch.tensor([rng.uniform(-1, 1) for _ in branches]))
    mods, cur = [], size
    for _ in range(rng.randint(1, 5)):
        if cur > 3 and rng.random() < 0.3:  # a size-changing conv at the top level
            k, st = rng.choice([2, 3]), rng.choice([1, 2])
            pad = rng.choice([0, 1])
            out = (cur + 2 * pad - (k - 1) - 1) // st + 1
            if out >= 1:
                mods.append(Conv2d(k, stride=st, padding=pad, var_weight=rng.uniform(0.5, 2.0)))
                cur = out
                continue
        mods.append(preserving(0))
    if cur > 1:
        mods.append(Conv2d(cur, padding=0, var_bias=rng.choice([0.0, 0.1])))
    return Sequential(*mods)


def test_compiled_program_equals_tree_walk_on_random_trees():
    """program.compile_model flattens Sequential / Sum / Mixture trees into ops over reusable slots;
    the reference evaluates the tree recursively (kernels.py:184-187, 221-225, 252-254).  Both must
    give the same numbers for arbitrary trees: same C primitives on both sides, float64."""
    import random
    from oracle import oracle
    rng = random.Random(2024)
    g = np.random.default_rng(7)
    worst_slots = 0
    for case in range(60):
        size = rng.choice([6, 9, 10])
        model = _random_tree(rng, size).double()  # Mixture takes its softmax in the parameter dtype (kernels.py:222)
        X, Z = g.random((3, 2, size, size)), g.standard_normal((2, 2, size, size))
        ops, ns = program.compile_model(model)
        worst_slots = max(worst_slots, ns)
        nat.Plan(ops, ns, size, size, nat.F64)  # shape inference and validation agree with the tree
        want = oracle.gram(model, X, Z)
        got = _run_program(ops, ns, X, Z)
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=0, err_msg=f"case {case}: {model}")
    assert worst_slots <= 6


# ---- the fused kernels' host translators, re-evaluated on the CPU --------------------------------
def _box(R, lo, hi, st, dil=1):
    """out[y, x] = sum of in[st*y + dil*dy, st*x + dil*dx], dy, dx in [-lo, hi], zero padding (both fused kernels)."""
    M, H, W = R.shape
    P = np.zeros((M, H + dil * (lo + hi), W + dil * (lo + hi)))
    P[:, dil * lo:dil * lo + H, dil * lo:dil * lo + W] = R
    k = lo + hi + 1
    full = sum(P[:, dil * dy:dil * dy + H, dil * dx:dil * dx + W] for dy in range(k) for dx in range(k))
    return full[:, ::st, ::st]


def _relu2(R, xx, yy, N1, N2):
    """The kernels keep the ReLU output doubled: 2 * arccos kernel of kernels.py:146-152."""
    from oracle import oracle
    kp = oracle.Patch(False, False, np.ascontiguousarray(R), np.ascontiguousarray(xx), np.ascontiguousarray(yy), N1, N2)
    return 2.0 * oracle._propagate(_ReLUTag(), kp).xy


def _parse(dump):
    lines = dump.strip().splitlines()
    out = []
    for ln in lines[1:]:
        kind, *kv = ln.split()
        out.append((kind, {k: float(v) if ("." in v or "e" in v or "inf" in v or "nan" in v) else int(v)
                           for k, v in (f.split("=") for f in kv)}))
    return lines[0], out


def _simulate_translation(plan, xy0, relu_vars, N1, N2):
    """What the fused / fused-net kernel will compute, from cnngp_plan_dump's register-level op list."""
    head, prog = _parse(plan.dump())
    px = [v[0].shape[1] * v[0].shape[2] for v in relu_vars]
    off = np.concatenate([[0], np.cumsum(px)])[:-1].tolist()                        # plain xx offsets (T_RELU)
    foff = np.concatenate([[0], np.cumsum([4 * ((q + 1) // 2) for q in px])])[:-1].tolist()  # fused section (RELU)
    R, T, tot, nth = xy0.copy(), {}, None, 0
    for kind, f in prog:
        if head.startswith("fused S="):
            if kind == "CONV":
                window = f["lo"] or f["hi"]
                if window:
                    R = _box(R, f["lo"], f["hi"], 1) + f["pre_bias"]
                if not window or f["scale"] != 1:
                    R = R * f["scale"] + f["bias"]
            elif kind == "RELU":
                xx, yy = relu_vars[nth]
                nth += 1
                a2 = float(f["aux_scale"]) ** 2
                R = _relu2(R, xx * a2, yy * a2, N1, N2)
            else:
                tot = R.sum(axis=(1, 2)) * f["scale"] + f["bias"]
            continue
        if kind == "CONV":
            # the second sliding sum carries pre_bias (the folded conv bias); scale / bias are the explicit pass
            R = (_box(R, f["lo"], f["hi"], f["st"], f["dil"]) + f["pre_bias"]) * f["scale"] + f["bias"]
            assert R.shape[1] == f["so"]
        elif kind == "AFFINE":
            R = R * f["scale"] + f["bias"]
        elif kind == "RELU":
            xx, yy = relu_vars[foff.index(f["aux"])]
            assert xx.shape[1] == f["si"] and f["half"] == (f["si"] ** 2 + 1) // 2
            a2 = float(f["aux_scale"]) ** 2  # the factor the host puts on this layer's s maps (carried conv taps)
            R = _relu2(R, xx * a2, yy * a2, N1, N2)
        elif kind == "STASH":
            T[f["slot"]] = R.copy()
        elif kind == "UNSTASH":
            R = T[f["slot"]].copy()
        elif kind == "ADD":
            R = R + f["scale"] * T[f["slot"]] + f["bias"]
        elif kind == "TRANSPOSE":
            pass
        elif kind == "DENSE":
            tot = R.sum(axis=(1, 2)) * f["scale"] + f["bias"]
        elif kind == "T_AFFINE":
            tot = tot * f["scale"] + f["bias"]
        elif kind == "T_RELU":
            xx, yy = relu_vars[off.index(f["aux"])]
            tot = _relu2(tot.reshape(-1, 1, 1), xx, yy, N1, N2).reshape(-1)
        else:
            raise AssertionError(kind)
    return tot.reshape(N1, N2)


def _random_straight_line(rng):
    mods, last_relu = [], True
    for _ in range(rng.randint(2, 7)):
        if not last_relu and rng.random() < 0.5:
            mods.append(ReLU())
            last_relu = True
        else:
            mods.append(Conv2d(rng.choice([1, 3, 4, 5, 7]), var_weight=rng.choice([0.01, 0.7, 2.79, 40.0]),
                               var_bias=rng.choice([0.0, 1e-3, 0.5, 7.86])))
            last_relu = False
    if rng.random() < 0.7 and not last_relu:
        mods.append(ReLU())
    mods.append(Conv2d(28, padding=0, var_weight=rng.uniform(0.5, 2), var_bias=rng.choice([0.0, 0.05])))
    return Sequential(*mods)


def _random_resnet(rng, S0):
    def conv(k=None, **kw):
        return Conv2d(k or rng.choice([3, 3, 3, 4, 5, 7]), var_weight=rng.uniform(0.5, 3.0),
                      var_bias=rng.choice([0.0, 0.1, 1.0]), **kw)
    mods, size = [conv()], S0
    for _ in range(rng.randint(1, 5)):
        kind = rng.choice(["identity", "projection", "strided", "sum", "mixture", "relu_conv", "dilated", "nested"])
        if kind == "identity":
            mods.append(resnet_block(stride=1))
        elif kind == "projection":
            mods.append(resnet_block(stride=1, projection_shortcut=True))
        elif kind == "strided" and size > S0 // 4:
            mods.append(resnet_block(stride=2, projection_shortcut=True))
            size //= 2
        elif kind == "sum":
            k = 3 if size < S0 else None
            mods.append(Sum([Sequential(), Sequential(ReLU(), conv(k), ReLU(), conv(k))]))
        elif kind == "dilated" and size == S0:
            mods += [ReLU(), Conv2d(3, dilation=2, var_weight=rng.uniform(0.5, 2.0), var_bias=rng.choice([0.0, 0.3]))]
        elif kind == "nested" and size == S0:
            mods.append(Sum([Sequential(), Sequential(ReLU(), Sum([Sequential(), conv(3)]), ReLU(), conv(3))]))
        elif kind == "mixture":
            k = 3 if size < S0 else None
            mods.append(Mixture([Sequential(), Sequential(ReLU(), conv(k))],
                                logit_proportions=torch.tensor([rng.uniform(-1, 1), rng.uniform(-1, 1)])))
        else:
            mods += [ReLU(), conv(3 if size < S0 else None)]
    if rng.random() < 0.8:
        mods.append(ReLU())
    mods.append(Conv2d(size, padding=0, var_bias=rng.choice([0.0, 0.2])))
    if rng.random() < 0.3:  # a tail on the 1 x 1 map
        mods += [ReLU(), Conv2d(1, var_weight=1.3, var_bias=0.1)]
    return Sequential(*mods)


@pytest.mark.parametrize("family", ["fused", "fused_net_28", "fused_net_32"])
def test_fused_translations_reevaluated_on_cpu(family):
    """The host translators turn a layer program into the register-level op lists the fused kernels
    run (carried conv factors and scaled variance maps in gram_fused.cu; stash / alias / doubled-ReLU
    bookkeeping in gram_fnet.cu).  cnngp_plan_dump exposes those lists; evaluated on the CPU in
    float64 they must reproduce the recursive tree walk of the reference (kernels.py:184-254) on
    random programs -- the only differences are the float32 roundings of the folded constants."""
    import random
    from oracle import oracle
    rng = random.Random({"fused": 1, "fused_net_28": 2, "fused_net_32": 3}[family])
    g = np.random.default_rng(11)
    S0 = 32 if family.endswith("32") else 28
    hits = 0
    for case in range(30 if family == "fused" else 40):
        model = (_random_straight_line(rng) if family == "fused" else _random_resnet(rng, S0)).double()
        ops, ns = program.compile_model(model)
        plan = nat.Plan(ops, ns, S0, S0, nat.F32)
        if plan.fused_kind != (2 if family == "fused" else 3):
            continue  # outside the kernel's set: the generic kernel runs it
        hits += 1
        N1, N2 = 2, 2
        X, Z = g.random((N1, 1, S0, S0)), g.random((N2, 1, S0, S0))
        relu_vars, first = [], []
        want = _run_program(ops, ns, X, Z, relu_vars, first)
        np.testing.assert_allclose(want, oracle.gram(model, X, Z), rtol=1e-12)
        got = _simulate_translation(plan, first[0], relu_vars, N1, N2)
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=0, err_msg=f"{family} case {case}:\n{plan.describe()}")
    assert hits >= 25, hits



# ---- plan cache (engine.plan_for) ------------------------------------------------------------------
def test_plan_cache_follows_hyperparameters_and_survives_copies():
    """The reference reads a layer's hyperparameters on every propagate (kernels.py:92-98); the compiled
    plan is therefore keyed by them, and lives outside the module so that deepcopy / pickle of a model
    that has been called keep working (a plan holds a ctypes handle)."""
    import copy
    import pickle
    from cnn_gp import engine
    model = Sequential(Conv2d(3, var_bias=0.5), ReLU(), Conv2d(28, padding=0))
    p1 = engine.plan_for(model, 28, 28, torch.float32)
    assert engine.plan_for(model, 28, 28, torch.float32) is p1
    model.mods[0].var_bias = 2.0
    p2 = engine.plan_for(model, 28, 28, torch.float32)
    assert p2 is not p1
    assert "2" in p2.dump() and p2.dump() != p1.dump()
    clone = copy.deepcopy(model)
    assert pickle.loads(pickle.dumps(model)).mods[0].var_bias == 2.0
    assert engine.plan_for(clone, 28, 28, torch.float32) is not p2
    assert not any(k.startswith("_cnngp") for k in model.__dict__)
