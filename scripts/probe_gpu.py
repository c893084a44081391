"""Pipe-throughput probes + a quick throughput reading of the Gram kernels (run on the GPU box)."""
import ctypes
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cnn-gp_b200"), ROOT]
import torch  # noqa: E402
from cnn_gp import engine  # noqa: E402


def probes():
    L = ctypes.CDLL(os.path.join(ROOT, "cnn-gp_b200", "libcnngp_bench.so"))
    L.mb_probe.restype = ctypes.c_double
    L.mb_probe.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
    names = {0: "ffma", 1: "ffma2(x2)", 2: "mufu.sqrt", 3: "mufu.rcp", 4: "7ffma+1mufu", 5: "fadd",
             6: "6ffma+2fmnmx", 10: "lds32", 11: "lds128(words)", 20: "dmma(fma-equiv)", 21: "dfma"}
    out = {}
    for k, n in names.items():
        for bps in (4, 8):
            v = L.mb_probe(k, bps, 4000)
            out[f"{n}@{bps}cta"] = v / 1e12
    return out


def gram_rate(cfg, n, path, C=1, S=28, reps=3):
    model = importlib.import_module("configs." + cfg).initial_model.cuda()
    X = torch.rand(n, C, S, S, device="cuda")
    prev = engine.set_path(path)
    try:
        model(X[:64])
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            model(X)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        return n * (n + 1) / 2 / best, engine.last_path()
    finally:
        engine.set_path(prev)


def relu_probes():
    L = ctypes.CDLL(os.path.join(ROOT, "cnn-gp_b200", "libcnngp_bench.so"))
    L.mb_relu.restype = ctypes.c_double
    L.mb_relu.argtypes = [ctypes.c_int, ctypes.c_int]
    return {f"relu_v{v}_Gpp_per_s": L.mb_relu(v, 2000) / 1e9 for v in (0, 1)}


if __name__ == "__main__":
    if sys.argv[1:] == ["relu"]:
        print(json.dumps(relu_probes(), indent=1))
        sys.exit(0)
    res = {"probes_Tops": probes()}
    for cfg, C, S in (("mnist_paper_convnet_gp", 1, 28), ("mnist_as_tf", 1, 28), ("cifar10", 3, 32)):
        for path in sys.argv[1:] or ["generic"]:
            try:
                r, used = gram_rate(cfg, 1500, path, C, S)
                res[f"{cfg}:{path}"] = {"pairs_per_s": r, "path": used}
            except Exception as e:  # noqa: BLE001
                res[f"{cfg}:{path}"] = {"error": str(e)}
    print(json.dumps(res, indent=1))
