"""Where the end-to-end call model(x_host) spends its time next to the device-resident step:
python scripts/e2e_breakdown.py [N]   (wall clock around synchronised phases, best of 5)"""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cnn-gp_b200"), ROOT]
import torch  # noqa: E402
from cnn_gp import engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
model = importlib.import_module("configs.mnist_paper_convnet_gp").initial_model.cuda()
Xh = torch.rand(n, 1, 28, 28).pin_memory()
out = torch.empty((n, n), dtype=torch.float32).pin_memory()


def best(fn, reps=5):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return round(min(ts), 3)


X = Xh.cuda()
plan = engine.plan_for(model, 28, 28, torch.float32)
aux, _, kd = engine.variances(plan, X)
K = torch.empty((n, n), device="cuda")
res = {"n": n}
res["upload_ms"] = best(lambda: Xh.to("cuda", non_blocking=True))
res["variances_ms"] = best(lambda: engine.variances(plan, X))
res["gram_device_ms"] = best(lambda: engine.gram_with_aux(plan, X, X, aux, aux, True, False, True, out=K, kdiag=kd))
res["d2h_last_band_ms"] = best(lambda: out[-424:].copy_(K[-424:], non_blocking=True))
res["d2h_all_ms"] = best(lambda: out.copy_(K, non_blocking=True))
model(Xh); model(Xh)
res["model_x_host_ms"] = best(lambda: model(Xh))
res["gram_host_out_given_ms"] = best(lambda: engine.gram_host(model, Xh, out=out))
print(res)

# the streamed call taken apart: the launch with progress counters timed by events on the compute stream,
# the copy stream's last band by an event on the copy stream
import ctypes  # noqa: E402
from cnn_gp import _native as nat  # noqa: E402
side = torch.cuda.Stream()
scratch = torch.empty(max(64, n // 32 + 8), dtype=torch.int32, device="cuda")
for _ in range(4):
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t0 = time.perf_counter()
    e0.record()
    rc = nat.lib().cnngp_gram_symmetric_to_host(
        plan.handle, X.data_ptr(), n, 1, aux.data_ptr(), kd.data_ptr(), K.data_ptr(), K.stride(0),
        out.data_ptr(), out.stride(0), scratch.data_ptr(), scratch.numel() * 4,
        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.c_void_p(side.cuda_stream))
    t1 = time.perf_counter()
    e1.record()
    e2.record(side)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print({"rc": rc, "enqueue_ms": round((t1 - t0) * 1e3, 3), "kernel_with_progress_ms": round(e0.elapsed_time(e1), 3),
           "copy_stream_done_after_kernel_start_ms": round(e0.elapsed_time(e2), 3), "wall_ms": round((t2 - t0) * 1e3, 3)})
