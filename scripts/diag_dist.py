import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cnn-gp_b200"), ROOT]
import torch, torch.distributed as dist
from cnn_gp import linalg_dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1])
out = {}
buf = torch.empty(256 * n, dtype=torch.float64, device=dev)
for _ in range(3): dist.broadcast(buf, src=0)
torch.cuda.synchronize(); t0 = time.perf_counter()
for k in range(10): dist.broadcast(buf, src=k % world)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
out["bcast_MB"] = buf.numel() * 8 / 1e6; out["bcast_ms"] = dt * 1e3; out["bcast_GBs"] = buf.numel() * 8 / dt / 1e9
ch = linalg_dist.DistributedCholesky(n, dev)
ch.local.normal_()
be = ch.backend
X = buf[:256 * (n - 256)].view(256, n - 256)
blocks = [i for i in ch.blocks if i > 0]
def step():
    for i in blocks:
        be.syrk(X, 256, n - 256, ch.rows_of(i), 256, (i - 1) * 256)
step(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record(); step(); e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
out["step0_update_gpu_ms"] = e0.elapsed_time(e1); out["step0_update_host_ms"] = (t1 - t0) * 1e3; out["launches"] = len(blocks)
flops = sum(2.0 * 256 * 256 * (n - 256 - (i - 1) * 256 - 128) for i in blocks)
out["step0_tflops"] = flops / (e0.elapsed_time(e1) * 1e-3) / 1e12
info = torch.zeros(1, dtype=torch.int32, device=dev)
ch.local.zero_(); rows = ch.rows_of(ch.blocks[0]); rows[:, ch.blocks[0]*256:ch.blocks[0]*256+256] = torch.eye(256, device=dev, dtype=torch.float64) * 4
be.panel(rows, ch.blocks[0] * 256, n, info); torch.cuda.synchronize()
rows[:, ch.blocks[0]*256:ch.blocks[0]*256+256] = torch.eye(256, device=dev, dtype=torch.float64) * 4
e0.record(); be.panel(rows, ch.blocks[0] * 256, n, info); e1.record(); torch.cuda.synchronize()
out["panel_ms"] = e0.elapsed_time(e1)
if rank == 0: print(json.dumps(out))
dist.destroy_process_group()
