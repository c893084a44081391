"""Distributed fp64 Cholesky solve under torchrun: correctness against the single-GPU solver and timing.
usage: python -m torch.distributed.run --nproc-per-node N scripts/check_dist_solve.py n [n ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cnn-gp_b200"), ROOT]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from cnn_gp import linalg, linalg_dist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
for n in [int(a) for a in sys.argv[1:]] or [8192]:
    K = Y = None
    if rank == 0:
        g = torch.Generator(device="cuda").manual_seed(n)
        B = torch.randn(n, 64, generator=g, device=dev, dtype=torch.float64)
        K = B @ B.T
        K.diagonal().add_(1.0)
        K = torch.triu(K) + torch.tril(torch.full_like(K, float("nan")), -1)
        Y = torch.randn(n, 10, generator=g, device=dev, dtype=torch.float64)
    out = {"n": n, "world": world}
    for la in (True, False):
        ch = linalg_dist.DistributedCholesky(n, dev)
        ch.scatter_from(K)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        info = ch.factorize(lookahead=la)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        assert info == 0, info
        out["potrf_ms_lookahead" if la else "potrf_ms_plain"] = float(ms)
        out["tflops_lookahead" if la else "tflops_plain"] = n ** 3 / 3 / (float(ms) * 1e-3) / 1e12
        if la:
            # the two sweeps on the distributed factor (no rank holds all of U), timed, then the factor for the check
            ch.solve(Y)  # first call: communicator warm-up
            torch.cuda.synchronize()
            e0.record()
            Xd = ch.solve(Y)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            out["potrs_distributed_ms"] = float(ms)
            U = ch.gather_to(0)
        del ch
    if rank == 0:
        X = linalg.potrs_upper_(U, Y.clone())
        out["distributed_vs_one_gpu_solution"] = float((Xd - X).abs().max() / X.abs().max())
        if n <= 40000:  # the symmetrised copy costs 3 more n x n buffers
            Kf = torch.triu(K) + torch.triu(K, 1).T
            out["residual"] = float((Kf @ X - Y).abs().max() / (Kf.abs().max() * X.abs().max()))
            del Kf
        U1 = K  # the single-GPU factorisation may overwrite the input now
        K = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        linalg.potrf_upper_(U1)
        e1.record()
        torch.cuda.synchronize()
        out["single_gpu_potrf_ms"] = e0.elapsed_time(e1)
        U.sub_(U1)
        del U1
        out["max_diff_vs_single_gpu_factor"] = float(torch.triu(U).abs().max())
        print(json.dumps(out))
        del U, X
    dist.barrier()
dist.destroy_process_group()
