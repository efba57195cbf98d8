"""Multi-GPU check (run under torchrun): distributed two-level Cholesky == single-GPU sa_potrf, and its timing.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_dist_potrf.py [M ...]
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sobolev_alignment_b200 import _native as nat, distributed  # noqa: E402
from sobolev_alignment_b200.engine import CudaOps  # noqa: E402
from sobolev_alignment_b200.solver import Comm  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
ops, comm = CudaOps(), Comm(None, enabled=True)
for M in [int(x) for x in sys.argv[1:]] or [8192, 50048]:
    g = torch.Generator(device=dev).manual_seed(1)
    Z = torch.randn(M, 64, device=dev, generator=g)
    K = torch.cdist(Z, Z).mul_(-1.0 / 8.0).exp_()
    K.diagonal().add_(1e-5 * M)
    del Z
    nblk = M // 128
    ref, inv_ref = K.clone(), torch.empty(nblk, 128, 128, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    nat.potrf(ref, inv_ref, info)
    torch.cuda.synchronize()
    t_single = []
    for _ in range(2):
        A = K.clone()
        torch.cuda.synchronize(); dist.barrier(); t = time.perf_counter()
        nat.potrf(A, inv_ref, info)
        torch.cuda.synchronize(); t_single.append(time.perf_counter() - t)
    t_dist = []
    for _ in range(3):
        A, inv = K.clone(), torch.empty(nblk, 128, 128, device=dev)
        torch.cuda.synchronize(); dist.barrier(); t = time.perf_counter()
        distributed.potrf(ops, A, inv, info, comm)
        torch.cuda.synchronize(); dist.barrier(); t_dist.append(time.perf_counter() - t)
    L, Lr = torch.tril(A), torch.tril(ref)
    err = float((L - Lr).abs().max() / Lr.abs().max())
    erri = float((inv - inv_ref).abs().max() / inv_ref.abs().max())
    # every rank must hold the same bits
    chk = torch.stack([L.double().sum(), inv.double().sum()])
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"M={M} world={world}: distributed potrf {min(t_dist) * 1e3:.1f} ms (single-GPU {min(t_single) * 1e3:.1f} ms), "
              f"max rel diff L {err:.2e} inv {erri:.2e}, info {int(info)}, identical across ranks: {bool((lo == hi).all())}",
              flush=True)
    del A, K, ref, L, Lr
    torch.cuda.empty_cache()
dist.destroy_process_group()
