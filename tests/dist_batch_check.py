"""torchrun --nproc-per-node N tests/dist_batch_check.py : qp_batch_distributed over NCCL on real GPUs,
checked against a single-GPU solve of the same batch on rank 0."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
from problems import dense_qp


def main():
    rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    import cvxopt_b200
    B, n, m = 11, 96, 200
    args = (None,) * 4
    if rank == 0:
        Ps, qs, Gs, hs = zip(*[dense_qp(n, m, seed=k) for k in range(B)])
        args = (np.stack(Ps), np.stack(qs), np.stack(Gs), np.stack(hs))
    res = cvxopt_b200.qp_batch_distributed(*args)
    if rank == 0:
        single = cvxopt_b200.qp_batch(*args, device=lr)
        full = res["all"]
        assert list(full["iterations"]) == list(single["iterations"]), (full["iterations"], single["iterations"])
        assert np.allclose(full["x"], single["x"], rtol=1e-10, atol=1e-12)
        assert all(s == "optimal" for s in full["status"])
        print("dist_batch_check OK: world %d, %d problems, shard0 %d, iterations %s" % (
            dist.get_world_size(), B, res["x"].shape[0], list(full["iterations"])))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
