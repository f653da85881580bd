"""Worker of tests/test_host_cpu.py::test_data_parallel_reduce_reproduces_global_masked_mean (one process per rank,
gloo backend): python dp_worker.py RANK WORLD PORT OUTFILE"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import loss_oracle as lo  # noqa: E402


def main(rank, world, port, outfile):
    import torch.distributed as dist
    from rift_b200.trainer import allreduce_grads_and_stats
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # two shards of one batch with unequal valid counts: the sharded (sum, count) + unscaled gradients must
        # reproduce the global masked mean of rift_trainer.py:173-178 exactly
        gen = torch.Generator().manual_seed(0)
        bs, R, Mo = 8, 3, 12
        z = torch.randn(bs, R, Mo, generator=gen)
        old = z + 0.3 * torch.randn(bs, R, Mo, generator=gen)
        adv = torch.randn(bs, R, Mo, generator=gen, dtype=torch.float64)
        nr = torch.tensor([3, 1, 2, 3, 1, 1, 2, 3])
        vm = (torch.arange(R)[None, :] < nr[:, None])[..., None].expand(bs, R, Mo).contiguous()
        r_pad = ~vm.any(-1)
        sl = slice(0, 5) if rank == 0 else slice(5, 8)             # uneven shards
        zs = z[sl].clone().requires_grad_(True)
        loss_local = lo.rift_loss(zs, old[sl], adv[sl] * vm[sl], vm[sl], r_pad[sl])
        cnt = float(vm[sl].sum())
        (loss_local * cnt).backward()                              # unscaled gradient = d(-sum obj)/dz
        n = zs.numel()
        n_max = 5 * R * Mo
        grads = torch.zeros(n_max + 64)
        grads[:n] = zs.grad.flatten()
        stats = torch.tensor([-float(loss_local) * cnt, cnt], dtype=torch.float64)
        s = allreduce_grads_and_stats(grads, n_max, stats)
        torch.save((rank, float(-s[0] / s[1]), float(s[1]), grads[:n_max].clone()), outfile)
    finally:
        dist.destroy_process_group()




if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
