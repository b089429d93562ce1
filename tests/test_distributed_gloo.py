"""CPU, world_size 2 over gloo: bucketed gradient all-reduce of flowtron_b200.distributed equals the
average of per-rank gradients (reference semantics: distributed.py:113-120 sum then /world_size)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


class Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.flows = torch.nn.ModuleList([torch.nn.Linear(4, 4), torch.nn.Linear(4, 4)])
        self.embedding = torch.nn.Embedding(5, 4)

    def forward(self, ids):
        x = self.embedding(ids)
        for f in self.flows:
            x = torch.tanh(f(x))
        return x.sum()


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    from flowtron_b200 import distributed as ftd
    ftd.init_distributed(rank, world, "gloo")
    torch.manual_seed(100 + rank)          # different init per rank: broadcast must make them equal
    m = Tiny()
    ftd.apply_gradient_allreduce(m)
    w0 = m.flows[0].weight.detach().clone()
    for it in range(2):                     # two iterations: buckets re-arm
        m.zero_grad_buckets()
        ids = torch.tensor([[rank, (rank + it + 1) % 5, 4]])
        m(ids).backward()
    torch.save({"w0": w0, "grads": [p.grad.clone() for p in m.parameters()]}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_allreduce_world2(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert torch.equal(r0["w0"], r1["w0"])                      # rank-0 broadcast
    for a, b in zip(r0["grads"], r1["grads"]):
        assert torch.allclose(a, b, atol=1e-7)                   # identical averaged grads on both ranks
    # recompute the expected average on one process
    torch.manual_seed(100)
    m = Tiny()
    expected = [torch.zeros_like(p) for p in m.parameters()]
    for rank in range(2):
        m.zero_grad()
        m(torch.tensor([[rank, (rank + 1 + 1) % 5, 4]])).backward()
        for e, p in zip(expected, m.parameters()):
            e += (p.grad if p.grad is not None else torch.zeros_like(p)) / 2
    for e, g in zip(expected, r0["grads"]):
        assert torch.allclose(e, g, atol=1e-6)
