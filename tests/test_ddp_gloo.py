"""CPU, world_size 2, gloo: the bucketed/overlapped gradient all-reduce of lav_b200.train averages gradients so that
two ranks on half batches reproduce the single-process full-batch step (the N>1 path of SURVEY §8e)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn


def _model():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(12, 64), nn.ReLU(), nn.Linear(64, 64), nn.ReLU(), nn.Linear(64, 3))


def _data():
    g = torch.Generator().manual_seed(1)
    return torch.randn(16, 12, generator=g), torch.randn(16, 3, generator=g)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lav_b200.train import GradAllReducer
    m = _model()
    red = GradAllReducer(m.parameters(), bucket_bytes=4096)      # several buckets
    assert len(red.buckets) > 1
    x, y = _data()
    n = x.shape[0] // world
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    for step in range(3):
        xs, ys = x[rank * n:(rank + 1) * n], y[rank * n:(rank + 1) * n]
        opt.zero_grad(set_to_none=True)
        ((m(xs) - ys) ** 2).mean().backward()
        red.finish()
        opt.step()
    if rank == 0:
        torch.save([p.detach().clone() for p in m.parameters()], out)
    dist.destroy_process_group()


def test_two_rank_allreduce_matches_full_batch(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "params.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    m = _model()
    x, y = _data()
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    for step in range(3):
        opt.zero_grad(set_to_none=True)
        ((m(x) - y) ** 2).mean().backward()          # mean over the full batch == mean of the two half-batch means
        opt.step()
    for a, b in zip(got, m.parameters()):
        assert torch.allclose(a, b.detach(), rtol=1e-5, atol=1e-6)


def test_single_process_is_a_noop():
    from lav_b200.train import GradAllReducer
    m = _model()
    red = GradAllReducer(m.parameters())
    x, y = _data()
    ((m(x) - y) ** 2).mean().backward()
    g0 = [p.grad.clone() for p in m.parameters()]
    red.finish()
    assert all(torch.equal(a, p.grad) for a, p in zip(g0, m.parameters()))


class _TwoBranch(nn.Module):
    """a module whose second branch only runs when asked: on the rank that skips it those parameters get NO gradient"""

    def __init__(self):
        super().__init__()
        torch.manual_seed(3)
        self.a, self.b, self.out = nn.Linear(12, 32), nn.Linear(12, 32), nn.Linear(32, 3)

    def forward(self, x, use_b):
        h = torch.relu(self.a(x))
        if use_b:
            h = h + torch.relu(self.b(x))
        return self.out(h)


def _worker_uneven(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lav_b200.train import GradAllReducer
    m = _TwoBranch()
    red = GradAllReducer(m.parameters(), bucket_bytes=512)       # one bucket per parameter tensor or so
    x, y = _data()
    n = x.shape[0] // world
    xs, ys = x[rank * n:(rank + 1) * n], y[rank * n:(rank + 1) * n]
    # rank 1 never touches branch b: its buckets complete in a different order (and some never complete) — the reducer must
    # still issue the same collectives in the same order on both ranks (no hang, correct average)
    ((m(xs, use_b=(rank == 0)) - ys) ** 2).mean().backward()
    red.finish()
    if rank == 0:
        torch.save({k: p.grad.detach().clone() for k, p in m.named_parameters()}, out)
    dist.destroy_process_group()


def test_rank_without_gradient_for_some_parameters(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "grads.pt")
    mp.spawn(_worker_uneven, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    x, y = _data()
    want = {}
    for rank in range(2):
        m = _TwoBranch()
        xs, ys = x[rank * 8:(rank + 1) * 8], y[rank * 8:(rank + 1) * 8]
        ((m(xs, use_b=(rank == 0)) - ys) ** 2).mean().backward()
        for k, p in m.named_parameters():
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            want[k] = want.get(k, 0) + g / 2
    for k in want:
        assert torch.allclose(got[k], want[k], rtol=1e-5, atol=1e-6), k
