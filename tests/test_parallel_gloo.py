"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes exercise the env sharding and
the one-collective gradient / moment reductions used by env-sharded SHAC."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffrl_b200.parallel import allreduce_gradients, allreduce_moments, shard_envs
    first, count = shard_envs(10)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Tanh(), torch.nn.Linear(3, 2))
    x = torch.arange(40, dtype=torch.float32).view(10, 4) / 10.0
    loss = net(x[first:first + count]).pow(2).sum() / 10.0          # loss normalised by the GLOBAL batch
    loss.backward()
    n = allreduce_gradients(list(net.parameters()), average=False)
    local = x[first:first + count]
    c, m, v = allreduce_moments(count, local.mean(0), local.var(0, unbiased=False))
    if rank == 0:
        torch.save({"grads": [p.grad.clone() for p in net.parameters()], "n": n, "shard": (first, count),
                    "moments": (c, m, v)}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gradients_equal_single_process(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Tanh(), torch.nn.Linear(3, 2))
    x = torch.arange(40, dtype=torch.float32).view(10, 4) / 10.0
    (net(x).pow(2).sum() / 10.0).backward()
    for g, p in zip(got["grads"], net.parameters()):
        assert torch.allclose(g, p.grad, rtol=1e-5, atol=1e-6)
    assert got["n"] == sum(p.numel() for p in net.parameters()) and got["shard"] == (0, 5)
    c, m, v = got["moments"]
    assert c == 10 and torch.allclose(m, x.mean(0), atol=1e-6) and torch.allclose(v, x.var(0, unbiased=False), atol=1e-5)
