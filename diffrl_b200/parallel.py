"""Data-parallel glue for env-sharded training: environments are independent, so
the simulation kernels never communicate; the only exchange is the per-rollout all-reduce of the policy
gradient (and, for exact equivalence, of critic gradients and observation-normaliser moments).
One process per GPU, ``torch.distributed`` (NCCL on GPUs; gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_envs(total_envs, rank=None, world_size=None):
    """Contiguous partition of ``total_envs`` environments; returns (first_env, count) for this rank."""
    rank = dist.get_rank() if rank is None else rank
    world_size = dist.get_world_size() if world_size is None else world_size
    base, extra = divmod(int(total_envs), world_size)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def allreduce_gradients(params, average=True, group=None):
    """Sum (or average) the ``.grad`` of ``params`` over ranks in ONE collective on a flat buffer.
    Call after ``loss.backward()`` and before gradient clipping (reference algorithms/shac.py:411-417)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not (dist.is_available() and dist.is_initialized()):
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    offset = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[offset:offset + n].view_as(g))
        offset += n
    return int(flat.numel())


def allreduce_moments(count, mean, var, group=None):
    """Combine per-rank batch moments (count, mean, var) into global ones (parallel-variance formula), for
    the running observation normaliser (reference utils/running_mean_std.py:32-36)."""
    if not (dist.is_available() and dist.is_initialized()):
        return count, mean, var
    n = torch.as_tensor(float(count), device=mean.device, dtype=mean.dtype)
    packed = torch.cat([n.reshape(1), (mean * n).reshape(-1), ((var + mean * mean) * n).reshape(-1)])
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    total = packed[0]
    k = mean.numel()
    g_mean = packed[1:1 + k].view_as(mean) / total
    g_var = packed[1 + k:1 + 2 * k].view_as(var) / total - g_mean * g_mean
    return float(total), g_mean, g_var
