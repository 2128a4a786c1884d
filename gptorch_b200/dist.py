"""Host-side helpers for one-process-per-GPU runs (torch.distributed, NCCL on the B200 box, gloo in CPU tests).

Only two things shard in this path: the rows of X / Y for the sparse models, and replicas of an
independent evaluation.  The collectives themselves live where the exchange happens (VfeStatsFn for VFE);
these helpers cover row partitioning, gradient all-reduce for data-parallel SVGP and max-over-ranks timing.
"""
import torch
import torch.distributed as dist


def shard_rows(n, rank, world):
    """Contiguous, balanced row range [start, stop) of rank `rank` out of `world` for n rows."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def allreduce_grads(model, group=None):
    """Sum the gradients of every trainable parameter over the ranks as ONE flat buffer (SVGP data parallel:
    M*D + M*dy + M^2 + D + 2 doubles).  Parameters without a gradient contribute zeros."""
    params = [p for p in model.parameters() if p.requires_grad]
    if not params:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, group=group)
    offset = 0
    for p in params:
        chunk = flat[offset: offset + p.numel()].reshape(p.shape)
        if p.grad is None:
            p.grad = chunk.clone()
        else:
            p.grad.copy_(chunk)
        offset += p.numel()


def allreduce_scalar_sum(value, device, group=None):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, group=group)
    return t.item()


def max_over_ranks(value, device, group=None):
    """Device-side max (multi-GPU timings are reported as the slowest rank)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t.item()
