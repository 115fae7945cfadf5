"""Batch sharding of the op path over the GPUs of one node.

Every op is independent per image, so the only exchange is the scalar loss reduction of the caller
(model/networks.py:377, `(mask*diff).sum() / mask.sum()`): one packed all-reduce of numerators and
denominators.  One process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world_size):
    """Contiguous block of `n_items` owned by `rank`: [lo, hi).  The first n_items % world_size ranks
    get one extra item, so any batch size shards."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank/world_size %r/%r" % (rank, world_size))
    base, extra = divmod(int(n_items), world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class ShardedLoss:
    """Accumulates (numerator, denominator) pairs of masked means computed on this rank's shard and
    reduces them all with ONE all-reduce, so N losses cost one latency, not N."""

    def __init__(self, device, group=None):
        self.device = device
        self.group = group
        self._pairs = []

    def add(self, numerator, denominator):
        self._pairs.append((numerator.reshape(()).to(torch.float32), denominator.reshape(()).to(torch.float32)))
        return len(self._pairs) - 1

    def pack(self):
        if not self._pairs:
            return torch.zeros(0, 2, device=self.device)
        return torch.stack([torch.stack(p) for p in self._pairs])

    def reduce(self, async_op=False):
        """Returns (values, work): values[i] = sum_ranks(num_i) / sum_ranks(den_i)."""
        packed = self.pack()
        work = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            work = dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        self._packed = packed
        if async_op and work is not None:
            return None, work
        return packed[:, 0] / packed[:, 1], None

    def result(self):
        return self._packed[:, 0] / self._packed[:, 1]
