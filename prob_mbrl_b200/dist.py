"""Particle sharding across GPUs (one process per GPU, torch.distributed over NCCL).

The reference has no distributed code; the particle axis of a no-moment-matching
rollout is embarrassingly parallel, so the N particles are split contiguously across ranks and the only
exchange is ONE all-reduce (sum) of the flat policy gradient per iteration, before clipping so that the
clip + Adam update is identical on every rank.

To make results independent of the world size every rank keeps drawing the FULL-N noise buffers from
identically seeded generators (exactly what a single process would draw) and then narrows each buffer
to its own rows; ``ShardedNoise`` does that narrowing/widening on the modules' own buffers so the same
code serves the reference's modules, this package's mirror modules, the fused and the eager backend.
"""
import torch


def world():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank(), torch.distributed.get_world_size()
    return 0, 1


def shard_rows(n_global, rank, world_size):
    if n_global % world_size != 0:
        raise ValueError("particle count %d is not divisible by the world size %d" % (n_global, world_size))
    n = n_global // world_size
    return rank * n, n


def _noise_slots(dynamics, policy):
    """(module, attribute) pairs of every per-particle noise buffer on the rollout path."""
    slots = []
    for top in (policy.model, dynamics.model):
        for m in top.modules():
            if hasattr(m, "concrete_noise") and hasattr(m, "logit_p"):
                slots += [(m, "noise"), (m, "concrete_noise")]
            elif hasattr(m, "noise") and hasattr(m, "rate"):
                slots.append((m, "noise"))
            elif hasattr(m, "max_log_std") and hasattr(m, "z"):
                slots.append((m, "z"))
    od = getattr(dynamics, "output_density", None)
    if od is not None and hasattr(od, "z"):
        slots.append((od, "z"))
    return slots


class ShardedNoise:
    """Narrow the modules' [N_global, .] noise buffers to this rank's rows, and back."""

    def __init__(self, dynamics, policy, n_global, rank, world_size):
        self.row0, self.n = shard_rows(n_global, rank, world_size)
        self.n_global = n_global
        self.slots = _noise_slots(dynamics, policy)
        self.full = None

    def narrow(self):
        if self.full is not None:
            return
        self.full = []
        for m, name in self.slots:
            t = getattr(m, name)
            whole = t.detach()                  # new tensor object aliasing the full storage
            self.full.append(whole)
            if t.dim() == 2 and t.shape[0] >= self.n_global:
                t.data = whole[self.row0:self.row0 + self.n]

    def widen(self):
        if self.full is None:
            return
        for (m, name), whole in zip(self.slots, self.full):
            getattr(m, name).data = whole
        self.full = None


def allreduce_gradient(flat, loss=None):
    """The single collective of an iteration: sum of the per-rank gradient contributions."""
    torch.distributed.all_reduce(flat)
    if loss is not None:
        torch.distributed.all_reduce(loss)
