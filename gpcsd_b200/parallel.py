"""Trial sharding over torch.distributed (one process per GPU, NCCL over NVLink; gloo for CPU tests).

Trials are i.i.d. given the hyperparameters (the sum over ``trial`` at gpcsd1d.py:124-126), so the LFP is
partitioned into contiguous trial slabs, the small factors are replicated, and one all-reduce of the
partial (loglik, gradient) vector -- P+1 doubles -- closes each evaluation.  ``predict`` has no exchange.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(ntrials, rank, world):
    """Contiguous, balanced [lo, hi) slab of trials for `rank` (first ntrials % world ranks get one more)."""
    base, rem = divmod(int(ntrials), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class TrialShard:
    """group: None/False -> single process (no sharding); True -> the default process group;
    otherwise a torch.distributed process group."""

    def __init__(self, group=None):
        self.enabled = group is not None and group is not False
        if self.enabled and not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("trial sharding requested but torch.distributed is not initialised")
        self.group = None if (group is True or not self.enabled) else group
        self.rank = dist.get_rank(self.group) if self.enabled else 0
        self.world = dist.get_world_size(self.group) if self.enabled else 1

    def bounds(self, ntrials):
        return shard_bounds(ntrials, self.rank, self.world)

    def det_fraction(self):
        """Share of the trial-independent (log-det) terms each rank contributes, so that the SUM over ranks
        counts them exactly once."""
        return 1.0 / self.world

    def allreduce_sum(self, vec, device=None):
        """Sum a small float64 host vector over the ranks (no-op for world == 1)."""
        vec = np.asarray(vec, dtype=np.float64)
        if not self.enabled or self.world == 1:
            return vec
        backend = dist.get_backend(self.group)
        t = torch.from_numpy(vec.copy())
        if backend == "nccl":
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()


class RestartShard:
    """Multi-start restarts are independent optimisations of the same model: shard restart indices round-robin
    over the ranks (every rank holds the full, small LFP), and exchange (nll, params, message) once at the end
    -- the 'allgather of n_restarts scalars' of SURVEY.md 8(e).  group: None/False -> no sharding."""

    def __init__(self, group=None):
        self.enabled = group is not None and group is not False
        if self.enabled and not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("restart sharding requested but torch.distributed is not initialised")
        self.group = None if (group is True or not self.enabled) else group
        self.rank = dist.get_rank(self.group) if self.enabled else 0
        self.world = dist.get_world_size(self.group) if self.enabled else 1

    def mine(self, restart_index):
        return restart_index % self.world == self.rank

    def gather(self, local):
        """local: {restart_index: result}; returns the merged dict (identical on every rank)."""
        if not self.enabled or self.world == 1:
            return dict(local)
        parts = [None] * self.world
        dist.all_gather_object(parts, local, group=self.group)
        merged = {}
        for p in parts:
            merged.update(p)
        return merged
