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


class CollectiveOrder:
    """Turnstile for models that are evaluated from different host threads of one process, each on its own communicator.

    Collectives of DIFFERENT communicators on one device still have to be issued in the same order on every rank (NCCL
    kernels of one device are not guaranteed to run concurrently): two free-running threads would issue
    (comm0, comm1) on one rank and (comm1, comm0) on another and dead-lock.  `turn(i)` lets member i enqueue its next
    collective only when it is its turn in the fixed rotation 0, 1, ..., n-1, 0, ... -- every member must therefore issue
    the same sequence of collectives between start() and stop() (true for models evaluated in lock step, as in bench.py).
    Only the ENQUEUE is inside the turn; waiting for the result happens outside."""

    def __init__(self, n):
        import threading
        self.n = int(n)
        self.seq = 0
        self.enabled = False
        self.cv = threading.Condition()

    def start(self):
        """Begin a phase in which the members run on concurrent threads (call from the coordinating thread while no
        member is evaluating).  Outside such phases -- members driven one after the other by a single thread, whose order is
        the same on every rank anyway -- turn() does not wait."""
        with self.cv:
            self.seq, self.enabled = 0, True

    def stop(self):
        with self.cv:
            self.enabled = False
            self.cv.notify_all()

    def turn(self, idx):
        return _Turn(self, int(idx)) if self.enabled else _NoTurn()


class _Turn:
    def __init__(self, order, idx):
        self.order, self.idx = order, idx

    def __enter__(self):
        o = self.order
        with o.cv:
            while o.seq % o.n != self.idx:
                o.cv.wait()

    def __exit__(self, *exc):
        o = self.order
        with o.cv:
            o.seq += 1
            o.cv.notify_all()
        return False


class _NoTurn:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


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
        self.order, self.order_index = None, 0

    def set_order(self, order, index):
        """Join a CollectiveOrder rotation (models driven from several host threads of one process)."""
        self.order, self.order_index = order, int(index)

    def _turn(self):
        return self.order.turn(self.order_index) if self.order is not None else _NoTurn()

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
        with self._turn():
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def allreduce_inplace(self, t):
        """Sum a small float64 DEVICE tensor over the ranks in place, on the caller's current stream (NCCL); the result
        stays on the device (the native plan reads it back itself).  No-op for a single rank."""
        if self.enabled and self.world > 1:
            if dist.get_backend(self.group) != "nccl":
                raise RuntimeError("device all-reduce needs the NCCL backend")
            with self._turn():
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def allreduce_device(self, t):
        """Sum a small float64 DEVICE vector over the ranks and return it on the host: with NCCL the all-reduce runs in
        place on the caller's stream (no host round trip before the collective), then ONE device->host read; with gloo
        (CPU tests) or a single rank the vector is read first."""
        if self.enabled and self.world > 1 and dist.get_backend(self.group) == "nccl":
            with self._turn():
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            return t.cpu().numpy()
        return self.allreduce_sum(t.cpu().numpy())


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
