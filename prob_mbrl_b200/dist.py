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


class PeerBuffer:
    """A zeroed device buffer of this rank that every rank of the node has mapped (CUDA IPC through
    ``pmb_peer_alloc`` / ``pmb_peer_open``).  Collective construction; ``ptrs[r]`` = rank r's buffer as seen here."""

    def __init__(self, nbytes, device):
        import ctypes as C
        from . import _lib
        self.lib = _lib.load()
        self.rank, self.world = world()
        if self.world > 16:
            raise ValueError("the peer-memory exchange serves one node (<= 16 GPUs)")
        # Every rank takes part in every collective below whatever happens locally (a rank that raised half-way would
        # leave the others hanging): failures are recorded in self.error and agreed on by the caller (all_ok()).
        self.error, self.own, self.opened = None, None, []
        self.ptrs = (C.c_void_p * self.world)()
        with torch.cuda.device(device):
            own, handle = C.c_void_p(), C.create_string_buffer(64)
            try:
                self._check(self.lib.pmb_peer_alloc(int(nbytes), C.byref(own), handle))
                self.own = own.value
            except _lib.LibraryError as e:
                self.error = e
            handles = [None] * self.world
            torch.distributed.all_gather_object(handles, bytes(handle.raw) if self.error is None else b"")
            if self.error is None and any(len(h) != 64 for h in handles):
                self.error = _lib.LibraryError(-4, "a peer could not allocate its exchange buffer")
            for r, h in enumerate(handles):
                if self.error is not None:
                    break
                if r == self.rank:
                    self.ptrs[r] = self.own
                else:
                    p = C.c_void_p()
                    try:
                        # fails across nodes, or between devices without peer access: the caller falls back
                        self._check(self.lib.pmb_peer_open(C.create_string_buffer(h, 64), C.byref(p)))
                    except _lib.LibraryError as e:
                        self.error = e
                        break
                    self.opened.append(p.value)
                    self.ptrs[r] = p.value
            torch.cuda.synchronize()
        torch.distributed.barrier()         # every buffer is zeroed and mapped before anyone writes

    def all_ok(self, device):
        """Collective: True when every rank allocated and mapped everything."""
        flag = torch.tensor([0 if self.error is not None else 1], dtype=torch.int32, device=device)
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
        return bool(int(flag.item()))

    def _check(self, rc):
        if rc != 0:
            from . import _lib
            raise _lib.LibraryError(rc, self.lib.pmb_peer_last_error().decode("utf-8", "replace"))

    def close(self):
        if getattr(self, "own", None) is None and not getattr(self, "opened", None):
            return
        torch.cuda.synchronize()
        for p in self.opened:
            self.lib.pmb_peer_close(p)
        if self.own is not None:
            self.lib.pmb_peer_free(self.own)
        self.own, self.opened = None, []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PeerAllReduce:
    """The gradient all-reduce over NVLink peer memory (``pmb_peer_allreduce``): stream-ordered kernels, no NCCL call,
    capturable in the iteration's CUDA graph, results bitwise identical on every rank.  Construction is collective
    (every rank of the default process group, same ``n``)."""

    def __init__(self, n, device):
        from . import _lib
        self.lib = _lib.load()
        self.n = int(n)
        self.rank, self.world = world()
        self.buf = PeerBuffer(self.lib.pmb_peer_buffer_bytes(self.n, min(self.world, 16)), device)
        self.state = torch.zeros(3, dtype=torch.int64, device=device)
        self.ok = self.buf.all_ok(device)

    def __call__(self, flat, loss=None):
        """In-place sum of ``flat`` (float32, ``n`` elements, contiguous) over the ranks, on the current stream."""
        from . import _lib
        assert flat.numel() == self.n and flat.dtype == torch.float32 and flat.is_contiguous()
        rc = self.lib.pmb_peer_allreduce(flat.data_ptr(), flat.data_ptr(), self.n, self.world, self.rank, self.buf.ptrs,
                                         self.state.data_ptr(), _lib.current_stream_ptr())
        if rc != 0:
            raise _lib.LibraryError(rc, self.lib.pmb_peer_last_error().decode("utf-8", "replace"))
        if loss is not None:
            torch.distributed.all_reduce(loss)

    def close(self):
        self.buf.close()


def gradient_sync(n, device):
    """What a sharded iteration uses for its one collective: the peer-memory exchange (default), or the NCCL all-reduce
    with ``PMB_GRAD_SYNC=nccl``.  None in a single-process run."""
    import os
    if world()[1] <= 1:
        return None
    if os.environ.get("PMB_GRAD_SYNC", "peer") == "nccl":
        return allreduce_gradient
    peer = PeerAllReduce(n, device)
    if not peer.ok:
        # e.g. ranks on different nodes, or devices without peer access: every rank agrees on the NCCL all-reduce
        import warnings
        warnings.warn("prob_mbrl_b200: peer-memory gradient exchange unavailable (%s); using the NCCL all-reduce"
                      % (peer.buf.error,))
        peer.close()
        return allreduce_gradient
    return peer
