"""torch.distributed helpers for one-process-per-GPU training on NVLink/NVSwitch.

Same function names and semantics as the reference's Miscellaneous/distributed.py
(get_rank :18, synchronize :28, get_world_size :43, reduce_sum :53, gather_grad :66,
all_gather :78, reduce_loss_dict :113), plus what the reference never had (it only ever ran
nn.DataParallel): ``init_distributed`` and a bucketed gradient all-reduce that
overlaps with backward (``GradBucketReducer``), replacing the per-parameter unbucketed
all_reduce of ``gather_grad`` (:72-75).

Inference shards by batch and needs none of this (no data-path collective).
"""
import os
import pickle

import torch
from torch import distributed as dist


def init_distributed(backend=None, device=None):
    """Initialise the default process group from the torchrun environment (RANK, WORLD_SIZE,
    LOCAL_RANK, MASTER_ADDR/PORT).  NCCL on CUDA, gloo on CPU.  Returns (rank, world, local_rank)."""
    if not dist.is_available():
        return 0, 1, 0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kwargs["device_id"] = torch.device("cuda", local_rank) if device is None else device
        dist.init_process_group(backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local_rank


def _active():
    return dist.is_available() and dist.is_initialized()


def get_rank():
    return dist.get_rank() if _active() else 0


def get_world_size():
    return dist.get_world_size() if _active() else 1


def synchronize():
    """Barrier across ranks (no-op for a single process)."""
    if _active() and dist.get_world_size() > 1:
        dist.barrier()


def reduce_sum(tensor):
    """Sum of ``tensor`` over ranks (returns a new tensor; the input is left untouched)."""
    if not _active():
        return tensor
    out = tensor.clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out


def shard_batch(n, rank=None, world=None):
    """Contiguous [start, stop) slice of a batch of n items owned by ``rank`` (batch sharding for
    inference sweeps, BASELINE config 5).  Remainders go to the lowest ranks."""
    rank = get_rank() if rank is None else rank
    world = get_world_size() if world is None else world
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


# ------------------------------------------------------------------ gradient averaging
def _bucketize(params, bucket_bytes):
    """Greedy buckets in REVERSE parameter order (gradients become ready back-to-front)."""
    buckets, cur, size = [], [], 0
    for p in reversed(list(params)):
        if not p.requires_grad:
            continue
        nbytes = p.numel() * p.element_size()
        if cur and (size + nbytes > bucket_bytes or cur[0].dtype != p.dtype or cur[0].device != p.device):
            buckets.append(cur)
            cur, size = [], 0
        cur.append(p)
        size += nbytes
    if cur:
        buckets.append(cur)
    return buckets


def gather_grad(params, bucket_bytes=32 << 20):
    """Average ``.grad`` over ranks.  Same contract as the reference (:66-75) but with flat
    ~32 MB buckets (one collective per bucket instead of one per parameter)."""
    world = get_world_size()
    if world == 1:
        return
    for bucket in _bucketize(params, bucket_bytes):
        grads = [p.grad for p in bucket if p.grad is not None]
        if not grads:
            continue
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()


class GradBucketReducer:
    """Bucketed gradient all-reduce overlapped with backward.

    Parameters are packed into ~``bucket_mb`` MB flat buckets in reverse registration order.  A
    post-accumulate-grad hook counts arrivals; when the last gradient of a bucket lands, the
    bucket's gradients are copied into its persistent flat buffer and all-reduced asynchronously
    (on NCCL's own stream, ordered after the producing kernels through an event), while autograd
    keeps computing earlier layers.  ``finish()`` waits for the outstanding collectives, divides by
    the world size and scatters the averages back into ``.grad``.  Two independent reducers are used
    for the two parameter groups that step at different times (G + encoders, D),
    train_3_encoder.py:476-477,557-558.

    Contract:
      * exactly ONE parameter-gradient-producing backward between two ``finish()`` calls (no gradient
        accumulation over micro-batches: a second arrival for a parameter raises).  Double-backward
        passes (R1, path length) are fine -- ``autograd.grad(create_graph=True)`` does not accumulate into
        ``.grad``, so the hooks fire only in the final backward;
      * every collective has the same size on every rank: a bucket always carries all of its parameters,
        zero-filled where a parameter received no gradient on this rank.  Such a parameter keeps
        ``grad = None`` locally, so the set of parameters that receive gradients must still be the same on
        every rank (as with DistributedDataParallel without ``find_unused_parameters``).
    """

    def __init__(self, params, bucket_mb=32):
        self.world = get_world_size()
        self.buckets = _bucketize(params, int(bucket_mb * (1 << 20)))
        self.index = {}
        self.hooks = []
        self.flat = [None] * len(self.buckets)          # persistent flat buffers, allocated on first use
        self.slots = []
        for bi, bucket in enumerate(self.buckets):
            off, sl = 0, []
            for p in bucket:
                self.index[p] = bi
                sl.append((off, p.numel()))
                off += p.numel()
                if self.world > 1:
                    self.hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
            self.slots.append(sl)
        self.exposed_ms = []                             # per finish(): time the step waited for communication
        self._timing = False
        self._reset()

    def bucket_bytes(self):
        return [sum(p.numel() * p.element_size() for p in b) for b in self.buckets]

    def enable_timing(self, on=True):
        """Record, per ``finish()``, how long the compute stream had to wait for the collectives that were still
        running when backward ended (the part of the all-reduce that backward did NOT hide)."""
        self._timing = on
        self.exposed_ms = []

    def _reset(self):
        self.pending = [len(b) for b in self.buckets]
        self.launched = [False] * len(self.buckets)
        self.inflight = []

    def _on_grad(self, p):
        bi = self.index[p]
        if self.pending[bi] <= 0 or self.launched[bi]:
            raise RuntimeError("GradBucketReducer: a parameter received a second gradient before finish() -- gradient "
                               "accumulation over several backward() calls is not supported; call finish() after each")
        self.pending[bi] -= 1
        if self.pending[bi] == 0:
            self._launch(bi)

    def _launch(self, bi):
        bucket = self.buckets[bi]
        flat = self.flat[bi]
        if flat is None:
            n = self.slots[bi][-1][0] + self.slots[bi][-1][1]
            flat = self.flat[bi] = torch.zeros(n, device=bucket[0].device, dtype=bucket[0].dtype)
        for p, (off, n) in zip(bucket, self.slots[bi]):
            if p.grad is not None:
                flat[off:off + n].copy_(p.grad.reshape(-1))
            else:
                flat[off:off + n].zero_()                # rank-invariant message size
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
        self.launched[bi] = True
        self.inflight.append((work, bi))

    def finish(self):
        """Call after ``loss.backward()`` and before ``optimizer.step()``."""
        if self.world > 1:
            for bi in range(len(self.buckets)):           # buckets some of whose parameters received no gradient
                if not self.launched[bi] and any(p.grad is not None for p in self.buckets[bi]):
                    self._launch(bi)
            timed = self._timing and self.inflight and self.flat[self.inflight[0][1]].is_cuda
            if timed:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            for work, bi in self.inflight:
                work.wait()
                flat = self.flat[bi]
                flat.div_(self.world)
                for p, (off, n) in zip(self.buckets[bi], self.slots[bi]):
                    if p.grad is not None:
                        p.grad.copy_(flat[off:off + n].view_as(p.grad))
            if timed:
                e1.record()
                self.exposed_ms.append((e0, e1))
        self._reset()

    def exposed_times_ms(self):
        """Milliseconds per finish() between the end of backward and the last averaged gradient (needs a device sync)."""
        out = []
        for e in self.exposed_ms:
            out.append(e[0].elapsed_time(e[1]) if isinstance(e, tuple) else e)
        return out

    def remove(self):
        for h in self.hooks:
            h.remove()
        self.hooks = []


def all_gather(data):
    """Gather arbitrary picklable objects from every rank (list indexed by rank)."""
    world = get_world_size()
    if world == 1:
        return [data]
    out = [None] * world
    dist.all_gather_object(out, data)
    return out


def reduce_loss_dict(loss_dict):
    """Average a dict of scalar loss tensors onto rank 0 (other ranks get the un-normalised sum),
    keys processed in sorted order as in the reference (:113-135)."""
    world = get_world_size()
    if world < 2:
        return loss_dict
    with torch.no_grad():
        keys = sorted(loss_dict.keys())
        losses = torch.stack([loss_dict[k] for k in keys], 0)
        dist.reduce(losses, dst=0)
        if dist.get_rank() == 0:
            losses /= world
        return {k: v for k, v in zip(keys, losses)}
