"""Gradient all-reduce of the training configuration (BASELINE configs[4]).

What the reference does (`LegacyDistributedDataParallel.all_reduce`, fairseq/legacy_distributed_data_parallel.py:94-178, selected
by `--ddp-backend no_c10d`, chimera/scripts/train-en2any-ST.sh:54): after the whole backward pass it packs every gradient into ONE
flat buffer (up to 2^28 elements), divides it by the world size, calls `dist.all_reduce` (sum) once, and copies the slices back --
no overlap with compute.

What this module does with the same result (sum of gradients pre-divided by the world size, every rank ends with identical
values): gradients are grouped into BUCKETS in the order the backward pass produces them (memory stage first, feature extractor
last); a bucket is flattened, scaled and handed to `torch.distributed.all_reduce(async_op=True)` as soon as its last gradient
exists (`ready()`), so on NCCL the transfers over NVLink / NVSwitch run on the communicator's own stream underneath the backward
kernels of the earlier layers; `finish()` waits and scatters the reduced values back.  NVSwitch gives every pair of GPUs full
bandwidth, so buckets are sized for launch latency and overlap (default 32 MB), not for link count.
The backend is whatever process group is initialised (NCCL on the GPUs; gloo in the CPU tests).
"""
from collections import OrderedDict

import torch
import torch.distributed as dist


def plan_buckets(named_sizes, bucket_bytes=32 << 20, elem_bytes=4):
    """[(name, numel)] in backward order -> list of buckets (lists of names); a tensor larger than a bucket travels alone
    (the reference's "all-reduce big params directly", legacy_distributed_data_parallel.py:159-161)."""
    cap = max(1, bucket_bytes // elem_bytes)
    buckets, cur, fill = [], [], 0
    for name, n in named_sizes:
        if n >= cap:
            if cur:
                buckets.append(cur)
                cur, fill = [], 0
            buckets.append([name])
            continue
        if fill + n > cap:
            buckets.append(cur)
            cur, fill = [], 0
        cur.append(name)
        fill += n
    if cur:
        buckets.append(cur)
    return buckets


class GradAllReducer:
    def __init__(self, named_sizes, world_size=None, process_group=None, bucket_bytes=32 << 20, comm_dtype=None, persistent=False):
        """comm_dtype: dtype of the flat buffers on the wire (None: the gradients' own fp32; torch.bfloat16 halves the bytes, the
        reference's --fp16 runs reduce fp16 gradients the same way).  persistent: the flat buffers are allocated once and reused, so
        the reduced gradients keep their addresses from step to step (a CUDA-graphed optimizer update can read them)."""
        self.comm_dtype = comm_dtype
        self.persistent = persistent
        self._stage, self._comm = {}, {}
        self.group = process_group
        self.world = world_size if world_size is not None else (dist.get_world_size(process_group) if dist.is_initialized() else 1)
        self.sizes = OrderedDict(named_sizes)
        self.buckets = plan_buckets(list(self.sizes.items()), bucket_bytes)
        self.where = {n: (bi, i) for bi, b in enumerate(self.buckets) for i, n in enumerate(b)}
        self.reset()

    def reset(self):
        self.have = [dict() for _ in self.buckets]
        self.inflight = {}                                          # bucket index -> (flat buffer, work handle)

    def ready(self, grads):
        """Hand over gradients ({name: tensor}) as the backward pass produces them; complete buckets start reducing at once."""
        touched = set()
        for name, g in grads.items():
            if name not in self.where:
                continue
            bi, _ = self.where[name]
            self.have[bi][name] = g
            touched.add(bi)
        for bi in sorted(touched):
            if bi not in self.inflight and len(self.have[bi]) == len(self.buckets[bi]):
                self._launch(bi)

    def _launch(self, bi):
        names = self.buckets[bi]
        parts = [self.have[bi][n].reshape(-1) for n in names]
        if self.persistent:
            if bi not in self._stage:
                self._stage[bi] = torch.empty(sum(p.numel() for p in parts), dtype=parts[0].dtype, device=parts[0].device)
                if self.comm_dtype is not None and self.comm_dtype != parts[0].dtype:
                    self._comm[bi] = torch.empty_like(self._stage[bi], dtype=self.comm_dtype)
            flat = self._stage[bi]
            torch.cat(parts, out=flat)
            flat.div_(self.world)
            if bi in self._comm:
                self._comm[bi].copy_(flat)
                flat = self._comm[bi]
        else:
            flat = torch.cat(parts) if len(names) > 1 else parts[0].clone()
            flat.div_(self.world)                                   # pre-divided, as the reference (:126-127)
            if self.comm_dtype is not None and flat.dtype != self.comm_dtype:
                flat = flat.to(self.comm_dtype)
        work = None
        if self.world > 1:
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.inflight[bi] = (flat, work)

    def finish(self, grads=None):
        """Wait for every bucket; missing gradients count as zeros (parameters without a gradient, :141-143).
        -> {name: reduced gradient} (views of the flat buffers, shaped like the inputs)."""
        out = {}
        for bi, names in enumerate(self.buckets):
            if bi not in self.inflight:
                for n in names:
                    if n not in self.have[bi]:
                        ref = next(iter(self.have[bi].values()), None)
                        if ref is None and grads:
                            ref = next(iter(grads.values()))
                        dev = ref.device if ref is not None else "cpu"
                        self.have[bi][n] = torch.zeros(self.sizes[n], dtype=torch.float32, device=dev)
                self._launch(bi)
        for bi, names in enumerate(self.buckets):
            flat, work = self.inflight[bi]
            if work is not None:
                work.wait()
            off = 0
            for n in names:
                sz = self.sizes[n]
                out[n] = flat[off:off + sz].view_as(self.have[bi][n]) if self.have[bi][n].numel() == sz else flat[off:off + sz]
                off += sz
        self.reset()
        return out


def all_reduce_gradients(grads, world_size=None, process_group=None, bucket_bytes=32 << 20):
    """One-shot form: {name: gradient} -> {name: sum over ranks / world size}, identical on every rank."""
    r = GradAllReducer([(n, g.numel()) for n, g in grads.items()], world_size, process_group, bucket_bytes)
    r.ready(grads)
    return r.finish(grads)
