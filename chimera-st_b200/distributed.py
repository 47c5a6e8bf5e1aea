"""One-process-per-GPU plumbing for the utterance-sharded path (torch.distributed; NCCL on GPUs, gloo in
the CPU tests).  Inference needs NO data-path collective (independent utterances): ranks
only agree on the batch -> rank assignment (deterministic, the reference's ShardedIterator rule) and reduce
two scalars for reporting (sum of audio seconds, max of elapsed time)."""
import os

import torch
import torch.distributed as dist

from . import batching


def env_rank_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend=None):
    rank, world, local_rank = env_rank_world()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kw)
    return rank, world, local_rank


def shard_utterances(lengths, world, rank, max_tokens=2000000, bsz_mult=8):
    """Global length-sorted token-budget batches, dealt round-robin: -> this rank's batches (index lists)."""
    n_total = len(batching.batch_by_size(batching.ordered_indices(lengths), lengths, max_tokens, 0, bsz_mult))
    return batching.plan_batches(lengths, max_tokens, 0, bsz_mult, world, rank), n_total


def _scalar(x, device):
    return torch.tensor([float(x)], dtype=torch.float64, device=device)


def reduce_max(x, device="cpu"):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = _scalar(x, device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(x, device="cpu"):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = _scalar(x, device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def finalize():
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
