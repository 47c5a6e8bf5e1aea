"""Length-bucketed batching and rank sharding of utterances (host side).

Restates the reference's rule so that a batch -> rank assignment is identical to what
`fairseq-generate` would do (outputs depend on batch composition):
  * `batch_by_size`: greedy token-budget packing over an ordered index list -- a batch is full when
    (n+1) * max_len_in_batch > max_tokens (or n == max_sentences); a full batch is cut down to a multiple
    of `bsz_mult` and the remainder carried over (fairseq/data/data_utils_fast.pyx:17-69).
  * `shard_batches`: batches dealt round-robin to ranks, short shards padded with None -> here: empty
    list (ShardedIterator, fairseq/data/iterators.py:470-500, as used by generate.py:145-160).
  * ordering: longest first (descending length, stable), the order the bench / tests use
    (the collater itself re-sorts each batch by length descending, triplet_dataset.py:174-179).
"""
import math


def ordered_indices(lengths):
    return sorted(range(len(lengths)), key=lambda i: -int(lengths[i]))


def batch_by_size(indices, lengths, max_tokens=2000000, max_sentences=0, bsz_mult=8):
    batches, batch, sample_lens = [], [], []
    sample_len = 0
    for idx in indices:
        n_tok = int(lengths[idx])
        sample_lens.append(n_tok)
        sample_len = max(sample_len, n_tok)
        if max_tokens > 0 and sample_len > max_tokens:
            raise ValueError("utterance %d of %d samples exceeds max_tokens=%d" % (idx, sample_len, max_tokens))
        num_tokens = (len(batch) + 1) * sample_len
        full = len(batch) > 0 and ((max_sentences > 0 and len(batch) == max_sentences)
                                   or (max_tokens > 0 and num_tokens > max_tokens))
        if full:
            mod_len = max(bsz_mult * (len(batch) // bsz_mult), len(batch) % bsz_mult)
            batches.append(batch[:mod_len])
            batch = batch[mod_len:]
            sample_lens = sample_lens[mod_len:]
            sample_len = max(sample_lens) if sample_lens else 0
        batch.append(idx)
    if batch:
        batches.append(batch)
    return batches


def shard_batches(batches, num_shards, shard_id):
    if not 0 <= shard_id < num_shards:
        raise ValueError("shard_id must be between 0 and num_shards")
    sharded_len = int(math.ceil(len(batches) / float(num_shards)))
    mine = list(batches[shard_id::num_shards])
    return mine + [[] for _ in range(sharded_len - len(mine))]
