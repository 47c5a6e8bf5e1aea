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
def ordered_indices(lengths):
    return sorted(range(len(lengths)), key=lambda i: -int(lengths[i]))


def batch_by_size(indices, lengths, max_tokens=2000000, max_sentences=0, bsz_mult=8):
    """Token-budget packing with the reference's outcome (fairseq/data/data_utils_fast.pyx:28-69; pinned to the
    reference's compiled Cython by tests/test_batching.py).  Batches are consecutive slices of the ordered index list, so
    only cut positions are tracked: the open slice order[lo:pos] is closed when taking utterance `pos` as well would
    exceed the budget ((count + 1) * widest > max_tokens) or the sentence cap; a closed slice keeps a whole multiple of
    `bsz_mult` utterances (all of them when it has fewer than one multiple) and the rest stay open for the next batch."""
    order = [int(i) for i in indices]
    width = [int(lengths[i]) for i in order]
    cuts, lo, widest = [], 0, 0
    for pos, w in enumerate(width):
        widest = max(widest, w)
        if 0 < max_tokens < widest:
            raise ValueError("utterance %d of %d samples exceeds max_tokens=%d" % (order[pos], widest, max_tokens))
        count = pos - lo
        if count and ((0 < max_sentences == count) or (max_tokens > 0 and (count + 1) * widest > max_tokens)):
            keep = count if count < bsz_mult else count - count % bsz_mult
            cuts.append((lo, lo + keep))
            lo += keep
            widest = max(width[lo:pos + 1])
    if lo < len(order):
        cuts.append((lo, len(order)))
    return [order[a:b] for a, b in cuts]


def shard_batches(batches, num_shards, shard_id):
    """Every num_shards-th batch starting at shard_id; all shards report the same number of batches, the short ones
    padded with empty batches (ShardedIterator, fairseq/data/iterators.py:470-500)."""
    if not 0 <= shard_id < num_shards:
        raise ValueError("shard_id must be between 0 and num_shards")
    per_shard = -(-len(batches) // num_shards)
    return [list(batches[j]) if j < len(batches) else [] for j in range(shard_id, per_shard * num_shards, num_shards)]


# ---- collation (the caller side of the encoder: SURVEY.md §8(f) row 3) ------------------------------------------------
def collate_waveforms(waves, ids=None, pin=False):
    """Padded batch of 1-D float waveforms as the reference's collater builds it: zero-padded to the longest
    (`_collate_frames(..., is_audio_input=True)`, fairseq/data/audio/speech_to_text_dataset.py:207-225), then rows
    re-ordered by descending length with torch's own sort (triplet_dataset.py:165-179).
    -> (ids [B] int64 in batch order, src_tokens [B, L] float32, src_lengths [B] int64); `pin` puts src_tokens in
    pinned host memory (the form `encoder.forward_many` copies on its lanes' own streams)."""
    import numpy as np
    import torch
    if len(waves) == 0:
        raise ValueError("empty batch")
    # int16 inputs (16-bit PCM, `audio_io.read_pcm16`) stay int16: the wire format the encoder de-quantises on the GPU
    pcm = all(getattr(w, "dtype", None) in (torch.int16, np.dtype("int16")) for w in waves)
    dt = torch.int16 if pcm else torch.float32
    waves = [torch.as_tensor(w, dtype=dt).reshape(-1) for w in waves]
    n = torch.tensor([w.numel() for w in waves], dtype=torch.long)
    out = torch.zeros(len(waves), int(n.max()), dtype=dt)
    for i, w in enumerate(waves):
        out[i, :w.numel()] = w
    n_sorted, order = n.sort(descending=True)
    idx = torch.arange(len(waves)) if ids is None else torch.as_tensor(ids, dtype=torch.long)
    out = out.index_select(0, order)
    if pin and torch.cuda.is_available():
        out = out.pin_memory()
    return idx.index_select(0, order), out, n_sorted


def plan_batches(lengths, max_tokens=2000000, max_sentences=0, bsz_mult=8, num_shards=1, shard_id=0):
    """Utterance indices of this rank's batches: longest-first order, the reference's token-budget packing, round-robin
    sharding (generate.py:145-160)."""
    batches = batch_by_size(ordered_indices(lengths), lengths, max_tokens, max_sentences, bsz_mult)
    return [b for b in shard_batches(batches, num_shards, shard_id) if b]


def encode_utterances(encoder, waves, max_tokens=2000000, bsz_mult=8, n_lanes=3, num_shards=1, shard_id=0):
    """waveform list -> {utterance index: memories [M, 512]} for this rank's share, through the public module API:
    reference batching + collation on the host, pinned batches, `encoder.forward_many` on stream lanes."""
    lengths = [len(w) for w in waves]
    todo = plan_batches(lengths, max_tokens, 0, bsz_mult, num_shards, shard_id)
    coll = [collate_waveforms([waves[i] for i in b], ids=b, pin=True) for b in todo]
    outs = encoder.forward_many([(w, n) for _, w, n in coll], n_lanes=n_lanes)
    result = {}
    for (ids, _, _), o in zip(coll, outs):
        mem = o.encoder_out                                     # [M, B, 512]
        for j, i in enumerate(ids.tolist()):
            result[i] = mem[:, j]
    return result
