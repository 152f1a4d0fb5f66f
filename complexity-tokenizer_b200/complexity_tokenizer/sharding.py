"""Document sharding across the GPUs of one box.

Documents are independent in the reference (`texts.par_iter()`, src/huggingface/mod.rs:694-696), so
the batch splits into contiguous document ranges, one per rank, balanced by BYTES (not by count:
config 4's 1 MiB documents next to tiny ones show why).  Every rank encodes its range on its own GPU
with its own replica of the (small) tables; there is NO collective on the data path.  The only thing
that crosses ranks is per-shard metadata -- (first_doc, n_docs, n_ids) -- from which every rank
derives the global id offsets of its shard.
"""
import numpy as np


def shard_ranges(offsets, n_shards):
    """offsets: uint64[n_docs+1] byte offsets of a packed batch -> list of (first_doc, end_doc), one per
    shard, contiguous, covering [0, n_docs), balanced by bytes.  Deterministic, no communication."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    n_docs = len(offsets) - 1
    total = int(offsets[-1]) if n_docs >= 0 else 0
    cuts = [0]
    for s in range(1, n_shards):
        target = total * s // n_shards
        # first document whose start is >= target (documents are never split)
        d = int(np.searchsorted(offsets[:n_docs + 1], target, side='left'))
        d = min(max(d, cuts[-1]), n_docs)
        cuts.append(d)
    cuts.append(n_docs)
    return [(cuts[i], cuts[i + 1]) for i in range(n_shards)]


def local_shard(text, offsets, rank, world):
    """-> (text_view, rebased_offsets, first_doc) of this rank's contiguous document range."""
    d0, d1 = shard_ranges(offsets, world)[rank]
    b0, b1 = int(offsets[d0]), int(offsets[d1])
    return text[b0:b1], (np.asarray(offsets[d0:d1 + 1], dtype=np.uint64) - np.uint64(b0)), d0


def exchange_shard_metadata(first_doc, n_docs, n_ids, group=None):
    """All ranks learn every shard's (first_doc, n_docs, n_ids) and derive ids_base = ids emitted by the
    shards before them.  One tiny all_gather (3 int64 per rank); works on NCCL (device tensors) and gloo."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [dict(first_doc=int(first_doc), n_docs=int(n_docs), n_ids=int(n_ids), ids_base=0)]
    world = dist.get_world_size(group)
    dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend(group) == 'nccl' else torch.device('cpu')
    mine = torch.tensor([int(first_doc), int(n_docs), int(n_ids)], dtype=torch.int64, device=dev)
    out = torch.empty(world * 3, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(out, mine, group=group)             # one collective, one read-back
    flat = out.tolist()
    rows = [tuple(flat[3 * r:3 * r + 3]) for r in range(world)]
    base, meta = 0, []
    for fd, nd, ni in rows:
        meta.append(dict(first_doc=fd, n_docs=nd, n_ids=ni, ids_base=base))
        base += ni
    return meta


def encode_batch_sharded(tok, text, offsets, rank, world, group=None):
    """Encode this rank's shard of a packed batch with `tok` (anything with encode_packed) and return
    (ids, global_ids_off_of_local_docs, meta).  global offsets = local offsets + ids_base of the shard."""
    t, o, d0 = local_shard(text, offsets, rank, world)
    ids, ioff = tok.encode_packed(t, o)
    meta = exchange_shard_metadata(d0, len(o) - 1, int(ioff[-1]), group)
    return ids, ioff + np.uint64(meta[rank]['ids_base']), meta


def gather_ids(ids, meta, rank, dst=0, group=None):
    """Optional: bring every shard's ids to ONE rank.  `ids` is this rank's torch tensor of ids
    (on its GPU with NCCL -- the copies then go GPU to GPU over NVLink/NVSwitch -- or on the CPU with gloo); `meta` is
    exchange_shard_metadata()'s result.  Returns on `dst` one tensor holding all ids in document order (shard r at
    meta[r]['ids_base']), elsewhere None.  Point-to-point sends: no collective, nothing moves that is not needed."""
    import torch
    import torch.distributed as dist
    world = len(meta)
    if world == 1 or not (dist.is_available() and dist.is_initialized()):
        return ids
    total = meta[-1]['ids_base'] + meta[-1]['n_ids']
    if rank == dst:
        out = torch.empty(total, dtype=ids.dtype, device=ids.device)
        base = meta[rank]['ids_base']
        out[base:base + meta[rank]['n_ids']].copy_(ids[:meta[rank]['n_ids']])
        ops = [dist.P2POp(dist.irecv, out[m['ids_base']:m['ids_base'] + m['n_ids']], r, group)
               for r, m in enumerate(meta) if r != dst and m['n_ids']]
        for req in (dist.batch_isend_irecv(ops) if ops else []):
            req.wait()
        return out
    if meta[rank]['n_ids']:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, ids[:meta[rank]['n_ids']].contiguous(), dst, group)]):
            req.wait()
    return None
