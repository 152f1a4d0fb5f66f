"""N > 1 path on CPU: world_size-2 gloo processes shard a batch by documents, encode their shard, exchange
only metadata and reassemble -- result must equal the whole-batch encode.  The device encoder is replaced
by the oracle here (tests may use it); the product code under test is complexity_tokenizer.sharding."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_properties():
    from complexity_tokenizer.sharding import shard_ranges
    rng = np.random.default_rng(0)
    for n_docs in (0, 1, 2, 7, 100, 1000):
        lens = rng.integers(0, 5000, size=n_docs)
        if n_docs > 3:
            lens[rng.integers(0, n_docs)] = 1 << 20                       # one huge document (config 4 shape)
        offs = np.zeros(n_docs + 1, dtype=np.uint64)
        offs[1:] = np.cumsum(lens)
        for world in (1, 2, 4, 8):
            rs = shard_ranges(offs, world)
            assert len(rs) == world and rs[0][0] == 0 and rs[-1][1] == n_docs
            for (a, b), (c, d) in zip(rs, rs[1:]):
                assert a <= b == c <= d                                    # contiguous, ordered, no gaps
            if n_docs >= 100 and world > 1:
                sizes = [int(offs[b] - offs[a]) for a, b in rs]
                assert max(sizes) <= int(offs[-1]) // world + int(lens.max()) + 1   # balanced up to one document


def _worker(rank, world, port, tok_json, result_path):
    sys.path[:0] = [os.path.join(ROOT, p) for p in ('complexity-tokenizer_b200', 'oracle', 'fixtures')]
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    import torch.distributed as dist
    import c_oracle
    import synth
    import torch
    from complexity_tokenizer.sharding import encode_batch_sharded, gather_ids, shard_ranges
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        orc = c_oracle.COracle.from_str(tok_json)
        text, offs = synth.gen_corpus('mixed', 99, 1 << 20, doc_median=900)   # every rank regenerates the same batch
        ids, gioff, meta = encode_batch_sharded(orc, text, offs, rank, world)
        d0, d1 = shard_ranges(offs, world)[rank]
        assert meta[rank]['first_doc'] == d0 and meta[rank]['n_docs'] == d1 - d0 and meta[rank]['n_ids'] == ids.size
        assert sum(m['n_docs'] for m in meta) == len(offs) - 1
        np.savez(result_path % rank, ids=ids, gioff=gioff, d0=d0, d1=d1)
        allids = gather_ids(torch.from_numpy(ids.astype(np.int64)), meta, rank, dst=0)      # optional gather of the ids to rank 0
        if rank == 0:
            np.save(result_path % 99, allids.numpy())
        else:
            assert allids is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_encode_equals_whole_batch(tmp_path, small_tok_json):
    import torch.multiprocessing as mp
    import c_oracle
    import synth
    world, port = 2, 29500 + os.getpid() % 2000
    result_path = str(tmp_path / 'shard%d.npz')
    mp.spawn(_worker, args=(world, port, small_tok_json, result_path), nprocs=world, join=True)
    orc = c_oracle.COracle.from_str(small_tok_json)
    text, offs = synth.gen_corpus('mixed', 99, 1 << 20, doc_median=900)
    wids, woff = orc.encode_packed(text, offs)
    parts = [np.load(result_path % r) for r in range(world)]
    assert np.array_equal(np.concatenate([p['ids'] for p in parts]), wids)
    goff = np.concatenate([p['gioff'][:-1] for p in parts] + [parts[-1]['gioff'][-1:]])
    assert np.array_equal(goff, woff)
    assert np.array_equal(np.load((result_path % 99) + '.npy').astype(np.uint32), wids)   # gather_ids: all ids on rank 0, in document order
