"""N>1 host logic on CPU: world_size-2 gloo process group.  The per-GPU encoder is replaced
by the C oracle (tests may use it as a stand-in; the product path uses the CUDA library),
so what is checked here is the partition + the single count exchange + global offsets:
concatenating the ranks' outputs must equal the single-process result
(reference semantics: encode_batch[i] == encode(text_i), tests/cl100k.rs:191-214)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir):
    for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import synth
    from splintr_b200 import presets as P
    from splintr_b200.distributed import encode_sharded
    from oracle.c_oracle import COracle
    p = P.PRESETS["cl100k_base"]
    vb = P.load_vocab_bytes(p.vocab_file)
    data, offsets = synth.cfg1(vb, 301)
    orc = COracle(vb, p.pattern, p.special_tokens, False)
    ids, goff, d0, base, total = encode_sharded(lambda b, o: orc.encode_packed(b, o, n_threads=1), data, offsets, rank, world)
    np.savez(os.path.join(outdir, f"r{rank}.npz"), ids=ids, goff=goff, d0=d0, base=base, total=total)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_encode(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    import synth
    from splintr_b200 import presets as P
    from conftest import c_oracle
    vb = P.load_vocab_bytes("cl100k_base.tiktoken")
    data, offsets = synth.cfg1(vb, 301)
    want_ids, want_off = c_oracle("cl100k_base").encode_packed(data, offsets)
    parts = [np.load(os.path.join(tmp_path, f"r{r}.npz")) for r in range(world)]
    assert all(int(p["total"]) == len(want_ids) for p in parts)
    got = np.concatenate([p["ids"] for p in parts])
    assert np.array_equal(got, want_ids)
    d = 0
    for p in parts:
        assert int(p["d0"]) == d
        g = p["goff"]
        assert np.array_equal(g, want_off[d:d + len(g)])
        assert int(p["base"]) == int(want_off[d])
        d += len(g) - 1
    assert d == len(offsets) - 1


def test_shard_bounds_properties():
    from splintr_b200.distributed import shard_bounds
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        for n_docs in (0, 1, 3, 100):
            lens = rng.integers(0, 50, size=n_docs)
            off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
            dlo = shard_bounds(off, world)
            assert dlo[0] == 0 and dlo[-1] == n_docs and np.all(np.diff(dlo) >= 0)
            if n_docs == 100 and world > 1:
                per = np.diff(off[dlo].astype(np.int64))
                assert per.max() - per.min() <= 2 * 50
