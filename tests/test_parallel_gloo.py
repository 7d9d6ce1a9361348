"""CPU: the N > 1 plumbing (utterance scatter / transcript gather) on a
world-size-2 gloo group.  The per-rank compute is replaced by a deterministic stand-in so that
only the sharding logic is under test (the CUDA compute cannot run without a GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from libreasr_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_transcribe(audio, lens, max_iters):
    """tokens of utterance i = [round(1000*mean), len or n, first sample as int] -> checks that the right
    rows (and lengths) reached the right rank and come back in global order."""
    n = audio.shape[0]
    U = 6
    tokens = torch.zeros(n, U, dtype=torch.int32)
    ntok = torch.zeros(n, dtype=torch.int32)
    for i in range(n):
        ln = int(lens[i]) if lens is not None else audio.shape[1]
        vals = [int(round(float(audio[i, :ln].sum()))), ln, int(audio[i, 0])]
        k = 1 + (int(audio[i, 0]) % 3)
        tokens[i, :k] = torch.tensor(vals[:k], dtype=torch.int32)
        ntok[i] = k
    return {"tokens": tokens, "ntok": ntok}


def _worker(rank, world, port, n_items, with_lens, q, balance=False, block=0):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 64
        audio = lens = None
        if rank == 0:
            audio = torch.arange(n_items, dtype=torch.float32)[:, None].repeat(1, n) + 1.0
            lens = torch.tensor([n - (i % 5) for i in range(n_items)], dtype=torch.int32) if with_lens else None
        out = parallel.transcribe_sharded(None, audio, lens, n_items, n, transcribe_fn=_fake_transcribe, balance=balance, block=block)
        if rank == 0:
            want = _fake_transcribe(audio, lens, 3)
            exp = [want["tokens"][i, : int(want["ntok"][i])].tolist() for i in range(n_items)]
            q.put(("ok", out == exp))
        else:
            q.put(("none", out is None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items,with_lens", [(7, True), (8, False), (1, True)])
def test_scatter_gather_world2(n_items, with_lens):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, with_lens, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [("none", True), ("ok", True)]


@pytest.mark.parametrize("n_items,block", [(11, 2), (8, 4), (3, 64), (1, 2)])
def test_block_pipelined_scatter_world2(n_items, block):
    """The block-pipelined scatter (bench.py strong-scaling leg): blocks of `block` utterances dealt round-robin over the
    ranks, each transcribed on arrival; the caller still gets every utterance back in global order."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, False, q, False, block)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [("none", True), ("ok", True)]


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 5, 32, 33, 2048):
        for w in (1, 2, 4, 8):
            b = [parallel.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_balanced_sharding_world2_returns_input_order():
    """Length-balanced dealing permutes what each rank computes but not what the caller gets back."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 9, True, q, True)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [("none", True), ("ok", True)]


def test_balanced_order_is_a_permutation_that_evens_out_lengths():
    import random

    rnd = random.Random(0)
    for n, w in ((9, 2), (32, 4), (33, 8), (5, 8), (2048, 8)):
        lens = [rnd.choice([160000, 240000]) - rnd.randrange(0, 20000) for _ in range(n)]
        order = parallel.balanced_order(lens, w)
        assert sorted(order) == list(range(n))
        loads, sizes = [], []
        for r in range(w):
            lo, hi = parallel.shard_bounds(n, w, r)
            loads.append(sum(lens[i] for i in order[lo:hi]))
            sizes.append(hi - lo)
        contiguous = [sum(lens[lo:hi]) for lo, hi in (parallel.shard_bounds(n, w, r) for r in range(w))]
        if n >= 4 * w:   # with enough utterances per rank the dealt split is tighter than (or as tight as) contiguous blocks
            full = [l for l, s in zip(loads, sizes) if s == max(sizes)]
            assert max(full) - min(full) <= max(lens)
            assert max(loads) <= max(contiguous) + max(lens) // 8
