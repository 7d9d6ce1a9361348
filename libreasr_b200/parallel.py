"""Utterance sharding over the GPUs of one box.

The path shards naturally: utterances / streams are independent, weights are replicated and
no collective sits inside the compute loop.  The only exchange steps are
  * ``scatter_audio``: rank 0 holds ``[N, n]`` audio (+ lengths) and sends each rank its
    contiguous block of ``ceil(N / world)`` utterances,
  * ``gather_tokens``: every rank returns its padded ``[n_local, U]`` int32 token array and
    counts to rank 0.
Both use ``torch.distributed`` point-to-point / collective calls, i.e. NCCL over NVLink on
the GPU box and gloo on CPU (the partition logic below is backend independent and covered by
world-size-2 gloo tests).
"""
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block partition: rank r owns [lo, hi); blocks differ by at most one item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def balanced_order(lens, world: int) -> List[int]:
    """Permutation that evens out the work of ragged batches ("optionally length-sorted then round-robin
    to balance T"): utterances sorted by length are dealt to the ranks in snake order, and each rank's share is laid out as
    its contiguous block of ``shard_bounds``.  ``order[i]`` = index of the utterance placed at position i."""
    n = len(lens)
    idx = sorted(range(n), key=lambda i: (-int(lens[i]), i))
    per_rank = [[] for _ in range(world)]
    caps = [shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world)]
    r, step = 0, 1
    for i in idx:
        while len(per_rank[r]) >= caps[r]:          # a rank that is full is skipped
            r, step = (r + step, step) if 0 <= r + step < world else (r, -step)
        per_rank[r].append(i)
        nr = r + step
        if nr < 0 or nr >= world:
            step = -step                             # snake: 0..w-1, w-1..0, ...
        else:
            r = nr
    return [i for part in per_rank for i in part]


def scatter_audio(audio: Optional[torch.Tensor], lens: Optional[torch.Tensor], n_items: int, n_samples: int,
                  device: torch.device, src: int = 0):
    """Rank ``src`` passes ``audio [N, n]`` (and optional ``lens [N]``); every rank returns its
    shard ``(audio_local [n_local, n], lens_local [n_local] | None)`` on ``device``."""
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_bounds(n_items, world, rank)
    local = torch.empty(hi - lo, n_samples, dtype=torch.float32, device=device)
    has_lens = torch.tensor([0 if lens is None else 1], device=device) if rank == src else torch.zeros(1, dtype=torch.long, device=device)
    has_lens = has_lens.to(torch.long)
    dist.broadcast(has_lens, src=src)
    local_lens = torch.empty(hi - lo, dtype=torch.int32, device=device) if int(has_lens) else None
    if rank == src:
        reqs = []
        for r in range(world):
            a, b = shard_bounds(n_items, world, r)
            if r == src:
                local.copy_(audio[a:b])
                if local_lens is not None:
                    local_lens.copy_(lens[a:b])
            elif b > a:
                reqs.append(dist.isend(audio[a:b].contiguous().to(device), dst=r))
                if local_lens is not None:
                    reqs.append(dist.isend(lens[a:b].to(device=device, dtype=torch.int32).contiguous(), dst=r))
        for q in reqs:
            q.wait()
    elif hi > lo:
        dist.recv(local, src=src)
        if local_lens is not None:
            dist.recv(local_lens, src=src)
    return local, local_lens


def gather_tokens(tokens: torch.Tensor, ntok: torch.Tensor, n_items: int, dst: int = 0) -> Optional[List[List[int]]]:
    """Every rank passes its ``tokens [n_local, U]`` / ``ntok [n_local]`` (int32); rank ``dst``
    gets the list of token lists of all ``n_items`` utterances in global order, others None."""
    world, rank = dist.get_world_size(), dist.get_rank()
    U = tokens.shape[1]
    n_max = shard_bounds(n_items, world, 0)[1]
    pad_t = torch.zeros(n_max, U, dtype=torch.int32, device=tokens.device)
    pad_n = torch.zeros(n_max, dtype=torch.int32, device=tokens.device)
    pad_t[: tokens.shape[0]] = tokens
    pad_n[: ntok.shape[0]] = ntok
    gt = [torch.empty_like(pad_t) for _ in range(world)] if rank == dst else None
    gn = [torch.empty_like(pad_n) for _ in range(world)] if rank == dst else None
    dist.gather(pad_t, gt, dst=dst)
    dist.gather(pad_n, gn, dst=dst)
    if rank != dst:
        return None
    out = []
    for r in range(world):
        lo, hi = shard_bounds(n_items, world, r)
        t, n = gt[r].cpu().numpy(), gn[r].cpu().numpy()
        out.extend(t[i, : int(n[i])].tolist() for i in range(hi - lo))
    return out


def scatter_audio_blocks(audio: Optional[torch.Tensor], n_items: int, n_samples: int, device: torch.device, block: int, src: int = 0):
    """Block-pipelined scatter: every rank's shard travels as blocks of ``block`` utterances, dealt round-robin over the
    ranks (block 0 of every rank first), so that each rank can start on its first block while the rest is still in flight
    (the NCCL send/recv kernels run on NCCL's own stream, next to the compute stream).  Returns the list of
    (local buffer slice, request | None) in order; ``request.wait()`` before using a slice."""
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_bounds(n_items, world, rank)
    local = torch.empty(hi - lo, n_samples, dtype=torch.float32, device=device)
    parts = []
    n_blocks = -(-(shard_bounds(n_items, world, 0)[1]) // block) if n_items else 0   # rank 0 owns the (joint) largest shard
    if rank == src:
        reqs = []
        for k in range(n_blocks):   # one batched group of sends per round: block k of every rank
            ops = []
            for r in range(world):
                a, b = shard_bounds(n_items, world, r)
                s0, s1 = a + k * block, min(b, a + (k + 1) * block)
                if s0 >= s1:
                    continue
                if r == src:
                    local[s0 - a:s1 - a].copy_(audio[s0:s1])
                else:
                    ops.append(dist.P2POp(dist.isend, audio[s0:s1], r))
            if ops:
                reqs += dist.batch_isend_irecv(ops)
        for b0 in range(0, hi - lo, block):
            parts.append((local[b0:min(hi - lo, b0 + block)], None))
        return parts, reqs
    for b0 in range(0, hi - lo, block):
        sl = local[b0:min(hi - lo, b0 + block)]
        parts.append((sl, dist.batch_isend_irecv([dist.P2POp(dist.irecv, sl, src)])[0]))
    return parts, []


def transcribe_sharded(engine, audio: Optional[torch.Tensor], lens: Optional[torch.Tensor], n_items: int, n_samples: int,
                       max_iters: int = 3, transcribe_fn=None, balance: bool = False, block: int = 0):
    """scatter -> per-rank ``engine.transcribe`` -> gather.  ``transcribe_fn(audio, lens, max_iters)``
    may replace the engine call (used by the CPU tests of the plumbing).  ``balance`` (rank 0 decides, needs ``lens``):
    deal the utterances by length (``balanced_order``) instead of contiguous blocks; results come back in input order.
    ``block`` > 0 (no ``lens``): block-pipelined scatter (``scatter_audio_blocks``), each block is transcribed as soon as it
    has arrived, overlapping the rest of the scatter with compute."""
    device = engine.device if engine is not None else torch.device("cpu")
    fn = transcribe_fn or (lambda x, ln, mi: engine.transcribe(x, ln, mi))
    order = None
    if block > 0 and lens is None and not balance:
        parts, reqs = scatter_audio_blocks(audio.to(device) if audio is not None else None, n_items, n_samples, device, block)
        toks, nts = [], []
        for sl, rq in parts:
            if rq is not None:
                rq.wait()
            r = fn(sl, None, max_iters)
            toks.append(r["tokens"])
            nts.append(r["ntok"])
        for q in reqs:
            q.wait()
        if toks:
            tokens, ntok = torch.cat(toks, 0), torch.cat(nts, 0)
        else:
            U = max_iters * max(1, (n_samples // 160 + 1 - 10) // 8 + 1)
            tokens = torch.zeros(0, U, dtype=torch.int32, device=device)
            ntok = torch.zeros(0, dtype=torch.int32, device=device)
        return gather_tokens(tokens, ntok, n_items)
    if balance and dist.get_rank() == 0 and lens is not None:
        order = balanced_order(lens.tolist(), dist.get_world_size())
        sel = torch.as_tensor(order, dtype=torch.long, device=audio.device)
        audio, lens = audio.index_select(0, sel), lens.index_select(0, sel.to(lens.device))
    a, l = scatter_audio(audio, lens, n_items, n_samples, device)
    if a.shape[0] > 0:
        r = fn(a, l, max_iters)
        tokens, ntok = r["tokens"], r["ntok"]
    else:
        U = max_iters * max(1, (n_samples // 160 + 1 - 10) // 8 + 1)
        tokens = torch.zeros(0, U, dtype=torch.int32, device=device)
        ntok = torch.zeros(0, dtype=torch.int32, device=device)
    out = gather_tokens(tokens, ntok, n_items)
    if out is not None and order is not None:
        restored = [None] * n_items
        for pos, i in enumerate(order):
            restored[i] = out[pos]
        out = restored
    return out
