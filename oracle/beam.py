"""TEST INFRASTRUCTURE ONLY -- CPU definition of the batched RNN-T beam search (BASELINE.json configs[3], [4]).

PARITY UNPINNED IN THE REFERENCE: iceychris/LibreASR has no beam search (``grep -i beam`` finds only an unused
``PriorityQueue`` import, libreasr/lib/models.py:8).  SURVEY.md section 8c therefore asks for a self-pinned restatement
that uses the reference's own ``Predictor`` / ``Joint`` modules (models.py:116-187).  This file IS that definition; the
fixtures ``tests/golden/*_beam.npz`` are produced by running it on the imported reference modules
(``oracle/make_golden.py``), and the same function runs on the oracle's restated modules in the tests.

Algorithm (a breadth-first, iteration-capped variant of Graves 2012 / "time-synchronous" RNN-T beam search that keeps
the reference's ``max_iters`` rule, models.py:369: at most ``max_iters`` symbols per encoder frame, and a symbol emitted
on the last allowed iteration is kept):

    hyps = [ (tokens = [], score = 0, predictor state after BOS) ]                  # at most W
    for every encoder frame t:
        A = hyps; leave = []
        for it in 0 .. max_iters - 1:
            for h in A:   lp = log_softmax(joint(h.g, enc[t]))
                          leave += [ h with score + lp[blank] ]                       # blank: h moves on to frame t + 1
                          cand  += [ (h, k, h.score + lp[k]) for k in the W best non-blank tokens ]
            A = the W best cand (score desc; ties: lower parent slot, then lower token id), each advanced through the predictor
        leave += A                                                                    # iteration cap: they move on as well
        merge entries of `leave` with the same token sequence (score = logaddexp, state of the higher-scoring one;
            ties: the earlier entry), keep the W best (score desc; ties: earlier entry)  -> hyps
    result = best of hyps (score desc; ties: earlier)

Every cut (top-W) reports its margin so that fixtures can be chosen away from ties.
"""
import math
from dataclasses import dataclass, field
from typing import Callable, List, Tuple

import torch
import torch.nn.functional as F

NEG = float("-inf")


@dataclass
class Hyp:
    tokens: Tuple[int, ...]
    score: float
    state: object          # predictor state (opaque)
    g: torch.Tensor        # predictor output [1, H]
    order: int = 0         # creation order inside the current frame (tie-breaking)


@dataclass
class BeamResult:
    tokens: List[int]
    score: float
    nbest: List[Tuple[List[int], float]]
    min_margin: float = math.inf
    cut_margins: List[float] = field(default_factory=list)


def _topk_stable(items, k, key):
    """k best by key desc, ties by original position (stable)."""
    idx = sorted(range(len(items)), key=lambda i: (-key(items[i]), i))
    return [items[i] for i in idx[:k]], (key(items[idx[k - 1]]) - key(items[idx[k]]) if len(idx) > k else math.inf)


def beam_search(enc: torch.Tensor, predict: Callable, joint_logits: Callable, bos: int, blank: int, width: int, max_iters: int = 3) -> BeamResult:
    """enc [T, H] encoder output of ONE utterance.  predict(token:int, state|None) -> (g [1,H], state);
    joint_logits(g [1,H], h_enc [1,H]) -> logits [1, V]."""
    res = BeamResult([], 0.0, [])
    with torch.no_grad():
        g0, s0 = predict(bos, None)
        hyps = [Hyp((), 0.0, s0, g0)]
        for t in range(enc.shape[0]):
            h_enc = enc[t][None]
            A, leave = hyps, []
            for it in range(max_iters):
                cand = []
                for slot, h in enumerate(A):
                    lp = F.log_softmax(joint_logits(h.g, h_enc), dim=-1)[0]
                    leave.append(Hyp(h.tokens, h.score + float(lp[blank]), h.state, h.g))
                    nb = lp.clone()
                    nb[blank] = NEG
                    vals, idx = torch.sort(nb, descending=True, stable=True)     # ties: lower token id first
                    # (the per-parent cut cannot change the result: the global top-W below is a subset of the per-parent top-W's)
                    for r in range(min(width, nb.numel() - 1)):
                        cand.append((slot, int(idx[r]), h.score + float(vals[r])))
                best, m = _topk_stable(cand, width, key=lambda c: c[2])      # cand order = (parent slot, token rank)
                res.cut_margins.append(m)
                nxt = []
                for slot, k, sc in best:
                    g, st = predict(k, A[slot].state)
                    nxt.append(Hyp(A[slot].tokens + (k,), sc, st, g))
                A = nxt
            leave += A
            merged = {}
            for e in leave:                       # insertion order = creation order
                if e.tokens in merged:
                    o = merged[e.tokens]
                    hi, lo = (o, e) if o.score >= e.score else (e, o)
                    tot = hi.score + math.log1p(math.exp(lo.score - hi.score))
                    merged[e.tokens] = Hyp(hi.tokens, tot, hi.state, hi.g)
                else:
                    merged[e.tokens] = e
            hyps, m = _topk_stable(list(merged.values()), width, key=lambda h: h.score)
            res.cut_margins.append(m)
        best, _ = _topk_stable(hyps, 1, key=lambda h: h.score)
    res.tokens, res.score = list(best[0].tokens), best[0].score
    res.nbest = [(list(h.tokens), h.score) for h in hyps]
    finite = [m for m in res.cut_margins if math.isfinite(m)]
    res.min_margin = min(finite) if finite else math.inf
    return res


def oracle_callables(orc):
    """(predict, joint_logits) on the oracle's restated Predictor / Joint (rnnt_oracle.py)."""
    def predict(tok, state):
        return orc.predictor(torch.tensor([tok]), state)

    def joint_logits(g, h_enc):
        return orc.joint(g, h_enc)
    return predict, joint_logits


def reference_callables(ref):
    """(predict, joint_logits) on the imported, unmodified reference modules (models.py:116-187)."""
    def predict(tok, state):
        out, st = ref.predictor(torch.LongTensor([[tok]]), state)     # [1,1,H]
        return out[:, 0], st

    def joint_logits(g, h_enc):
        return ref.joint(g, h_enc)
    return predict, joint_logits
