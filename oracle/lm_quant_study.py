"""TEST INFRASTRUCTURE ONLY -- how far is the as-served LM (qint8 dynamic quantisation) from the fp32 LM this repo fuses?

The reference serves its language model through ``maybe_quantize`` (libreasr/lib/lm.py:86-100,
libreasr/lib/utils.py:197-210: ``torch.quantization.quantize_dynamic(model, {nn.LSTM, nn.Linear}, qint8)``).  This repo
fuses the fp32 LM (policy).  This script runs the imported, unmodified reference on the `cfg2_lm`
fixture inputs twice -- fp32 LM (the fixture's own setting) and the quantised LM -- and records the token sequences of
the quantised run plus how many positions differ (``tests/golden/cfg2_lm_qint8.npz``).

    python -m oracle.lm_quant_study          (authoring container only: needs the reference tree)
"""
import difflib
import os
import sys

import numpy as np
import torch

from . import make_golden as G
from . import ref_shim, weights


def main():
    if not ref_shim.reference_available():
        sys.exit("reference tree missing")
    g = np.load(os.path.join(G.GOLDEN_DIR, "cfg2_lm.npz"))
    name, lm_name = str(g["config"]), str(g["lm_config"])
    cfg, lm_cfg = weights.CONFIGS[name], weights.LM_CONFIGS[lm_name]
    ref = G.build_reference(cfg)
    lm = G.build_reference_lm(lm_cfg)
    import importlib

    U = importlib.import_module("libreasr.lib.utils")
    qlm = U.maybe_quantize(G.build_reference_lm(lm_cfg))          # the reference's own call (lm.py:97)
    audio = weights.make_audio(int(g["n_utt"]), int(g["n_samples"]), int(g["audio_seed"]))
    out = {"config": name, "lm_config": lm_name, "n_utt": int(g["n_utt"])}
    tot, diff = 0, 0
    for b in range(int(g["n_utt"])):
        feats = G.ref_features_offline(torch.from_numpy(audio[b:b + 1]), cfg)[0]
        with torch.no_grad():
            ref.lm = lm
            t32 = [int(t) for t in ref.decode_greedy(feats, max_iters=3)[0]]
            ref.lm = qlm
            tq = [int(t) for t in ref.decode_greedy(feats, max_iters=3)[0]]
        assert t32 == g[f"tokens_{b}"].tolist(), "fp32 run must reproduce the committed fixture"
        sm = difflib.SequenceMatcher(a=t32, b=tq, autojunk=False)
        same = sum(m.size for m in sm.get_matching_blocks())
        d = max(len(t32), len(tq)) - same
        print(f"utt {b}: fp32-LM tokens {len(t32)}, qint8-LM tokens {len(tq)}, non-matching positions {d}")
        out[f"tokens_qint8_{b}"] = np.asarray(tq, dtype=np.int32)
        out[f"n_diff_{b}"] = d
        tot += max(len(t32), len(tq))
        diff += d
    a = weights.make_audio(1, int(g["n_chunks"]) * G.CHUNK, int(g["stream_seed"]))[0]
    a[:G.CHUNK] = 0.0
    rows = [c for c in G.ref_stream_chunks(torch.from_numpy(a), cfg)]
    with torch.no_grad():
        ref.lm = qlm
        ys = [([int(t) for t in y]) for (y, _ys, _r) in ref.transcribe_stream(iter(rows), lambda t: list(t), max_iters=10)]
    sq = ys[-1] if ys else []
    s32 = g["stream_tokens_all"].tolist()
    sm = difflib.SequenceMatcher(a=s32, b=sq, autojunk=False)
    ds = max(len(s32), len(sq)) - sum(m.size for m in sm.get_matching_blocks())
    print(f"stream: fp32-LM tokens {len(s32)}, qint8-LM tokens {len(sq)}, non-matching positions {ds}")
    out["stream_tokens_qint8"] = np.asarray(sq, dtype=np.int32)
    out["n_diff_stream"] = ds
    out["n_diff_total"], out["n_tokens_total"] = diff + ds, tot + max(len(s32), len(sq))
    np.savez_compressed(os.path.join(G.GOLDEN_DIR, "cfg2_lm_qint8.npz"), **out)
    print(f"total: {diff + ds} of {tot + max(len(s32), len(sq))} token positions differ between the fp32 and the qint8 LM")


if __name__ == "__main__":
    main()
