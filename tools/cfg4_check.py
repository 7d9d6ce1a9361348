"""Diagnostic: determinism of the decode kernel for shapes where some CTAs sit phases out (cfg4: 6x1536, 96 CTAs)."""
import dataclasses, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libreasr_b200 import synth
from libreasr_b200.engine import Engine, EngineConfig, tokens_to_lists

def mk(cfg, mode=1):
    ec = EngineConfig(n_mels=cfg.n_mels, n_stack=cfg.n_stack, downsample=cfg.downsample, enc_layers=cfg.enc_layers,
                      pred_layers=cfg.pred_layers, hidden_sz=cfg.hidden_sz, embed_sz=cfg.embed_sz, joint_sz=cfg.joint_sz,
                      vocab_sz=cfg.vocab_sz, gemm_mode=mode)
    return Engine(ec).load_state_dict(synth.make_state_dict(cfg, 1234))

base = synth.CONFIGS["cfg4"]
variants = {"cfg4 (J=1024: 64 of 96 CTAs in A; V=2048: 86 in B)": base,
            "J=1536 (all in A), V=2048": dataclasses.replace(base, joint_sz=1536),
            "J=1024, V=2304 (all in B)": dataclasses.replace(base, vocab_sz=2304),
            "J=1536, V=2304 (all CTAs in every phase)": dataclasses.replace(base, joint_sz=1536, vocab_sz=2304),
            "2 encoder layers only (same decode shape)": dataclasses.replace(base, enc_layers=2)}
B, n = int(os.environ.get("BATCH", "32")), int(os.environ.get("SECONDS", "10")) * 16000
audio = torch.from_numpy(synth.make_audio(B, n, seed=4)).cuda()
for name, cfg in variants.items():
    eng = mk(cfg)
    enc, _ = eng.encode(eng.features(audio))
    outs = []
    for i in range(6):
        d = eng.decode_greedy(enc, max_iters=3)
        outs.append(tokens_to_lists(d["tokens"], d["ntok"]))
    ref = outs[0]
    bad = []
    for i in range(1, 6):
        for b in range(B):
            if outs[i][b] != ref[b]:
                k = next((j for j, (x, y) in enumerate(zip(outs[i][b], ref[b])) if x != y), min(len(outs[i][b]), len(ref[b])))
                bad.append((i, b, k, len(ref[b])))
    print(f"{name}: tokens {[sum(len(t) for t in o) for o in outs]} mismatches(run,utt,first_diff,len)={bad[:8]}")
    eng.close()
