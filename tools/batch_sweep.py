"""Supplementary measurement (not the bench line): offline RTFx of the cfg2 model as a function of
the utterance batch on ONE GPU, inputs resident in HBM, 10 s utterances, CUDA-event timing.
bench.py's workload is B = 32 (BASELINE.json configs[1]); this shows how the per-step latency
floor of the recurrences amortises when a server batches more utterances per call
(encode runs in chunks of <= 128 utterances, decode in chunks of <= 64: rnnt_b200_transcribe)."""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libreasr_b200 import synth
from libreasr_b200.engine import Engine, EngineConfig


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, nargs="+", default=[1, 8, 32, 64, 128, 256])
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--gemm-mode", type=int, default=1)
    ap.add_argument("--config", default="cfg2", help="cfg2 (BASELINE configs[1]) or cfg4 (configs[3] shape: 6x1536, 15 s)")
    a = ap.parse_args()
    cfg = synth.CONFIGS[a.config]
    ec = EngineConfig(n_mels=cfg.n_mels, n_stack=cfg.n_stack, downsample=cfg.downsample, enc_layers=cfg.enc_layers,
                      pred_layers=cfg.pred_layers, hidden_sz=cfg.hidden_sz, embed_sz=cfg.embed_sz, joint_sz=cfg.joint_sz,
                      vocab_sz=cfg.vocab_sz, gemm_mode=a.gemm_mode)
    eng = Engine(ec).load_state_dict(synth.make_state_dict(cfg, 1234))
    n = int(a.seconds * 16000)
    rows = []
    for B in a.batches:
        # two rotating input sets; at B >= 64 one set (>= 41 MB with its workspaces) plus the 170 MB of weights
        # already exceeds what stays in L2 between steps
        sets = [torch.from_numpy(synth.make_audio(B, n, seed=synth.BENCH_AUDIO_SEED + i)).to(eng.device) for i in range(2)]
        for i in range(3):
            out = eng.transcribe(sets[i & 1], max_iters=3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            out = eng.transcribe(sets[i & 1], max_iters=3)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        rows.append({"batch": B, "ms_per_step": round(ms, 3), "rtfx": round(B * a.seconds / (ms * 1e-3), 1),
                     "tokens": int(out["ntok"].sum().item())})
    print(json.dumps({"metric": "offline greedy RTFx vs utterance batch (%s, %.0f s utterances, 1 GPU, device-resident inputs)" % (a.config, a.seconds),
                      "gemm_mode": a.gemm_mode, "rows": rows}))


if __name__ == "__main__":
    main()
