"""Supplementary measurement for SURVEY section 8 row a16 (LM shallow fusion): the bench workload
(cfg2 model, 32 x 10 s utterances, greedy max_iters 3) with the shipped LM shape attached
(4 x 768 LSTM, tied 2048 x 768 embedding; config/testing.yaml:306-313), device-resident inputs,
CUDA-event timing.  With an LM the decode loop runs in the cooperative fp32 kernel (decode.cu);
the encoder stays on the tcgen05 path."""
import argparse, dataclasses, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libreasr_b200 import synth
from libreasr_b200.engine import Engine, EngineConfig


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--lm", default="en", help="en = shipped override 4x768 (testing.yaml:306-313), default = 6x1024 (testing.yaml:293-299)")
    a = ap.parse_args()
    cfg, lc = synth.CONFIGS["cfg2"], synth.LM_CONFIGS[a.lm]
    base = EngineConfig(n_mels=cfg.n_mels, n_stack=cfg.n_stack, downsample=cfg.downsample, enc_layers=cfg.enc_layers,
                        pred_layers=cfg.pred_layers, hidden_sz=cfg.hidden_sz, embed_sz=cfg.embed_sz, joint_sz=cfg.joint_sz,
                        vocab_sz=cfg.vocab_sz)
    sd = synth.make_state_dict(cfg, 1234)
    n = int(a.seconds * 16000)
    sets = [torch.from_numpy(synth.make_audio(a.batch, n, seed=synth.BENCH_AUDIO_SEED + i)).cuda() for i in range(2)]
    out = {}
    for tag, ec, lsd in (("no_lm", base, None),
                         ("lm_%dx%d" % (lc.num_layers, lc.hidden_sz), dataclasses.replace(base, lm_layers=lc.num_layers, lm_hidden_sz=lc.hidden_sz, lm_embed_sz=lc.embed_sz),
                          synth.make_lm_state_dict(lc, 4321))):
        eng = Engine(ec).load_state_dict(sd, lm_state_dict=lsd)
        eng.lib.rnnt_b200_set_profiling(eng._h, 1)
        for i in range(3):
            r = eng.transcribe(sets[i & 1], max_iters=3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            r = eng.transcribe(sets[i & 1], max_iters=3)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        out[tag] = {"ms_per_step": round(ms, 3), "rtfx": round(a.batch * a.seconds / (ms * 1e-3), 1),
                    "tokens": int(r["ntok"].sum().item()), "stage_ms": {k: round(v, 3) for k, v in eng.stage_times_ms().items()}}
        eng.close()
    print(json.dumps({"metric": "offline RTFx with and without LM shallow fusion (cfg2, %d x %.0f s, 1 GPU)" % (a.batch, a.seconds), **out}))


if __name__ == "__main__":
    main()
