"""BASELINE.json configs[2]: 64 concurrent synthetic streams, 80 ms chunks, cfg2 model, greedy
(max_iters = 10), reference windowing (3-chunk window, Buffer(2) -> encoder T = 2 every 160 ms).
Prints one JSON line: streaming RTFx = audio-seconds processed / wall-second over all streams,
plus the per-tick latency.  Inputs are generated on the host and copied per tick (H2D inside the
timed region), results (new tokens) are read back every tick like a server would."""
import argparse, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libreasr_b200 import synth
from libreasr_b200.api import StreamBatch
from libreasr_b200.engine import Engine, EngineConfig


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=20.0)
    ap.add_argument("--gemm-mode", type=int, default=1)
    ap.add_argument("--lm", default="", help="fuse a language model: en (4x768) or default (6x1024)")
    a = ap.parse_args()
    cfg = synth.CONFIGS["cfg2"]
    ec = EngineConfig(n_mels=cfg.n_mels, n_stack=cfg.n_stack, downsample=cfg.downsample, enc_layers=cfg.enc_layers,
                      pred_layers=cfg.pred_layers, hidden_sz=cfg.hidden_sz, embed_sz=cfg.embed_sz, joint_sz=cfg.joint_sz,
                      vocab_sz=cfg.vocab_sz, gemm_mode=a.gemm_mode)
    lsd = None
    if a.lm:
        import dataclasses
        lc = synth.LM_CONFIGS[a.lm]
        ec = dataclasses.replace(ec, lm_layers=lc.num_layers, lm_hidden_sz=lc.hidden_sz, lm_embed_sz=lc.embed_sz)
        lsd = synth.make_lm_state_dict(lc, 4321)
    eng = Engine(ec).load_state_dict(synth.make_state_dict(cfg, 1234), lm_state_dict=lsd)
    chunk = 1280
    n_chunks = int(a.seconds * 16000) // chunk
    audio = synth.make_audio(a.streams, n_chunks * chunk, seed=1)   # seed 1: BASELINE.md config 3
    # [n_chunks, B, chunk] pinned: each tick's slab is one contiguous host block (what a server's receive buffer is)
    host = torch.from_numpy(audio.reshape(a.streams, n_chunks, chunk).transpose(1, 0, 2).copy()).pin_memory()
    sb = StreamBatch(eng, a.streams, max_iters=10)
    warm = 10
    lat = []
    for j in range(n_chunks):
        if j == warm:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        t1 = time.perf_counter()
        new = sb.push(host[j])
        if new is not None and j >= warm:
            lat.append(time.perf_counter() - t1)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    audio_s = a.streams * (n_chunks - warm) * chunk / 16000.0
    print(json.dumps({"metric": "streaming RTFx (audio-s/wall-s), 64 concurrent 80 ms-chunk streams", "value": round(audio_s / dt, 1),
                      "streams": a.streams, "audio_s_per_stream": round((n_chunks - warm) * 0.08, 2), "wall_s": round(dt, 3),
                      "model_tick_ms_mean": round(1e3 * float(np.mean(lat)), 3), "model_tick_ms_p50": round(1e3 * float(np.median(lat)), 3), "model_tick_ms_p99": round(1e3 * float(np.quantile(lat, 0.99)), 3),
                      "ticks_with_model_step": len(lat), "tokens_total": int(sum(len(t) for t in sb.tokens)),
                      "gemm_mode": a.gemm_mode, "lm": a.lm or None, "note": "rnnt_b200_stream_push per tick (host chunks in, host tokens out); one model step (encoder T=2 -> greedy decode) every second 80 ms chunk"}))


if __name__ == "__main__":
    main()
