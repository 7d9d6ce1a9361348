"""Stage-by-stage numerical diagnosis of the CUDA path against the CPU oracle (run on the
GPU box: `python tools/gpu_diag.py --config tiny cfg2`).  Test infrastructure: prints the
max abs error of every stage and never stops at the first failure, so that one `gpurun`
call yields a full picture.  Results are also written to gpurun_out/diag_<config>.json.
"""
import argparse
import json
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import rnnt_oracle as O  # noqa: E402
from oracle import weights  # noqa: E402
from libreasr_b200.engine import Engine, EngineConfig, tokens_to_lists  # noqa: E402


def engine_for(cfg, sd, gemm_mode=0):
    ec = EngineConfig(n_mels=cfg.n_mels, n_stack=cfg.n_stack, downsample=cfg.downsample, enc_layers=cfg.enc_layers,
                      pred_layers=cfg.pred_layers, hidden_sz=cfg.hidden_sz, embed_sz=cfg.embed_sz,
                      joint_sz=cfg.joint_sz, vocab_sz=cfg.vocab_sz, gemm_mode=gemm_mode)
    return Engine(ec).load_state_dict(sd)


def maxdiff(a, b):
    a = a.detach().cpu().double() if torch.is_tensor(a) else torch.as_tensor(a).double()
    b = b.detach().cpu().double() if torch.is_tensor(b) else torch.as_tensor(b).double()
    return float((a - b).abs().max())


def run(name, n_samples, n_utt, gemm_mode, results):
    cfg = weights.CONFIGS[name]
    sd = weights.make_state_dict(cfg, 1234)
    orc = O.OracleTransducer(cfg, sd)
    dev = torch.device("cuda:0")
    t0 = time.time()
    eng = engine_for(cfg, sd, gemm_mode)
    torch.cuda.synchronize()
    print(f"[{name}] engine ready in {time.time() - t0:.2f}s", flush=True)
    audio = weights.make_audio(n_utt, n_samples, seed=31)
    a_cpu = torch.from_numpy(audio)
    a_dev = a_cpu.to(dev)
    res = {}

    def stage(label, fn):
        try:
            v = fn()
            torch.cuda.synchronize()
            res[label] = v
            print(f"[{name}] {label}: {v}", flush=True)
        except Exception as e:  # noqa: BLE001
            res[label] = "EXC " + repr(e)
            print(f"[{name}] {label}: EXCEPTION {e!r}", flush=True)
            traceback.print_exc()

    with torch.no_grad():
        f_ref = O.features_offline(a_cpu, cfg)
        lm_ref = O.log_mel(a_cpu, cfg)
        stage("features max|d|", lambda: maxdiff(eng.features(a_dev), f_ref))
        stage("logmel max|d|", lambda: maxdiff(eng.logmel(a_dev), lm_ref))
        win = a_cpu[:, 1000:1000 + 3840].contiguous()
        stage("features_stream max|d|", lambda: maxdiff(eng.features_stream(win.to(dev)), O.features_stream_window(win, cfg)[:, 0]))
        lens = torch.tensor([n_samples - 137 * b for b in range(n_utt)], dtype=torch.int32)

        def ragged():
            out = eng.features(a_dev, lens)
            worst = 0.0
            for b in range(n_utt):
                fr = O.features_offline(a_cpu[b:b + 1, : int(lens[b])], cfg)[0]
                worst = max(worst, maxdiff(out[b, : fr.shape[0]], fr), float(out[b, fr.shape[0]:].abs().max()) if fr.shape[0] < out.shape[1] else 0.0)
            return worst
        stage("features ragged max|d|", ragged)

        enc_ref, st_ref = orc.encoder(f_ref, None, "aten")

        def enc_check():
            enc, st = eng.encode(f_ref.to(dev), want_state=True)
            return {"enc": maxdiff(enc, enc_ref), "h": maxdiff(st[0], torch.cat([s[0] for s in st_ref], 0)),
                    "c": maxdiff(st[1], torch.cat([s[1] for s in st_ref], 0))}
        stage("encoder max|d|", enc_check)

        def enc_chunked():
            T = f_ref.shape[1]
            h = T // 2
            e1, s1 = eng.encode(f_ref[:, :h].to(dev), want_state=True)
            e2, s2 = eng.encode(f_ref[:, h:].to(dev), state=s1, want_state=True)
            return maxdiff(torch.cat([e1, e2], 1), enc_ref)
        stage("encoder chunked-state max|d|", enc_chunked)

        def pred_check():
            g = torch.Generator().manual_seed(3)
            toks = torch.randint(1, cfg.vocab_sz, (n_utt, 5), generator=g)
            st_o, st_g, worst = None, None, 0.0
            for j in range(toks.shape[1]):
                o_ref, st_o = orc.predictor(toks[:, j], st_o)
                o_gpu, st_g = eng.predict(toks[:, j].to(dev), st_g)
                worst = max(worst, maxdiff(o_gpu, o_ref), maxdiff(st_g, torch.cat(st_o, 0)))
            return worst
        stage("predictor 5 steps max|d|", pred_check)

        def joint_check():
            hp, _ = orc.predictor(torch.full((n_utt,), 2))
            he = enc_ref[:, 3]
            return maxdiff(eng.joint(hp.to(dev), he.to(dev)), orc.joint(hp, he))
        stage("joint logits max|d|", joint_check)

        # decode on the ORACLE's encoder output (isolates the loop)
        def decode_check():
            r = eng.decode_greedy(enc_ref.to(dev), max_iters=3, trace_cap=3 * enc_ref.shape[1])
            toks = tokens_to_lists(r["tokens"], r["ntok"])
            out = {"tok_match": [], "logp_maxd": 0.0, "nlp_d": 0.0, "iters_match": []}
            for b in range(n_utt):
                ro = orc.decode_greedy(f_ref[b], max_iters=3, impl="aten", keep_logits=True)
                out["tok_match"].append(toks[b] == ro["tokens"])
                ne = len(ro["margins"])
                if toks[b] == ro["tokens"]:
                    out["logp_maxd"] = max(out["logp_maxd"], maxdiff(r["trace"][b, :ne], ro["logp"]))
                out["nlp_d"] = max(out["nlp_d"], abs(float(r["neg_logp"][b]) - ro["neg_log_p"]))
                out["iters_match"].append(r["iters"][b].cpu().tolist() == ro["iters"])
                if toks[b] != ro["tokens"]:
                    k = next((i for i, (x, y) in enumerate(zip(toks[b], ro["tokens"])) if x != y), min(len(toks[b]), len(ro["tokens"])))
                    print(f"   utt {b}: first token mismatch at {k}: gpu {toks[b][k:k+5]} ref {ro['tokens'][k:k+5]} "
                          f"(len {len(toks[b])} vs {len(ro['tokens'])}); min margin {min(ro['margins']):.2e}")
            return out
        stage("decode (oracle enc)", decode_check)

        def full_check():
            r = eng.transcribe(a_dev, max_iters=3)
            toks = tokens_to_lists(r["tokens"], r["ntok"])
            ok = []
            for b in range(n_utt):
                ro = orc.decode_greedy(f_ref[b], max_iters=3, impl="aten")
                ok.append(toks[b] == ro["tokens"])
            return {"tok_match": ok, "ntok": [len(t) for t in toks]}
        stage("transcribe (full path)", full_check)

        def host_check():
            r1 = eng.transcribe(a_dev, max_iters=3)
            r2 = eng.transcribe_host(a_cpu.pin_memory(), max_iters=3)
            return tokens_to_lists(r1["tokens"], r1["ntok"]) == tokens_to_lists(r2["tokens"], r2["ntok"])
        stage("transcribe_host == transcribe", host_check)

        def stream_check():
            from libreasr_b200.api import StreamBatch
            n_chunks = 30
            aud = weights.make_audio(2, n_chunks * 1280, seed=41)
            aud[:, :1280] = 0.0
            sb = StreamBatch(eng, 2, max_iters=10)
            for j in range(n_chunks):
                sb.push(torch.from_numpy(aud[:, j * 1280:(j + 1) * 1280]).to(dev))
            ok = []
            for b in range(2):
                fe = O.StreamFrontend(cfg)
                rows = [fe.push(torch.from_numpy(aud[b:b + 1, j * 1280:(j + 1) * 1280])) for j in range(n_chunks)]
                ys = list(orc.transcribe_stream(iter(rows), max_iters=10))
                ok.append(sb.tokens[b] == (ys[-1][0] if ys else []))
            return {"tok_match": ok, "ntok": [len(t) for t in sb.tokens]}
        stage("stream 2x30 chunks", stream_check)

        def timing():
            eng.set_profiling(True)
            for _ in range(2):
                eng.transcribe(a_dev, max_iters=3)
            torch.cuda.synchronize()
            t = eng.stage_times_ms()
            eng.set_profiling(False)
            return t
        stage("stage times ms", timing)
    results[name] = res
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", nargs="+", default=["tiny"])
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--utts", type=int, default=3)
    ap.add_argument("--gemm-mode", type=int, default=1)
    args = ap.parse_args()
    print("device:", torch.cuda.get_device_name(0), flush=True)
    results = {}
    for name in args.config:
        try:
            run(name, int(args.seconds * 16000), args.utts, args.gemm_mode, results)
        except Exception as e:  # noqa: BLE001
            print(f"[{name}] FATAL {e!r}")
            traceback.print_exc()
            results[name] = "FATAL " + repr(e)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"diag_mode{args.gemm_mode}.json"), "w") as f:
        json.dump(results, f, indent=1, default=str)


if __name__ == "__main__":
    main()
