#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 RNN-T inference path (driver contract).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[1]): batch = 32 synthetic 16 kHz utterances of 10 s,
80-mel x 10-stack features, 4x1024 LSTM encoder, 2x1024 GRU predictor, joint 1024,
vocab 2048, greedy decode (max_iters = 3).  One "step" = the whole hot path over one batch:
audio -> log-mel/stack -> LayerNorm -> LSTM stack -> joint/predictor greedy loop -> tokens.
Metric: streaming RTFx = audio-seconds processed / wall-second (whole job, all GPUs).

Prints ONE JSON line (rank 0).  Under torchrun every rank processes its own 32-utterance
batch (weak scaling; utterances are independent, no data-path collective); the only
collective is the NCCL gather of the token arrays inside the end-to-end measurement.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "cfg2"
BATCH = 32
SECONDS = 10.0
MAX_ITERS = 3
N_ROTATE = 8  # distinct input batches cycled through the timed region (8 x 20.5 MB > 126 MB L2)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def model_flops_bytes(cfg, B, T, evals, emitted):
    """Algorithmic work of one step (DESIGN.md section 'Algorithmic bytes and flops')."""
    H, X, J, V, L = cfg.hidden_sz, cfg.feature_sz, cfg.joint_sz, cfg.vocab_sz, cfg.enc_layers
    enc = B * T * (2 * (X + H) * 4 * H + (L - 1) * 2 * (2 * H) * 4 * H)
    joint_enc = B * T * 2 * H * J
    dec = evals * 2 * J * V + emitted * (2 * H * J + 2 * H * 3 * H * (2 * cfg.pred_layers - 1))
    return {"encoder_flops": enc, "joint_enc_flops": joint_enc, "decode_flops": dec}


def run_product(args):
    from libreasr_b200 import synth
    from libreasr_b200.engine import Engine, EngineConfig, tokens_to_lists

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    cfg = synth.CONFIGS[WORKLOAD]
    n = int(SECONDS * cfg.sample_rate)
    ec = EngineConfig(n_mels=cfg.n_mels, n_stack=cfg.n_stack, downsample=cfg.downsample, enc_layers=cfg.enc_layers,
                      pred_layers=cfg.pred_layers, hidden_sz=cfg.hidden_sz, embed_sz=cfg.embed_sz, joint_sz=cfg.joint_sz,
                      vocab_sz=cfg.vocab_sz, gemm_mode=args.gemm_mode)
    eng = Engine(ec, device=dev).load_state_dict(synth.make_state_dict(cfg, 1234))
    eng.reserve(BATCH, n)
    T = eng.num_steps(n)
    U = MAX_ITERS * T

    # synthetic input: N_ROTATE distinct batches per rank, in pinned host memory and in HBM
    base = synth.make_audio(BATCH, n, seed=synth.BENCH_AUDIO_SEED + 1000 * rank)  # see synth.BENCH_AUDIO_SEED
    host = [torch.from_numpy(np.roll(base, 997 * r, axis=1).copy()).pin_memory() for r in range(N_ROTATE)]
    devb = [h.to(dev) for h in host]
    out_host = Engine.alloc_host_outputs(BATCH, U)

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput (value) ----
    for i in range(args.warmup):
        res = eng.transcribe(devb[i % N_ROTATE], max_iters=MAX_ITERS)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.kernel_launches()
    eng.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        res = eng.transcribe(devb[i % N_ROTATE], max_iters=MAX_ITERS)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    stage = eng.stage_times_ms()
    eng.set_profiling(False)
    launches = (eng.kernel_launches() - l0) // max(args.steps, 1)
    clocks = sampler.stop() if rank == 0 else None
    toks = tokens_to_lists(res["tokens"], res["ntok"])
    iters = res["iters"].cpu().numpy()
    evals, emitted = int(iters.sum()), int(sum(len(t) for t in toks))

    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": ms_total / args.steps, "launches_per_step": int(launches),
                              "stage_ms": stage}), flush=True)
        return None

    # ---- end-to-end through the public host-buffer call (e2e) ----
    def gather_tokens():
        if not dist:
            return
        t = out_host["tokens"].to(dev, non_blocking=True)
        gl = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, gl, dst=0)
        if rank == 0:
            torch.stack(gl).cpu()

    for i in range(min(args.warmup, 3)):
        eng.transcribe_host(host[i % N_ROTATE], max_iters=MAX_ITERS, out=out_host)
        gather_tokens()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        eng.transcribe_host(host[i % N_ROTATE], max_iters=MAX_ITERS, out=out_host)
        gather_tokens()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0

    # max over ranks
    if dist:
        tt = torch.tensor([ms_total, e2e_s * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, e2e_s = float(tt[0]), float(tt[1]) / 1e3

    audio_s = BATCH * SECONDS * world
    value = audio_s * args.steps / (ms_total / 1e3)
    e2e_value = audio_s * args.steps / e2e_s

    line = None
    if rank == 0:
        peaks = load_peaks()
        work = model_flops_bytes(cfg, BATCH, T, evals, emitted)
        # per-kernel rooflines from the live CUDA-event stage times of the timed region.  Algorithmic work
        # (DESIGN.md section 4): tensor-bound kernels in flops vs the sustained bf16 peak (the kernels sit inside
        # a long step), the front end in bytes vs the measured HBM copy bandwidth.
        tc = args.gemm_mode == 1
        peak_t = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
        H, X, L = cfg.hidden_sz, cfg.feature_sz, cfg.enc_layers
        hoist_flops = BATCH * T * (2 * X * 4 * H + (L - 1) * 2 * H * 4 * H)
        rec_flops = work["encoder_flops"] - hoist_flops
        rec_ms = max(stage["encoder"] - stage["encoder_input_gemms"], 1e-6)
        fe_bytes = BATCH * (4 * n + 4 * T * X)

        def tens(name, flops, ms, launches):
            a = flops / (ms * 1e-3) / 1e12
            return {"kernel": name, "bound": "tensor", "achieved": round(a, 3), "peak": peak_t, "unit": "TFLOP/s",
                    "frac": round(a / peak_t, 5), "ms_per_step": round(ms, 4), "launches_per_step": launches,
                    "algorithmic_flops_per_step": int(flops)}
        per_kernel = [
            tens("decode_tc_kernel" if tc else "decode_greedy_kernel", work["decode_flops"], stage["decode"], 1),
            tens("lstm_layer_tc_kernel" if tc else "lstm_step_kernel", rec_flops, rec_ms, L if tc else L * T),
            tens("gemm_tc_f16x3_kernel" if tc else "gemm_nt_f32_kernel", hoist_flops, max(stage["encoder_input_gemms"], 1e-6), L),
            {"kernel": "mel_stack_kernel", "bound": "hbm", "achieved": round(fe_bytes / (stage["features"] * 1e-3) / 1e9, 1),
             "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": round(fe_bytes / (stage["features"] * 1e-3) / 1e9 / peaks["hbm_gbs"], 5),
             "ms_per_step": round(stage["features"], 4), "launches_per_step": 1, "algorithmic_bytes_per_step": fe_bytes,
             "note": "FFT-compute-bound, not HBM-bound (DESIGN.md section 4)"},
        ]
        dom = max(per_kernel[:3], key=lambda k: k["ms_per_step"])
        # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/)
        ncu_traffic = {"decode_tc_kernel": {"bytes": 76083456 + 11085824, "source": "profiles/r1_ncu_full_decode_tc_kernel.csv"},
                       "lstm_layer_tc_kernel": {"bytes": 81903360 + 6575104, "source": "profiles/r1_ncu_full_lstm_layer_tc_kernel.csv"}}
        roofline = {"kernel": dom["kernel"], "bound": dom["bound"], "achieved": dom["achieved"], "peak": dom["peak"], "unit": dom["unit"],
                    "frac": dom["frac"], "traffic": ncu_traffic.get(dom["kernel"], {}).get("bytes"),
                    "traffic_source": ncu_traffic.get(dom["kernel"], {}).get("source"), "peak_source": peaks["source"] + ", sustained bf16 GEMM",
                    "ms_per_launch": round(dom["ms_per_step"] / dom["launches_per_step"], 4),
                    "flops_per_launch": int(dom["algorithmic_flops_per_step"] / dom["launches_per_step"]),
                    "note": ("latency-bound at batch 32 (DESIGN.md section 4): fraction = algorithmic flops / sustained bf16 tensor peak; "
                             "3xFP16 split issues 4 fp16 MACs per algorithmic MAC" if tc else "fp32 CUDA-core mode; fraction vs the bf16 tensor peak"),
                    "per_kernel": per_kernel}
        cpu = cpu_baseline(cfg, n, budget_s=args.cpu_budget)
        line = {
            "metric": "streaming RTFx (audio-s/wall-s)", "value": round(value, 1), "unit": "x real-time",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (3xFP16 split on tcgen05, fp32 accumulate)" if args.gemm_mode == 1 else "f32", "data": "synthetic",
            "utterances_per_s": round(BATCH * world * args.steps / (ms_total / 1e3), 1),
            "config": {"workload": "BASELINE.json configs[1]: batch=32 x 10 s synthetic 16 kHz, 80-mel, 4x1024 LSTM encoder, 2x1024 GRU predictor, greedy",
                       "global_batch": BATCH * world, "audio_s_per_utt": SECONDS, "enc_steps": T, "max_iters": MAX_ITERS,
                       "gemm_mode": {0: "fp32_simt", 1: "tc_fp16x3", 2: "tc_bf16"}[args.gemm_mode],
                       "parallelism": f"dp{world} (utterance shards, no data-path collective)",
                       "l2": f"rotating {N_ROTATE} distinct input batches ({N_ROTATE * BATCH * n * 4 / 1e6:.0f} MB) > 126 MB L2",
                       "joint_evals_per_step": evals, "tokens_per_step": emitted,
                       "emission_rate_tok_per_frame": round(emitted / (BATCH * T), 3)},
            "e2e": {"value": round(e2e_value, 1), "unit": "x real-time", "h2d_bytes_per_step": BATCH * n * 4,
                    "d2h_bytes_per_step": BATCH * U * 4 + BATCH * 4 + BATCH * 8, "ms_per_step": round(e2e_s * 1e3 / args.steps, 4),
                    "api": "rnnt_b200_transcribe_host (pinned host audio in, host tokens out)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "stage_ms": {k: round(v, 4) for k, v in stage.items()},
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return line


def pick_cpu_threads(orc, cfg, probe_audio):
    """torch's intra-op pool is slow when oversubscribed on batch-1 GEMVs (128 threads on a
    128-core host are ~500x slower than 8), so the CPU arm uses the fastest of a few pool sizes
    (the reference itself pins 2 threads, inference.py:21)."""
    from oracle import rnnt_oracle as O

    best, best_t = None, None
    ncpu = os.cpu_count() or 1
    for th in sorted({2, 4, 8, 16, 32} & set(range(1, ncpu + 1)) | {min(ncpu, 8)}):
        torch.set_num_threads(th)
        O.transcribe_batch(orc, probe_audio, max_iters=MAX_ITERS)  # warm the pool
        t0 = time.perf_counter()
        O.transcribe_batch(orc, probe_audio, max_iters=MAX_ITERS)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = th, dt
        if dt > 4 * best_t:
            break
    torch.set_num_threads(best)
    return best


def cpu_baseline(cfg, n, budget_s=12.0, max_utts=512, pool=16, seed=100):
    """The reference's CPU path (oracle port: torch fp32, utterance by utterance as the
    reference serves them) on this host's cores, on a bounded sample of the workload."""
    from oracle import rnnt_oracle as O
    from libreasr_b200 import synth

    orc = O.OracleTransducer(cfg, synth.make_state_dict(cfg, 1234))
    audio = synth.make_audio(pool, n, seed=seed)   # the sample cycles over a pool of distinct utterances
    threads = pick_cpu_threads(orc, cfg, audio[:1, : n // 5])
    done, t0 = 0, time.perf_counter()
    while done < max_utts and (done < 1 or time.perf_counter() - t0 < budget_s):
        O.transcribe_batch(orc, audio[done % pool:done % pool + 1], max_iters=MAX_ITERS)
        done += 1
    dt = time.perf_counter() - t0
    # as shipped: the reference pins 2 intra-op threads (inference.py:21); a short sample of the same workload
    torch.set_num_threads(2)
    O.transcribe_batch(orc, audio[:1, : n // 5], max_iters=MAX_ITERS)
    k2, t2 = 0, time.perf_counter()
    while k2 < 3 and (k2 < 1 or time.perf_counter() - t2 < 4.0):
        O.transcribe_batch(orc, audio[k2 % pool:k2 % pool + 1], max_iters=MAX_ITERS)
        k2 += 1
    d2 = time.perf_counter() - t2
    torch.set_num_threads(threads)
    return {"value": round(done * n / cfg.sample_rate / dt, 2), "unit": "x real-time", "cores": threads,
            "as_shipped_2_threads": round(k2 * n / cfg.sample_rate / d2, 2),
            "host_cpus": os.cpu_count(), "kind": "port",
            "sample": f"{done} utterances x {n / cfg.sample_rate:.0f} s of the same workload, sequential (bs=1 as the reference serves), {dt:.1f} s wall"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port,
    all host threads), same metric/config; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import rnnt_oracle as O
    from libreasr_b200 import synth

    cfg = synth.CONFIGS[WORKLOAD]
    n = int(SECONDS * cfg.sample_rate)
    orc = O.OracleTransducer(cfg, synth.make_state_dict(cfg, 1234))
    per_step = 2  # bounded sample: 2 of the 32 utterances per step
    audio = synth.make_audio(per_step * (args.steps + args.warmup), n, seed=100)
    pick_cpu_threads(orc, cfg, audio[:1, : n // 5])
    k = 0
    for _ in range(args.warmup):
        O.transcribe_batch(orc, audio[k:k + per_step], max_iters=MAX_ITERS)
        k += per_step
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.transcribe_batch(orc, audio[k:k + per_step], max_iters=MAX_ITERS)
        k += per_step
    dt = time.perf_counter() - t0
    value = per_step * args.steps * SECONDS / dt
    sample = f"{per_step} of {BATCH} utterances per step, sequential bs=1, torch fp32 CPU"
    print(json.dumps({
        "impl": "reference", "metric": "streaming RTFx (audio-s/wall-s)", "value": round(value, 2), "unit": "x real-time",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt * 1e3 / args.steps, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[1] (bounded sample): 10 s synthetic 16 kHz utterances, 80-mel, 4x1024 LSTM, greedy",
                   "sample": sample},
        "cpu_baseline": {"value": round(value, 2), "unit": "x real-time", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 2), "unit": "x real-time", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gemm-mode", type=int, default=1, help="1 = tcgen05 3xFP16 (default), 0 = fp32 CUDA cores")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--profile", action="store_true",
                    help="for runs under ncu: device-resident steps only, no e2e / CPU legs, warm-up not forced to 3 "
                         "(numbers printed in this mode are not bench values)")
    args = ap.parse_args()
    if not args.profile:
        args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
