"""Soak test of the full-size hot path (tuning / debugging aid): many back-to-back transcribe calls on rotating batches with a
watchdog that reports the call, batch and stage a stall happened in.  STRESS_LIB=<path of a librnnt_b200 build> picks the library."""
import faulthandler, os, sys, threading, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libreasr_b200 import _capi
if os.environ.get("STRESS_LIB"):
    _capi.LIB_PATH = os.environ["STRESS_LIB"]
from libreasr_b200 import synth
from libreasr_b200.engine import Engine, EngineConfig

calls = int(os.environ.get("STRESS_CALLS", "600")); staged = os.environ.get("STRESS_STAGED", "0") == "1"
B, n, NR = 32, 160000, 8
cfg = synth.CONFIGS["cfg2"]
ec = EngineConfig(n_mels=cfg.n_mels, n_stack=cfg.n_stack, downsample=cfg.downsample, enc_layers=cfg.enc_layers, pred_layers=cfg.pred_layers,
                  hidden_sz=cfg.hidden_sz, embed_sz=cfg.embed_sz, joint_sz=cfg.joint_sz, vocab_sz=cfg.vocab_sz)
eng = Engine(ec).load_state_dict(synth.make_state_dict(cfg, 1234))
eng.reserve(B, n)
base = synth.make_audio(B, n, seed=synth.BENCH_AUDIO_SEED)
dev = [torch.from_numpy(np.roll(base, 997 * r, axis=1).copy()).cuda() for r in range(NR)]
progress = [time.time(), "start"]

def dog():
    while True:
        time.sleep(0.5)
        if time.time() - progress[0] > 6.0:
            print("STALL at", progress[1], flush=True)
            os._exit(3)
threading.Thread(target=dog, daemon=True).start()
want = {}
t0 = time.time()
for i in range(calls):
    r = i % NR
    if staged:
        progress[:] = [time.time(), f"call {i} batch {r} features"]
        f = eng.features(dev[r]); torch.cuda.synchronize()
        progress[:] = [time.time(), f"call {i} batch {r} encode"]
        e = eng.encode(f); torch.cuda.synchronize()
        progress[:] = [time.time(), f"call {i} batch {r} decode"]
        res = eng.decode_greedy(e if not isinstance(e, tuple) else e[0]); torch.cuda.synchronize()
    else:
        progress[:] = [time.time(), f"call {i} batch {r}"]
        res = eng.transcribe(dev[r]); torch.cuda.synchronize()
    nt = res["ntok"].tolist()
    if r in want:
        assert nt == want[r], f"call {i} batch {r}: token counts changed"
    want[r] = nt
print(f"OK {calls} calls, {(time.time()-t0)*1e3/calls:.3f} ms/call, lib={os.path.basename(_capi.LIB_PATH)} staged={staged}", flush=True)
