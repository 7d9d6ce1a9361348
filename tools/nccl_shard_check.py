"""Run under torchrun (NCCL, one rank per GPU): N-GPU sharded transcription == 1-GPU transcription.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/nccl_shard_check.py --utterances 70

Rank 0 holds `--n` utterances (cfg2 model, ragged lengths optional); they go through
libreasr_b200.parallel.transcribe_sharded three ways -- contiguous blocks, length-balanced dealing and the
block-pipelined scatter -- and must come back identical (token for token, in input order) to rank 0 transcribing the
whole batch alone.  Prints one JSON line on rank 0; exit code 1 on any mismatch."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from libreasr_b200 import parallel, synth  # noqa: E402
from libreasr_b200.engine import Engine, EngineConfig, tokens_to_lists  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utterances", type=int, default=70)
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--config", default="cfg2")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = synth.CONFIGS[a.config]
    ec = EngineConfig(n_mels=cfg.n_mels, n_stack=cfg.n_stack, downsample=cfg.downsample, enc_layers=cfg.enc_layers,
                      pred_layers=cfg.pred_layers, hidden_sz=cfg.hidden_sz, embed_sz=cfg.embed_sz, joint_sz=cfg.joint_sz,
                      vocab_sz=cfg.vocab_sz)
    eng = Engine(ec, device=dev).load_state_dict(synth.make_state_dict(cfg, 1234))
    n = int(a.seconds * 16000)
    audio = lens = want = None
    if rank == 0:
        audio = torch.from_numpy(synth.make_audio(a.utterances, n, seed=151)).to(dev)
        lens = torch.tensor([n - 800 * (i % 23) for i in range(a.utterances)], dtype=torch.int32, device=dev)
        r = eng.transcribe(audio, lens, 3)
        want = tokens_to_lists(r["tokens"], r["ntok"])
        r2 = eng.transcribe(audio, None, 3)
        want_full = tokens_to_lists(r2["tokens"], r2["ntok"])
    res = {}
    got = parallel.transcribe_sharded(eng, audio, lens, a.utterances, n, max_iters=3)
    if rank == 0:
        res["contiguous_ragged"] = got == want
    got = parallel.transcribe_sharded(eng, audio, lens, a.utterances, n, max_iters=3, balance=True)
    if rank == 0:
        res["balanced_ragged"] = got == want
    got = parallel.transcribe_sharded(eng, audio, None, a.utterances, n, max_iters=3, block=16)
    if rank == 0:
        res["block_pipelined"] = got == want_full
    ok = True
    if rank == 0:
        ok = all(res.values())
        print(json.dumps({"world": world, "n_utterances": a.utterances, "tokens_total": sum(len(t) for t in want), "checks": res, "ok": ok}), flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
