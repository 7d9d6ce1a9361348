"""Supplementary: BASELINE configs[2] through the WIRE path -- 64 gRPC clients (`ASR.ASR/TranscribeStream`, loopback), each
streaming synthetic 16 kHz audio in 80 ms `Audio` frames as fast as the server takes them; cfg2 model, reference windowing,
max_iters 10.  Reports audio-seconds served per wall-second and the scheduler's tick statistics.  The clients, the gRPC
message pump and the scheduler all run in this one Python process, so this measures the serving stack, not the kernels."""
import argparse, json, os, sys, threading, time
from concurrent import futures
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import grpc
from libreasr_b200 import synth, serve as S
from libreasr_b200.api import LibreASR
from libreasr_b200.lib.models import Transducer
import torch


class Lang:
    def denumericalize(self, ids):
        return list(ids)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=20.0)
    a = ap.parse_args()
    cfg = synth.CONFIGS["cfg2"]
    m = Transducer(cfg.feature_sz, cfg.embed_sz, cfg.vocab_sz, cfg.hidden_sz, cfg.out_sz, cfg.joint_sz, Lang(),
                   encoder_kwargs={"num_layers": cfg.enc_layers}, predictor_kwargs={"num_layers": cfg.pred_layers})
    m.load_state_dict({k: torch.as_tensor(v) for k, v in synth.make_state_dict(cfg, 1234).items()}, strict=True)
    m = m.to("cuda:0")
    server, servicer, port = S.serve(LibreASR(m), address="127.0.0.1:0", n_streams=a.streams)
    chunk = 1280
    n_chunks = int(a.seconds * 16000) // chunk
    audio = synth.make_audio(a.streams, n_chunks * chunk, seed=1)
    frames = [[S.encode_audio(audio[b, j * chunk:(j + 1) * chunk].astype("<f4").tobytes(), 16000) for j in range(n_chunks)]
              for b in range(a.streams)]
    ident = lambda b: b  # noqa: E731
    chans = [grpc.insecure_channel(f"127.0.0.1:{port}") for _ in range(min(8, a.streams))]
    n_msgs = [0] * a.streams

    def client(b):
        call = chans[b % len(chans)].stream_stream(f"/{S.SERVICE}/TranscribeStream", request_serializer=ident, response_deserializer=ident)
        for r in call(iter(frames[b]), timeout=600):
            n_msgs[b] += 1

    ths = [threading.Thread(target=client, args=(b,)) for b in range(a.streams)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    sch = servicer.scheduler
    print(json.dumps({"metric": "streaming RTFx through gRPC (audio-s/wall-s), %d clients, un-paced" % a.streams,
                      "value": round(a.streams * n_chunks * 0.08 / dt, 1), "wall_s": round(dt, 3), "audio_s_per_stream": round(n_chunks * 0.08, 2),
                      "scheduler_ticks": sch.ticks, "model_ticks": sch.model_ticks, "transcript_messages": int(sum(n_msgs)),
                      "audio_frames": a.streams * n_chunks}))
    server.stop(0)
    servicer.close()


if __name__ == "__main__":
    main()
