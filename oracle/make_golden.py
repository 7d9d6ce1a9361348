"""TEST INFRASTRUCTURE ONLY -- generate ``tests/golden/*.npz`` from the REAL reference.

Run in the authoring container (needs ``/root/reference``):

    python -m oracle.make_golden

What runs here is the reference's own code, imported unmodified through
``oracle/ref_shim.py``: ``Transducer.decode_greedy`` (models.py:369-455),
``Transducer.transcribe_stream`` (models.py:457-577), ``Encoder`` /
``Predictor`` / ``Joint`` ``forward`` (models.py:105-113, 181-187, 132-140).
The reference's feature transforms cannot be imported (transforms.py needs
fastai2 / fastcore / fastai2_audio), so the feature side of every fixture is
produced with the very ``torchaudio.transforms.MelSpectrogram`` call the
reference makes (transforms.py:290-296) followed by the literal tensor ops of
``TransformTime`` / ``StreamPostprocess`` / ``StackDownsample`` / ``Buffer``
(transforms.py:311-323, 335-342, 436-441, 463-471) and the serving window of
api-server.py:26,83-115.

Weights and audio are synthetic and seed-defined (``oracle/weights.py``), so the
fixtures only store seeds plus the reference's OUTPUTS.
"""
import os
import sys
import warnings

import numpy as np
import torch
import torchaudio

from . import ref_shim, weights

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
WEIGHT_SEED = 1234
CHUNK = 1280  # 80 ms @ 16 kHz (api-client.py:14)


def build_reference(cfg):
    M = ref_shim.import_reference_models()
    sd = weights.make_state_dict(cfg, WEIGHT_SEED)
    ref = M.Transducer.from_config(ref_shim.reference_conf(cfg), ref_shim.FakeLang())
    ref.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=True)
    ref.eval()
    return ref


def ref_logmel(audio, cfg):
    """transforms.py:290-296 + 311-323 with deltas=0."""
    op = torchaudio.transforms.MelSpectrogram(
        sample_rate=cfg.sample_rate, win_length=cfg.win_length, hop_length=cfg.hop_length,
        n_fft=cfg.n_fft, n_mels=cfg.n_mels,
    )
    with torch.no_grad():
        res = torch.log(op(audio) + 1e-6)
    return res.permute(0, 2, 1)


def ref_stack(t, cfg):
    """transforms.py:436-441."""
    uf = t.unfold(-2, cfg.n_stack, cfg.downsample).contiguous()
    return uf.view(uf.size(0), uf.size(1), -1).contiguous()


def ref_features_offline(audio, cfg):
    return ref_stack(ref_logmel(audio, cfg), cfg).unsqueeze(-1)  # FixDimensions, transforms.py:450-452


def ref_stream_chunks(audio_1d, cfg):
    """api-server.py:83-115 windowing + stream transform pipeline incl. Buffer(n_buffer=2):
    yields [2, X, 1] tensors or None, one per 80 ms chunk."""
    frames, saved = [], []
    n_chunks = audio_1d.shape[0] // CHUNK
    for j in range(n_chunks):
        frames.append(audio_1d[None, j * CHUNK:(j + 1) * CHUNK])
        if len(frames) < 3:
            yield None  # server: `continue` (nothing reaches the model)
            continue
        aud = torch.cat(frames, dim=1)
        frames.pop(0)
        sp = ref_logmel(aud, cfg)
        a = sp.shape[1] // 3 + 1  # StreamPostprocess, transforms.py:335-342
        sp = sp[:, a:, :][:, : cfg.n_stack, :]
        saved.append(ref_stack(sp, cfg).unsqueeze(-1))
        if len(saved) == 2:  # Buffer, transforms.py:463-471
            catted = torch.cat(saved, dim=1)
            saved.clear()
            yield catted[0]
        else:
            yield None


def offline_fixture(name, n_utt, n_samples, audio_seed, full=False, enc_stride=8, n_logp=8):
    cfg = weights.CONFIGS[name]
    ref = build_reference(cfg)
    audio = weights.make_audio(n_utt, n_samples, audio_seed)
    out = {
        "config": name, "weight_seed": WEIGHT_SEED, "audio_seed": audio_seed,
        "n_utt": n_utt, "n_samples": n_samples, "max_iters": 3,
    }
    for b in range(n_utt):
        feats = ref_features_offline(torch.from_numpy(audio[b:b + 1]), cfg)[0]  # [T, X, 1]
        with torch.no_grad():
            toks, nlp, metrics, extra = ref.decode_greedy(feats, max_iters=3)
            enc = ref.encoder(feats[None])[0]
        logp = torch.stack([o.reshape(-1) for o in extra["outs"]])
        top2 = torch.topk(logp, 2, dim=-1).values
        out[f"tokens_{b}"] = np.asarray(toks, dtype=np.int32)
        out[f"neg_log_p_{b}"] = np.float64(nlp)
        out[f"iters_{b}"] = np.asarray(extra["iters"], dtype=np.int32)
        out[f"alignment_score_{b}"] = np.float64(metrics["alignment_score"])
        out[f"margins_{b}"] = (top2[:, 0] - top2[:, 1]).numpy().astype(np.float32)
        out[f"argmax_{b}"] = logp.argmax(-1).numpy().astype(np.int32)
        out[f"maxlogp_{b}"] = logp.max(-1).values.numpy().astype(np.float32)
        if full:
            out[f"feats_{b}"] = feats[..., 0].numpy()
            out[f"enc_{b}"] = enc.numpy()
            out[f"logp_{b}"] = logp.numpy()
        else:
            out[f"feats_sub_{b}"] = feats[::enc_stride, ::7, 0].numpy()
            out[f"enc_sub_{b}"] = enc[::enc_stride, ::16].numpy()
            out[f"logp_first_{b}"] = logp[:n_logp].numpy()
        print(f"[{name}] utt {b}: T={feats.shape[0]} evals={logp.shape[0]} tokens={len(toks)} "
              f"min margin={float(out[f'margins_{b}'].min()):.2e}")
    out["enc_stride"], out["n_logp"] = enc_stride, n_logp
    return out


def stream_fixture(name, n_chunks, audio_seed, lead_zero_chunks=1, reset_after=()):
    """api-client.py:32-47 sends one leading zero chunk, then the audio in 80 ms chunks.  ``reset_after``: yield indices
    after which the consumer calls the yielded reset_fn, as the server's silence logic does (api-server.py:133-135)."""
    cfg = weights.CONFIGS[name]
    ref = build_reference(cfg)
    audio = weights.make_audio(1, n_chunks * CHUNK, audio_seed)[0]
    audio[: lead_zero_chunks * CHUNK] = 0.0
    a = torch.from_numpy(audio)
    rows = [c for c in ref_stream_chunks(a, cfg)]
    feats = [c[..., 0].numpy() for c in rows if c is not None]
    with torch.no_grad():
        yields = []
        for i, (y, ys, reset_fn) in enumerate(ref.transcribe_stream(iter(rows), lambda t: list(t), max_iters=10)):
            yields.append((list(y), list(ys)))
            if i in reset_after:
                reset_fn()
    out = {
        "reset_after": np.asarray(sorted(reset_after), dtype=np.int32),
        "config": name, "weight_seed": WEIGHT_SEED, "audio_seed": audio_seed, "n_chunks": n_chunks,
        "lead_zero_chunks": lead_zero_chunks, "max_iters": 10,
        "n_yields": len(yields),
        "tokens_all": np.asarray(yields[-1][0] if yields else [], dtype=np.int32),
        "chunk_counts": np.asarray([len(ys) for _, ys in yields], dtype=np.int32),
        "feats": np.stack(feats) if name == "tiny" else np.stack(feats)[:, :, ::7],
    }
    print(f"[{name} stream] chunks={n_chunks} yields={len(yields)} tokens={len(out['tokens_all'])}")
    return out


def modules_fixture(name="tiny", T=12, N=3):
    """Direct ``Encoder`` / ``Predictor`` / ``Joint`` calls with explicit state (section 8b)."""
    cfg = weights.CONFIGS[name]
    ref = build_reference(cfg)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(N, 2 * T, cfg.feature_sz, 1, generator=g)
    toks = torch.randint(1, cfg.vocab_sz, (N, 4), generator=g)
    with torch.no_grad():
        e1, s1 = ref.encoder(x[:, :T], return_state=True)
        e2, s2 = ref.encoder(x[:, T:], state=s1, return_state=True)
        efull = ref.encoder(x)
        p_out, p_state, outs = None, None, []
        for j in range(toks.shape[1]):
            p_out, p_state = ref.predictor(toks[:, j:j + 1], state=p_state)
            outs.append(p_out[:, 0])
        jl = ref.joint(outs[-1], e2[:, -1])
    return {
        "config": name, "weight_seed": WEIGHT_SEED, "x": x[..., 0].numpy(), "tokens": toks.numpy().astype(np.int32),
        "enc_first": e1.numpy(), "enc_second": e2.numpy(), "enc_full": efull.numpy(),
        "enc_h": np.stack([s[0][0].numpy() for s in s2]), "enc_c": np.stack([s[1][0].numpy() for s in s2]),
        "pred_outs": torch.stack(outs, 1).numpy(), "pred_h": np.stack([s[0].numpy() for s in p_state]),
        "joint_logits": jl.numpy(),
    }


def demo_fixture(name="ref"):
    """BASELINE.json configs[0]: the reference's demo utterance (api-client.py:13; 16 kHz mono 16-bit FLAC, 330,400
    samples) through the reference-shape model, greedy (max_iters 3).  No pretrained weights exist offline, so this is
    plumbing parity on real speech with the seeded synthetic weights.  The PCM is stored (int16) because the
    reference tree does not exist on the GPU box."""
    from . import flac

    cfg = weights.CONFIGS[name]
    ref = build_reference(cfg)
    pcm, sr, bps = flac.decode(os.path.join(ref_shim.REFERENCE_ROOT, "demo", "3729-6852-0035.flac"))
    assert (sr, bps, pcm.shape) == (16000, 16, (1, 330400))
    audio = torch.from_numpy(pcm.astype(np.float32) / 32768.0)          # torchaudio.load normalisation
    feats = ref_features_offline(audio, cfg)[0]
    with torch.no_grad():
        toks, nlp, metrics, extra = ref.decode_greedy(feats, max_iters=3)
    logp = torch.stack([o.reshape(-1) for o in extra["outs"]])
    top2 = torch.topk(logp, 2, dim=-1).values
    margins = (top2[:, 0] - top2[:, 1]).numpy()
    print(f"[demo/{name}] T={feats.shape[0]} evals={logp.shape[0]} tokens={len(toks)} min margin={margins.min():.2e}")
    return {"config": name, "weight_seed": WEIGHT_SEED, "pcm16": pcm[0].astype(np.int16), "max_iters": 3,
            "tokens": np.asarray(toks, dtype=np.int32), "iters": np.asarray(extra["iters"], dtype=np.int32),
            "neg_log_p": np.float64(nlp), "alignment_score": np.float64(metrics["alignment_score"]),
            "maxlogp": logp.max(-1).values.numpy().astype(np.float32), "margins": margins.astype(np.float32)}


def beam_fixture(name, widths, n_utt, n_samples, audio_seed, max_iters=3):
    """Beam search (BASELINE configs[3], [4]) -- PARITY UNPINNED IN THE REFERENCE (it has none): the algorithm is defined by
    oracle/beam.py; what is pinned here is that definition evaluated on the imported reference's own Encoder / Predictor /
    Joint modules (models.py:105-187)."""
    from . import beam

    cfg = weights.CONFIGS[name]
    ref = build_reference(cfg)
    audio = weights.make_audio(n_utt, n_samples, audio_seed)
    out = {"config": name, "weight_seed": WEIGHT_SEED, "audio_seed": audio_seed, "n_utt": n_utt, "n_samples": n_samples,
           "max_iters": max_iters, "widths": np.asarray(widths, dtype=np.int32)}
    predict, joint_logits = beam.reference_callables(ref)
    for b in range(n_utt):
        feats = ref_features_offline(torch.from_numpy(audio[b:b + 1]), cfg)[0]
        with torch.no_grad():
            enc = ref.encoder(feats[None])[0]
            greedy = ref.decode_greedy(feats, max_iters=max_iters)[0]
        for W in widths:
            r = beam.beam_search(enc, predict, joint_logits, ref.bos, ref.blank, W, max_iters)
            out[f"tokens_w{W}_{b}"] = np.asarray(r.tokens, dtype=np.int32)
            out[f"score_w{W}_{b}"] = np.float64(r.score)
            out[f"nbest_scores_w{W}_{b}"] = np.asarray([s for _, s in r.nbest], dtype=np.float64)
            out[f"min_margin_w{W}_{b}"] = np.float64(r.min_margin)
            print(f"[{name} beam W={W}] utt {b}: T={enc.shape[0]} tokens={len(r.tokens)} (greedy {len(greedy)}, equal={list(greedy) == r.tokens}) "
                  f"score={r.score:.4f} min cut margin={r.min_margin:.2e}")
    return out


def forward_fixture(name="tiny", N=3, T=14, Umax=5, seed=51):
    """``Transducer.forward`` (models.py:308-359) of the imported reference in eval mode: the training-time joint lattice.
    Lengths are ragged; the lattice is compared on the valid region only."""
    cfg = weights.CONFIGS[name]
    ref = build_reference(cfg)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, T, cfg.feature_sz, 1, generator=g)
    y = torch.randint(3, cfg.vocab_sz, (N, Umax), generator=g)
    xl = torch.tensor([T, T - 3, T - 5][:N])
    yl = torch.tensor([Umax, Umax - 2, Umax - 1][:N])
    for n in range(N):
        y[n, int(yl[n]):] = 0
    with torch.no_grad():
        lat = ref((x, y, xl, yl))
    print(f"[{name} forward] lattice {tuple(lat.shape)}")
    return {"config": name, "weight_seed": WEIGHT_SEED, "x": x[..., 0].numpy(), "y": y.numpy().astype(np.int32), "xl": xl.numpy().astype(np.int32),
            "yl": yl.numpy().astype(np.int32), "lattice": lat.numpy()}


LM_WEIGHT_SEED = 4321


def build_reference_lm(lm_cfg):
    """The reference ``LM`` (lm.py:20-41) with the seeded synthetic weights; attached as ``Transducer.lm``
    (models.py:234, api-server.py:158-161)."""
    import importlib

    ref_shim.import_reference_models()
    L = importlib.import_module("libreasr.lib.lm")
    lm = L.LM(lm_cfg.vocab_sz, lm_cfg.embed_sz, lm_cfg.hidden_sz, lm_cfg.num_layers)
    sd = weights.make_lm_state_dict(lm_cfg, LM_WEIGHT_SEED)
    lm.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=True)
    lm.eval()
    return lm


def lm_fixture(name, lm_name, n_utt, n_samples, audio_seed, n_chunks, stream_seed, lm_rows=4):
    """Shallow fusion: ``decode_greedy`` and ``transcribe_stream`` of the reference with
    ``m.lm`` set.  Also stores the first standardised LM rows the fuser holds (lm.py:50-54) and the fused-score margins."""
    from . import rnnt_oracle as O

    cfg, lm_cfg = weights.CONFIGS[name], weights.LM_CONFIGS[lm_name]
    ref = build_reference(cfg)
    with torch.no_grad():
        plain = []
        audio = weights.make_audio(n_utt, n_samples, audio_seed)
        feats_all = [ref_features_offline(torch.from_numpy(audio[b:b + 1]), cfg)[0] for b in range(n_utt)]
        for f in feats_all:
            plain.append(ref.decode_greedy(f, max_iters=3)[0])
    ref.lm = build_reference_lm(lm_cfg)
    orc = O.OracleTransducer(cfg, weights.make_state_dict(cfg, WEIGHT_SEED))
    olm = O.OracleLM(lm_cfg, weights.make_lm_state_dict(lm_cfg, LM_WEIGHT_SEED))
    out = {"config": name, "lm_config": lm_name, "weight_seed": WEIGHT_SEED, "lm_weight_seed": LM_WEIGHT_SEED,
           "audio_seed": audio_seed, "n_utt": n_utt, "n_samples": n_samples, "max_iters": 3,
           "n_chunks": n_chunks, "stream_seed": stream_seed}
    for b, feats in enumerate(feats_all):
        with torch.no_grad():
            toks, nlp, metrics, extra = ref.decode_greedy(feats, max_iters=3)
        toks = [int(t) for t in toks]
        out[f"tokens_{b}"] = np.asarray(toks, dtype=np.int32)
        out[f"tokens_nolm_{b}"] = np.asarray(plain[b], dtype=np.int32)
        out[f"neg_log_p_{b}"] = np.float64(nlp)
        out[f"iters_{b}"] = np.asarray(extra["iters"], dtype=np.int32)
        # the rows the reference's fuser held after each of the first emitted tokens (reference LM + standardize)
        fz = importlib_lm().LMFuser(ref.lm)
        rows = []
        with torch.no_grad():
            for t in toks[:lm_rows]:
                fz.advance(torch.LongTensor([[t]]))
                rows.append(fz.lm_logits.reshape(-1).clone())
        out[f"lm_rows_{b}"] = torch.stack(rows).numpy() if rows else np.zeros((0, lm_cfg.vocab_sz), np.float32)
        # fused-score margins along the reference's own path (informational: tie risk of the fused arg max)
        res = orc.decode_greedy(feats[..., 0], max_iters=3, impl="aten", lm=olm, keep_logits=True)
        assert res["tokens"] == toks, "oracle restatement of the fusion differs from the reference"
        n_flip = sum(1 for x, y in zip(toks, plain[b]) if x != y) + abs(len(toks) - len(plain[b]))
        out[f"fused_margins_{b}"] = np.asarray(olm.fused_margins, dtype=np.float32)
        out[f"margins_{b}"] = np.asarray(res["margins"], dtype=np.float32)
        print(f"[{name}+lm:{lm_name}] utt {b}: tokens={len(toks)} (no-LM {len(plain[b])}), positions that differ from the no-LM decode: {n_flip}; "
              f"min margin joint {min(res['margins']):.2e}, fused {min(olm.fused_margins) if olm.fused_margins else float('nan'):.2e}")
    a = weights.make_audio(1, n_chunks * CHUNK, stream_seed)[0]
    a[:CHUNK] = 0.0
    rows = [c for c in ref_stream_chunks(torch.from_numpy(a), cfg)]
    with torch.no_grad():
        yields = [([int(t) for t in y], [int(t) for t in ys]) for (y, ys, _r) in ref.transcribe_stream(iter(rows), lambda t: list(t), max_iters=10)]
    out["stream_tokens_all"] = np.asarray(yields[-1][0] if yields else [], dtype=np.int32)
    out["stream_chunk_counts"] = np.asarray([len(ys) for _, ys in yields], dtype=np.int32)
    print(f"[{name}+lm:{lm_name} stream] yields={len(yields)} tokens={len(out['stream_tokens_all'])}")
    return out


def importlib_lm():
    import importlib

    return importlib.import_module("libreasr.lib.lm")


def main():
    if not ref_shim.reference_available():
        sys.exit("reference tree missing; fixtures can only be generated in the authoring container")
    warnings.filterwarnings("ignore")
    torch.manual_seed(0)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    jobs = {
        "tiny_offline": lambda: offline_fixture("tiny", n_utt=3, n_samples=40000, audio_seed=21, full=True),
        "tiny_stream": lambda: stream_fixture("tiny", n_chunks=40, audio_seed=22),
        "tiny_stream_reset": lambda: stream_fixture("tiny", n_chunks=44, audio_seed=23, reset_after=(4, 11)),
        "tiny_modules": lambda: modules_fixture("tiny"),
        "tiny_forward": lambda: forward_fixture("tiny"),
        "cfg2_offline": lambda: offline_fixture("cfg2", n_utt=2, n_samples=80000, audio_seed=105),
        "cfg2_stream": lambda: stream_fixture("cfg2", n_chunks=50, audio_seed=47),
        "ref_offline": lambda: offline_fixture("ref", n_utt=1, n_samples=48000, audio_seed=25),
        "cfg4_offline": lambda: offline_fixture("cfg4", n_utt=1, n_samples=32000, audio_seed=26),
        "cfg4_b8": lambda: offline_fixture("cfg4", n_utt=8, n_samples=160000, audio_seed=27),   # BASELINE configs[3] shape, >= 10 s
        "tiny_beam": lambda: beam_fixture("tiny", widths=(1, 2, 4, 8), n_utt=3, n_samples=32000, audio_seed=41),
        "cfg2_beam": lambda: beam_fixture("cfg2", widths=(4,), n_utt=2, n_samples=48000, audio_seed=43),
        "cfg4_beam": lambda: beam_fixture("cfg4", widths=(4, 8), n_utt=2, n_samples=40000, audio_seed=45),
        "cfg1_demo": lambda: demo_fixture("ref"),
        "tiny_lm": lambda: lm_fixture("tiny", "tiny", n_utt=3, n_samples=40000, audio_seed=21, n_chunks=40, stream_seed=22),
        "tiny_lm_untied": lambda: lm_fixture("tiny", "tiny_untied", n_utt=2, n_samples=40000, audio_seed=31, n_chunks=30, stream_seed=32),
        "cfg2_lm": lambda: lm_fixture("cfg2", "en", n_utt=2, n_samples=80000, audio_seed=105, n_chunks=50, stream_seed=47),
    }
    only = sys.argv[1:]
    if not only or "testing_conf" in only:
        # the reference's shipped configuration (config/testing.yaml), verbatim as JSON: the GPU box has no reference tree
        import json

        import yaml

        with open(os.path.join(ref_shim.REFERENCE_ROOT, "config", "testing.yaml")) as f:
            conf = yaml.safe_load(f)
        with open(os.path.join(GOLDEN_DIR, "testing_conf.json"), "w") as f:
            json.dump(conf, f, indent=1, sort_keys=True)
        print("wrote testing_conf.json")
    for k, fn in jobs.items():
        if only and k not in only:
            continue
        np.savez_compressed(os.path.join(GOLDEN_DIR, k + ".npz"), **fn())
        print("wrote", k, os.path.getsize(os.path.join(GOLDEN_DIR, k + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
