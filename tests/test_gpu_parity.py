"""GPU parity tests (run with -m gpu on the B200 box).  Every call goes through the C ABI
(include/rnnt_b200.h) via the reference-named Python surface; the checker is the CPU oracle
(pinned to the reference by tests/test_oracle_golden.py) and the golden fixtures produced
by the imported reference itself.

Tolerances (fp32 path, gemm_mode 0): features 5e-5 abs in the log domain, encoder 5e-5,
joint logits / log-softmax rows 2e-4 abs; token sequences, per-frame iteration counts:
exact."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import rnnt_oracle as O
from oracle import weights

pytestmark = pytest.mark.gpu
CHUNK = 1280
GEMM_MODE = int(os.environ.get("RNNT_GEMM_MODE", "1"))  # 0 = fp32 CUDA cores, 1 = tcgen05 3xFP16


class Lang:
    def denumericalize(self, ids):
        return list(ids)


_models = {}


def model_for(name):
    """Reference-surface Transducer on cuda:0 with the synthetic state_dict loaded."""
    from libreasr_b200.lib.models import Transducer

    if name not in _models:
        cfg = weights.CONFIGS[name]
        sd = weights.make_state_dict(cfg, 1234)
        m = Transducer(cfg.feature_sz, cfg.embed_sz, cfg.vocab_sz, cfg.hidden_sz, cfg.out_sz, cfg.joint_sz, Lang(),
                       encoder_kwargs={"num_layers": cfg.enc_layers}, predictor_kwargs={"num_layers": cfg.pred_layers},
                       gemm_mode=GEMM_MODE)
        m.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()}, strict=True)
        m = m.to("cuda:0")
        _models[name] = (cfg, sd, m, O.OracleTransducer(cfg, sd))
    return _models[name]


def cpu(t):
    return t.detach().float().cpu().numpy()


# ---------------- golden fixtures from the imported reference ----------------
@pytest.mark.parametrize("name", ["tiny_offline", "cfg2_offline", "ref_offline", "cfg4_offline"])
def test_offline_greedy_matches_reference_fixture(name):
    g = load_golden(name)
    cfg, sd, m, _ = model_for(str(g["config"]))
    eng = m.engine()
    audio = weights.make_audio(int(g["n_utt"]), int(g["n_samples"]), int(g["audio_seed"]))
    for b in range(int(g["n_utt"])):
        feats = eng.features(torch.from_numpy(audio[b:b + 1]).cuda())[0]  # [T, X]
        toks, nlp, metrics, extra = m.decode_greedy(feats.unsqueeze(-1), max_iters=int(g["max_iters"]), return_logp=True)
        assert toks == g[f"tokens_{b}"].tolist()
        assert extra["iters"] == g[f"iters_{b}"].tolist()
        assert abs(nlp - float(g[f"neg_log_p_{b}"])) < 2e-3
        assert abs(metrics["alignment_score"] - float(g[f"alignment_score_{b}"])) < 1e-9
        logp = torch.stack([o.reshape(-1) for o in extra["outs"]])
        np.testing.assert_allclose(cpu(logp.max(-1).values), g[f"maxlogp_{b}"], atol=2e-4)
        assert cpu(logp.argmax(-1)).tolist() == g[f"argmax_{b}"].tolist()
        enc = m.encoder(feats[None, :, :, None])[0]
        if f"feats_{b}" in g:
            np.testing.assert_allclose(cpu(feats), g[f"feats_{b}"], atol=5e-5)
            np.testing.assert_allclose(cpu(enc), g[f"enc_{b}"], atol=5e-5)
            np.testing.assert_allclose(cpu(logp), g[f"logp_{b}"], atol=2e-4)
        else:
            s, n = int(g["enc_stride"]), int(g["n_logp"])
            np.testing.assert_allclose(cpu(feats[::s, ::7]), g[f"feats_sub_{b}"], atol=5e-5)
            np.testing.assert_allclose(cpu(enc[::s, ::16]), g[f"enc_sub_{b}"], atol=1e-4)
            np.testing.assert_allclose(cpu(logp[:n]), g[f"logp_first_{b}"], atol=5e-4)


@pytest.mark.parametrize("name", ["tiny_stream", "cfg2_stream"])
def test_transcribe_stream_matches_reference_fixture(name):
    """Transducer.transcribe_stream fed by the reference-named stream transforms."""
    from libreasr_b200.lib.transforms import FusedStreamFeatures

    g = load_golden(name)
    cfg, sd, m, _ = model_for(str(g["config"]))
    n_chunks = int(g["n_chunks"])
    audio = weights.make_audio(1, n_chunks * CHUNK, int(g["audio_seed"]))[0]
    audio[: int(g["lead_zero_chunks"]) * CHUNK] = 0.0
    tfm = FusedStreamFeatures(m.engine(), n_buffer=2)
    frames, rows = [], []
    for j in range(n_chunks):  # api-server.py:83-115
        frames.append(torch.from_numpy(audio[None, j * CHUNK:(j + 1) * CHUNK]))
        if len(frames) < 3:
            rows.append(None)
            continue
        rows.append(tfm(torch.cat(frames, dim=1)))
        frames.pop(0)
    feats = np.stack([cpu(r[..., 0]) for r in rows if r is not None])
    np.testing.assert_allclose(feats if name.startswith("tiny") else feats[:, :, ::7], g["feats"], atol=5e-5)
    yields = [(list(y), list(ys)) for y, ys, _ in m.transcribe_stream(iter(rows), lambda t: list(t), max_iters=int(g["max_iters"]))]
    assert len(yields) == int(g["n_yields"])
    assert [len(ys) for _, ys in yields] == g["chunk_counts"].tolist()
    assert yields[-1][0] == g["tokens_all"].tolist()


def test_config1_demo_utterance_through_libreasr_facade():
    """BASELINE.json configs[0]: the reference's demo utterance through ``LibreASR.transcribe()`` with the
    reference-shape model -- token ids identical to the imported reference's ``decode_greedy``."""
    from libreasr_b200 import LibreASR

    g = load_golden("cfg1_demo")
    cfg, sd, m, _ = model_for(str(g["config"]))
    audio = g["pcm16"].astype(np.float32) / 32768.0
    asr = LibreASR(m)
    assert asr.transcribe(audio) == g["tokens"].tolist()
    # and the reference entry point (api-server.py:75-78): features [T, X, 1] -> model.transcribe
    feats = m.engine().features(torch.from_numpy(audio)[None].cuda())[0].unsqueeze(-1)
    toks, metrics = m.transcribe(feats)
    assert toks == g["tokens"].tolist()
    assert abs(metrics["alignment_score"] - float(g["alignment_score"])) < 1e-9


def test_modules_match_reference_fixture():
    """Encoder / Predictor / Joint forward with explicit state."""
    g = load_golden("tiny_modules")
    cfg, sd, m, _ = model_for("tiny")
    x = torch.from_numpy(g["x"])[..., None]
    T = x.shape[1] // 2
    e1, s1 = m.encoder(x[:, :T], return_state=True)
    e2, s2 = m.encoder(x[:, T:], state=s1, return_state=True)
    ef = m.encoder(x)
    np.testing.assert_allclose(cpu(e1), g["enc_first"], atol=5e-5)
    np.testing.assert_allclose(cpu(e2), g["enc_second"], atol=5e-5)
    np.testing.assert_allclose(cpu(ef), g["enc_full"], atol=5e-5)
    assert s2[0][0].shape == (1, x.shape[0], cfg.hidden_sz)
    np.testing.assert_allclose(np.stack([cpu(s[0][0]) for s in s2]), g["enc_h"], atol=5e-5)
    np.testing.assert_allclose(np.stack([cpu(s[1][0]) for s in s2]), g["enc_c"], atol=5e-5)
    toks = torch.from_numpy(g["tokens"]).long()
    st, outs = None, []
    for j in range(toks.shape[1]):
        o, st = m.predictor(toks[:, j:j + 1], state=st)
        assert o.shape == (toks.shape[0], 1, cfg.hidden_sz)
        outs.append(o[:, 0])
    np.testing.assert_allclose(cpu(torch.stack(outs, 1)), g["pred_outs"], atol=5e-5)
    np.testing.assert_allclose(np.stack([cpu(s[0]) for s in st]), g["pred_h"], atol=5e-5)
    jl = m.joint(outs[-1], e2[:, -1])
    np.testing.assert_allclose(cpu(jl), g["joint_logits"], atol=2e-4)
    # broadcasting form used inside the reference decode loop (models.py:413-415)
    jb = m.joint(outs[-1][:1][None], e2[0, -1][None, None, None])
    assert jb.shape == (1, 1, 1, cfg.vocab_sz)
    np.testing.assert_allclose(cpu(jb[0, 0, 0]), g["joint_logits"][0], atol=2e-4)


# ---------------- oracle comparisons on fresh seeded inputs ----------------
@pytest.mark.parametrize("name,n_utt,seconds", [("tiny", 5, 4.0), ("cfg2", 4, 6.0)])
def test_batched_transcribe_matches_oracle(name, n_utt, seconds):
    """Batched path (one call for all utterances) == the reference's utterance-by-utterance path."""
    cfg, sd, m, orc = model_for(name)
    eng = m.engine()
    n = int(seconds * 16000)
    audio = weights.make_audio(n_utt, n, seed=77)
    r = eng.transcribe(torch.from_numpy(audio).cuda(), max_iters=3)
    from libreasr_b200.engine import tokens_to_lists

    got = tokens_to_lists(r["tokens"], r["ntok"])
    for b in range(n_utt):
        feats = O.features_offline(torch.from_numpy(audio[b:b + 1]), cfg)[0]
        ro = orc.decode_greedy(feats, max_iters=3, impl="aten")
        assert got[b] == ro["tokens"], f"utt {b}: min oracle margin {min(ro['margins']):.2e}"
        assert r["iters"][b].cpu().tolist() == ro["iters"]
        assert abs(float(r["neg_logp"][b]) - ro["neg_log_p"]) < 2e-3


def test_ragged_batch_matches_per_utterance():
    """lens[b] < n: each utterance reflect-pads at its own end and stops at its own T_b."""
    cfg, sd, m, orc = model_for("tiny")
    eng = m.engine()
    from libreasr_b200.engine import tokens_to_lists

    n = 48000
    audio = weights.make_audio(4, n, seed=5)
    lens = torch.tensor([48000, 40001, 16000, 1700], dtype=torch.int32)
    r = eng.transcribe(torch.from_numpy(audio).cuda(), lens.cuda(), max_iters=3)
    got = tokens_to_lists(r["tokens"], r["ntok"])
    for b in range(4):
        want = O.transcribe_batch(orc, audio[b:b + 1, : int(lens[b])], max_iters=3, impl="explicit")[0]
        assert got[b] == want, b


def test_host_api_equals_device_api():
    cfg, sd, m, _ = model_for("tiny")
    eng = m.engine()
    from libreasr_b200.engine import tokens_to_lists

    audio = torch.from_numpy(weights.make_audio(3, 32000, seed=9))
    a = eng.transcribe(audio.cuda(), max_iters=3)
    b = eng.transcribe_host(audio.pin_memory(), max_iters=3)
    assert tokens_to_lists(a["tokens"], a["ntok"]) == tokens_to_lists(b["tokens"], b["ntok"])
    np.testing.assert_array_equal(a["neg_logp"].cpu().numpy(), b["neg_logp"].numpy())


def test_libreasr_facade_transcribe_and_stream():
    from libreasr_b200 import LibreASR

    cfg, sd, m, orc = model_for("tiny")
    asr = LibreASR(m)
    audio = weights.make_audio(1, 40 * CHUNK, seed=13)[0]
    want = O.transcribe_batch(orc, audio[None], max_iters=3, impl="explicit")[0]
    assert asr.transcribe(audio) == want
    assert asr.transcribe(audio.tobytes()) == want  # raw little-endian f32 PCM, utils.py:149-153
    fe = O.StreamFrontend(cfg)
    rows = [fe.push(torch.from_numpy(audio[None, j * CHUNK:(j + 1) * CHUNK])) for j in range(40)]
    ys = list(orc.transcribe_stream(iter(rows), max_iters=10))
    out = list(asr.stream(audio[j * CHUNK:(j + 1) * CHUNK] for j in range(40)))
    assert len(out) == len(ys)
    assert out[-1][0] == ys[-1][0]


# ---------------- size-independent properties at BASELINE sizes ----------------
def test_streaming_state_carry_equals_full_sequence_cfg2():
    """Chunked encoder with carried (h, c) == one pass over the full sequence (config 2 size)."""
    cfg, sd, m, _ = model_for("cfg2")
    eng = m.engine()
    g = torch.Generator().manual_seed(0)
    feats = torch.randn(8, 124, cfg.feature_sz, generator=g).cuda()
    full, _ = eng.encode(feats)
    st, outs = None, []
    for t0 in range(0, 124, 31):
        o, st = eng.encode(feats[:, t0:t0 + 31].contiguous(), state=st, want_state=True)
        outs.append(o)
    assert float((torch.cat(outs, 1) - full).abs().max()) == 0.0  # same kernels, same order -> bit equal


def test_batch_invariance_full_size_cfg2():
    """32 x 10 s (BASELINE config 2): every utterance decodes to the same tokens alone and in the batch;
    a permuted batch gives permuted results."""
    from libreasr_b200.engine import tokens_to_lists

    cfg, sd, m, _ = model_for("cfg2")
    eng = m.engine()
    audio = torch.from_numpy(weights.make_audio(32, 160000, seed=0)).cuda()
    r = eng.transcribe(audio, max_iters=3)
    toks = tokens_to_lists(r["tokens"], r["ntok"])
    assert sum(map(len, toks)) > 0
    perm = torch.randperm(32, generator=torch.Generator().manual_seed(1))
    rp = eng.transcribe(audio[perm.cuda()].contiguous(), max_iters=3)
    tp = tokens_to_lists(rp["tokens"], rp["ntok"])
    assert [tp[i] for i in range(32)] == [toks[int(perm[i])] for i in range(32)]
    for b in (0, 13, 31):
        r1 = eng.transcribe(audio[b:b + 1].contiguous(), max_iters=3)
        assert tokens_to_lists(r1["tokens"], r1["ntok"])[0] == toks[b]
    it = r["iters"].cpu().numpy().astype(np.int64)
    assert it.min() >= 1 and it.max() <= 3
    # every evaluation but the last of a frame emitted a token; the last one emitted iff it was not blank,
    # which can only happen on the max_iters-th evaluation (models.py:408-437)
    ntok = r["ntok"].cpu().numpy().astype(np.int64)
    lo = (it - 1).sum(1)
    hi = lo + (it == 3).sum(1)
    assert (lo <= ntok).all() and (ntok <= hi).all()


def test_error_behaviour():
    """Shape validation mirrors the reference's ValueErrors (haste/base_rnn.py:81-117)."""
    cfg, sd, m, _ = model_for("tiny")
    eng = m.engine()
    with pytest.raises(ValueError):
        eng.encode(torch.zeros(1, 4, cfg.feature_sz + 4, device="cuda"))
    with pytest.raises(ValueError):
        eng.encode(torch.zeros(2, 4, cfg.feature_sz, device="cuda"), state=(torch.zeros(1, 2, 64, device="cuda"),) * 2)
    with pytest.raises(ValueError):
        eng.features(torch.zeros(1, 300, device="cuda"))  # shorter than the reflect padding
    with pytest.raises(NotImplementedError):
        m(None)


def test_bench_workload_tokens_identical_to_oracle():
    """The headline workload itself (bench.py: 32 x 10 s, cfg2, synth.BENCH_AUDIO_SEED): every utterance's greedy
    token sequence and per-frame evaluation counts equal the CPU oracle's."""
    from libreasr_b200.engine import tokens_to_lists

    cfg, sd, m, orc = model_for("cfg2")
    eng = m.engine()
    audio = weights.make_audio(32, 160000, seed=weights.BENCH_AUDIO_SEED)
    r = eng.transcribe(torch.from_numpy(audio).cuda(), max_iters=3)
    got = tokens_to_lists(r["tokens"], r["ntok"])
    torch.set_num_threads(min(16, torch.get_num_threads()))
    n_tok = 0
    for b in range(32):
        ro = orc.decode_greedy(O.features_offline(torch.from_numpy(audio[b:b + 1]), cfg)[0], max_iters=3, impl="aten")
        assert got[b] == ro["tokens"], f"utt {b}: oracle min margin {min(ro['margins']):.2e}"
        assert r["iters"][b].cpu().tolist() == ro["iters"]
        assert abs(float(r["neg_logp"][b]) - ro["neg_log_p"]) < 5e-3
        n_tok += len(got[b])
    assert n_tok > 500


# ---------------- batch-size dependent kernel paths ----------------
@pytest.mark.parametrize("name,B,seconds", [("tiny", 48, 1.5), ("tiny", 100, 1.2), ("tiny", 150, 1.0), ("cfg2", 40, 1.5), ("cfg2", 72, 1.2)])
def test_large_batches_match_single_utterance_runs(name, B, seconds):
    """B in (32, 64]: M = 128 stacked tiles; B in (64, 128]: unfused LSTM tiles + decode sub-batches of 64;
    B > 128: encoder sub-batches.  Every utterance must decode exactly as it does alone."""
    from libreasr_b200.engine import tokens_to_lists

    cfg, sd, m, orc = model_for(name)
    eng = m.engine()
    n = int(seconds * 16000)
    audio = torch.from_numpy(weights.make_audio(B, n, seed=200 + B)).cuda()
    r = eng.transcribe(audio, max_iters=3)
    toks = tokens_to_lists(r["tokens"], r["ntok"])
    assert sum(map(len, toks)) > 0
    for b in list(range(0, B, max(1, B // 7))) + [B - 1]:
        r1 = eng.transcribe(audio[b:b + 1].contiguous(), max_iters=3)
        assert tokens_to_lists(r1["tokens"], r1["ntok"])[0] == toks[b], b
        assert r1["iters"][0].cpu().tolist() == r["iters"][b].cpu().tolist()
    # and two of them against the CPU oracle
    for b in (1, B - 2):
        want = O.transcribe_batch(orc, audio[b:b + 1].cpu().numpy(), max_iters=3, impl="aten")[0]
        assert toks[b] == want, b


def test_gemm_arithmetic_modes_against_fp64():
    """The library GEMM behind the gate / joint projections: both arithmetic modes stay at fp32-grade error."""
    cfg, sd, m, _ = model_for("tiny")
    eng = m.engine()
    g = torch.Generator(device="cuda").manual_seed(0)
    for (M_, N_, K_) in [(130, 260, 160), (777, 1024, 800)]:
        A = torch.randn(M_, K_, device="cuda", generator=g)
        W = (torch.rand(N_, K_, device="cuda", generator=g) * 2 - 1) * 0.1
        b = torch.randn(N_, device="cuda", generator=g)
        ref = A.double() @ W.double().t() + b.double()
        scale = float(ref.abs().mean())
        for mode in (0, 1):
            C = eng.selftest_gemm(A, W, b, gemm_mode=mode)
            rel = float((C.double() - ref).pow(2).mean().sqrt()) / scale
            assert rel < 4e-6, (mode, M_, N_, K_, rel)


# ---------------- LM shallow fusion (lm.py:43-83) ----------------
_lm_models = {}


def lm_model_for(name, lm_name):
    """Transducer with the reference-named ``LM`` attached (``m.lm = lm``, api-server.py:158-161)."""
    from libreasr_b200.lib.lm import LM
    from libreasr_b200.lib.models import Transducer

    key = (name, lm_name)
    if key not in _lm_models:
        cfg, lc = weights.CONFIGS[name], weights.LM_CONFIGS[lm_name]
        sd, lsd = weights.make_state_dict(cfg, 1234), weights.make_lm_state_dict(lc, 4321)
        m = Transducer(cfg.feature_sz, cfg.embed_sz, cfg.vocab_sz, cfg.hidden_sz, cfg.out_sz, cfg.joint_sz, Lang(),
                       encoder_kwargs={"num_layers": cfg.enc_layers}, predictor_kwargs={"num_layers": cfg.pred_layers},
                       gemm_mode=GEMM_MODE)
        m.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()}, strict=True)
        lm = LM(lc.vocab_sz, lc.embed_sz, lc.hidden_sz, lc.num_layers)
        lm.load_state_dict({k: torch.as_tensor(v) for k, v in lsd.items()}, strict=True)
        m.lm = lm
        m = m.to("cuda:0")
        _lm_models[key] = (cfg, lc, m, O.OracleTransducer(cfg, sd), O.OracleLM(lc, lsd))
    return _lm_models[key]


@pytest.mark.parametrize("name", ["tiny_lm", "tiny_lm_untied", "cfg2_lm"])
def test_lm_fusion_matches_reference_fixture(name):
    """decode_greedy and transcribe_stream with ``m.lm`` set: token ids identical to the imported reference's."""
    from libreasr_b200.lib.transforms import FusedStreamFeatures

    g = load_golden(name)
    cfg, lc, m, _, _ = lm_model_for(str(g["config"]), str(g["lm_config"]))
    eng = m.engine()
    audio = weights.make_audio(int(g["n_utt"]), int(g["n_samples"]), int(g["audio_seed"]))
    for b in range(int(g["n_utt"])):
        feats = eng.features(torch.from_numpy(audio[b:b + 1]).cuda())[0]
        toks, nlp, metrics, extra = m.decode_greedy(feats.unsqueeze(-1), max_iters=int(g["max_iters"]))
        assert toks == g[f"tokens_{b}"].tolist()
        assert extra["iters"] == g[f"iters_{b}"].tolist()
        assert abs(nlp - float(g[f"neg_log_p_{b}"])) < 2e-3   # accumulates the PRE-fusion log-probs (models.py:422)
    n_chunks = int(g["n_chunks"])
    a = weights.make_audio(1, n_chunks * CHUNK, int(g["stream_seed"]))[0]
    a[:CHUNK] = 0.0
    tfm = FusedStreamFeatures(eng, n_buffer=2)
    frames, rows = [], []
    for j in range(n_chunks):
        frames.append(torch.from_numpy(a[None, j * CHUNK:(j + 1) * CHUNK]))
        if len(frames) < 3:
            rows.append(None)
            continue
        rows.append(tfm(torch.cat(frames, dim=1)))
        frames.pop(0)
    yields = [(list(y), list(ys)) for y, ys, _ in m.transcribe_stream(iter(rows), lambda t: list(t), max_iters=10)]
    assert [len(ys) for _, ys in yields] == g["stream_chunk_counts"].tolist()
    assert (yields[-1][0] if yields else []) == g["stream_tokens_all"].tolist()


def test_lm_fusion_batched_ragged_matches_oracle():
    """A ragged batch through Engine.transcribe with the LM on: every utterance has its own fuser."""
    cfg, lc, m, orc, olm = lm_model_for("tiny", "tiny")
    eng = m.engine()
    n = 48000
    audio = weights.make_audio(7, n, seed=91)
    lens = np.array([n, 30000, 41000, n, 17000, 25000, 36000], dtype=np.int32)
    r = eng.transcribe(torch.from_numpy(audio).cuda(), lens=torch.from_numpy(lens))
    got = [r["tokens"][b, : int(r["ntok"][b])].tolist() for b in range(7)]
    plain = []
    for b in range(7):
        feats = O.features_offline(torch.from_numpy(audio[b:b + 1, : lens[b]]), cfg)[0]
        want = orc.decode_greedy(feats, max_iters=3, impl="aten", lm=olm)
        assert got[b] == want["tokens"], f"utterance {b}"
        plain.append(orc.decode_greedy(feats, max_iters=3, impl="aten")["tokens"])
    assert got != plain


def test_lm_fuser_state_carries_across_calls_and_resets():
    """Streaming: chunked decode with a registered fuser blob == one decode over the whole sequence (bit-equal tokens),
    a zeroed blob == a fresh fuser, and two streams in one batch keep separate fusers."""
    cfg, lc, m, _, _ = lm_model_for("tiny", "tiny")
    eng = m.engine()
    audio = torch.from_numpy(weights.make_audio(2, 64000, seed=93)).cuda()
    feats = eng.features(audio)
    enc, _ = eng.encode(feats)
    T = enc.shape[1]
    full = eng.decode_greedy(enc, max_iters=10)
    want = [full["tokens"][b, : int(full["ntok"][b])].tolist() for b in range(2)]
    blob = eng.new_lm_state(2)
    state, got = None, [[], []]
    for t0 in range(0, T, 7):
        r = eng.decode_greedy(enc[:, t0:t0 + 7].contiguous(), max_iters=10, state=state, want_state=True, lm_state=blob)
        state = r["state"]
        for b in range(2):
            got[b] += r["tokens"][b, : int(r["ntok"][b])].tolist()
    assert got == want
    blob.zero_()
    again = eng.decode_greedy(enc, max_iters=10, lm_state=blob)
    assert [again["tokens"][b, : int(again["ntok"][b])].tolist() for b in range(2)] == want
    # per-stream fusers: decoding stream 1 alone gives the same tokens as inside the batch
    solo = eng.decode_greedy(enc[1:2].contiguous(), max_iters=10)
    assert solo["tokens"][0, : int(solo["ntok"][0])].tolist() == want[1]


def test_lm_errors():
    cfg, lc, m, _, _ = lm_model_for("tiny", "tiny")
    eng = m.engine()
    enc = torch.zeros(3, 4, cfg.hidden_sz, device="cuda")
    with pytest.raises(ValueError):
        eng.decode_greedy(enc, lm_state=eng.new_lm_state(2))     # blob sized for another stream count
    _, _, plain, _ = model_for("tiny")
    with pytest.raises(Exception):
        plain.engine().new_lm_state(1)                            # handle without a language model


# ---------------- streaming sessions (rnnt_b200_stream_*) ----------------
def _oracle_stream_tokens(orc, cfg, audio_1d, n_chunks, lm=None):
    fe = O.StreamFrontend(cfg)
    rows = [fe.push(torch.from_numpy(audio_1d[None, j * CHUNK:(j + 1) * CHUNK])) for j in range(n_chunks)]
    ys = list(orc.transcribe_stream(iter(rows), max_iters=10, lm=lm))
    return [list(y) for _, y in ys]


@pytest.mark.parametrize("with_lm", [False, True])
def test_stream_session_many_streams_match_oracle_and_reset(with_lm):
    """5 concurrent streams through the C streaming session (host chunks in, host tokens out): every stream's per-tick
    tokens equal the oracle's transcribe_stream run on that stream alone; reset() starts all streams over."""
    from libreasr_b200.api import StreamBatch

    if with_lm:
        cfg, lc, m, orc, olm = lm_model_for("tiny", "tiny")
    else:
        cfg, sd, m, orc = model_for("tiny")
        olm = None
    eng = m.engine()
    S, n_chunks = 5, 24
    audio = weights.make_audio(S, n_chunks * CHUNK, seed=71)
    want = [_oracle_stream_tokens(orc, cfg, audio[b], n_chunks, olm) for b in range(S)]
    sb = StreamBatch(eng, S, max_iters=10)
    for rep in range(2):
        ticks = [[] for _ in range(S)]
        for j in range(n_chunks):
            slab = torch.from_numpy(np.ascontiguousarray(audio[:, j * CHUNK:(j + 1) * CHUNK]))
            new = sb.push(slab if j % 2 else slab.cuda())   # host and device chunk sources
            if new is not None:
                for b in range(S):
                    ticks[b].append(new[b])
        assert ticks == want, f"pass {rep}"
        sb.reset()
    sb.close()


def test_decode_is_deterministic_when_ctas_sit_phases_out():
    """cfg4 shape (H = 1536 -> 96 CTAs, J = 1024 -> only 64 of them own a slice of the joint's first projection): CTAs
    that sit a grid phase out must not run ahead of the phase counter.  Repeated decodes are identical, equal the four
    8-utterance sub-batches, and equal the fp32 cooperative kernel (cooperative-groups grid sync) on the first 8."""
    from libreasr_b200.engine import Engine, EngineConfig, tokens_to_lists

    cfg = weights.CONFIGS["cfg4"]
    sd = weights.make_state_dict(cfg, 1234)

    def mk(mode):
        ec = EngineConfig(n_mels=cfg.n_mels, n_stack=cfg.n_stack, downsample=cfg.downsample, enc_layers=cfg.enc_layers,
                          pred_layers=cfg.pred_layers, hidden_sz=cfg.hidden_sz, embed_sz=cfg.embed_sz,
                          joint_sz=cfg.joint_sz, vocab_sz=cfg.vocab_sz, gemm_mode=mode)
        return Engine(ec).load_state_dict(sd)

    eng = mk(GEMM_MODE)
    audio = torch.from_numpy(weights.make_audio(32, 6 * 16000, seed=4)).cuda()
    enc, _ = eng.encode(eng.features(audio))
    runs = []
    for _ in range(5):
        d = eng.decode_greedy(enc, max_iters=3)
        runs.append(tokens_to_lists(d["tokens"], d["ntok"]))
    assert all(r == runs[0] for r in runs[1:])
    sub = []
    for b0 in range(0, 32, 8):
        d = eng.decode_greedy(enc[b0:b0 + 8].contiguous(), max_iters=3)
        sub += tokens_to_lists(d["tokens"], d["ntok"])
    assert sub == runs[0]
    eng0 = mk(0)
    enc0, _ = eng0.encode(eng0.features(audio[:8].contiguous()))
    d0 = eng0.decode_greedy(enc0, max_iters=3)
    assert tokens_to_lists(d0["tokens"], d0["ntok"]) == runs[0][:8]
    eng.close(); eng0.close()


@pytest.mark.parametrize("B", [40, 72])
def test_lm_fusion_wide_batches_match_oracle(B):
    """LM fusion beyond 32 utterances: 33..64 take the 128-row MMA tiles of the tcgen05 kernel (2 threads per stream),
    more than 64 run as sub-batches; every utterance still equals the oracle's single-utterance decode with its own fuser."""
    cfg, lc, m, orc, olm = lm_model_for("tiny", "tiny")
    eng = m.engine()
    n = 24000
    audio = weights.make_audio(B, n, seed=97)
    r = eng.transcribe(torch.from_numpy(audio).cuda())
    got = [r["tokens"][b, : int(r["ntok"][b])].tolist() for b in range(B)]
    for b in list(range(0, B, 5)) + [B - 1]:
        feats = O.features_offline(torch.from_numpy(audio[b:b + 1]), cfg)[0]
        assert got[b] == orc.decode_greedy(feats, max_iters=3, impl="aten", lm=olm)["tokens"], f"utterance {b}"


def test_stream_session_beyond_64_streams():
    """70 lock-step streams: more than one launch of the tcgen05 decode kernel takes (64), so the stateful decode runs as two
    sub-batches over row slices of the predictor state -- the fp32 cooperative kernel is never launched in gemm_mode 1;
    spot-checked streams equal the oracle."""
    from libreasr_b200.api import StreamBatch

    cfg, sd, m, orc = model_for("tiny")
    S, n_chunks = 70, 14
    audio = weights.make_audio(S, n_chunks * CHUNK, seed=73)
    f32_before = m.engine().fp32_decode_launches()
    sb = StreamBatch(m.engine(), S, max_iters=10)
    ticks = [[] for _ in range(S)]
    for j in range(n_chunks):
        new = sb.push(torch.from_numpy(np.ascontiguousarray(audio[:, j * CHUNK:(j + 1) * CHUNK])))
        if new is not None:
            for b in range(S):
                ticks[b].append(new[b])
    sb.close()
    if GEMM_MODE == 1:
        assert m.engine().fp32_decode_launches() == f32_before      # the tensor-core kernels served every tick
    for b in (0, 33, 63, 64, 69):
        assert ticks[b] == _oracle_stream_tokens(orc, cfg, audio[b], n_chunks), f"stream {b}"


def test_stream_session_independent_lifecycles():
    """Streams of one session start, pause and restart independently: stream 1 connects 3 ticks late, stream 2 skips two
    ticks in the middle (no chunk), stream 0 is reset (new connection) at tick 11.  Every stream's tokens equal the
    oracle's transcribe_stream on the chunks that stream actually received since its (re)start."""
    from libreasr_b200.api import StreamBatch

    cfg, sd, m, orc = model_for("tiny")
    S, n_ticks = 3, 26
    audio = weights.make_audio(S, n_ticks * CHUNK, seed=75)
    sb = StreamBatch(m.engine(), S, max_iters=10)
    got = [[] for _ in range(S)]
    fed = [[] for _ in range(S)]          # chunk indices each stream consumed since its last (re)start
    for j in range(n_ticks):
        if j == 11:
            sb.reset(0)
            got[0], fed[0] = [], []
        active = [True, j >= 3, j not in (8, 9)]
        for b in range(S):
            if active[b]:
                fed[b].append(j)
        new = sb.push(torch.from_numpy(np.ascontiguousarray(audio[:, j * CHUNK:(j + 1) * CHUNK])), active=active)
        if new is not None:
            for b in range(S):
                got[b].append(new[b])
    sb.close()
    for b in range(S):
        seq = np.concatenate([audio[b, j * CHUNK:(j + 1) * CHUNK] for j in fed[b]])
        want = _oracle_stream_tokens(orc, cfg, seq, len(fed[b]))
        flat_got = [t for tick in got[b] for t in tick]
        flat_want = [t for tick in want for t in tick]
        assert flat_got == flat_want, f"stream {b}"


@pytest.mark.parametrize("orig", [8000, 44100, 22050, 48000, 16000])
def test_resample_matches_oracle(orig):
    """Row a2: Resample to 16 kHz on the GPU vs the oracle's restatement of torchaudio.transforms.Resample (2e-6 abs), and the
    facade resamples before transcribing."""
    from libreasr_b200 import LibreASR

    cfg, sd, m, orc = model_for("tiny")
    eng = m.engine()
    x = weights.make_audio(3, int(orig * 1.3), seed=6)
    got = eng.resample(torch.from_numpy(x).cuda(), orig)
    want = O.resample(torch.from_numpy(x), orig, 16000)
    assert tuple(got.shape) == tuple(want.shape)
    np.testing.assert_allclose(cpu(got), want.numpy(), atol=2e-6)
    toks = LibreASR(m).transcribe(torch.from_numpy(x[0]), sr=orig)
    feats = O.features_offline(want[:1], cfg)[0]
    assert toks == orc.decode_greedy(feats, max_iters=3, impl="aten")["tokens"]


def test_grpc_streaming_end_to_end():
    """The serving wire path (libreasr_b200/serve.py) on the real engine: gRPC `ASR.ASR/TranscribeStream` over loopback, three
    concurrent clients sending `Audio{data, sr}` frames of 80 ms; each client's `Transcript` messages equal the oracle's
    transcribe_stream on that client's audio passed through the reference's diffing (api-server.py:117-135).  Unary
    `Transcribe` at 8 kHz goes through the resampler."""
    import itertools as it
    import threading
    import time
    from concurrent import futures

    grpc = pytest.importorskip("grpc")
    from libreasr_b200 import LibreASR
    from libreasr_b200 import serve as S

    cfg, sd, m, orc = model_for("tiny")
    denum = lambda ids: "".join(chr(0x100 + int(t)) for t in ids)   # noqa: E731
    asr = LibreASR(m)
    server = grpc.server(futures.ThreadPoolExecutor(max_workers=12))
    servicer = S.ASRServicer(asr, n_streams=4, denumericalize=denum)
    S.add_servicer_to_server(servicer, server)
    port = server.add_insecure_port("127.0.0.1:0")
    server.start()
    try:
        ch = grpc.insecure_channel(f"127.0.0.1:{port}")
        ident = lambda b: b  # noqa: E731
        stream = ch.stream_stream(f"/{S.SERVICE}/TranscribeStream", request_serializer=ident, response_deserializer=ident)
        n_chunks = 30
        audio = weights.make_audio(3, n_chunks * CHUNK, seed=77)

        def client(b, out):
            def gen():
                for j in range(n_chunks):
                    yield S.encode_audio(audio[b, j * CHUNK:(j + 1) * CHUNK].astype("<f4").tobytes(), 16000)
                    time.sleep(0.001 * (b + 1))
            out.extend(S.decode_transcript(r) for r in stream(gen(), timeout=60))

        # a fourth client streams 8 kHz frames (80 ms = 640 samples): every frame is resampled on its own, like the reference's
        # stream pipeline applies Resample per frame (transforms.py:135-144, testing.yaml:357)
        audio8 = weights.make_audio(1, n_chunks * (CHUNK // 2), seed=79)[0]
        want16 = np.concatenate([O.resample(torch.from_numpy(audio8[None, j * (CHUNK // 2):(j + 1) * (CHUNK // 2)]), 8000, 16000)[0].numpy()
                                 for j in range(n_chunks)])
        assert want16.shape[0] == n_chunks * CHUNK

        def client8(out):
            def gen():
                for j in range(n_chunks):
                    yield S.encode_audio(audio8[j * (CHUNK // 2):(j + 1) * (CHUNK // 2)].astype("<f4").tobytes(), 8000)
                    time.sleep(0.002)
            out.extend(S.decode_transcript(r) for r in stream(gen(), timeout=60))

        outs = [[] for _ in range(4)]
        ths = [threading.Thread(target=client, args=(b, outs[b])) for b in range(3)] + [threading.Thread(target=client8, args=(outs[3],))]
        for t in ths:
            t.start()
        for t in ths:
            t.join(timeout=90)
        for b in range(4):
            steps = _oracle_stream_tokens(orc, cfg, audio[b] if b < 3 else want16, n_chunks)
            want, last, last_diff, y = [], "", "", []
            for new in steps:
                y = y + new
                if denum(new) != "":
                    now = denum(y)
                    diff = "".join(c2 for c1, c2 in it.zip_longest(last, now) if c1 != c2)
                    last = now
                    if diff == last_diff:
                        continue
                    last_diff = diff
                    want.append(diff)
            assert outs[b] == want, f"client {b}"
        assert len(outs[3]) > 0
        unary = ch.unary_unary(f"/{S.SERVICE}/Transcribe", request_serializer=ident, response_deserializer=ident)
        x8 = weights.make_audio(1, 12000, seed=78)[0]
        got = S.decode_transcript(unary(S.encode_audio(x8.astype("<f4").tobytes(), 8000), timeout=30))
        feats = O.features_offline(O.resample(torch.from_numpy(x8[None]), 8000, 16000), cfg)[0]
        assert got == denum(orc.decode_greedy(feats, max_iters=3, impl="aten")["tokens"])
    finally:
        server.stop(0)
        servicer.close()
