"""GPU parity tests at the BASELINE.json sizes that round 1 left thin ("Next round" item 1):

* configs[2]: the C streaming session (rnnt_b200_stream_push) at the cfg2 shape with 64 concurrent streams for
  >= 40 ticks, every stream against the oracle's single-stream transcribe_stream -- plain and with the shipped
  4 x 768 LM fused;
* configs[3] shape (6 x 1536): a 8 x 10 s fixture produced by the imported reference, and a 128-utterance batch
  equal to per-utterance runs;
* Transducer.from_config on the reference's real config/testing.yaml (stored verbatim as tests/golden/testing_conf.json).

Token sequences and per-frame iteration counts: exact.  All calls go through the C ABI.
"""
import copy
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN as GOLDEN_DIR, load_golden
from oracle import rnnt_oracle as O
from oracle import weights
from test_gpu_parity import CHUNK, Lang, _oracle_stream_tokens, lm_model_for, model_for

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("with_lm", [False, True])
def test_stream_session_cfg2_64_streams(with_lm):
    """BASELINE configs[2]: 64 concurrent 80 ms-chunk streams, 4 x 1024 encoder, reference windowing (3-chunk window,
    Buffer of 2), max_iters 10, 42 ticks, host chunks in / host tokens out through the C session."""
    from libreasr_b200.api import StreamBatch

    if with_lm:
        cfg, lc, m, orc, olm = lm_model_for("cfg2", "en")
    else:
        cfg, sd, m, orc = model_for("cfg2")
        olm = None
    S, n_chunks = 64, 42
    audio = weights.make_audio(S, n_chunks * CHUNK, seed=131)
    sb = StreamBatch(m.engine(), S, max_iters=10)
    ticks = [[] for _ in range(S)]
    for j in range(n_chunks):
        new = sb.push(torch.from_numpy(np.ascontiguousarray(audio[:, j * CHUNK:(j + 1) * CHUNK])))
        if new is not None:
            for b in range(S):
                ticks[b].append(new[b])
    sb.close()
    assert sum(len(t) for tk in ticks for t in tk) > S * 4   # the streams do emit
    check = range(S) if not with_lm else list(range(0, S, 5)) + [S - 1]
    for b in check:
        assert ticks[b] == _oracle_stream_tokens(orc, cfg, audio[b], n_chunks, olm), f"stream {b}"


def test_cfg4_fixture_8x10s_from_reference():
    """BASELINE configs[3] shape (6 x 1536 LSTM): 8 utterances of 10 s decoded by the imported reference
    (oracle/make_golden.py cfg4_b8), here as ONE batch through Engine.transcribe."""
    g = load_golden("cfg4_b8")
    cfg, sd, m, orc = model_for("cfg4")
    eng = m.engine()
    n_utt = int(g["n_utt"])
    audio = weights.make_audio(n_utt, int(g["n_samples"]), int(g["audio_seed"]))
    r = eng.transcribe(torch.from_numpy(audio).cuda(), max_iters=int(g["max_iters"]))
    for b in range(n_utt):
        assert r["tokens"][b, : int(r["ntok"][b])].tolist() == g[f"tokens_{b}"].tolist(), f"utterance {b}"
        assert r["iters"][b].tolist()[: len(g[f"iters_{b}"])] == g[f"iters_{b}"].tolist(), f"utterance {b}"
        assert abs(float(r["neg_logp"][b]) - float(g[f"neg_log_p_{b}"])) < 5e-3


def test_cfg4_batch_128_equals_per_utterance_runs():
    """configs[3] batch size: 128 utterances (the 8 reference-fixture utterances first) in one call == the same
    utterances run one at a time (sub-batching, ragged lengths) and == the reference fixture."""
    g = load_golden("cfg4_b8")
    cfg, sd, m, orc = model_for("cfg4")
    eng = m.engine()
    n = int(g["n_samples"])
    fix = weights.make_audio(int(g["n_utt"]), n, int(g["audio_seed"]))
    rest = weights.make_audio(120, n, seed=133)
    audio = np.concatenate([fix, rest], 0)
    lens = np.full(128, n, dtype=np.int32)
    lens[8:] = n - 1600 * (np.arange(120) % 50)       # ragged tail; the fixture utterances keep their full length
    a_dev = torch.from_numpy(audio).cuda()
    r = eng.transcribe(a_dev, lens=torch.from_numpy(lens), max_iters=3)
    got = [r["tokens"][b, : int(r["ntok"][b])].tolist() for b in range(128)]
    for b in range(8):
        assert got[b] == g[f"tokens_{b}"].tolist(), f"fixture utterance {b}"
    for b in list(range(8, 128, 7)) + [127]:
        one = eng.transcribe(a_dev[b:b + 1, : int(lens[b])].contiguous(), max_iters=3)
        assert got[b] == one["tokens"][0, : int(one["ntok"][0])].tolist(), f"utterance {b}"


def test_from_config_with_the_reference_testing_yaml():
    """Row a17: enter through Transducer.from_config(conf, lang) with the reference's shipped config/testing.yaml
    (models.py:236-259; key set verbatim), load the reference-keyed state_dict, decode the `ref_offline` fixture."""
    from libreasr_b200.lib.models import Transducer

    with open(os.path.join(GOLDEN_DIR, "testing_conf.json")) as f:
        conf = json.load(f)
    assert conf["cuda"]["device"].startswith("cuda")
    m = Transducer.from_config(copy.deepcopy(conf), Lang())
    assert sum(p.numel() for p in m.parameters()) == 69_829_120      # the reference's parameter count for this conf
    e = m._ecfg
    assert (e.n_mels, e.n_fft, e.win_length, e.hop_length, e.n_stack, e.downsample) == (128, 1024, 400, 160, 10, 8)
    g = load_golden("ref_offline")
    cfg = weights.CONFIGS["ref"]
    sd = weights.make_state_dict(cfg, int(g["weight_seed"]))
    m.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()}, strict=True)
    audio = weights.make_audio(int(g["n_utt"]), int(g["n_samples"]), int(g["audio_seed"]))
    feats = m.engine().features(torch.from_numpy(audio[0:1]).cuda())[0]
    toks, nlp, metrics, extra = m.decode_greedy(feats.unsqueeze(-1), max_iters=int(g["max_iters"]))
    assert toks == g["tokens_0"].tolist()
    assert extra["iters"] == g["iters_0"].tolist()
    # a conf the built path does not cover fails loudly instead of decoding with other settings
    bad = copy.deepcopy(conf)
    bad["model"]["encoder"]["layer_norm"] = True
    with pytest.raises(NotImplementedError):
        Transducer.from_config(bad, Lang())
    bad = copy.deepcopy(conf)
    bad["melkwargs"]["n_mels"] = 80
    with pytest.raises(ValueError):
        Transducer.from_config(bad, Lang())
    m._drop_engine()


def test_stream_session_mid_stream_state_reset_matches_reference_fixture():
    """ADVICE r1 (serve.py:216): the silence reset of the serving loop is the yielded reset_fn (models.py:480-500) --
    model state only; window and Buffer persist.  Fixture `tiny_stream_reset`: the imported reference's
    transcribe_stream with reset_fn called after yields 4 and 11.  The C session with rnnt_b200_stream_reset_state at
    the same points gives the same per-yield token counts and the same transcript; a second stream of the same session
    that is never reset is unaffected."""
    from libreasr_b200.api import StreamBatch

    g = load_golden("tiny_stream_reset")
    cfg, sd, m, orc = model_for("tiny")
    n_chunks = int(g["n_chunks"])
    a = weights.make_audio(1, n_chunks * CHUNK, int(g["audio_seed"]))[0]
    a[: int(g["lead_zero_chunks"]) * CHUNK] = 0.0
    audio = np.stack([a, a])
    reset_after = set(g["reset_after"].tolist())
    sb = StreamBatch(m.engine(), 2, max_iters=10)
    yields, other, n_yield = [], [], 0
    for j in range(n_chunks):
        new = sb.push(torch.from_numpy(np.ascontiguousarray(audio[:, j * CHUNK:(j + 1) * CHUNK])))
        if new is not None:
            yields.append(list(new[0]))
            other.append(list(new[1]))
            if n_yield in reset_after:
                sb.reset_state(0)
            n_yield += 1
    sb.close()
    assert [len(y) for y in yields] == g["chunk_counts"].tolist()
    assert [t for y in yields for t in y] == g["tokens_all"].tolist()
    assert other == _oracle_stream_tokens(orc, cfg, a, n_chunks)       # stream 1: no reset
    assert other != yields


def test_checkpoint_bundle_then_decode(tmp_path):
    """Row f4: libreasr-model-<lang>.tar.gz (fastai `{"model", "opt"}` dict inside `<lang>/model.pth`, model_utils.py:20-95)
    -> load_asr_model -> the tokens of the `tiny_offline` reference fixture."""
    from libreasr_b200.lib import model_utils as MU
    from libreasr_b200.lib.models import Transducer

    g = load_golden("tiny_offline")
    cfg = weights.CONFIGS["tiny"]
    sd = {k: torch.as_tensor(v) for k, v in weights.make_state_dict(cfg, int(g["weight_seed"])).items()}

    def fresh():
        return Transducer(cfg.feature_sz, cfg.embed_sz, cfg.vocab_sz, cfg.hidden_sz, cfg.out_sz, cfg.joint_sz, Lang(),
                          encoder_kwargs={"num_layers": cfg.enc_layers}, predictor_kwargs={"num_layers": cfg.pred_layers})
    src = fresh()
    src.load_state_dict(sd, strict=True)
    dest = tmp_path / "tmp"
    (dest / "xx").mkdir(parents=True)
    torch.save({"model": src.state_dict(), "opt": {}}, dest / "xx" / "model.pth")
    (dest / "xx" / "tokenizer.yttm-model").write_bytes(b"\x00")
    arc = tmp_path / "libreasr-model-xx.tar.gz"
    MU.save_asr_model("xx", path_archive=arc, path_dest=dest)
    MU.extract_tars([str(arc)], path_dest=tmp_path / "out")
    m = MU.load_asr_model(fresh(), "xx", Lang(), device="cuda:0", path_dest=tmp_path / "out")
    audio = weights.make_audio(int(g["n_utt"]), int(g["n_samples"]), int(g["audio_seed"]))
    for b in range(int(g["n_utt"])):
        feats = m.engine().features(torch.from_numpy(audio[b:b + 1]).cuda())[0]
        toks = m.decode_greedy(feats.unsqueeze(-1), max_iters=int(g["max_iters"]))[0]
        assert toks == g[f"tokens_{b}"].tolist()
    m._drop_engine()


def test_forward_lattice_and_rnnt_loss():
    """Row f3: Transducer.forward (models.py:308-359, eval mode) on the GPU == the imported reference's lattice (fixture
    tiny_forward) on the valid region (2e-4), and the RNN-T loss (loss.py:72-110 -> warp_rnnt, restated in oracle/rnnt_loss.py:
    unpinned) == the fp64 CPU recursion; a second, larger random case against the oracle end to end."""
    from oracle import rnnt_loss as RL
    from libreasr_b200.lib.loss import get_loss_func

    g = load_golden("tiny_forward")
    cfg, sd, m, orc = model_for("tiny")
    m.eval()
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    xl, yl = torch.from_numpy(g["xl"]), torch.from_numpy(g["yl"])
    lat = m((x[..., None].cuda(), y.cuda(), xl.cuda(), yl.cuda()))
    ref = torch.from_numpy(g["lattice"])
    assert tuple(lat.shape) == tuple(ref.shape)
    for n in range(x.shape[0]):
        T, U = int(xl[n]), int(yl[n]) + 1
        assert float((lat[n, :T, :U].cpu() - ref[n, :T, :U]).abs().max()) < 2e-4
    want = RL.rnnt_loss(ref, g["y"], g["xl"], g["yl"])
    got = m._last_loss.cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-3)
    loss_fn = get_loss_func("rnnt", m.engine())
    per_seq = loss_fn(lat, (y.cuda(), yl.cuda(), xl.cuda()), reduction="none").cpu().numpy()
    np.testing.assert_allclose(per_seq, want, atol=2e-3)
    assert abs(float(loss_fn(lat, (y.cuda(), yl.cuda(), xl.cuda()))) - want.mean()) < 2e-3
    m.train()
    with pytest.raises(NotImplementedError):
        m((x[..., None].cuda(), y.cuda(), xl.cuda(), yl.cuda()))
    m.eval()
    # larger case, chunked lattice (N*T*U > one 8192-row chunk), against the oracle
    gen = torch.Generator().manual_seed(61)
    N, T, Umax = 5, 60, 33
    x2 = torch.randn(N, T, cfg.feature_sz, generator=gen)
    y2 = torch.randint(3, cfg.vocab_sz, (N, Umax), generator=gen)
    xl2 = torch.tensor([60, 51, 60, 37, 44])
    yl2 = torch.tensor([33, 20, 1, 17, 30])
    for n in range(N):
        y2[n, int(yl2[n]):] = 0
    r = m.engine().forward_loss(x2.cuda(), xl2.cuda(), y2.cuda(), yl2.cuda(), want_lattice=True)
    lat2 = RL.forward_lattice(orc, x2, y2, xl2.numpy(), yl2.numpy())
    for n in range(N):
        Tn, Un = int(xl2[n]), int(yl2[n]) + 1
        assert float((r["lattice"][n, :Tn, :Un].cpu() - lat2[n, :Tn, :Un]).abs().max()) < 3e-4
    np.testing.assert_allclose(r["loss"].cpu().numpy(), RL.rnnt_loss(lat2, y2.numpy(), xl2.numpy(), yl2.numpy()), atol=2e-2, rtol=1e-4)


@pytest.mark.timeout(180, method="thread")
def test_soak_full_size_batches_do_not_stall_and_stay_deterministic():
    """Regression test for a phase-lapping hang of decode_tc2_kernel (all 128 epilogue threads used to wait on the `ctlack`
    mbarrier whose next phase one of them could start: a late warp spun forever about once in 100 full-size calls).
    300 back-to-back calls on the bench workload (32 x 10 s, 8 rotated batches); every repeat of a batch must reproduce
    its tokens."""
    from libreasr_b200 import synth

    cfg, sd, m, orc = model_for("cfg2")
    eng = m.engine()
    n = 160000
    base = weights.make_audio(32, n, seed=synth.BENCH_AUDIO_SEED)
    batches = [torch.from_numpy(np.roll(base, 997 * r, axis=1).copy()).cuda() for r in range(8)]
    first = {}
    for i in range(300):
        r = eng.transcribe(batches[i % 8], max_iters=3)
        got = (r["ntok"].tolist(), r["tokens"][:, :8].tolist())
        if i % 8 in first:
            assert got == first[i % 8], f"call {i}"
        first[i % 8] = got
    assert sum(sum(f[0]) for f in first.values()) > 0


@pytest.mark.timeout(240, method="thread")
@pytest.mark.parametrize("name,B,n", [("tiny", 5, 24000), ("cfg2", 6, 48000), ("cfg2", 32, 160000)])
def test_two_deep_pipeline_equals_plain_calls(name, B, n):
    """rnnt_b200_pipeline_submit / _collect (host->device copy and front end of batch i+1 on a side stream under the
    recurrent kernels of batch i): tokens, counts and scores of every batch equal the plain end-to-end call on the same
    batch, for host and device inputs and ragged lengths (incl. the bench size, 24 batches back to back); the plain calls
    refuse to run while a slot is in flight."""
    cfg, sd, m, orc = model_for(name)
    eng = m.engine()
    nb = 5
    batches = [weights.make_audio(B, n, seed=300 + i) for i in range(nb)]
    lens = [None, torch.tensor([n - 800 * (b % 7) for b in range(B)], dtype=torch.int32), None, None,
            torch.tensor([n - 1600 * (b % 3) for b in range(B)], dtype=torch.int32)]
    want = []
    for a, l in zip(batches, lens):
        r = eng.transcribe_host(torch.from_numpy(a).pin_memory(), lens_host=l, max_iters=3)
        want.append({k: v.clone() for k, v in r.items()})
    reps = 5 if B == 32 else 1     # the bench size: 25 batches back to back
    host = [torch.from_numpy(a).pin_memory() for a in batches]
    got = list(eng.transcribe_pipelined(host * reps, max_iters=3, lens=lens * reps))
    dev = [h.cuda() for h in host]
    dlens = [None if l is None else l.cuda() for l in lens]
    got_dev = list(eng.transcribe_pipelined(dev * reps, max_iters=3, lens=dlens * reps))
    for j in range(nb * reps):
        i = j % nb
        for g in (got[j], got_dev[j]):
            assert g["ntok"].tolist() == want[i]["ntok"].tolist(), f"batch {j}"
            for b in range(B):
                k = int(want[i]["ntok"][b])
                assert g["tokens"][b, :k].tolist() == want[i]["tokens"][b, :k].tolist(), f"batch {j} utterance {b}"
            assert torch.equal(g["neg_logp"], want[i]["neg_logp"]), f"batch {j}"
    assert sum(int(w["ntok"].sum()) for w in want) > 0
    # a slot in flight blocks the plain path and a second submit on the same slot
    out = eng.pipeline_submit(host[0], 0)
    with pytest.raises(RuntimeError):
        eng.transcribe_host(host[1])
    with pytest.raises(RuntimeError):
        eng.pipeline_submit(host[1], 0)
    eng.pipeline_collect(0)
    assert out["ntok"].tolist() == want[0]["ntok"].tolist()
    with pytest.raises(RuntimeError):
        eng.pipeline_collect(0)
    eng.transcribe_host(host[1])
