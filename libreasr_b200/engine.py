"""Thin, typed Python veneer over the C ABI: one ``Engine`` = one ``rnnt_b200_handle``.

PyTorch is used for device memory and CUDA streams only; every array handed to the
library is a contiguous CUDA tensor whose ``data_ptr()`` is passed through, and all
kernels are enqueued on ``torch.cuda.current_stream()``.
"""
import ctypes as C
import math
from dataclasses import dataclass

import numpy as np
import torch

from . import _capi


@dataclass
class EngineConfig:
    """Python mirror of ``rnnt_b200_config``; names follow config/testing.yaml."""

    n_mels: int = 128
    n_stack: int = 10
    downsample: int = 8
    enc_layers: int = 6
    pred_layers: int = 2
    hidden_sz: int = 1024
    embed_sz: int = 512
    joint_sz: int = 1024
    vocab_sz: int = 2048
    blank: int = 0
    bos: int = 2
    sample_rate: int = 16000
    n_fft: int = 1024
    win_length: int = 400
    hop_length: int = 160
    gemm_mode: int = _capi.GEMM_TC_FP16X3
    log_offset: float = 1e-6
    ln_eps: float = 1e-5
    bn_eps: float = 1e-5
    # fused language model (config/testing.yaml:293-313, lm.py); lm_layers == 0: no LM (``m.lm is None``)
    lm_layers: int = 0
    lm_hidden_sz: int = 768
    lm_embed_sz: int = 768
    lm_alpha: float = 0.1
    lm_theta: float = 1.0

    @property
    def feature_sz(self):
        return self.n_mels * self.n_stack

    @staticmethod
    def from_conf(conf, **over):
        """From a reference-style ``conf`` dict (config/testing.yaml keys)."""
        m = conf["model"]
        sr = conf.get("sr", 16000)
        st = _find_stack_args(conf)
        kw = dict(
            n_mels=conf.get("melkwargs", {}).get("n_mels", 128), n_fft=conf.get("melkwargs", {}).get("n_fft", 1024),
            win_length=int(conf.get("win_length", 0.025) * sr), hop_length=int(conf.get("hop_length", 0.01) * sr),
            sample_rate=sr, n_stack=st.get("n_stack", 10), downsample=st.get("downsample", 8),
            enc_layers=m["encoder"]["num_layers"], pred_layers=m["predictor"]["num_layers"],
            hidden_sz=m["hidden_sz"], embed_sz=m["embed_sz"], joint_sz=m["joint_sz"], vocab_sz=m["vocab_sz"],
        )
        lm = conf.get("lm") or {}
        if lm.get("enable"):  # config/testing.yaml:293-299 (+ the per-language override :306-313)
            kw.update(lm_layers=lm["num_layers"], lm_hidden_sz=lm["hidden_sz"], lm_embed_sz=lm["embed_sz"])
        kw.update(over)
        return EngineConfig(**kw)


def _find_stack_args(conf):
    try:
        for t in conf["transforms"]["x"]:
            if t.get("name") == "StackDownsample":
                return t.get("args", {})
    except (KeyError, TypeError):
        pass
    return {}


def melscale_fbanks_htk(n_freqs, n_mels, sample_rate):
    """[n_freqs, n_mels] HTK triangular filterbank with f_min=0, f_max=sr/2, norm=None --
    the table ``torchaudio.transforms.MelSpectrogram`` builds for the reference's call
    (transforms.py:290-296).  Passed to the library as the "frontend.mel_fb" tensor."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + 0.0 / 700.0)
    m_max = 2595.0 * math.log10(1.0 + (float(sample_rate // 2) / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Engine:
    def __init__(self, cfg: EngineConfig, device="cuda:0"):
        self.lib = _capi.load_library()
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("libreasr_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        c = _capi.Config()
        self.lib.rnnt_b200_default_config(C.byref(c))
        for k in ("n_mels", "n_stack", "downsample", "enc_layers", "pred_layers", "hidden_sz", "embed_sz", "joint_sz",
                  "vocab_sz", "blank", "bos", "sample_rate", "n_fft", "win_length", "hop_length", "gemm_mode",
                  "log_offset", "ln_eps", "bn_eps", "lm_layers", "lm_hidden_sz", "lm_embed_sz", "lm_alpha", "lm_theta"):
            setattr(c, k, getattr(cfg, k))
        c.device = self.device.index or 0
        self._h = C.c_void_p(0)
        st = self.lib.rnnt_b200_create(C.byref(c), C.byref(self._h))
        if st != 0:
            _capi.check(self.lib, None, st)
        self.finalized = False

    # ---- lifecycle ------------------------------------------------------------------
    def close(self):
        """Destroys the handle.  Streaming sessions opened on it are closed first (the C ABI refuses to destroy a handle that
        sessions still reference), so a StreamBatch can never outlive the engine it points into."""
        for ref in list(getattr(self, "_sessions", ())):
            sb = ref()
            if sb is not None:
                sb.close()
        self._sessions = []
        if getattr(self, "_h", None) and self._h.value:
            self.lib.rnnt_b200_destroy(self._h)
            self._h = C.c_void_p(0)

    def _register_session(self, sb):
        import weakref

        if not hasattr(self, "_sessions"):
            self._sessions = []
        self._sessions = [r for r in self._sessions if r() is not None]
        self._sessions.append(weakref.ref(sb))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        _capi.check(self.lib, self._h, st)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def load_state_dict(self, sd, lm_state_dict=None):
        """``sd``: reference ``state_dict`` (name -> tensor / ndarray).  Adds the two
        front-end tensors (hann window, mel filterbank) and finalizes.  ``lm_state_dict``: the
        ``LM.state_dict()`` of the fused language model (lm.py:20-29), required iff ``cfg.lm_layers > 0``."""
        cfg = self.cfg
        items = dict(sd)
        if (lm_state_dict is not None) != (cfg.lm_layers > 0):
            raise ValueError("lm_state_dict must be given exactly when cfg.lm_layers > 0")
        for k, v in (lm_state_dict or {}).items():
            items["lm." + k] = v
        items["frontend.window"] = torch.hann_window(cfg.win_length, periodic=True)
        items["frontend.mel_fb"] = melscale_fbanks_htk(cfg.n_fft // 2 + 1, cfg.n_mels, cfg.sample_rate)
        for name, v in items.items():
            if name.endswith("num_batches_tracked"):
                continue
            a = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
            a = np.ascontiguousarray(a, dtype=np.float32)
            self._ck(self.lib.rnnt_b200_set_weight(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), a.size))
        with torch.cuda.device(self.device):
            self._ck(self.lib.rnnt_b200_finalize(self._h, self._stream()))
        self.finalized = True
        return self

    def reserve(self, max_batch, max_samples):
        self._ck(self.lib.rnnt_b200_reserve(self._h, max_batch, max_samples))

    # ---- LM fuser state (streaming; lm.py:43-48,81-83) ------------------------------------
    def new_lm_state(self, B):
        """A zero-filled (= fresh fuser) device blob carrying the LM state of B streams between decode calls."""
        n = C.c_int64(0)
        self._ck(self.lib.rnnt_b200_lm_state_bytes(self._h, B, C.byref(n)))
        blob = torch.zeros(n.value // 4, dtype=torch.float32, device=self.device)
        blob._rnnt_streams = B  # the library checks the stream count of every call against the registration
        return blob

    def set_lm_state(self, blob, B=0):
        """Registers ``blob`` (from ``new_lm_state``) for the following decode calls; ``None`` = fresh fuser per call."""
        self._ck(self.lib.rnnt_b200_set_lm_state(self._h, _ptr(blob), getattr(blob, "_rnnt_streams", B) if blob is not None else 0))
        self._lm_blob = blob  # keep it alive while registered

    # ---- shapes ------------------------------------------------------------------------
    def num_frames(self, n):
        return int(self.lib.rnnt_b200_num_frames(self._h, n))

    def num_steps(self, n):
        return int(self.lib.rnnt_b200_num_steps(self._h, n))

    def _f32(self, t):
        assert t.is_cuda and t.dtype == torch.float32, "expected a float32 CUDA tensor"
        return t.contiguous()

    def _i32(self, t):
        if t is None:
            return None
        return t.to(device=self.device, dtype=torch.int32).contiguous()

    # ---- features ------------------------------------------------------------------------
    def features(self, audio, lens=None):
        """audio [B, n] -> [B, T, X] stacked log-mel (TransformTime + StackDownsample)."""
        audio = self._f32(audio)
        B, n = audio.shape
        lens = self._i32(lens)
        out = torch.empty(B, max(self.num_steps(n), 0), self.cfg.feature_sz, device=self.device)
        self._ck(self.lib.rnnt_b200_features(self._h, _ptr(audio), _ptr(lens), B, n, _ptr(out), self._stream()))
        return out

    def logmel(self, audio, lens=None):
        """audio [B, n] -> [B, F, n_mels] (TransformTime alone)."""
        audio = self._f32(audio)
        B, n = audio.shape
        lens = self._i32(lens)
        out = torch.empty(B, self.num_frames(n), self.cfg.n_mels, device=self.device)
        self._ck(self.lib.rnnt_b200_logmel(self._h, _ptr(audio), _ptr(lens), B, n, _ptr(out), self._stream()))
        return out

    def features_stream(self, window):
        """window [B, W] (3-chunk serving window) -> [B, X] (one stacked row)."""
        window = self._f32(window)
        B, W = window.shape
        out = torch.empty(B, self.cfg.feature_sz, device=self.device)
        self._ck(self.lib.rnnt_b200_features_stream(self._h, _ptr(window), B, W, _ptr(out), self._stream()))
        return out

    # ---- resampling (Resample, transforms.py:135-144) ---------------------------------------------
    def resample(self, audio, orig_sr):
        """audio [B, n] at ``orig_sr`` -> [B, ceil(sample_rate * n / orig_sr)] at the model's sample rate."""
        audio = self._f32(audio)
        B, n = audio.shape
        L = int(self.lib.rnnt_b200_resample_len(self._h, n, int(orig_sr)))
        out = torch.empty(B, L, device=self.device)
        self._ck(self.lib.rnnt_b200_resample(self._h, _ptr(audio), B, n, int(orig_sr), _ptr(out), self._stream()))
        return out

    # ---- encoder -------------------------------------------------------------------------
    def encode(self, feats, lens_T=None, state=None, want_state=False):
        """feats [B, T, X]; state = (h [L,B,H], c [L,B,H]) or None.  Returns (enc [B,T,H], state|None)."""
        feats = self._f32(feats)
        B, T, X = feats.shape
        if X != self.cfg.feature_sz:
            raise ValueError(f"feature size mismatch; expected {self.cfg.feature_sz} got {X}")
        L, H = self.cfg.enc_layers, self.cfg.hidden_sz
        lens_T = self._i32(lens_T)
        use_in = 0
        sh = sc = None
        if state is not None:
            sh, sc = self._f32(state[0]).clone(), self._f32(state[1]).clone()
            if tuple(sh.shape) != (L, B, H) or tuple(sc.shape) != (L, B, H):
                raise ValueError(f"RNN state size mismatch; expected {(L, B, H)} got {tuple(sh.shape)}")
            use_in = 1
        elif want_state:
            sh = torch.empty(L, B, H, device=self.device)
            sc = torch.empty(L, B, H, device=self.device)
        out = torch.empty(B, T, H, device=self.device)
        self._ck(self.lib.rnnt_b200_encode(self._h, _ptr(feats), _ptr(lens_T), B, T, _ptr(sh), _ptr(sc), use_in,
                                           _ptr(out), self._stream()))
        return out, ((sh, sc) if sh is not None else None)

    # ---- predictor / joint ------------------------------------------------------------------
    def predict(self, tokens, state=None):
        """tokens [B] -> (out [B,H], state [Lp,B,H])."""
        tokens = self._i32(tokens)
        B = tokens.shape[0]
        Lp, H = self.cfg.pred_layers, self.cfg.hidden_sz
        if state is None:
            st, use_in = torch.empty(Lp, B, H, device=self.device), 0
        else:
            st, use_in = self._f32(state).clone(), 1
            if tuple(st.shape) != (Lp, B, H):
                raise ValueError(f"RNN state size mismatch; expected {(Lp, B, H)} got {tuple(st.shape)}")
        out = torch.empty(B, H, device=self.device)
        self._ck(self.lib.rnnt_b200_predict(self._h, _ptr(tokens), B, _ptr(st), use_in, _ptr(out), self._stream()))
        return out, st

    def joint(self, h_pred, h_enc):
        """[B,H], [B,H] -> logits [B,V]."""
        h_pred, h_enc = self._f32(h_pred), self._f32(h_enc)
        B = h_pred.shape[0]
        out = torch.empty(B, self.cfg.vocab_sz, device=self.device)
        self._ck(self.lib.rnnt_b200_joint(self._h, _ptr(h_pred), _ptr(h_enc), B, _ptr(out), self._stream()))
        return out

    # ---- decode ----------------------------------------------------------------------------------
    def decode_greedy(self, enc, lens_T=None, max_iters=3, state=None, want_state=False, trace_cap=0, lm_state=None):
        """enc [B,T,H].  state = (pred_h [Lp,B,H], pred_out [B,H]) or None (-> BOS from the
        learnable state).  lm_state: blob from ``new_lm_state(B)`` carrying the LM fuser of B streams
        across calls (updated in place); None = fresh fuser (offline decode).  Returns dict of device tensors:
        tokens [B,U], ntok [B], neg_logp [B] (f64), iters [B,T] (u8), trace [B,trace_cap,V] | None, state | None."""
        enc = self._f32(enc)
        B, T, H = enc.shape
        Lp, V = self.cfg.pred_layers, self.cfg.vocab_sz
        lens_T = self._i32(lens_T)
        U = max_iters * T
        tokens = torch.zeros(B, U, dtype=torch.int32, device=self.device)
        ntok = torch.zeros(B, dtype=torch.int32, device=self.device)
        nlp = torch.zeros(B, dtype=torch.float64, device=self.device)
        iters = torch.zeros(B, T, dtype=torch.uint8, device=self.device)
        trace = torch.zeros(B, trace_cap, V, device=self.device) if trace_cap > 0 else None
        use_in = 0
        ph = po = None
        if state is not None:
            ph, po = self._f32(state[0]).clone(), self._f32(state[1]).clone()
            use_in = 1
        elif want_state:
            ph = torch.empty(Lp, B, H, device=self.device)
            po = torch.empty(B, H, device=self.device)
        if lm_state is not None:
            self.set_lm_state(lm_state, B)
        try:
            self._ck(self.lib.rnnt_b200_decode_greedy(
                self._h, _ptr(enc), _ptr(lens_T), B, T, max_iters, _ptr(ph), _ptr(po), use_in, _ptr(tokens), U,
                _ptr(ntok), _ptr(nlp), _ptr(iters), _ptr(trace), trace_cap, self._stream()))
        finally:
            if lm_state is not None:
                self.set_lm_state(None)
        return {"tokens": tokens, "ntok": ntok, "neg_logp": nlp, "iters": iters, "trace": trace,
                "state": (ph, po) if ph is not None else None}

    # ---- whole path --------------------------------------------------------------------------------
    # ---- training-time forward: joint lattice + RNN-T loss (eval mode, no gradients) ---------------------
    def forward_loss(self, feats, lens_T, labels, label_lens, want_lattice=False):
        """``Transducer.forward`` (models.py:308-359, eval mode) + the RNN-T loss of loss.py:72-110 on its output.
        feats [N,T,X], lens_T [N] | None, labels [N,Umax] (padded), label_lens [N] ->
        dict(loss [N] fp64 = -log p(labels | audio), lattice [N,T,Umax+1,V] log-probabilities | None)."""
        feats = self._f32(feats)
        N, T, X = feats.shape
        if X != self.cfg.feature_sz:
            raise ValueError(f"feature size mismatch; expected {self.cfg.feature_sz} got {X}")
        labels = self._i32(labels)
        Umax = labels.shape[1]
        lens_T, label_lens = self._i32(lens_T), self._i32(label_lens)
        loss = torch.zeros(N, dtype=torch.float64, device=self.device)
        lat = torch.empty(N, T, Umax + 1, self.cfg.vocab_sz, device=self.device) if want_lattice else None
        self._ck(self.lib.rnnt_b200_forward_loss(self._h, _ptr(feats), _ptr(lens_T), _ptr(labels), _ptr(label_lens), N, T, Umax,
                                                 _ptr(lat), _ptr(loss), self._stream()))
        return {"loss": loss, "lattice": lat}

    def rnnt_loss(self, lattice, lens_T, labels, label_lens):
        """Per-sequence RNN-T negative log-likelihood from a log-probability lattice [N,T,U,V] (U = label columns + 1)."""
        lattice = self._f32(lattice)
        N, T, U, V = lattice.shape
        if V != self.cfg.vocab_sz:
            raise ValueError("lattice vocabulary size mismatch")
        labels = self._i32(labels)
        if labels.shape[1] != U - 1:
            raise ValueError(f"labels must have U - 1 = {U - 1} columns")
        loss = torch.zeros(N, dtype=torch.float64, device=self.device)
        self._ck(self.lib.rnnt_b200_rnnt_loss(self._h, _ptr(lattice), _ptr(self._i32(lens_T)), _ptr(labels), _ptr(self._i32(label_lens)),
                                              N, T, U, _ptr(loss), self._stream()))
        return loss

    def decode_beam(self, enc, lens_T=None, width=4, max_iters=3):
        """Beam search over encoder output [B,T,H] (not in the reference; algorithm: oracle/beam.py).  Returns the best
        hypothesis per utterance: dict(tokens [B,U], ntok [B], score [B] fp64 log-probability)."""
        enc = self._f32(enc)
        B, T, H = enc.shape
        U = max_iters * max(T, 1)
        lens_T = self._i32(lens_T)
        tokens = torch.zeros(B, U, dtype=torch.int32, device=self.device)
        ntok = torch.zeros(B, dtype=torch.int32, device=self.device)
        score = torch.zeros(B, dtype=torch.float64, device=self.device)
        self._ck(self.lib.rnnt_b200_decode_beam(self._h, _ptr(enc), _ptr(lens_T), B, T, int(width), int(max_iters), _ptr(tokens), U,
                                                _ptr(ntok), _ptr(score), self._stream()))
        return {"tokens": tokens, "ntok": ntok, "score": score}

    def transcribe_beam(self, audio, lens=None, width=4, max_iters=3):
        """Device-resident audio [B, n] -> features -> encoder -> beam search."""
        audio = self._f32(audio)
        feats = self.features(audio, lens)
        lens_T = None
        if lens is not None:
            lens_T = torch.clamp((self._i32(lens).to(torch.int64) // self.cfg.hop_length + 1 - self.cfg.n_stack) // self.cfg.downsample + 1, min=0).to(torch.int32)
        enc, _ = self.encode(feats, lens_T)
        return self.decode_beam(enc, lens_T, width, max_iters)

    def transcribe(self, audio, lens=None, max_iters=3):
        """Device-resident audio [B, n] -> dict(tokens, ntok, neg_logp, iters) device tensors."""
        audio = self._f32(audio)
        B, n = audio.shape
        T = self.num_steps(n)
        U = max_iters * max(T, 1)
        lens = self._i32(lens)
        tokens = torch.zeros(B, U, dtype=torch.int32, device=self.device)
        ntok = torch.zeros(B, dtype=torch.int32, device=self.device)
        nlp = torch.zeros(B, dtype=torch.float64, device=self.device)
        iters = torch.zeros(B, max(T, 1), dtype=torch.uint8, device=self.device)
        self._ck(self.lib.rnnt_b200_transcribe(self._h, _ptr(audio), _ptr(lens), B, n, max_iters, _ptr(tokens), U,
                                               _ptr(ntok), _ptr(nlp), _ptr(iters), self._stream()))
        return {"tokens": tokens, "ntok": ntok, "neg_logp": nlp, "iters": iters}

    def transcribe_host(self, audio_host, lens_host=None, max_iters=3, out=None):
        """HOST audio [B, n] (pin it for full speed) -> host tensors; the end-to-end call."""
        assert not audio_host.is_cuda and audio_host.dtype == torch.float32 and audio_host.is_contiguous()
        B, n = audio_host.shape
        T = self.num_steps(n)
        U = max_iters * max(T, 1)
        if out is None:
            out = self.alloc_host_outputs(B, U)
        lens_p = C.c_void_p(lens_host.data_ptr()) if lens_host is not None else C.c_void_p(0)
        with torch.cuda.device(self.device):
            self._ck(self.lib.rnnt_b200_transcribe_host(
                self._h, C.c_void_p(audio_host.data_ptr()), lens_p, B, n, max_iters,
                C.c_void_p(out["tokens"].data_ptr()), U, C.c_void_p(out["ntok"].data_ptr()),
                C.c_void_p(out["neg_logp"].data_ptr()), self._stream()))
        return out

    # ---- two-deep pipeline (throughput): front end of batch i+1 under the recurrent kernels of batch i ----
    def pipeline_submit(self, audio, slot, lens=None, max_iters=3, out=None):
        """Queue one batch on pipeline slot 0/1 (rnnt_b200_pipeline_submit).  ``audio`` [B, n] fp32: a CUDA tensor, or a
        host tensor (pin it) that is copied inside the pipeline.  ``lens`` lives where ``audio`` lives.  Returns the host
        output dict, valid after ``pipeline_collect(slot)``."""
        assert audio.dtype == torch.float32 and audio.is_contiguous() and audio.dim() == 2
        on_host = not audio.is_cuda
        if lens is not None:
            assert lens.dtype == torch.int32 and lens.is_cuda == audio.is_cuda
        B, n = audio.shape
        U = max_iters * max(self.num_steps(n), 1)
        if out is None:
            out = self.alloc_host_outputs(B, U)
        lens_p = C.c_void_p(lens.data_ptr()) if lens is not None else C.c_void_p(0)
        with torch.cuda.device(self.device):
            self._ck(self.lib.rnnt_b200_pipeline_submit(
                self._h, C.c_void_p(audio.data_ptr()), 1 if on_host else 0, lens_p, B, n, max_iters, slot,
                C.c_void_p(out["tokens"].data_ptr()), out["tokens"].shape[1], C.c_void_p(out["ntok"].data_ptr()),
                C.c_void_p(out["neg_logp"].data_ptr()), self._stream()))
        return out

    def pipeline_collect(self, slot):
        self._ck(self.lib.rnnt_b200_pipeline_collect(self._h, slot))

    def pipeline_query(self, slot):
        """(front_done, back_done) of a slot's batch, without blocking."""
        f, b = C.c_int32(0), C.c_int32(0)
        self._ck(self.lib.rnnt_b200_pipeline_query(self._h, slot, C.byref(f), C.byref(b)))
        return bool(f.value), bool(b.value)

    def transcribe_pipelined(self, batches, max_iters=3, lens=None):
        """Run a sequence of [B, n] batches (host or device tensors) through the two-deep pipeline; yields one host
        output dict per batch, in order."""
        pending = None
        for i, a in enumerate(batches):
            out = self.pipeline_submit(a, i & 1, lens=None if lens is None else lens[i], max_iters=max_iters)
            if pending is not None:
                self.pipeline_collect(pending[0])
                yield pending[1]
            pending = (i & 1, out, a)   # the input stays referenced until its batch has been collected (asynchronous copy)
        if pending is not None:
            self.pipeline_collect(pending[0])
            yield pending[1]

    @staticmethod
    def alloc_host_outputs(B, U, pin=True):
        mk = (lambda *s, dtype: torch.zeros(*s, dtype=dtype).pin_memory()) if pin else (lambda *s, dtype: torch.zeros(*s, dtype=dtype))
        return {"tokens": mk(B, U, dtype=torch.int32), "ntok": mk(B, dtype=torch.int32), "neg_logp": mk(B, dtype=torch.float64)}

    def selftest_gemm(self, A, W, bias=None, gemm_mode=None):
        """C = A @ W.T + bias with the library's GEMM in the given arithmetic mode (test hook)."""
        A, W = self._f32(A), self._f32(W)
        M, K = A.shape
        N = W.shape[0]
        out = torch.empty(M, N, device=self.device)
        mode = self.cfg.gemm_mode if gemm_mode is None else gemm_mode
        self._ck(self.lib.rnnt_b200_selftest_gemm(self._h, _ptr(A), _ptr(W), _ptr(bias), _ptr(out), M, N, K, mode, self._stream()))
        return out

    # ---- introspection ---------------------------------------------------------------------------------
    def kernel_launches(self):
        return int(self.lib.rnnt_b200_kernel_launches(self._h))

    def fp32_decode_launches(self):
        """Launches of the fp32 cooperative decode kernel (the slow twin): 0 on the tensor-core path."""
        return int(self.lib.rnnt_b200_fp32_decode_launches(self._h))

    def set_profiling(self, on):
        self._ck(self.lib.rnnt_b200_set_profiling(self._h, 1 if on else 0))

    def stage_times_ms(self):
        buf = (C.c_float * 5)()
        self._ck(self.lib.rnnt_b200_stage_times_ms(self._h, buf))
        return dict(zip(("features", "encoder", "encoder_input_gemms", "joint_enc_gemm", "decode"), [float(x) for x in buf]))


def tokens_to_lists(tokens, ntok):
    """Device/host [B,U] + [B] -> list of python int lists."""
    t, n = tokens.cpu().numpy(), ntok.cpu().numpy()
    return [t[b, : int(n[b])].tolist() for b in range(t.shape[0])]
