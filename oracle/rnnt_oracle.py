"""TEST INFRASTRUCTURE ONLY -- CPU (torch fp32) restatement of the reference's
streaming RNN-T inference path.  Never imported by ``libreasr_b200``.

Every function cites the reference file:line it restates (paths relative to
``/root/reference``).  Third-party arithmetic the reference calls but does not
vendor is restated from its published definition and named here:

* ``torchaudio.transforms.MelSpectrogram`` (pinned torchaudio==0.6.0,
  docker/requirements.inference.txt:5; call site transforms.py:290-296):
  ``torch.stft`` (n_fft 1024, hop 160, periodic hann(400) zero-padded to 1024,
  center + reflect padding, power 2) followed by an HTK triangular filterbank
  (``melscale_fbanks``: f in [0, sr/2], norm=None).
* ``torch.nn.LSTM`` (pinned torch==1.6.0+cpu, requirements.inference.txt:3; call
  sites custom_rnn.py:26-45,244,268): gate order i,f,g,o; the explicit arithmetic
  is the one spelled out in the reference's own haste/lstm.py:34-68.
* ``nn.LayerNorm`` / ``nn.BatchNorm1d`` (eval) / ``nn.Linear`` / ``nn.Embedding`` /
  ``F.log_softmax``: standard definitions.

Pinned: ``tests/test_oracle_golden.py`` checks this module against fixtures
produced by the imported, unmodified reference (``oracle/make_golden.py``).
Beam search does not exist in the reference and is not restated here.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .weights import ModelConfig

LOG_OFFSET = 1e-6  # transforms.py:311-313
LN_EPS = 1e-5  # nn.LayerNorm default (models.py:84)
BN_EPS = 1e-5  # nn.BatchNorm1d default (custom_rnn.py:124)
BLANK = 0  # models.py:203,225
BOS = 2  # models.py:226-227


# ----------------------------------------------------------------------------
# features (a3, a4, a5, a6)
# ----------------------------------------------------------------------------
def resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """Polyphase windowed-sinc filter bank of ``torchaudio.transforms.Resample(orig_freq, new_freq)`` with its defaults
    (sinc_interp_hann) -- the call the reference makes per utterance (transforms.py:135-144; the pinned torchaudio 0.6.0 routed
    it through kaldi.resample_waveform, the same algorithm).  Returns (kernels [new, K] float32, width, orig/gcd, new/gcd)."""
    import math

    g = math.gcd(int(orig_freq), int(new_freq))
    o, n = int(orig_freq) // g, int(new_freq) // g
    base = min(o, n) * rolloff
    width = math.ceil(lowpass_filter_width * o / base)
    idx = np.arange(-width, width + o, dtype=np.float64)[None, :] / o
    # the phase term is formed in float32 (int64 / python int under torch's default dtype) before it meets the float64 grid
    t = (np.arange(0, -n, -1).astype(np.float32) / np.float32(n)).astype(np.float64)[:, None] + idx
    t = t * base
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    k = k * window * (base / o)
    return k.astype(np.float32), width, o, n


def resample(audio: torch.Tensor, orig_freq: int, new_freq: int = 16000) -> torch.Tensor:
    """``Resample.encodes`` (transforms.py:135-144): audio [C, n] at orig_freq -> [C, ceil(new * n / orig)]."""
    if int(orig_freq) == int(new_freq):
        return audio
    k, width, o, n = resample_kernel(orig_freq, new_freq)
    x = F.pad(audio, (width, width + o))
    y = F.conv1d(x[:, None], torch.from_numpy(k)[:, None], stride=o)          # [C, new, frames]
    y = y.transpose(1, 2).reshape(audio.shape[0], -1)
    return y[:, : int(np.ceil(n * audio.shape[1] / o))]


def mel_fbanks_htk(n_freqs: int, n_mels: int, sample_rate: int) -> torch.Tensor:
    """[n_freqs, n_mels] HTK triangular filterbank, f_min=0, f_max=sr/2, norm=None
    (what ``MelSpectrogram(sample_rate, n_fft, n_mels)`` builds by default;
    transforms.py:290-296 passes only sr/win/hop/n_fft/n_mels)."""
    f_min, f_max = 0.0, float(sample_rate // 2)
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


def log_mel(audio: torch.Tensor, cfg: ModelConfig) -> torch.Tensor:
    """``TransformTime.encodes`` (transforms.py:301-323) with deltas=0
    (config/testing.yaml:143): [C, n] f32 -> [C, F, M], F = n // hop + 1."""
    window = torch.hann_window(cfg.win_length, periodic=True)
    spec = torch.stft(
        audio, n_fft=cfg.n_fft, hop_length=cfg.hop_length, win_length=cfg.win_length,
        window=window, center=True, pad_mode="reflect", normalized=False, onesided=True,
        return_complex=True,
    )
    power = spec.real * spec.real + spec.imag * spec.imag  # power=2.0
    fb = mel_fbanks_htk(cfg.n_fft // 2 + 1, cfg.n_mels, cfg.sample_rate)
    mel = torch.matmul(power.transpose(-1, -2), fb).transpose(-1, -2)  # [C, M, F]
    res = torch.log(mel + LOG_OFFSET)
    return res.permute(0, 2, 1).contiguous()


def stream_postprocess(spectro: torch.Tensor, n_stack: int) -> torch.Tensor:
    """``StreamPostprocess.encodes`` (transforms.py:335-342)."""
    t = spectro.shape[1]
    a = t // 3 + 1
    return spectro[:, a:, :][:, :n_stack, :]


def stack_downsample(t: torch.Tensor, n_stack: int, downsample: int) -> torch.Tensor:
    """``StackDownsample.encodes`` (transforms.py:436-441): feature index = m*S + s."""
    uf = t.unfold(-2, n_stack, downsample).contiguous()
    return uf.view(uf.size(0), uf.size(1), -1).contiguous()


def features_offline(audio: torch.Tensor, cfg: ModelConfig) -> torch.Tensor:
    """Inference ``x`` pipeline (config/testing.yaml:339-354) for 16 kHz mono input
    (Resample / ChannelCut are identities there): [C, n] -> [C, T, X]
    (``FixDimensions`` only appends a unit axis, transforms.py:444-452)."""
    return stack_downsample(log_mel(audio, cfg), cfg.n_stack, cfg.downsample)


def features_stream_window(window: torch.Tensor, cfg: ModelConfig) -> torch.Tensor:
    """Inference ``stream`` pipeline up to (excluding) ``Buffer``
    (config/testing.yaml:356-374): [C, 3*chunk] -> [C, 1, X]."""
    sp = stream_postprocess(log_mel(window, cfg), cfg.n_stack)
    return stack_downsample(sp, cfg.n_stack, cfg.downsample)


class StreamFrontend:
    """The serving loop's windowing (api-server.py:26,83-115) + ``Buffer``
    (transforms.py:455-471): 3-chunk sliding window, one feature row per chunk once
    3 chunks are buffered, rows released in pairs (``n_buffer`` = 2)."""

    def __init__(self, cfg: ModelConfig, n_window: int = 3, n_buffer: int = 2):
        self.cfg, self.n_window, self.n_buffer = cfg, n_window, n_buffer
        self.frames, self.saved = [], []

    def push(self, chunk: torch.Tensor):
        """chunk [1, n_chunk]; returns [n_buffer, X] or None."""
        self.frames.append(chunk)
        if len(self.frames) < self.n_window:
            return None
        window = torch.cat(self.frames, dim=1)
        self.frames.pop(0)
        self.saved.append(features_stream_window(window, self.cfg))
        if len(self.saved) == self.n_buffer:
            catted = torch.cat(self.saved, dim=1)
            self.saved.clear()
            return catted[0]
        return None


# ----------------------------------------------------------------------------
# model (a7 - a12)
# ----------------------------------------------------------------------------
def _t(a):
    return a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))


class OracleLM:
    """The fused language model and its fuser (reference libreasr/lib/lm.py:20-83).

    ``LM.forward`` (lm.py:31-41): Embedding -> nn.LSTM (zero initial state when none is carried) -> Dropout (identity in
    eval) -> Linear (weights tied to the embedding when embed_sz == hidden_sz, lm.py:27-29) -> log_softmax.
    ``LMFuser.advance`` (lm.py:50-54) standardises that row in place ((t - mean) / (std + 1e-5), unbiased std,
    utils.py:162-164) and pins the blank entry to -10; ``LMFuser.fuse`` (lm.py:56-79) does the same to the joint's
    log_softmax row and takes arg max of ``alpha * lm + theta * joint``.  Before the first ``advance`` there is no LM row
    and ``fuse`` returns its inputs (lm.py:58,79)."""

    MIN_VAL = -10.0  # lm.py:15

    def __init__(self, lm_cfg, state_dict: dict):
        self.cfg = lm_cfg
        self.sd = {k: _t(v) for k, v in state_dict.items()}
        self.reset()

    def reset(self):  # lm.py:81-83
        self.logits = None
        self.state = None
        self.fused_margins = []  # top-1 minus top-2 of every fused row (tie-risk bookkeeping for the tests)

    @staticmethod
    def standardize(t: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:  # utils.py:162-164
        t = t - t.mean()
        return t / (t.std() + eps)

    def forward(self, token: int):
        """One LM step on one token: returns the log_softmax row [V] and updates the carried (h, c) per layer."""
        sd, L, H = self.sd, self.cfg.num_layers, self.cfg.hidden_sz
        x = sd["embed.weight"][token]
        if self.state is None:
            self.state = [(torch.zeros(H), torch.zeros(H)) for _ in range(L)]
        new = []
        for i in range(L):
            h, c = self.state[i]
            v = sd[f"rnn.weight_ih_l{i}"] @ x + sd[f"rnn.bias_ih_l{i}"] + sd[f"rnn.weight_hh_l{i}"] @ h + sd[f"rnn.bias_hh_l{i}"]
            ig, fg, gg, og = v.chunk(4)  # torch.nn.LSTM gate order i, f, g, o
            c = torch.sigmoid(fg) * c + torch.sigmoid(ig) * torch.tanh(gg)
            h = torch.sigmoid(og) * torch.tanh(c)
            new.append((h, c))
            x = h
        self.state = new
        return F.log_softmax(sd["linear.weight"] @ x + sd["linear.bias"], dim=-1)

    def advance(self, token: int):  # lm.py:50-54
        row = self.standardize(self.forward(token))
        row[0] = self.MIN_VAL
        self.logits = row

    def fuse(self, logp: torch.Tensor, pred: int) -> int:  # lm.py:56-79
        if self.logits is None:
            return pred
        j = self.standardize(logp)
        j[0] = self.MIN_VAL
        fused = self.cfg.alpha * self.logits + self.cfg.theta * j
        top2 = torch.topk(fused, 2).values
        self.fused_margins.append(float(top2[0] - top2[1]))
        return int(fused.argmax(-1))


class OracleTransducer:
    """Restates ``Transducer`` inference (models.py:190-577) on a reference ``state_dict``."""

    def __init__(self, cfg: ModelConfig, state_dict: dict):
        self.cfg = cfg
        self.sd = {k: _t(v) for k, v in state_dict.items()}
        self.blank, self.bos = BLANK, BOS

    # -- building blocks -------------------------------------------------------
    def _bn(self, prefix: str, x: torch.Tensor) -> torch.Tensor:
        """BatchNorm1d eval over the channel (last) axis (custom_rnn.py:210-213)."""
        sd = self.sd
        inv = torch.rsqrt(sd[prefix + ".running_var"] + BN_EPS)
        return (x - sd[prefix + ".running_mean"]) * inv * sd[prefix + ".weight"] + sd[prefix + ".bias"]

    def lstm_layer(self, i: int, x: torch.Tensor, state, impl: str = "explicit"):
        """One ``nn.LSTM`` layer over all T (custom_rnn.py:140-175; arithmetic
        haste/lstm.py:51-60 with native gate order i,f,g,o).  x [N,T,in];
        state (h,c) each [1,N,H] or None -> learnable ``hs[i]`` (custom_rnn.py:152-156)."""
        sd, H = self.sd, self.cfg.hidden_sz
        p = f"encoder.rnn_stack.rnns.{i}."
        N, T = x.shape[0], x.shape[1]
        if state is None:
            hs = sd[f"encoder.rnn_stack.hs.{i}"]
            h = hs[0].expand(1, N, H).contiguous()
            c = hs[1].expand(1, N, H).contiguous()
        else:
            h, c = state
        w_ih, w_hh = sd[p + "weight_ih_l0"], sd[p + "weight_hh_l0"]
        b_ih, b_hh = sd[p + "bias_ih_l0"], sd[p + "bias_hh_l0"]
        if impl == "aten":  # what the reference executes (torch.nn.LSTM kernel)
            out, hn, cn = torch._VF.lstm(x, (h, c), [w_ih, w_hh, b_ih, b_hh], True, 1, 0.0, False, False, True)
            return out, (hn, cn)
        h, c = h[0], c[0]
        wx = x @ w_ih.t()  # hoisted input projection (haste/lstm.py:51)
        outs = []
        for t in range(T):
            v = h @ w_hh.t() + wx[:, t] + b_ih + b_hh
            gi, gf, gg, go = torch.chunk(v, 4, 1)
            c = torch.sigmoid(gf) * c + torch.sigmoid(gi) * torch.tanh(gg)
            h = torch.sigmoid(go) * torch.tanh(c)
            outs.append(h)
        return torch.stack(outs, 1), (h[None], c[None])

    def encoder(self, x: torch.Tensor, state=None, impl: str = "explicit"):
        """``Encoder.forward`` (models.py:105-113) + ``CustomRNN.forward``
        (custom_rnn.py:177-232) in eval mode.  x [N,T,X(,1)] -> ([N,T,H], [(h,c)]*L)."""
        sd, cfg = self.sd, self.cfg
        x = x.reshape(x.size(0), x.size(1), -1)
        x = F.layer_norm(x, (cfg.feature_sz,), sd["encoder.input_norm.weight"], sd["encoder.input_norm.bias"], LN_EPS)
        new_states = []
        for i in range(cfg.enc_layers):
            x, st = self.lstm_layer(i, x, None if state is None else state[i], impl)
            x = self._bn(f"encoder.rnn_stack.bns.{i}", x)
            new_states.append(st)
        return x, new_states  # Dropout eval = id; Linear absent since hidden_sz == out_sz (models.py:97-100)

    def gru_cell(self, i: int, x: torch.Tensor, h: torch.Tensor) -> torch.Tensor:
        """``NBRCScript`` one step (haste/nbrc.py:46-56): gate layout z,r,g. x,h [N,H]."""
        sd = self.sd
        p = f"predictor.rnn_stack.rnns.{i}."
        wx = x @ sd[p + "kernel"] + sd[p + "bias"]
        rh = h @ sd[p + "recurrent_kernel"] + sd[p + "recurrent_bias"]
        vx, vh = torch.chunk(wx, 3, 1), torch.chunk(rh, 3, 1)
        z = torch.sigmoid(vx[0] + vh[0])
        r = torch.sigmoid(vx[1] + vh[1])
        g = torch.tanh(vx[2] + r * vh[2])
        return z * h + (1 - z) * g

    def predictor(self, tokens: torch.Tensor, state=None):
        """``Predictor.forward`` for one step (models.py:181-187): tokens [N] int64,
        state list of [1,N,H] or None -> (out [N,H], new state)."""
        sd, cfg = self.sd, self.cfg
        x = sd["predictor.embed.weight"][tokens]
        x = x @ sd["predictor.ffn.weight"].t() + sd["predictor.ffn.bias"]
        N, new_state = x.shape[0], []
        for i in range(cfg.pred_layers):
            if state is None:
                h = sd[f"predictor.rnn_stack.hs.{i}"][0].expand(1, N, cfg.hidden_sz)[0]
            else:
                h = state[i][0]
            h = self.gru_cell(i, x, h)
            new_state.append(h[None])
            x = self._bn(f"predictor.rnn_stack.bns.{i}", h)
        return x, new_state

    def joint(self, h_pred: torch.Tensor, h_enc: torch.Tensor) -> torch.Tensor:
        """``Joint.forward`` concat method (models.py:132-140): [...,O],[...,O] -> logits [...,V]."""
        sd = self.sd
        x = torch.cat((h_pred, h_enc), dim=-1)
        x = torch.tanh(x @ sd["joint.joint.0.weight"].t() + sd["joint.joint.0.bias"])
        return x @ sd["joint.joint.2.weight"].t() + sd["joint.joint.2.bias"]

    # -- decode loops ------------------------------------------------------------
    def decode_greedy(self, x: torch.Tensor, max_iters: int = 3, impl: str = "explicit", keep_logits: bool = False, lm=None):
        """``Transducer.decode_greedy`` (models.py:369-455); ``lm`` = an ``OracleLM`` for shallow fusion (models.py:401,431,440).
        x [T,X] features of ONE utterance.  Returns dict(tokens, neg_log_p, iters,
        alignment_score, margins[, logp])."""
        with torch.no_grad():
            enc, _ = self.encoder(x[None], None, impl)
            enc = enc[0]
            tok = torch.tensor([self.bos])
            h_pred, pstate = self.predictor(tok)
            y_seq, log_p, iters_all, margins, outs = [], 0.0, [], [], []
            if lm is not None:
                lm.reset()  # a fresh LMFuser per call (models.py:401)
            for h_enc in enc:
                iters = 0
                while iters < max_iters:
                    iters += 1
                    logp = F.log_softmax(self.joint(h_pred, h_enc[None]), dim=-1)[0]
                    top2 = torch.topk(logp, 2).values
                    margins.append(float(top2[0] - top2[1]))
                    if keep_logits:
                        outs.append(logp.clone())
                    prob, pred = logp.max(-1)
                    pred = int(pred)
                    log_p += float(prob)
                    if pred == self.blank:
                        break
                    if lm is not None:
                        pred = lm.fuse(logp, pred)  # only after the blank test (models.py:427-431)
                    y_seq.append(pred)
                    h_pred, pstate = self.predictor(torch.tensor([pred]), pstate)
                    if lm is not None:
                        lm.advance(pred)
                iters_all.append(iters)
            align = np.array(iters_all)
            _sum = align.sum()
            ones = int((align == 1).sum())
            res = {
                "tokens": y_seq, "neg_log_p": -log_p, "iters": iters_all,
                "alignment_score": float((_sum - ones) / (_sum + 1e-4)),  # models.py:445-453
                "margins": margins, "enc": enc,
            }
            if keep_logits:
                res["logp"] = torch.stack(outs)
            return res

    def transcribe_stream(self, stream, max_iters: int = 10, impl: str = "explicit", lm=None, reset_after=()):
        """``Transducer.transcribe_stream`` (models.py:457-577; LM fusion at :558,569): generator over
        chunks ([T_c, X] or None) yielding (all tokens so far, this chunk's tokens).  ``reset_after``: yield indices
        after which the consumer calls the yielded ``reset_fn`` (models.py:480-500: encoder state -> None,
        predictor -> BOS, fuser reset; ``y`` keeps accumulating)."""
        with torch.no_grad():
            enc_state = None
            h_pred, pstate = self.predictor(torch.tensor([self.bos]))
            y = []
            n_yield = 0
            if lm is not None:
                lm.reset()
            for chunk in stream:
                if chunk is None:  # models.py:509
                    continue
                enc, enc_state = self.encoder(chunk[None], enc_state, impl)
                y_seq = []
                for i in range(enc.shape[1]):
                    h_enc = enc[0, i]
                    iters = 0
                    while iters < max_iters:
                        iters += 1
                        logp = F.log_softmax(self.joint(h_pred, h_enc[None]), dim=-1)[0]
                        pred = int(logp.argmax(-1))
                        if pred == self.blank:
                            break
                        if lm is not None:
                            pred = lm.fuse(logp, pred)
                        y_seq.append(pred)
                        h_pred, pstate = self.predictor(torch.tensor([pred]), pstate)
                        if lm is not None:
                            lm.advance(pred)
                y = y + y_seq
                yield list(y), list(y_seq)
                if n_yield in reset_after:   # reset() of models.py:494-497, called by the consumer between two chunks
                    enc_state = None
                    h_pred, pstate = self.predictor(torch.tensor([self.bos]))
                    if lm is not None:
                        lm.reset()
                n_yield += 1


def transcribe_batch(model: OracleTransducer, audio: np.ndarray, max_iters: int = 3, impl: str = "aten", lm=None):
    """The reference serving path applied utterance by utterance (api-server.py:64-80;
    the reference has no batched decode, config/testing.yaml:380 bs=1):
    audio [B, n] -> list of token lists."""
    out = []
    for b in range(audio.shape[0]):
        feats = features_offline(torch.from_numpy(audio[b : b + 1]), model.cfg)[0]
        out.append(model.decode_greedy(feats, max_iters=max_iters, impl=impl, lm=lm)["tokens"])
    return out
