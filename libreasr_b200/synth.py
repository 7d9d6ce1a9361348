"""Deterministic synthetic weights and audio (no algorithm of the path lives here).

Used by bench.py / smoke() (random-init weights of the named architecture and synthetic
16 kHz audio, as there is no network for checkpoints or datasets) and re-exported as
``oracle.weights`` for the parity tests.

The reference ships neither pretrained weights (``tmp/``, ``models/`` are
git-ignored, .gitignore:109-113) nor test fixtures, and there is no network.
Both sides of every parity test (the imported reference / the CPU oracle and
the CUDA path) therefore load the SAME synthetic ``state_dict``, generated
here from a seed and the parameter NAME only (order independent), using the
reference's exact key set (SURVEY.md section 8b "Checkpoint contract").

Distributions follow the PyTorch / Haste default initialisers used by the
reference constructors, except where the defaults would leave a code path
unexercised (SURVEY.md section 8d "Weights"):
  * BatchNorm running stats are randomised so the eval-mode affine is not an identity,
  * learnable initial states ``hs`` are non-zero,
  * biases of the GRU ("NBRC") cells are non-zero,
  * the blank logit bias is raised so that roughly 70-80 % of joint evaluations
    emit blank (with pure default init every frame emits ``max_iters`` symbols).
"""
from dataclasses import dataclass, asdict
import zlib

import numpy as np


@dataclass(frozen=True)
class ModelConfig:
    """Shape parameters of the path (names follow config/testing.yaml:202-229)."""

    n_mels: int = 80
    n_stack: int = 10
    downsample: int = 8
    enc_layers: int = 4
    pred_layers: int = 2
    hidden_sz: int = 1024
    out_sz: int = 1024
    embed_sz: int = 512
    joint_sz: int = 1024
    vocab_sz: int = 2048
    blank_bias: float = 0.0
    joint_gain1: float = 1.0  # scales joint.joint.0.weight (synthetic-weight realism knob)
    joint_gain2: float = 1.0  # scales joint.joint.2.weight
    blank_gain: float = 1.0  # extra scale on the blank row of joint.joint.2.weight (makes emission input-driven)
    rnn_gain: float = 1.0  # scales encoder LSTM weight matrices (trained nets are not at init scale)
    sample_rate: int = 16000
    n_fft: int = 1024
    win_length: int = 400
    hop_length: int = 160

    @property
    def feature_sz(self) -> int:
        return self.n_mels * self.n_stack

    def to_dict(self):
        return asdict(self)


# Realism knobs shared by every synthetic model (calibrated so that greedy decode
# mixes blank / non-blank decisions and hits every ``max_iters`` branch).
_GAINS = dict(joint_gain1=3.0, joint_gain2=6.0, blank_gain=4.0, rnn_gain=3.0)

# The shapes BASELINE.json / SURVEY.md section 8 name.
CONFIGS = {
    # tiny: fast CPU unit tests, every intermediate stored in the fixtures
    "tiny": ModelConfig(n_mels=16, enc_layers=2, pred_layers=2, hidden_sz=64, out_sz=64,
                        embed_sz=32, joint_sz=64, vocab_sz=64, blank_bias=3.0, **_GAINS),
    # reference's shipped shape (config/testing.yaml:133-135,202-229)
    "ref": ModelConfig(n_mels=128, enc_layers=6, blank_bias=6.0, **_GAINS),
    # BASELINE.json configs[1], configs[2]
    "cfg2": ModelConfig(n_mels=80, enc_layers=4, blank_bias=4.0, **_GAINS),  # ~75 % blank evaluations on the bench audio
    # BASELINE.json configs[3]
    "cfg4": ModelConfig(n_mels=80, enc_layers=6, hidden_sz=1536, out_sz=1536, blank_bias=1.5, **_GAINS),
}


# Audio seed of the headline workload (bench.py, BASELINE.json configs[1]): chosen among the first seeds so that the
# fp32 CPU path's own greedy decisions all have a top-1/top-2 log-probability margin >= 1e-3 (seed 0 contains a
# 1.5e-5 near-tie that two fp32 implementations may legitimately resolve differently); 77 % blank evaluations.
BENCH_AUDIO_SEED = 4


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed & 0xFFFFFFFF, zlib.crc32(name.encode())]))


def _uniform(seed, name, shape, bound):
    r = _rng(seed, name)
    return ((r.random(size=shape, dtype=np.float32) * 2.0 - 1.0) * np.float32(bound)).astype(np.float32)


def _normal(seed, name, shape, std, mean=0.0):
    r = _rng(seed, name)
    return (r.standard_normal(size=shape, dtype=np.float32) * np.float32(std) + np.float32(mean)).astype(np.float32)


def make_state_dict(cfg: ModelConfig, seed: int = 1234) -> dict:
    """name -> np.ndarray, with exactly the reference ``state_dict`` keys."""
    H, X, E, J, V, O = cfg.hidden_sz, cfg.feature_sz, cfg.embed_sz, cfg.joint_sz, cfg.vocab_sz, cfg.out_sz
    assert O == H, "reference configs all use out_sz == hidden_sz (models.py:97-100,173-176 -> identity)"
    sd = {}

    def bn(prefix):
        sd[prefix + ".weight"] = _uniform(seed, prefix + ".weight", (H,), 0.25) + np.float32(1.0)
        sd[prefix + ".bias"] = _normal(seed, prefix + ".bias", (H,), 0.1)
        sd[prefix + ".running_mean"] = _normal(seed, prefix + ".running_mean", (H,), 0.1)
        sd[prefix + ".running_var"] = _uniform(seed, prefix + ".running_var", (H,), 0.5) + np.float32(1.0)
        sd[prefix + ".num_batches_tracked"] = np.array(1000, dtype=np.int64)

    # --- encoder (models.py:68-100, custom_rnn.py:85-132, 259-269)
    sd["encoder.input_norm.weight"] = _uniform(seed, "encoder.input_norm.weight", (X,), 0.2) + np.float32(1.0)
    sd["encoder.input_norm.bias"] = _normal(seed, "encoder.input_norm.bias", (X,), 0.1)
    k = 1.0 / np.sqrt(H)
    for i in range(cfg.enc_layers):
        inp = X if i == 0 else H
        p = f"encoder.rnn_stack.rnns.{i}."
        sd[p + "weight_ih_l0"] = _uniform(seed, p + "weight_ih_l0", (4 * H, inp), k * cfg.rnn_gain)
        sd[p + "weight_hh_l0"] = _uniform(seed, p + "weight_hh_l0", (4 * H, H), k * cfg.rnn_gain)
        sd[p + "bias_ih_l0"] = _uniform(seed, p + "bias_ih_l0", (4 * H,), k)
        sd[p + "bias_hh_l0"] = _uniform(seed, p + "bias_hh_l0", (4 * H,), k)
        sd[f"encoder.rnn_stack.hs.{i}"] = _normal(seed, f"encoder.rnn_stack.hs.{i}", (2, 1, 1, H), 0.1)
        bn(f"encoder.rnn_stack.bns.{i}")

    # --- predictor (models.py:143-176, haste/nbrc.py:134-138,201-212)
    sd["predictor.embed.weight"] = _normal(seed, "predictor.embed.weight", (V, E), 1.0)
    sd["predictor.embed.weight"][0] = 0.0  # padding_idx=blank row (models.py:159)
    sd["predictor.ffn.weight"] = _uniform(seed, "predictor.ffn.weight", (H, E), 1.0 / np.sqrt(E))
    sd["predictor.ffn.bias"] = _uniform(seed, "predictor.ffn.bias", (H,), 1.0 / np.sqrt(E))
    for i in range(cfg.pred_layers):
        p = f"predictor.rnn_stack.rnns.{i}."
        sd[p + "kernel"] = _uniform(seed, p + "kernel", (H, 3 * H), np.sqrt(6.0 / (2 * H)))
        sd[p + "recurrent_kernel"] = _normal(seed, p + "recurrent_kernel", (H, 3 * H), 1.0 / np.sqrt(H))
        sd[p + "bias"] = _normal(seed, p + "bias", (3 * H,), 0.05)
        sd[p + "recurrent_bias"] = _normal(seed, p + "recurrent_bias", (3 * H,), 0.05)
        sd[f"predictor.rnn_stack.hs.{i}"] = _normal(seed, f"predictor.rnn_stack.hs.{i}", (1, 1, 1, H), 0.1)
        bn(f"predictor.rnn_stack.bns.{i}")

    # --- joint (models.py:116-127), concat method
    kj = 1.0 / np.sqrt(2 * O)
    sd["joint.joint.0.weight"] = _uniform(seed, "joint.joint.0.weight", (J, 2 * O), kj * cfg.joint_gain1)
    sd["joint.joint.0.bias"] = _uniform(seed, "joint.joint.0.bias", (J,), kj)
    k2 = 1.0 / np.sqrt(J)
    sd["joint.joint.2.weight"] = _uniform(seed, "joint.joint.2.weight", (V, J), k2 * cfg.joint_gain2)
    sd["joint.joint.2.weight"][0] *= np.float32(cfg.blank_gain)
    b2 = _uniform(seed, "joint.joint.2.bias", (V,), k2)
    b2[0] += np.float32(cfg.blank_bias)
    sd["joint.joint.2.bias"] = b2
    return sd


@dataclass(frozen=True)
class LMConfig:
    """Shape of the fused language model (reference libreasr/lib/lm.py:20-29, config/testing.yaml:293-313)."""

    vocab_sz: int = 2048
    embed_sz: int = 768
    hidden_sz: int = 768
    num_layers: int = 4
    alpha: float = 0.1  # lm.py:13 (the decode loops call fuse() with its defaults, models.py:431,558)
    theta: float = 1.0  # lm.py:14
    rnn_gain: float = 3.0

    @property
    def tied(self) -> bool:  # lm.py:27-29
        return self.embed_sz == self.hidden_sz


LM_CONFIGS = {
    "tiny": LMConfig(vocab_sz=64, embed_sz=64, hidden_sz=64, num_layers=2),  # tied embedding / output weights
    "tiny_untied": LMConfig(vocab_sz=64, embed_sz=32, hidden_sz=64, num_layers=3),
    "en": LMConfig(),  # the shipped inference override (testing.yaml:306-313): 4 x 768, tied
    "default": LMConfig(embed_sz=1024, hidden_sz=1024, num_layers=6),  # testing.yaml:293-299
}


def make_lm_state_dict(lm: LMConfig, seed: int = 4321) -> dict:
    """name -> np.ndarray with the reference ``LM.state_dict()`` keys (lm.py:20-29)."""
    V, E, H = lm.vocab_sz, lm.embed_sz, lm.hidden_sz
    sd = {}
    sd["embed.weight"] = _normal(seed, "lm.embed.weight", (V, E), 1.0)
    sd["embed.weight"][0] = 0.0  # padding_idx=0
    k = 1.0 / np.sqrt(H)
    for i in range(lm.num_layers):
        inp = E if i == 0 else H
        sd[f"rnn.weight_ih_l{i}"] = _uniform(seed, f"lm.rnn.weight_ih_l{i}", (4 * H, inp), k * lm.rnn_gain)
        sd[f"rnn.weight_hh_l{i}"] = _uniform(seed, f"lm.rnn.weight_hh_l{i}", (4 * H, H), k * lm.rnn_gain)
        sd[f"rnn.bias_ih_l{i}"] = _uniform(seed, f"lm.rnn.bias_ih_l{i}", (4 * H,), k)
        sd[f"rnn.bias_hh_l{i}"] = _uniform(seed, f"lm.rnn.bias_hh_l{i}", (4 * H,), k)
    sd["linear.weight"] = sd["embed.weight"] if lm.tied else _uniform(seed, "lm.linear.weight", (V, H), k * 4.0)
    sd["linear.bias"] = _uniform(seed, "lm.linear.bias", (V,), k)
    return sd


def make_audio(batch: int, n_samples: int, seed: int = 0) -> np.ndarray:
    """[batch, n_samples] float32 synthetic 16 kHz audio.  Speech-like non-stationarity:
    a N(0, 0.02^2) noise floor plus back-to-back 40-300 ms segments, each a harmonic
    stack (random f0, spectral tilt) and band-passed noise burst with random gains and
    an attack/decay envelope; ~15 % of segments are silence."""
    out = np.empty((batch, n_samples), dtype=np.float32)
    for b in range(batch):
        r = _rng(seed, f"audio.{b}")
        x = r.standard_normal(size=n_samples, dtype=np.float32).astype(np.float64) * 0.02
        pos = 0
        while pos < n_samples:
            seg = int(16000 * (0.04 + 0.26 * float(r.random())))
            end = min(n_samples, pos + seg)
            m = end - pos
            u = [float(v) for v in r.random(size=8)]
            if u[0] > 0.15 and m > 8:
                t = np.arange(m, dtype=np.float64) / 16000.0
                env = np.minimum(1.0, np.minimum(t / 0.01, (t[-1] - t) / 0.02 + 1e-3))
                f0 = 80.0 + 320.0 * u[1]
                tilt = 0.3 + 1.7 * u[2]
                nh = int(min(20, 7600.0 // f0))
                sig = np.zeros(m)
                for h in range(1, nh + 1):
                    sig += (h ** -tilt) * np.sin(2 * np.pi * f0 * h * t + 2 * np.pi * u[3] * h)
                sig *= (0.02 + 0.25 * u[4])
                nz = r.standard_normal(size=m, dtype=np.float32).astype(np.float64)
                fc = 300.0 + 6000.0 * u[5]  # crude band emphasis: modulate noise to fc
                nz = nz * np.cos(2 * np.pi * fc * t) * (0.2 * u[6])
                x[pos:end] += env * (sig + nz)
            pos = end
        out[b] = x.astype(np.float32)
    return out
