"""Inference-side transforms with the reference's class names and ``encodes`` semantics
(libreasr/lib/transforms.py), computing on the CUDA engine.

The reference assembles these by name from YAML (config.py:45-69,
config/testing.yaml:339-374).  ``TransformTime`` runs the mel kernel; the remaining
transforms are pure re-indexing and operate on the CUDA tensor it returns.  The fused
hot path (``Engine.features`` / ``Engine.features_stream``) produces the same rows in
one kernel; ``FusedFeatures`` / ``FusedStreamFeatures`` expose it in pipeline form.
"""
import torch

__all__ = ["Resample", "ChannelCut", "TransformTime", "StreamPostprocess", "StackDownsample", "FixDimensions",
           "Buffer", "FusedFeatures", "FusedStreamFeatures", "tensorize"]


def tensorize(x):
    """bytes (little-endian f32 PCM) -> [1, n] tensor (utils.py:149-153)."""
    import numpy as np

    return torch.from_numpy(np.frombuffer(x, dtype=np.float32).copy())[None]


class Resample:
    """transforms.py:135-144: ``torchaudio.transforms.Resample(orig_freq=i.sr, new_freq=target_sr)`` on the GPU
    (``rnnt_b200_resample``: the same windowed-sinc filter bank).  ``engine`` may be omitted for the identity case."""

    def __init__(self, target_sr=16000, engine=None, **kwargs):
        self.sr = target_sr
        self.engine = engine

    def encodes(self, i, sr=None):
        sr = getattr(i, "sr", sr) if sr is None else sr
        if sr is None or int(sr) == int(self.sr):
            return i
        if self.engine is None:
            raise RuntimeError("Resample needs the CUDA engine for sr != target_sr (there is no CPU path): Resample(sr, engine=...)")
        if int(self.engine.cfg.sample_rate) != int(self.sr):
            raise ValueError("target_sr differs from the engine's sample rate")
        x = i if i.dim() == 2 else i[None]
        return self.engine.resample(x.to(self.engine.device, torch.float32), int(sr))

    __call__ = encodes


class ChannelCut:
    """transforms.py:122-132: keep the first ``channels`` channels."""

    def __init__(self, channels=1, **kwargs):
        self.n_chans = channels

    def encodes(self, i):
        return i if i.size(0) == self.n_chans else i[: self.n_chans]

    __call__ = encodes


class TransformTime:
    """transforms.py:269-323 with deltas=0: [C, n] audio -> [C, F, n_mels] log-mel."""

    def __init__(self, engine, **kwargs):
        self.engine = engine

    def encodes(self, sig):
        return self.engine.logmel(sig.to(self.engine.device, torch.float32))

    __call__ = encodes


class StreamPostprocess:
    """transforms.py:326-342: keep n_stack frames after the first third."""

    def __init__(self, n_stack, **kwargs):
        self.n_stack = n_stack

    def encodes(self, spectro):
        a = spectro.shape[1] // 3 + 1
        return spectro[:, a:, :][:, : self.n_stack, :]

    __call__ = encodes


class StackDownsample:
    """transforms.py:429-441."""

    def __init__(self, n_stack=6, downsample=3, **kwargs):
        self.n_stack, self.downsample = n_stack, downsample

    def encodes(self, t):
        uf = t.unfold(-2, self.n_stack, self.downsample).contiguous()
        return uf.view(uf.size(0), uf.size(1), -1).contiguous()

    __call__ = encodes


class FixDimensions:
    """transforms.py:444-452."""

    def __init__(self, **kwargs):
        pass

    def encodes(self, t):
        return t.unsqueeze(-1)

    __call__ = encodes


class Buffer:
    """transforms.py:455-471.  One instance per stream (the reference shares one across
    its gRPC worker threads, which is only correct for a single stream)."""

    def __init__(self, n_buffer, **kwargs):
        self.n_buffer = n_buffer
        self.saved = []

    def encodes(self, t):
        self.saved.append(t)
        if len(self.saved) == self.n_buffer:
            catted = torch.cat(self.saved, dim=1)
            self.saved.clear()
            return catted[0]
        return None

    __call__ = encodes


class FusedFeatures:
    """Resample(id) -> ChannelCut -> TransformTime -> StackDownsample -> FixDimensions
    (config/testing.yaml:339-354) in one kernel: [C, n] -> [C, T, X, 1]."""

    def __init__(self, engine):
        self.engine = engine

    def __call__(self, sig):
        return self.engine.features(sig.to(self.engine.device, torch.float32)[:1]).unsqueeze(-1)


class FusedStreamFeatures:
    """The ``stream`` pipeline (config/testing.yaml:356-374) incl. ``Buffer``:
    [1, 3*chunk] window -> [n_buffer, X, 1] every n_buffer-th call, else None."""

    def __init__(self, engine, n_buffer=2):
        self.engine, self.buffer = engine, Buffer(n_buffer)

    def __call__(self, window):
        row = self.engine.features_stream(window.to(self.engine.device, torch.float32)[:1])  # [1, X]
        return self.buffer(row[:, None, :, None])
