"""TEST / BENCH INFRASTRUCTURE ONLY -- run the UNMODIFIED reference on CPU, utterance by utterance.

Used by ``bench.py``'s CPU arm (``cpu_baseline.kind == "reference"``) when the reference package is importable
(``/root/reference`` in the authoring container, ``baseline/_ref`` -- the offline install of the same files -- on
the GPU box).  What runs is the reference's own ``Transducer.decode_greedy`` (libreasr/lib/models.py:369-455)
on features produced by the very ``torchaudio.transforms.MelSpectrogram`` call the reference makes
(transforms.py:290-296) followed by the tensor ops of TransformTime / StackDownsample / FixDimensions
(transforms.py:311-323, 436-441, 450-452) -- the reference's transforms.py itself needs fastai2 and cannot be
imported (see make_golden.py).  bs = 1, as the reference serves (config/testing.yaml:380).
"""
import numpy as np
import torch

from . import ref_shim, weights


def available() -> bool:
    return ref_shim.reference_available()


class ReferenceRunner:
    def __init__(self, cfg, weight_seed=1234):
        from . import make_golden as G

        self.cfg = cfg
        self.G = G
        M = ref_shim.import_reference_models()
        sd = weights.make_state_dict(cfg, weight_seed)
        self.ref = M.Transducer.from_config(ref_shim.reference_conf(cfg), ref_shim.FakeLang())
        self.ref.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=True)
        self.ref.eval()
        self.root = ref_shim.REFERENCE_ROOT

    def transcribe(self, audio_1d: np.ndarray, max_iters=3):
        """One utterance [n] -> token ids, through the reference's own decode_greedy."""
        with torch.no_grad():
            feats = self.G.ref_features_offline(torch.from_numpy(audio_1d[None]), self.cfg)[0]   # [T, X, 1]
            toks = self.ref.decode_greedy(feats, max_iters=max_iters)[0]
        return [int(t) for t in toks]

    def transcribe_batch(self, audio: np.ndarray, max_iters=3):
        return [self.transcribe(audio[b], max_iters) for b in range(audio.shape[0])]
