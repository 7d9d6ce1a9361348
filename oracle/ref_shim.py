"""TEST INFRASTRUCTURE ONLY -- import the unmodified reference in this container.

The reference's ``libreasr/lib/models.py`` imports fastai2, IPython and
matplotlib at module scope (models.py:16-21, utils.py:9, custom_rnn.py:10);
none is installed here and there is no network.  This module registers the
few names those imports need in ``sys.modules`` and then imports the reference
source *as it lies* under ``/root/reference`` (nothing is copied).

The only behavioural piece is ``fastai2.torch_core.Module``: fastai's
``PrePostInitMeta`` calls ``nn.Module.__init__`` before the user ``__init__``
(the reference classes never call ``super().__init__()``, e.g. models.py:68-100).
"""
import os
import sys
import types

import torch.nn as nn

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_reference_root():
    """The authoring container has the tree at /root/reference; the GPU box only has the offline install of the same
    unmodified package under baseline/_ref (git-ignored, travels with the snapshot; recipe in DESIGN.md section 2)."""
    for cand in (os.environ.get("LIBREASR_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "libreasr", "lib", "models.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_reference_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "libreasr", "lib", "models.py"))


class _PrePostInitMeta(type):
    """Metaclass with fastai ``PrePostInitMeta`` semantics (pre-init hook only)."""

    def __call__(cls, *args, **kwargs):
        obj = cls.__new__(cls)
        nn.Module.__init__(obj)
        obj.__init__(*args, **kwargs)
        return obj


class _FastaiModule(nn.Module, metaclass=_PrePostInitMeta):
    def __init__(self):  # pragma: no cover - user classes override
        pass


def _mod(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def install_stubs():
    class _Cancel(Exception):
        pass

    _mod("fastai2")
    _mod("fastai2.vision")
    _mod("fastai2.vision.models")
    _mod("fastai2.vision.models.xresnet", xresnet18=lambda *a, **k: None)
    _mod("fastai2.layers", Debugger=object, ResBlock=object)
    _mod("fastai2.torch_core", Module=_FastaiModule)
    _mod("fastai2.learner", CancelBatchException=_Cancel)
    _mod("IPython")
    _mod("IPython.core")
    _mod("IPython.core.debugger", set_trace=lambda *a, **k: None)
    _mod("matplotlib")
    _mod("matplotlib.pyplot")


def import_reference_models():
    """Returns the reference's ``libreasr.lib.models`` module (unmodified source)."""
    if not reference_available():
        raise RuntimeError(
            f"reference tree not found under {REFERENCE_ROOT}; the shim only works "
            "in the authoring container"
        )
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    return importlib.import_module("libreasr.lib.models")


class FakeLang:
    """Stands in for ``TokenizedLanguage`` (language.py:115-151): token ids are the
    parity contract, so ``denumericalize`` returns the ids themselves."""

    def denumericalize(self, ids):
        return list(ids)


def reference_conf(cfg):
    """``conf`` dict with the keys ``Transducer.from_config`` reads (models.py:236-259),
    filled from an oracle ``ModelConfig``."""
    return {
        "model": {
            "feature_sz": cfg.feature_sz,
            "embed_sz": cfg.embed_sz,
            "vocab_sz": cfg.vocab_sz,
            "hidden_sz": cfg.hidden_sz,
            "out_sz": cfg.out_sz,
            "joint_sz": cfg.joint_sz,
            "encoder": {
                "rnn_type": "LSTM",
                "num_layers": cfg.enc_layers,
                "dropout": 0.05,
                "layer_norm": False,
                "use_tmp_state_pcent": 0.99,
            },
            "predictor": {
                "rnn_type": "NBRC",
                "num_layers": cfg.pred_layers,
                "dropout": 0.05,
                "layer_norm": False,
                "use_tmp_state_pcent": 0.99,
            },
            "joint": {"method": "concat", "dropout": 0.0},
            "use_tmp_bos": False,
            "use_tmp_bos_pcent": 0.2,
        },
        "bs": 1,
        "mp": False,
        "cuda": {"device": "cpu"},
    }
