"""Checkpoint bundle ingestion: the reference's ``libreasr/lib/model_utils.py`` for the inference path.

The reference ships a model as ``libreasr-model[-<lang>].tar.gz`` holding ``<lang>/model.pth`` (what fastai's
``learn.save`` wrote: either a bare ``state_dict`` or ``{"model": state_dict, "opt": ...}``, fastai2 ``load_model``) and
``<lang>/tokenizer.yttm-model`` (model_utils.py:20-95, docs/docs.md:139-140).  ``load_asr_model`` here does what the
reference's does -- extract, read ``model.pth``, ``load_state_dict`` with the reference's key set (strict) -- on the
B200 ``Transducer`` (libreasr_b200.lib.models), whose engine repacks the weights at its next use.

Not taken over: ``maybe_quantize`` (model_utils.py:90-93, a CPU int8 path; the B200 path computes in fp32-grade 3xFP16)
and the YouTokenToMe tokenizer (``tokenizer.yttm-model`` is extracted next to the weights but youtokentome is not a
dependency of this package: token ids are the contract, ``lang.denumericalize`` stays the caller's).
"""
import glob
import io
import os
import tarfile
from pathlib import Path

import torch

_PATH_ARCHIVE = Path("libreasr-model.tar.gz")
_PATH_TOKENIZER = Path("tokenizer.yttm-model")
_PATH_MODEL = Path("model.pth")
_PATH_DEST = Path("./tmp")


def save_asr_model(lang, path_tokenizer=_PATH_TOKENIZER, path_model=_PATH_MODEL, path_archive=_PATH_ARCHIVE, path_dest=_PATH_DEST):
    """Bundles ``<dest>/<lang>/tokenizer.yttm-model`` and ``<dest>/<lang>/model.pth`` into ``path_archive`` with the
    member names the reference uses (model_utils.py:32-49)."""
    base_real, base_arc = Path(path_dest) / Path(lang), Path(lang)
    with tarfile.open(path_archive, mode="w:gz") as tar:
        for name in (path_tokenizer, path_model):
            tar.add(str(base_real / name), arcname=str(base_arc / name))


def extract_tars(paths_archive=None, path_dest=_PATH_DEST):
    """model_utils.py:52-60: every ``./libreasr-model-*.tar.gz`` (or the given archives) into ``path_dest``.  Members
    that would land outside ``path_dest`` are refused."""
    if paths_archive is None:
        paths_archive = glob.glob("./libreasr-model-*.tar.gz")
    dest = os.path.realpath(str(path_dest))
    for arc in paths_archive:
        with tarfile.open(arc) as tar:
            for m in tar.getmembers():
                target = os.path.realpath(os.path.join(dest, m.name))
                if not (target == dest or target.startswith(dest + os.sep)) or m.issym() or m.islnk():
                    raise ValueError(f"refusing archive member {m.name!r}")
            tar.extractall(path=dest)


def read_model_state(path_model, device="cpu"):
    """fastai2 ``load_model`` semantics: ``{"model": sd, "opt": ...}`` or a bare state_dict."""
    state = torch.load(str(path_model), map_location=device, weights_only=False)
    if isinstance(state, dict) and set(state.keys()) == {"model", "opt"}:
        state = state["model"]
    return state


def load_asr_model(model, lang_name, lang, device="cuda:0", lm=None, path_tokenizer=_PATH_TOKENIZER, path_archive=_PATH_ARCHIVE,
                   path_dest=_PATH_DEST):
    """model_utils.py:63-95: weights of ``<dest>/<lang_name>/model.pth`` into ``model`` (reference key set, strict), moved to
    ``device``.  Like the reference it clears ``model.lang`` / ``model.lm`` first; the caller re-attaches them."""
    model.lang = None
    model.lm = None
    try:
        sd = read_model_state(Path(path_dest) / Path(lang_name) / _PATH_MODEL, device="cpu")
        sd = {k: v for k, v in sd.items() if not k.startswith("lm.")}
        model.load_state_dict(sd, strict=True)
    except Exception as e:
        print("Unable to load_model(...)")
        raise e
    model.lang = lang
    return model.to(device)
