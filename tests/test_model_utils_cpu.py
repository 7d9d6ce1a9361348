"""CPU: checkpoint bundle ingestion (reference libreasr/lib/model_utils.py:20-95)."""
import os
import tarfile

import numpy as np
import pytest
import torch

from libreasr_b200 import synth
from libreasr_b200.lib import model_utils as MU
from libreasr_b200.lib.models import Transducer


class Lang:
    def denumericalize(self, ids):
        return list(ids)


def _model(cfg):
    return Transducer(cfg.feature_sz, cfg.embed_sz, cfg.vocab_sz, cfg.hidden_sz, cfg.out_sz, cfg.joint_sz, Lang(),
                      encoder_kwargs={"num_layers": cfg.enc_layers}, predictor_kwargs={"num_layers": cfg.pred_layers})


@pytest.mark.parametrize("fastai_dict", [True, False])
def test_bundle_roundtrip(tmp_path, fastai_dict):
    cfg = synth.CONFIGS["tiny"]
    sd = {k: torch.as_tensor(v) for k, v in synth.make_state_dict(cfg, 99).items()}
    src = _model(cfg)
    src.load_state_dict(sd, strict=True)
    dest = tmp_path / "tmp"
    (dest / "en").mkdir(parents=True)
    full = src.state_dict()
    torch.save({"model": full, "opt": {"state": [], "hypers": []}} if fastai_dict else full, dest / "en" / "model.pth")   # what learn.save writes
    (dest / "en" / "tokenizer.yttm-model").write_bytes(b"\x00yttm")
    arc = tmp_path / "libreasr-model-en.tar.gz"
    MU.save_asr_model("en", path_archive=arc, path_dest=dest)
    with tarfile.open(arc) as t:
        assert sorted(t.getnames()) == ["en/model.pth", "en/tokenizer.yttm-model"]     # the reference's member names
    out = tmp_path / "extracted"
    MU.extract_tars([str(arc)], path_dest=out)
    assert (out / "en" / "tokenizer.yttm-model").exists()
    m = _model(cfg)
    m.lm = object()
    lang = Lang()
    m = MU.load_asr_model(m, "en", lang, device="cpu", path_dest=out)
    assert m.lang is lang and m.lm is None
    got = m.state_dict()
    assert set(got) == set(full)
    for k in full:
        assert torch.equal(got[k], full[k]), k


def test_extract_refuses_members_outside_the_destination(tmp_path):
    arc = tmp_path / "evil.tar.gz"
    p = tmp_path / "x"
    p.write_bytes(b"1")
    with tarfile.open(arc, "w:gz") as t:
        t.add(str(p), arcname="../escape")
    with pytest.raises(ValueError):
        MU.extract_tars([str(arc)], path_dest=tmp_path / "d")


def test_wrong_checkpoint_fails_loudly(tmp_path):
    cfg = synth.CONFIGS["tiny"]
    dest = tmp_path / "tmp" / "en"
    dest.mkdir(parents=True)
    torch.save({"encoder.input_norm.weight": torch.zeros(3)}, dest / "model.pth")
    with pytest.raises(Exception):
        MU.load_asr_model(_model(cfg), "en", Lang(), device="cpu", path_dest=tmp_path / "tmp")
