"""CPU: the C-ABI library builds, loads and exports every symbol include/rnnt_b200.h
declares; host-side validation and the no-CPU-fallback rule.  No compute calls."""
import ctypes as C
import os
import re

import pytest
import torch

from libreasr_b200 import _capi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "rnnt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rnnt_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    lib = C.CDLL(path)
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rnnt_b200.h but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(_capi.SIGNATURES) == names


def test_abi_version_and_default_config():
    lib = _capi.load_library()
    assert lib.rnnt_b200_abi_version() == 2
    c = _capi.Config()
    assert lib.rnnt_b200_default_config(C.byref(c)) == 0
    # reference's shipped shape (config/testing.yaml:133-135,202-229)
    assert (c.n_mels, c.n_stack, c.downsample, c.enc_layers, c.pred_layers) == (128, 10, 8, 6, 2)
    assert (c.hidden_sz, c.embed_sz, c.joint_sz, c.vocab_sz, c.blank, c.bos) == (1024, 512, 1024, 2048, 0, 2)
    assert (c.sample_rate, c.n_fft, c.win_length, c.hop_length) == (16000, 1024, 400, 160)
    # no fused LM by default (m.lm is None, models.py:234); fuser constants of lm.py:13-14
    assert c.lm_layers == 0 and abs(c.lm_alpha - 0.1) < 1e-7 and c.lm_theta == 1.0


@pytest.mark.parametrize("field,value", [("n_fft", 512), ("hidden_sz", 100), ("vocab_sz", 33), ("pred_layers", 9),
                                         ("joint_sz", 0), ("gemm_mode", 7), ("bos", 5000), ("lm_layers", 9)])
def test_create_rejects_bad_config(field, value):
    lib = _capi.load_library()
    c = _capi.Config()
    lib.rnnt_b200_default_config(C.byref(c))
    setattr(c, field, value)
    h = C.c_void_p(0)
    assert lib.rnnt_b200_create(C.byref(c), C.byref(h)) == -1
    assert not h.value
    assert len(lib.rnnt_b200_last_error(None)) > 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour WITHOUT a GPU")
def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, not compute on the CPU."""
    from libreasr_b200.engine import Engine, EngineConfig
    from libreasr_b200.lib.models import Transducer

    with pytest.raises(RuntimeError):
        Engine(EngineConfig())
    m = Transducer(160, 32, 64, 64, 64, 64, lang=None, encoder_kwargs={"num_layers": 2}, predictor_kwargs={"num_layers": 2})
    with pytest.raises(RuntimeError):
        m.engine()
    with pytest.raises(RuntimeError):
        m.convert_to_cpu()
    with pytest.raises(RuntimeError):
        m.encoder(torch.zeros(1, 4, 160, 1))


def test_product_code_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "libreasr_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", s, flags=re.M) or "rnnt_oracle" in s:
                    bad.append(f)
    assert not bad, bad


def test_state_dict_keys_match_reference_contract():
    """SURVEY.md section 8b checkpoint contract == keys of oracle.weights == keys of the mirror."""
    from oracle import weights
    from libreasr_b200.lib.models import Transducer

    cfg = weights.CONFIGS["tiny"]
    sd = weights.make_state_dict(cfg)
    m = Transducer(cfg.feature_sz, cfg.embed_sz, cfg.vocab_sz, cfg.hidden_sz, cfg.out_sz, cfg.joint_sz, lang=None,
                   encoder_kwargs={"num_layers": cfg.enc_layers}, predictor_kwargs={"num_layers": cfg.pred_layers})
    assert sorted(m.state_dict().keys()) == sorted(sd.keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    m.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()}, strict=True)


def test_engine_config_from_reference_style_conf():
    """``EngineConfig.from_conf`` reads the keys of config/testing.yaml (values below are that file's: melkwargs :131-135,
    StackDownsample :352-355, model :202-229, lm :293-299 and the `en` inference override :306-313)."""
    from libreasr_b200.engine import EngineConfig

    conf = {
        "sr": 16000, "win_length": 0.025, "hop_length": 0.01, "melkwargs": {"n_fft": 1024, "n_mels": 128},
        "transforms": {"x": [{"name": "TransformTime"}, {"name": "StackDownsample", "args": {"n_stack": 10, "downsample": 8}}]},
        "model": {"feature_sz": 1280, "embed_sz": 512, "vocab_sz": 2048, "hidden_sz": 1024, "out_sz": 1024, "joint_sz": 1024,
                  "encoder": {"num_layers": 6}, "predictor": {"num_layers": 2}, "joint": {"method": "concat"}},
        "lm": {"enable": True, "vocab_sz": 2048, "embed_sz": 1024, "hidden_sz": 1024, "num_layers": 6, "p": 0.2},
    }
    ec = EngineConfig.from_conf(conf)
    assert (ec.n_mels, ec.n_stack, ec.downsample, ec.win_length, ec.hop_length) == (128, 10, 8, 400, 160)
    assert (ec.enc_layers, ec.pred_layers, ec.hidden_sz, ec.embed_sz, ec.joint_sz, ec.vocab_sz) == (6, 2, 1024, 512, 1024, 2048)
    assert (ec.lm_layers, ec.lm_hidden_sz, ec.lm_embed_sz) == (6, 1024, 1024)
    conf["lm"].update({"embed_sz": 768, "hidden_sz": 768, "num_layers": 4, "p": 0.3})      # overrides.en.lm
    ec = EngineConfig.from_conf(conf)
    assert (ec.lm_layers, ec.lm_hidden_sz, ec.lm_embed_sz) == (4, 768, 768)
    conf["lm"]["enable"] = False
    assert EngineConfig.from_conf(conf).lm_layers == 0
    assert ec.feature_sz == conf["model"]["feature_sz"]
