"""Drop-in module surface of the reference's ``libreasr/lib/models.py`` for the inference
path: ``Transducer`` / ``Encoder`` / ``Predictor`` / ``Joint`` with the same constructor
arguments, ``state_dict`` key set and method signatures, executing
on the sm_100a CUDA library through the C ABI (``libreasr_b200.engine``).

The ``nn.Module`` objects here are parameter CONTAINERS (so ``load_state_dict`` /
``state_dict`` / ``.to()`` keep working with reference checkpoints); they never run a
PyTorch forward.  Anything that needs a CPU or a non-CUDA device raises: there is no
fallback path.
"""
import numpy as np
import torch
from torch import nn

from ..engine import Engine, EngineConfig, tokens_to_lists

__all__ = ["Transducer", "Encoder", "Predictor", "Joint", "NBRC", "CustomRNNParams"]


class NBRC(nn.Module):
    """Parameter container for the Haste-style GRU ("NBRC") cell
    (reference libreasr/lib/layers/haste/nbrc.py:134-138): gate layout z,r,g."""

    def __init__(self, input_size, hidden_size, batch_first=True):
        super().__init__()
        self.input_size, self.hidden_size, self.batch_first = input_size, hidden_size, batch_first
        self.kernel = nn.Parameter(torch.empty(input_size, hidden_size * 3))
        self.recurrent_kernel = nn.Parameter(torch.empty(hidden_size, hidden_size * 3))
        self.bias = nn.Parameter(torch.zeros(hidden_size * 3))
        self.recurrent_bias = nn.Parameter(torch.zeros(hidden_size * 3))
        for i in range(3):  # nbrc.py:201-212
            nn.init.xavier_uniform_(self.kernel.data[:, i * hidden_size:(i + 1) * hidden_size])
            nn.init.orthogonal_(self.recurrent_kernel.data[:, i * hidden_size:(i + 1) * hidden_size])


class CustomRNNParams(nn.Module):
    """Parameter container mirroring ``CustomCPURNN`` (custom_rnn.py:85-132, 259-269):
    ``hs`` learnable initial states, ``bns`` BatchNorm1d per layer, ``rnns`` single-layer
    ``nn.LSTM`` (weight_ih_l0 ...) or ``NBRC``."""

    def __init__(self, input_size, hidden_size, num_layers=1, rnn_type="LSTM", **_):
        super().__init__()
        assert rnn_type in ("LSTM", "NBRC"), "the path implements LSTM encoders and NBRC (GRU) predictors"
        self.rnn_type, self.hidden_size, self.num_layers = rnn_type, hidden_size, num_layers
        ins = [input_size] + [hidden_size] * (num_layers - 1)
        self.hs = nn.ParameterList(
            [nn.Parameter(torch.zeros(2 if rnn_type == "LSTM" else 1, 1, 1, hidden_size)) for _ in ins])
        self.bns = nn.ModuleList([nn.BatchNorm1d(hidden_size) for _ in ins])
        if rnn_type == "LSTM":
            self.rnns = nn.ModuleList([nn.LSTM(i, hidden_size, batch_first=True) for i in ins])
        else:
            self.rnns = nn.ModuleList([NBRC(i, hidden_size, batch_first=True) for i in ins])


def _owner(mod):
    o = getattr(mod, "_owner", None)
    if o is None:
        raise RuntimeError(
            f"{type(mod).__name__} must belong to a Transducer (its CUDA engine holds the repacked weights)")
    return o()


class Encoder(nn.Module):
    """``Encoder`` (models.py:68-113)."""

    def __init__(self, feature_sz, hidden_sz, out_sz, dropout=0.01, num_layers=2, trace=True, device="cuda:0",
                 layer_norm=False, rnn_type="LSTM", use_tmp_state_pcent=0.9, **kwargs):
        super().__init__()
        if hidden_sz != out_sz:
            raise NotImplementedError("hidden_sz != out_sz (extra Linear, models.py:97-98) is outside the built path")
        if layer_norm:
            raise NotImplementedError("layer_norm=True selects the Haste LayerNormLSTM (custom_rnn.py:30-36), outside the built path")
        self.num_layers = num_layers
        self.input_norm = nn.LayerNorm(feature_sz)
        self.rnn_stack = CustomRNNParams(feature_sz, hidden_sz, num_layers, rnn_type=rnn_type)

    def param_groups(self):
        return [p for p in self.parameters() if p.requires_grad]

    def forward(self, x, state=None, lengths=None, return_state=False):
        """x [N,T,X(,1)]; state list of (h,c) each [1,N,H] (models.py:105-113)."""
        eng = _owner(self).engine()
        x = x.reshape((x.size(0), x.size(1), -1)).to(eng.device, torch.float32)
        st = None
        if state is not None:
            st = (torch.cat([s[0] for s in state], 0).to(eng.device), torch.cat([s[1] for s in state], 0).to(eng.device))
        out, ns = eng.encode(x, lens_T=lengths, state=st, want_state=return_state)
        if return_state:
            return out, [(ns[0][i:i + 1], ns[1][i:i + 1]) for i in range(self.num_layers)]
        return out


class Predictor(nn.Module):
    """``Predictor`` (models.py:143-187)."""

    def __init__(self, vocab_sz, embed_sz, hidden_sz, out_sz, dropout=0.01, num_layers=2, blank=0, layer_norm=False,
                 rnn_type="NBRC", use_tmp_state_pcent=0.9):
        super().__init__()
        if hidden_sz != out_sz:
            raise NotImplementedError("hidden_sz != out_sz (extra Linear, models.py:173-174) is outside the built path")
        if layer_norm:
            raise NotImplementedError("layer_norm=True selects the Haste LayerNorm cell (custom_rnn.py:37-44), outside the built path")
        self.vocab_sz, self.num_layers = vocab_sz, num_layers
        self.embed = nn.Embedding(vocab_sz, embed_sz, padding_idx=blank)
        self.ffn = nn.Linear(embed_sz, hidden_sz) if embed_sz != hidden_sz else nn.Sequential()
        self.rnn_stack = CustomRNNParams(hidden_sz, hidden_sz, num_layers, rnn_type=rnn_type)

    def param_groups(self):
        return [p for p in self.parameters() if p.requires_grad]

    def forward(self, x, state=None, lengths=None):
        """x [N,1] tokens; state list of [1,N,H] -> (out [N,1,H], new state) (models.py:181-187)."""
        eng = _owner(self).engine()
        if x.dim() != 2 or x.size(1) != 1:
            raise ValueError("Predictor.forward takes one token per sequence: x of shape [N, 1]")
        st = None if state is None else torch.cat(list(state), 0).to(eng.device)
        out, ns = eng.predict(x[:, 0], st)
        return out[:, None, :], [ns[i:i + 1] for i in range(self.num_layers)]


class Joint(nn.Module):
    """``Joint`` (models.py:116-140)."""

    def __init__(self, out_sz, joint_sz, vocab_sz, joint_method):
        super().__init__()
        self.joint_method = joint_method
        if joint_method == "concat":
            input_sz = 2 * out_sz
        elif joint_method == "add":
            raise NotImplementedError('joint_method "add" is outside the built path (config uses "concat", testing.yaml:225)')
        else:
            raise Exception("No such joint_method")  # models.py:124
        self.joint = nn.Sequential(nn.Linear(input_sz, joint_sz), nn.Tanh(), nn.Linear(joint_sz, vocab_sz))

    def param_groups(self):
        return [p for p in self.parameters() if p.requires_grad]

    def forward(self, h_pred, h_enc):
        """Broadcasting ``cat`` + MLP (models.py:132-140): [...,H] x [...,H] -> logits [...,V]."""
        eng = _owner(self).engine()
        hp, he = torch.broadcast_tensors(h_pred.to(eng.device), h_enc.to(eng.device))
        lead = hp.shape[:-1]
        H = hp.shape[-1]
        out = eng.joint(hp.reshape(-1, H).float(), he.reshape(-1, H).float())
        return out.reshape(*lead, -1)


class Transducer(nn.Module):
    """``Transducer`` (models.py:190-577), inference methods."""

    def __init__(self, feature_sz, embed_sz, vocab_sz, hidden_sz, out_sz, joint_sz, lang, l_e=6, l_p=2, p_j=0.0,
                 blank=0, joint_method="concat", perf=False, act=None, use_tmp_bos=True, use_tmp_bos_pcent=0.99,
                 encoder_kwargs={}, predictor_kwargs={}, n_stack=10, downsample=8, gemm_mode=1, **kwargs):
        super().__init__()
        import weakref

        self.encoder = Encoder(feature_sz, hidden_sz=hidden_sz, out_sz=out_sz, **encoder_kwargs)
        self.predictor = Predictor(vocab_sz, embed_sz=embed_sz, hidden_sz=hidden_sz, out_sz=out_sz, **predictor_kwargs)
        self.joint = Joint(out_sz, joint_sz, vocab_sz, joint_method)
        for m in (self.encoder, self.predictor, self.joint):
            object.__setattr__(m, "_owner", weakref.ref(self))
        self.lang = lang
        self.blank = blank
        self.bos = 2  # models.py:226-227
        self.perf = perf
        self.mp = False
        self.vocab_sz = vocab_sz
        self.lm = None
        if feature_sz % n_stack:
            raise ValueError("feature_sz must be n_mels * n_stack")
        self._ecfg = EngineConfig(
            n_mels=feature_sz // n_stack, n_stack=n_stack, downsample=downsample,
            enc_layers=self.encoder.num_layers, pred_layers=self.predictor.num_layers, hidden_sz=hidden_sz,
            embed_sz=embed_sz, joint_sz=joint_sz, vocab_sz=vocab_sz, blank=blank, bos=self.bos, gemm_mode=gemm_mode)
        self._engine = None

    # ---- construction ---------------------------------------------------------------
    @staticmethod
    def from_config(conf, lang, lm=None):
        """models.py:236-259; ``conf["cuda"]["device"]`` must be a CUDA device."""
        ecfg = EngineConfig.from_conf(conf)
        m = Transducer(
            conf["model"]["feature_sz"], conf["model"]["embed_sz"], conf["model"]["vocab_sz"],
            conf["model"]["hidden_sz"], conf["model"]["out_sz"], conf["model"]["joint_sz"], lang,
            p_j=conf["model"]["joint"].get("dropout", 0.0), joint_method=conf["model"]["joint"]["method"],
            encoder_kwargs=conf["model"]["encoder"], predictor_kwargs=conf["model"]["predictor"],
            n_stack=ecfg.n_stack, downsample=ecfg.downsample, gemm_mode=conf.get("gemm_mode", 1),
        )
        # the front end (sr / window / hop / n_fft / n_mels) comes from the conf too, not from defaults
        import dataclasses

        if ecfg.feature_sz != conf["model"]["feature_sz"]:
            raise ValueError(f'model.feature_sz {conf["model"]["feature_sz"]} != n_mels * n_stack = {ecfg.feature_sz}')
        m._ecfg = dataclasses.replace(ecfg, gemm_mode=conf.get("gemm_mode", 1), blank=m.blank, bos=m.bos, lm_layers=0)
        m = m.to(conf["cuda"]["device"])
        m.mp = conf.get("mp", False)
        return m

    def param_groups(self):
        return [self.encoder.param_groups(), self.predictor.param_groups(), self.joint.param_groups()]

    def convert_to_cpu(self):
        raise RuntimeError("libreasr_b200 is CUDA (sm_100a) only: there is no CPU inference path")

    def convert_to_gpu(self):
        self.engine()
        return self

    def load_state_dict(self, state_dict, strict=True):
        r = super().load_state_dict(state_dict, strict=strict)
        self._drop_engine()
        return r

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._drop_engine()
        return r

    def _drop_engine(self):
        if getattr(self, "_engine", None) is not None:
            self._engine.close()
        self._engine = None

    def engine(self) -> Engine:
        """The CUDA engine holding the repacked weights (built lazily from the parameters)."""
        if self._engine is None:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("Transducer parameters are on %s; move the model to a CUDA device "
                                   "(libreasr_b200 has no CPU path)" % dev)
            import dataclasses

            ecfg, lm_sd = self._ecfg, None
            sd = {k: v for k, v in self.state_dict().items() if not k.startswith("lm.")}
            if self.lm is not None:  # fused LM (models.py:234, api-server.py:158-161): weights go to the engine as "lm.*"
                lm = self.lm
                ecfg = dataclasses.replace(ecfg, lm_layers=lm.num_layers, lm_hidden_sz=lm.hidden_sz, lm_embed_sz=lm.embed_sz)
                lm_sd = {k: v for k, v in lm.state_dict().items()}
            eng = Engine(ecfg, device=dev)
            eng.load_state_dict(sd, lm_state_dict=lm_sd)
            self._engine = eng
            object.__setattr__(self, "_engine_lm", self.lm)
        return self._engine

    def __setattr__(self, name, value):
        super().__setattr__(name, value)
        if name == "lm" and getattr(self, "_engine", None) is not None and value is not getattr(self, "_engine_lm", None):
            self._drop_engine()  # the engine bakes the LM in at finalize

    def start_perf(self):
        """models.py:278-280 (the device is synchronised so that the figure is the stage's time, not its launch)."""
        if self.perf:
            import time

            torch.cuda.synchronize()
            self.t = time.time()

    def stop_perf(self, name="unknown"):
        if self.perf:  # models.py:282-285
            import time

            torch.cuda.synchronize()
            t = (time.time() - self.t) * 1000.0
            print(f"{name.ljust(10, ' ')} | {t:4.2f}ms")

    def forward(self, tpl):
        """``(x, y, xl, yl)`` -> the log_softmax joint lattice [N, T, U, V] (models.py:308-359), EVAL MODE ONLY: the engine has no
        autograd, so this serves validation / loss evaluation (``libreasr_b200.lib.loss``), not a training step.  Like the
        reference outside training, the predictor is teacher-forced over ``cat(bos, y)`` (grab_bos, models.py:286-306)."""
        if self.training:
            raise NotImplementedError("Transducer.forward in training mode needs gradients; the B200 path is inference-only "
                                      "(call .eval() for the validation-time lattice / loss)")
        x, y, xl, yl = tpl
        eng = self.engine()
        x = x.to(eng.device, torch.float32)
        x = x.reshape(x.size(0), x.size(1), -1)                      # models.py:319
        r = eng.forward_loss(x, xl, y, yl, want_lattice=True)
        self._last_loss = r["loss"]
        return r["lattice"]

    # ---- inference (models.py:361-455) ---------------------------------------------------
    def decode(self, *args, **kwargs):
        res, log_p, _ = self.decode_greedy(*args, **kwargs)[:3]
        return res, log_p

    def transcribe(self, *args, **kwargs):
        res, _, metrics, _ = self.decode_greedy(*args, **kwargs)
        return res, metrics

    def decode_greedy(self, x, max_iters=3, alpha=0.005, theta=1.0, return_logp=False):
        """x [T,X,1] (what the inference transform pipeline yields, api-server.py:75-78) or
        [T,X]: features of one utterance (models.py:369-455).  With ``self.lm`` set the LM is fused inside the
        decode loop with the fuser's constants (lm.py:13-14; like the reference, the ``alpha``/``theta`` arguments are
        not forwarded to ``fuse``, models.py:431)."""
        eng = self.engine()
        x = x.to(eng.device, torch.float32)
        if x.dim() == 2:
            x = x[:, :, None]
        x = x[None]  # models.py:391
        x = x.reshape(x.size(0), x.size(1), -1)  # Encoder.forward, models.py:106
        enc, _ = eng.encode(x)
        T = enc.shape[1]
        r = eng.decode_greedy(enc, max_iters=max_iters, trace_cap=(max_iters * T if return_logp else 0))
        y_seq = tokens_to_lists(r["tokens"], r["ntok"])[0]
        iters = r["iters"][0].cpu().numpy().astype(np.int64)
        _sum = iters.sum()
        _ones = int((iters == 1).sum())
        metrics = {"alignment_score": (_sum - _ones) / (_sum + 1e-4)}  # models.py:445-453
        extra = {"iters": iters.tolist(), "outs": []}
        if return_logp:
            n_eval = int(_sum)
            extra["outs"] = [o[None, None, None] for o in r["trace"][0, :n_eval]]
        return self.lang.denumericalize(y_seq), float(r["neg_logp"][0]), metrics, extra

    def decode_beam(self, x, width=4, max_iters=3):
        """Beam search for one utterance x [T,X(,1)] -> (text, score).  The reference has no beam search (models.py:8 is
        an unused PriorityQueue import); the algorithm is the one defined in oracle/beam.py on the reference's own
        Predictor / Joint, keeping ``decode_greedy``'s max_iters rule."""
        eng = self.engine()
        x = x.to(eng.device, torch.float32)
        if x.dim() == 2:
            x = x[:, :, None]
        x = x[None].reshape(1, x.size(0), -1)
        enc, _ = eng.encode(x)
        r = eng.decode_beam(enc, None, width=width, max_iters=max_iters)
        y_seq = tokens_to_lists(r["tokens"], r["ntok"])[0]
        return self.lang.denumericalize(y_seq), float(r["score"][0])

    def transcribe_stream(self, stream, denumericalizer, max_iters=10, alpha=0.3, theta=1.0):
        """Generator over chunks of shape [T_c, X(,1)] or None (models.py:457-577): yields
        (all tokens so far, denumericalizer(tokens of this chunk), reset_fn).  One LM fuser lives for the whole
        stream (models.py:478) until ``reset_fn`` is called (models.py:491-500)."""
        eng = self.engine()
        st = {"enc": None, "pred": None, "lm": eng.new_lm_state(1) if self.lm is not None else None}

        def reset():  # models.py:480-500
            st["enc"], st["pred"] = None, None
            if st["lm"] is not None:
                st["lm"].zero_()  # LMFuser.reset (lm.py:81-83)

        y = []
        for chunk in stream:
            if chunk is None:  # models.py:509
                continue
            x = chunk.to(eng.device, torch.float32)
            x = x.reshape(1, x.size(0), -1)
            self.start_perf()
            enc, st["enc"] = eng.encode(x, state=st["enc"], want_state=True)
            self.stop_perf("encoder")  # models.py:516-524
            self.start_perf()
            r = eng.decode_greedy(enc, max_iters=max_iters, state=st["pred"], want_state=True, lm_state=st["lm"])
            self.stop_perf("joint + predictor")  # models.py:526-577
            st["pred"] = r["state"]
            y_seq = tokens_to_lists(r["tokens"], r["ntok"])[0]
            y = y + y_seq
            yield y, denumericalizer(y_seq), reset
