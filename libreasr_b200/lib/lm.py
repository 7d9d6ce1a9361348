"""Fused language model: the reference module surface of ``libreasr/lib/lm.py``.

``LM`` keeps the reference constructor and ``state_dict`` keys (lm.py:20-29) so that the checkpoints
``load_lm`` reads (lm.py:86-101) load unchanged.  It is a parameter container: the LM never runs as a
PyTorch module here.  Attached to a ``Transducer`` (``m.lm = lm``, api-server.py:158-161) its weights
are handed to the CUDA engine under the ``lm.`` prefix, and the fuser of ``LMFuser`` (lm.py:43-83 --
standardise both rows, pin blank to -10, arg max of ``0.1 * lm + 1.0 * joint`` after the blank test,
advance the LM after every emitted token) runs inside the persistent decode kernel
(``libreasr_b200/csrc/decode.cu``), per utterance / stream.
"""
import torch
import torch.nn as nn

ALPHA = 0.1  # lm.py:13
THETA = 1.0  # lm.py:14
MIN_VAL = -10.0  # lm.py:15


class LM(nn.Module):
    """Parameter container with the reference's attribute names (they define the ``state_dict`` keys: ``embed.weight``,
    ``rnn.weight_ih_l{k}`` ..., ``linear.weight`` / ``linear.bias``)."""

    def __init__(self, vocab_sz, embed_sz, hidden_sz, num_layers, p=0.2, **kwargs):
        super().__init__()
        self.vocab_sz, self.embed_sz, self.hidden_sz, self.num_layers = int(vocab_sz), int(embed_sz), int(hidden_sz), int(num_layers)
        shapes = dict(embed=nn.Embedding(self.vocab_sz, self.embed_sz, padding_idx=0),                  # blank row stays zero
                      rnn=nn.LSTM(input_size=self.embed_sz, hidden_size=self.hidden_sz, num_layers=self.num_layers, batch_first=True),
                      drop=nn.Dropout(p),                                                               # identity at inference
                      linear=nn.Linear(self.hidden_sz, self.vocab_sz))
        for name, mod in shapes.items():
            setattr(self, name, mod)
        if self.embed_sz == self.hidden_sz:   # weight tying (lm.py:27-29): one tensor under both keys
            self.linear.weight = self.embed.weight

    def forward(self, x, state=None):
        raise NotImplementedError("the language model runs fused inside the CUDA decode loop (attach it with "
                                  "`transducer.lm = lm`); there is no PyTorch/CPU evaluation path")


def load_lm(conf, lang=None):
    """lm.py:86-101 without the CPU quantisation step (``maybe_quantize`` targets the reference's CPU path)."""
    kw = {k: v for k, v in conf["lm"].items() if k in ("vocab_sz", "embed_sz", "hidden_sz", "num_layers", "p")}
    lm = LM(**kw)
    lm.load_state_dict(torch.load(conf["lm"]["path"], map_location="cpu"))
    lm.eval()
    return lm
