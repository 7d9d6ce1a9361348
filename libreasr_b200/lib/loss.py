"""RNN-T loss for the B200 path: the inference-side counterpart of ``libreasr/lib/loss.py:72-110``.

``get_loss_func("rnnt", ...)`` returns a function with the reference's ``_loss_func(inp, tgt, reduction)`` signature: ``inp`` is
the log-probability lattice ``Transducer.forward`` returns, ``tgt = (labels, label_lens, input_lens)``.  The value is what
``warp_rnnt.rnnt_loss(..., average_frames=False)`` defines -- the per-sequence negative log-likelihood by the forward
recursion over the lattice -- computed on the GPU (``rnnt_b200_rnnt_loss``); no gradients (validation / scoring).
"""
import torch


def get_loss_func(loss_type, engine, reduction_factor=1, zero_nan=False, zero_inf=False, **_):
    if loss_type != "rnnt":
        raise Exception(f"no such loss type: {loss_type}")          # loss.py:88-89 (CTC is outside the built path)

    def _loss_func(inp, tgt, reduction="mean", **kwargs):
        tgt_labels, tgt_lens, inp_lens = tgt
        inp_lens = torch.as_tensor(inp_lens).to(torch.int32) // reduction_factor      # loss.py:100-102
        if zero_nan:
            inp = torch.where(torch.isnan(inp), torch.zeros_like(inp), inp)
        if zero_inf:
            inp = torch.where(torch.isinf(inp), torch.zeros_like(inp), inp)
        loss = engine.rnnt_loss(inp, inp_lens, tgt_labels, tgt_lens)
        return loss.mean() if reduction == "mean" else loss
    return _loss_func
