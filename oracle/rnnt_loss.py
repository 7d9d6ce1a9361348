"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the training-time forward (joint lattice) and the RNN-T loss.

``forward_lattice`` follows ``Transducer.forward`` (libreasr/lib/models.py:308-359) in eval mode: encoder over the padded
batch with lengths, the predictor teacher-forced over ``cat(bos, y)`` (``grab_bos`` returns the plain BOS column outside
training, models.py:286-306), the joint network on the broadcast ``[N, T, U, H]`` pair, ``log_softmax``.  Pinned by
``tests/golden/tiny_forward.npz`` (the imported reference's own ``forward``).

``rnnt_loss`` restates what ``get_loss_func("rnnt")`` computes per sequence (libreasr/lib/loss.py:72-110:
``warp_rnnt.rnnt_loss(log_probs, labels, frames_lengths, labels_lengths, average_frames=False)``): the negative
log-likelihood of the label sequence under the transducer lattice (Graves 2012), by the forward recursion

    alpha(0, 0) = 0
    alpha(t, u) = logaddexp(alpha(t-1, u) + lp[t-1, u, blank],  alpha(t, u-1) + lp[t, u-1, y_u])
    loss = -(alpha(T-1, U) + lp[T-1, U, blank])

PARITY UNPINNED for the loss value: ``warp_rnnt`` is an un-vendored dependency that is not installed here; the recursion
above is its published definition (and the reduction ``mean`` is applied by the caller, loss.py:63-66).
"""
import numpy as np
import torch
import torch.nn.functional as F


def forward_lattice(orc, feats: torch.Tensor, y: torch.Tensor, xl, yl):
    """feats [N,T,X], y [N,Umax] int64 (padded with 0), xl [N] encoder steps, yl [N] label lengths
    -> log_softmax lattice [N, T, Umax+1, V] (positions beyond (xl, yl) hold the values of the padded computation)."""
    N, T, _ = feats.shape
    with torch.no_grad():
        encs = []
        for n in range(N):
            e, _ = orc.encoder(feats[n:n + 1, : int(xl[n])], None, "aten")
            pad = torch.zeros(T - e.shape[1], e.shape[2])
            encs.append(torch.cat([e[0], pad], 0))
        enc = torch.stack(encs)                                            # [N,T,H]
        bos = torch.full((N, 1), orc.bos, dtype=torch.long)
        yc = torch.cat([bos, y.long()], 1)                                  # models.py:331-332
        outs, state = [], None
        for u in range(yc.shape[1]):
            g, state = orc.predictor(yc[:, u], state)
            outs.append(g)
        pred = torch.stack(outs, 1)                                         # [N,U,H]
        U = pred.shape[1]
        joint = orc.joint(pred[:, None].expand(N, T, U, -1), enc[:, :, None].expand(N, T, U, -1))
        return F.log_softmax(joint, -1)


def rnnt_loss(lattice, y, xl, yl, blank=0):
    """lattice [N,T,U,V] log-probabilities; -> np.float64 [N] negative log-likelihoods."""
    lat = lattice.detach().double().numpy() if torch.is_tensor(lattice) else np.asarray(lattice, dtype=np.float64)
    out = np.zeros(lat.shape[0])
    for n in range(lat.shape[0]):
        T, U = int(xl[n]), int(yl[n])
        a = np.full((T, U + 1), -np.inf)
        a[0, 0] = 0.0
        for t in range(T):
            for u in range(U + 1):
                if t == 0 and u == 0:
                    continue
                v = -np.inf
                if t > 0:
                    v = np.logaddexp(v, a[t - 1, u] + lat[n, t - 1, u, blank])
                if u > 0:
                    v = np.logaddexp(v, a[t, u - 1] + lat[n, t, u - 1, int(y[n][u - 1])])
                a[t, u] = v
        out[n] = -(a[T - 1, U] + lat[n, T - 1, U, blank])
    return out
