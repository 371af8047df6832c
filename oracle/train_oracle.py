"""TEST INFRASTRUCTURE ONLY (oracle) -- CPU restatement of one training step of the VGG builders.

PARITY UNPINNED: the arithmetic belongs to Keras/TensorFlow (absent, un-pinned; see models_oracle.py).
Restated with torch autograd in float64 from the reference call sites:
  * compile defaults loss='binary_crossentropy', optimizer='adam', metrics=['accuracy']
    (flypylib/fplnetwork.py:74-79); one batch of fit_generator (fplnetwork.py:120-121)
  * graph in training mode: Conv3D(no bias) -> BatchNormalization (batch statistics, biased variance,
    eps 1e-3, momentum 0.99: Keras non-fused path for 5-D tensors) -> ReLU, Dropout(0.5) after full1/full2
    (fplmodels.py:102-136, :138-172), final Conv3D + bias + sigmoid
  * towers: every GPU normalises with its own batch statistics; the loss is the mean over the global
    batch, so tower gradients add up (flypylib/multi_gpu.py:20-61)
  * Adam, Keras defaults lr 1e-3, beta 0.9/0.999, epsilon 1e-7 (K.epsilon), no decay:
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps)
Dropout masks use the same counter-based hash as the device code so both sides drop the same units.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import models_oracle as M

MASK = (1 << 64) - 1


def _splitmix(z):
    z = np.asarray(z, dtype=np.uint64)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def dropout_keep(seed, layer, n):
    """keep bit for flat element indices 0..n-1 of a (N,z,y,x,C) tensor."""
    with np.errstate(over="ignore"):
        base = np.uint64(((seed + layer * 0x632BE59BD9B4E019) * 0x9E3779B97F4A7C15) & MASK)
        idx = np.arange(n, dtype=np.uint64) + base
        return (_splitmix(idx) & np.uint64(1)).astype(bool)


def split_params(arch, flat):
    shapes = M.weight_shapes(arch)
    out, o = [], 0
    for s in shapes:
        n = int(np.prod(s))
        out.append(np.asarray(flat[o:o + n]).reshape(s))
        o += n
    assert o == len(flat)
    return out


def flatten_params(ws):
    return np.concatenate([np.asarray(w, dtype=np.float64).ravel() for w in ws])


def forward_backward(arch, weights, x, labels, global_batch, seed):
    """Returns (loss_sum, n_correct, grads list in Keras order (zeros for moving stats), bn_batch list)."""
    ops, _, _, final_bias = M.ARCHS[arch]
    ws = [torch.tensor(np.asarray(w), dtype=torch.float64, requires_grad=True) for w in weights]
    t = torch.as_tensor(np.asarray(x), dtype=torch.float64)[:, None]
    wi, li = 0, 0
    bn_batch = []
    for op in ops:
        if op[0] == "C":
            kern = ws[wi].permute(4, 3, 0, 1, 2)
            gamma, beta = ws[wi + 1], ws[wi + 2]
            wi += 5
            t = F.conv3d(t, kern)
            mean = t.mean(dim=(0, 2, 3, 4))
            var = t.var(dim=(0, 2, 3, 4), unbiased=False)
            bn_batch.append((mean.detach().numpy().copy(), var.detach().numpy().copy()))
            t = (t - mean.view(1, -1, 1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1, 1) + M.BN_EPS)
            t = torch.relu(t * gamma.view(1, -1, 1, 1, 1) + beta.view(1, -1, 1, 1, 1))
            if li in (5, 6):
                cl = t.permute(0, 2, 3, 4, 1)                      # channels-last flat order of the device code
                keep = dropout_keep(seed, li, cl.numel()).reshape(tuple(cl.shape))
                t = (cl * torch.as_tensor(keep, dtype=torch.float64) * 2.0).permute(0, 4, 1, 2, 3)
            li += 1
        elif op[0] == "P":
            t = F.max_pool3d(t, 2)
        elif op[0] == "F":
            kern = ws[wi].permute(4, 3, 0, 1, 2)
            t = F.conv3d(t, kern) + ws[wi + 1].view(1, -1, 1, 1, 1)
            wi += 2
    logit = t.reshape(-1)
    y = torch.as_tensor(np.asarray(labels).reshape(-1), dtype=torch.float64)
    bce = F.binary_cross_entropy_with_logits(logit, y, reduction="sum")
    (bce / global_batch).backward()
    p = torch.sigmoid(logit).detach().numpy()
    correct = int(((p > 0.5) == (y.numpy() > 0.5)).sum())
    grads = [w.grad.numpy() if w.grad is not None else np.zeros(tuple(w.shape)) for w in ws]
    return float(bce.detach()), correct, grads, bn_batch


def adam_step(weights, grads, m, v, bn_batch, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7, momentum=0.99):
    """In-place Keras Adam on trainable arrays + moving-average update of the BN statistics."""
    lr_t = lr * np.sqrt(1.0 - b2 ** step) / (1.0 - b1 ** step)
    bi = 0
    i = 0
    while i < len(weights):
        w = weights[i]
        is_block = w.ndim == 5 and i + 4 < len(weights) and weights[i + 1].ndim == 1 and weights[i + 1].shape[0] == w.shape[4] \
            and not (w.shape[4] == 1)
        idxs = [i, i + 1, i + 2] if is_block else [i]
        for j in idxs:
            m[j] = b1 * m[j] + (1 - b1) * grads[j]
            v[j] = b2 * v[j] + (1 - b2) * grads[j] ** 2
            weights[j] = weights[j] - lr_t * m[j] / (np.sqrt(v[j]) + eps)
        if is_block:
            mean, var = bn_batch[bi]
            weights[i + 3] = weights[i + 3] * momentum + mean * (1 - momentum)
            weights[i + 4] = weights[i + 4] * momentum + var * (1 - momentum)
            bi += 1
            i += 5
        else:
            i += 1
    return weights, m, v
