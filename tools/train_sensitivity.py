"""How badly conditioned is the training step the parity tests run?  CPU only, oracle only (test infrastructure).

Re-runs oracle/train_oracle.py's float64 step with i.i.d. relative noise on every convolution output (1e-7: fp32
arithmetic; 7e-6: one bf16 hi/lo x3 tensor-core contraction; 2.5e-3: one bf16 contraction) and prints how far the
gradients move -- the basis of the tolerances in tests/test_train_gpu.py.   python tools/train_sensitivity.py
"""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from oracle import models_oracle as M, train_oracle as T

def fb(arch, weights, x, labels, global_batch, seed, dt, noise=0.0, bf=False):
    ops, _, _, final_bias = M.ARCHS[arch]
    ws = [torch.tensor(np.asarray(w), dtype=dt, requires_grad=True) for w in weights]
    t = torch.as_tensor(np.asarray(x), dtype=dt)[:, None]
    wi, li = 0, 0
    g = torch.Generator().manual_seed(5)
    for op in ops:
        if op[0] == "C":
            kern = ws[wi].permute(4, 3, 0, 1, 2); gamma, beta = ws[wi + 1], ws[wi + 2]; wi += 5
            t = F.conv3d(t, kern)
            if noise: t = t * (1 + noise * torch.randn(t.shape, generator=g, dtype=dt))
            mean = t.mean(dim=(0, 2, 3, 4)); var = t.var(dim=(0, 2, 3, 4), unbiased=False)
            t = (t - mean.view(1, -1, 1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1, 1) + M.BN_EPS)
            t = torch.relu(t * gamma.view(1, -1, 1, 1, 1) + beta.view(1, -1, 1, 1, 1))
            if li in (5, 6):
                cl = t.permute(0, 2, 3, 4, 1)
                keep = T.dropout_keep(seed, li, cl.numel()).reshape(tuple(cl.shape))
                t = (cl * torch.as_tensor(keep, dtype=dt) * 2.0).permute(0, 4, 1, 2, 3)
            li += 1
        elif op[0] == "P": t = F.max_pool3d(t, 2)
        elif op[0] == "F":
            kern = ws[wi].permute(4, 3, 0, 1, 2); t = F.conv3d(t, kern) + ws[wi + 1].view(1, -1, 1, 1, 1); wi += 2
    logit = t.reshape(-1)
    y = torch.as_tensor(np.asarray(labels).reshape(-1), dtype=dt)
    bce = F.binary_cross_entropy_with_logits(logit, y, reduction="sum")
    (bce / global_batch).backward()
    return [w.grad.numpy().astype(np.float64) if w.grad is not None else None for w in ws]

arch, batch = "vgg_like", 8
rf = M.ARCHS[arch][1][0]
w = M.random_weights(arch, seed=17)
rng = np.random.default_rng(18)
x = rng.standard_normal((batch, rf, rf, rf)).astype(np.float32)
y = (rng.random(batch) < 0.5).astype(np.uint8)
g64 = fb(arch, w, x, y, 32, 12345, torch.float64)
for name, kw in [("f32", dict(dt=torch.float32)), ("noise1e-7", dict(dt=torch.float64, noise=1e-7)), ("noise7e-6", dict(dt=torch.float64, noise=7e-6)), ("noise2e-3", dict(dt=torch.float64, noise=2e-3))]:
    g = fb(arch, w, x, y, 32, 12345, **kw)
    out = []
    for a, b in zip(g, g64):
        if a is None or np.abs(b).max() == 0: continue
        out.append(np.abs(a - b).max() / np.abs(b).max())
    print(name, " ".join("%.1e" % v for v in out))
print("--- batch sweep, noise 7e-6 (tf32 path model) and 2.5e-3 (bf16 model)")
for arch, batches in [("vgg_like", (8, 32, 64)), ("vgg_like2", (6, 24))]:
    rf = M.ARCHS[arch][1][0]
    for batch in batches:
        rng = np.random.default_rng(18)
        x = rng.standard_normal((batch, rf, rf, rf)).astype(np.float32)
        y = (rng.random(batch) < 0.5).astype(np.uint8)
        g64 = fb(arch, w if arch == "vgg_like" else M.random_weights(arch, seed=17), x, y, batch, 12345, torch.float64)
        for nz in (7e-6, 2.5e-3):
            g = fb(arch, w if arch == "vgg_like" else M.random_weights(arch, seed=17), x, y, batch, 12345, torch.float64, noise=nz)
            out = [np.abs(a - b).max() / np.abs(b).max() for a, b in zip(g, g64) if a is not None and np.abs(b).max() > 0]
            print(arch, batch, nz, "max %.1e  median %.1e  first %.1e" % (max(out), np.median(out), out[0]))
print("--- structured task (blob at the centre = positive)")
for arch, batch in [("vgg_like", 8), ("vgg_like2", 6)]:
    rf = M.ARCHS[arch][1][0]
    rng = np.random.default_rng(18)
    y = (np.arange(batch) % 2).astype(np.uint8)
    x = rng.standard_normal((batch, rf, rf, rf)).astype(np.float32) * 0.5
    c = rf // 2
    x[y == 1, c-2:c+2, c-2:c+2, c-2:c+2] += 2.0
    ww = M.random_weights(arch, seed=17)
    g64 = fb(arch, ww, x, y, batch, 12345, torch.float64)
    for nz in (1e-7, 7e-6, 2.5e-3):
        g = fb(arch, ww, x, y, batch, 12345, torch.float64, noise=nz)
        out = [np.abs(a - b).max() / np.abs(b).max() for a, b in zip(g, g64) if a is not None and np.abs(b).max() > 0]
        print(arch, batch, nz, "max %.1e  median %.1e  first %.1e" % (max(out), np.median(out), out[0]))
print("--- L2 metrics, white-noise inputs (the test's inputs)")
for arch, batch in [("vgg_like", 8), ("vgg_like2", 6)]:
    rf = M.ARCHS[arch][1][0]
    ww = M.random_weights(arch, seed=17)
    rng = np.random.default_rng(18)
    x = rng.standard_normal((batch, rf, rf, rf)).astype(np.float32)
    y = (rng.random(batch) < 0.5).astype(np.uint8)
    g64 = fb(arch, ww, x, y, 4*batch, 12345, torch.float64)
    for nz in (7e-6, 2e-5, 2.5e-3):
        g = fb(arch, ww, x, y, 4*batch, 12345, torch.float64, noise=nz)
        pairs = [(a, b) for a, b in zip(g, g64) if a is not None and np.abs(b).max() > 0]
        l2 = [np.linalg.norm(a - b) / np.linalg.norm(b) for a, b in pairs]
        fa = np.concatenate([a.ravel() for a, b in pairs]); fb_ = np.concatenate([b.ravel() for a, b in pairs])
        cos = fa @ fb_ / np.linalg.norm(fa) / np.linalg.norm(fb_)
        print(arch, nz, "relL2 max %.1e median %.1e  global relL2 %.1e  1-cos %.1e" % (max(l2), np.median(l2), np.linalg.norm(fa-fb_)/np.linalg.norm(fb_), 1-cos))
