"""TEST INFRASTRUCTURE ONLY (oracle) -- CPU restatement of the CNN half of the hot path.

PARITY UNPINNED for the network arithmetic: the reference delegates it to Keras / TensorFlow
(un-vendored, un-pinned third-party packages: conda-recipe/meta.yaml:18-19 lists bare ``keras`` and
``tensorflow-gpu``; era Keras 2.0-2.1 / TF 1.3-1.8), neither is installed in this image and there is
no network, and the reference has no tests or golden vectors at this boundary.  This module restates
the published Keras-2 layer semantics on torch-CPU float64 and anchors on the reference's own call
sites:

  * graph definitions   flypylib/fplmodels.py:102-136 (vgg_like), :138-172 (vgg_like2),
                        :258-304 (unet_like2), helper _bn_relu :67-71
  * inference rebuild   flypylib/fplnetwork.py:99-110 (UpSampling3D(rf_stride) appended for VGGs)
  * tiling / scatter    flypylib/fplnetwork.py:136-189 -- this part IS pinned: ``infer_tiler`` below
                        is checked against the unmodified reference method run with a fake network
                        (tests/golden/infer_tiler_golden.npz)

Keras-2 semantics used: Conv3D padding='valid', stride 1, channels_last, kernel (kd,kh,kw,Cin,Cout),
cross-correlation; BatchNormalization(axis=-1, epsilon=1e-3) inference with moving statistics,
weights [gamma, beta, moving_mean, moving_variance]; MaxPooling3D 2/2 'valid' (floor);
UpSampling3D nearest repeat; Cropping3D symmetric; concatenate(axis=-1) in argument order;
Dropout identity at inference; sigmoid; glorot_uniform kernel init, zero bias.
"""
import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3

# (kind, k, cin, cout) chains; 'P' pool, 'U' up-sample+concat with a stored skip (crop), 'S' store skip;
# unet_like_vol: 'CN' conv + ReLU without BatchNormalization;
# resnet_like adds: 'CB' conv + BN without ReLU, 'CS' plain conv (no BN, no activation) applied to a stored skip,
# 'A' add the symmetrically cropped skip, then ReLU
ARCHS = {
    # name: (ops, (rf_size, rf_offset, rf_stride), infer_sz, final_has_bias)
    "vgg_like": ([("C", 3, 1, 48), ("C", 1, 48, 48), ("P",), ("C", 3, 48, 48), ("C", 1, 48, 48), ("P",),
                  ("C", 3, 48, 48), ("C", 1, 48, 96), ("C", 1, 96, 96), ("F", 96)],
                 (18, 7, 4), 102, True),
    "vgg_like2": ([("C", 3, 1, 48), ("C", 3, 48, 48), ("P",), ("C", 3, 48, 48), ("C", 3, 48, 48), ("P",),
                   ("C", 3, 48, 48), ("C", 1, 48, 96), ("C", 1, 96, 96), ("F", 96)],
                  (24, 10, 4), 100, True),
    "unet_like2": ([("C", 3, 1, 32), ("C", 3, 32, 32), ("S", "conv1"), ("P",),
                    ("C", 3, 32, 64), ("C", 3, 64, 64), ("S", "conv2"), ("P",),
                    ("C", 1, 64, 128), ("U", "conv2", 0),
                    ("C", 3, 192, 64), ("C", 1, 64, 64), ("U", "conv1", 6),
                    ("C", 3, 96, 32), ("C", 1, 32, 32), ("F", 32)],
                   (24, 9, 1), 100, False),
    # fplmodels.py:174-208.  Weight order = Keras model.layers order: the shortcut convolution (created after conv3b, but
    # reached first when Keras walks the graph back from the output: add([crop_pool2, conv3])) precedes conv3b.
    "resnet_like": ([("C", 3, 1, 32), ("P",), ("S", "pool1"),
                     ("C", 3, 32, 32), ("CB", 1, 32, 32), ("A", "pool1", 1), ("P",), ("S", "pool2"),
                     ("C", 3, 32, 64), ("CS", 1, 32, 64, "pool2"), ("CB", 1, 64, 64), ("A", "pool2", 1), ("F", 64)],
                    (18, 7, 4), 102, True),
    # fplmodels.py:470-526: no BatchNormalization anywhere (Conv3D(activation='relu', use_bias=False)); trains volume to
    # volume, rf tuple (62, 6, 1) as the builder returns it
    "unet_like_vol": ([("CN", 3, 1, 16), ("CN", 1, 16, 16), ("S", "conv1"), ("P",),
                       ("CN", 3, 16, 32), ("CN", 1, 32, 32), ("S", "conv2"), ("P",),
                       ("CN", 1, 32, 64), ("U", "conv2", 0),
                       ("CN", 3, 96, 64), ("CN", 1, 64, 64), ("U", "conv1", 4),
                       ("CN", 3, 80, 32), ("CN", 1, 32, 32), ("F", 32)],
                      (62, 6, 1), 102, False),
    # further builders with the same layer vocabulary: fplmodels.py:73-100, :206-256, :306-357, :359-410, :412-467
    "baseline_model": ([("C", 3, 1, 32), ("P",), ("C", 3, 32, 32), ("P",), ("C", 3, 32, 32), ("C", 1, 32, 64), ("F", 64)],
                       (18, 7, 4), 102, True),
    "unet_like": ([("C", 3, 1, 32), ("C", 1, 32, 32), ("S", "conv1"), ("P",),
                   ("C", 3, 32, 64), ("C", 1, 64, 64), ("S", "conv2"), ("P",),
                   ("C", 1, 64, 128), ("U", "conv2", 0),
                   ("C", 3, 192, 64), ("C", 1, 64, 64), ("U", "conv1", 4),
                   ("C", 3, 96, 32), ("C", 1, 32, 32), ("F", 32)],
                  (18, 6, 1), 102, False),
    "unet_like3": ([("C", 3, 1, 32), ("C", 3, 32, 32), ("S", "conv1"), ("P",),
                    ("C", 3, 32, 64), ("C", 3, 64, 64), ("S", "conv2"), ("P",),
                    ("C", 3, 64, 128), ("C", 1, 128, 128), ("U", "conv2", 2),
                    ("C", 3, 192, 64), ("C", 1, 64, 64), ("U", "conv1", 10),
                    ("C", 3, 96, 32), ("C", 1, 32, 32), ("F", 32)],
                   (32, 13, 1), 100, False),
    "unet_like4": ([("C", 3, 1, 32), ("C", 3, 32, 32), ("S", "conv1"), ("P",),
                    ("C", 3, 32, 64), ("C", 3, 64, 64), ("S", "conv2"), ("P",),
                    ("C", 3, 64, 128), ("C", 3, 128, 128), ("U", "conv2", 4),
                    ("C", 3, 192, 64), ("C", 1, 64, 64), ("U", "conv1", 14),
                    ("C", 3, 96, 32), ("C", 1, 32, 32), ("F", 32)],
                   (40, 17, 1), 100, False),
    "unet_like4b": ([("C", 3, 1, 32), ("C", 3, 32, 32), ("S", "conv1"), ("P",),
                     ("C", 3, 32, 64), ("C", 1, 64, 32), ("C", 3, 32, 64), ("S", "conv2"), ("P",),
                     ("C", 1, 64, 48), ("C", 3, 48, 128), ("C", 1, 128, 48), ("C", 3, 48, 128), ("C", 1, 128, 48),
                     ("U", "conv2", 4),
                     ("C", 3, 112, 64), ("C", 1, 64, 64), ("U", "conv1", 14),
                     ("C", 3, 96, 32), ("C", 1, 32, 32), ("F", 32)],
                    (40, 17, 1), 100, False),
}


def weight_shapes(arch):
    """Shapes of Model.get_weights() in Keras order."""
    ops, _, _, final_bias = ARCHS[arch]
    shapes = []
    for op in ops:
        if op[0] in ("C", "CB"):
            _, k, cin, cout = op
            shapes.append((k, k, k, cin, cout))
            shapes += [(cout,)] * 4
        elif op[0] in ("CS", "CN"):
            shapes.append((op[1],) * 3 + (op[2], op[3]))
        elif op[0] == "F":
            shapes.append((1, 1, 1, op[1], 1))
            if final_bias:
                shapes.append((1,))
    return shapes


def random_weights(arch, seed=4321, trained_like=True):
    """glorot_uniform kernels; BN statistics non-trivial when trained_like."""
    rng = np.random.default_rng(seed)
    ops, _, _, final_bias = ARCHS[arch]
    ws = []
    for op in ops:
        if op[0] in ("CS", "CN"):
            k, cin, cout = op[1], op[2], op[3]
            lim = np.sqrt(6.0 / (k ** 3 * cin + k ** 3 * cout))
            ws.append(rng.uniform(-lim, lim, (k, k, k, cin, cout)).astype(np.float32))
        elif op[0] in ("C", "CB"):
            _, k, cin, cout = op
            lim = np.sqrt(6.0 / (k ** 3 * cin + k ** 3 * cout))
            ws.append(rng.uniform(-lim, lim, (k, k, k, cin, cout)).astype(np.float32))
            if trained_like:
                ws.append(rng.uniform(0.5, 1.5, cout).astype(np.float32))       # gamma
                ws.append((0.1 * rng.standard_normal(cout)).astype(np.float32))  # beta
                ws.append((0.1 * rng.standard_normal(cout)).astype(np.float32))  # moving_mean
                ws.append(rng.uniform(0.5, 1.5, cout).astype(np.float32))       # moving_var
            else:
                ws += [np.ones(cout, np.float32), np.zeros(cout, np.float32),
                       np.zeros(cout, np.float32), np.ones(cout, np.float32)]
        elif op[0] == "F":
            cin = op[1]
            lim = np.sqrt(6.0 / (cin + 1))
            ws.append(rng.uniform(-lim, lim, (1, 1, 1, cin, 1)).astype(np.float32))
            if final_bias:
                ws.append(np.zeros(1, np.float32))
    return ws


def forward(arch, weights, x, dtype=torch.float64, upsample=True):
    """x: (N, D, H, W) array (single channel). Returns (N, o, o, o) numpy array of `dtype`.

    upsample=True applies the inference-time UpSampling3D(rf_stride) of fplnetwork.py:99-105."""
    ops, rf, _, final_bias = ARCHS[arch]
    t = torch.as_tensor(np.asarray(x)).to(dtype)[:, None]          # N,1,D,H,W
    wi = 0
    skips = {}
    for op in ops:
        if op[0] in ("C", "CB"):
            kern = torch.as_tensor(weights[wi]).to(dtype).permute(4, 3, 0, 1, 2)
            gamma, beta, mean, var = (torch.as_tensor(w).to(dtype) for w in weights[wi + 1:wi + 5])
            wi += 5
            t = F.conv3d(t, kern)
            t = (t - mean.view(1, -1, 1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1, 1) + BN_EPS) \
                * gamma.view(1, -1, 1, 1, 1) + beta.view(1, -1, 1, 1, 1)
            if op[0] == "C":
                t = torch.relu(t)
        elif op[0] == "CN":                       # Conv3D(activation='relu', use_bias=False), no BatchNormalization
            kern = torch.as_tensor(weights[wi]).to(dtype).permute(4, 3, 0, 1, 2)
            wi += 1
            t = torch.relu(F.conv3d(t, kern))
        elif op[0] == "CS":                       # plain convolution of a stored tensor (resnet shortcut)
            kern = torch.as_tensor(weights[wi]).to(dtype).permute(4, 3, 0, 1, 2)
            wi += 1
            skips[op[4]] = F.conv3d(skips[op[4]], kern)
        elif op[0] == "A":                        # add([Cropping3D(skip), t]) -> ReLU
            sk = skips[op[1]]
            c = op[2]
            if c:
                sk = sk[:, :, c:-c, c:-c, c:-c]
            t = torch.relu(sk + t)
        elif op[0] == "P":
            t = F.max_pool3d(t, 2)
        elif op[0] == "S":
            skips[op[1]] = t
        elif op[0] == "U":
            up = t.repeat_interleave(2, 2).repeat_interleave(2, 3).repeat_interleave(2, 4)
            sk = skips[op[1]]
            c = op[2]
            if c:
                sk = sk[:, :, c:-c, c:-c, c:-c]
            t = torch.cat([up, sk], 1)
        elif op[0] == "F":
            kern = torch.as_tensor(weights[wi]).to(dtype).permute(4, 3, 0, 1, 2)
            wi += 1
            t = F.conv3d(t, kern)
            if final_bias:
                t = t + torch.as_tensor(weights[wi]).to(dtype).view(1, -1, 1, 1, 1)
                wi += 1
            t = torch.sigmoid(t)
    assert wi == len(weights)
    t = t[:, 0]
    s = rf[2]
    if upsample and s != 1:
        t = t.repeat_interleave(s, 1).repeat_interleave(s, 2).repeat_interleave(s, 3)
    return t.numpy()


class TorchNet:
    """Duck-typed ``infer_network`` for infer_tiler / the reference tiler: input_shape + predict."""

    def __init__(self, arch, weights, infer_sz=None, dtype=torch.float32):
        self.arch, self.weights, self.dtype = arch, weights, dtype
        s = infer_sz if infer_sz is not None else ARCHS[arch][2]
        self.input_shape = (None, s, s, s, 1)

    def predict(self, x, batch_size=1):
        outs = []
        for i in range(0, x.shape[0], max(1, batch_size)):
            xb = np.asarray(x[i:i + batch_size, ..., 0])
            outs.append(forward(self.arch, self.weights, xb, self.dtype).astype(np.float32)[..., None])
        return np.concatenate(outs, 0)


def infer_tiler(image, infer_network, infer_sz, rf_offset, n_gpu=1):
    """Restatement of FplNetwork.infer (flypylib/fplnetwork.py:136-189): tile grid with origins
    k*(infer_sz-2*off), zero-padded far-edge tiles (float64 staging batch, padded to a multiple of
    n_gpu), predict, scatter of the valid interior; the off-wide border stays 0."""
    image = np.asarray(image)
    size = np.array(image.shape)
    isz = np.array(infer_sz)
    off = np.array(rf_offset)
    out = isz - 2 * off
    axes = [np.arange(off[a], size[a] - off[a], out[a]) for a in range(3)]
    origins = np.stack(np.meshgrid(*axes, indexing="ij"), 0).reshape(3, -1)
    n = origins.shape[1]
    n_batch = int(np.ceil(n / float(n_gpu)) * n_gpu)
    batch = np.zeros((n_batch, isz[0], isz[1], isz[2], 1))
    spans = []
    for i in range(n):
        lo = origins[:, i] - off
        hi = np.minimum(origins[:, i] + out + off, size)
        ext = hi - lo
        batch[i, :ext[0], :ext[1], :ext[2], 0] = image[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]
        spans.append((origins[:, i], hi - off, ext - 2 * off))
    pb = infer_network.predict(batch, batch_size=n_gpu)
    pred = np.zeros(image.shape, dtype="float32")
    for i, (lo, hi, ext) in enumerate(spans):
        pred[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = pb[i, :ext[0], :ext[1], :ext[2], 0]
    return pred


def detector_weights(arch, sample, seed=7, final_logit=(-3.0, 6.0)):
    """Weights that make `arch` a (crude but genuine) bright-blob detector, for the detection-F1 test.

    Every kernel is |glorot_uniform| (non-negative) and every BatchNormalization is *calibrated* on
    `sample` ((N,D,H,W) float32, already normalised): moving_mean / moving_variance are the measured
    per-channel statistics of the convolution output, gamma = 1, beta = 0 -- what a trained network's
    moving statistics look like.  Non-negative kernels, positive BN scale, ReLU, max-pool, up-sampling
    and concatenation are all monotone, so the output probability is a monotone function of local
    brightness: planted bright blobs become peaks of the probability map.  The final bias (VGGs) shifts
    the median logit to final_logit[0]; the final kernel is scaled so that the sample's logit spread
    (99.9th percentile - median) is final_logit[1]."""
    rng = np.random.default_rng(seed)
    ops, _, _, final_bias = ARCHS[arch]
    t = torch.as_tensor(np.asarray(sample, dtype=np.float32))[:, None]
    ws, skips = [], {}
    for op in ops:
        if op[0] == "C":
            _, k, cin, cout = op
            lim = np.sqrt(6.0 / (k ** 3 * cin + k ** 3 * cout))
            kern = rng.uniform(0, lim, (k, k, k, cin, cout)).astype(np.float32)
            t = F.conv3d(t, torch.as_tensor(kern).permute(4, 3, 0, 1, 2))
            mean = t.mean(dim=(0, 2, 3, 4)); var = t.var(dim=(0, 2, 3, 4), unbiased=False)
            ws += [kern, np.ones(cout, np.float32), np.zeros(cout, np.float32), mean.numpy().astype(np.float32),
                   var.numpy().astype(np.float32)]
            t = torch.relu((t - mean.view(1, -1, 1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1, 1) + BN_EPS))
        elif op[0] == "P":
            t = F.max_pool3d(t, 2)
        elif op[0] == "S":
            skips[op[1]] = t
        elif op[0] == "U":
            up = t.repeat_interleave(2, 2).repeat_interleave(2, 3).repeat_interleave(2, 4)
            sk = skips[op[1]]
            if op[2]:
                sk = sk[:, :, op[2]:-op[2], op[2]:-op[2], op[2]:-op[2]]
            t = torch.cat([up, sk], 1)
        elif op[0] == "F":
            cin = op[1]
            kern = rng.uniform(0, np.sqrt(6.0 / (cin + 1)), (1, 1, 1, cin, 1)).astype(np.float32)
            logit = F.conv3d(t, torch.as_tensor(kern).permute(4, 3, 0, 1, 2)).flatten()
            med = float(logit.median()); hi = float(torch.quantile(logit[:: max(1, logit.numel() // 1000000)], 0.999))
            scale = final_logit[1] / max(hi - med, 1e-6)
            ws.append((kern * np.float32(scale)).astype(np.float32))
            if final_bias:
                ws.append(np.asarray([final_logit[0] - med * scale], np.float32))
    return ws
