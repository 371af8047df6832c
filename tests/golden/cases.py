"""Seeded synthetic inputs shared by the golden-vector generator, the parity tests and bench.py.

Only numpy is used so the arrays are reproducible on any machine with the same numpy.
"""
import numpy as np


def prob_map(shape, seed, kind="blobs", peaks_per_50cube=3.0, blob_sigma=3.0):
    """Synthetic float32 probability map (Z,Y,X).

    kind = 'blobs'    sparse planted peaks (separable Gaussian bumps, ~0.9-1.0, a few clipped
                      to exactly 1.0 to create exact ties) + U(0,0.02) noise
           'uniform'  i.i.d. U(0,1)  (adversarial: many "rim" detections)
           'ties'     blobs quantised to 1/16 steps (massive exact ties after smoothing)
           'saturated' >3 % of voxels share the maximum value -> reference returns 0 detections
           'zeros'    all zero
    """
    rng = np.random.default_rng(seed)
    Z, Y, X = shape
    if kind == "uniform":
        return rng.random(shape, dtype=np.float32)
    if kind == "zeros":
        return np.zeros(shape, dtype=np.float32)
    if kind == "saturated":
        a = rng.random(shape, dtype=np.float32) * np.float32(0.01)
        a[: max(1, Z // 2)] = 1.0
        return a
    vol = rng.random(shape, dtype=np.float32) * np.float32(0.02)
    n_peaks = max(1, int(round(peaks_per_50cube * Z * Y * X / 50.0 ** 3)))
    w = int(3 * blob_sigma)
    ax = np.arange(-w, w + 1, dtype=np.float64)
    g = np.exp(-0.5 * (ax / blob_sigma) ** 2)
    bump = (g[:, None, None] * g[None, :, None] * g[None, None, :]).astype(np.float32)
    cz = rng.integers(0, Z, n_peaks); cy = rng.integers(0, Y, n_peaks); cx = rng.integers(0, X, n_peaks)
    amp = (0.85 + 0.3 * rng.random(n_peaks)).astype(np.float32)
    for z, y, x, a in zip(cz, cy, cx, amp):
        z0, z1 = max(0, z - w), min(Z, z + w + 1)
        y0, y1 = max(0, y - w), min(Y, y + w + 1)
        x0, x1 = max(0, x - w), min(X, x + w + 1)
        vol[z0:z1, y0:y1, x0:x1] += a * bump[z0 - z + w:z1 - z + w, y0 - y + w:y1 - y + w,
                                              x0 - x + w:x1 - x + w]
    np.clip(vol, 0, 1, out=vol)
    if kind == "ties":
        vol = (np.round(vol * 16) / 16).astype(np.float32)
    return vol


def em_volume(shape, seed=1234):
    """EM-like uint8 volume: clip(128 + 33 * smooth unit-variance noise, 0, 255)  (SURVEY 8d)."""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal(shape, dtype=np.float32)
    # cheap separable box-smoothing (3 passes of width 3 per axis ~ sigma 2), numpy only
    for axis in range(3):
        for _ in range(3):
            a = (np.roll(a, 1, axis) + a + np.roll(a, -1, axis)) / np.float32(3)
    a = (a - a.mean()) / a.std()
    return np.clip(128 + 33 * a, 0, 255).astype(np.uint8)


# (name, shape, seed, kind, r, sigma, thd, buffer_sz, volume_offset)
VOXEL2OBJ_CASES = [
    ("blobs_48_r5_s1p5",      (48, 48, 48),  1, "blobs",     5, 1.5, 0,    2,          (0, 0, 0)),
    ("blobs_40x52x64_r6_s2",  (40, 52, 64),  2, "blobs",     6, 2.0, 0,    (3, 4, 5),  (100, 200, 300)),
    ("uniform_56_r7_s2",      (56, 56, 56),  3, "uniform",   7, 2.0, 0,    0,          (0, 0, 0)),
    ("ties_48_r5_s1",         (48, 48, 48),  4, "ties",      5, 1.0, 0,    0,          (0, 0, 0)),
    ("ties_40_r4_sigma0",     (40, 40, 40),  5, "ties",      4, 0.0, 0,    0,          (0, 0, 0)),
    ("reflect_32_r3_s4",      (32, 36, 40),  6, "blobs",     3, 4.0, 0,    1,          (0, 0, 0)),
    ("thd_64_r8_s2",          (64, 64, 64),  7, "blobs",     8, 2.0, 0.05, 5,          (-7, 0, 9)),
    ("saturated_32_r4_s1",    (32, 32, 32),  8, "saturated", 4, 1.0, 0,    0,          (0, 0, 0)),
    ("zeros_24_r3_s1",        (24, 24, 24),  9, "zeros",     3, 1.0, 0,    0,          (0, 0, 0)),
    ("uniform_96_r27_s5",     (96, 96, 96), 10, "uniform",  27, 5.0, 0,   30,          (0, 0, 0)),
    ("blobs_128_r27_s5",      (128, 128, 128), 11, "blobs", 27, 5.0, 0,   30,          (0, 0, 0)),
    ("blobs_thin_9x70x33_r4", (9, 70, 33),  12, "blobs",     4, 1.5, 0,    0,          (1, 2, 3)),
    ("uniform_160_r27_s5",    (160, 160, 160), 13, "uniform", 27, 5.0, 0, 15,          (0, 0, 0)),
]


class FakeNet:
    """Stands in for the Keras model inside the reference FplNetwork.infer: output voxel =
    mean of a (2*off+1)-cube would be costly; use centre-crop * 0.5 + tile-local ramp so that
    any mistake in tile origin, padding or scatter shows up."""

    def __init__(self, infer_sz, off, stride):
        self.input_shape = (None,) + tuple(infer_sz) + (1,)
        self.off = off
        self.calls = []

    def predict(self, x, batch_size=1):
        self.calls.append((x.shape, str(x.dtype), batch_size))
        o = self.off
        core = x[:, o[0]:x.shape[1] - o[0], o[1]:x.shape[2] - o[1], o[2]:x.shape[3] - o[2], :]
        zz, yy, xx = np.meshgrid(*[np.arange(n) for n in core.shape[1:4]], indexing="ij")
        ramp = (zz * 1e-3 + yy * 1e-5 + xx * 1e-7)[None, ..., None]
        return (core * 0.5 + ramp).astype(np.float32)




def blob_volume(shape, n_side, seed=5, amp=100.0, sigma=4.0, margin=34, noise=0.3):
    """EM-like uint8 volume with planted bright blobs (synthetic "T-bars") on a jittered n_side^3 grid;
    returns (volume, centres as (K,3) float64 in (x,y,z) order)."""
    rng = np.random.default_rng(seed)
    vol = 128.0 + noise * (em_volume(shape, seed=seed + 1).astype(np.float32) - 128.0)
    Z, Y, X = shape
    w = int(3 * sigma)
    ax = np.arange(-w, w + 1, dtype=np.float64)
    g = np.exp(-0.5 * (ax / sigma) ** 2)
    bump = (g[:, None, None] * g[None, :, None] * g[None, None, :]).astype(np.float32)
    centres = []
    for iz in range(n_side):
        for iy in range(n_side):
            for ix in range(n_side):
                if rng.random() < 0.25:
                    continue
                c = []
                for i, n in zip((iz, iy, ix), (Z, Y, X)):
                    cell = (n - 2 * margin) / n_side
                    c.append(int(margin + (i + 0.5) * cell + rng.uniform(-0.15, 0.15) * cell))
                z, y, x = c
                vol[z - w:z + w + 1, y - w:y + w + 1, x - w:x + w + 1] += amp * bump
                centres.append((x, y, z))
    return np.clip(vol, 0, 255).astype(np.uint8), np.asarray(centres, dtype=np.float64)


def segmentation(shape, seed, n_seeds=40):
    """Synthetic label volume (Z,Y,X) int64: nearest-seed (Voronoi) cells with a few tiny extra segments and some
    background (label 0)."""
    rng = np.random.default_rng(seed)
    Z, Y, X = shape
    pts = np.stack([rng.integers(0, Z, n_seeds), rng.integers(0, Y, n_seeds), rng.integers(0, X, n_seeds)], axis=1)
    zz, yy, xx = np.meshgrid(np.arange(Z), np.arange(Y), np.arange(X), indexing="ij")
    best = np.full(shape, np.inf)
    lab = np.zeros(shape, dtype=np.int64)
    for i, (pz, py, px) in enumerate(pts):
        d = (zz - pz) ** 2 + 1.7 * (yy - py) ** 2 + 0.6 * (xx - px) ** 2
        m = d < best
        best[m] = d[m]
        lab[m] = 1000 + 7 * i
    lab[best > (0.35 * max(shape)) ** 2] = 0                      # background far from every seed
    for k in range(6):                                           # tiny segments (below any size threshold)
        z, y, x = rng.integers(1, Z - 2), rng.integers(1, Y - 2), rng.integers(1, X - 2)
        lab[z:z + 2, y:y + 2, x:x + 2] = 50 + k
    return lab


# name, shape, seed, kind, r, sigma, thd, buffer, offset, seg_dilate, seg_sz_thd, seg_force
VOXEL2OBJ_SEG_CASES = [
    ("seg_blobs_56_r6_d3",     (56, 56, 56), 21, "blobs",   6, 1.5, 0, 2, (0, 0, 0),    3,    None, None),
    ("seg_uniform_48_r7_d8",   (48, 52, 60), 22, "uniform", 7, 2.0, 0, 0, (5, 6, 7),    8,    None, None),
    ("seg_uniform_48_r5_none", (48, 48, 48), 23, "uniform", 5, 1.0, 0, 0, (0, 0, 0),    None, None, None),
    ("seg_uniform_48_r6_f3",   (48, 48, 48), 24, "uniform", 6, 1.5, 0, 3, (0, 0, 0),    2,    None, 3),
    ("seg_blobs_64_r8_sz",     (64, 64, 64), 25, "blobs",   8, 2.0, 0, 4, (0, 0, 0),    4,    20,   None),
    ("seg_uniform_40_r27_d8",  (70, 64, 66), 26, "uniform", 27, 5.0, 0, 5, (0, 0, 0),   8,    None, 4),
]
