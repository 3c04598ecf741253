"""Seeded synthetic inputs for the BASELINE.json configurations.  numpy only;
used by bench.py, tests/ and __graft_entry__.smoke().  No NaN/Inf, HDR values <= 65504."""
import numpy as np

SEED = 0xD1EC7E0


def _tile(a, ty, tx):
    return np.kron(a, np.ones((ty, tx) + (1,) * (a.ndim - 2), a.dtype))


def c2_rgba32f(w=4096, h=4096, seed=SEED):
    """C2: RGBA32F in [0,1]: low-frequency sum of 4 sinusoids per channel + uniform noise (amp 0.02) +
    12.5% of blocks hard two-colour edges; alpha 1.0 on 75% of 64x64 tiles, smooth ramp on 25%."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.empty((h, w, 4), np.float32)
    for ch in range(3):
        v = np.zeros((h, w), np.float32)
        for _ in range(4):
            fx, fy = rng.uniform(0.002, 0.05, 2).astype(np.float32)
            ph = np.float32(rng.uniform(0, 6.28))
            amp = np.float32(rng.uniform(0.08, 0.25))
            v += np.sin(x * (fx * np.float32(6.28)) + y * (fy * np.float32(6.28)) + ph) * amp
        img[..., ch] = np.float32(0.5) + v
    img[..., :3] += rng.uniform(-0.02, 0.02, (h, w, 3)).astype(np.float32)
    bh, bw = (h + 3) // 4, (w + 3) // 4
    edge = rng.random((bh, bw)) < 0.125
    ca = rng.random((bh, bw, 3)).astype(np.float32)
    cb = rng.random((bh, bw, 3)).astype(np.float32)
    ang = rng.uniform(0, np.pi, (bh, bw)).astype(np.float32)
    off = rng.uniform(-1, 1, (bh, bw)).astype(np.float32)
    up = lambda a: _tile(a, 4, 4)[:h, :w]
    side = ((x % 4 - 1.5) * np.cos(up(ang)) + (y % 4 - 1.5) * np.sin(up(ang))) > up(off)
    eb = up(edge.astype(np.float32)) > 0
    img[..., :3] = np.where(eb[..., None], np.where(side[..., None], up(ca), up(cb)), img[..., :3])
    th, tw = (h + 63) // 64, (w + 63) // 64
    tile = (rng.random((th, tw)) < 0.25).astype(np.float32)
    tm = _tile(tile, 64, 64)[:h, :w] > 0
    ramp = np.float32(0.5) + np.float32(0.5) * np.sin(x * np.float32(0.05) + y * np.float32(0.031))
    img[..., 3] = np.where(tm, ramp, np.float32(1.0))
    return np.ascontiguousarray(np.clip(img, 0, 1).astype(np.float32))


def c1_rgba8(w=256, h=256, seed=SEED):
    """C1: RGBA8 in 8x8 tiles of (a) uniform random bytes, (b) 2-colour noise, (c) smooth gradients;
    alpha 255 except one tile column with random alpha (exercises the BC1 colour key)."""
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w, 4), np.uint8)
    ty, tx = (h + 7) // 8, (w + 7) // 8
    kind = rng.integers(0, 3, (ty, tx))
    y, x = np.mgrid[0:h, 0:w]
    rnd = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    c0 = _tile(rng.integers(0, 256, (ty, tx, 3), dtype=np.uint8), 8, 8)[:h, :w]
    c1 = _tile(rng.integers(0, 256, (ty, tx, 3), dtype=np.uint8), 8, 8)[:h, :w]
    two = np.where(rng.random((h, w, 1)) < 0.5, c0, c1)
    t = (((x % 8) + (y % 8)) / 14.0)[..., None]
    grad = (c0 * (1 - t) + c1 * t).astype(np.uint8)
    k = _tile(kind.astype(np.uint8), 8, 8)[:h, :w][..., None]
    img[..., :3] = np.where(k == 0, rnd, np.where(k == 1, two, grad))
    img[..., 3] = 255
    col = (tx // 2) * 8
    img[:, col:col + 8, 3] = rng.integers(0, 256, (h, min(8, w - col)), dtype=np.uint8)
    return np.ascontiguousarray(img)


def c3_rgba16f(w=2048, h=2048, seed=SEED):
    """C3: HDR exp2(uniform(-6,6)) modulated by a smooth field, clamped to [0,65504], alpha 1."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    field = 0.6 + 0.4 * np.sin(x * 0.01) * np.cos(y * 0.013)
    v = np.exp2(rng.uniform(-6, 6, (h, w, 3))).astype(np.float32) * field[..., None]
    img = np.concatenate([np.clip(v, 0, 65504), np.ones((h, w, 1), np.float32)], -1)
    return np.ascontiguousarray(img.astype(np.float16))


def c5_r8(w=8192, h=8192, seed=SEED):
    """C5: R8 smooth field + noise with 1% texels forced to 0 and 1% to 255 (BC4 4- vs 6-interp branch)."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    v = 128 + 90 * np.sin(x * 0.003) * np.cos(y * 0.004) + rng.normal(0, 6, (h, w)).astype(np.float32)
    img = np.clip(v, 1, 254).astype(np.uint8)
    r = rng.random((h, w))
    img[r < 0.01] = 0
    img[r > 0.99] = 255
    return np.ascontiguousarray(img)


def photo_rgba32f(w, h, seed, alpha=False):
    """random-walk 'photo-like' texture in [0,1] (tests)."""
    rng = np.random.default_rng(seed)
    nch = 4 if alpha else 3
    v = np.cumsum(rng.normal(0, 0.02, (h, w, nch)), 1) + np.cumsum(rng.normal(0, 0.02, (h, w, nch)), 0)
    v = (v - v.min((0, 1))) / (v.max((0, 1)) - v.min((0, 1)) + 1e-9)
    v += rng.normal(0, 0.01, v.shape)
    if not alpha:
        v = np.concatenate([v, np.ones((h, w, 1))], -1)
    return np.ascontiguousarray(np.clip(v, 0, 1).astype(np.float32))
