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


# ---------------------------------------------------------------------------------------------------------------
# Content classes for the BC7 / BC6H tolerance corpus (tests/golden/make_golden.py, tests, bench parity leg).
# Every class is seeded and numpy-only, so the GPU box regenerates the inputs and only the reference encoder's
# per-block errors need to be committed as fixtures.
LDR_CLASSES = ("noise", "gradient", "voronoi", "cluster2", "cluster3", "cluster4", "alpha_uncorr", "chan_uncorr",
               "dark", "text", "c2", "photo", "alpha_photo", "cutout")


def _palette_field(rng, h, w, n, cell, jitter=0.0):
    """per-cell palettes of n random colours; returns (h, w, n, 3)"""
    ch, cw = (h + cell - 1) // cell, (w + cell - 1) // cell
    pal = rng.random((ch, cw, n, 3)).astype(np.float32)
    pal = np.kron(pal, np.ones((cell, cell, 1, 1), np.float32))[:h, :w]
    if jitter:
        pal = pal + rng.normal(0, jitter, (h, w, n, 3)).astype(np.float32)
    return pal


def content_ldr(kind, w=256, h=256, seed=1):
    """RGBA32F image in [0,1] of content class `kind` (LDR_CLASSES)."""
    rng = np.random.default_rng(seed * 7919 + sum(map(ord, kind)))
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    one = np.ones((h, w, 1), np.float32)
    if kind == "noise":
        img = np.concatenate([rng.random((h, w, 3)).astype(np.float32), one], -1)
    elif kind == "gradient":
        r = 0.5 + 0.5 * np.sin(x * 0.011 + y * 0.007)
        g = (x / max(w - 1, 1)) * 0.8 + 0.1 * np.cos(y * 0.02)
        b = (y / max(h - 1, 1)) * (x / max(w - 1, 1))
        img = np.stack([r, g, b, np.ones_like(r)], -1)
    elif kind == "voronoi":
        n = max(8, (w * h) // 600)
        pts = rng.random((n, 2)).astype(np.float32) * np.array([w, h], np.float32)
        cols = rng.random((n, 3)).astype(np.float32)
        d = (x[..., None] - pts[:, 0]) ** 2 + (y[..., None] - pts[:, 1]) ** 2
        img = np.concatenate([cols[np.argmin(d, -1)], one], -1)
        img[..., :3] += rng.normal(0, 0.004, (h, w, 3)).astype(np.float32)
    elif kind in ("cluster2", "cluster3", "cluster4"):
        n = int(kind[-1])
        pal = _palette_field(rng, h, w, n, 16)
        # periodic assignment: stripes / checker patterns whose period does not align with the 4x4 grid
        per = rng.integers(2, 6)
        sel = ((x * 0.9 + y * 0.6) // per).astype(np.int64) % n
        sel2 = rng.integers(0, n, (h, w))
        mix = (rng.random((h, w)) < 0.15)
        sel = np.where(mix, sel2, sel)
        rgb = np.take_along_axis(pal, sel[..., None, None].repeat(3, -1), 2)[:, :, 0]
        rgb = rgb + rng.normal(0, 0.01, (h, w, 3)).astype(np.float32)
        img = np.concatenate([rgb, one], -1)
    elif kind == "alpha_uncorr":
        base = photo_rgba32f(w, h, seed + 101)
        img = base.copy()
        img[..., 3] = rng.random((h, w)).astype(np.float32)
    elif kind == "chan_uncorr":
        base = photo_rgba32f(w, h, seed + 102)
        img = base.copy()
        img[..., 2] = rng.random((h, w)).astype(np.float32)      # blue independent of red / green, block stays opaque
    elif kind == "dark":
        base = photo_rgba32f(w, h, seed + 103)
        img = base.copy()
        img[..., :3] *= np.float32(0.06)
    elif kind == "text":
        bg = 0.75 + 0.2 * np.sin(x * 0.02)[..., None] * np.array([1.0, 0.9, 0.7], np.float32)
        glyph = rng.random(((h + 1) // 2, (w + 2) // 3)) < 0.45
        glyph = np.kron(glyph, np.ones((2, 3), bool))[:h, :w]
        glyph &= ((y.astype(np.int64) % 12) < 9)
        ink = np.array([0.05, 0.05, 0.1], np.float32)
        rgb = np.where(glyph[..., None], ink, bg).astype(np.float32)
        # anti-aliased edges
        rgb = 0.5 * rgb + 0.25 * np.roll(rgb, 1, 1) + 0.25 * np.roll(rgb, -1, 1)
        img = np.concatenate([rgb, one], -1)
    elif kind == "c2":
        img = c2_rgba32f(w, h, seed)
    elif kind == "photo":
        img = photo_rgba32f(w, h, seed + 104)
    elif kind == "alpha_photo":
        img = photo_rgba32f(w, h, seed + 105, alpha=True)
    elif kind == "cutout":
        base = photo_rgba32f(w, h, seed + 106)
        img = base.copy()
        mask = (np.sin(x * 0.13) * np.cos(y * 0.09) + 0.3 * rng.normal(size=(h, w))) > 0.1
        img[..., 3] = mask.astype(np.float32)
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(np.clip(img, 0, 1).astype(np.float32))


HDR_CLASSES = ("c3", "smooth", "edges", "hdr_noise", "sky", "flat_hdr", "ldr_in_hdr", "bright_spots")
HDR_SIGNED_CLASSES = ("sincos", "normal_map", "noise_pm10", "gauss3000", "signed_smooth", "cross_edges")


def content_hdr(kind, w=256, h=256, seed=1):
    """RGBA32F HDR image of content class `kind` (HDR_CLASSES: non-negative, for BC6H_UF16; HDR_SIGNED_CLASSES: values of
    both signs, for BC6H_SF16).  |v| <= 65504, alpha 1."""
    rng = np.random.default_rng(seed * 104729 + sum(map(ord, kind)))
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    smooth = np.stack([np.exp2(4 * np.sin(x * 0.05) + 2 * np.cos(y * 0.03)), np.exp2(3 * np.cos(x * 0.02 + y * 0.04)),
                       np.exp2(2 * np.sin(y * 0.06))], -1).astype(np.float32)
    if kind == "c3":
        return np.ascontiguousarray(c3_rgba16f(w, h, seed=seed).astype(np.float32))
    if kind == "smooth":
        rgb = smooth * (1 + 0.02 * rng.normal(size=smooth.shape))
    elif kind == "edges":
        rgb = np.where(((x // 4 + y // 4) % 3 == 0)[..., None] & ((x % 4) < 2)[..., None], np.float32(50.0), smooth)
    elif kind == "hdr_noise":
        rgb = np.exp2(rng.uniform(-4, 8, (h, w, 3)))
    elif kind == "sky":
        sun = 4000.0 * np.exp(-(((x - w * 0.6) ** 2 + (y - h * 0.3) ** 2) / (2 * (0.05 * w) ** 2)))
        base = np.stack([0.2 + 0.3 * y / h, 0.4 + 0.3 * y / h, 0.9 - 0.2 * y / h], -1)
        rgb = base + sun[..., None] * np.array([1.0, 0.9, 0.7]) + 0.01 * rng.random((h, w, 3))
    elif kind == "flat_hdr":
        cell = np.exp2(rng.uniform(-3, 10, ((h + 7) // 8, (w + 7) // 8, 3)))
        rgb = np.kron(cell, np.ones((8, 8, 1)))[:h, :w]
    elif kind == "ldr_in_hdr":
        rgb = photo_rgba32f(w, h, seed + 201)[..., :3]
    elif kind == "bright_spots":
        rgb = photo_rgba32f(w, h, seed + 202)[..., :3] * 0.5
        spots = rng.random((h, w)) < 0.01
        rgb = np.where(spots[..., None], rng.uniform(100, 60000, (h, w, 3)), rgb)
    elif kind == "sincos":
        rgb = np.stack([np.sin(x * 0.11 + y * 0.05), np.cos(x * 0.07 - y * 0.13), np.sin(x * 0.19) * np.cos(y * 0.17)], -1)
    elif kind == "normal_map":
        n = rng.normal(size=(h, w, 3))
        n = n + 2.0 * np.stack([np.sin(x * 0.05), np.cos(y * 0.05), np.ones_like(x)], -1)
        rgb = n / np.linalg.norm(n, axis=-1, keepdims=True)
        rgb[..., 2] -= 0.5
    elif kind == "noise_pm10":
        rgb = rng.uniform(-10, 10, (h, w, 3))
    elif kind == "gauss3000":
        rgb = rng.normal(0, 3000, (h, w, 3))
    elif kind == "signed_smooth":
        rgb = smooth - np.float32(3.0)
    elif kind == "cross_edges":
        rgb = np.where(((x // 4 + y // 4) % 2 == 0)[..., None] & ((x % 4) < 2)[..., None], np.float32(-20.0), smooth * 0.1)
    else:
        raise ValueError(kind)
    rgb = np.clip(np.asarray(rgb, np.float32), -65504, 65504)
    return np.ascontiguousarray(np.concatenate([rgb, np.ones((h, w, 1), np.float32)], -1).astype(np.float32))
