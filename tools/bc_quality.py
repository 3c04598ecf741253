#!/usr/bin/env python
"""Development tool (test infrastructure): BC7 / BC6H quality of the host lock-step emulator of the CUDA encoder
(bit-identical to the device, checked by the GPU tests) against the oracle (the unmodified reference encoder),
per content class.  Oracle encodes are cached under /tmp/dxb_oracle_cache (they are slow and deterministic).

    python tools/bc_quality.py bc7 [--size 128] [--flags 0] [class ...]
    python tools/bc_quality.py bc6h [--size 128] [class ...]
"""
import argparse
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from directxtex_b200 import synth  # noqa: E402
from tests import oracle_lib  # noqa: E402

CACHE = "/tmp/dxb_oracle_cache"


def ref_compress_cached(ref, img, w, h, sf, df, flags):
    os.makedirs(CACHE, exist_ok=True)
    key = hashlib.sha1(img.tobytes() + bytes("%d_%d_%d_%d_%d" % (w, h, sf, df, flags), "ascii")).hexdigest()
    p = os.path.join(CACHE, key + ".npy")
    if os.path.exists(p):
        return np.load(p)
    hr, out = ref.compress(img, w, h, sf, df, flags)
    assert hr == 0
    np.save(p, out)
    return out


def block_sse(dec255, src255, w, h):
    """per-block sum of squared errors, shape (h/4, w/4)"""
    d = (dec255 - src255) ** 2
    return d.reshape(h // 4, 4, w // 4, 4, -1).sum((1, 3, 4))


def bc7_modes(blocks):
    b0 = blocks.reshape(-1, 16)[:, 0]
    modes = np.full(b0.shape, 8)
    for m in range(7, -1, -1):
        modes[(b0 & ((1 << (m + 1)) - 1)) == (1 << m)] = m
    return np.bincount(modes, minlength=9)


def run_bc7(args):
    ref, emu = oracle_lib.load_ref(), oracle_lib.load_emul()
    n = args.size
    kinds = args.classes or synth.LDR_CLASSES
    worst = 0
    print("%-13s %9s %9s %7s %7s %8s %8s  modes(emul) | modes(ref)" % ("class", "mse_emul", "mse_ref", "ratio", "dB", ">2x+16", ">1.5x+8"))
    for k in kinds:
        img = synth.content_ldr(k, n, n, args.seed)
        t0 = time.time()
        rb = ref_compress_cached(ref, img, n, n, 2, 98, args.flags)
        t1 = time.time()
        hr, eb = emu.compress(img, n, n, 2, 98, args.flags)
        assert hr == 0
        src = oracle_lib.bc7_ldr(img).astype(np.float64)
        de = ref.decode_blocks(98, eb, n, n).astype(np.float64) * 255.0
        dr = ref.decode_blocks(98, rb, n, n).astype(np.float64) * 255.0
        se, sr = block_sse(de, src, n, n), block_sse(dr, src, n, n)
        me, mr = se.sum() / (n * n * 4), sr.sum() / (n * n * 4)
        ratio = me / max(mr, 1e-9)
        bad2 = float((se > 2 * sr + 16).mean())
        bad15 = float((se > 1.5 * sr + 8).mean())
        worst = max(worst, ratio)
        print("%-13s %9.4f %9.4f %7.4f %+7.3f %7.2f%% %7.2f%%  %s | %s" % (
            k, me, mr, ratio, 10 * np.log10(max(mr, 1e-9) / max(me, 1e-9)), 100 * bad2, 100 * bad15,
            " ".join(str(v) for v in bc7_modes(eb)[:8]), " ".join(str(v) for v in bc7_modes(rb)[:8])), flush=True)
    print("worst ratio %.4f" % worst)


def half_to_float(bits):
    return bits.astype(np.uint16).view(np.float16).astype(np.float64)


def run_bc6h(args):
    ref, emu = oracle_lib.load_ref(), oracle_lib.load_emul()
    n = args.size
    sets = [(k, 95) for k in synth.HDR_CLASSES] + [(k, 96) for k in synth.HDR_SIGNED_CLASSES] + [(k, 96) for k in ("c3", "smooth")]
    if args.classes:
        sets = [s for s in sets if s[0] in args.classes]
    print("%-14s %3s %11s %11s %7s | %11s %11s %7s | %9s %9s  %7s" % ("class", "fmt", "int_emul", "int_ref", "ratio", "fmse_emul", "fmse_ref", "fratio", "fmax_emul", "fmax_ref", ">2x"))
    for k, fmt in sets:
        img = synth.content_hdr(k, n, n, args.seed)
        signed = fmt == 96
        rb = ref_compress_cached(ref, img, n, n, 2, fmt, 0)
        hr, eb = emu.compress(img, n, n, 2, fmt, 0)
        assert hr == 0
        de = ref.decode_blocks(fmt, eb, n, n)
        dr = ref.decode_blocks(fmt, rb, n, n)
        s16 = np.clip(img[..., :3], -65504 if signed else 0, 65504).astype(np.float16).astype(np.float64)
        ie, ir = oracle_lib.bc6h_int_mse(de, img, signed), oracle_lib.bc6h_int_mse(dr, img, signed)
        fe = ((de[..., :3].astype(np.float64) - s16) ** 2)
        fr = ((dr[..., :3].astype(np.float64) - s16) ** 2)
        a = oracle_lib.bc6h_to_int(de[..., :3], signed).astype(np.float64)
        b = oracle_lib.bc6h_to_int(dr[..., :3], signed).astype(np.float64)
        s = oracle_lib.bc6h_to_int(np.clip(img[..., :3], -65504 if signed else 0, 65504), signed).astype(np.float64)
        se, sr = block_sse(a, s, n, n), block_sse(b, s, n, n)
        bad2 = float((se > 2 * sr + 48 * 16).mean())
        print("%-14s %3d %11.4g %11.4g %7.4f | %11.4g %11.4g %7.3f | %9.4g %9.4g  %6.2f%%" % (
            k, fmt, ie, ir, ie / max(ir, 1e-9), fe.mean(), fr.mean(), fe.mean() / max(fr.mean(), 1e-30),
            np.sqrt(fe.max()), np.sqrt(fr.max()), 100 * bad2), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("codec", choices=["bc7", "bc6h"])
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--flags", type=lambda s: int(s, 0), default=0)
    ap.add_argument("classes", nargs="*")
    a = ap.parse_intermixed_args()
    (run_bc7 if a.codec == "bc7" else run_bc6h)(a)
