#!/usr/bin/env python
"""Generates tests/golden/golden_bc67_v2.npz: the tolerance anchors of the two non-bit-exact codecs (BC7, BC6H).

For every content class of directxtex_b200.synth (LDR_CLASSES -> BC7; HDR_CLASSES -> BC6H_UF16; HDR_SIGNED_CLASSES and two
non-negative classes -> BC6H_SF16) at 256x256 the UNMODIFIED reference encoder (oracle/_ref/libdxtex_ref.so) compresses the
image and the reference decoder decodes it; stored per class:
    sha1 of the input bytes (the tests regenerate the input from the seeded numpy generators and check it),
    per-block error of the reference: BC7  : sum of squared 8-bit differences over RGBA          (64 x 64 float32)
                                      BC6H : the same over the half bit patterns of RGB (the reference encoder's own metric,
                                             BC6HBC7.cpp:1167-1173) and over the decoded float values, and the largest
                                             absolute float error of the image.
Run in the build container only (needs /root/reference to build the oracle):   python tests/golden/make_golden_bc67.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from directxtex_b200 import synth  # noqa: E402
from tests import oracle_lib, tolerance  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_bc67_v2.npz")


def main():
    ref = oracle_lib.load_ref()
    data = {}
    for kind, flags in tolerance.bc7_cases():
        img = synth.content_ldr(kind, tolerance.SIZE, tolerance.SIZE, tolerance.SEED)
        hr, blocks = ref.compress(img, tolerance.SIZE, tolerance.SIZE, 2, 98, flags)
        assert hr == 0
        key = tolerance.bc7_key(kind, flags)
        data[key + "_sha1"] = np.frombuffer(hashlib.sha1(img.tobytes()).digest(), np.uint8)
        data[key + "_sse"] = tolerance.bc7_block_sse(ref, blocks, img).astype(np.float32)
        print(key, "ref mse %.4f" % (data[key + "_sse"].sum() / (tolerance.SIZE ** 2 * 4)), flush=True)
    for kind, fmt in tolerance.bc6h_cases():
        img = synth.content_hdr(kind, tolerance.SIZE, tolerance.SIZE, tolerance.SEED)
        hr, blocks = ref.compress(img, tolerance.SIZE, tolerance.SIZE, 2, fmt, 0)
        assert hr == 0
        key = tolerance.bc6h_key(kind, fmt)
        isse, fsse, fmax = tolerance.bc6h_block_errors(ref, blocks, img, fmt)
        data[key + "_sha1"] = np.frombuffer(hashlib.sha1(img.tobytes()).digest(), np.uint8)
        data[key + "_isse"] = isse.astype(np.float64)
        data[key + "_fsse"] = fsse.astype(np.float64)
        data[key + "_fmax"] = np.array([fmax], np.float64)
        print(key, "ref int mse %.5g float mse %.5g max %.5g" % (isse.sum() / (tolerance.SIZE ** 2 * 3), fsse.sum() / (tolerance.SIZE ** 2 * 3), fmax), flush=True)
    np.savez_compressed(OUT, **data)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
