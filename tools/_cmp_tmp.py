import sys, numpy as np
sys.path.insert(0,'/root/repo')
from directxtex_b200 import capi, synth
from tests import oracle_lib, tolerance
capi.lib.dxb200_init(0)
emu = oracle_lib.load_emul()
for kind in synth.LDR_CLASSES:
    img = synth.content_ldr(kind, 256, 256, 1)
    got = capi.compress(img, 256, 256, 2, 98, 0)
    he, em = emu.compress(img, 256, 256, 2, 98, 0)
    d = np.nonzero((got.reshape(-1,16) != em.reshape(-1,16)).any(1))[0]
    print(kind, len(d), d[:6])
