import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import *
import torch
from dealiiscale_b200 import capi
capi.load()
lib = C.CDLL(os.path.join(ROOT, "dealiiscale_b200", "_build", "libmsdrivers.so"))
lib.msdrv_bio_create.restype = C.c_void_p
prm, cells, limit = sys.argv[1], (int(sys.argv[2]), int(sys.argv[3])), int(sys.argv[4])
o = OracleBio(prm, *cells, threads=8)
nc = o.run(limit)
h = C.c_void_p(lib.msdrv_bio_create(os.path.join(INPUT, prm).encode(), *cells))
got = lib.msdrv_bio_run(h, limit)
print("cycles oracle", nc, "gpu", got)
N, n = C.c_int(), C.c_int(); lib.msdrv_bio_sizes(h, C.byref(N), C.byref(n)); N, n = N.value, n.value
for c in range(min(nc, got)):
    ref = o.history(c)
    sc = np.zeros(6); su, sw, cu, cw = (np.zeros(N) for _ in range(4)); micro, its = np.zeros((N, n)), np.zeros(N, dtype=np.int32)
    lib.msdrv_bio_history(h, c, dptr(sc), dptr(su), dptr(sw), dptr(cu), dptr(cw), dptr(micro), iptr(its))
    print(c, "su %.1e sw %.1e cu %.1e cw %.1e micro %.1e | mL2 %.6e/%.6e ML2 %.6e/%.6e res %.3e/%.3e %.3e/%.3e its %d/%d" % (
        rel_l2(su, ref["sol_u"]), rel_l2(sw, ref["sol_w"]), rel_l2(cu, ref["cu"]), rel_l2(cw, ref["cw"]), rel_l2(micro, ref["micro"]),
        sc[0], ref["micro_l2"], sc[2], ref["macro_l2"], sc[4], ref["macro_residual"], sc[5], ref["micro_residual"], its.max(), ref["micro_iters"].max()))
