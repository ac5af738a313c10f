#!/usr/bin/env python3
"""A/B timing of the diagonal encoder + image build (MatMult4StreamPreprocess of a resident cache) for builds of the library given
by SFG_B200_LIB:  python profiles/microbench/ab_encode.py PN14QP438 8192 65536   ->  one line: diagonals, best-of-5 seconds, us / diagonal."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from bench import CKKS  # noqa: E402
from sfgwas_b200 import CryptoParams, GenoFileStream, MatMult4StreamPreprocess  # noqa: E402

pname, nrows, ncols = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
P = CKKS[pname]
cps = CryptoParams(P["logN"], P["Q"], P["P"], P["scale"], device=0)
X = np.random.default_rng(1).integers(0, 3, (nrows, ncols), dtype=np.int8)
gfs = GenoFileStream.from_matrix(cps, X)
best, npoly = 1e9, 0
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cache = MatMult4StreamPreprocess(cps, gfs, 5)
    torch.cuda.synchronize()
    best = min(best, time.perf_counter() - t0)
    npoly = cache.num_polys if hasattr(cache, "num_polys") else 0
    del cache
print("%s %s %dx%d: %d diagonals, %.4f s, %.3f us per diagonal" % (os.environ.get("SFG_B200_LIB", "in-tree"), pname, nrows, ncols, npoly, best,
                                                                   1e6 * best / max(1, npoly)))
