import numpy as np, sys
sys.path.insert(0, "/root/repo")
from oracle.oracle import Oracle, sweep_params
from sfgwas_b200 import CryptoParams
p = sweep_params(15)
o = Oracle.from_params(p); cps = CryptoParams(p["logN"], p["Q"], p["P"], p["scale"])
rng = np.random.default_rng(0)
mods = o.Q + o.P
for idx in ([0], [3, 4, 5], [18, 19, 20], list(range(21))):
    for G in (1, 2, 3):
        polys = np.stack([np.stack([rng.integers(0, mods[l], o.N, dtype=np.uint64) for l in idx]) for _ in range(G)])
        fw = cps.NTT(polys, idx)
        want = np.stack([np.stack([o.ntt(l, polys[g, k].copy()) for k, l in enumerate(idx)]) for g in range(G)])
        bk = cps.NTT(fw, idx, inverse=True)
        print(idx[:3], G, "fwd", bool((fw == want).all()), "inv", bool((bk == polys).all()))
