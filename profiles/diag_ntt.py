import sys, json, random
sys.path.insert(0, '.')
import numpy as np
from sfgwas_b200 import CryptoParams
from oracle.oracle import Oracle, gen_primes
for logN in [6, 8, 9, 10, 12, 13, 14]:
    qs = gen_primes(logN, 45, 1) + gen_primes(logN, 30, 4) + gen_primes(logN, 31, 2)
    ps = gen_primes(logN, 55, 1)
    o = Oracle(logN, qs, ps, 2.0**30)
    cps = CryptoParams(logN, qs, ps, 2.0**30)
    rng = np.random.default_rng(logN)
    for idx, q in enumerate(qs + ps):
        a = rng.integers(0, q, 1 << logN, dtype=np.uint64)
        want = o.ntt(idx, a)
        got = cps.NTT(a[None], [idx])[0]
        back = cps.NTT(want[None], [idx], inverse=True)[0]
        print(logN, hex(q), 'kind', 1 if q < 2**30 else (2 if q < 2**31 else 0), 'fwd', bool((got == want).all()), 'inv', bool((back == a).all()),
              'nbad', int((got != want).sum()))
    cps.close()
