"""the local arithmetic of the collective bootstrap (mpc/mhe.go:262-341 -> dckks.RefreshProtocol GenShares / Decrypt / Recode /
Recrypt) on the GPU against the big-integer restatement in oracle/refresh.py, bit for bit, and by meaning (a 3-party refresh of a level-1
ciphertext decrypts to the same message at the top level)."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(params, level, nct, nparties, seed):
    from oracle.oracle import Oracle
    from sfgwas_b200 import CryptoParams

    o = Oracle.from_params(params)
    cps = CryptoParams(params["logN"], params["Q"], params["P"], params["scale"])
    rng, pr = np.random.default_rng(seed), random.Random(seed)
    sks = [o.keygen_secret(10 + p) for p in range(nparties)]
    sk = np.zeros_like(sks[0])
    for l in range(o.nQ + o.nP):  # aggregate key = sum of the shards (crypto/crypto.go:166-169)
        q = int((o.Q + o.P)[l])
        acc = np.zeros(o.N, dtype=object)
        for s_ in sks:
            acc = acc + s_[l].astype(object)
        sk[l] = (acc % q).astype(np.uint64)
    vals = rng.normal(size=(nct, o.slots))
    cts = np.stack([o.encrypt_vector(sk, vals[t], level, seed=100 + t) for t in range(nct)])  # [nct][2][level+1][N]
    Ql = 1
    for q in o.Q[: level + 1]:
        Ql *= int(q)
    B = Ql // (2 * nparties)
    crp = np.stack([np.stack([rng.integers(0, int(q), o.N, dtype=np.uint64) for q in o.Q]) for _ in range(nct)])
    draws = []
    for p in range(nparties):
        mask = [[pr.randrange(B) - B // 2 for _ in range(o.N)] for _ in range(nct)]
        e0 = np.array([[pr.randrange(-19, 20) for _ in range(o.N)] for _ in range(nct)], dtype=np.int64)
        e1 = np.array([[pr.randrange(-19, 20) for _ in range(o.N)] for _ in range(nct)], dtype=np.int64)
        draws.append((mask, e0, e1))
    return o, cps, sks, sk, vals, cts, crp, draws


def _mont(o, sk):
    out = sk.copy()
    for l in range(o.nQ):
        q = int(o.Q[l])
        out[l] = ((sk[l].astype(object) << 64) % q).astype(np.uint64)
    return out


@pytest.mark.parametrize("shape,logN,level,nct", [("pn13", 8, 1, 3), ("pn13", 8, 4, 2), ("pn14", 9, 2, 2)])
def test_refresh_bit_exact_and_meaning(shape, logN, level, nct):
    from oracle import refresh
    from oracle.oracle import small_params
    from sfgwas_b200 import RefreshFinish, RefreshGenShares

    nparties = 3
    o, cps, sks, sk, vals, cts, crp, draws = _setup(small_params(logN, shape), level, nct, nparties, seed=logN * 10 + level)
    agg0 = np.zeros((nct, level + 1, o.N), dtype=object)
    agg1 = np.zeros((nct, o.nQ, o.N), dtype=object)
    in_scale = o.scale * 1.37  # not a power of two: mask and plaintext are rescaled by floor(targetScale) / floor(ct.Scale)
    for p in range(nparties):
        mask, e0, e1 = draws[p]
        h0, h1 = RefreshGenShares(cps, level, cts[:, 1], _mont(o, sks[p]), crp, mask, e0, e1, in_scale, o.scale)
        for t in range(nct):
            w0, w1 = refresh.gen_shares(o, level, sks[p], cts[t, 1], crp[t], mask[t], e0[t], e1[t], in_scale, o.scale)
            assert (h0[t] == w0).all() and (h1[t] == w1).all(), "GenShares differs from the oracle (party %d, ct %d)" % (p, t)
        agg0 = agg0 + h0.astype(object)
        agg1 = agg1 + h1.astype(object)
    for l in range(o.nQ):  # AggregateRefreshShare: mod-q sums (mpc/aggregate.go:291-336)
        q = int(o.Q[l])
        if l <= level:
            agg0[:, l] %= q
        agg1[:, l] %= q
    agg0, agg1 = agg0.astype(np.uint64), agg1.astype(np.uint64)
    out = RefreshFinish(cps, level, cts[:, 0], in_scale, agg0, agg1, crp, out_scale=o.scale)
    for t in range(nct):
        want = refresh.finish(o, level, cts[t, 0], in_scale, o.scale, agg0[t], agg1[t], crp[t])
        assert (out[t] == want).all(), "Decrypt/Recode/Recrypt differs from the oracle (ct %d)" % t
        got = o.decrypt_vector(sk, out[t], o.scale).real * 1.37  # the message was encoded at scale, declared as 1.37 * scale
        assert np.abs(got - vals[t]).max() < 1e-4
    cps.close()


def test_refresh_real_parameter_set():
    """PN13QP218, a MatMult output's level and scale: level 4, scale ~ 2^60 (> 2^53: exercised the 128-bit divisor of the recode)."""
    from oracle import refresh
    from oracle.oracle import PARAMS
    from sfgwas_b200 import RefreshFinish, RefreshGenShares

    level, nct, nparties = 4, 1, 2
    o, cps, sks, sk, vals, cts, crp, draws = _setup(PARAMS["PN13QP218"], level, nct, nparties, seed=77)
    agg0 = np.zeros((nct, level + 1, o.N), dtype=object)
    agg1 = np.zeros((nct, o.nQ, o.N), dtype=object)
    in_scale = o.scale * o.scale  # A.scale * params.Scale (gwas/matmult.go:1045): 2^60, the ciphertext "holds" vals / 2^30
    for p in range(nparties):
        mask, e0, e1 = draws[p]
        h0, h1 = RefreshGenShares(cps, level, cts[:, 1], _mont(o, sks[p]), crp, mask, e0, e1, in_scale, o.scale)
        w0, w1 = refresh.gen_shares(o, level, sks[p], cts[0, 1], crp[0], mask[0], e0[0], e1[0], in_scale, o.scale)
        assert (h0[0] == w0).all() and (h1[0] == w1).all()
        agg0 = agg0 + h0.astype(object)
        agg1 = agg1 + h1.astype(object)
    for l in range(o.nQ):
        q = int(o.Q[l])
        if l <= level:
            agg0[:, l] %= q
        agg1[:, l] %= q
    agg0, agg1 = agg0.astype(np.uint64), agg1.astype(np.uint64)
    out = RefreshFinish(cps, level, cts[:, 0], in_scale, agg0, agg1, crp)
    want = refresh.finish(o, level, cts[0, 0], in_scale, o.scale, agg0[0], agg1[0], crp[0])
    assert (out[0] == want).all()
    cps.close()
