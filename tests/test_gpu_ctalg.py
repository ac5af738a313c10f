"""GPU parity tests of the callers' ciphertext algebra: crypto.CMult / CMultScalar / InnerSumAll / MaskTrunc /
Sub and the lazy-normalisation wrappers QXLazyNormStream / QXtLazyNormStream (gwas/matmult.go:27-116), CUDA through the C ABI
against the CPU oracle on the same seeded inputs -- BIT-EXACT -- plus the decrypted result against the plain formula
(CKKS tolerance 1e-3 absolute at scale 2^30 on O(1) inputs, observed ~1e-5)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class _Net:
    """Stand-in for mpcObj.Network.BootstrapMatAll: the collective bootstrap is network protocol (out of scope); the tests refresh
    a ciphertext by decrypt + re-encrypt with the oracle so that the wrappers' level budget matches the reference's."""

    def __init__(self, o, sk, wrap):
        self.o, self.sk, self.wrap = o, sk, wrap
        self.seed = 1000

    def BootstrapMatAll(self, cps, M):
        o = self.o
        out = []
        for row in M:
            r = []
            for ct in row:
                v, sc = (ct.value, ct.scale) if hasattr(ct, "value") else ct
                vals = o.decrypt_vector(self.sk, v, sc).real
                self.seed += 1
                r.append(self.wrap(o.encrypt_vector(self.sk, vals, o.nQ - 1, seed=self.seed), o.scale))
            out.append(r)
        return out


class _Mpc:
    def __init__(self, net, pid=1):
        self.Network, self.pid = net, pid

    def GetPid(self):
        return self.pid


@pytest.fixture(scope="module", params=["pn13", "pn14"])
def env(request):
    from oracle.oracle import Oracle, small_params
    from sfgwas_b200 import CryptoParams, SetRelinKey

    p = small_params(8 if request.param == "pn13" else 9, request.param)
    o = Oracle.from_params(p)
    cps = CryptoParams(p["logN"], p["Q"], p["P"], p["scale"])
    sk = o.keygen_secret(1)
    keys = o.gen_bsgs_keys(sk)
    for k in o.pow2_rotations():
        if k not in keys:
            keys[k] = o.gen_rotation_key(sk, k)
    cps.SetRotKeys(keys)
    rlk = o.gen_relin_key(sk)
    SetRelinKey(cps, rlk)
    return o, cps, sk, keys, rlk


def enc(o, sk, vals, level, seed):
    return o.encrypt_vector(sk, vals, level, seed=seed)


def test_mul_relin_rescale_bit_exact(env):
    from sfgwas_b200 import Ciphertext, CMult

    o, cps, sk, keys, rlk = env
    rng = np.random.default_rng(1)
    top = o.nQ - 1
    a = [rng.normal(size=o.slots) for _ in range(3)]
    b = [rng.normal(size=o.slots) for _ in range(3)]
    X = [enc(o, sk, a[i], top, 10 + i) for i in range(3)]
    Y = [enc(o, sk, b[i], top - 1, 20 + i) for i in range(3)]  # different level: MulRelin works at the minimum
    want = o.CMult([(x, o.scale) for x in X], [(y, o.scale) for y in Y], rlk)
    got = CMult(cps, [Ciphertext(x, o.scale) for x in X], [Ciphertext(y, o.scale) for y in Y])
    for g, w in zip(got, want):
        assert g.value.shape == w[0].shape == (2, top - 1, o.N)
        assert (g.value == w[0]).all()
        assert g.scale == w[1]
    err = np.abs(o.decrypt_vector(sk, got[0].value, got[0].scale).real - a[0] * b[0]).max()
    assert err < 1e-3, err
    # broadcast of a length-1 side (crypto/basics.go:390-415) and CMultScalar
    want = o.CMult([(X[0], o.scale)], [(y, o.scale) for y in Y], rlk)
    got = CMult(cps, [Ciphertext(X[0], o.scale)], [Ciphertext(y, o.scale) for y in Y])
    assert all((g.value == w[0]).all() for g, w in zip(got, want))


def test_rescale_matches_bignum_rounding(env):
    """ring.DivRoundByLastModulusNTT is round(x / q_L) with the centred remainder: checked on CRT-reconstructed integers."""
    o, cps, sk, keys, rlk = env
    rng = np.random.default_rng(2)
    lvl = o.nQ - 1
    ct = np.stack([np.stack([rng.integers(0, o.Q[l], o.N, dtype=np.uint64) for l in range(lvl + 1)]) for _ in range(2)])
    out = np.zeros((1, 2, lvl, o.N), dtype=np.uint64)
    cps._check(cps.L.sfg_ct_rescale(cps.h, lvl, ct.ctypes.data, 1, 1, out.ctypes.data), "sfg_ct_rescale")
    want1 = np.zeros((2, lvl, o.N), dtype=np.uint64)
    o.L.orc_rescale_once(o.ctx, lvl, ct.ctypes.data, want1.ctypes.data)
    assert (out[0] == want1).all()
    qL = o.Q[lvl]
    for comp in range(2):
        xc = np.stack([o.intt(l, ct[comp, l].copy()) for l in range(lvl + 1)])
        yc = np.stack([o.intt(l, out[0, comp, l].copy()) for l in range(lvl)])
        big, got = o.crt_center(xc), o.crt_center(yc)
        for x, g in zip(big[:64], got[:64]):
            r = x % qL
            if r > (qL - 1) // 2:
                r -= qL
            assert (x - r) // qL == g
    # two steps in one call == two calls
    out2 = np.zeros((1, 2, lvl - 1, o.N), dtype=np.uint64)
    cps._check(cps.L.sfg_ct_rescale(cps.h, lvl, ct.ctypes.data, 1, 2, out2.ctypes.data), "sfg_ct_rescale")
    out3 = np.zeros((1, 2, lvl - 1, o.N), dtype=np.uint64)
    cps._check(cps.L.sfg_ct_rescale(cps.h, lvl - 1, out.ctypes.data, 1, 1, out3.ctypes.data), "sfg_ct_rescale")
    assert (out2 == out3).all()


def test_inner_sum_all_and_mask_trunc(env):
    from sfgwas_b200 import Ciphertext, InnerSumAll, MaskTrunc

    o, cps, sk, keys, rlk = env
    rng = np.random.default_rng(3)
    top = o.nQ - 1
    vals = [rng.normal(size=o.slots) for _ in range(3)]
    X = [enc(o, sk, v, top, 30 + i) for i, v in enumerate(vals)]
    want = o.InnerSumAll([(x, o.scale) for x in X], keys)
    got = InnerSumAll(cps, [Ciphertext(x, o.scale) for x in X])
    assert (got.value == want[0]).all()
    dec = o.decrypt_vector(sk, got.value, got.scale).real
    assert np.abs(dec - sum(v.sum() for v in vals)).max() < 1e-2
    # MaskTrunc: device-encoded mask == the oracle's correctly rounded encoding; result bit-exact
    n_keep = o.slots // 3
    m = np.zeros(o.slots)
    m[:n_keep] = 1.0
    want = o.MaskTrunc((X[0], o.scale), n_keep)
    got = MaskTrunc(cps, Ciphertext(X[0], o.scale), n_keep)
    assert (got.value == want[0]).all() and got.scale == want[1]
    dec = o.decrypt_vector(sk, got.value, got.scale).real
    assert np.abs(dec[:n_keep] - vals[0][:n_keep]).max() < 1e-3 and np.abs(dec[n_keep:]).max() < 1e-3
    assert MaskTrunc(cps, got, o.slots) is got  # N == slots returns the input (crypto/basics.go:111-113)


def test_missing_relin_key_fails_loudly():
    from oracle.oracle import small_params
    from sfgwas_b200 import Ciphertext, CMult, CryptoParams, SfgError

    p = small_params(8, "pn13")
    cps = CryptoParams(p["logN"], p["Q"], p["P"], p["scale"])
    z = np.zeros((2, 6, cps.N), dtype=np.uint64)
    with pytest.raises(SfgError, match="relinearisation key"):
        CMult(cps, [Ciphertext(z, cps.scale)], [Ciphertext(z, cps.scale)])


def _setup_lazy(o, sk, rng, nsnp, nind, kp):
    X = rng.integers(0, 3, (nsnp, nind)).astype(np.int8)
    mean = X.mean(axis=1)
    stdinv = 1.0 / (X.std(axis=1) + 0.5)
    Qp = rng.normal(size=(kp, nsnp)) * 0.3
    return X, mean, stdinv, Qp


def _enc_vec(o, sk, v, level, seed0):
    n = (len(v) - 1) // o.slots + 1
    return [o.encrypt_vector(sk, v[b * o.slots:(b + 1) * o.slots], level, seed=seed0 + b) for b in range(n)]


def test_qx_lazy_norm_stream(env):
    """gwas/matmult.go:27-77 end to end: bit-exact vs the oracle composition, decrypts to Q*S*(X - m 1^T)."""
    from sfgwas_b200 import Ciphertext, GenoFileStream, MatMult4StreamPreprocess, QXLazyNormStream

    o, cps, sk, keys, rlk = env
    rng = np.random.default_rng(4)
    nsnp, nind, kp = o.slots + 37, o.slots + 11, 2
    X, mean, stdinv, Qp = _setup_lazy(o, sk, rng, nsnp, nind, kp)
    top = o.nQ - 1
    Q = [_enc_vec(o, sk, Qp[i], top, 100 + 10 * i) for i in range(kp)]
    XMean, XStdInv = _enc_vec(o, sk, mean, top, 200), _enc_vec(o, sk, stdinv, top, 300)
    cache = MatMult4StreamPreprocess(cps, GenoFileStream.from_matrix(cps, X), 5)
    if o.nQ - 2 < 5:  # a 6-limb chain leaves Q*S at level 4 < 5: the reference dies in DropLevel (crypto/basics.go:806-824)
        from sfgwas_b200 import SfgError

        with pytest.raises(SfgError, match="smaller than the requested level"):
            QXLazyNormStream(cps, _Mpc(_Net(o, sk, Ciphertext)), [[Ciphertext(c, o.scale) for c in q] for q in Q], cache,
                             [Ciphertext(c, o.scale) for c in XMean], [Ciphertext(c, o.scale) for c in XStdInv], nind)
        return
    dc = o.preprocess(X, 5, nproc=1)

    def compute(QS):  # the oracle's MatMult4StreamCompute on (value, scale) pairs
        A = np.ascontiguousarray(np.stack([np.stack([v[:, :6] for v, _ in row]) for row in QS]))
        out = o.compute(A, dc, keys, 5, nproc=1)
        sc = QS[0][0][1] * o.scale
        return [[(out[i, j], sc) for j in range(out.shape[1])] for i in range(out.shape[0])]

    net = _Net(o, sk, lambda v, s: (v, s))
    want = o.QXLazyNormStream([[(c, o.scale) for c in q] for q in Q], compute, lambda M: net.BootstrapMatAll(None, M),
                              [(c, o.scale) for c in XMean], [(c, o.scale) for c in XStdInv], nind, rlk, keys)
    got = QXLazyNormStream(cps, _Mpc(_Net(o, sk, Ciphertext)), [[Ciphertext(c, o.scale) for c in q] for q in Q], cache,
                           [Ciphertext(c, o.scale) for c in XMean], [Ciphertext(c, o.scale) for c in XStdInv], nind)
    o.cache_free(dc)
    ref = (Qp * stdinv) @ (X - mean[:, None])
    for i in range(kp):
        for j in range(len(got[i])):
            assert (got[i][j].value == want[i][j][0]).all(), (i, j)
            assert got[i][j].scale == want[i][j][1]
            dec = o.decrypt_vector(sk, got[i][j].value, got[i][j].scale).real
            seg = ref[i, j * o.slots:(j + 1) * o.slots]
            assert np.abs(dec[: len(seg)] - seg).max() < 2e-2, np.abs(dec[: len(seg)] - seg).max()
            if len(seg) < o.slots:
                assert np.abs(dec[len(seg):]).max() < 1e-3  # MaskTrunc zeroed the tail
    assert QXLazyNormStream(cps, _Mpc(None, pid=0), None, None, None, None, 0) is None  # party 0 returns immediately (:28-30)


def test_qxt_lazy_norm_stream(env):
    """gwas/matmult.go:83-116: bit-exact vs the oracle composition, decrypts to Q*(X^T - 1 m^T)*S."""
    from sfgwas_b200 import Ciphertext, GenoFileStream, MatMult4StreamPreprocess, QXtLazyNormStream

    o, cps, sk, keys, rlk = env
    rng = np.random.default_rng(5)
    nsnp, nind, kp = o.slots + 21, o.slots - 9, 2
    X, mean, stdinv, _ = _setup_lazy(o, sk, rng, nsnp, nind, kp)
    Qp = rng.normal(size=(kp, nind)) * 0.3
    XT = np.ascontiguousarray(X.T)
    top = o.nQ - 1
    Q = [_enc_vec(o, sk, Qp[i], top, 400 + 10 * i) for i in range(kp)]
    XMean, XStdInv = _enc_vec(o, sk, mean, top, 500), _enc_vec(o, sk, stdinv, top, 600)
    cache = MatMult4StreamPreprocess(cps, GenoFileStream.from_matrix(cps, XT), 5)
    dc = o.preprocess(XT, 5, nproc=1)

    def compute(QQ):
        A = np.ascontiguousarray(np.stack([np.stack([v[:, :6] for v, _ in row]) for row in QQ]))
        out = o.compute(A, dc, keys, 5, nproc=1)
        sc = QQ[0][0][1] * o.scale
        return [[(out[i, j], sc) for j in range(out.shape[1])] for i in range(out.shape[0])]

    net = _Net(o, sk, lambda v, s: (v, s))
    want = o.QXtLazyNormStream([[(c, o.scale) for c in q] for q in Q], compute, lambda M: net.BootstrapMatAll(None, M),
                               [(c, o.scale) for c in XMean], [(c, o.scale) for c in XStdInv], rlk, keys)
    got = QXtLazyNormStream(cps, _Mpc(_Net(o, sk, Ciphertext)), [[Ciphertext(c, o.scale) for c in q] for q in Q], cache,
                            [Ciphertext(c, o.scale) for c in XMean], [Ciphertext(c, o.scale) for c in XStdInv])
    o.cache_free(dc)
    ref = (Qp @ (X.T - mean[None, :])) * stdinv[None, :]
    for i in range(kp):
        for j in range(len(got[i])):
            assert (got[i][j].value == want[i][j][0]).all(), (i, j)
            dec = o.decrypt_vector(sk, got[i][j].value, got[i][j].scale).real
            seg = ref[i, j * o.slots:(j + 1) * o.slots]
            assert np.abs(dec[: len(seg)] - seg).max() < 2e-2, np.abs(dec[: len(seg)] - seg).max()
    # the same wrapper with Q, XMean and XStdInv resident in HBM (sfg_cts handles): the MatMult, the row sums, the mean correction and the
    # final scaling chain on the device; bit-identical to the host-buffer path
    from sfgwas_b200 import DeviceCipherVector, QXtLazyNormStreamDevice

    dQ = DeviceCipherVector.upload(cps, [Ciphertext(c, o.scale) for q in Q for c in q])
    dMean = DeviceCipherVector.upload(cps, [Ciphertext(c, o.scale) for c in XMean])
    dStd = DeviceCipherVector.upload(cps, [Ciphertext(c, o.scale) for c in XStdInv])
    rows = QXtLazyNormStreamDevice(cps, _Mpc(_Net(o, sk, Ciphertext)), dQ, kp, cache, dMean, dStd)
    for i in range(kp):
        host = rows[i].download()
        assert len(host) == len(got[i])
        for j in range(len(host)):
            assert (host[j].value == got[i][j].value).all() and abs(host[j].scale / got[i][j].scale - 1) < 1e-12, (i, j)
