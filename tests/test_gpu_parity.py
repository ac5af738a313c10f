"""GPU parity tests: the CUDA path, called through the C ABI (sfgwas_b200.gwas mirrors the Go entry points), against the
CPU oracle on the same seeded inputs, and against the committed golden vectors.  Integer work must be BIT-EXACT."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    with open(os.path.join(G, name + ".json")) as f:
        return json.load(f)


def sha(xs):
    h = hashlib.sha256()
    for x in xs:
        h.update(int(x).to_bytes(8, "little"))
    return h.hexdigest()


def u64(xs):
    return np.array(xs, dtype=np.uint64)


def make(params):
    from oracle.oracle import Oracle
    from sfgwas_b200 import CryptoParams

    o = Oracle.from_params(params)
    cps = CryptoParams(params["logN"], params["Q"], params["P"], params["scale"])
    return o, cps


@pytest.fixture(scope="module")
def small13():
    from oracle.oracle import small_params

    o, cps = make(small_params(8, "pn13"))
    sk = o.keygen_secret(1)
    keys = o.gen_bsgs_keys(sk)
    cps.SetRotKeys(keys)
    return o, cps, sk, keys


@pytest.fixture(scope="module")
def small14():
    from oracle.oracle import small_params

    o, cps = make(small_params(9, "pn14"))
    sk = o.keygen_secret(1)
    keys = o.gen_bsgs_keys(sk)
    cps.SetRotKeys(keys)
    return o, cps, sk, keys


# ---- K1 / K2 / K3 ----------------------------------------------------------------------------------------------------
def test_k1_k2_k3_golden():
    from sfgwas_b200 import CryptoParams

    cases = load("primitives")["cases"]
    for case in cases:
        q = case["q"]
        # a context whose modulus 0 is q (any NTT-friendly logN that divides q-1)
        logN = 8
        cps = CryptoParams(logN, [q], [0x800004001 if q != 0x800004001 else 0x1FFFEC001], 2.0 ** 30)
        a, b = u64(case["a"]), u64(case["b"])
        acc = u64(case["acc_in"])
        cps.MulCoeffsAndAdd128(a, b, acc)
        assert acc.tolist() == case["acc_out"]
        out = u64(case["red_in"])
        cps.ReduceAndAddUint128(acc, out, 0)
        assert out.tolist() == case["red_out"]
        p = np.zeros((1, cps.N), dtype=np.uint64)
        p[0, : len(case["a"])] = a
        assert cps.MFormLvl(0, p)[0, : len(case["a"])].tolist() == case["mform"]
        cps.close()


def test_k1_k2_vs_oracle_random(small13):
    o, cps, _, _ = small13
    rng = np.random.default_rng(0)
    n = 4096
    for limb in (0, 1, o.nQ):
        q = o.moduli[limb]
        a = rng.integers(0, q, n, dtype=np.uint64)
        b = rng.integers(0, q, n, dtype=np.uint64)
        acc_g = rng.integers(0, 1 << 63, (n, 2), dtype=np.uint64) * np.uint64(2) + np.uint64(1)
        acc_o = acc_g.copy()
        for _ in range(3):
            cps.MulCoeffsAndAdd128(a, b, acc_g)
            o.mul_coeffs_and_add128(a, b, acc_o)
        assert (acc_g == acc_o).all()
        out_g = rng.integers(0, q, n, dtype=np.uint64)
        out_o = out_g.copy()
        cps.ReduceAndAddUint128(acc_g, out_g, limb)
        o.reduce_and_add_uint128(acc_o, out_o, limb)
        assert (out_g == out_o).all()


# ---- NTT -------------------------------------------------------------------------------------------------------------
def test_ntt_golden():
    from sfgwas_b200 import CryptoParams

    by_logn = {}
    for case in load("ntt")["cases"]:
        by_logn.setdefault(case["logN"], []).append(case)
    for logN, cases in by_logn.items():
        qs = [c["q"] for c in cases]
        cps = CryptoParams(logN, qs[:-1] if len(qs) > 1 else qs, qs[-1:], 2.0 ** 30)
        assert cps.psi() == [c["psi"] for c in cases]
        for idx, c in enumerate(cases):
            q = c["q"]
            if "input" in c:
                a = u64(c["input"])
            else:
                r2 = random.Random(c["seed"])
                a = u64([r2.randrange(q) for _ in range(1 << logN)])
            A = cps.NTT(a[None], [idx])[0]
            assert A[:8].tolist() == c["first"]
            assert sha(A.tolist()) == c["sha256"]
            assert (cps.NTT(A[None], [idx], inverse=True)[0] == a).all()
        cps.close()


@pytest.mark.parametrize("logN", [6, 10, 12, 13, 14, 15, 16])
def test_ntt_vs_oracle_all_sizes(logN):
    from oracle.oracle import gen_primes

    qs = gen_primes(logN, 45, 2) + gen_primes(logN, 30, 1)
    ps = gen_primes(logN, 55, 1)
    o, cps = make(dict(logN=logN, Q=qs, P=ps, scale=2.0 ** 30))
    rng = np.random.default_rng(logN)
    polys = np.stack([rng.integers(0, m, o.N, dtype=np.uint64) for m in o.moduli] * 2)  # 2 groups of nQP limbs
    got = cps.NTT(polys, list(range(o.nQP)))
    want = np.stack([o.ntt(i % o.nQP, polys[i]) for i in range(polys.shape[0])])
    assert (got == want).all()
    back = cps.NTT(got, list(range(o.nQP)), inverse=True)
    assert (back == polys).all()
    cps.close()


# ---- encoder ---------------------------------------------------------------------------------------------------------
def test_encode_golden_small():
    from oracle.oracle import small_params
    from sfgwas_b200 import CryptoParams, GenoFileStream

    for case in load("encode")["cases"]:
        if "values" not in case or case["logN"] < 8:
            continue
        logN = case["logN"]
        p = small_params(logN, "pn13")
        cps = CryptoParams(logN, p["Q"], p["P"], case["scale"])
        n = cps.slots
        v = np.array(case["values"], dtype=np.int8)
        # a slots x slots block whose main diagonal (shift 0) is v
        X = np.zeros((n, n), dtype=np.int8)
        X[np.arange(n), np.arange(n)] = v
        g = GenoFileStream.from_matrix(cps, X)
        _, present, co = g.EncodeDiag(0, 0, case["nrot"], 5, mont=False, want_coeffs=True)
        assert present.all()
        assert co[0].tolist() == case["coeffs"]
        cps.close()


@pytest.mark.parametrize("name,nshift", [("PN13QP218", 6), ("PN14QP438", 3)])
def test_encode_vs_oracle_real_params(name, nshift):
    """Plaintext diagonals at the real parameter sets: bit-exact against the quad-precision oracle (and, on a few
    coefficients, against the 200-bit mpmath golden values)."""
    from oracle.oracle import PARAMS
    from sfgwas_b200 import GenoFileStream

    o, cps = make(PARAMS[name])
    rng = np.random.default_rng(13)
    n = o.slots
    nr, nc = n + 37, n + 11  # ragged second block row / column
    X = rng.integers(0, 3, (nr, nc)).astype(np.int8)
    g = GenoFileStream.from_matrix(cps, X)
    d = o.d
    shifts = [0, 1, d, n - 1] + [int(x) for x in rng.integers(0, n, nshift)]
    for bi in (0, 1):
        for shift in shifts[: (len(shifts) if bi == 0 else 3)]:
            nrot = d * (shift // d)
            pv, present = g.EncodeDiag(bi, shift, nrot, 5, mont=True)
            for bj in range(2):
                blk = X[bi * n : (bi + 1) * n, bj * n : (bj + 1) * n]
                ok, diag = o.get_diag(blk, -shift)
                assert ok == bool(present[bj])
                if ok:
                    want = o.mform_lvl(5, o.encode_ntt(diag, nrot, 5))
                    assert (pv[bj] == want).all(), (bi, shift, bj)
    # golden picks (mpmath, 200 bit)
    for case in load("encode")["cases"]:
        if case["logN"] == o.logN and "picks" in case:
            r2 = random.Random(case["seed"])
            v = np.array([r2.choice([0, 1, 2]) for _ in range(n)], dtype=np.int8)
            Xd = np.zeros((n, n), dtype=np.int8)
            Xd[np.arange(n), np.arange(n)] = v
            g2 = GenoFileStream.from_matrix(cps, Xd)
            _, _, co = g2.EncodeDiag(0, 0, 0, 0, mont=False, want_coeffs=True)
            assert [int(co[0][k]) for k in case["picks"]] == case["pick_coeffs"]
    rechecked, unresolved = cps.encoder_stats()
    assert unresolved == 0
    cps.close()


# ---- rotation / key-switch ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", ["pn13", "pn14", "pn15"])
def test_rotate_vs_python_bignum_tiny(shape):
    import pyref
    from oracle.oracle import small_params

    logN = 6
    o, cps = make(small_params(logN, shape))
    sk = o.keygen_secret(9)
    k = 5
    swk = o.gen_rotation_key(sk, k, seed=21)
    cps.SetRotKey(k, swk)
    rng = np.random.default_rng(4)
    for level in (5, 4):
        ct = np.stack([np.stack([rng.integers(0, o.Q[l], o.N, dtype=np.uint64) for l in range(level + 1)]) for _ in range(2)])
        got = cps.RotateRightWithEvaluator(ct, -k)
        want = pyref.rotate_right_ref(o.Q, o.P, logN, level, ct.tolist(), -k, swk.tolist())
        assert got.tolist() == want
        assert (got == o.rotate_right(ct, -k, swk)).all()
    cps.close()


@pytest.mark.parametrize("fix", ["small13", "small14"])
def test_rotate_vs_oracle(fix, request):
    o, cps, sk, keys = request.getfixturevalue(fix)
    rng = np.random.default_rng(2)
    for level in (5, 4):
        cts = np.stack([np.stack([np.stack([rng.integers(0, o.Q[l], o.N, dtype=np.uint64) for l in range(level + 1)])
                                  for _ in range(2)]) for _ in range(3)])
        for k in (1, o.d - 1, o.d, 2 * o.d):
            got = cps.RotateRightWithEvaluator(cts, -k)
            for t in range(3):
                assert (got[t] == o.rotate_right(cts[t], -k, keys[k])).all(), (level, k, t)
        assert (cps.RotateRightWithEvaluator(cts, 0) == cts).all()
        assert (cps.RotateRightWithEvaluator(cts, o.slots) == cts).all()


def test_rotate_real_params_pn13():
    from oracle.oracle import PARAMS

    o, cps = make(PARAMS["PN13QP218"])
    sk = o.keygen_secret(3)
    rng = np.random.default_rng(8)
    for k in (1, 64):
        swk = o.gen_rotation_key(sk, k)
        cps.SetRotKey(k, swk)
        for level in (5, 4):
            ct = np.stack([np.stack([rng.integers(0, o.Q[l], o.N, dtype=np.uint64) for l in range(level + 1)]) for _ in range(2)])
            assert (cps.RotateRightWithEvaluator(ct, -k) == o.rotate_right(ct, -k, swk)).all()
    cps.close()


def test_missing_key_fails_loudly(small13):
    from sfgwas_b200 import SfgError

    o, cps, _, _ = small13
    ct = np.zeros((2, 6, o.N), dtype=np.uint64)
    with pytest.raises(SfgError):
        cps.RotateRightWithEvaluator(ct, -(o.d + 1))  # no key for this rotation


# ---- the stream entry points ---------------------------------------------------------------------------------------------
def enc_matrix(o, sk, Ap, level=5):
    s, nr = Ap.shape
    nbr = (nr - 1) // o.slots + 1
    A = np.zeros((s, nbr, 2, level + 1, o.N), dtype=np.uint64)
    for i in range(s):
        for b in range(nbr):
            A[i, b] = o.encrypt_vector(sk, Ap[i, b * o.slots : (b + 1) * o.slots], level, seed=100 + 17 * i + b)
    return A


SHAPES = [(200, 300, 2), (128, 128, 1), (77, 50, 3), (1, 5, 1), (129, 1, 2), (300, 129, 5), (260, 520, 10)]


@pytest.mark.parametrize("nr,nc,s", SHAPES)
def test_preprocess_compute_bit_exact(small13, nr, nc, s):
    from sfgwas_b200 import GenoFileStream, MatMult4StreamCompute, MatMult4StreamPreprocess

    o, cps, sk, keys = small13
    rng = np.random.default_rng(nr * 7 + nc)
    X = rng.integers(0, 3, (nr, nc)).astype(np.int8)
    Ap = rng.normal(size=(s, nr))
    A = enc_matrix(o, sk, Ap)
    gfs = GenoFileStream.from_matrix(cps, X)
    cache = MatMult4StreamPreprocess(cps, gfs, 5, "unused_prefix")
    assert cache.materialised
    dc = o.preprocess(X, 5, nproc=1)
    assert cache.num_polys == o.L.orc_diag_cache_num_polys(dc)
    # cached plaintexts: bit-exact (limbs 0..4 are what the MAC reads)
    for bi in range(cache.num_block_rows):
        for shift in (0, 1, o.d, o.slots - 1):
            for bj in range(cache.m_ct):
                want = o.cache_get(dc, bi, shift, bj)
                got = cache.get_diag(bi, shift, bj)
                assert (want is None) == (got is None)
                if want is not None:
                    assert (got == want[:5]).all()
    out = MatMult4StreamCompute(cps, A, 5, cache)
    want = o.compute(A, dc, keys, 5, nproc=1)
    o.cache_free(dc)
    assert out.shape == want.shape
    if not (out == want).all():
        # A mismatch is a HARD failure.  Before failing, say where and which side moved (round 1 saw a rare first-call mismatch whose
        # root cause was a legacy-stream table upload racing the non-blocking context stream: DESIGN.md 6b).
        bad = out != want
        again = MatMult4StreamCompute(cps, A, 5, cache)
        dc1 = o.preprocess(X, 5, nproc=1)
        want1 = o.compute(A, dc1, keys, 5, nproc=1)
        o.cache_free(dc1)
        raise AssertionError("CUDA != oracle: %d of %d words, first %s; cuda repeatable=%s, oracle repeatable=%s, cuda==oracle(rerun)=%s"
                             % (int(bad.sum()), bad.size, np.argwhere(bad)[:3].tolist(), bool((again == out).all()),
                                bool((want1 == want).all()), bool((out == want1).all())))
    # numerical meaning (the reference's CPMatMult0 notion): decrypt(out) ~= A_plain . X ; tolerance 1e-4 relative
    ref = Ap @ X.astype(float)
    tol = 1e-4 * max(1.0, np.abs(ref).max())
    got = o.decrypt_vector(sk, out[0, 0], o.scale * o.scale).real
    w = ref[0, : o.slots]
    assert np.abs(got[: len(w)] - w).max() < tol


def test_compute_on_the_fly_cache_identical(small13):
    """A cache that does not fit the HBM budget regenerates diagonals per giant chunk: same bits."""
    from sfgwas_b200 import GenoFileStream, MatMult4StreamCompute, MatMult4StreamPreprocess

    o, cps, sk, keys = small13
    rng = np.random.default_rng(5)
    X = rng.integers(0, 3, (200, 300)).astype(np.int8)
    A = enc_matrix(o, sk, rng.normal(size=(2, 200)))
    gfs = GenoFileStream.from_matrix(cps, X)
    c1 = MatMult4StreamPreprocess(cps, gfs, 5)
    cps.set_cache_budget(1)
    c2 = MatMult4StreamPreprocess(cps, gfs, 5)
    cps.set_cache_budget(0)
    assert c1.materialised and not c2.materialised
    assert (MatMult4StreamCompute(cps, A, 5, c1) == MatMult4StreamCompute(cps, A, 5, c2)).all()


@pytest.mark.parametrize("s", [3, 15])  # 15 (kp = 15): 30 rows x 5-6 byte planes exceed half of TMEM -> the MAC runs as two row halves
def test_pn14_shape_bit_exact(small14, s):
    from sfgwas_b200 import GenoFileStream, MatMult4StreamCompute, MatMult4StreamPreprocess

    o, cps, sk, keys = small14
    rng = np.random.default_rng(14)
    X = rng.integers(0, 3, (300, 520)).astype(np.int8)
    A = enc_matrix(o, sk, rng.normal(size=(s, 300)), level=7)  # level 7 input is dropped to 5
    gfs = GenoFileStream.from_matrix(cps, X)
    cache = MatMult4StreamPreprocess(cps, gfs, 5)
    out = MatMult4StreamCompute(cps, A, 5, cache)
    dc = o.preprocess(X, 5, nproc=1)
    want = o.compute(A, dc, keys, 5, nproc=1)
    o.cache_free(dc)
    assert (out == want).all()


@pytest.mark.parametrize("s", [9, 13, 15, 16])
def test_pair_mode_shares_the_p_stream_bit_exact(small14, s, monkeypatch):
    """kp = 13 .. 16 at the logN-14 limb widths: the 2 kp rows do not fit TMEM twice, so the MAC runs as two row parts, by default as two
    launches.  SFG_TC_PAIR=1 runs them as ONE launch of 2-CTA clusters (each CTA fetches half of every P stage and multicasts it to
    both: P leaves HBM once).  Same bits, several column tiles per cluster, resident cache and on-the-fly diagonals; kp = 9 (one part)
    must not be affected by the switch."""
    from sfgwas_b200 import GenoFileStream, MatMult4StreamCompute, MatMult4StreamPreprocess

    o, cps, sk, keys = small14
    rng = np.random.default_rng(140 + s)
    nr, nc = 2 * o.slots + 11, 13 * o.slots + 5  # 3 block rows; 14 block columns x d giants > 128 columns: more than one column tile
    X = rng.integers(0, 3, (nr, nc)).astype(np.int8)
    A = enc_matrix(o, sk, rng.normal(size=(s, nr)))
    gfs = GenoFileStream.from_matrix(cps, X)
    cache = MatMult4StreamPreprocess(cps, gfs, 5)
    two_launches = MatMult4StreamCompute(cps, A, 5, cache)
    launches0 = cps.launch_count()
    monkeypatch.setenv("SFG_TC_PAIR", "1")
    paired = MatMult4StreamCompute(cps, A, 5, cache)
    launches_paired = cps.launch_count() - launches0
    cps.set_cache_budget(1)
    otf = MatMult4StreamPreprocess(cps, gfs, 5)
    cps.set_cache_budget(0)
    assert not otf.materialised
    paired_otf = MatMult4StreamCompute(cps, A, 5, otf)
    monkeypatch.delenv("SFG_TC_PAIR")
    launches0 = cps.launch_count()
    MatMult4StreamCompute(cps, A, 5, cache)
    launches_two = cps.launch_count() - launches0
    assert (paired == two_launches).all()
    assert (paired_otf == two_launches).all()
    assert (launches_paired < launches_two) if s >= 13 else (launches_paired == launches_two)  # one MAC launch instead of two when the rows are split
    if s == 15:
        dc = o.preprocess(X, 5, nproc=1)
        want = o.compute(A, dc, keys, 5, nproc=1)
        o.cache_free(dc)
        assert (paired == want).all()


@pytest.mark.parametrize("nc", [140, 144])  # 144: the 16-byte genotype scan (ncols % 16 == 0); 140: the byte-wise one
def test_matmult4_stream_fused(small13, nc):
    from sfgwas_b200 import GenoFileStream, MatMult4Stream

    o, cps, sk, keys = small13
    rng = np.random.default_rng(21)
    X = rng.integers(-1, 3, (150, nc)).astype(np.int8)  # -1 = missing
    A = enc_matrix(o, sk, rng.normal(size=(2, 150)))
    gfs = GenoFileStream.from_matrix(cps, X)
    for sq_sum, square in ((True, True), (True, False), (False, False)):
        out, sm, sq = MatMult4Stream(cps, A, gfs, 5, sq_sum, square, 0)
        want, wsm, wsq = o.matmult4_stream(A, X, keys, 5, sq_sum, square, nproc=1)
        assert (out == want).all()
        if sq_sum:
            assert (sm == wsm).all() and (sq == wsq).all()
        else:
            assert sm is None and sq is None


def test_error_behaviour(small13):
    from sfgwas_b200 import GenoFileStream, MatMult4StreamCompute, MatMult4StreamPreprocess, SfgError

    o, cps, sk, keys = small13
    X = np.ones((10, 10), dtype=np.int8)
    gfs = GenoFileStream.from_matrix(cps, X)
    cache = MatMult4StreamPreprocess(cps, gfs, 5)
    with pytest.raises(SfgError):  # fewer than maxLevel limbs: the reference indexes past Coeffs[] (gwas/matmult.go:393) and panics
        MatMult4StreamCompute(cps, np.zeros((1, 1, 2, 4, o.N), dtype=np.uint64), 5, cache)
    with pytest.raises(SfgError):  # wrong number of block rows
        MatMult4StreamCompute(cps, np.zeros((1, 2, 2, 6, o.N), dtype=np.uint64), 5, cache)
    g2 = GenoFileStream(cps, 4, 4)
    g2.push_rows(np.zeros((2, 4), dtype=np.int8))
    with pytest.raises(SfgError):  # incomplete stream
        MatMult4StreamPreprocess(cps, g2, 5)


def test_input_one_level_below_maxlevel(small13):
    """gwas/matmult.go:1053-1056 drops A only when Level() > maxLevel: an input at level maxLevel-1 (e.g. QS = CMult(Q, XStdInv) landing one
    level lower) is valid in the reference -- rotated at its own level, output at level maxLevel-1."""
    from sfgwas_b200 import GenoFileStream, MatMult4StreamCompute, MatMult4StreamPreprocess

    o, cps, sk, keys = small13
    rng = np.random.default_rng(41)
    X = rng.integers(0, 3, (200, 300)).astype(np.int8)
    Ap = rng.normal(size=(2, 200))
    A = enc_matrix(o, sk, Ap, level=4)
    cache = MatMult4StreamPreprocess(cps, GenoFileStream.from_matrix(cps, X), 5)
    out = MatMult4StreamCompute(cps, A, 5, cache)
    dc = o.preprocess(X, 5, nproc=1)
    want = o.compute(A, dc, keys, 5, nproc=1)
    o.cache_free(dc)
    assert (out == want).all()
    got = o.decrypt_vector(sk, out[1, 0], o.scale * o.scale).real
    ref = (Ap @ X.astype(float))[1, : o.slots]
    assert np.abs(got[: len(ref)] - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("s", [17, 35])
def test_more_than_16_rows(small13, s):
    """The reference has no limit on len(A) (gwas/assoc.go:700-718 passes len(Q)+2 rows; a PCA with kp > 16): more rows than one
    tensor-core pass holds run as balanced row passes with identical bits."""
    from sfgwas_b200 import GenoFileStream, MatMult4StreamCompute, MatMult4StreamPreprocess

    o, cps, sk, keys = small13
    rng = np.random.default_rng(1000 + s)
    X = rng.integers(0, 3, (150, 260)).astype(np.int8)
    A = enc_matrix(o, sk, rng.normal(size=(s, 150)))
    cache = MatMult4StreamPreprocess(cps, GenoFileStream.from_matrix(cps, X), 5)
    out = MatMult4StreamCompute(cps, A, 5, cache)
    dc = o.preprocess(X, 5, nproc=1)
    want = o.compute(A, dc, keys, 5, nproc=1)
    o.cache_free(dc)
    assert out.shape == want.shape and (out == want).all()


def test_ptrs_entry_points_equal_flat(small13):
    """The entry points the cgo shim binds (go/gwas/matmult_b200.go): one PAGEABLE array per limb for A, the result, the rotation keys
    and the relinearisation key -- exactly what Go hands over ([][]uint64 Coeffs).  Bit-equal to the flat entry points."""
    import ctypes as C

    from sfgwas_b200 import CryptoParams, GenoFileStream, MatMult4StreamCompute, MatMult4StreamPreprocess

    o, cps, sk, keys = small13
    p = dict(logN=o.logN, Q=o.Q, P=o.P, scale=o.scale)
    cps2 = CryptoParams(p["logN"], p["Q"], p["P"], p["scale"])
    L = cps2.L
    keep = []

    def limb_ptrs(arr):  # every limb a separate heap allocation, like a Go slice
        flat = arr.reshape(-1, o.N)
        limbs = [np.array(flat[k], dtype=np.uint64, copy=True) for k in range(flat.shape[0])]
        keep.append(limbs)
        return (C.c_void_p * len(limbs))(*[x.ctypes.data for x in limbs]), limbs

    for k, v in keys.items():
        ptrs, _ = limb_ptrs(v)
        cps2._check(L.sfg_ctx_set_rotation_key_ptrs(cps2.h, k, ptrs), "set_rotation_key_ptrs")
    rlk = o.gen_relin_key(sk)
    ptrs, _ = limb_ptrs(rlk)
    cps2._check(L.sfg_ctx_set_relin_key_ptrs(cps2.h, ptrs), "set_relin_key_ptrs")
    assert L.sfg_ctx_has_rotation_key(cps2.h, 1) == 1

    rng = np.random.default_rng(77)
    for nr, nc, s in ((260, 520, 10), (130, 77, 3), (150, 260, 18)):
        X = rng.integers(0, 3, (nr, nc)).astype(np.int8)
        A = enc_matrix(o, sk, rng.normal(size=(s, nr)))
        want = MatMult4StreamCompute(cps, A, 5, MatMult4StreamPreprocess(cps, GenoFileStream.from_matrix(cps, X), 5))
        cache = MatMult4StreamPreprocess(cps2, GenoFileStream.from_matrix(cps2, X), 5)
        a_ptrs, _ = limb_ptrs(A)
        out = np.full((s, cache.m_ct, 2, 5, o.N), 0xA5A5A5A5A5A5A5A5, dtype=np.uint64)
        o_ptrs, o_limbs = limb_ptrs(out)
        for rep in range(2):  # twice: the staging buffers are reused
            cps2._check(L.sfg_matmult4_stream_compute_ptrs(cps2.h, a_ptrs, s, A.shape[1], 5, 5, cache.h, o_ptrs), "compute_ptrs")
            got = np.stack(o_limbs).reshape(want.shape)
            assert (got == want).all(), (nr, nc, s, rep)
    # the relinearisation key uploaded through _ptrs multiplies like the flat one
    from sfgwas_b200 import Ciphertext, CMult, SetRelinKey

    SetRelinKey(cps, rlk)
    v = rng.normal(size=o.slots)
    ct = Ciphertext(o.encrypt_vector(sk, v, 5, seed=5), o.scale)
    assert (CMult(cps, [ct], [ct])[0].value == CMult(cps2, [ct], [ct])[0].value).all()
    cps2.close()


def test_seven_k_groups_transposed_shape(small13):
    """The transposed orientation of BASELINE config 2 (100k x 10k: 25 block rows, K = 1 600 baby-step slots = 7 K groups of the
    tensor-core MAC) at the same SHAPE on a small ring: enough block rows for 7 K groups, 3 block columns, s = 10."""
    from sfgwas_b200 import GenoFileStream, MatMult4StreamCompute, MatMult4StreamPreprocess

    o, cps, sk, keys = small13
    nbr = (6 * 256) // o.d + 1  # K = nbr * d > 6 * 256  ->  7 K groups
    nr, nc, s = (nbr - 1) * o.slots + 17, 2 * o.slots + 30, 10
    rng = np.random.default_rng(7)
    X = rng.integers(0, 3, (nr, nc)).astype(np.int8)
    A = enc_matrix(o, sk, rng.normal(size=(s, nr)))
    cache = MatMult4StreamPreprocess(cps, GenoFileStream.from_matrix(cps, X), 5)
    assert cache.num_block_rows == nbr and nbr * o.d > 6 * 256
    out = MatMult4StreamCompute(cps, A, 5, cache)
    dc = o.preprocess(X, 5, nproc=8)
    want = o.compute(A, dc, keys, 5, nproc=8)
    o.cache_free(dc)
    assert (out == want).all()


def test_k_groups_split_over_launches_accumulate_mod_q(small13, monkeypatch):
    """K groups normally accumulate in TMEM inside ONE launch; beyond the s32 range of the partial sums (K > ~6 600) the K range is cut into
    several launches that add into cv mod q in the epilogue.  SFG_TC_MAXGROUPS forces that path on a 7-group shape (2 + 2 + 2 + 1 groups):
    same bits as the fused launch and as the oracle (covers the read-modify-write of the 64-byte output runs)."""
    from sfgwas_b200 import GenoFileStream, MatMult4StreamCompute, MatMult4StreamPreprocess

    o, cps, sk, keys = small13
    nbr = (6 * 256) // o.d + 1
    nr, nc, s = (nbr - 1) * o.slots + 5, o.slots + 9, 3
    rng = np.random.default_rng(71)
    X = rng.integers(0, 3, (nr, nc)).astype(np.int8)
    A = enc_matrix(o, sk, rng.normal(size=(s, nr)))
    cache = MatMult4StreamPreprocess(cps, GenoFileStream.from_matrix(cps, X), 5)
    fused = MatMult4StreamCompute(cps, A, 5, cache)
    monkeypatch.setenv("SFG_TC_MAXGROUPS", "2")
    split = MatMult4StreamCompute(cps, A, 5, cache)
    monkeypatch.delenv("SFG_TC_MAXGROUPS")
    assert (split == fused).all()
    dc = o.preprocess(X, 5, nproc=8)
    want = o.compute(A, dc, keys, 5, nproc=8)
    o.cache_free(dc)
    assert (fused == want).all()


def test_on_the_fly_image_chunked_by_k_group(small13, monkeypatch):
    """Diagonals re-encoded inside the call, many block rows: when the K groups of ONE column tile exceed the temporary image (config 4
    at full size: 5 groups x 14 GB), the image holds one tile and as many K groups as fit, and the MAC adds group chunk after group
    chunk into cv.  SFG_OTF_IMG_MB shrinks the image so that a 7-group shape runs as 2 + 2 + 2 + 1: same bits as the resident cache."""
    from sfgwas_b200 import GenoFileStream, MatMult4StreamCompute, MatMult4StreamPreprocess

    o, cps, sk, keys = small13
    nbr = (6 * 256) // o.d + 1
    nr, nc, s = (nbr - 1) * o.slots + 5, 2 * o.slots + 9, 3
    rng = np.random.default_rng(72)
    X = rng.integers(0, 3, (nr, nc)).astype(np.int8)
    A = enc_matrix(o, sk, rng.normal(size=(s, nr)))
    gfs = GenoFileStream.from_matrix(cps, X)
    resident = MatMult4StreamCompute(cps, A, 5, MatMult4StreamPreprocess(cps, gfs, 5))
    cps.set_cache_budget(1)
    c2 = MatMult4StreamPreprocess(cps, gfs, 5)
    cps.set_cache_budget(0)
    assert not c2.materialised
    whole_tiles = MatMult4StreamCompute(cps, A, 5, c2)
    per_tg_mb = o.N * sum((int(q) - 1).bit_length() + 7 >> 3 for q in o.Q[:5]) * 128 * 256 >> 20
    monkeypatch.setenv("SFG_OTF_IMG_MB", str(2 * per_tg_mb + 1))
    by_group = MatMult4StreamCompute(cps, A, 5, c2)
    monkeypatch.delenv("SFG_OTF_IMG_MB")
    assert (whole_tiles == resident).all()
    assert (by_group == resident).all()


@pytest.mark.parametrize("case", ["pn13_2x4_s10", "pn14_2x2_s15"])
def test_benchmarked_geometry_golden(case):
    """Preprocess + Compute at the REAL parameter sets and the benchmarked geometry (several block rows, full 128-column tiles, giant
    chunks, row chunks; PN14: two-half MAC, half-ring key-switch, nP = 2) against the CPU oracle: the oracle ran offline
    (tests/golden/gen_parity_big.py, minutes on 8 cores) and left SHA-256 digests of every output ciphertext; the same seeded inputs are
    regenerated here and ALL of `out` must hash identically."""
    import importlib.util

    from oracle.oracle import PARAMS, Oracle
    from sfgwas_b200 import CryptoParams, GenoFileStream, MatMult4StreamCompute, MatMult4StreamPreprocess

    path = os.path.join(G, "parity_%s.json" % case)
    if not os.path.exists(path):
        pytest.skip("fixture %s not generated" % path)
    spec = importlib.util.spec_from_file_location("gen_parity_big", os.path.join(G, "gen_parity_big.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    with open(path) as f:
        gold = json.load(f)
    cs = gold["case"]
    prm = PARAMS[cs["params"]]
    o = Oracle.from_params(prm)
    X, Ap, sk, keys, A = gen.inputs(o, cs)
    assert hashlib.sha256(X.tobytes()).hexdigest() == gold["input_digests"]["X"], "seeded genotypes differ from the fixture's"
    assert gen.digest(A) == gold["input_digests"]["A"], "seeded ciphertexts differ from the fixture's"
    cps = CryptoParams(prm["logN"], prm["Q"], prm["P"], prm["scale"])
    cps.SetRotKeys(keys)
    cache = MatMult4StreamPreprocess(cps, GenoFileStream.from_matrix(cps, X), 5)
    assert cache.materialised and cache.num_polys == gold["num_polys"]
    for key, want in gold["diag_sha256"].items():
        bi, shift, bj = (int(x) for x in key.split(","))
        got = cache.get_diag(bi, shift, bj)
        assert (got is None) == (want is None), key
        if want is not None:
            assert gen.digest(got) == want, "cached diagonal %s differs from the oracle's" % key
    out = MatMult4StreamCompute(cps, A, 5, cache)
    assert list(out.shape) == gold["shape"]
    s, m_ct = out.shape[:2]
    bad = [(i, bj) for i in range(s) for bj in range(m_ct) if gen.digest(out[i, bj]) != gold["out_sha256"][i][bj]]
    assert not bad, "output ciphertexts differ from the oracle's: %s" % bad[:8]
    ref = Ap @ X.astype(float)
    got = o.decrypt_vector(sk, out[s - 1, m_ct - 1], o.scale * o.scale).real
    w = ref[s - 1, (m_ct - 1) * o.slots:]
    assert np.abs(got[: len(w)] - w).max() < 1e-3 * max(1.0, np.abs(ref).max())  # CKKS tolerance
    # the non-materialised path (diagonals regenerated per chunk: the config 4 / 5 regime) gives the same bits at this geometry
    cps.set_cache_budget(1)
    c2 = MatMult4StreamPreprocess(cps, GenoFileStream.from_matrix(cps, X), 5)
    cps.set_cache_budget(0)
    assert not c2.materialised
    assert (MatMult4StreamCompute(cps, A, 5, c2) == out).all()
    cps.close()


@pytest.mark.parametrize("pname", ["PN13QP218", "PN14QP438"])
def test_linearity_property_full_size_pn13(pname):
    """Size-independent property at the real logN=13 / logN=14 rings: MatMult(A1 + A2) == MatMult(A1) + MatMult(A2) is NOT bitwise
    (key-switch rounding), but the MAC stage is: cv is linear mod q. Checked through one block with s=2 rows where row 1 =
    2*row 0 (mod q): out row 1 decrypts to twice row 0; and idempotence: the same call twice gives identical bits."""
    from oracle.oracle import PARAMS
    from sfgwas_b200 import GenoFileStream, MatMult4StreamCompute, MatMult4StreamPreprocess

    o, cps = make(PARAMS[pname])
    sk = o.keygen_secret(5)
    d = o.d
    # a narrow matrix keeps the number of needed keys small: nr = 40 rows -> shifts 0..39 and 4096-c+1.. ; use c = 1 column block
    nr, nc = 40, 30
    rng = np.random.default_rng(99)
    X = rng.integers(0, 3, (nr, nc)).astype(np.int8)
    need = set()
    for shift in range(o.slots):
        if shift <= nr - 1 or shift >= o.slots - nc + 1:
            need.add(shift % d)
            need.add((shift // d) * d)
    need.discard(0)
    keys = {k: o.gen_rotation_key(sk, k) for k in sorted(need)}
    cps.SetRotKeys(keys)
    Ap = rng.normal(size=(1, nr))
    A1 = enc_matrix(o, sk, Ap)
    A = np.concatenate([A1, A1], axis=0)
    for l in range(6):
        q = np.uint64(o.Q[l])
        A[1, :, :, l] = (A[1, :, :, l] * np.uint64(2)) % q
    gfs = GenoFileStream.from_matrix(cps, X)
    cache = MatMult4StreamPreprocess(cps, gfs, 5)
    out = MatMult4StreamCompute(cps, A, 5, cache)
    out2 = MatMult4StreamCompute(cps, A, 5, cache)
    assert (out == out2).all()
    ref = (Ap @ X.astype(float))[0]
    g0 = o.decrypt_vector(sk, out[0, 0], o.scale * o.scale).real[:nc]
    g1 = o.decrypt_vector(sk, out[1, 0], o.scale * o.scale).real[:nc]
    assert np.abs(g0 - ref).max() < 1e-3 and np.abs(g1 - 2 * ref).max() < 2e-3
    # and bit-exact against the oracle at the real parameter set
    dc = o.preprocess(X, 5, nproc=1)
    want = o.compute(A, dc, keys, 5, nproc=1)
    o.cache_free(dc)
    assert (out == want).all()
    cps.close()


@pytest.mark.gpu
def test_many_block_rows_multiple_k_groups(small13):
    """17 block rows -> K = 17*d = 272 > 256 baby-step slots: the tensor-core MAC runs 2 K groups and accumulates mod q."""
    from sfgwas_b200 import GenoFileStream, MatMult4StreamCompute, MatMult4StreamPreprocess

    o, cps, sk, keys = small13
    rng = np.random.default_rng(17)
    nr, nc, s = 16 * o.slots + 9, 40, 2
    X = rng.integers(0, 3, (nr, nc)).astype(np.int8)
    A = enc_matrix(o, sk, rng.normal(size=(s, nr)))
    cache = MatMult4StreamPreprocess(cps, GenoFileStream.from_matrix(cps, X), 5)
    out = MatMult4StreamCompute(cps, A, 5, cache)
    dc = o.preprocess(X, 5, nproc=1)
    want = o.compute(A, dc, keys, 5, nproc=1)
    o.cache_free(dc)
    assert (out == want).all()


@pytest.mark.gpu
@pytest.mark.parametrize("nparts", [2, 5, 16])
def test_giant_sharded_partials_sum_to_compute(small13, nparts):
    """Giant-step sharding (bench.py --gpus N, dist.GiantSharded) on ONE GPU: the caches of the `nparts` shares are built one after the
    other, every share's Compute is a partial sum over its giant steps, and the shares add up mod q to the unsharded result bit for
    bit (16 shares > 12 giant steps: some shares are empty)."""
    import ctypes as C

    from sfgwas_b200 import DiagCache, GenoFileStream, MatMult4StreamCompute, MatMult4StreamPreprocess

    o, cps, sk, keys = small13
    rng = np.random.default_rng(400 + nparts)
    nr, nc, s = 300, 260, 3
    X = rng.integers(0, 3, (nr, nc)).astype(np.int8)
    A = enc_matrix(o, sk, rng.normal(size=(s, nr)))
    gfs = GenoFileStream.from_matrix(cps, X)
    full = MatMult4StreamCompute(cps, A, 5, MatMult4StreamPreprocess(cps, gfs, 5))
    acc = np.zeros(full.shape, dtype=object)
    for part in range(nparts):
        h = C.c_void_p()
        cps._check(cps.L.sfg_matmult4_stream_preprocess_giants(cps.h, gfs.h, 5, part, nparts, C.byref(h)), "preprocess_giants")
        cache = DiagCache(cps, h)
        acc = acc + MatMult4StreamCompute(cps, A, 5, cache).astype(object)
        cache.close()
    for l in range(5):
        acc[:, :, :, l] %= o.Q[l]
    assert (acc.astype(np.uint64) == full).all()


@pytest.mark.gpu
@pytest.mark.parametrize("nparts", [1, 3, 8])
def test_baby_sharded_rotation_cache_equals_compute(small13, nparts):
    """Baby-step sharding on ONE GPU: the shares of the rotation cache are computed one after the other into the same buffer (what the
    all-gather assembles across ranks), then the rest of Compute runs on it: bit-identical to MatMult4StreamCompute."""
    import ctypes as C

    import torch

    from sfgwas_b200 import GenoFileStream, MatMult4StreamCompute, MatMult4StreamPreprocess

    o, cps, sk, keys = small13
    rng = np.random.default_rng(900 + nparts)
    nr, nc, s = 300, 260, 4
    X = rng.integers(0, 3, (nr, nc)).astype(np.int8)
    A = enc_matrix(o, sk, rng.normal(size=(s, nr)))
    cache = MatMult4StreamPreprocess(cps, GenoFileStream.from_matrix(cps, X), 5)
    full = MatMult4StreamCompute(cps, A, 5, cache)
    L = cps.L
    chunk = int(L.sfg_matmult4_baby_chunk_bytes(cps.h, cache.h, s, nparts))
    R = torch.full((nparts * chunk,), 0xA5, dtype=torch.uint8, device="cuda")
    d_A = torch.from_numpy(A.view(np.int64)).cuda()
    torch.cuda.synchronize()
    for part in range(nparts):
        cps._check(L.sfg_matmult4_baby_dev(cps.h, C.c_void_p(d_A.data_ptr()), s, A.shape[1], 5, 5, cache.h, part, nparts, C.c_void_p(R.data_ptr())),
                   "baby_dev")
    d_out = torch.zeros((s, cache.m_ct, 2, 5, o.N), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    cps._check(L.sfg_matmult4_stream_compute_r_dev(cps.h, C.c_void_p(R.data_ptr()), s, 5, cache.h, C.c_void_p(d_out.data_ptr())), "compute_r_dev")
    assert (d_out.cpu().numpy().view(np.uint64) == full).all()


@pytest.mark.gpu
def test_partial_mod_reduce_finish_equals_compute(small13):
    """The multi-GPU pieces on one GPU: partial sums per block-row range, integer sum + mod q, giant ranges, modular output sum."""
    import ctypes as C

    import torch

    from sfgwas_b200 import GenoFileStream, MatMult4StreamCompute, MatMult4StreamPreprocess
    from sfgwas_b200.gwas import _p

    o, cps, sk, keys = small13
    rng = np.random.default_rng(23)
    nr, nc, s = 600, 300, 2
    X = rng.integers(0, 3, (nr, nc)).astype(np.int8)
    A = enc_matrix(o, sk, rng.normal(size=(s, nr)))
    cache = MatMult4StreamPreprocess(cps, GenoFileStream.from_matrix(cps, X), 5)
    single = MatMult4StreamCompute(cps, A, 5, cache)
    L = cps.L
    nbr = A.shape[1]
    n_cv = int(L.sfg_cv_elems(cps.h, cache.h, s, 5))
    dev = torch.device("cuda", cps.device)
    parts = []
    for lo, hi in ((0, 1), (1, nbr)):
        cv = torch.empty(n_cv, dtype=torch.int64, device=dev)
        cps._check(L.sfg_matmult4_partial(cps.h, _p(A), s, nbr, 5, 5, cache.h, lo, hi, C.c_void_p(cv.data_ptr())), "partial")
        parts.append(cv)
    cv = parts[0] + parts[1]
    torch.cuda.synchronize()
    cps._check(L.sfg_cv_mod_reduce(cps.h, cache.h, s, 5, C.c_void_p(cv.data_ptr()), 0, n_cv), "mod_reduce")
    per_g = cache.m_ct * 2 * s * 5 * cps.N
    ng = n_cv // per_g
    total = np.zeros(single.shape, dtype=object)
    for g_lo, g_hi in ((0, ng // 2), (ng // 2, ng)):
        out = np.zeros(single.shape, dtype=np.uint64)
        cps._check(L.sfg_matmult4_finish(cps.h, cache.h, s, 5, C.c_void_p(cv.data_ptr()), g_lo, g_hi, _p(out)), "finish")
        total = total + out.astype(object)
    for l in range(5):
        total[:, :, :, l, :] %= cps.Q[l]
    assert (total.astype(np.uint64) == single).all()


def test_row_restricted_caches_equal_compute(small13):
    """Block-row sharding with per-rank caches (sfg_matmult4_stream_preprocess_rows): each 'rank' holds the diagonals of its own block
    rows only; partial sums + mod q + finish equal the single-cache result bit for bit, and a restricted cache refuses what it cannot do."""
    import ctypes as C

    import torch

    from sfgwas_b200 import DiagCache, GenoFileStream, MatMult4StreamCompute, MatMult4StreamPreprocess, SfgError
    from sfgwas_b200.gwas import _p

    o, cps, sk, keys = small13
    rng = np.random.default_rng(29)
    nr, nc, s = 3 * o.slots + 17, 300, 2  # 4 block rows, the last one ragged
    X = rng.integers(0, 3, (nr, nc)).astype(np.int8)
    A = enc_matrix(o, sk, rng.normal(size=(s, nr)))
    gfs = GenoFileStream.from_matrix(cps, X)
    full = MatMult4StreamPreprocess(cps, gfs, 5)
    single = MatMult4StreamCompute(cps, A, 5, full)
    L, nbr = cps.L, A.shape[1]
    dev = torch.device("cuda", cps.device)
    n_cv = int(L.sfg_cv_elems(cps.h, full.h, s, 5))
    acc, caches = None, []
    for lo, hi in ((0, 1), (1, 3), (3, nbr)):
        h = C.c_void_p()
        cps._check(L.sfg_matmult4_stream_preprocess_rows(cps.h, gfs.h, 5, lo, hi, C.byref(h)), "preprocess_rows")
        ca = DiagCache(cps, h)
        caches.append(ca)
        assert ca.bytes < full.bytes and int(L.sfg_cv_elems(cps.h, ca.h, s, 5)) == n_cv  # smaller image, same accumulator layout
        cv = torch.empty(n_cv, dtype=torch.int64, device=dev)
        cps._check(L.sfg_matmult4_partial(cps.h, _p(A), s, nbr, 5, 5, ca.h, lo, hi, C.c_void_p(cv.data_ptr())), "partial")
        acc = cv if acc is None else acc + cv
    torch.cuda.synchronize()
    cps._check(L.sfg_cv_mod_reduce(cps.h, caches[0].h, s, 5, C.c_void_p(acc.data_ptr()), 0, n_cv), "mod_reduce")
    per_g = full.m_ct * 2 * s * 5 * cps.N
    out = np.zeros(single.shape, dtype=np.uint64)
    cps._check(L.sfg_matmult4_finish(cps.h, caches[1].h, s, 5, C.c_void_p(acc.data_ptr()), 0, n_cv // per_g, _p(out)), "finish")
    assert (out == single).all()
    with pytest.raises(SfgError, match="holds block rows"):
        MatMult4StreamCompute(cps, A, 5, caches[0])
    cv = torch.empty(n_cv, dtype=torch.int64, device=dev)
    assert L.sfg_matmult4_partial(cps.h, _p(A), s, nbr, 5, 5, caches[0].h, 0, 2, C.c_void_p(cv.data_ptr())) != 0


@pytest.mark.parametrize("nr,nc", [(300, 520), (128, 77), (700, 130)])
def test_diag_cache_files_interop(small13, tmp_path, nr, nc, monkeypatch):
    """the reference's on-disk cache format (gwas/filestream.go:19-282).  GPU-written files are byte-identical to
    the oracle's, and a cache loaded from the oracle's files (records shuffled, as the reference's goroutines may write them, and
    staged in small chunks) computes the same ciphertexts bit for bit."""
    from sfgwas_b200 import DiagCache, GenoFileStream, MatMult4StreamCompute, MatMult4StreamPreprocess, SfgError

    o, cps, sk, keys = small13
    rng = np.random.default_rng(nr * 1000 + nc)
    X = rng.integers(0, 3, (nr, nc)).astype(np.int8)
    dc = o.preprocess(X, 5, nproc=1)
    ref_prefix, gpu_prefix = str(tmp_path / "ref"), str(tmp_path / "gpu")
    o.cache_write_files(dc, ref_prefix)
    o.cache_free(dc)
    monkeypatch.setenv("SFG_CACHEFILE_CHUNK_POLYS", "7")  # exercise the chunked staging on small inputs
    cache = MatMult4StreamPreprocess(cps, GenoFileStream.from_matrix(cps, X), 5, gpu_prefix)
    nbr = (nr - 1) // o.slots + 1
    for bi in range(nbr):
        a = open("%s_%d.bin" % (ref_prefix, bi), "rb").read()
        b = open("%s_%d.bin" % (gpu_prefix, bi), "rb").read()
        assert len(a) == len(b) and a == b, "block row %d: files differ" % bi
    # shuffle the records of the reference files (header + tables stay)
    d = o.d
    for bi in range(nbr):
        raw = open("%s_%d.bin" % (ref_prefix, bi), "rb").read()
        head, pos, recs = raw[: 48 + 2 * d], 48 + 2 * d, []
        while pos < len(raw):
            n = int.from_bytes(raw[pos:pos + 8], "little")
            recs.append(raw[pos:pos + 8 + n])
            pos += 8 + n
        random.Random(bi).shuffle(recs)
        open("%s_%d.bin" % (ref_prefix, bi), "wb").write(head + b"".join(recs))
    loaded = DiagCache.load_files(cps, ref_prefix, nr, nc, 5)
    assert loaded.num_polys == cache.num_polys and loaded.materialised
    s = 2
    A = np.zeros((s, nbr, 2, 6, o.N), dtype=np.uint64)
    for i in range(s):
        for b in range(nbr):
            A[i, b] = o.encrypt_vector(sk, rng.normal(size=o.slots), 5, seed=50 + 7 * i + b)
    want = MatMult4StreamCompute(cps, A, 5, cache)
    assert (MatMult4StreamCompute(cps, A, 5, loaded) == want).all()
    # shape taken from the files alone (what the reference's MatMult4StreamCompute has in hand: only the prefix)
    inferred = DiagCache.load_files(cps, ref_prefix, 0, 0, 5)
    assert inferred.num_polys == cache.num_polys and inferred.m_ct == cache.m_ct and inferred.num_block_rows == nbr
    assert (MatMult4StreamCompute(cps, A, 5, inferred) == want).all()
    with pytest.raises(SfgError, match="open .*missing_0.bin"):
        DiagCache.load_files(cps, str(tmp_path / "missing"), nr, nc, 5)
    with pytest.raises(SfgError, match="does not match"):
        DiagCache.load_files(cps, ref_prefix, nr, nc + o.slots, 5)


@pytest.mark.parametrize("nr,nc,kp", [(300, 520, 7), (1000, 4096 + 48, 15), (77, 131, 3), (5, 16, 20)])
def test_count_sketch_bit_exact(small13, nr, nc, kp):
    """SURVEY 8f row 1 (gwas/pca.go:152-162): count sketch + column sums, exact integers; aligned (ncols % 16 == 0) and ragged widths,
    empty buckets, more buckets than rows."""
    from oracle.oracle import count_sketch
    from sfgwas_b200 import GenoFileStream, SfgError

    o, cps, sk, keys = small13
    rng = np.random.default_rng(nr + nc + kp)
    X = rng.integers(0, 3, (nr, nc)).astype(np.int8)
    ri = rng.integers(0, kp, nr).astype(np.int32)
    if kp > 2:
        ri[ri == 1] = 0  # an empty bucket
    sg = (rng.integers(0, 2, nr) * 2 - 1).astype(np.int8)
    gfs = GenoFileStream.from_matrix(cps, X)
    got = gfs.CountSketch(ri, sg, kp)
    want = count_sketch(X, ri, sg, kp)
    for g, w in zip(got, want):
        assert g.dtype == w.dtype and (g == w).all()
    Xbad = X.copy()
    Xbad[nr // 2, nc // 2] = -1  # a missing value that was not replaced
    with pytest.raises(SfgError, match="outside"):
        GenoFileStream.from_matrix(cps, Xbad).CountSketch(ri, sg, kp)
    with pytest.raises(SfgError, match="randIndex"):
        gfs.CountSketch(np.full(nr, kp, dtype=np.int32), sg, kp)


def test_count_sketch_full_size_properties(small13):
    """Config-2 shaped scan (10 000 x 100 000 int8, kp = 10): checked through size-independent properties -- the bucket rows of the sketch
    with all signs +1 sum to xsum, x2sum = xsum + 2 * (#twos), and the signed sketch equals (sum of + rows) - (sum of - rows) on a column
    sample computed by numpy."""
    from sfgwas_b200 import GenoFileStream

    o, cps, sk, keys = small13
    rng = np.random.default_rng(7)
    nr, nc, kp = 10000, 100000, 10
    X = rng.integers(0, 3, (nr, nc), dtype=np.int8)
    ri = rng.integers(0, kp, nr).astype(np.int32)
    sg = (rng.integers(0, 2, nr) * 2 - 1).astype(np.int8)
    gfs = GenoFileStream.from_matrix(cps, X)
    sk1, xs, x2, ms = gfs.CountSketch(ri, np.ones(nr, dtype=np.int8), kp, want_ms=True)
    assert (sk1.sum(axis=0) == xs.astype(np.float64)).all()
    cols = rng.choice(nc, 64, replace=False)
    Xc = X[:, cols].astype(np.int64)
    assert (xs[cols] == Xc.sum(axis=0).astype(np.uint64)).all()
    assert (x2[cols] == (Xc * Xc).sum(axis=0).astype(np.uint64)).all()
    sk2, xs2, x22, ms2 = gfs.CountSketch(ri, sg, kp, want_ms=True)
    assert (xs2 == xs).all() and (x22 == x2).all()
    for b in range(kp):
        m = ri == b
        assert (sk2[b, cols] == (Xc[m] * sg[m, None].astype(np.int64)).sum(axis=0)).all()
    print("count sketch scan: %.3f ms -> %.0f GB/s of int8 genotypes" % (ms2, nr * nc / ms2 / 1e6))


@pytest.mark.parametrize("logN", [12, 13, 14, 15, 16])
def test_sweep_full_chain_ntt_and_rotation(logN):
    """BASELINE config 3: NTT / INTT and rotation (hybrid key-switch + automorphism) at logN 12-16 with the FULL modulus chain
    ((nQ, nP) = (2,1), (6,1), (10,2), (18,3), (34,4)), top level, bit-exact against the oracle.  logN 15 and 16 run the unfused
    large-ring key-switch (a 2^15-coefficient limb does not fit one CTA's shared memory)."""
    from oracle.oracle import sweep_params

    o, cps = make(sweep_params(logN))
    rng = np.random.default_rng(logN)
    top = o.nQ - 1
    # NTT / INTT over every modulus of the chain (Q and P)
    idx = list(range(o.nQP))
    mods = o.Q + o.P
    polys = np.stack([rng.integers(0, mods[l], o.N, dtype=np.uint64) for l in idx])
    fw = cps.NTT(polys, idx)
    assert (fw == np.stack([o.ntt(l, polys[l].copy()) for l in idx])).all()
    assert (cps.NTT(fw, idx, inverse=True) == polys).all()
    # rotation by 1 and by d with random keys of the right shape, a batch of 2 ciphertexts at the top level
    sk = o.keygen_secret(5)
    cts = np.stack([np.stack([np.stack([rng.integers(0, o.Q[l], o.N, dtype=np.uint64) for l in range(top + 1)]) for _ in range(2)])
                    for _ in range(2)])
    for k in (1, o.d):
        swk = o.gen_rotation_key(sk, k)
        cps.SetRotKey(k, swk)
        got = cps.RotateRightWithEvaluator(cts, -k)
        for t in range(2):
            assert (got[t] == o.rotate_right(cts[t], -k, swk)).all(), (logN, k, t)
    cps.close()


def test_unfused_keyswitch_equals_fused_on_small_rings():
    """The large-ring (logN 15 / 16) key-switch kernels, forced onto small rings (SFG_KS_UNFUSED=1 is read once per process, hence the
    subprocess), must pass the same oracle / Python-bignum rotation tests as the fused kernels."""
    import subprocess
    import sys

    env = dict(os.environ, SFG_KS_UNFUSED="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-k",
                        "test_rotate_vs_oracle or test_rotate_vs_python_bignum_tiny"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
