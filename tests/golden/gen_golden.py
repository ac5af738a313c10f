#!/usr/bin/env python3
"""Generate tests/golden/*.json -- INDEPENDENT pure-Python (bignum / mpmath) known-answer vectors.

The reference (hhcho/sfgwas) ships no tests or golden vectors and cannot be compiled here (no Go toolchain, Lattigo
fork not vendored), so these vectors are derived from
  (a) the closed-form integer functions written in the reference itself (gwas/matmult.go:247-324, 433-440, 627-672), and
  (b) the mathematical definitions of the Lattigo pieces (NTT = evaluation at psi^(2*brv(i)+1) with Lattigo's root rule,
      encode = correctly rounded canonical-embedding inverse, hybrid key-switch with the float64 quotient estimate).
Nothing here calls the C oracle or the CUDA library: the vectors pin BOTH.

Run:  python tests/golden/gen_golden.py      (a few minutes; needs mpmath)
"""
import hashlib
import json
import math
import os
import random

import mpmath

HERE = os.path.dirname(os.path.abspath(__file__))
M64 = (1 << 64) - 1
M128 = (1 << 128) - 1


def brv(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


def prime_factors(n):
    fs, f = [], 2
    while f * f <= n:
        if n % f == 0:
            fs.append(f)
            while n % f == 0:
                n //= f
        f += 1 if f == 2 else 2
    if n > 1:
        fs.append(n)
    return fs


def lattigo_primitive_root(q):
    """ring.primitiveRoot: g = 2; loop { g++; test } -> first candidate is 3."""
    fs = prime_factors(q - 1)
    g = 2
    while True:
        g += 1
        if all(pow(g, (q - 1) // f, q) != 1 for f in fs):
            return g


def psi_for(q, N):
    return pow(lattigo_primitive_root(q), (q - 1) // (2 * N), q)


def ntt_ref(a, q, psi, logN):
    """Independent O(N log N) NTT from the definition: out[i] = a(psi^(2*brv(i)+1)).
    Recursive even/odd split: a(x) = e(x^2) + x*o(x^2)."""
    N = 1 << logN

    def evaluate(coeffs, roots):  # roots: list of points; len(roots) == len(coeffs); points come in +/- pairs
        n = len(coeffs)
        if n == 1:
            return [coeffs[0] % q] * len(roots)
        half = len(roots) // 2
        # roots arranged so that roots[k + half] = -roots[k]
        sq = [r * r % q for r in roots[:half]]
        e = evaluate(coeffs[0::2], sq)
        o = evaluate(coeffs[1::2], sq)
        out = [0] * len(roots)
        for k in range(half):
            t = roots[k] * o[k] % q
            out[k] = (e[k] + t) % q
            out[k + half] = (e[k] - t) % q
        return out

    # natural-order points psi^(2j+1), j < N, arranged in +/- pairs: psi^(2(j+N/2)+1) = -psi^(2j+1)
    def build_points(n_pts, base, step):
        # points base*step^j for j < n_pts; second half is the negation of the first when step^(n_pts/2) = -1
        pts, cur = [], base
        for _ in range(n_pts):
            pts.append(cur)
            cur = cur * step % q
        return pts

    pts = build_points(N, psi, psi * psi % q)
    # recursion needs sq-roots to again be in +/- pair order: squares of psi^(2j+1), j<N/2 are (psi^2)^(2j+1): same shape
    vals = evaluate([x % q for x in a], pts)  # vals[j] = a(psi^(2j+1))
    return [vals[brv(i, logN)] for i in range(N)]


def sha(xs):
    h = hashlib.sha256()
    for x in xs:
        h.update(int(x).to_bytes(8, "little"))
    return h.hexdigest()


def mform(a, q):
    """gwas/matmult.go:433-440 evaluated literally with u = floor(2^128/q) = {hi, lo}."""
    u = (1 << 128) // q
    uhi, ulo = u >> 64, u & M64
    mhi = (a * ulo) >> 64
    r = ((-((a * uhi + mhi) & M64)) & M64) * q & M64
    if r >= q:
        r -= q
    return r


def gen_primitives(rng):
    moduli = [0x3FFF4001, 0x40020001, 0x1FFFEC001, 0x800004001, 0x400018001, 0x200000008001, 0x7FFFFFD8001,
              0x4000000120001, 0x80000000080001]
    out = {"moduli": moduli, "cases": []}
    for q in moduli:
        qinv = pow(q, -1, 1 << 64)
        n = 16
        a = [rng.randrange(q) for _ in range(n)]
        b = [rng.randrange(q) for _ in range(n)]
        acc0 = [rng.randrange(1 << 128) for _ in range(n)]
        # force wrap-around in a few lanes
        acc0[0] = M128
        acc0[1] = M128 - 5
        # K1 literal (gwas/matmult.go:264-266): lo, carry = Add64(z.lo, lo); hi += hi + carry   (wraps mod 2^128)
        acc1 = [(x + ai * bi) & M128 for x, ai, bi in zip(acc0, a, b)]
        # K2 literal (gwas/matmult.go:300-301): out += in.hi - hi64((in.lo*qInv mod 2^64)*q) + q   (u64 wrap)
        out0 = [rng.randrange(q) for _ in range(n)]
        out1 = []
        for x, o in zip(acc1, out0):
            hi, lo = x >> 64, x & M64
            hhi = (((lo * qinv) & M64) * q) >> 64
            out1.append((o + hi - hhi + q) & M64)
        mf = [mform(x, q) for x in a]
        assert all(m == x * (1 << 64) % q for m, x in zip(mf, a))
        # lazy-MAC end to end: K products with MForm'ed plaintext, reduce, canonical
        K = 200
        aa = [rng.randrange(q) for _ in range(K)]
        bb = [rng.randrange(q) for _ in range(K)]
        acc = 0
        for x, y in zip(aa, bb):
            acc = (acc + x * mform(y, q)) & M128
        hi, lo = acc >> 64, acc & M64
        hhi = (((lo * qinv) & M64) * q) >> 64
        y = (hi - hhi + q) & M64
        lazy = y % q
        assert lazy == sum(x * y_ for x, y_ in zip(aa, bb)) % q
        out["cases"].append(dict(q=q, qinv=qinv, bred=[((1 << 128) // q) >> 64, ((1 << 128) // q) & M64], a=a, b=b,
                                 acc_in=[[x >> 64, x & M64] for x in acc0], acc_out=[[x >> 64, x & M64] for x in acc1],
                                 red_in=out0, red_out=out1, mform=mf, lazy_a=aa, lazy_b=bb, lazy_result=lazy))
    return out


def gen_ntt(rng):
    out = {"cases": []}
    sets = [(6, [0x1FFFEC001, 0x3FFF4001, 0x800004001]), (10, [0x40020001, 0x200000008001]),
            (13, [0x1FFFEC001, 0x3FFF4001, 0x3FFE8001, 0x40020001, 0x40038001, 0x3FFC0001, 0x800004001]),
            (14, [0x200000008001, 0x400018001, 0x7FFFFFD8001])]
    for logN, qs in sets:
        N = 1 << logN
        for q in qs:
            assert (q - 1) % (2 * N) == 0
            g = lattigo_primitive_root(q)
            psi = pow(g, (q - 1) // (2 * N), q)
            seed = rng.randrange(1 << 30)
            r2 = random.Random(seed)
            a = [r2.randrange(q) for _ in range(N)]
            A = ntt_ref(a, q, psi, logN)
            # spot-check against the naive definition
            for i in (0, 1, N // 2 + 3, N - 1):
                pt = pow(psi, 2 * brv(i, logN) + 1, q)
                if N <= 1024:
                    assert A[i] == sum(c * pow(pt, k, q) for k, c in enumerate(a)) % q
            case = dict(logN=logN, q=q, g=g, psi=psi, seed=seed, first=A[:8], sha256=sha(A))
            if N <= 64:
                case["input"] = a
                case["output"] = A
            out["cases"].append(case)
    return out


def special_invfft_exact(v, N, prec=200):
    """w_k = (1/n) sum_j v_j * exp(-2 pi i 5^j k / 2N)  (inverse of Lattigo's decode), mpmath."""
    mpmath.mp.prec = prec
    n = N // 2
    M = 2 * N
    rot = [pow(5, j, M) for j in range(n)]
    res = []
    for k in range(n):
        re = mpmath.mpf(0)
        im = mpmath.mpf(0)
        for j in range(n):
            if v[j] == 0:
                continue
            t = (rot[j] * k) % M
            ang = 2 * mpmath.pi * t / M
            re += v[j] * mpmath.cos(ang)
            im -= v[j] * mpmath.sin(ang)
        res.append((re / n, im / n))
    return res


def round_away(x):
    return int(mpmath.floor(x + mpmath.mpf(1) / 2)) if x >= 0 else -int(mpmath.floor(-x + mpmath.mpf(1) / 2))


def gen_encode(rng):
    out = {"cases": []}
    for logN, scale in ((6, 2.0 ** 30), (8, 2.0 ** 30), (8, 2.0 ** 34)):
        N, n = 1 << logN, 1 << (logN - 1)
        for nrot in (0, 5):
            v = [rng.choice([0, 1, 2, 2, 4]) for _ in range(n)]
            vr = [0] * n
            for i in range(n):
                vr[(i + nrot) % n] = v[i]  # gwas/matmult.go:666-672
            w = special_invfft_exact(vr, N)
            coeffs = [round_away(re * scale) for re, _ in w] + [round_away(im * scale) for _, im in w]
            out["cases"].append(dict(logN=logN, scale=scale, nrot=nrot, values=v, coeffs=coeffs))
    # a few coefficients at the real sizes (direct sums, 200-bit)
    for logN, scale in ((13, 2.0 ** 30), (14, 2.0 ** 34)):
        N, n, M = 1 << logN, 1 << (logN - 1), 2 << logN
        seed = rng.randrange(1 << 30)
        r2 = random.Random(seed)
        v = [r2.choice([0, 1, 2]) for _ in range(n)]
        rot = [pow(5, j, M) for j in range(n)]
        mpmath.mp.prec = 200
        picks = [0, 1, 2, n - 1, n, n + 1, N - 1] + [rng.randrange(N) for _ in range(9)]
        vals = []
        for k in picks:
            kk = k % n
            acc = mpmath.mpf(0)
            for j in range(n):
                if v[j]:
                    ang = 2 * mpmath.pi * ((rot[j] * kk) % M) / M
                    acc += v[j] * (mpmath.cos(ang) if k < n else -mpmath.sin(ang))
            vals.append(round_away(acc / n * scale))
        out["cases"].append(dict(logN=logN, scale=scale, nrot=0, seed=seed, picks=picks, pick_coeffs=vals))
    return out


def gen_diag(rng):
    """GetDiag / GetDiagBool (gwas/matmult.go:627-664) evaluated literally."""
    out = {"cases": []}
    dim = 16
    for r, c in ((16, 16), (5, 16), (16, 3), (7, 9), (1, 1), (1, 16), (16, 1)):
        X = [[rng.randrange(3) for _ in range(c)] for _ in range(r)]
        diags = {}
        for index in range(-dim + 1, dim):
            idx = index % dim
            ok = (dim + 1 - r) <= idx or idx <= c - 1
            if ok:
                i = (-idx) % dim
                dst = []
                for j in range(dim):
                    dst.append(X[i][j] if (i < r and j < c) else 0)
                    i = (i + 1) % dim
                diags[str(index)] = dst
            else:
                diags[str(index)] = None
        out["cases"].append(dict(dim=dim, r=r, c=c, X=X, diags=diags))
    return out


def main():
    rng = random.Random(20261017)
    for name, fn in (("primitives", gen_primitives), ("ntt", gen_ntt), ("diag", gen_diag), ("encode", gen_encode)):
        data = fn(rng)
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(data, f)
        print("wrote", name)


if __name__ == "__main__":
    main()
