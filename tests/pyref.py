"""Independent pure-Python (bignum) references used to pin BOTH the C oracle and the CUDA library at small sizes.

Nothing here imports the oracle or the product.  Each function restates the mathematical definition of a Lattigo v2.1
routine ([UNVERIFIED vs the fork]) or the literal code of gwas/matmult.go.
"""
import math

M64 = (1 << 64) - 1


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
    fs = prime_factors(q - 1)
    g = 2
    while True:
        g += 1
        if all(pow(g, (q - 1) // f, q) != 1 for f in fs):
            return g


def psi_for(q, logN):
    return pow(lattigo_primitive_root(q), (q - 1) // (2 << logN), q)


def ntt_naive(a, q, psi, logN):
    """out[i] = a(psi^(2*brv(i)+1))  -- O(N^2), N <= 256."""
    N = 1 << logN
    out = []
    for i in range(N):
        pt = pow(psi, 2 * brv(i, logN) + 1, q)
        acc, pw = 0, 1
        for k in range(N):
            acc += a[k] * pw
            pw = pw * pt % q
        out.append(acc % q)
    return out


def intt_naive(A, q, psi, logN):
    """inverse of ntt_naive: a_k = N^-1 * sum_i A_i * pt_i^-k."""
    N = 1 << logN
    ninv = pow(N, -1, q)
    pts_inv = [pow(pow(psi, 2 * brv(i, logN) + 1, q), -1, q) for i in range(N)]
    out = []
    for k in range(N):
        acc = 0
        for i in range(N):
            acc += A[i] * pow(pts_inv[i], k, q)
        out.append(acc * ninv % q)
    return out


def base_convert(xs, smods, t):
    """Lattigo fast exact base conversion (ring.modUpExact / Decomposer), SURVEY App. B.5: one coefficient."""
    S = 1
    for s in smods:
        S *= s
    vi = 0.0
    acc = 0
    for x, s in zip(xs, smods):
        y = x * pow(S // s, -1, s) % s
        vi += float(y) / float(s)
        acc += y * ((S // s) % t)
    v = int(vi)
    return (acc - v * (S % t)) % t


def keyswitch_ref(Q, P, logN, level, c1, swk):
    """Lattigo v2.1 switchKeysInPlace. c1: [level+1][N] NTT; swk[beta][2][nQ+nP][N] (NTT + Montgomery). -> (out0, out1)."""
    N = 1 << logN
    nQ, nP = len(Q), len(P)
    mods = list(Q) + list(P)
    psis = [psi_for(q, logN) for q in mods]
    nl = level + 1
    alpha = nP
    beta = (nl + alpha - 1) // alpha
    c2 = [intt_naive(list(c1[l]), mods[l], psis[l], logN) for l in range(nl)]
    targets = list(range(nl)) + [nQ + p for p in range(nP)]
    acc = {t: ([0] * N, [0] * N) for t in targets}
    for i in range(beta):
        st, ed = i * alpha, min((i + 1) * alpha, nl)
        smods = mods[st:ed]
        for t in targets:
            q = mods[t]
            if st <= t < ed:
                dn = list(c1[t])
            else:
                if ed - st == 1:
                    d = [x % q for x in c2[st]]
                else:
                    d = [base_convert([c2[k][x] for k in range(st, ed)], smods, q) for x in range(N)]
                dn = ntt_naive(d, q, psis[t], logN)
            rinv = pow(1 << 64, -1, q)
            for comp in range(2):
                key = swk[i][comp][t]
                a = acc[t][comp]
                for x in range(N):
                    a[x] = (a[x] + dn[x] * (int(key[x]) * rinv % q)) % q
    outs = []
    for comp in range(2):
        tp = [intt_naive(acc[nQ + p][comp], mods[nQ + p], psis[nQ + p], logN) for p in range(nP)]
        res = []
        for l in range(nl):
            q = mods[l]
            ext = [base_convert([tp[p][x] for p in range(nP)], list(P), q) for x in range(N)]
            extn = ntt_naive(ext, q, psis[l], logN)
            Pm = 1
            for p in P:
                Pm = Pm * p % q
            pinv = pow(Pm, -1, q)
            res.append([(acc[l][comp][x] - extn[x]) * pinv % q for x in range(N)])
        outs.append(res)
    return outs[0], outs[1]


def galois_element(logN, k):
    return pow(5, k & ((2 << logN) - 1), 2 << logN)


def permute_index(logN, galEl):
    N = 1 << logN
    mask = 2 * N - 1
    return [brv((((galEl * (2 * brv(i, logN) + 1)) & mask) - 1) >> 1, logN) for i in range(N)]


def rotate_right_ref(Q, P, logN, level, ct, nrot, swk):
    """crypto.RotateRightWithEvaluator -> RotateNew(ct, slots - nrot) -> permuteNTT (KS, + c0, permute)."""
    N = 1 << logN
    slots = N // 2
    nrot %= slots
    if nrot == 0:
        return [[list(map(int, ct[c][l])) for l in range(level + 1)] for c in range(2)]
    galEl = galois_element(logN, slots - nrot)
    o0, o1 = keyswitch_ref(Q, P, logN, level, [list(map(int, ct[1][l])) for l in range(level + 1)], swk)
    idx = permute_index(logN, galEl)
    out = [[None] * (level + 1), [None] * (level + 1)]
    for l in range(level + 1):
        q = Q[l]
        t0 = [(o0[l][x] + int(ct[0][l][x])) % q for x in range(N)]
        out[0][l] = [t0[idx[j]] for j in range(N)]
        out[1][l] = [o1[l][idx[j]] for j in range(N)]
    return out


def slot_matmult_sim(Aplain, X, slots, d):
    """Slot-domain simulation of the BSGS diagonal algebra of gwas/matmult.go:1043-1236 (no encryption):
    returns out[s][m_ct*slots] using only rotations / elementwise products of slot vectors."""
    import numpy as np

    s, nrows = Aplain.shape
    ncols = X.shape[1]
    nbr, m_ct = (nrows - 1) // slots + 1, (ncols - 1) // slots + 1
    out = np.zeros((s, m_ct * slots))
    for i in range(s):
        acc = {}
        for bi in range(nbr):
            a = np.zeros(slots)
            seg = Aplain[i, bi * slots:(bi + 1) * slots]
            a[: len(seg)] = seg
            r = min((bi + 1) * slots, nrows) - bi * slots
            for shift in range(slots):
                baby, giant = shift % d, shift // d
                rot = np.roll(a, -baby)  # RotateRight(-baby) = left rotation by baby
                for bj in range(m_ct):
                    c = min((bj + 1) * slots, ncols) - bj * slots
                    index = (slots - shift) % slots
                    if not ((slots + 1 - r) <= index or index <= c - 1):
                        continue
                    diag = np.zeros(slots)
                    for j in range(slots):
                        row = (shift + j) % slots
                        if row < r and j < c:
                            diag[j] = X[bi * slots + row, bj * slots + j]
                    pt = np.roll(diag, d * giant)  # right rotation by d*giant before encoding
                    acc.setdefault((giant, bj), np.zeros(slots))
                    acc[(giant, bj)] += rot * pt
        for (giant, bj), v in acc.items():
            out[i, bj * slots:(bj + 1) * slots] += np.roll(v, -giant * d)
    return out
