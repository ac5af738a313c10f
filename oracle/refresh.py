"""CPU oracle (TEST INFRASTRUCTURE ONLY) for SURVEY 8f row 4: the LOCAL arithmetic of the collective bootstrap, i.e. what every party
computes per ciphertext around the two network aggregations of ``mpc/mhe.go:262-341`` (CollectiveBootstrap / CollectiveBootstrapMat):

    refProtocol.GenShares(skShard, levelStart, nParties-1, ct, scale, crp, share1, share2)      mhe.go:303-311
    ... AggregateRefreshShare(Mat) over the network (stays in Go) ...                              mhe.go:313-314
    refProtocol.Decrypt(ct, agg1); refProtocol.Recode(ct, scale); refProtocol.Recrypt(ct, crp, agg2)   mhe.go:316-318

The protocol itself lives in the un-vendored Lattigo fork (``dckks/refresh.go`` of github.com/hcholab/lattigo/v2 @ e8d68c24b94a, go.mod:5,12).
It is restated here from the published v2.1 algorithm -- **[UNVERIFIED vs the fork]**, parity unpinned like the rest of the Lattigo pieces:

  GenShares : mask_k uniform in [-B/2, B/2), B = Q_level / (2 nParties)  (drawn by the CALLER: crypto/rand in Go)
              h0 = NTT(mask mod Q_level) + sk * c1 + NTT(e0)                       (level+1 limbs)
              mask' = Quo(mask * floor(targetScale), floor(ct.Scale))               ("scales the mask by the ratio between the two scales")
              h1 = -( NTT(mask' mod Q_max) + sk * a + NTT(e1) )                    (all nQ limbs; a = the common reference polynomial)
  Decrypt   : c0 += agg(h0)                                                         (level+1 limbs)
  Recode    : x = centred CRT reconstruction of INTT(c0) mod Q_level;  x = Quo(x * floor(targetScale), floor(ct.Scale))  (big.Int.Quo:
              truncation toward zero);  c0 = NTT(x mod q_j) for ALL nQ limbs (big.Int.Mod: non-negative)
  Recrypt   : c0 += agg(h1);  c1 = a

Pure Python big integers + the C oracle's NTT: only for small cases in the tests."""
from __future__ import annotations

import numpy as np


def _prod(xs):
    p = 1
    for x in xs:
        p *= int(x)
    return p


def residues(o, vals, nl):
    """ring.SetCoefficientsBigintLvl: big.Int.Mod (Euclidean, non-negative) of every coefficient into limbs 0..nl-1 -> [nl][N] uint64."""
    out = np.zeros((nl, o.N), dtype=np.uint64)
    for l in range(nl):
        q = int(o.Q[l])
        out[l] = np.array([int(v) % q for v in vals], dtype=np.uint64)
    return out


def scale_quo(x: int, so: int, si: int) -> int:
    """big.Int: x.Mul(x, so); x.Quo(x, si) -- Quo truncates toward zero."""
    y = x * so
    z = abs(y) // si
    return -z if y < 0 else z


def gen_shares(o, level, sk_ntt, c1, crp, mask, e0, e1, in_scale=None, out_scale=None):
    """-> (shareDecrypt [level+1][N], shareRecrypt [nQ][N]).  sk_ntt: plain NTT-domain secret key [nQ+nP][N] (the oracle's form; Lattigo
    stores it in Montgomery form and multiplies with MRed, which gives the same plain product); mask: N Python ints; e0 / e1: N small ints."""
    nl, nQ = level + 1, o.nQ
    si = int(o.scale if in_scale is None else in_scale)
    so = int(o.scale if out_scale is None else out_scale)
    h0 = residues(o, mask, nl)
    h1 = residues(o, [scale_quo(int(v), so, si) for v in mask], nQ)
    n0 = residues(o, e0, nl)
    n1 = residues(o, e1, nQ)
    for l in range(nQ):
        q = int(o.Q[l])
        if l < nl:
            t = o.ntt(l, h0[l]).astype(object) + sk_ntt[l].astype(object) * c1[l].astype(object) + o.ntt(l, n0[l]).astype(object)
            h0[l] = (t % q).astype(np.uint64)
        t = o.ntt(l, h1[l]).astype(object) + sk_ntt[l].astype(object) * crp[l].astype(object) + o.ntt(l, n1[l]).astype(object)
        h1[l] = ((-t) % q).astype(np.uint64)
    return h0, h1


def finish(o, level, c0, in_scale, out_scale, agg0, agg1, crp):
    """Decrypt + Recode + Recrypt -> ct' [2][nQ][N] at the top level with scale out_scale."""
    nl, nQ, N = level + 1, o.nQ, o.N
    coeff = np.zeros((nl, N), dtype=np.uint64)
    for l in range(nl):
        q = int(o.Q[l])
        coeff[l] = o.intt(l, ((c0[l].astype(object) + agg0[l].astype(object)) % q).astype(np.uint64))
    Qs = [int(x) for x in o.Q[:nl]]
    Q = _prod(Qs)
    qhalf = Q >> 1
    crt = [(Q // q) * pow(Q // q, -1, q) for q in Qs]
    si, so = int(in_scale), int(out_scale)  # big.Float.Int: truncation of the float64 scales
    vals = []
    for j in range(N):
        x = sum(int(coeff[l, j]) * crt[l] for l in range(nl)) % Q
        if x >= qhalf:  # sign == 1 || sign == 0
            x -= Q
        vals.append(scale_quo(x, so, si))
    out = np.zeros((2, nQ, N), dtype=np.uint64)
    r = residues(o, vals, nQ)
    for l in range(nQ):
        q = int(o.Q[l])
        out[0, l] = ((o.ntt(l, r[l]).astype(object) + agg1[l].astype(object)) % q).astype(np.uint64)
        out[1, l] = crp[l]
    return out
