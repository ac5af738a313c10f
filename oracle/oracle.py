"""ctypes binding of the CPU ORACLE (oracle/libsfg_oracle.so).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (sfgwas_b200/) never imports this.
Parity status: see oracle/sfg_oracle.h ("parity unpinned" for the Lattigo-fork parts).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsfg_oracle.so")

u64p = C.POINTER(C.c_uint64)


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "sfg_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "sfg_oracle.h"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        vp = C.c_void_p
        L.orc_ctx_new.restype = vp
        L.orc_ctx_new.argtypes = [C.c_int, u64p, C.c_int, u64p, C.c_int, C.c_double]
        L.orc_ctx_free.argtypes = [vp]
        for f in ("orc_ctx_N", "orc_ctx_nQ", "orc_ctx_nP", "orc_beta"):
            getattr(L, f).restype = C.c_int
            getattr(L, f).argtypes = [vp]
        for f in ("orc_ctx_modulus", "orc_ctx_mred", "orc_ctx_psi"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [vp, C.c_int]
        L.orc_ctx_bred.argtypes = [vp, C.c_int, u64p]
        L.orc_primitive_root.restype = C.c_uint64
        L.orc_primitive_root.argtypes = [C.c_uint64]
        L.orc_mred.restype = C.c_uint64
        L.orc_mred.argtypes = [C.c_uint64] * 4
        L.orc_mform.restype = C.c_uint64
        L.orc_mform.argtypes = [C.c_uint64, C.c_uint64, u64p]
        L.orc_bred_add.restype = C.c_uint64
        L.orc_bred_add.argtypes = [C.c_uint64, C.c_uint64, u64p]
        L.orc_ntt.argtypes = [vp, C.c_int, vp]
        L.orc_intt.argtypes = [vp, C.c_int, vp]
        L.orc_mul_coeffs_and_add128.argtypes = [vp, vp, vp, C.c_size_t]
        L.orc_reduce_and_add_uint128.argtypes = [vp, vp, C.c_uint64, C.c_uint64, C.c_size_t]
        L.orc_mform_lvl.argtypes = [vp, C.c_int, vp]
        L.orc_reduce_canonical.argtypes = [vp, C.c_int, vp]
        L.orc_get_diag_bool.restype = C.c_int
        L.orc_get_diag_bool.argtypes = [C.c_int] * 4
        L.orc_get_diag.restype = C.c_int
        L.orc_get_diag.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_encode_ntt.argtypes = [vp, vp, C.c_int, C.c_int, vp]
        L.orc_encode_coeffs.argtypes = [vp, vp, C.c_int, vp]
        L.orc_keygen_secret.argtypes = [vp, C.c_uint64, vp]
        L.orc_gen_switching_key.argtypes = [vp, vp, vp, C.c_uint64, vp]
        L.orc_galois_element.restype = C.c_uint64
        L.orc_galois_element.argtypes = [vp, C.c_int]
        L.orc_gen_rotation_key.argtypes = [vp, vp, C.c_uint64, C.c_uint64, vp]
        L.orc_encrypt_sk.argtypes = [vp, vp, vp, C.c_int, C.c_uint64, vp]
        L.orc_decrypt_coeffs.argtypes = [vp, vp, vp, C.c_int, vp]
        L.orc_permute_ntt_index.argtypes = [C.c_int, C.c_uint64, vp]
        L.orc_keyswitch.argtypes = [vp, C.c_int, vp, vp, vp, vp]
        L.orc_rotate_right.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp]
        L.orc_mul_relin.argtypes = [vp, C.c_int, vp, vp, vp, vp]
        L.orc_mul_plain.argtypes = [vp, C.c_int, vp, vp, vp]
        L.orc_rescale_once.argtypes = [vp, C.c_int, vp, vp]
        L.orc_ct_addsub.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp]
        L.orc_matmult4_stream_preprocess.restype = vp
        L.orc_matmult4_stream_preprocess.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_diag_cache_free.argtypes = [vp]
        L.orc_diag_cache_num_polys.restype = C.c_size_t
        L.orc_diag_cache_num_polys.argtypes = [vp]
        L.orc_diag_cache_mct.restype = C.c_int
        L.orc_diag_cache_mct.argtypes = [vp]
        L.orc_diag_cache_get.restype = vp
        L.orc_diag_cache_get.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        L.orc_diag_cache_write_files.restype = C.c_int
        L.orc_diag_cache_write_files.argtypes = [vp, vp, C.c_char_p]
        L.orc_matmult4_stream_compute.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, vp,
                                                  C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_matmult4_stream.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                          vp, C.c_int, vp, vp, vp]
        L.orc_bench_mac.restype = C.c_double
        L.orc_bench_mac.argtypes = [C.c_int] * 5
        _lib = L
    return _lib


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _u64arr(xs):
    arr = (C.c_uint64 * len(xs))(*[int(x) for x in xs])
    return arr


# ----------------------------------------------------------------------------------------------
# Parameter sets.  PN13QP218 / PN14QP438 are the recalled Lattigo v2 defaults (SURVEY App. B.1,
# [UNVERIFIED] -- the library always takes the chain from the caller).  TEST* are small rings with
# the same limb-width structure, used so that CPU tests run in seconds.
# ----------------------------------------------------------------------------------------------
PARAMS = {
    "PN13QP218": dict(logN=13, Q=[0x1FFFEC001, 0x3FFF4001, 0x3FFE8001, 0x40020001, 0x40038001, 0x3FFC0001],
                      P=[0x800004001], scale=float(1 << 30)),
    "PN14QP438": dict(logN=14, Q=[0x200000008001, 0x400018001, 0x3FFFD0001, 0x400060001, 0x400068001, 0x3FFF90001,
                                  0x400080001, 0x4000A8001, 0x400108001, 0x3FFEB8001],
                      P=[0x7FFFFFD8001, 0x7FFFFFC8001], scale=float(1 << 34)),
}


def _is_prime(n: int) -> bool:
    if n < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % p == 0:
            return n == p
    d, r = n - 1, 0
    while d % 2 == 0:
        d //= 2
        r += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(r - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def gen_primes(logN: int, bits: int, count: int, avoid=()) -> list[int]:
    """NTT-friendly primes q = 1 mod 2N just below/above 2^bits (alternating), deterministic."""
    twoN = 1 << (logN + 1)
    out, k = [], 1
    base = 1 << bits
    while len(out) < count:
        for cand in (base - k * twoN + 1, base + k * twoN + 1):
            if len(out) < count and cand not in avoid and cand not in out and _is_prime(cand):
                out.append(cand)
        k += 1
    return out


def small_params(logN: int, shape: str = "pn13") -> dict:
    """Small rings mirroring the limb-width structure of PN13QP218 (33+5x30 | 36) or PN14QP438 (46+9x34 | 43,43)."""
    if shape == "pn13":
        q0 = gen_primes(logN, 33, 1)
        q = q0 + gen_primes(logN, 30, 5, avoid=q0)
        p = gen_primes(logN, 36, 1, avoid=q)
        return dict(logN=logN, Q=q, P=p, scale=float(1 << 30))
    if shape == "pn14":
        q0 = gen_primes(logN, 46, 1)
        q = q0 + gen_primes(logN, 34, 7, avoid=q0)
        p = gen_primes(logN, 43, 2, avoid=q)
        return dict(logN=logN, Q=q, P=p, scale=float(1 << 34))
    if shape == "pn15":  # alpha = 3, wide limbs
        q0 = gen_primes(logN, 51, 1)
        q = q0 + gen_primes(logN, 40, 8, avoid=q0)
        p = gen_primes(logN, 50, 3, avoid=q)
        return dict(logN=logN, Q=q, P=p, scale=float(1 << 40))
    raise ValueError(shape)


def count_sketch(X: np.ndarray, randIndex, sgn, kp: int):
    """gwas/pca.go:152-162 restated: localSketch[randIndex[i]][j] += sgn[i] * float64(row[j]); xsum[j] += uint64(row[j]);
    x2sum[j] += uint64(row[j] * row[j]) (int8 product, like the Go expression)."""
    nind, nsnp = X.shape
    sketch = np.zeros((kp, nsnp), dtype=np.float64)
    xsum = np.zeros(nsnp, dtype=np.uint64)
    x2sum = np.zeros(nsnp, dtype=np.uint64)
    for i in range(nind):
        row = X[i]
        sketch[randIndex[i]] += float(sgn[i]) * row.astype(np.float64)
        xsum += row.astype(np.int64).astype(np.uint64)
        x2sum += (row * row).astype(np.int8).astype(np.int64).astype(np.uint64)
    return sketch, xsum, x2sum


def sweep_params(logN: int) -> dict:
    """Full modulus chains of the BASELINE sweep (config 3: logN 12-16), shaped like the recalled Lattigo v2 defaults PN12QP109 ..
    PN16QP1761 ([UNVERIFIED]): (nQ, nP) = (2,1), (6,1), (10,2), (18,3), (34,4).  Primes are generated here (q = 1
    mod 2N) with the default sets' bit widths; the library always takes the chain from the caller."""
    shape = {12: (37, 32, 2, 38, 1, 32), 13: (33, 30, 6, 36, 1, 30), 14: (45, 34, 10, 43, 2, 34), 15: (50, 40, 18, 50, 3, 40),
             16: (55, 45, 34, 55, 4, 45)}[logN]
    q0b, qb, nQ, pb, nP, sc = shape
    q0 = gen_primes(logN, q0b, 1)
    q = q0 + gen_primes(logN, qb, nQ - 1, avoid=q0)
    p = gen_primes(logN, pb, nP, avoid=q)
    return dict(logN=logN, Q=q, P=p, scale=float(1 << sc))


def marshal_ciphertext(value: np.ndarray, scale: float) -> bytes:
    """Lattigo v2.1 ckks.Element.MarshalBinary [UNVERIFIED vs the fork; SURVEY App. B.8]: u8 degree+1, f64 LE scale, u8 isNTT, then per
    polynomial ring.Poly.WriteTo = u8 log2(N), u8 numModuli, coefficients limb-major as 8-byte BIG-endian words (ring.WriteCoeffsTo)."""
    import struct

    deg1, nl, N = value.shape
    out = [struct.pack("<BdB", deg1, float(scale), 1)]
    for k in range(deg1):
        out.append(struct.pack("<BB", int(N).bit_length() - 1, nl))
        out.append(np.ascontiguousarray(value[k], dtype=">u8").tobytes())
    return b"".join(out)


def save_cipher_matrix(cm, filename: str):
    """crypto.SaveCipherMatrixToFile (crypto/utilities.go:82-113) with MarshalCM (:35-56) and CipherMatrix/CipherVector.MarshalBinary
    (crypto/crypto.go:542-595): cm[i][j] = (value [2][nl][N] uint64, scale)."""
    import struct

    blobs = [[marshal_ciphertext(v, sc) for (v, sc) in row] for row in cm]
    sbytes = b"".join(struct.pack("<Q", len(b)) for row in blobs for b in row)      # MarshalCM: ctSizes as little-endian u64
    cmbytes = b"".join(b for row in blobs for b in row)
    with open(filename, "wb") as f:
        f.write(struct.pack("<I", len(cm)))          # nrbuf
        f.write(struct.pack("<I", len(cm[0])))       # ncbuf
        f.write(struct.pack("<Q", len(sbytes)))      # sbuf
        f.write(sbytes)
        f.write(struct.pack("<Q", len(cmbytes)))     # cmbuf
        f.write(cmbytes)


class Oracle:
    """One CKKS ring context of the oracle (Lattigo ring.Ring + ckks.Parameters restated)."""

    def __init__(self, logN: int, Q, P, scale: float):
        self.L = lib()
        self.logN, self.N, self.slots = logN, 1 << logN, 1 << (logN - 1)
        self.Q, self.P = [int(x) for x in Q], [int(x) for x in P]
        self.nQ, self.nP = len(Q), len(P)
        self.nQP = self.nQ + self.nP
        self.scale = float(scale)
        self.ctx = self.L.orc_ctx_new(logN, _u64arr(self.Q), self.nQ, _u64arr(self.P), self.nP, self.scale)
        assert self.ctx
        self.d = int(math.ceil(math.sqrt(self.slots)))
        self.beta = self.L.orc_beta(self.ctx)
        self.moduli = self.Q + self.P

    @classmethod
    def from_params(cls, p: dict) -> "Oracle":
        return cls(p["logN"], p["Q"], p["P"], p["scale"])

    def __del__(self):
        try:
            self.L.orc_ctx_free(self.ctx)
        except Exception:
            pass

    # -- ring constants -------------------------------------------------------------------------
    def mred_param(self, i):
        return int(self.L.orc_ctx_mred(self.ctx, i))

    def bred_param(self, i):
        out = (C.c_uint64 * 2)()
        self.L.orc_ctx_bred(self.ctx, i, out)
        return [int(out[0]), int(out[1])]

    def psi(self, i):
        return int(self.L.orc_ctx_psi(self.ctx, i))

    # -- transforms -----------------------------------------------------------------------------
    def ntt(self, idx: int, a: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint64).copy()
        self.L.orc_ntt(self.ctx, idx, _p(a))
        return a

    def intt(self, idx: int, a: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint64).copy()
        self.L.orc_intt(self.ctx, idx, _p(a))
        return a

    def ntt_poly(self, p: np.ndarray, idxs=None) -> np.ndarray:
        p = np.ascontiguousarray(p, dtype=np.uint64).copy()
        idxs = range(p.shape[0]) if idxs is None else idxs
        for k, i in enumerate(idxs):
            self.L.orc_ntt(self.ctx, i, _p(p[k]))
        return p

    # -- K1/K2/K3 -------------------------------------------------------------------------------
    def mul_coeffs_and_add128(self, a, b, c):
        """c: uint64 array [n,2] holding (hi, lo) pairs (gwas/matmult.go:196-199), updated in place."""
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = np.ascontiguousarray(b, dtype=np.uint64)
        assert c.dtype == np.uint64 and c.shape == (a.shape[0], 2)
        self.L.orc_mul_coeffs_and_add128(_p(a), _p(b), _p(c), a.shape[0])

    def reduce_and_add_uint128(self, acc, out, limb):
        self.L.orc_reduce_and_add_uint128(_p(acc), _p(out), self.mred_param(limb), self.moduli[limb], acc.shape[0])

    def mform_lvl(self, level, p):
        p = np.ascontiguousarray(p, dtype=np.uint64).copy()
        self.L.orc_mform_lvl(self.ctx, level, _p(p))
        return p

    def reduce_canonical(self, p):
        p = np.ascontiguousarray(p, dtype=np.uint64).copy()
        self.L.orc_reduce_canonical(self.ctx, p.shape[0], _p(p))
        return p

    # -- diagonals / encoder -----------------------------------------------------------------------
    def get_diag(self, X: np.ndarray, index: int):
        """GetDiag(dst, X, dim=slots, index) for an r x c int8 block. Returns (ok, dst)."""
        X = np.ascontiguousarray(X, dtype=np.int8)
        dst = np.zeros(self.slots, dtype=np.float64)
        ok = self.L.orc_get_diag(_p(dst), _p(X), X.shape[1], X.shape[0], X.shape[1], self.slots, index)
        return bool(ok), dst

    def encode_coeffs(self, values, nrot=0) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.float64)
        assert v.shape == (self.slots,)
        out = np.zeros(self.N, dtype=np.int64)
        self.L.orc_encode_coeffs(self.ctx, _p(v), nrot, _p(out))
        return out

    def encode_ntt(self, values, nrot=0, level=None) -> np.ndarray:
        level = self.nQ - 1 if level is None else level
        v = np.ascontiguousarray(values, dtype=np.float64)
        out = np.zeros((level + 1, self.N), dtype=np.uint64)
        self.L.orc_encode_ntt(self.ctx, _p(v), nrot, level, _p(out))
        return out

    # -- test-only CKKS ----------------------------------------------------------------------------
    def keygen_secret(self, seed=1) -> np.ndarray:
        sk = np.zeros((self.nQP, self.N), dtype=np.uint64)
        self.L.orc_keygen_secret(self.ctx, seed, _p(sk))
        return sk

    def galois_element(self, k: int) -> int:
        return int(self.L.orc_galois_element(self.ctx, k))

    def gen_rotation_key(self, sk, k: int, seed=7) -> np.ndarray:
        """Switching key for LEFT rotation by k: [beta][2][nQP][N], NTT + Montgomery form."""
        swk = np.zeros((self.beta, 2, self.nQP, self.N), dtype=np.uint64)
        self.L.orc_gen_rotation_key(self.ctx, _p(sk), self.galois_element(k), seed, _p(swk))
        return swk

    def encrypt(self, sk, pt_ntt, level, seed=3) -> np.ndarray:
        ct = np.zeros((2, level + 1, self.N), dtype=np.uint64)
        pt = None if pt_ntt is None else np.ascontiguousarray(pt_ntt[: level + 1], dtype=np.uint64)
        self.L.orc_encrypt_sk(self.ctx, _p(sk), _p(pt) if pt is not None else None, level, seed, _p(ct))
        return ct

    def decrypt_coeffs(self, sk, ct) -> np.ndarray:
        level = ct.shape[1] - 1
        ct = np.ascontiguousarray(ct, dtype=np.uint64)
        out = np.zeros((level + 1, self.N), dtype=np.uint64)
        self.L.orc_decrypt_coeffs(self.ctx, _p(sk), _p(ct), level, _p(out))
        return out

    def crt_center(self, res: np.ndarray) -> list[int]:
        """CRT-reconstruct centred big integers from residues [nl][N] (python ints)."""
        nl = res.shape[0]
        mods = self.Q[:nl]
        Qall = 1
        for m in mods:
            Qall *= m
        coef = []
        for m in mods:
            Mi = Qall // m
            coef.append(Mi * pow(Mi, -1, m))
        out = []
        cols = [res[i].tolist() for i in range(nl)]
        half = Qall // 2
        for j in range(self.N):
            x = 0
            for i in range(nl):
                x += cols[i][j] * coef[i]
            x %= Qall
            if x > half:
                x -= Qall
            out.append(x)
        return out

    def decode(self, coeffs, scale) -> np.ndarray:
        """CKKS decode (Lattigo decode: v_j = m(zeta^(5^j)), zeta = exp(i pi / N)) in float64."""
        n = self.slots
        m = np.array([float(x) for x in coeffs], dtype=np.float64) / scale
        w = m[:n] + 1j * m[n:]
        k = np.arange(n)
        u = w * np.exp(2j * np.pi * k / (4 * n))
        V = np.fft.ifft(u) * n
        five, idx = 1, np.zeros(n, dtype=np.int64)
        for j in range(n):
            idx[j] = ((five - 1) // 4) % n
            five = five * 5 % (4 * n)
        return V[idx]

    def encrypt_vector(self, sk, values, level, seed=3, scale=None):
        """Encode (scale) + encrypt a real vector of <= slots entries."""
        v = np.zeros(self.slots)
        v[: len(values)] = values
        if scale is None or scale == self.scale:
            pt = self.encode_ntt(v, 0, level)
        else:
            raise NotImplementedError
        return self.encrypt(sk, pt, level, seed)

    def decrypt_vector(self, sk, ct, scale) -> np.ndarray:
        return self.decode(self.crt_center(self.decrypt_coeffs(sk, ct)), scale)

    # -- rotation ------------------------------------------------------------------------------------
    def permute_ntt_index(self, galEl) -> np.ndarray:
        idx = np.zeros(self.N, dtype=np.uint32)
        self.L.orc_permute_ntt_index(self.logN, galEl, _p(idx))
        return idx

    def keyswitch(self, c1, swk):
        level = c1.shape[0] - 1
        c1 = np.ascontiguousarray(c1, dtype=np.uint64)
        o0, o1 = np.zeros_like(c1), np.zeros_like(c1)
        self.L.orc_keyswitch(self.ctx, level, _p(c1), _p(swk), _p(o0), _p(o1))
        return o0, o1

    def rotate_right(self, ct, nrot, swk):
        """crypto.RotateRightWithEvaluator(ct, nrot). swk = key for LEFT rotation by (slots - nrot mod slots)."""
        ct = np.ascontiguousarray(ct, dtype=np.uint64)
        out = np.zeros_like(ct)
        self.L.orc_rotate_right(self.ctx, ct.shape[1] - 1, _p(ct), nrot, _p(swk) if swk is not None else None, _p(out))
        return out

    # -- ct x ct / ct x pt algebra of the callers (gwas/matmult.go:27-116, crypto/basics.go) ----------------
    # A ciphertext is a pair (value [2][level+1][N] uint64, scale float) like *ckks.Ciphertext's Value / Scale.
    def gen_relin_key(self, sk, seed=13) -> np.ndarray:
        """Lattigo keygen.GenRelinearizationKey: switching key s^2 -> s, [beta][2][nQP][N], NTT + Montgomery form."""
        mods = self.Q + self.P
        sk2 = np.zeros_like(sk)
        for l in range(self.nQP):
            v = sk[l].astype(object)
            sk2[l] = np.array((v * v) % mods[l], dtype=np.uint64)
        swk = np.zeros((self.beta, 2, self.nQP, self.N), dtype=np.uint64)
        self.L.orc_gen_switching_key(self.ctx, _p(sk2), _p(np.ascontiguousarray(sk)), seed, _p(swk))
        return swk

    def pow2_rotations(self):
        """Left rotations 1, 2, 4, .. < slots that InnerSumAll needs (crypto/crypto.go:232-249)."""
        return [1 << i for i in range(self.logN - 1)]

    @staticmethod
    def _at_level(ct, level):
        return np.ascontiguousarray(ct[:, : level + 1])

    def mul_relin(self, a, b, rlk):
        """evaluator.MulRelinNew(a, b): level = min level, scale = product."""
        (va, sa), (vb, sb) = a, b
        level = min(va.shape[1], vb.shape[1]) - 1
        out = np.zeros((2, level + 1, self.N), dtype=np.uint64)
        self.L.orc_mul_relin(self.ctx, level, _p(self._at_level(va, level)), _p(self._at_level(vb, level)), _p(rlk), _p(out))
        return out, sa * sb

    def mul_plain(self, pt, ct):
        """evaluator.MulRelinNew(plaintext, ct); pt = (value [nl][N], scale)."""
        (vp_, sp), (vc, sc) = pt, ct
        level = min(vp_.shape[0], vc.shape[1]) - 1
        out = np.zeros((2, level + 1, self.N), dtype=np.uint64)
        self.L.orc_mul_plain(self.ctx, level, _p(np.ascontiguousarray(vp_[: level + 1])), _p(self._at_level(vc, level)), _p(out))
        return out, sp * sc

    def rescale(self, ct, threshold=None):
        """evaluator.Rescale(ct, threshold, ct) (Lattigo v2.1): divide by the last modulus while scale >= threshold*q_level/2."""
        v, sc = ct
        threshold = self.scale if threshold is None else threshold
        while v.shape[1] - 1 > 0 and sc >= threshold * float(self.Q[v.shape[1] - 1]) / 2:
            level = v.shape[1] - 1
            out = np.zeros((2, level, self.N), dtype=np.uint64)
            self.L.orc_rescale_once(self.ctx, level, _p(np.ascontiguousarray(v)), _p(out))
            v, sc = out, sc / float(self.Q[level])
        return v, sc

    def addsub(self, a, b, sub=False):
        """evaluator.Add / Sub for operands whose scales differ by less than a factor 2 (Lattigo v2.1 evaluateInPlace: no
        scale matching unless floor(ratio) > 1): level = min, scale = max."""
        (va, sa), (vb, sb) = a, b
        r = max(sa, sb) / min(sa, sb)
        assert math.floor(r) <= 1, "scale matching by constant multiplication is not restated"
        level = min(va.shape[1], vb.shape[1]) - 1
        out = np.zeros((2, level + 1, self.N), dtype=np.uint64)
        self.L.orc_ct_addsub(self.ctx, level + 1, _p(self._at_level(va, level)), _p(self._at_level(vb, level)), int(sub), _p(out))
        return out, max(sa, sb)

    def CMult(self, X, Y, rlk):
        """crypto.CMult (crypto/basics.go:386-427): element-wise MulRelinNew + Rescale(params.Scale), broadcasting a length-1 side."""
        n = max(len(X), len(Y))
        if len(X) == 1:
            return [self.rescale(self.mul_relin(Y[i], X[0], rlk)) for i in range(len(Y))]
        if len(Y) == 1:
            return [self.rescale(self.mul_relin(X[i], Y[0], rlk)) for i in range(len(X))]
        return [self.rescale(self.mul_relin(X[i], Y[i], rlk)) for i in range(n)]

    def CMultScalar(self, X, ct, rlk):
        """crypto.CMultScalar (crypto/basics.go:553-566)."""
        return [self.rescale(self.mul_relin(x, ct, rlk)) for x in X]

    def rotate_left(self, ct, k, keys):
        v, sc = ct
        return self.rotate_right(v, self.slots - k, keys[k]), sc

    def InnerSumAll(self, X, keys):
        """crypto.InnerSumAll (crypto/basics.go:278-293) -> RotateAndAdd (:236-246): sum of the vector's ciphertexts, then
        ct += RotL_r(ct) for r = 1, 2, 4, .. < slots."""
        acc = X[0]
        for x in X[1:]:
            acc = self.addsub(x, acc)
        r = 1
        while r < self.slots:
            acc = self.addsub(self.rotate_left(acc, r, keys), acc)
            r *= 2
        return acc

    def InnerProd(self, X, Y, rlk, keys):
        return self.InnerSumAll(self.CMult(X, Y, rlk), keys)

    def MaskTrunc(self, ct, n_keep, mask_pt=None):
        """crypto.MaskTrunc (crypto/basics.go:110-127): keep the first n_keep slots.  mask_pt: NTT-domain plaintext of the 0/1 mask at
        the maximum level (the reference encodes it with the float64 encoder; here the correctly rounded encoding)."""
        if n_keep == self.slots:
            return ct
        if mask_pt is None:
            m = np.zeros(self.slots)
            m[:n_keep] = 1.0
            mask_pt = self.encode_ntt(m, 0, self.nQ - 1)
        return self.rescale(self.mul_plain((mask_pt, self.scale), ct))

    def QXLazyNormStream(self, Q, compute, bootstrap, XMean, XStdInv, numInd, rlk, keys):
        """gwas/matmult.go:27-77.  Q: list of CipherVectors; compute(QS) -> out (MatMult4StreamCompute), bootstrap(out) -> out."""
        QS = [self.CMult(q, XStdInv, rlk) for q in Q]
        out = bootstrap(compute(QS))
        QSm = [self.InnerProd(qs, XMean, rlk, keys) for qs in QS]
        res = []
        for i in range(len(QS)):
            row = []
            for j in range(len(out[i])):
                t = self.addsub(out[i][j], QSm[i], sub=True)
                n = self.slots if j < len(out[i]) - 1 else ((numInd - 1) % self.slots) + 1
                row.append(self.MaskTrunc(t, n))
            res.append(row)
        return res

    def QXtLazyNormStream(self, Q, compute, bootstrap, XMean, XStdInv, rlk, keys):
        """gwas/matmult.go:83-116."""
        out = bootstrap(compute(Q))
        res = []
        for i in range(len(out)):
            rowSum = self.InnerSumAll(Q[i], keys)
            Q1m = self.CMultScalar(XMean, rowSum, rlk)
            row = [self.addsub(out[i][j], Q1m[j], sub=True) for j in range(len(out[i]))]
            res.append(self.CMult(row, XStdInv, rlk))
        return res

    # -- hot path ------------------------------------------------------------------------------------
    def bsgs_rotations(self):
        """Left-rotation amounts the path needs: 1..d-1 and d,2d,..  (crypto/crypto.go:251-264)."""
        d = self.d
        ks = set(range(1, d))
        for g in range(1, d):
            if g * d < self.slots:
                ks.add(g * d)
        return sorted(ks)

    def gen_bsgs_keys(self, sk, seed=11) -> dict:
        return {k: self.gen_rotation_key(sk, k, seed) for k in self.bsgs_rotations()}

    def _swk_table(self, keys: dict):
        tab = (C.c_void_p * self.slots)()
        for k, v in keys.items():
            tab[k] = v.ctypes.data
        return tab

    def preprocess(self, X: np.ndarray, maxLevel=5, nproc=1, shift_lo=0, shift_hi=0):
        X = np.ascontiguousarray(X, dtype=np.int8)
        return self.L.orc_matmult4_stream_preprocess(self.ctx, _p(X), X.shape[0], X.shape[1], maxLevel, nproc, shift_lo, shift_hi)

    def cache_get(self, dc, bi, shift, bj, maxLevel=5):
        ptr = self.L.orc_diag_cache_get(dc, bi, shift, bj)
        if not ptr:
            return None
        n = (maxLevel + 1) * self.N
        return np.ctypeslib.as_array(C.cast(ptr, u64p), shape=(n,)).reshape(maxLevel + 1, self.N).copy()

    def cache_write_files(self, dc, prefix: str):
        """DiagCacheStream.WriteDiag for every active shift (gwas/filestream.go:140-234): <prefix>_<bi>.bin."""
        assert self.L.orc_diag_cache_write_files(self.ctx, dc, str(prefix).encode()) == 0

    def cache_free(self, dc):
        self.L.orc_diag_cache_free(dc)

    def compute(self, A: np.ndarray, dc, keys: dict, maxLevel=5, nproc=1, timings=False):
        """A: [s][numBlockRows][2][levelA+1][N] -> S [s][m_ct][2][maxLevel][N]."""
        A = np.ascontiguousarray(A, dtype=np.uint64)
        s, nbr, _, nlA, N = A.shape
        m_ct = self.L.orc_diag_cache_mct(dc)
        S = np.zeros((s, m_ct, 2, maxLevel, N), dtype=np.uint64)
        tab = self._swk_table(keys)
        t = [C.c_double(0), C.c_double(0), C.c_double(0)]
        self.L.orc_matmult4_stream_compute(self.ctx, _p(A), s, nbr, nlA - 1, maxLevel, dc, tab, nproc, _p(S),
                                           C.byref(t[0]), C.byref(t[1]), C.byref(t[2]))
        if timings:
            return S, [x.value for x in t]
        return S

    def matmult4_stream(self, A: np.ndarray, X: np.ndarray, keys: dict, maxLevel=5, computeSquaredSum=False, square=False,
                        nproc=1):
        A = np.ascontiguousarray(A, dtype=np.uint64)
        X = np.ascontiguousarray(X, dtype=np.int8)
        s, nbr, _, nlA, N = A.shape
        m_ct = (X.shape[1] - 1) // self.slots + 1
        assert nbr == (X.shape[0] - 1) // self.slots + 1
        S = np.zeros((s, m_ct, 2, maxLevel, N), dtype=np.uint64)
        sm = np.zeros(X.shape[1], dtype=np.float64)
        sq = np.zeros(X.shape[1], dtype=np.float64)
        tab = self._swk_table(keys)
        self.L.orc_matmult4_stream(self.ctx, _p(A), s, nlA - 1, _p(X), X.shape[0], X.shape[1], maxLevel,
                                   int(computeSquaredSum), int(square), tab, nproc, _p(S), _p(sm), _p(sq))
        if computeSquaredSum:
            return S, sm, sq
        return S, None, None
