"""Host-side mirror of the reference's Go API for the MatMult hot path, on top of the C ABI.

The reference boundary is the exported Go functions of package ``gwas`` and the slice types of package ``crypto``
(SURVEY.md 8b).  Go is not installed in this image, so the same names, argument meaning and error behaviour are
mirrored here in Python over ctypes for tests and benchmarks; the cgo shim a maintainer would add is in
``go/gwas/matmult_b200.go`` / INTEGRATION.md.  All arithmetic happens in libsfgwas_b200.so on the GPU.

Types (numpy, uint64, limb-major exactly like Lattigo's ``Poly.Coeffs[l][j]``):
    CipherMatrix  -> ndarray [s][n_ct][2][level+1][N]      (crypto/crypto.go:32-42)
    PlainVector   -> ndarray [n_ct][level+1][N]
Failures raise ``SfgError`` (the reference panics / log.Fatal's, gwas/matmult.go:360-362).
"""
from __future__ import annotations

import ctypes as C
import math
import weakref

import numpy as np

from ._lib import SfgError, load

__all__ = [
    "DeviceCipherVector", "QXtLazyNormStreamDevice",
    "SaveCipherMatrixToFile", "LoadCipherMatrixFromFile", "RefreshGenShares", "RefreshFinish",
    "CryptoParams", "GenoFileStream", "DiagCache", "MatMult4StreamPreprocess", "MatMult4StreamCompute", "MatMult4Stream",
    "SfgError", "Ciphertext", "SetRelinKey", "CMult", "CMultScalar", "CSub", "CAdd", "InnerSumAll", "InnerProd", "MaskTrunc",
    "QXLazyNormStream", "QXtLazyNormStream",
]


def _p(a: np.ndarray):
    if not a.flags["C_CONTIGUOUS"]:
        raise SfgError("array must be C-contiguous")
    return a.ctypes.data_as(C.c_void_p)


class CryptoParams:
    """The subset of crypto.CryptoParams the path reads: Params (moduli, scale) and RotKs (crypto/crypto.go:45-60)."""

    def __init__(self, logN: int, Q, P, scale: float, device: int = 0, psi=None):
        self.L = load()
        self.logN, self.N, self.slots = int(logN), 1 << logN, 1 << (logN - 1)
        self.Q, self.P = [int(x) for x in Q], [int(x) for x in P]
        self.nQ, self.nP = len(self.Q), len(self.P)
        self.nQP = self.nQ + self.nP
        self.beta = (self.nQ + self.nP - 1) // self.nP
        self.scale = float(scale)
        self.d = int(math.ceil(math.sqrt(self.slots)))
        self.device = device
        h = C.c_void_p()
        q = (C.c_uint64 * self.nQ)(*self.Q)
        p = (C.c_uint64 * self.nP)(*self.P)
        ps = None if psi is None else (C.c_uint64 * self.nQP)(*[int(x) for x in psi])
        rc = self.L.sfg_ctx_create(device, self.logN, q, self.nQ, p, self.nP, self.scale, ps, C.byref(h))
        if rc != 0:
            raise SfgError("sfg_ctx_create: " + self.L.sfg_last_error(None).decode())
        self.h = h
        self._children = weakref.WeakSet()

    def close(self):
        if getattr(self, "h", None):
            for ch in list(self._children):  # handles that live on this context go first
                ch.close()
            self.L.sfg_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise SfgError(f"{what}: {self.L.sfg_last_error(self.h).decode()}")

    def GetSlots(self) -> int:  # crypto/crypto.go:281-284
        return self.slots

    # -- keys ------------------------------------------------------------------------------------
    def SetRotKey(self, rot_left: int, key: np.ndarray):
        """cryptoParams.RotKs.Keys[galEl] for the left rotation by rot_left: [beta][2][nQ+nP][N], NTT + Montgomery."""
        key = np.ascontiguousarray(key, dtype=np.uint64)
        if key.shape != (self.beta, 2, self.nQP, self.N):
            raise SfgError(f"rotation key shape {key.shape} != {(self.beta, 2, self.nQP, self.N)}")
        self._check(self.L.sfg_ctx_set_rotation_key(self.h, rot_left, _p(key)), "sfg_ctx_set_rotation_key")

    def SetRotKeys(self, keys: dict):
        for k, v in keys.items():
            self.SetRotKey(k, v)

    def psi(self):
        out = (C.c_uint64 * self.nQP)()
        self.L.sfg_ctx_psi(self.h, out)
        return [int(x) for x in out]

    def set_cache_budget(self, nbytes: int):
        self.L.sfg_ctx_set_cache_budget(self.h, int(nbytes))

    def launch_count(self) -> int:
        return int(self.L.sfg_ctx_launch_count(self.h))

    def encoder_stats(self):
        out = (C.c_ulonglong * 2)()
        self._check(self.L.sfg_ctx_encoder_stats(self.h, out), "sfg_ctx_encoder_stats")
        return int(out[0]), int(out[1])

    def last_timings(self):
        out = (C.c_float * 5)()
        self.L.sfg_ctx_last_timings(self.h, out)
        return dict(baby_ms=out[0], mac_ms=out[1], giant_ms=out[2], total_ms=out[3], mac_kernel_ms=out[4])

    # -- lattice primitives (parity-test surface) ----------------------------------------------------
    def NTT(self, polys: np.ndarray, limb_idx, inverse=False) -> np.ndarray:
        """ring.NTT / ring.InvNTT on polys [..., len(limb_idx), N]; modulus of poly p is limb_idx[p % len]."""
        a = np.ascontiguousarray(polys, dtype=np.uint64).copy()
        npoly = a.size // self.N
        idx = (C.c_int * len(limb_idx))(*limb_idx)
        self._check(self.L.sfg_ntt(self.h, _p(a), npoly, idx, len(limb_idx), int(inverse)), "sfg_ntt")
        return a

    def MulCoeffsAndAdd128(self, a, b, acc):
        """gwas/matmult.go:247-289; acc [n][2] = (hi, lo), updated in place."""
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = np.ascontiguousarray(b, dtype=np.uint64)
        self._check(self.L.sfg_mul_coeffs_and_add128(self.h, _p(a), _p(b), _p(acc), a.shape[0]), "sfg_mul_coeffs_and_add128")

    def ReduceAndAddUint128(self, acc, out, limb):
        """gwas/matmult.go:291-324; out updated in place."""
        self._check(self.L.sfg_reduce_and_add_uint128(self.h, _p(acc), _p(out), limb, out.shape[0]), "sfg_reduce_and_add_uint128")

    def MFormLvl(self, level, p) -> np.ndarray:
        p = np.ascontiguousarray(p, dtype=np.uint64).copy()
        self._check(self.L.sfg_mform_lvl(self.h, level, _p(p)), "sfg_mform_lvl")
        return p

    def RotateRightWithEvaluator(self, cts: np.ndarray, nrot: int) -> np.ndarray:
        """crypto/basics.go:201-210 on cts [nct][2][level+1][N] (or a single [2][level+1][N])."""
        a = np.ascontiguousarray(cts, dtype=np.uint64)
        single = a.ndim == 3
        if single:
            a = a[None]
        out = np.zeros_like(a)
        self._check(self.L.sfg_rotate_right(self.h, a.shape[2] - 1, _p(a), a.shape[0], int(nrot), _p(out)), "sfg_rotate_right")
        return out[0] if single else out


class GenoFileStream:
    """What gwas.GenoFileStream (gwas/filestream.go:284-494) delivers -- int8 rows, filters already applied -- kept in HBM."""

    def __init__(self, cps: CryptoParams, nrows: int, ncols: int):
        self.cps, self.nrows, self.ncols = cps, int(nrows), int(ncols)
        h = C.c_void_p()
        cps._check(cps.L.sfg_geno_create(cps.h, self.nrows, self.ncols, C.byref(h)), "sfg_geno_create")
        self.h = h
        cps._children.add(self)

    @classmethod
    def from_matrix(cls, cps: CryptoParams, X: np.ndarray, chunk_rows: int = 1 << 14) -> "GenoFileStream":
        X = np.ascontiguousarray(X, dtype=np.int8)
        g = cls(cps, X.shape[0], X.shape[1])
        for r0 in range(0, X.shape[0], chunk_rows):
            g.push_rows(X[r0 : r0 + chunk_rows])
        return g

    def push_rows(self, rows: np.ndarray):
        rows = np.ascontiguousarray(rows, dtype=np.int8)
        if rows.ndim != 2 or rows.shape[1] != self.ncols:
            raise SfgError("rows must be [k][ncols] int8")
        self.cps._check(self.cps.L.sfg_geno_push_rows(self.h, _p(rows), rows.shape[0]), "sfg_geno_push_rows")

    def NumRows(self):
        return self.nrows

    def NumCols(self):
        return self.ncols

    def EncodeDiag(self, block_row: int, shift: int, nrot: int, level: int, mont: bool = True, want_coeffs: bool = False):
        """EncodeDiagWithEncoder(blockVec, -shift, nrot, level) (+ ToMontgomeryForm). Returns (pv [m_ct][level+1][N], present)."""
        cps = self.cps
        m_ct = (self.ncols - 1) // cps.slots + 1
        out = np.zeros((m_ct, level + 1, cps.N), dtype=np.uint64)
        present = np.zeros(m_ct, dtype=np.uint8)
        co = np.zeros((m_ct, cps.N), dtype=np.int64) if want_coeffs else None
        cps._check(cps.L.sfg_encode_diag(cps.h, self.h, block_row, shift, nrot, level, int(mont), _p(out), _p(present),
                                         _p(co) if want_coeffs else None), "sfg_encode_diag")
        if want_coeffs:
            return out, present.astype(bool), co
        return out, present.astype(bool)

    def CountSketch(self, randIndex, sgn, kp: int, want_ms: bool = False):
        """The sketching loop of gwas/pca.go:152-162 over all rows: returns (localSketch [kp][ncols] float64, xsum, x2sum [ncols] uint64).
        ``randIndex[i]`` in [0, kp) and ``sgn[i]`` = +-1 are the caller's PRG draws (pca.go:129-131)."""
        ri = np.ascontiguousarray(randIndex, dtype=np.int32)
        sg = np.ascontiguousarray(sgn, dtype=np.int8)
        if ri.shape != (self.nrows,) or sg.shape != (self.nrows,):
            raise SfgError("randIndex / sgn must have one entry per row")
        sk = np.zeros((kp, self.ncols), dtype=np.float64)
        xs = np.zeros(self.ncols, dtype=np.uint64)
        x2 = np.zeros(self.ncols, dtype=np.uint64)
        ms = C.c_float()
        self.cps._check(self.cps.L.sfg_geno_count_sketch(self.cps.h, self.h, _p(ri), _p(sg), int(kp), _p(sk), _p(xs), _p(x2), C.byref(ms)),
                        "sfg_geno_count_sketch")
        return (sk, xs, x2, ms.value) if want_ms else (sk, xs, x2)

    def close(self):
        if getattr(self, "h", None):
            self.cps.L.sfg_geno_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DiagCache:
    """Device-resident replacement of the DiagCacheStream files written by MatMult4StreamPreprocess."""

    def __init__(self, cps: CryptoParams, h):
        self.cps, self.h = cps, h
        cps._children.add(self)
        n, b, m, mc, nb = C.c_size_t(), C.c_size_t(), C.c_int(), C.c_int(), C.c_int()
        cps.L.sfg_cache_info(h, C.byref(n), C.byref(b), C.byref(m), C.byref(mc), C.byref(nb))
        self.num_polys, self.bytes, self.materialised, self.m_ct, self.num_block_rows = n.value, b.value, bool(m.value), mc.value, nb.value

    def get_diag(self, bi, shift, bj, max_level=5):
        out = np.zeros((max_level, self.cps.N), dtype=np.uint64)
        pr = C.c_int()
        self.cps._check(self.cps.L.sfg_cache_get_diag(self.cps.h, self.h, bi, shift, bj, _p(out), C.byref(pr)), "sfg_cache_get_diag")
        return out if pr.value else None

    def write_files(self, cacheFilePrefix: str):
        """Write ``<prefix>_<bi>.bin`` in the reference's DiagCacheStream format (gwas/filestream.go:19-282), i.e. what the reference's
        MatMult4StreamPreprocess leaves on disk for MatMult4StreamCompute."""
        self.cps._check(self.cps.L.sfg_cache_write_files(self.cps.h, self.h, str(cacheFilePrefix).encode()), "sfg_cache_write_files")

    @classmethod
    def load_files(cls, cps: "CryptoParams", cacheFilePrefix: str, nrows: int, ncols: int, maxLevel: int = 5) -> "DiagCache":
        """Build the HBM cache from DiagCacheStream files written by the reference for an nrows x ncols matrix."""
        h = C.c_void_p()
        cps._check(cps.L.sfg_cache_load_files(cps.h, str(cacheFilePrefix).encode(), int(nrows), int(ncols), int(maxLevel), C.byref(h)),
                   "sfg_cache_load_files")
        return cls(cps, h)

    def close(self):
        if getattr(self, "h", None):
            self.cps.L.sfg_cache_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def MatMult4StreamPreprocess(cryptoParams: CryptoParams, gfs: GenoFileStream, maxLevel: int, cacheFilePrefix=None) -> DiagCache:
    """gwas/matmult.go:914-1041.  The cache lives in HBM; when ``cacheFilePrefix`` is given the reference's ``<prefix>_<bi>.bin``
    files are written as well (skipped, like the reference does, when the first file already exists: gwas/filestream.go:47-53)."""
    h = C.c_void_p()
    cryptoParams._check(cryptoParams.L.sfg_matmult4_stream_preprocess(cryptoParams.h, gfs.h, maxLevel, C.byref(h)),
                        "MatMult4StreamPreprocess")
    dc = DiagCache(cryptoParams, h)
    if cacheFilePrefix is not None:
        import os

        if not os.path.exists("%s_0.bin" % cacheFilePrefix):
            dc.write_files(cacheFilePrefix)
    return dc


def MatMult4StreamCompute(cryptoParams: CryptoParams, A: np.ndarray, maxLevel: int, cache: DiagCache) -> np.ndarray:
    """gwas/matmult.go:1043-1236. A: [s][numBlockRows][2][levelA+1][N] -> [s][m_ct][2][maxLevel][N] (level maxLevel-1)."""
    A = np.ascontiguousarray(A, dtype=np.uint64)
    if A.ndim != 5 or A.shape[2] != 2 or A.shape[4] != cryptoParams.N:
        raise SfgError("A must be [s][numBlockRows][2][level+1][N]")
    s, nbr, _, nlA, _ = A.shape
    out = np.zeros((s, cache.m_ct, 2, maxLevel, cryptoParams.N), dtype=np.uint64)
    cryptoParams._check(cryptoParams.L.sfg_matmult4_stream_compute(cryptoParams.h, _p(A), s, nbr, nlA - 1, maxLevel, cache.h, _p(out)),
                        "MatMult4StreamCompute")
    return out


def MatMult4Stream(cryptoParams: CryptoParams, A: np.ndarray, gfs: GenoFileStream, maxLevel: int, computeSquaredSum: bool,
                   square: bool, nproc: int = 0):
    """gwas/matmult.go:1238-1505. Returns (out, sum, sqSum); sum/sqSum are None unless computeSquaredSum (like the nil slices)."""
    A = np.ascontiguousarray(A, dtype=np.uint64)
    if A.ndim != 5 or A.shape[2] != 2 or A.shape[4] != cryptoParams.N:
        raise SfgError("A must be [s][numBlockRows][2][level+1][N]")
    s, nbr, _, nlA, _ = A.shape
    if nbr != (gfs.nrows - 1) // cryptoParams.slots + 1:
        raise SfgError("A has %d block rows but the genotype stream has %d" % (nbr, (gfs.nrows - 1) // cryptoParams.slots + 1))
    m_ct = (gfs.ncols - 1) // cryptoParams.slots + 1
    out = np.zeros((s, m_ct, 2, maxLevel, cryptoParams.N), dtype=np.uint64)
    sm = np.zeros(gfs.ncols, dtype=np.float64) if computeSquaredSum else None
    sq = np.zeros(gfs.ncols, dtype=np.float64) if computeSquaredSum else None
    cryptoParams._check(
        cryptoParams.L.sfg_matmult4_stream(cryptoParams.h, _p(A), s, nlA - 1, gfs.h, maxLevel, int(computeSquaredSum), int(square),
                                           _p(out), _p(sm) if computeSquaredSum else None, _p(sq) if computeSquaredSum else None),
        "MatMult4Stream")
    return out, sm, sq


# ------------------------------------------------------------------------------------------------------------------------
# The callers' ciphertext algebra around the path: crypto.CMult, CMultScalar, InnerSumAll, InnerProd,
# MaskTrunc, CSub and the two lazy-normalisation wrappers QXLazyNormStream / QXtLazyNormStream (gwas/matmult.go:27-116).
# A ciphertext carries its limbs and its scale like *ckks.Ciphertext; CipherVector = list[Ciphertext], CipherMatrix = list of those.
# ------------------------------------------------------------------------------------------------------------------------
class Ciphertext:
    """*ckks.Ciphertext of degree 1: ``value`` [2][level+1][N] uint64 (Value()[k].Coeffs[l][j]) and ``scale``."""

    __slots__ = ("value", "scale")

    def __init__(self, value: np.ndarray, scale: float):
        self.value = np.ascontiguousarray(value, dtype=np.uint64)
        self.scale = float(scale)

    def Level(self) -> int:
        return self.value.shape[1] - 1

    def Scale(self) -> float:
        return self.scale

    def CopyNew(self) -> "Ciphertext":
        return Ciphertext(self.value.copy(), self.scale)


def SaveCipherMatrixToFile(cps, cm, filename: str):
    """crypto.SaveCipherMatrixToFile (crypto/utilities.go:82-113): ``cm`` is a CipherMatrix (list of CipherVectors) of degree-1 ciphertexts
    at one level -- what MatMult4StreamCompute returns and gwas/assoc.go:317-333 caches as ``assoc_cache_mult.%d.bin``."""
    L = load()
    nr, nc = len(cm), len(cm[0])
    lvl = cm[0][0].Level()
    if any(len(row) != nc or any(c.Level() != lvl for c in row) for row in cm):
        raise SfgError("SaveCipherMatrixToFile: ragged matrix or mixed levels")
    cts = np.ascontiguousarray(np.stack([np.stack([c.value for c in row]) for row in cm]), dtype=np.uint64)
    sc = np.ascontiguousarray([[c.scale for c in row] for row in cm], dtype=np.float64)
    logN = int(cts.shape[-1]).bit_length() - 1
    if L.sfg_cipher_matrix_save(str(filename).encode(), logN, _p(cts), _p(sc), nr, nc, lvl) != 0:
        raise SfgError("SaveCipherMatrixToFile: " + L.sfg_last_error(None).decode())


def LoadCipherMatrixFromFile(cps, filename: str):
    """crypto.LoadCipherMatrixFromFile (crypto/utilities.go:115-141) -> CipherMatrix (list of lists of Ciphertext)."""
    L = load()
    nr, nc, lvl, logN = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    if L.sfg_cipher_matrix_info(str(filename).encode(), C.byref(nr), C.byref(nc), C.byref(lvl), C.byref(logN)) != 0:
        raise SfgError("LoadCipherMatrixFromFile: " + L.sfg_last_error(None).decode())
    if cps is not None and logN.value != cps.logN:
        raise SfgError("LoadCipherMatrixFromFile: file has logN %d, parameters have %d" % (logN.value, cps.logN))
    cts = np.zeros((nr.value, nc.value, 2, lvl.value + 1, 1 << logN.value), dtype=np.uint64)
    sc = np.zeros((nr.value, nc.value), dtype=np.float64)
    if L.sfg_cipher_matrix_load(str(filename).encode(), logN.value, _p(cts), _p(sc), nr.value, nc.value, lvl.value) != 0:
        raise SfgError("LoadCipherMatrixFromFile: " + L.sfg_last_error(None).decode())
    return [[Ciphertext(cts[i, j], sc[i, j]) for j in range(nc.value)] for i in range(nr.value)]


def _bigints_to_words(vals, nwords=None):
    """Python ints [nct][N] -> (magnitude words [nct][N][nwords] little-endian uint64, sign [nct][N] int8): big.Int.Bits() / Sign()."""
    nct, N = len(vals), len(vals[0])
    if nwords is None:
        nwords = max(1, max((abs(int(v)).bit_length() + 63) // 64 for row in vals for v in row))
    mag = np.zeros((nct, N, nwords), dtype=np.uint64)
    sign = np.zeros((nct, N), dtype=np.int8)
    mask64 = (1 << 64) - 1
    for t in range(nct):
        for j in range(N):
            v = int(vals[t][j])
            sign[t, j] = (v > 0) - (v < 0)
            a = abs(v)
            for w in range(nwords):
                mag[t, j, w] = (a >> (64 * w)) & mask64
    return mag, sign, nwords


def RefreshGenShares(cps: CryptoParams, level: int, c1: np.ndarray, sk_mont: np.ndarray, crp: np.ndarray, mask, e0, e1, in_scale: float = None,
                     out_scale: float = None):
    """dckks.RefreshProtocol.GenShares for nct ciphertexts (mpc/mhe.go:303-311 inside CollectiveBootstrap / CollectiveBootstrapMat).
    c1 [nct][level+1][N]; sk_mont [nQ][N] (cps.Sk.Value.Coeffs: NTT + Montgomery form); crp [nct][nQ][N] (crpGen.ReadNew()); mask [nct][N]
    Python ints (ring.RandInt draws, centred); e0 / e1 [nct][N] Gaussian noise; in_scale = ct.Scale(), out_scale = targetScale.  Returns (shareDecrypt [nct][level+1][N], shareRecrypt
    [nct][nQ][N]).  The random draws and the network aggregation stay with the caller."""
    c1 = np.ascontiguousarray(c1, dtype=np.uint64)
    nct = c1.shape[0]
    mag, sign, nw = _bigints_to_words(mask)
    sk_mont = np.ascontiguousarray(sk_mont[: cps.nQ], dtype=np.uint64)
    crp = np.ascontiguousarray(crp, dtype=np.uint64)
    e0 = np.ascontiguousarray(e0, dtype=np.int64)
    e1 = np.ascontiguousarray(e1, dtype=np.int64)
    h0 = np.zeros((nct, level + 1, cps.N), dtype=np.uint64)
    h1 = np.zeros((nct, cps.nQ, cps.N), dtype=np.uint64)
    cps._check(cps.L.sfg_refresh_gen_shares(cps.h, level, nct, _p(c1), _p(sk_mont), _p(crp), _p(mag), _p(sign), nw,
                                            float(cps.scale if in_scale is None else in_scale), float(cps.scale if out_scale is None else out_scale),
                                            _p(e0), _p(e1), _p(h0), _p(h1)),
               "sfg_refresh_gen_shares")
    return h0, h1


def RefreshFinish(cps: CryptoParams, level: int, c0: np.ndarray, in_scale: float, agg_decrypt: np.ndarray, agg_recrypt: np.ndarray, crp: np.ndarray,
                  out_scale: float = None):
    """refProtocol.Decrypt + Recode + Recrypt (mpc/mhe.go:316-318) on nct ciphertexts: c0 [nct][>= level+1][N] and the aggregated shares ->
    refreshed ciphertexts [nct][2][nQ][N] at the top level with scale ``out_scale`` (params.Scale())."""
    c0 = np.ascontiguousarray(c0, dtype=np.uint64)
    nct, c0_nl = c0.shape[0], c0.shape[1]
    out = np.zeros((nct, 2, cps.nQ, cps.N), dtype=np.uint64)
    a0 = np.ascontiguousarray(agg_decrypt, dtype=np.uint64)
    a1 = np.ascontiguousarray(agg_recrypt, dtype=np.uint64)
    crp = np.ascontiguousarray(crp, dtype=np.uint64)
    cps._check(cps.L.sfg_refresh_finish(cps.h, level, nct, _p(c0), c0_nl, float(in_scale), float(cps.scale if out_scale is None else out_scale),
                                        _p(a0), _p(a1), _p(crp), _p(out)), "sfg_refresh_finish")
    return out


def SetRelinKey(cps: CryptoParams, rlk: np.ndarray):
    """cryptoParams.Rlk.Keys[0]: [beta][2][nQ+nP][N], NTT + Montgomery form."""
    rlk = np.ascontiguousarray(rlk, dtype=np.uint64)
    if rlk.shape != (cps.beta, 2, cps.nQP, cps.N):
        raise SfgError(f"relinearisation key shape {rlk.shape} != {(cps.beta, 2, cps.nQP, cps.N)}")
    cps._check(cps.L.sfg_ctx_set_relin_key(cps.h, _p(rlk)), "sfg_ctx_set_relin_key")


def _nrescale(cps: CryptoParams, scale: float, level: int, threshold=None):
    """evaluator.Rescale(ct, threshold, ct) (Lattigo v2.1): steps taken while scale >= threshold*q_level/2 and level > 0."""
    threshold = cps.scale if threshold is None else threshold
    n = 0
    while level - n > 0 and scale >= threshold * float(cps.Q[level - n]) / 2:
        scale /= float(cps.Q[level - n])
        n += 1
    return n, scale


def _stack(cts, level):
    """cts (same stored limb count or not) -> one array [n][2][level+1][N]: DropLevel to `level` = limb truncation."""
    return np.ascontiguousarray(np.stack([c.value[:, : level + 1] for c in cts]))


def _mul_relin_batch(cps, X, Y):
    n = max(len(X), len(Y))
    if not ((len(X) in (1, n)) and (len(Y) in (1, n))):
        raise SfgError("CMult: vector lengths %d and %d do not broadcast" % (len(X), len(Y)))
    xs = [X[i if len(X) > 1 else 0] for i in range(n)]
    ys = [Y[i if len(Y) > 1 else 0] for i in range(n)]
    out = [None] * n
    groups = {}
    for i in range(n):  # one device batch per (level, scale product)
        lvl = min(xs[i].Level(), ys[i].Level())
        groups.setdefault((lvl, xs[i].scale * ys[i].scale), []).append(i)
    for (lvl, sc), idx in groups.items():
        k, sc2 = _nrescale(cps, sc, lvl)
        bx = len(X) == 1
        by = len(Y) == 1
        ax = _stack([X[0]] if bx else [xs[i] for i in idx], lvl)
        ay = _stack([Y[0]] if by else [ys[i] for i in idx], lvl)
        res = np.zeros((len(idx), 2, lvl + 1 - k, cps.N), dtype=np.uint64)
        cps._check(cps.L.sfg_ct_mul_relin(cps.h, lvl, _p(ax), ax.shape[0], lvl + 1, _p(ay), ay.shape[0], lvl + 1, k, _p(res)), "sfg_ct_mul_relin")
        for j, i in enumerate(idx):
            out[i] = Ciphertext(res[j], sc2)
    return out


def CMult(cryptoParams: CryptoParams, X, Y):
    """crypto.CMult (crypto/basics.go:386-427): element-wise MulRelinNew + Rescale(params.Scale); a length-1 side is broadcast."""
    return _mul_relin_batch(cryptoParams, X, Y)


def CMultScalar(cryptoParams: CryptoParams, X, ct: Ciphertext):
    """crypto.CMultScalar (crypto/basics.go:553-566)."""
    return _mul_relin_batch(cryptoParams, X, [ct])


def _addsub(cps, a: Ciphertext, b: Ciphertext, sub: bool) -> Ciphertext:
    r = max(a.scale, b.scale) / min(a.scale, b.scale)
    if math.floor(r) > 1:  # Lattigo v2.1 evaluateInPlace would first multiply the smaller-scale operand by floor(ratio)
        raise SfgError("ciphertext add/sub with scales %g and %g needs scale matching, which this path never does" % (a.scale, b.scale))
    lvl = min(a.Level(), b.Level())
    out = np.zeros((2, lvl + 1, cps.N), dtype=np.uint64)
    fn = cps.L.sfg_ct_sub if sub else cps.L.sfg_ct_add2
    cps._check(fn(cps.h, lvl, _p(a.value), 1, a.Level() + 1, _p(b.value), 1, b.Level() + 1, _p(out)), "sfg_ct_sub" if sub else "sfg_ct_add2")
    return Ciphertext(out, max(a.scale, b.scale))


def CSub(cryptoParams: CryptoParams, X, Y):
    """crypto.CSub (crypto/basics.go:580-590)."""
    return [_addsub(cryptoParams, X[i], Y[i], True) for i in range(len(Y))]


def CAdd(cryptoParams: CryptoParams, X, Y):
    """crypto.CAdd (crypto/basics.go:568-578)."""
    return [_addsub(cryptoParams, X[i], Y[i], False) for i in range(max(len(X), len(Y)))]


def InnerSumAll(cryptoParams: CryptoParams, X) -> Ciphertext:
    """crypto.InnerSumAll (crypto/basics.go:278-293): every slot of the result holds the sum of all slots of all ciphertexts of X."""
    cps = cryptoParams
    lvl = min(c.Level() for c in X)
    a = _stack(X, lvl)
    out = np.zeros((1, 2, lvl + 1, cps.N), dtype=np.uint64)
    cps._check(cps.L.sfg_inner_sum_all(cps.h, lvl, _p(a), 1, len(X), _p(out)), "sfg_inner_sum_all")
    return Ciphertext(out[0], max(c.scale for c in X))


def InnerProd(cryptoParams: CryptoParams, X, Y) -> Ciphertext:
    """crypto.InnerProd (crypto/basics.go:274-276)."""
    return InnerSumAll(cryptoParams, CMult(cryptoParams, X, Y))


def MaskTrunc(cryptoParams: CryptoParams, ct: Ciphertext, N: int, mask_pt: np.ndarray = None) -> Ciphertext:
    """crypto.MaskTrunc (crypto/basics.go:110-127): retain the first N slots (consumes a level).  ``mask_pt`` lets the caller pass the
    reference-encoded mask plaintext ([nQ][N], NTT domain); by default the mask is encoded on the device (correctly rounded)."""
    cps = cryptoParams
    if N == cps.slots:
        return ct
    if mask_pt is None:
        m = np.zeros(cps.slots, dtype=np.int8)
        m[:N] = 1
        mask_pt = np.zeros((cps.nQ, cps.N), dtype=np.uint64)
        cps._check(cps.L.sfg_encode_slots_i8(cps.h, _p(m), cps.nQ - 1, 0, _p(mask_pt)), "sfg_encode_slots_i8")
    mask_pt = np.ascontiguousarray(mask_pt, dtype=np.uint64)
    lvl = min(ct.Level(), mask_pt.shape[0] - 1)
    k, sc = _nrescale(cps, ct.scale * cps.scale, lvl)
    out = np.zeros((1, 2, lvl + 1 - k, cps.N), dtype=np.uint64)
    cps._check(cps.L.sfg_ct_mul_plain(cps.h, lvl, _p(mask_pt), 1, mask_pt.shape[0], _p(ct.value), 1, ct.Level() + 1, k, _p(out)), "sfg_ct_mul_plain")
    return Ciphertext(out[0], sc)


def _matmult_cm(cps, M, cache: DiagCache, maxLevel=5):
    """MatMult4StreamCompute on a CipherMatrix (list of CipherVectors): DropLevel to maxLevel, run, re-wrap with the output scale."""
    lvl = min(c.Level() for row in M for c in row)
    if lvl < maxLevel:
        raise SfgError("input level %d is smaller than the requested level %d" % (lvl, maxLevel))  # crypto/basics.go:806-824
    A = np.ascontiguousarray(np.stack([_stack(row, maxLevel) for row in M]))
    out = MatMult4StreamCompute(cps, A, maxLevel, cache)
    sc = M[0][0].scale * cps.scale  # gwas/matmult.go:1045
    return [[Ciphertext(out[i, j], sc) for j in range(out.shape[1])] for i in range(out.shape[0])]


def QXLazyNormStream(cps: CryptoParams, mpcObj, Q, Xcache: DiagCache, XMean, XStdInv, numInd: int):
    """gwas/matmult.go:27-77: Q*S*(X - m*1^T) = (Q*S)*X - ((Q*S)*m)*1^T.  ``mpcObj`` supplies GetPid() and Network.BootstrapMatAll(cps, M)
    (the collective bootstrap is the reference's network protocol and stays outside this library)."""
    if mpcObj.GetPid() == 0:
        return None
    slots = cps.GetSlots()
    QS = [CMult(cps, Q[i], XStdInv) for i in range(len(Q))]
    out = _matmult_cm(cps, QS, Xcache, 5)
    out = mpcObj.Network.BootstrapMatAll(cps, out)
    QSm = [InnerProd(cps, QS[i], XMean) for i in range(len(Q))]
    for i in range(len(QS)):
        for j in range(len(out[i])):
            out[i][j] = _addsub(cps, out[i][j], QSm[i], True)
        for j in range(len(out[i])):
            n = slots if j < len(out[i]) - 1 else ((numInd - 1) % slots) + 1
            out[i][j] = MaskTrunc(cps, out[i][j], n)
    return out


def QXtLazyNormStream(cps: CryptoParams, mpcObj, Q, XTcache: DiagCache, XMean, XStdInv):
    """gwas/matmult.go:83-116: Q*(X^T - 1*m^T)*S = ((Q*X^T) - ((Q*1)*m^T))*S."""
    if mpcObj.GetPid() == 0:
        return None
    out = _matmult_cm(cps, Q, XTcache, 5)
    out = mpcObj.Network.BootstrapMatAll(cps, out)
    for i in range(len(out)):
        rowSum = InnerSumAll(cps, Q[i])
        Q1m = CMultScalar(cps, XMean, rowSum)
        for j in range(len(out[i])):
            out[i][j] = _addsub(cps, out[i][j], Q1m[j], True)
    for i in range(len(out)):
        out[i] = CMult(cps, out[i], XStdInv)
    return out


# ------------------------------------------------------------------------------------------------------------------------
# Device-resident ciphertext vectors ("keeps Q on device across a power iteration").  Same algebra as above, but the
# operands live in HBM between calls (sfg_cts handles); only what has to reach the host (the network bootstrap, final results) moves.
# ------------------------------------------------------------------------------------------------------------------------
class DeviceCipherVector:
    """n degree-1 ciphertexts at one level and scale, resident in HBM (a CipherVector, or a CipherMatrix stored row-major)."""

    def __init__(self, cps: CryptoParams, h, scale: float):
        self.cps, self.h, self.scale = cps, h, float(scale)
        n, nl = C.c_int(), C.c_int()
        cps.L.sfg_cts_shape(h, C.byref(n), C.byref(nl))
        self.n, self.nl = n.value, nl.value
        cps._children.add(self)

    @classmethod
    def upload(cls, cps: CryptoParams, cts) -> "DeviceCipherVector":
        """cts: list of Ciphertext at one level and scale."""
        lvl = min(c.Level() for c in cts)
        a = _stack(cts, lvl)
        h = C.c_void_p()
        cps._check(cps.L.sfg_cts_upload(cps.h, _p(a), a.shape[0], lvl + 1, C.byref(h)), "sfg_cts_upload")
        return cls(cps, h, cts[0].scale)

    def download(self):
        out = np.zeros((self.n, 2, self.nl, self.cps.N), dtype=np.uint64)
        self.cps._check(self.cps.L.sfg_cts_download(self.cps.h, self.h, _p(out)), "sfg_cts_download")
        return [Ciphertext(out[k], self.scale) for k in range(self.n)]

    def Level(self) -> int:
        return self.nl - 1

    def slice(self, first: int, count: int) -> "DeviceCipherVector":
        h = C.c_void_p()
        self.cps._check(self.cps.L.sfg_cts_slice(self.cps.h, self.h, first, count, C.byref(h)), "sfg_cts_slice")
        return DeviceCipherVector(self.cps, h, self.scale)

    def close(self):
        if getattr(self, "h", None):
            self.cps.L.sfg_cts_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _dev_cmult(cps, X: DeviceCipherVector, Y: DeviceCipherVector) -> DeviceCipherVector:
    """crypto.CMult / CMultScalar on handles (a side of length 1 broadcasts)."""
    lvl = min(X.Level(), Y.Level())
    k, sc = _nrescale(cps, X.scale * Y.scale, lvl)
    h = C.c_void_p()
    cps._check(cps.L.sfg_cts_mul_relin(cps.h, lvl, X.h, Y.h, k, C.byref(h)), "sfg_cts_mul_relin")
    return DeviceCipherVector(cps, h, sc)


def _dev_sub(cps, a: DeviceCipherVector, b: DeviceCipherVector) -> DeviceCipherVector:
    r = max(a.scale, b.scale) / min(a.scale, b.scale)
    if math.floor(r) > 1:
        raise SfgError("ciphertext sub with scales %g and %g needs scale matching, which this path never does" % (a.scale, b.scale))
    lvl = min(a.Level(), b.Level())
    h = C.c_void_p()
    cps._check(cps.L.sfg_cts_addsub(cps.h, lvl, a.h, b.h, 1, C.byref(h)), "sfg_cts_addsub")
    return DeviceCipherVector(cps, h, max(a.scale, b.scale))


def QXtLazyNormStreamDevice(cps: CryptoParams, mpcObj, Q: DeviceCipherVector, s: int, XTcache: DiagCache, XMean: DeviceCipherVector,
                            XStdInv: DeviceCipherVector) -> DeviceCipherVector:
    """gwas/matmult.go:83-116 with Q (s rows x num_block_rows ciphertexts, row-major), XMean and XStdInv resident in HBM: MatMult, the
    row sums, the mean correction and the final scaling chain on the device; the only host round trip is the network bootstrap
    (``mpcObj.Network.BootstrapMatAll``), which the reference performs on the host as well.  Bit-identical to QXtLazyNormStream."""
    if mpcObj.GetPid() == 0:
        return None
    L = cps.L
    nbr = Q.n // s
    h = C.c_void_p()
    cps._check(L.sfg_cts_matmult4_stream_compute(cps.h, Q.h, s, nbr, 5, XTcache.h, C.byref(h)), "sfg_cts_matmult4_stream_compute")
    out = DeviceCipherVector(cps, h, Q.scale * cps.scale)
    m_ct = out.n // s
    # the collective bootstrap is a network protocol: download, bootstrap, upload (what the reference does with every MatMult output)
    host = out.download()
    boot = mpcObj.Network.BootstrapMatAll(cps, [host[i * m_ct:(i + 1) * m_ct] for i in range(s)])
    out = DeviceCipherVector.upload(cps, [c for row in boot for c in row])
    rows = []
    for i in range(s):
        qi = Q.slice(i * nbr, nbr)
        hs = C.c_void_p()
        cps._check(L.sfg_cts_inner_sum_all(cps.h, qi.Level(), qi.h, 1, nbr, C.byref(hs)), "sfg_cts_inner_sum_all")
        rowSum = DeviceCipherVector(cps, hs, qi.scale)
        Q1m = _dev_cmult(cps, XMean, rowSum)            # CMultScalar(XMean, rowSum)
        oi = _dev_sub(cps, out.slice(i * m_ct, m_ct), Q1m)
        rows.append(_dev_cmult(cps, oi, XStdInv))       # CMult(out[i], XStdInv)
    return rows
