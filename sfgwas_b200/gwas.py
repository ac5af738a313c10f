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
    "CryptoParams", "GenoFileStream", "DiagCache", "MatMult4StreamPreprocess", "MatMult4StreamCompute", "MatMult4Stream",
    "SfgError",
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
    """gwas/matmult.go:914-1041. ``cacheFilePrefix`` is accepted for signature parity; the cache lives in HBM."""
    h = C.c_void_p()
    cryptoParams._check(cryptoParams.L.sfg_matmult4_stream_preprocess(cryptoParams.h, gfs.h, maxLevel, C.byref(h)),
                        "MatMult4StreamPreprocess")
    return DiagCache(cryptoParams, h)


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
