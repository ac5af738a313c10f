"""Multi-GPU sharding of the MatMult hot path on one box: one process per GPU, torch.distributed for the plumbing.

Three shardings, all bit-identical to the single-GPU result:

* ``GiantSharded`` -- STRONG scaling of one product: rank r owns a contiguous share of the active giant steps g, i.e. of the columns
  (g, bj) of the dense contraction: 1/world of the diagonal cache, of the MAC and of the giant-step key-switches.
  ``out[i][bj] = sum_g RotL_{g d}(cv[i][g][bj])`` is a mod-q sum over g (gwas/matmult.go:1223-1227), so the per-rank partial outputs
  are combined by ONE modular-add all-reduce of ``s * m_ct`` ciphertexts over NVLink (NCCL integer SUM + ``sfg_ct_mod_reduce``); the
  d giant steps balance over 2 / 4 / 8 ranks whatever the number of block columns.  The baby-step rotations (7 % of a single-GPU step,
  but a third of an 8-GPU one if every rank repeated them) are sharded as well: 1/world of the (block row, baby step) entries per rank,
  completed by one all-gather of the rotation cache.  This is what ``bench.py --gpus N`` measures.

* ``ColumnSharded`` -- SNP-block (= block-column) sharding for Q.X-shaped products with X = nind x nsnp: rank r owns a contiguous
  range of block columns; every accumulator (i, giant, bj), its reduce, its giant rotations and the final add are independent
  across bj, so there is NO data-path collective (weak scaling; this is what ``bench.py --gpus N`` measures).  ``gather()`` is an
  optional all-gather of the finished output ciphertexts.

* ``RowSharded`` -- block-row sharding: rank r multiplies only the block rows [bi_lo, bi_hi) (its baby rotations and its part of
  the K loop).  Partial sums must be combined BEFORE the giant-step rotations (key-switching is not bit-linear), i.e. between K2
  and K6: an integer SUM all-reduce of canonical residues (at most 2^8 ranks of residues < 2^56 fit a u64 -- NCCL ``ncclSum`` on
  int64 is exactly that) followed by one ``mod q`` pass (``sfg_cv_mod_reduce``), then every rank rotates and adds its own range
  of giant steps (``sfg_matmult4_finish``) and the per-rank outputs are summed the same way.
  Every rank builds the diagonal cache of its own block rows only (``sfg_matmult4_stream_preprocess_rows``): HBM and preprocessing
  time are proportional to the rank's share.

Reference context: the reference has no intra-box parallelism beyond goroutines (gwas/matmult.go:983-1036, 1138-1169); the
cross-party sum stays in Go (mpc/aggregate.go:466-500).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from .gwas import CryptoParams, DiagCache, GenoFileStream, MatMult4StreamCompute, MatMult4StreamPreprocess, SfgError, _p


def partition(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced ranges [lo, hi) of n units over `world` ranks (the first n % world ranks get one more)."""
    if world < 1:
        raise ValueError("world must be >= 1")
    base, extra = divmod(n, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def mod_allreduce_(t, moduli: Sequence[int], N: int, group=None, cps: CryptoParams = None, cache: DiagCache = None, s: int = 0,
                   max_level: int = 0):
    """In-place modular sum over ranks of a tensor of canonical residues laid out [..., L, N] (int64 view of u64).

    Integer SUM all-reduce + one canonicalisation pass.  CUDA tensors are reduced by NCCL and canonicalised by the library
    (``sfg_cv_mod_reduce``); CPU tensors (gloo, tests) are canonicalised with torch ops.
    """
    import torch
    import torch.distributed as dist

    L = len(moduli)
    world = dist.get_world_size(group)
    if max(moduli) * world >= 1 << 63:
        raise SfgError("sum of %d residues may overflow int64 for a %d-bit modulus" % (world, max(moduli).bit_length()))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    if t.is_cuda:
        if cps is None or cache is None:
            raise SfgError("CUDA tensors need the context and cache handles for the canonicalisation pass")
        torch.cuda.current_stream().synchronize()
        cps._check(cps.L.sfg_cv_mod_reduce(cps.h, cache.h, s, max_level, C.c_void_p(t.data_ptr()), 0, t.numel()), "sfg_cv_mod_reduce")
    else:
        v = t.view(-1, L, N)
        for l, q in enumerate(moduli):
            v[:, l, :] %= q
    return t


def ct_mod_allreduce_(t, cps: CryptoParams, nl: int, group=None):
    """In-place modular sum over ranks of a CUDA tensor of canonical residues [..., nl, N] (int64 view of u64): NCCL integer SUM all-reduce
    (at most 2^8 ranks of residues < 2^56) + the library's canonicalisation kernel.  Stream-ordered on torch's current stream."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if max(cps.Q[:nl]) * world >= 1 << 63:
        raise SfgError("sum of %d residues may overflow int64 for a %d-bit modulus" % (world, max(cps.Q[:nl]).bit_length()))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    torch.cuda.current_stream().synchronize()
    cps._check(cps.L.sfg_ct_mod_reduce(cps.h, C.c_void_p(t.data_ptr()), t.numel() // (nl * cps.N), nl), "sfg_ct_mod_reduce")
    return t


def giant_share(ng: int, world: int) -> int:
    """Giants per rank of the reduce-scattered accumulator image: equal shares (the last ranks' tail is zero padding)."""
    return (ng + world - 1) // world


def mod_reduce_scatter_(t, share: int, cps: CryptoParams = None, nl: int = 0, group=None, moduli: Sequence[int] = None, N: int = 0):
    """Modular-add reduce-scatter of canonical residues: `t` holds world * share int64 words laid out [..., nl, N]; returns this rank's
    share summed over ranks and canonicalised.  NCCL ``reduce_scatter_tensor`` + ``sfg_ct_mod_reduce`` on CUDA tensors; on CPU tensors
    (gloo has no reduce-scatter: the world-size-2 CPU tests of the host logic) an all-reduce + slice with torch ops."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if t.numel() != world * share:
        raise SfgError("reduce-scatter: %d words do not split into %d shares of %d" % (t.numel(), world, share))
    if t.is_cuda:
        if max(cps.Q[:nl]) * world >= 1 << 63:
            raise SfgError("sum of %d residues may overflow int64" % world)
        mine = torch.empty(share, dtype=t.dtype, device=t.device)
        dist.reduce_scatter_tensor(mine, t, op=dist.ReduceOp.SUM, group=group)
        torch.cuda.current_stream().synchronize()
        cps._check(cps.L.sfg_ct_mod_reduce(cps.h, C.c_void_p(mine.data_ptr()), share // (nl * cps.N), nl), "sfg_ct_mod_reduce")
        return mine
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    mine = t[rank * share:(rank + 1) * share].clone()
    v = mine.view(-1, len(moduli), N)
    for l, q in enumerate(moduli):
        v[:, l, :] %= q
    return mine


def _compute_shared_baby_steps(cps: CryptoParams, cache: DiagCache, d_A, d_out, s: int, nbr: int, level_a: int, max_level: int, rank: int,
                               world: int, holder: dict, group=None, check_chunks: bool = False):
    """MatMult4StreamCompute over this rank's `cache` with the baby-step rotations shared between the ranks: every rank rotates 1/world
    of the (block row, baby step) entries (``sfg_matmult4_baby_dev``), ONE all-gather over NVLink completes the rotation cache, then MAC +
    giant-step sums run on it (``sfg_matmult4_stream_compute_r_dev``).  Falls back to the plain device call when there is nothing to share
    (world == 1, s > 16).  `holder` keeps the gathered buffer between calls; returns the library's phase timings of the call."""
    import torch
    import torch.distributed as dist

    L = cps.L
    ph = dict(baby_ms=0.0, mac_ms=0.0, giant_ms=0.0, mac_kernel_ms=0.0)

    def add_timings():
        t = cps.last_timings()
        for k in ph:
            ph[k] += t[k]

    # the library's stream is non-blocking: whatever torch still has in flight for d_A / d_out (fills, copies) must be done first
    torch.cuda.current_stream().synchronize()
    if world > 1 and s <= 16:
        chunk = int(L.sfg_matmult4_baby_chunk_bytes(cps.h, cache.h, s, world))
        R = holder.get("R")
        if R is None or R.numel() != world * chunk:
            if check_chunks:  # caches built from different columns must still see the same baby steps
                cmm = torch.tensor([chunk, -chunk], device=d_A.device)
                dist.all_reduce(cmm, op=dist.ReduceOp.MAX, group=group)
                if int(cmm[0]) != chunk or -int(cmm[1]) != chunk:
                    raise SfgError("ranks disagree on the rotation-cache share (a rank without a full-width block column?)")
            R = holder["R"] = torch.empty(world * chunk, dtype=torch.uint8, device=d_A.device)
        cps._check(L.sfg_matmult4_baby_dev(cps.h, C.c_void_p(d_A.data_ptr()), s, nbr, level_a, max_level, cache.h, rank, world,
                                           C.c_void_p(R.data_ptr())), "sfg_matmult4_baby_dev")
        add_timings()
        mine = R[rank * chunk:(rank + 1) * chunk].clone()
        dist.all_gather_into_tensor(R, mine, group=group)
        torch.cuda.current_stream().synchronize()
        cps._check(L.sfg_matmult4_stream_compute_r_dev(cps.h, C.c_void_p(R.data_ptr()), s, max_level, cache.h, C.c_void_p(d_out.data_ptr())),
                   "sfg_matmult4_stream_compute_r_dev")
        add_timings()
    else:
        cps._check(L.sfg_matmult4_stream_compute_dev(cps.h, C.c_void_p(d_A.data_ptr()), s, nbr, level_a, max_level, cache.h,
                                                     C.c_void_p(d_out.data_ptr())), "sfg_matmult4_stream_compute_dev")
        add_timings()
    return ph


class GiantSharded:
    """Strong scaling of ONE MatMult4StreamCompute over the GPUs of a box: rank `rank` of `world` holds the diagonals of its share of the
    giant steps and produces the partial sum over them; ``compute`` all-reduces the partial outputs mod q (device-resident throughout).
    With ``shard_baby`` (default) the baby-step rotations -- which every rank would otherwise repeat -- are sharded too: every rank rotates
    1/world of the (block row, baby step) entries and ONE all-gather over NVLink completes the rotation cache on every rank."""

    def __init__(self, cps: CryptoParams, gfs: GenoFileStream, rank: int, world: int, max_level: int = 5, shard_baby: bool = True):
        self.cps, self.rank, self.world, self.max_level = cps, rank, world, max_level
        self.shard_baby = shard_baby and world > 1
        h = C.c_void_p()
        cps._check(cps.L.sfg_matmult4_stream_preprocess_giants(cps.h, gfs.h, max_level, rank, world, C.byref(h)),
                   "sfg_matmult4_stream_preprocess_giants")
        self.cache = DiagCache(cps, h)
        self._hold = {}
        self.phases_ms = dict(baby_ms=0.0, mac_ms=0.0, giant_ms=0.0, mac_kernel_ms=0.0)  # of the last compute_dev

    def compute_dev(self, d_A, d_out, s: int, nbr: int, level_a: int, group=None):
        """d_A [s][nbr][2][level_a+1][N], d_out [s][m_ct][2][max_level][N]: int64 CUDA tensors; d_out = the FULL product on every rank."""
        cps = self.cps
        ph = _compute_shared_baby_steps(cps, self.cache, d_A, d_out, s, nbr, level_a, self.max_level, self.rank,
                                        self.world if self.shard_baby else 1, self._hold, group)
        if self.world > 1:
            ct_mod_allreduce_(d_out, cps, self.max_level, group)
        self.phases_ms = ph
        return d_out

    def compute(self, A: np.ndarray, group=None) -> np.ndarray:
        import torch

        cps = self.cps
        A = np.ascontiguousarray(A, dtype=np.uint64)
        s, nbr, _, nlA, _ = A.shape
        dev = torch.device("cuda", cps.device)
        d_A = torch.from_numpy(A.view(np.int64)).to(dev)
        d_out = torch.zeros((s, self.cache.m_ct, 2, self.max_level, cps.N), dtype=torch.int64, device=dev)
        self.compute_dev(d_A, d_out, s, nbr, nlA - 1, group)
        return d_out.cpu().numpy().view(np.uint64)


class ColumnSharded:
    """SNP-block sharding: rank `rank` of `world` owns block columns [c_lo, c_hi) of the genotype matrix."""

    def __init__(self, cps: CryptoParams, X_local_cols: np.ndarray, ncols_total: int, rank: int, world: int, max_level: int = 5):
        self.cps, self.rank, self.world, self.max_level = cps, rank, world, max_level
        self.m_ct_total = (ncols_total - 1) // cps.slots + 1
        self.ranges = partition(self.m_ct_total, world)
        lo, hi = self.ranges[rank]
        want = min(hi * cps.slots, ncols_total) - lo * cps.slots
        if hi > lo and X_local_cols.shape[1] != want:
            raise SfgError("rank %d owns %d columns, got %d" % (rank, want, X_local_cols.shape[1]))
        self.empty = hi <= lo
        self._hold = {}
        if not self.empty:
            self.gfs = GenoFileStream.from_matrix(cps, X_local_cols)
            self.cache = MatMult4StreamPreprocess(cps, self.gfs, max_level)

    @staticmethod
    def local_columns(ncols_total: int, slots: int, rank: int, world: int) -> Tuple[int, int]:
        m_ct = (ncols_total - 1) // slots + 1
        lo, hi = partition(m_ct, world)[rank]
        return min(lo * slots, ncols_total), min(hi * slots, ncols_total)

    def compute(self, A: np.ndarray) -> np.ndarray:
        """out[:, c_lo:c_hi] of MatMult4StreamCompute(A, X): [s][hi-lo][2][maxLevel][N]."""
        if self.empty:
            return np.zeros((A.shape[0], 0, 2, self.max_level, self.cps.N), dtype=np.uint64)
        return MatMult4StreamCompute(self.cps, A, self.max_level, self.cache)

    def compute_dev(self, d_A, d_out, s: int, nbr: int, level_a: int, group=None, share_baby: bool = True):
        """Device-resident: d_A [s][nbr][2][level_a+1][N] (the same A on every rank), d_out [s][hi-lo][2][max_level][N] = this rank's block
        columns of the product.  With ``share_baby`` the baby-step rotations -- identical on every rank -- are split 1/world per rank and
        completed by ONE all-gather (the layout the full-size config-4 run uses: bench.py --col-sharding); every rank must then own at
        least one full-width block column so that all caches see the same baby steps (checked)."""
        if self.empty:
            raise SfgError("rank %d owns no block column" % self.rank)
        self.phases_ms = _compute_shared_baby_steps(self.cps, self.cache, d_A, d_out, s, nbr, level_a, self.max_level, self.rank,
                                                    self.world if share_baby else 1, self._hold, group, check_chunks=True)
        return d_out

    def gather(self, out_local: np.ndarray, group=None) -> np.ndarray:
        """All ranks get the full [s][m_ct][2][maxLevel][N] output (all-gather of equal-sized padded pieces)."""
        import torch
        import torch.distributed as dist

        s = out_local.shape[0]
        width = max(hi - lo for lo, hi in self.ranges)
        dev = torch.device("cuda", self.cps.device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        piece = torch.zeros((s, width, 2, self.max_level, self.cps.N), dtype=torch.int64, device=dev)
        if out_local.shape[1]:
            piece[:, : out_local.shape[1]] = torch.from_numpy(out_local.view(np.int64)).to(dev)
        pieces = [torch.empty_like(piece) for _ in range(self.world)]
        dist.all_gather(pieces, piece, group=group)
        full = np.zeros((s, self.m_ct_total, 2, self.max_level, self.cps.N), dtype=np.uint64)
        for r, (lo, hi) in enumerate(self.ranges):
            full[:, lo:hi] = pieces[r][:, : hi - lo].cpu().numpy().view(np.uint64)
        return full


class RowSharded:
    """Block-row sharding with a modular-add all-reduce between the MAC (K1 + K2) and the giant-step rotations (K6)."""

    def __init__(self, cps: CryptoParams, X: np.ndarray, rank: int, world: int, max_level: int = 5):
        self.cps, self.rank, self.world, self.max_level = cps, rank, world, max_level
        self.gfs = GenoFileStream.from_matrix(cps, X)  # int8, 1 byte per genotype: every rank holds the matrix, but encodes only its rows
        nbr = (X.shape[0] - 1) // cps.slots + 1
        self.row_ranges = partition(nbr, world)
        lo, hi = self.row_ranges[rank]
        h = C.c_void_p()
        cps._check(cps.L.sfg_matmult4_stream_preprocess_rows(cps.h, self.gfs.h, max_level, lo, hi, C.byref(h)),
                   "sfg_matmult4_stream_preprocess_rows")
        self.cache = DiagCache(cps, h)  # diagonals of block rows [lo, hi) only

    def compute(self, A: np.ndarray, group=None) -> np.ndarray:
        import torch
        import torch.distributed as dist

        cps, L = self.cps, self.cps.L
        A = np.ascontiguousarray(A, dtype=np.uint64)
        s, nbr, _, nlA, _ = A.shape
        dev = torch.device("cuda", cps.device)
        n_cv = int(L.sfg_cv_elems(cps.h, self.cache.h, s, self.max_level))
        per_g = self.cache.m_ct * 2 * s * self.max_level * cps.N
        ng = n_cv // per_g
        # the image is laid out [giant][bj][row][L][N]: equal shares of `ch` giants per rank (zero-padded tail) make the combination a
        # modular-add REDUCE-SCATTER -- every rank receives only the giants it then rotates (gwas/matmult.go:1203-1227)
        ch = giant_share(ng, self.world)
        cv = torch.zeros(self.world * ch * per_g, dtype=torch.int64, device=dev)
        lo, hi = self.row_ranges[self.rank]
        cps._check(L.sfg_matmult4_partial(cps.h, _p(A), s, nbr, nlA - 1, self.max_level, self.cache.h, lo, hi, C.c_void_p(cv.data_ptr())),
                   "sfg_matmult4_partial")
        mine = mod_reduce_scatter_(cv, ch * per_g, cps, self.max_level, group)
        g_lo, g_hi = min(ng, self.rank * ch), min(ng, (self.rank + 1) * ch)
        t = torch.zeros((s, self.cache.m_ct, 2, self.max_level, cps.N), dtype=torch.int64, device=dev)
        cps._check(L.sfg_matmult4_finish_dev(cps.h, self.cache.h, s, self.max_level, C.c_void_p(mine.data_ptr()), g_lo, g_hi,
                                             C.c_void_p(t.data_ptr())), "sfg_matmult4_finish_dev")
        ct_mod_allreduce_(t, cps, self.max_level, group)  # everything stays on the device until the final read-back
        return t.cpu().numpy().view(np.uint64)
