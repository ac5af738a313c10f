#!/usr/bin/env python3
"""bench.py -- genotype x CKKS-ciphertext MatMult throughput (BASELINE.json metric) on B200.

  python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path (torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU algorithm (oracle port) on the host cores

Default workload (BASELINE.json configs[1]): X = 10 000 samples x 100 000 SNPs int8 (Binomial(2, p_j), p_j ~ U(0.05, 0.5)),
A = 10 rows x 3 ciphertexts at level 5, PN13QP218 (logN = 13, scale 2^30), orientation A.X -> 10 x 25 ciphertexts at level 4.
A "step" is one MatMult4StreamCompute call over the preprocessed diagonal cache (the reference calls Compute ~40x per
Preprocess inside the PCA power iterations, gwas/pca.go:112-113,288-352).  Other workloads (--workload): the transposed
orientation of config 2 (25 block rows, 7 K groups), PCA-shaped logN-14 slices (configs 4/5), and "_otf" variants that force the
diagonals to be re-encoded on the fly inside every call (the regime of configs 4/5, whose cache exceeds HBM).

metric   = B_alg / t in GB/s, B_alg = diag_polys*L'*N*8 + s*nbr*2*6*N*8 + s*m_ct*2*L'*N*8
value    = inputs already resident in HBM
e2e      = the same call through the C ABI with pinned HOST buffers for A and out (H2D + D2H inside the timed region)
e2e_ptrs = the same through the entry point the cgo shim binds: one PAGEABLE host array per limb (what Go hands over)
verify   = inputs are REAL encryptions of a known matrix A under seeded keys; after the timed region a few output ciphertexts are
           decrypted and compared with A.X (tolerance 1e-3 relative to max |A.X|): the number carries its own correctness proof.
           Key generation / encryption / decryption are the Go side's job in the reference; here they come from the CPU oracle in its
           checker role (never inside a timed region).
Multi-GPU (--gpus N under torchrun): STRONG scaling of the same product -- giant-step sharding (sfgwas_b200.dist.GiantSharded): every
rank holds 1/N of the diagonal cache, computes the partial sum over its giant steps and the partial outputs are combined by a
modular-add all-reduce over NVLink (NCCL integer SUM + canonicalisation kernel) inside the timed region.
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PN13 = dict(logN=13, Q=[0x1FFFEC001, 0x3FFF4001, 0x3FFE8001, 0x40020001, 0x40038001, 0x3FFC0001], P=[0x800004001],
            scale=float(1 << 30))
PN14 = dict(logN=14, Q=[0x200000008001, 0x400018001, 0x3FFFD0001, 0x400060001, 0x400068001, 0x3FFF90001, 0x400080001, 0x4000A8001,
                         0x400108001, 0x3FFEB8001], P=[0x7FFFFFD8001, 0x7FFFFFC8001], scale=float(1 << 34))
CKKS = {"PN13QP218": PN13, "PN14QP438": PN14}
WORKLOADS = {
    # name: nrows, ncols, s, CKKS parameter set, orientation label, otf = force the non-materialised (encode-on-the-fly) path
    "mm_10k_x_100k_k10_logN13": dict(nrows=10000, ncols=100000, s=10, params="PN13QP218", orientation="A.X"),
    # config 2 transposed (SURVEY 8d-2 "both orientations"): 25 block rows -> K = 1 600 baby-step slots = 7 K groups, 15 750 baby rotations
    "mm_100k_x_10k_k10_logN13_T": dict(nrows=100000, ncols=10000, s=10, params="PN13QP218", orientation="A'.X^T"),
    "mm_2k_x_20k_k10_logN13": dict(nrows=2000, ncols=20000, s=10, params="PN13QP218", orientation="A.X"),   # quick check only
    # PCA-shaped blocks (BASELINE configs 4 / 5 run at logN 14, kp = 15)
    "mm_16k_x_64k_k15_logN14": dict(nrows=16384, ncols=65536, s=15, params="PN14QP438", orientation="Q.X^T slice"),
    # the regime of configs 4 / 5: the cache cannot be HBM-resident, every call re-encodes the diagonals chunk by chunk
    "mm_10k_x_100k_k10_logN13_otf": dict(nrows=10000, ncols=100000, s=10, params="PN13QP218", orientation="A.X", otf=True),
    "mm_16k_x_64k_k15_logN14_otf": dict(nrows=16384, ncols=65536, s=15, params="PN14QP438", orientation="Q.X^T slice", otf=True),
    # BASELINE config 4 at FULL size (one party's PCA half-iteration, Q.X^T direction: 13 x 62 blocks, 6.6 M diagonals per call): meant
    # for --gpus 8.  colshard: every rank holds only ITS block columns of the int8 matrix (SNP-block sharding, 6.5 of the 50 GB) and
    # produces those block columns of the output -- no all-reduce; the baby-step rotations are still shared 1/N + one all-gather
    "pca_100k_x_500k_k15_logN14_otf": dict(nrows=100000, ncols=500000, s=15, params="PN14QP438", orientation="Q.X^T (config 4, full size)", otf=True,
                                           colshard=True),
    "mm_32k_x_128k_k15_logN14_otf": dict(nrows=32768, ncols=131072, s=15, params="PN14QP438", orientation="Q.X^T slice (config 4, 4 x 16 of 13 x 62 blocks)", otf=True),
}
# complete small call the CPU arm times (same parameter set and s; one block row): (nrows, ncols)
CPU_SAMPLE = {"PN13QP218": (4096, 8192), "PN14QP438": (8192, 8192)}


def work_figures(nrows, ncols, s, logN, maxLevel=5):
    N = 1 << logN
    slots = N // 2
    d = int(math.ceil(math.sqrt(slots)))
    nbr, m_ct = (nrows - 1) // slots + 1, (ncols - 1) // slots + 1
    Lp = maxLevel
    # all `slots` shifts are active whenever any block is full width
    diag_polys = 0
    for bi in range(nbr):
        r = min((bi + 1) * slots, nrows) - bi * slots
        for bj in range(m_ct):
            c = min((bj + 1) * slots, ncols) - bj * slots
            diag_polys += min(slots, r + c - 1)
    b_diag = diag_polys * Lp * N * 8
    b_alg = b_diag + s * nbr * 2 * (maxLevel + 1) * N * 8 + s * m_ct * 2 * Lp * N * 8
    mac_alg = diag_polys * s * 2 * Lp * N
    ks_baby, ks_giant = nbr * (d - 1) * s, s * (d - 1) * m_ct
    return dict(N=N, slots=slots, d=d, nbr=nbr, m_ct=m_ct, diag_polys=diag_polys, b_diag=b_diag, b_alg=b_alg, mac_alg=mac_alg,
                ks_baby=ks_baby, ks_giant=ks_giant)


def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def read_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_mac_tc launch from the committed ncu --set full capture of the default
    workload (newest round first); None when no summary is found."""
    for tag in ("r2", "r1_final"):
        path = os.path.join(ROOT, "profiles", tag, "ncu_k_mac_tc.txt")
        try:
            rd = wr = None
            with open(path) as f:
                for ln in f:
                    t = ln.split()
                    if len(t) >= 3 and t[0] == "dram__bytes_read.sum" and rd is None:
                        rd = float(t[1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[t[2]]
                    if len(t) >= 3 and t[0] == "dram__bytes_write.sum" and wr is None:
                        wr = float(t[1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[t[2]]
            if rd is not None and wr is not None:
                return rd + wr, "ncu --set full, profiles/%s/ncu_k_mac_tc.txt (dram read + write per launch)" % tag
        except Exception:
            pass
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.proc, self.lines = gpu, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100", "-i",
                                          str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------------------------
# Checker role of the CPU oracle: what the Go side of the reference supplies (keys, encryption) and verifies (decryption)
# ------------------------------------------------------------------------------------------------------------------
def real_inputs(P, nrows, s, seed=1):
    """Seeded secret key, BSGS Galois keys and a real encryption A of a known s x nrows matrix (level 5, scale = params.Scale)."""
    import numpy as np

    from oracle.oracle import Oracle

    o = Oracle.from_params(P)
    sk = o.keygen_secret(seed)
    keys = o.gen_bsgs_keys(sk)
    rng = np.random.default_rng(seed)
    Ap = rng.normal(size=(s, nrows))
    nbr = (nrows - 1) // o.slots + 1
    A = np.zeros((s, nbr, 2, 6, o.N), dtype=np.uint64)
    for i in range(s):
        for b in range(nbr):
            A[i, b] = o.encrypt_vector(sk, Ap[i, b * o.slots:(b + 1) * o.slots], 5, seed=1000 + 131 * i + b)
    return o, sk, keys, Ap, A


def verify_decrypt(o, sk, Ap, xcols, out, picks):
    """decrypt(out[i][bj]) ~= (A_plain . X)[i, columns of bj] for the picked (i, bj); xcols[bj] = those columns of X as float64."""
    import numpy as np

    worst, scale_ref = 0.0, 1.0
    for i, bj in picks:
        xc = xcols[bj]  # int8 on the host; converted in row chunks (the full-size config-4 column is 100k x 8192)
        ref = sum(Ap[i][r:r + 8192] @ xc[r:r + 8192].astype(np.float64) for r in range(0, xc.shape[0], 8192))
        got = o.decrypt_vector(sk, out[i, bj], o.scale * o.scale).real[: len(ref)]
        worst = max(worst, float(np.abs(got - ref).max()))
        scale_ref = max(scale_ref, float(np.abs(ref).max()))
    tol = 1e-3 * scale_ref
    return dict(decrypt_max_abs_err=worst, tolerance=tol, max_abs_reference=scale_ref, checked=[list(p) for p in picks], ok=bool(worst < tol))


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port of gwas/matmult.go:1043-1236) -- ONE COMPLETE MatMult4StreamCompute per step on a
# bounded sample of the workload (same parameter set, same s, one block row), cache resident in RAM, all host cores
# ------------------------------------------------------------------------------------------------------------------
class CpuSample:
    def __init__(self, pname, s):
        import numpy as np

        from oracle.oracle import Oracle

        self.P = CKKS[pname]
        self.nthreads = os.cpu_count() or 1
        self.nrows, self.ncols = CPU_SAMPLE[pname]
        self.s = s
        o = self.o = Oracle.from_params(self.P)
        self.wf = work_figures(self.nrows, self.ncols, s, self.P["logN"])
        rng = np.random.default_rng(7)
        maf = rng.uniform(0.05, 0.5, self.ncols)
        X = (rng.random((self.nrows, self.ncols)) < maf).astype(np.int8) + (rng.random((self.nrows, self.ncols)) < maf).astype(np.int8)
        sk = o.keygen_secret(3)
        self.keys = o.gen_bsgs_keys(sk)
        nbr = self.wf["nbr"]
        self.A = np.zeros((s, nbr, 2, 6, o.N), dtype=np.uint64)
        for i in range(s):
            for b in range(nbr):
                self.A[i, b] = o.encrypt_vector(sk, rng.normal(size=o.slots), 5, seed=50 + 13 * i + b)
        t0 = time.perf_counter()
        self.dc = o.preprocess(X, 5, nproc=self.nthreads)  # MatMult4StreamPreprocess: not part of a step (like the GPU arm's cache)
        self.t_prep = time.perf_counter() - t0

    def step(self):
        t0 = time.perf_counter()
        _, ph = self.o.compute(self.A, self.dc, self.keys, 5, nproc=self.nthreads, timings=True)
        return time.perf_counter() - t0, ph

    def describe(self, full_wf, t, ph):
        wf = self.wf
        ksb = (wf["ks_baby"] + wf["ks_giant"]) / wf["diag_polys"]
        ksf = (full_wf["ks_baby"] + full_wf["ks_giant"]) / full_wf["diag_polys"]
        # the same measured phases scaled to the full workload's structure (context only; `value` is the measured sample itself)
        t_full = (ph[0] * full_wf["ks_baby"] / wf["ks_baby"] + ph[1] * full_wf["diag_polys"] / wf["diag_polys"]
                  + ph[2] * full_wf["ks_giant"] / wf["ks_giant"])
        desc = ("ONE COMPLETE MatMult4StreamCompute of the oracle port (gwas/matmult.go:1043-1236: baby rotations, K1 lazy MAC under the "
                "per-(row, giant) mutexes, K2 reduce, giant rotations, sum) per step on a %d x %d sample (same CKKS set, s = %d, %d x %d "
                "blocks, %d diagonal polys resident in RAM, %d + %d key-switches), %d threads: %.2f s = baby %.2f + MAC %.2f + giant %.2f; "
                "value = B_alg(sample) / t(sample), nothing extrapolated.  The sample has %.1fx more key-switches per algorithmic byte than "
                "the full workload, so it UNDERSTATES the CPU: scaling the measured phases to the full workload's structure gives %.2f GB/s "
                "(est_full_workload_gbs)." % (self.nrows, self.ncols, self.s, wf["nbr"], wf["m_ct"], wf["diag_polys"], wf["ks_baby"],
                                              wf["ks_giant"], self.nthreads, t, ph[0], ph[1], ph[2], ksb / ksf, full_wf["b_alg"] / t_full / 1e9))
        return desc, full_wf["b_alg"] / t_full / 1e9

    def close(self):
        self.o.cache_free(self.dc)


def run_reference(args, rank, world):
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    full_wf = work_figures(w["nrows"], w["ncols"], w["s"], CKKS[w["params"]]["logN"])
    t_start = time.perf_counter()
    cs = CpuSample(w["params"], w["s"])
    times, phases = [], []
    for it in range(args.warmup + args.steps):
        t, ph = cs.step()
        if it >= args.warmup:
            times.append(t)
            phases.append(ph)
    t = sum(times) / len(times)
    ph = [sum(p[k] for p in phases) / len(phases) for k in range(3)]
    v = cs.wf["b_alg"] / t / 1e9
    desc, est_full = cs.describe(full_wf, t, ph)
    line = dict(metric="genotype x ciphertext MatMult GB/s (B_alg / t)", value=v, unit="GB/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=t * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="u64",
                data="synthetic", impl="reference",
                config=dict(workload=args.workload, ckks_params=w["params"], logN=CKKS[w["params"]]["logN"], s=w["s"], orientation=w["orientation"],
                            max_level=5, num_block_rows=full_wf["nbr"], m_ct=full_wf["m_ct"], diag_polys=full_wf["diag_polys"],
                            b_alg_bytes=full_wf["b_alg"], mac_alg=full_wf["mac_alg"], key_switches=[full_wf["ks_baby"], full_wf["ks_giant"]],
                            note="CPU arm (rank 0 only): a complete call on a bounded sample of the workload, see cpu_baseline.sample",
                            sample_rows=cs.nrows, sample_cols=cs.ncols, sample_b_alg_bytes=cs.wf["b_alg"], preprocess_s=cs.t_prep),
                cpu_baseline=dict(value=v, unit="GB/s", cores=cs.nthreads, kind="port", sample=desc, est_full_workload_gbs=est_full),
                e2e=dict(value=v, unit="GB/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                wall_s=time.perf_counter() - t_start)
    cs.close()
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    """Our arm.  The measurement itself is _run_ours(); this wrapper owns the teardown ORDER: torch's current stream is the library's
    stream during the run, so every tensor has to be released (the frame of _run_ours is gone) and torch given its own stream back
    BEFORE the library context -- and with it that stream -- is destroyed.  (Freeing a tensor that NCCL has touched queries the current
    stream; with the stream already destroyed every rank of an N > 1 run aborted at exit with "CUDA error: context is destroyed" after
    the JSON line was out -- a non-zero exit code under torchrun.)"""
    import gc

    import torch
    import torch.distributed as dist

    keep = {}
    try:
        _run_ours(args, rank, local_rank, world, keep)
    finally:
        try:
            torch.cuda.synchronize()
            torch.cuda.set_stream(torch.cuda.default_stream(torch.device("cuda", local_rank)))
        except Exception:
            pass
        gc.collect()
        if world > 1 and dist.is_initialized():
            dist.destroy_process_group()
        if keep.get("cps") is not None:
            keep["cps"].close()


def _run_ours(args, rank, local_rank, world, keep):
    import numpy as np
    import torch
    import torch.distributed as dist

    from sfgwas_b200 import CryptoParams
    from sfgwas_b200.dist import ct_mod_allreduce_, partition

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = WORKLOADS[args.workload]
    nrows, ncols, s, pname = w["nrows"], w["ncols"], w["s"], w["params"]
    P = CKKS[pname]
    wf = work_figures(nrows, ncols, s, P["logN"])
    N, slots, d, nbr, m_ct = wf["N"], wf["slots"], wf["d"], wf["nbr"], wf["m_ct"]
    cps = keep["cps"] = CryptoParams(P["logN"], P["Q"], P["P"], P["scale"], device=local_rank)
    L = cps.L
    mods = P["Q"] + P["P"]
    # The library's stream is NON-BLOCKING: it is not ordered after torch's default stream.  Everything torch produces for the library
    # (the genotype chunks, synthetic keys) is therefore generated ON the library's stream.  (Round 1 generated the genotypes on torch's
    # stream and pushed them unsynchronised: the pushes raced the generator -- harmless for random data, caught by the decrypt check.)
    ext = torch.cuda.ExternalStream(L.sfg_ctx_stream(cps.h), device=dev)
    torch.cuda.set_stream(ext)
    otf = bool(w.get("otf")) or args.cache_budget_gb is not None
    colshard = (bool(w.get("colshard")) or args.col_sharding) and world > 1
    bc_lo, bc_hi = partition(m_ct, world)[rank] if colshard else (0, m_ct)  # block columns this rank holds and produces
    c_lo, c_hi = bc_lo * slots, min(bc_hi * slots, ncols)
    m_loc = bc_hi - bc_lo
    if otf:
        cps.set_cache_budget(int((args.cache_budget_gb if args.cache_budget_gb is not None else 0.001) * 1e9))

    # ---- inputs: real keys and a real encryption of a known A (identical on every rank: same seeds), or uniformly random residues ----
    real = not args.synthetic_inputs
    t0 = time.perf_counter()
    if real:
        o, sk, keys, Ap, A_np = real_inputs(P, nrows, s)
        for k, v in keys.items():
            cps.SetRotKey(k, v)
        del keys
        d_A = torch.from_numpy(A_np.view(np.int64)).to(dev)
    else:
        gen = torch.Generator(device=dev)
        gen.manual_seed(1234)

        def rand_res(shape_prefix, limb_ids):
            out = torch.empty(*shape_prefix, len(limb_ids), N, dtype=torch.int64, device=dev)
            for k, li in enumerate(limb_ids):
                out[..., k, :] = torch.randint(0, mods[li], (*shape_prefix, N), generator=gen, device=dev, dtype=torch.int64)
            return out

        # Galois keys for the BSGS rotations (crypto/crypto.go:251-264): left rotations 1..d-1 and d, 2d, ...
        for k in sorted(set(range(1, d)) | {g * d for g in range(1, d) if g * d < slots}):
            key = rand_res((cps.beta, 2), list(range(cps.nQP)))
            cps._check(L.sfg_ctx_set_rotation_key(cps.h, k, C.c_void_p(key.data_ptr())), "set_rotation_key")
        d_A = rand_res((s, nbr, 2), list(range(6)))
    t_inputs = time.perf_counter() - t0

    # ---- genotype matrix on the device (same on every rank), pushed through the ABI in row chunks; the block columns used by the
    #      decrypt check are kept on the host ----
    g = C.c_void_p()
    cps._check(L.sfg_geno_create(cps.h, nrows, c_hi - c_lo, C.byref(g)), "geno_create")
    gx = torch.Generator(device=dev)
    gx.manual_seed(1)
    maf = torch.rand(ncols, generator=gx, device=dev) * 0.45 + 0.05
    picks = sorted({(0, 0), (s // 2, m_ct // 2), (s - 1, m_ct - 1)})
    keeper = (lambda bj: bc_lo <= bj < bc_hi) if colshard else (lambda bj: rank == 0)  # who checks (and so keeps) a block column
    xcols = {bj: [] for _, bj in picks if real and keeper(bj)}
    for r0 in range(0, nrows, 512):
        r1 = min(nrows, r0 + 512)
        x = (torch.rand(r1 - r0, ncols, generator=gx, device=dev) < maf).to(torch.int8) + \
            (torch.rand(r1 - r0, ncols, generator=gx, device=dev) < maf).to(torch.int8)
        xl = x[:, c_lo:c_hi].contiguous() if colshard else x
        cps._check(L.sfg_geno_push_rows(g, C.c_void_p(xl.data_ptr()), r1 - r0), "geno_push_rows")
        for bj in xcols:
            xcols[bj].append(x[:, bj * slots:(bj + 1) * slots].cpu().numpy())  # int8: converted when the check runs
    del x, xl
    xcols = {bj: np.concatenate(v) for bj, v in xcols.items()}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cache = C.c_void_p()
    if world > 1 and not colshard:
        cps._check(L.sfg_matmult4_stream_preprocess_giants(cps.h, g, 5, rank, world, C.byref(cache)), "preprocess_giants")
    else:
        cps._check(L.sfg_matmult4_stream_preprocess(cps.h, g, 5, C.byref(cache)), "preprocess")
    t_prep = time.perf_counter() - t0
    npoly, cbytes, mat = C.c_size_t(), C.c_size_t(), C.c_int()
    L.sfg_cache_info(cache, C.byref(npoly), C.byref(cbytes), C.byref(mat), None, None)
    assert colshard or npoly.value == wf["diag_polys"], (npoly.value, wf["diag_polys"])
    assert bool(mat.value) != otf or cbytes.value == 0, "cache materialisation does not match the workload (otf=%s)" % otf

    d_out = torch.zeros(s, m_loc, 2, 5, N, dtype=torch.int64, device=dev)
    h_A = torch.empty(d_A.shape, dtype=torch.int64, pin_memory=True)
    h_A.copy_(d_A)
    h_out = torch.empty(d_out.shape, dtype=torch.int64, pin_memory=True)
    row_lo, row_hi = partition(s, world)[rank]  # e2e at N > 1: every rank reads back its stripe of output rows

    def compute_dev():
        cps._check(L.sfg_matmult4_stream_compute_dev(cps.h, C.c_void_p(d_A.data_ptr()), s, nbr, 5, 5, cache, C.c_void_p(d_out.data_ptr())),
                   "compute_dev")

    R_buf = [None]
    step_phases = dict(baby_ms=0.0, mac_ms=0.0, giant_ms=0.0, mac_kernel_ms=0.0)

    def add_timings():
        t = cps.last_timings()
        for kx in step_phases:
            step_phases[kx] += t[kx]

    def step_dev():
        for kx in step_phases:
            step_phases[kx] = 0.0
        if world == 1 or args.no_baby_sharding or s > 16:
            compute_dev()
            add_timings()
        else:
            # baby-step sharding: 1/world of the rotation-cache entries per rank, ONE all-gather over NVLink, then MAC + giant-step sums
            chunk = int(L.sfg_matmult4_baby_chunk_bytes(cps.h, cache, s, world))
            if R_buf[0] is None:
                if colshard:  # every rank's own cache must see the same baby steps (true when each holds a full-width block column)
                    cmm = torch.tensor([chunk, -chunk], device=dev)
                    dist.all_reduce(cmm, op=dist.ReduceOp.MAX)
                    assert int(cmm[0]) == chunk == -int(cmm[1]), "ranks disagree on the rotation-cache share"
                R_buf[0] = torch.empty(world * chunk, dtype=torch.uint8, device=dev)
            R = R_buf[0]
            cps._check(L.sfg_matmult4_baby_dev(cps.h, C.c_void_p(d_A.data_ptr()), s, nbr, 5, 5, cache, rank, world, C.c_void_p(R.data_ptr())),
                       "baby_dev")
            add_timings()
            mine = R[rank * chunk:(rank + 1) * chunk].clone()
            dist.all_gather_into_tensor(R, mine)
            torch.cuda.current_stream().synchronize()
            cps._check(L.sfg_matmult4_stream_compute_r_dev(cps.h, C.c_void_p(R.data_ptr()), s, 5, cache, C.c_void_p(d_out.data_ptr())),
                       "compute_r_dev")
            add_timings()
        if world > 1 and not colshard:  # partial sums over this rank's giant steps -> the full product on every rank (modular-add all-reduce)
            ct_mod_allreduce_(d_out, cps, 5)

    def step_e2e():
        if world == 1:
            cps._check(L.sfg_matmult4_stream_compute(cps.h, C.c_void_p(h_A.data_ptr()), s, nbr, 5, 5, cache, C.c_void_p(h_out.data_ptr())),
                       "compute")
            return
        d_A.copy_(h_A, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        step_dev()
        if colshard:
            h_out.copy_(d_out, non_blocking=True)
        elif row_hi > row_lo:
            h_out[row_lo:row_hi].copy_(d_out[row_lo:row_hi], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    # the entry point the cgo shim binds: one pageable array per limb for A and for the result
    a_limbs = o_limbs = None
    if world == 1:
        a_host = h_A.numpy().reshape(-1, N)
        a_limbs = [np.array(a_host[k], copy=True) for k in range(a_host.shape[0])]
        o_limbs = [np.empty(N, dtype=np.int64) for _ in range(s * m_ct * 2 * 5)]
        a_ptrs = (C.c_void_p * len(a_limbs))(*[x.ctypes.data for x in a_limbs])
        o_ptrs = (C.c_void_p * len(o_limbs))(*[x.ctypes.data for x in o_limbs])

    def step_ptrs():
        cps._check(L.sfg_matmult4_stream_compute_ptrs(cps.h, a_ptrs, s, nbr, 5, 5, cache, o_ptrs), "compute_ptrs")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sample_clocks=False):
        barrier()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = cps.launch_count()
        mac_ms, phases = 0.0, dict(baby_ms=0.0, mac_ms=0.0, giant_ms=0.0)
        e0.record(ext)
        for _ in range(steps):
            fn()
            t = step_phases if fn is step_dev or world > 1 else cps.last_timings()
            mac_ms += t["mac_kernel_ms"]
            for kx in phases:
                phases[kx] += t[kx]
        e1.record(ext)
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, cps.launch_count() - launches0, mac_ms, phases, clocks

    # torch copies and NCCL collectives are ordered on (and timed with) the library's stream (set as torch's current stream above)
    for _ in range(args.warmup):
        step_dev()
    ms, launches, mac_ms, phases, clocks = timed(step_dev, args.steps, sample_clocks=True)
    for _ in range(max(1, args.warmup - 2)):
        step_e2e()
    ms_e2e, _, _, _, _ = timed(step_e2e, args.steps)
    ms_ptrs = None
    if world == 1:
        for _ in range(max(1, args.warmup - 2)):
            step_ptrs()
        ms_ptrs, _, _, _, _ = timed(step_ptrs, args.steps)
    barrier()
    enc_stats = cps.encoder_stats()

    # ---- correctness of what was just timed (outside the timed region) ----
    verify = None
    if colshard:  # every rank decrypt-checks the picked outputs among its own block columns; rank 0 merges
        mine = None
        if real:
            out_np = h_out.numpy().view(np.uint64)
            mp = [(i, bj) for i, bj in picks if keeper(bj)]
            mine = verify_decrypt(o, sk, Ap, {bj - bc_lo: v for bj, v in xcols.items()}, out_np, [(i, bj - bc_lo) for i, bj in mp])
            mine["checked"] = [list(p) for p in mp]
            mine["dev_equals_e2e"] = bool((d_out.cpu().numpy().view(np.uint64) == out_np).all())
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        if real and rank == 0:
            verify = dict(decrypt_max_abs_err=max(p["decrypt_max_abs_err"] for p in parts),
                          tolerance=min(p["tolerance"] for p in parts if p["checked"]),  # each rank's own: 1e-3 of ITS largest reference value
                          max_abs_reference=max(p["max_abs_reference"] for p in parts), checked=sum((p["checked"] for p in parts), []),
                          ok=all(p["ok"] for p in parts), dev_equals_e2e=all(p["dev_equals_e2e"] for p in parts))
            if not (verify["ok"] and verify["dev_equals_e2e"] and len(verify["checked"]) == len(picks)):
                print(json.dumps(dict(error="verification failed", verify=verify)), flush=True)
                sys.exit(1)
    elif world > 1:  # the host copy is striped over the ranks: gather the stripes on rank 0 for the check
        h_full = d_out.cpu()
    if real and rank == 0 and not colshard:
        out_np = (h_out if world == 1 else h_full).numpy().view(np.uint64)
        verify = verify_decrypt(o, sk, Ap, xcols, out_np, picks)
        verify["dev_equals_e2e"] = bool((d_out.cpu().numpy().view(np.uint64) == out_np).all())
        if world == 1:
            verify["ptrs_equals_flat"] = bool((np.stack(o_limbs).reshape(out_np.shape).view(np.uint64) == out_np).all())
        if not (verify["ok"] and verify["dev_equals_e2e"] and verify.get("ptrs_equals_flat", True)):
            print(json.dumps(dict(error="verification failed", verify=verify)), flush=True)
            sys.exit(1)

    if rank == 0:
        peak, peak_src = read_peaks()
        t_step = ms / args.steps / 1e3
        value = wf["b_alg"] / t_step / 1e9
        e2e_v = wf["b_alg"] / (ms_e2e / args.steps / 1e3) / 1e9
        mac_avg_s = mac_ms / args.steps / 1e3
        # the roofline kernel streams this rank's share of the diagonals (1 / world of them under giant-step sharding)
        mac_gbs = wf["b_diag"] / world / mac_avg_s / 1e9
        traffic, traffic_src = read_traffic()
        ph = {k: v / args.steps for k, v in phases.items()}
        if otf:
            ph["encode_ms"] = ph["mac_ms"] - mac_ms / args.steps  # diagonal re-encoding + image build inside the MAC phase
        line = dict(
            metric="genotype x ciphertext MatMult GB/s (B_alg / t)", value=value, unit="GB/s", n_gpus=world, steps=args.steps,
            warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="u64",
            data="synthetic",
            config=dict(workload=args.workload, ckks_params=pname, logN=P["logN"], s=s, orientation=w["orientation"], max_level=5,
                        num_block_rows=nbr, m_ct=m_ct, diag_polys=wf["diag_polys"], b_alg_bytes=wf["b_alg"], mac_alg=wf["mac_alg"],
                        key_switches=[wf["ks_baby"], wf["ks_giant"]],
                        step="MatMult4StreamCompute over the HBM-resident diagonal cache" if not otf else
                             "MatMult4StreamCompute with the diagonals re-encoded from the int8 genotypes inside every call (cache over budget)",
                        cache_bytes=cbytes.value, cache_materialised=bool(mat.value), preprocess_s=t_prep, input_setup_s=t_inputs,
                        hbm_in_use_gb=round((lambda fr, tot: (tot - fr) / 1e9)(*torch.cuda.mem_get_info()), 1),
                        inputs="real encryptions of a known matrix under seeded keys (decrypt-checked, see verify)" if real else
                               "uniformly random residues",
                        l2_policy="inputs (%.1f GB cache per rank) larger than L2; no flush needed" % (cbytes.value / 1e9) if not otf else
                                  "every call streams freshly encoded diagonals (GBs per chunk) through HBM; larger than L2",
                        sharding=("SNP-block (block-column) sharding of ONE product over %d ranks: each holds %d-%d of the %d block columns of the "
                                  "int8 matrix and produces those columns of the output (no all-reduce); baby-step rotations sharded 1/%d "
                                  "per rank + one all-gather of the rotation cache inside the timed region"
                                  % (world, m_ct // world, -(-m_ct // world), m_ct, world)) if colshard else
                                 ("giant-step sharding of ONE product over %d ranks: 1/%d of the cache, MAC and giant-step key-switches per "
                                  "rank, baby-step rotations %s, modular-add all-reduce of the %d output ciphertexts (NCCL SUM + mod q) inside "
                                  "the timed region" % (world, world, "replicated" if args.no_baby_sharding else
                                                        "sharded 1/%d per rank + one all-gather of the rotation cache" % world, s * m_ct)) if world > 1 else "single GPU",
                        encoder_rechecked_coeffs=enc_stats[0], encoder_unresolved=enc_stats[1]),
            e2e=dict(value=e2e_v, unit="GB/s", ms_per_step=ms_e2e / args.steps, h2d_bytes_per_step=int(h_A.numel() * 8),
                     d2h_bytes_per_step=int(h_out.numel() * 8) if world == 1 or colshard else int(h_out[row_lo:row_hi].numel() * 8),
                     path="sfg_matmult4_stream_compute (pinned host A and out, D2H overlapped with the giant-step sums)" if world == 1 else
                          "per rank: pinned A -> device, compute, this rank's block columns of the output -> pinned host" if colshard else
                          "per rank: pinned A -> device, compute + all-reduce, this rank's stripe of output rows -> pinned host"),
            gpu_launches=int(launches),
            roofline=dict(bound="hbm", kernel="k_mac_tc (K1+K2: tcgen05 kind::i8 byte-plane MAC + recombine + modular reduce)",
                          achieved=mac_gbs, peak=peak, unit="GB/s", frac=mac_gbs / peak,
                          traffic=traffic if args.workload == "mm_10k_x_100k_k10_logN13" and world == 1 else None,  # the capture is of the default workload
                          traffic_source=traffic_src, peak_source=peak_src + " (MEASURED_PEAKS.json hbm_gbs)",
                          algorithmic_bytes_per_launch=wf["b_diag"] // world, stored_bytes_per_launch=int(cbytes.value),
                          avg_launch_ms=mac_avg_s * 1e3, share_of_step=mac_ms / ms,
                          note="algorithmic bytes count 8 B per cached residue (the reference's streamed volume); the image stores "
                               "4-6 byte planes per residue, so achieved can exceed the copy peak; stored+written bytes / time is "
                               "the physical HBM rate"),
            phases_ms_per_step=ph,
            whole_step_hbm_frac=value / world / peak,
            int8_genotype_gbs=nrows * ncols / t_step / 1e9, gmacs_per_s=wf["mac_alg"] / t_step / 1e9,
            clocks=clocks, verify=verify,
        )
        if ms_ptrs is not None:
            line["e2e_ptrs"] = dict(value=wf["b_alg"] / (ms_ptrs / args.steps / 1e3) / 1e9, unit="GB/s", ms_per_step=ms_ptrs / args.steps,
                                    path="sfg_matmult4_stream_compute_ptrs: one PAGEABLE host array per limb (%d in, %d out), the Go-side figure"
                                         % (len(a_limbs), len(o_limbs)))
        if world == 1 and not args.no_cpu_baseline:
            cs = CpuSample(pname, s)
            t, phs = cs.step()
            desc, est_full = cs.describe(wf, t, phs)
            line["cpu_baseline"] = dict(value=cs.wf["b_alg"] / t / 1e9, unit="GB/s", cores=cs.nthreads, kind="port", sample=desc,
                                        est_full_workload_gbs=est_full)
            cs.close()
        print(json.dumps(line), flush=True)
    L.sfg_cache_destroy(cache)
    L.sfg_geno_destroy(g)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mm_10k_x_100k_k10_logN13", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--synthetic-inputs", action="store_true", help="uniformly random residues instead of real keys / encryptions (no decrypt check)")
    ap.add_argument("--no-baby-sharding", action="store_true", help="N > 1: every rank repeats all baby-step rotations (no all-gather)")
    ap.add_argument("--col-sharding", action="store_true", help="N > 1: SNP-block sharding (each rank holds and produces its block columns) instead of giant-step sharding")
    ap.add_argument("--cache-budget-gb", type=float, default=None, help="HBM budget of the diagonal cache; below the image size the diagonals are re-encoded on the fly")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
