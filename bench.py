#!/usr/bin/env python3
"""bench.py -- genotype x CKKS-ciphertext MatMult throughput (BASELINE.json metric) on B200.

  python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU algorithm (oracle port) on the host cores

Workload (BASELINE.json configs[1]): X = 10 000 samples x 100 000 SNPs int8 (Binomial(2, p_j), p_j ~ U(0.05, 0.5)),
A = 10 rows x 3 ciphertexts at level 5, PN13QP218 (logN = 13, scale 2^30), orientation A.X -> 10 x 25 ciphertexts at level 4.
A "step" is one MatMult4StreamCompute call over the preprocessed, HBM-resident diagonal cache (the reference calls
Compute ~40x per Preprocess inside the PCA power iterations, gwas/pca.go:112-113,288-352).

metric  = B_alg / t  in GB/s with B_alg = diag_polys*L'*N*8 + s*nbr*2*6*N*8 + s*m_ct*2*L'*N*8
value   = inputs already resident in HBM; e2e = the same call through the C ABI with pinned HOST buffers for A and out.
Synthetic data: ciphertext and Galois-key residues are uniformly random (which is what real ones look like); the
kernels have no data-dependent control flow, so the work is identical.  Multi-GPU: every rank owns its own block of
100 000 SNP columns of a 10 000 x (N*100 000) matrix (SNP-block sharding, no data-path collective) -> weak scaling.
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
    # name: (nrows, ncols, s, CKKS parameter set)
    "mm_10k_x_100k_k10_logN13": (10000, 100000, 10, "PN13QP218"),
    "mm_2k_x_20k_k10_logN13": (2000, 20000, 10, "PN13QP218"),   # quick check only
    # a PCA-shaped block (BASELINE configs 4/5 run at logN 14, kp = 15): extra data point, not the default bench line
    "mm_16k_x_64k_k15_logN14": (16384, 65536, 15, "PN14QP438"),
}


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
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_mac_tc launch from the committed ncu --set full capture
    (profiles/r1_final/ncu_k_mac_tc.txt, same workload); None when the summary is missing."""
    try:
        rd = wr = None
        with open(os.path.join(ROOT, "profiles", "r1_final", "ncu_k_mac_tc.txt")) as f:
            for ln in f:
                t = ln.split()
                if len(t) >= 3 and t[0] == "dram__bytes_read.sum" and rd is None:
                    rd = float(t[1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[t[2]]
                if len(t) >= 3 and t[0] == "dram__bytes_write.sum" and wr is None:
                    wr = float(t[1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[t[2]]
        return None if rd is None or wr is None else rd + wr
    except Exception:
        return None


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
# CPU baseline: the reference's algorithm (oracle port, gwas/matmult.go:1138-1236) on a BOUNDED sample of the workload
# ------------------------------------------------------------------------------------------------------------------
def cpu_sample(wf, s, nthreads, mac_diags=12288, rots_per_thread=12, P=None):
    """Times (i) the K1 lazy-MAC loop with the reference's per-(row, giant) locks on `mac_diags` diagonal polynomials
    resident in RAM, and (ii) level-5 / level-4 rotations (key-switch + automorphism) on all threads; extrapolates linearly
    to the full call.  Returns (GB/s, description)."""
    import numpy as np

    from oracle.oracle import Oracle

    o = Oracle.from_params(P or PN13)
    t_mac = o.L.orc_bench_mac(o.N, 5, s, mac_diags, nthreads)
    mac_rate = mac_diags * s * 2 * 5 * o.N / t_mac
    sk = o.keygen_secret(1)
    swk = o.gen_rotation_key(sk, 1)
    rng = np.random.default_rng(0)
    times = {}
    for level in (5, 4):
        cts = [np.stack([np.stack([rng.integers(0, o.Q[l], o.N, dtype=np.uint64) for l in range(level + 1)]) for _ in range(2)])
               for _ in range(nthreads)]

        def work(t):
            for _ in range(rots_per_thread):
                o.rotate_right(cts[t], -1, swk)

        th = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
        t0 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        times[level] = (time.perf_counter() - t0) / (nthreads * rots_per_thread)  # seconds per rotation at full occupancy
    t_full = wf["mac_alg"] / mac_rate + wf["ks_baby"] * times[5] + wf["ks_giant"] * times[4]
    desc = ("oracle port of MatMult4StreamCompute: K1 lazy MAC on %d of %d diagonal polys (RAM-resident, %d threads, per-(row,giant) "
            "mutex) = %.2f GMAC/s; %d+%d rotations timed (%.1f / %.1f ms each at level 5 / 4 with all threads busy); extrapolated "
            "linearly to %d MAC-polys + %d + %d rotations" % (mac_diags, wf["diag_polys"], nthreads, mac_rate / 1e9,
                                                              nthreads * rots_per_thread, nthreads * rots_per_thread,
                                                              times[5] * 1e3 * nthreads, times[4] * 1e3 * nthreads,
                                                              wf["diag_polys"], wf["ks_baby"], wf["ks_giant"]))
    return wf["b_alg"] / t_full / 1e9, t_full, desc


def run_reference(args, rank, world):
    if rank != 0:
        return
    nrows, ncols, s, pname = WORKLOADS[args.workload]
    wf = work_figures(nrows, ncols, s, CKKS[pname]["logN"])
    nthreads = os.cpu_count() or 1
    vals, t0 = [], time.perf_counter()
    for it in range(args.warmup + args.steps):
        v, t_full, desc = cpu_sample(wf, s, nthreads, P=CKKS[pname])
        if it >= args.warmup:
            vals.append((v, t_full))
    v = sum(x[0] for x in vals) / len(vals)
    t_full = sum(x[1] for x in vals) / len(vals)
    line = dict(metric="genotype x ciphertext MatMult GB/s (B_alg / t)", value=v, unit="GB/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=t_full * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u64",
                data="synthetic", impl="reference",
                config=dict(workload=args.workload, ckks_params=pname, orientation="A.X", s=s, note="CPU arm: the per-GPU workload timed on the host cores of the box (rank 0 only)"),
                cpu_baseline=dict(value=v, unit="GB/s", cores=nthreads, kind="port", sample=desc),
                e2e=dict(value=v, unit="GB/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                wall_s=time.perf_counter() - t0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist

    from sfgwas_b200 import CryptoParams, SfgError

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nrows, ncols, s, pname = WORKLOADS[args.workload]
    P = CKKS[pname]
    wf = work_figures(nrows, ncols, s, P["logN"])
    N, slots, d, nbr, m_ct = wf["N"], wf["slots"], wf["d"], wf["nbr"], wf["m_ct"]
    cps = CryptoParams(P["logN"], P["Q"], P["P"], P["scale"], device=local_rank)
    L = cps.L
    mods = P["Q"] + P["P"]
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)

    def rand_res(shape_prefix, limb_ids):
        out = torch.empty(*shape_prefix, len(limb_ids), N, dtype=torch.int64, device=dev)
        for k, li in enumerate(limb_ids):
            out[..., k, :] = torch.randint(0, mods[li], (*shape_prefix, N), generator=gen, device=dev, dtype=torch.int64)
        return out

    # Galois keys for the BSGS rotations (crypto/crypto.go:251-264): left rotations 1..d-1 and d, 2d, ...
    rots = sorted(set(range(1, d)) | {g * d for g in range(1, d) if g * d < slots})
    for k in rots:
        key = rand_res((cps.beta, 2), list(range(cps.nQP)))
        cps._check(L.sfg_ctx_set_rotation_key(cps.h, k, C.c_void_p(key.data_ptr())), "set_rotation_key")
    del key
    # genotype matrix on the device, pushed through the ABI in row chunks
    g = C.c_void_p()
    cps._check(L.sfg_geno_create(cps.h, nrows, ncols, C.byref(g)), "geno_create")
    gx = torch.Generator(device=dev)
    gx.manual_seed(1 + rank)
    maf = torch.rand(ncols, generator=gx, device=dev) * 0.45 + 0.05
    for r0 in range(0, nrows, 512):
        r1 = min(nrows, r0 + 512)
        x = (torch.rand(r1 - r0, ncols, generator=gx, device=dev) < maf).to(torch.int8) + \
            (torch.rand(r1 - r0, ncols, generator=gx, device=dev) < maf).to(torch.int8)
        cps._check(L.sfg_geno_push_rows(g, C.c_void_p(x.data_ptr()), r1 - r0), "geno_push_rows")
    del x
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cache = C.c_void_p()
    cps._check(L.sfg_matmult4_stream_preprocess(cps.h, g, 5, C.byref(cache)), "preprocess")
    t_prep = time.perf_counter() - t0
    npoly, cbytes, mat = C.c_size_t(), C.c_size_t(), C.c_int()
    L.sfg_cache_info(cache, C.byref(npoly), C.byref(cbytes), C.byref(mat), None, None)
    assert npoly.value == wf["diag_polys"], (npoly.value, wf["diag_polys"])

    d_A = rand_res((s, nbr, 2), list(range(6)))
    d_out = torch.zeros(s, m_ct, 2, 5, N, dtype=torch.int64, device=dev)
    h_A = torch.empty(d_A.shape, dtype=torch.int64, pin_memory=True)
    h_A.copy_(d_A)
    h_out = torch.empty(d_out.shape, dtype=torch.int64, pin_memory=True)
    ext = torch.cuda.ExternalStream(L.sfg_ctx_stream(cps.h), device=dev)

    def step_dev():
        cps._check(L.sfg_matmult4_stream_compute_dev(cps.h, C.c_void_p(d_A.data_ptr()), s, nbr, 5, 5, cache, C.c_void_p(d_out.data_ptr())),
                   "compute_dev")

    def step_e2e():
        cps._check(L.sfg_matmult4_stream_compute(cps.h, C.c_void_p(h_A.data_ptr()), s, nbr, 5, 5, cache, C.c_void_p(h_out.data_ptr())),
                   "compute")

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
            t = cps.last_timings()
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

    for _ in range(args.warmup):
        step_dev()
    ms, launches, mac_ms, phases, clocks = timed(step_dev, args.steps, sample_clocks=True)
    for _ in range(max(1, args.warmup - 2)):
        step_e2e()
    ms_e2e, _, _, _, _ = timed(step_e2e, args.steps)
    enc_stats = cps.encoder_stats()

    if rank == 0:
        peak, peak_src = read_peaks()
        t_step = ms / args.steps / 1e3
        value = world * wf["b_alg"] / t_step / 1e9
        e2e_v = world * wf["b_alg"] / (ms_e2e / args.steps / 1e3) / 1e9
        mac_avg_s = mac_ms / args.steps / 1e3
        mac_gbs = wf["b_diag"] / mac_avg_s / 1e9
        line = dict(
            metric="genotype x ciphertext MatMult GB/s (B_alg / t)", value=value, unit="GB/s", n_gpus=world, steps=args.steps,
            warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u64",
            data="synthetic",
            config=dict(workload=args.workload, ckks_params=pname, logN=P["logN"], s=s, orientation="A.X", max_level=5,
                        num_block_rows=nbr, m_ct=m_ct, diag_polys=wf["diag_polys"], b_alg_bytes=wf["b_alg"], mac_alg=wf["mac_alg"],
                        key_switches=[wf["ks_baby"], wf["ks_giant"]], step="MatMult4StreamCompute over the HBM-resident diagonal cache",
                        cache_bytes=cbytes.value, cache_materialised=bool(mat.value), preprocess_s=t_prep,
                        l2_policy="inputs (%.1f GB cache) larger than L2; no flush needed" % (cbytes.value / 1e9),
                        sharding="SNP-block (block-column) per rank, no collective" if world > 1 else "single GPU",
                        encoder_rechecked_coeffs=enc_stats[0], encoder_unresolved=enc_stats[1]),
            e2e=dict(value=e2e_v, unit="GB/s", ms_per_step=ms_e2e / args.steps, h2d_bytes_per_step=int(h_A.numel() * 8),
                     d2h_bytes_per_step=int(h_out.numel() * 8)),
            gpu_launches=int(launches),
            roofline=dict(bound="hbm", kernel="k_mac_tc (K1+K2: tcgen05 kind::i8 byte-plane MAC + recombine + modular reduce)",
                          achieved=mac_gbs, peak=peak, unit="GB/s", frac=mac_gbs / peak,
                          traffic=read_traffic() if args.workload == "mm_10k_x_100k_k10_logN13" else None,  # the capture is of the default workload
                          traffic_source="ncu --set full, profiles/r1_final/ncu_k_mac_tc.txt (dram read + write per launch)",
                          peak_source=peak_src + " (MEASURED_PEAKS.json hbm_gbs)",
                          algorithmic_bytes_per_launch=wf["b_diag"], stored_bytes_per_launch=int(cbytes.value),
                          avg_launch_ms=mac_avg_s * 1e3, share_of_step=mac_ms / ms,
                          note="algorithmic bytes count 8 B per cached residue (the reference's streamed volume); the image stores "
                               "4-5 byte planes per residue, so achieved can exceed the copy peak; stored+written bytes / time is "
                               "the physical HBM rate"),
            phases_ms_per_step={k: v / args.steps for k, v in phases.items()},
            int8_genotype_gbs=world * nrows * ncols / t_step / 1e9, gmacs_per_s=world * wf["mac_alg"] / t_step / 1e9,
            clocks=clocks,
        )
        if world == 1 and not args.no_cpu_baseline:
            nthreads = os.cpu_count() or 1
            v, t_full, desc = cpu_sample(wf, s, nthreads, P=P)
            line["cpu_baseline"] = dict(value=v, unit="GB/s", cores=nthreads, kind="port", sample=desc, est_s_per_step=t_full)
        print(json.dumps(line), flush=True)
    L.sfg_cache_destroy(cache)
    L.sfg_geno_destroy(g)
    cps.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mm_10k_x_100k_k10_logN13", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
