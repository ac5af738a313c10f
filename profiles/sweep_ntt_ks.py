#!/usr/bin/env python3
"""BASELINE.json config 3: NTT / INTT + rotation (hybrid key-switch + automorphism) throughput sweep, logN 12-16, full RNS modulus
chain, a batch of uniformly random top-level ciphertexts resident in HBM.  Device time by CUDA events on the context's stream, 3 warm-up
+ 5 timed repetitions (the batch is larger than L2 for logN >= 13; stated per line).  Parity of exactly these shapes is
tests/test_gpu_parity.py::test_sweep_full_chain_ntt_and_rotation.

    python profiles/sweep_ntt_ks.py [--out profiles/r1_final/sweep_ntt_ks.json]

NTT bytes = 16 B per coefficient (read + write, 8-byte residues); a rotation moves 2 limbs-vectors in and out plus the key once per
ciphertext (beta * 2 * (nQ + nP) * N * 8 bytes), which is what bounds it from memory; both are reported next to the measured HBM peak."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from oracle.oracle import sweep_params  # parameter generator only (prime search); nothing of the oracle is timed or used as a result
    from sfgwas_b200 import CryptoParams

    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2", "sweep_ntt_ks.json"))
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    rows = []
    for logN in (12, 13, 14, 15, 16):
        p = sweep_params(logN)
        cps = CryptoParams(p["logN"], p["Q"], p["P"], p["scale"])
        L, N, nQ, nP = cps.L, cps.N, cps.nQ, cps.nP
        mods = p["Q"] + p["P"]
        top = nQ - 1
        ct_bytes = 2 * nQ * N * 8
        batch = max(8, min(1024, (1 << 30) // ct_bytes))  # ~1 GiB of ciphertexts, 64..1024 at the sweep's sizes
        gen = torch.Generator(device=dev)
        gen.manual_seed(logN)

        def rand_res(prefix, limbs):
            out = torch.empty(*prefix, len(limbs), N, dtype=torch.int64, device=dev)
            for k, li in enumerate(limbs):
                out[..., k, :] = torch.randint(0, mods[li], (*prefix, N), generator=gen, device=dev, dtype=torch.int64)
            return out

        cts = rand_res((batch, 2), list(range(nQ)))
        out = torch.empty_like(cts)
        torch.cuda.synchronize()  # the library's stream is non-blocking: it is not ordered after torch's generator kernels
        ext = torch.cuda.ExternalStream(L.sfg_ctx_stream(cps.h), device=dev)
        idx = (C.c_int * nQ)(*range(nQ))

        def timed(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            for _ in range(args.reps):
                fn()
            e1.record(ext)
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / args.reps

        def ntt(inv):
            cps._check(L.sfg_ntt_dev(cps.h, C.c_void_p(cts.data_ptr()), batch * 2 * nQ, idx, nQ, inv), "sfg_ntt_dev")

        ms_f = timed(lambda: ntt(0))
        ms_i = timed(lambda: ntt(1))
        row = dict(logN=logN, nQ=nQ, nP=nP, batch=batch, batch_bytes=batch * ct_bytes, larger_than_l2=batch * ct_bytes > 126e6,
                   ntt_ms=ms_f, intt_ms=ms_i, ntt_polys_per_s=batch * 2 * nQ / ms_f * 1e3, intt_polys_per_s=batch * 2 * nQ / ms_i * 1e3,
                   ntt_gbs=batch * ct_bytes * 2 / ms_f / 1e6, intt_gbs=batch * ct_bytes * 2 / ms_i / 1e6)
        row["ntt_frac_hbm"], row["intt_frac_hbm"] = row["ntt_gbs"] / peak, row["intt_gbs"] / peak
        d = cps.d
        for name, k in (("rot1", 1), ("rotd", d)):
            key = rand_res((cps.beta, 2), list(range(cps.nQP)))
            torch.cuda.synchronize()
            cps._check(L.sfg_ctx_set_rotation_key(cps.h, k, C.c_void_p(key.data_ptr())), "set_rotation_key")
            del key
            ms = timed(lambda: cps._check(L.sfg_rotate_right_dev(cps.h, top, C.c_void_p(cts.data_ptr()), batch, -k, C.c_void_p(out.data_ptr())),
                                          "sfg_rotate_right_dev"))
            row[name + "_ms"] = ms
            row[name + "_per_s"] = batch / ms * 1e3
            row[name + "_us_each"] = ms * 1e3 / batch
        row["key_bytes"] = cps.beta * 2 * cps.nQP * N * 8
        row["keyswitch_path"] = "fused (one CTA per ciphertext x target modulus)" if logN <= 14 else "unfused large-ring kernels around the four-step transforms"
        rows.append(row)
        print(json.dumps(row), flush=True)
        del cts, out
        cps.close()
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(dict(hbm_peak_gbs=peak, rows=rows), f, indent=1)
    txt = ["# NTT / INTT / rotation sweep (BASELINE config 3), B200, device time (CUDA events), batch resident in HBM",
           "%5s %3s %2s %6s %10s %9s %9s %8s %8s %12s %12s" % ("logN", "nQ", "nP", "batch", "NTT poly/s", "NTT GB/s", "INTT GB/s", "frac", "frac", "rot(1)/s", "rot(d)/s")]
    for r in rows:
        txt.append("%5d %3d %2d %6d %10.3g %9.0f %9.0f %8.2f %8.2f %12.0f %12.0f" % (r["logN"], r["nQ"], r["nP"], r["batch"], r["ntt_polys_per_s"], r["ntt_gbs"],
                                                                                  r["intt_gbs"], r["ntt_frac_hbm"], r["intt_frac_hbm"], r["rot1_per_s"], r["rotd_per_s"]))
    with open(args.out.replace(".json", ".txt"), "w") as f:
        f.write("\n".join(txt) + "\n")
    print("\n".join(txt))


if __name__ == "__main__":
    main()
