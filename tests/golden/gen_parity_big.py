#!/usr/bin/env python3
"""Golden fixtures for Preprocess + Compute parity at the BENCHMARKED geometries.

At the real rings the CPU oracle needs minutes (PN13QP218: 12 ms per encoded diagonal, 50 ms per rotation on one core), too slow to
run inside the GPU test session.  This script runs the oracle HERE (no GPU needed) on seeded inputs and commits only SHA-256 digests of
every output ciphertext (+ a few raw words and cached-diagonal digests); tests/test_gpu_parity.py::test_benchmarked_geometry_golden
regenerates the same seeded inputs on the GPU box, runs the CUDA path and compares digests -- bit-exactness of all of `out`.

Cases (shapes chosen to exercise what the toy rings cannot):
  pn13_2x4_s10 : PN13QP218, X = (4096+40) x (3*4096+30), s = 10  -> 2 block rows, 4 block columns, K = 128, 256 columns = 2 FULL
                 128-column tiles, 63 giant steps in 21-giant chunks, two row chunks of 5, narrow + FP64-class key-switch kernels.
  pn14_2x2_s15 : PN14QP438, X = (8192+40) x (8192+30), s = 15    -> two-half tensor-core MAC (30 rows), half-ring key-switch CTAs,
                 k_ks_extd / k_md_extp (nP = 2), 182 columns.

    python tests/golden/gen_parity_big.py [case ...]        # writes tests/golden/parity_<case>.json
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CASES = {
    "pn13_2x4_s10": dict(params="PN13QP218", nr=4096 + 40, nc=3 * 4096 + 30, s=10, seed=2024),
    "pn14_2x2_s15": dict(params="PN14QP438", nr=8192 + 40, nc=8192 + 30, s=15, seed=2025),
}


def inputs(o, case):
    """Seeded inputs, regenerated identically by the GPU test: secret key, BSGS keys, genotypes, encrypted A."""
    rng = np.random.default_rng(case["seed"])
    nr, nc, s = case["nr"], case["nc"], case["s"]
    maf = rng.uniform(0.05, 0.5, nc)
    X = (rng.random((nr, nc)) < maf).astype(np.int8) + (rng.random((nr, nc)) < maf).astype(np.int8)
    Ap = rng.normal(size=(s, nr))
    sk = o.keygen_secret(case["seed"])
    keys = o.gen_bsgs_keys(sk)
    nbr = (nr - 1) // o.slots + 1
    A = np.zeros((s, nbr, 2, 6, o.N), dtype=np.uint64)
    for i in range(s):
        for b in range(nbr):
            A[i, b] = o.encrypt_vector(sk, Ap[i, b * o.slots:(b + 1) * o.slots], 5, seed=100 + 17 * i + b)
    return X, Ap, sk, keys, A


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<u8").tobytes()).hexdigest()


def diag_probes(o, nbr, m_ct):
    """(bi, shift, bj) of the cached plaintexts whose digests are pinned."""
    out = []
    for bi in range(nbr):
        for shift in (0, 1, o.d - 1, o.d, 39, o.slots // 2 + 3, o.slots - 1):
            for bj in (0, m_ct - 1):
                out.append((bi, shift, bj))
    return out


def main():
    from oracle.oracle import PARAMS, Oracle

    names = sys.argv[1:] or list(CASES)
    nproc = os.cpu_count() or 1
    for name in names:
        case = CASES[name]
        o = Oracle.from_params(PARAMS[case["params"]])
        t0 = time.time()
        X, Ap, sk, keys, A = inputs(o, case)
        t1 = time.time()
        dc = o.preprocess(X, 5, nproc=nproc)
        t2 = time.time()
        out = o.compute(A, dc, keys, 5, nproc=nproc)
        t3 = time.time()
        npolys = int(o.L.orc_diag_cache_num_polys(dc))
        s, m_ct = out.shape[:2]
        nbr = A.shape[1]
        diags = {}
        for bi, shift, bj in diag_probes(o, nbr, m_ct):
            d = o.cache_get(dc, bi, shift, bj)
            diags["%d,%d,%d" % (bi, shift, bj)] = None if d is None else digest(d[:5])
        o.cache_free(dc)
        # numerical meaning, checked here once: decrypt(out) ~= A_plain . X
        ref = Ap @ X.astype(float)
        errs = []
        for i, bj in ((0, 0), (s - 1, m_ct - 1)):
            got = o.decrypt_vector(sk, out[i, bj], o.scale * o.scale).real
            w = ref[i, bj * o.slots:(bj + 1) * o.slots]
            errs.append(float(np.abs(got[:len(w)] - w).max()))
        assert max(errs) < 1e-3 * max(1.0, float(np.abs(ref).max())), errs
        rec = dict(case=case, oracle_threads=nproc, seconds=dict(inputs=t1 - t0, preprocess=t2 - t1, compute=t3 - t2),
                   shape=list(out.shape), num_polys=npolys,
                   input_digests=dict(X=hashlib.sha256(X.tobytes()).hexdigest(), A=digest(A)),
                   out_sha256=[[digest(out[i, bj]) for bj in range(m_ct)] for i in range(s)],
                   out_first_words=[[out[i, bj, c, l, :4].tolist() for c in range(2) for l in range(5)] for i, bj in ((0, 0), (s - 1, m_ct - 1))],
                   diag_sha256=diags, decrypt_max_abs_err=errs)
        with open(os.path.join(HERE, "parity_%s.json" % name), "w") as f:
            json.dump(rec, f, indent=1)
        print(name, "inputs %.0fs preprocess %.0fs compute %.0fs decrypt err %s" % (t1 - t0, t2 - t1, t3 - t2, errs), flush=True)


if __name__ == "__main__":
    main()
