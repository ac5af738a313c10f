"""Diagnostic at the BENCHMARKED geometry (config 2: 10 000 x 100 000, s = 10, PN13QP218) with real encryptions: decrypts ALL 250 output
ciphertexts against A.X and compares the row-chunk, single-row, column-subset and on-the-fly variants of the same product bit for bit.
    python profiles/diag_bench_geometry.py [nrows ncols]        (GPU box; ~20 s)"""
import sys, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import bench
from sfgwas_b200 import CryptoParams, GenoFileStream, MatMult4StreamCompute, MatMult4StreamPreprocess
P = bench.PN13
nrows, ncols, s = 10000, 100000, 10
if len(sys.argv) > 1:
    nrows, ncols = int(sys.argv[1]), int(sys.argv[2])
o, sk, keys, Ap, A = bench.real_inputs(P, nrows, s)
cps = CryptoParams(P["logN"], P["Q"], P["P"], P["scale"])
cps.SetRotKeys(keys)
rng = np.random.default_rng(5)
maf = rng.uniform(0.05, 0.5, ncols)
X = (rng.random((nrows, ncols)) < maf).astype(np.int8) + (rng.random((nrows, ncols)) < maf).astype(np.int8)
gfs = GenoFileStream.from_matrix(cps, X)
cache = MatMult4StreamPreprocess(cps, gfs, 5)
out = MatMult4StreamCompute(cps, A, 5, cache)
m_ct = out.shape[1]
ref = Ap @ X.astype(np.float64)
err = np.zeros((s, m_ct))
for i in range(s):
    for bj in range(m_ct):
        w = ref[i, bj * o.slots:(bj + 1) * o.slots]
        got = o.decrypt_vector(sk, out[i, bj], o.scale * o.scale).real[:len(w)]
        err[i, bj] = np.abs(got - w).max()
np.set_printoptions(linewidth=250, precision=2, suppress=False)
print("max |ref|", np.abs(ref).max())
print("decrypt err by (row i, block col bj):")
print(err)
# one row chunk at a time
oa = MatMult4StreamCompute(cps, A[:5], 5, cache)
ob = MatMult4StreamCompute(cps, A[5:], 5, cache)
print("rows 0..4 (s=5 call) == full:", bool((oa == out[:5]).all()), " rows 5..9:", bool((ob == out[5:]).all()))
o1 = MatMult4StreamCompute(cps, A[:1], 5, cache)
print("row 0 alone == full:", bool((o1 == out[:1]).all()))
# a column subset: block columns 0..3 only
sub = X[:, :4 * o.slots].copy()
c2 = MatMult4StreamPreprocess(cps, GenoFileStream.from_matrix(cps, sub), 5)
o2 = MatMult4StreamCompute(cps, A, 5, c2)
print("block cols 0..3 from a 4-column-block matrix == full[:, :4]:", bool((o2 == out[:, :4]).all()))
e2 = max(np.abs(o.decrypt_vector(sk, o2[i, bj], o.scale * o.scale).real[:o.slots] - ref[i, bj * o.slots:(bj + 1) * o.slots]).max() for i in range(s) for bj in range(4))
print("decrypt err of the 4-block-column product:", e2)
# non-materialised
cps.set_cache_budget(1)
c3 = MatMult4StreamPreprocess(cps, gfs, 5)
cps.set_cache_budget(0)
o3 = MatMult4StreamCompute(cps, A, 5, c3)
print("on-the-fly == materialised:", bool((o3 == out).all()))
