"""Multi-GPU host logic (sfgwas_b200/dist.py): partitions and the modular-add all-reduce on CPU with gloo (world size 2),
and the two shardings against the single-GPU result on 2 GPUs (skipped on boxes with fewer)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_partition_covers_everything_once():
    from sfgwas_b200.dist import ColumnSharded, partition

    for n in (0, 1, 3, 25, 62, 63):
        for world in (1, 2, 3, 4, 8):
            r = partition(n, world)
            assert len(r) == world and r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    # SNP-block sharding: the ranks' column ranges tile [0, ncols) at block-column granularity
    for ncols, slots, world in ((100000, 4096, 8), (520, 256, 2), (5, 256, 4)):
        cov = [ColumnSharded.local_columns(ncols, slots, r, world) for r in range(world)]
        assert cov[0][0] == 0 and cov[-1][1] == ncols
        assert all(a[1] == b[0] for a, b in zip(cov, cov[1:]))
        assert all(c0 % slots == 0 for c0, c1 in cov if c1 > c0)


def _gloo_worker(rank, world, port, moduli, N, seed, q):
    import torch
    import torch.distributed as dist

    from sfgwas_b200.dist import giant_share, mod_allreduce_, mod_reduce_scatter_

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = len(moduli)
        parts = []
        for r in range(world):  # every rank can regenerate every rank's input
            rng = np.random.default_rng(seed + r)
            parts.append(np.stack([rng.integers(0, m, (3, N), dtype=np.uint64) for m in moduli], axis=1))  # [3][L][N]
        t = torch.from_numpy(parts[rank].view(np.int64).copy())
        mod_allreduce_(t, moduli, N)
        want = np.zeros_like(parts[0])
        for l, m in enumerate(moduli):
            acc = np.zeros((3, N), dtype=object)
            for p in parts:
                acc = acc + p[:, l, :].astype(object)
            want[:, l, :] = (acc % m).astype(np.uint64)
        ok = bool((t.numpy().view(np.uint64) == want).all())
        # modular-add reduce-scatter of an accumulator image of ng = 3 "giants" ([3][L][N] each rank): shares of giant_share(3, 2) = 2
        # giants, zero-padded tail (RowSharded's combination between the MAC and the giant-step rotations)
        share = giant_share(3, world)
        per_g = L * N
        padded = torch.zeros(world * share * per_g, dtype=torch.int64)
        padded[: 3 * per_g] = torch.from_numpy(parts[rank].view(np.int64).reshape(-1).copy())
        mine = mod_reduce_scatter_(padded, share * per_g, group=None, moduli=moduli, N=N).numpy().view(np.uint64).reshape(share, L, N)
        g_lo, g_hi = min(3, rank * share), min(3, (rank + 1) * share)
        ok = ok and bool((mine[: g_hi - g_lo] == want[g_lo:g_hi]).all()) and bool((mine[g_hi - g_lo:] == 0).all())
        q.put((rank, ok, int(L)))
    finally:
        dist.destroy_process_group()


def test_mod_allreduce_gloo_world2():
    import torch.multiprocessing as mp

    moduli = [0x1FFFEC001, 0x3FFF4001, 0x3FFE8001, 0x40020001, (1 << 56) - 5]  # up to 56-bit residues
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, moduli, 64, 11, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1] and all(r[1] for r in res)


def _gpu_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    from oracle.oracle import Oracle, small_params
    from sfgwas_b200 import CryptoParams, GenoFileStream, MatMult4StreamCompute, MatMult4StreamPreprocess
    from sfgwas_b200.dist import ColumnSharded, GiantSharded, RowSharded

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        p = small_params(9, "pn13")
        o = Oracle.from_params(p)
        cps = CryptoParams(p["logN"], p["Q"], p["P"], p["scale"], device=rank)
        sk = o.keygen_secret(1)
        cps.SetRotKeys(o.gen_bsgs_keys(sk))
        rng = np.random.default_rng(3)
        nr, nc, s = 700, 900, 3
        X = rng.integers(0, 3, (nr, nc)).astype(np.int8)
        Ap = rng.normal(size=(s, nr))
        nbr = (nr - 1) // o.slots + 1
        A = np.zeros((s, nbr, 2, 6, o.N), dtype=np.uint64)
        for i in range(s):
            for b in range(nbr):
                A[i, b] = o.encrypt_vector(sk, Ap[i, b * o.slots:(b + 1) * o.slots], 5, seed=5 + 7 * i + b)
        single = MatMult4StreamCompute(cps, A, 5, MatMult4StreamPreprocess(cps, GenoFileStream.from_matrix(cps, X), 5))
        c0, c1 = ColumnSharded.local_columns(nc, o.slots, rank, world)
        cs = ColumnSharded(cps, X[:, c0:c1], nc, rank, world)
        col_ok = bool((cs.gather(cs.compute(A)) == single).all())
        # device-resident, baby-step rotations shared between the ranks (the layout of the full-size config-4 run)
        dev = torch.device("cuda", rank)
        lo, hi = cs.ranges[rank]
        d_A = torch.from_numpy(A.view(np.int64)).to(dev)
        for share in (True, False):
            d_out = torch.zeros((s, hi - lo, 2, 5, o.N), dtype=torch.int64, device=dev)
            cs.compute_dev(d_A, d_out, s, nbr, 5, share_baby=share)
            col_ok = col_ok and bool((d_out.cpu().numpy().view(np.uint64) == single[:, lo:hi]).all())
        rs = RowSharded(cps, X, rank, world)
        row_ok = bool((rs.compute(A) == single).all())
        gs = GiantSharded(cps, GenoFileStream.from_matrix(cps, X), rank, world)
        giant_ok = bool((gs.compute(A) == single).all())
        gs2 = GiantSharded(cps, GenoFileStream.from_matrix(cps, X), rank, world, shard_baby=False)
        giant_ok = giant_ok and bool((gs2.compute(A) == single).all())
        q.put((rank, col_ok, row_ok and giant_ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpu_shardings_bit_exact():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(r[1] and r[2] for r in res), res
