"""Host-side logic of bench.py (no GPU): the algorithmic work figures of SURVEY 8(d) that every reported number is divided by, the
workload table, the committed ncu traffic figure the roofline line quotes, and the SNP-block split of the full-size config-4 run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import bench  # noqa: E402


def test_work_figures_config2_match_design():
    w = bench.WORKLOADS["mm_10k_x_100k_k10_logN13"]
    wf = bench.work_figures(w["nrows"], w["ncols"], w["s"], bench.CKKS[w["params"]]["logN"])
    assert (wf["nbr"], wf["m_ct"], wf["d"], wf["slots"]) == (3, 25, 64, 4096)
    assert wf["diag_polys"] == 306607
    assert wf["b_diag"] == 306607 * 5 * 8192 * 8
    assert round(wf["b_alg"] / 1e9, 2) == 100.66
    assert (wf["ks_baby"], wf["ks_giant"]) == (1890, 15750)
    assert wf["mac_alg"] == 306607 * 10 * 2 * 5 * 8192


def test_work_figures_config4_full_size():
    w = bench.WORKLOADS["pca_100k_x_500k_k15_logN14_otf"]
    assert w["otf"] and w["colshard"] and w["s"] == 15
    wf = bench.work_figures(w["nrows"], w["ncols"], w["s"], bench.CKKS[w["params"]]["logN"])
    assert (wf["nbr"], wf["m_ct"], wf["d"], wf["slots"]) == (13, 62, 91, 8192)
    # every block but the ragged corner (1 696 rows x 288 columns) has all 8 192 generalized diagonals
    assert wf["diag_polys"] == (13 * 62 - 1) * 8192 + (1696 + 288 - 1)
    assert wf["ks_baby"] == 13 * 90 * 15 and wf["ks_giant"] == 15 * 90 * 62


def test_transposed_workload_is_the_transpose():
    a, b = bench.WORKLOADS["mm_10k_x_100k_k10_logN13"], bench.WORKLOADS["mm_100k_x_10k_k10_logN13_T"]
    assert (a["nrows"], a["ncols"]) == (b["ncols"], b["nrows"]) and a["params"] == b["params"] and a["s"] == b["s"]
    wf = bench.work_figures(b["nrows"], b["ncols"], b["s"], 13)
    assert wf["nbr"] == 25 and wf["nbr"] * wf["d"] == 1600  # 7 K groups of <= 256 baby-step slots
    assert wf["ks_baby"] == 15750


def test_snp_block_split_gives_every_rank_a_full_width_column():
    """bench.py --col-sharding shares the baby steps between the ranks: every rank's own cache must then see all of them, i.e. own at
    least one block column of full width (the library-side check is the all-reduce of the chunk size)."""
    from sfgwas_b200.dist import partition

    w = bench.WORKLOADS["pca_100k_x_500k_k15_logN14_otf"]
    slots, m_ct = 8192, 62
    for world in (2, 4, 8):
        parts = partition(m_ct, world)
        assert parts[0][0] == 0 and parts[-1][1] == m_ct
        for lo, hi in parts:
            widths = [min((bj + 1) * slots, w["ncols"]) - bj * slots for bj in range(lo, hi)]
            assert widths and max(widths) == slots


def test_roofline_traffic_comes_from_the_committed_capture():
    traffic, src = bench.read_traffic()
    assert traffic and 5e10 < traffic < 9e10, (traffic, src)  # dram read + write of one k_mac_tc launch of config 2: ~66.5 GB
    assert "profiles/r2" in src


def test_run_ours_teardown_closes_the_context_last(monkeypatch):
    """The wrapper around the measurement owns the teardown order (tensors, torch's current stream, process group, THEN the library
    context).  Without a GPU the measurement is replaced by a stub: the context must be closed exactly once, after the stub returned or
    raised, and an exception of the measurement must propagate."""
    import pytest

    events = []

    class FakeCtx:
        def close(self):
            events.append("close")

    def fake_run(args, rank, local_rank, world, keep):
        keep["cps"] = FakeCtx()
        events.append("run")

    monkeypatch.setattr(bench, "_run_ours", fake_run)
    bench.run_ours(None, 0, 0, 1)
    assert events == ["run", "close"]

    def failing_run(args, rank, local_rank, world, keep):
        keep["cps"] = FakeCtx()
        raise RuntimeError("measurement failed")

    events.clear()
    monkeypatch.setattr(bench, "_run_ours", failing_run)
    with pytest.raises(RuntimeError):
        bench.run_ours(None, 0, 0, 1)
    assert events == ["close"]
