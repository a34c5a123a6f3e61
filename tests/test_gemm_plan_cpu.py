"""CPU: the host-side launch plan of the five front-end contractions (st_debug_gemm_plan -- no device work).  The tile width /
split-K heuristic of the cta_group::2 GEMM must reproduce, for BASELINE configs[1] (B = 200), exactly the grids the ncu capture of
the real step recorded (profiles/r01_v8_gemm_tails_ncu_full_summary.csv: launch__grid_size 148, 144, 108, 120, 148), and behave at
the edges: one M-tile -> 1-CTA kernel, every plan covers N exactly, split-K never exceeds what the finalisation kernel sums."""
import csv
import ctypes
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["analysis", "synthesis", "synthesis_dgrad", "synthesis_wgrad", "analysis_wgrad"]


def _plan(scale, knobs, batch, sms=148):
    from signaltrain_b200 import _lib
    from signaltrain_b200.engine import Geometry
    lib = _lib.load()
    g = Geometry(scale, 4, knobs)
    cfg = _lib.StConfig(g.C, g.N, g.H, g.T, g.OT, g.K, g.R)
    out = (ctypes.c_int * 20)()
    assert lib.st_debug_gemm_plan(ctypes.byref(cfg), batch, sms, out) == 0
    return {n: dict(pair=out[4 * i], bn=out[4 * i + 1], splits=out[4 * i + 2], grid=out[4 * i + 3]) for i, n in enumerate(NAMES)}, g


def test_plan_at_bench_size_matches_the_ncu_capture():
    plan, _ = _plan(1, 4, 200)
    assert all(p["pair"] == 1 for p in plan.values())
    assert [plan[n]["grid"] for n in NAMES] == [148, 144, 108, 120, 148]
    assert [plan[n]["bn"] for n in NAMES] == [176, 128, 176, 256, 256]
    assert [plan[n]["splits"] for n in NAMES] == [1, 1, 1, 3, 7]
    # the committed ncu summary holds the same step: its five gemm_tc2_kernel launches, in launch order
    rows = list(csv.reader(open(os.path.join(ROOT, "profiles", "r01_v8_gemm_tails_ncu_full_summary.csv"))))
    cols = [i for i, name in enumerate(rows[0]) if "gemm_tc2_kernel" in name]
    grid_row = next(r for r in rows if r[0] == "launch__grid_size")
    ncu_grids = [int(float(grid_row[i])) for i in cols]          # analysis, synthesis, dgrad, synthesis wgrad, analysis wgrad
    assert ncu_grids == [plan[n]["grid"] for n in NAMES]


@pytest.mark.parametrize("scale,knobs,batch", [(1, 4, 1), (1, 4, 3), (1, 4, 12), (1, 4, 37), (1, 4, 512), (2, 2, 256), (1, 1, 256)])
def test_plan_invariants(scale, knobs, batch):
    plan, g = _plan(scale, knobs, batch)
    f2 = {"analysis": None, "synthesis": g.N, "synthesis_dgrad": None, "synthesis_wgrad": g.N, "analysis_wgrad": g.N}
    for name, p in plan.items():
        n = f2[name]
        if n is not None:
            assert n % p["bn"] == 0, name                       # tiles cover N exactly
        assert 64 <= p["bn"] <= 256 and p["bn"] % 16 == 0
        assert 1 <= p["splits"] <= 8                             # finalize kernels sum at most kMaxSplits planes
        assert 1 <= p["grid"] <= 148 and (p["pair"] == 0 or p["grid"] % 2 == 0)
        if name in ("synthesis", "synthesis_wgrad", "analysis_wgrad") and p["pair"]:
            assert p["bn"] % 64 == 0, name                       # MN-major B: each CTA's half is whole 32-column slabs
    frame_rows = batch * ((g.C + 2 * g.N + g.H - 1) // g.H)
    assert plan["analysis"]["pair"] == (1 if frame_rows > 128 else 0)      # a single M-tile cannot be paired
    assert plan["analysis"]["splits"] == plan["synthesis"]["splits"] == plan["synthesis_dgrad"]["splits"] == 1
