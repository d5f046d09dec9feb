"""Host logic of bench.py that needs no GPU: the workload generator (every rank count holds the SAME scene), the developed state,
the per-kernel roofline bookkeeping, the clock sampler's fall-back."""
import numpy as np
import pytest

import bench as B


def test_rank_pieces_of_the_dam_break_are_the_one_gpu_scene():
    """dam_break_positions draws one random stream per lattice row: the rows of any y-slab decomposition, concatenated, are the
    whole scene bit for bit, and the global id base of a piece is its offset in the whole scene's order."""
    res = 32
    whole, base = B.dam_break_positions(res)
    assert base == 0 and whole.shape == (6 * 6 * 28 * 8, 3)                      # x, z in [2h, 0.25): 6 cells; y in [2h, 1 - 2h): 28 cells; 8 per cell
    cells = np.floor(whole * res).astype(int)
    assert cells.min(axis=0).tolist() == [2, 2, 2] and cells.max(axis=0).tolist() == [7, 29, 7]
    counts = np.bincount((cells[:, 2] * res + cells[:, 1]) * res + cells[:, 0])
    assert set(counts[counts > 0].tolist()) == {8}                               # one particle per sub-cell
    for world in (2, 3, 8):
        lo_c, hi_c = 2, res - 2
        bounds = [0] + [lo_c + round(r * (hi_c - lo_c) / world) for r in range(1, world)] + [res]     # distributed.bench_main's slabs
        pieces = [B.dam_break_positions(res, y_cells=(bounds[r], bounds[r + 1])) for r in range(world)]
        assert np.array_equal(np.concatenate([p[0] for p in pieces]), whole)
        off = 0
        for (x, b), r in zip(pieces, range(world)):
            assert b == off and (np.floor(x[:, 1] * res) >= bounds[r]).all() and (np.floor(x[:, 1] * res) < bounds[r + 1]).all()
            off += len(x)


def test_developed_state_is_a_closed_form_of_the_position():
    x, _ = B.dam_break_positions(16)
    arrs, keep = B.packed_rest_state(x, 1.0)

    class A:
        state = "flowing"; perturb = 0.0
    B.set_state(arrs, A)
    v = arrs[1]
    assert np.allclose(v[0], B.FLOW_U * x[:, 2] / B.FLOW_H) and np.allclose(v[1], B.FLOW_V * np.sin(2 * np.pi * x[:, 2] / B.FLOW_H)) and not v[2].any()
    A.state = "rest"; B.set_state(arrs, A)
    assert not arrs[1].any() and np.array_equal(arrs[5][:, [0, 4, 8]], np.ones((len(x), 3)))       # F_E = I


def test_dominant_kernel_roofline_bookkeeping():
    n, nodes, peak = 64520064, 8.5e6, 6550.4
    stage = {"forces": 3.0, "forces_list": 0.3, "force_scatter": 3.6, "g2p": 5.9, "g2p_list": 0.7, "p2g": 3.9, "grid": 0.13, "sort": 6.0}
    dom, r = B.kernel_roofline(stage, n, nodes, peak, "measured")
    assert dom == "g2p" and r["kernel"].startswith("k_g2p2g<SCATTER=0>")                                  # the re-sort is overhead, never the roofline kernel
    assert r["algorithmic_bytes_per_launch"] == 224.0 * n + 24.0 * nodes and r["launch_ms"] == 5.9
    assert r["achieved"] == pytest.approx(r["algorithmic_bytes_per_launch"] / 5.9e-3 / 1e9) and r["frac"] == pytest.approx(r["achieved"] / peak)
    f = r["force_stage"]
    assert f["ms"] == pytest.approx(3.0 + 0.3 + 3.6) and f["algorithmic_bytes"] == 52.0 * n + 24.0 * nodes
    t = B.transfers_roofline(stage, n, nodes, peak)
    assert t["ms"] == pytest.approx(3.9 + 5.9 + 0.7) and t["algorithmic_bytes"] == (224.0 + 64.0) * n + (24.0 + 44.0) * nodes


def test_clock_sampler_reports_nothing_without_a_gpu_instead_of_failing():
    c = B.ClockSampler(0)
    out = c.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons"}
