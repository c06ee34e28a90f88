"""The reference's example workload (examples/rot_vs_unrot_surface_code.py there) through this package."""
import importlib.util
import json
import os

import numpy as np
import pytest

from graphqec_b200.stats import TaskStats

_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "rot_vs_unrot_surface_code.py")


def _load():
    spec = importlib.util.spec_from_file_location("rot_vs_unrot_example", _PATH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_fit_and_per_round_conversion():
    ex = _load()
    # sinter.shot_error_rate_to_piece_error_rate: r rounds of rate q compose to (1 - (1 - 2q)^r) / 2
    for q, r in ((1e-3, 11), (0.02, 5), (0.3, 3)):
        shot = 0.5 * (1 - (1 - 2 * q) ** r)
        assert abs(ex.shot_error_rate_to_piece_error_rate(shot, r) - q) < 1e-12
    assert abs(ex.shot_error_rate_to_piece_error_rate(0.9, 5) + ex.shot_error_rate_to_piece_error_rate(0.1, 5) - 1) < 1e-12
    stats = [TaskStats(str(d), "pymatching", {"n": n, "r": d, "error": 0.005}, 10**6, e, 0, 1.0)
             for d, n, e in ((5, 49, 6000), (7, 97, 3000), (9, 161, 1500), (11, 241, 0))]
    stats.append(TaskStats("other", "pymatching", {"n": 49, "r": 5, "error": 0.006}, 10**6, 9000, 0, 1.0))
    f, xs, ys = ex.fit(stats, 0.005)
    assert len(xs) == 3 and xs[0] == 7.0                    # the point without errors and the other error rate are left out
    slope, intercept = np.polyfit(xs, np.log(ys), 1)
    assert abs(f.slope - slope) < 1e-12 and abs(f.intercept - intercept) < 1e-12 and f.slope < 0


@pytest.mark.gpu
def test_example_runs_on_the_device(tmp_path):
    """a small corner of the reference's grid (its own runs d = 9 .. 23 with 10^6 shots per point)"""
    ex = _load()
    errors = np.array([0.004, 0.006])
    rows = ex.main(configs=[{"distance": d} for d in (3, 5, 7)], errors=errors, max_shots=400_000, max_errors=5000,
                   logic_check="Z", out_prefix=str(tmp_path / "out"))
    assert len(rows) == 4 and {r["type"] for r in rows} == {"Rotated", "Unrotated"}
    saved = json.load(open(tmp_path / "out.json"))
    assert saved == json.loads(json.dumps(rows))
    for r in rows:
        assert r["n"] == ([17, 49, 97] if r["type"] == "Rotated" else [25, 81, 169])
        assert all(0 < y < 0.1 for y in r["logical_error"])
        if r["error_rate"] == 0.004:
            assert r["fit"]["slope"] < 0        # below threshold the per-round rate falls with the code size
