"""The GPU path against the statistics the REFERENCE itself recorded (sinter + Stim + PyMatching runs kept in its notebooks).

Fixture: tests/golden/notebook_ler_readoffs.json (tests/golden/make_notebook_ler_readoffs.py):
* ``toric_taskstats`` -- the exact ``sinter.TaskStats`` rows of ``/root/reference/notebooks/toric_code.ipynb:181-220``
  (Toric L = 8, 12, 16, 20 through ``CssCode``, 10^6 shots or 10^5 errors per row, decoder "pymatching");
* ``points`` -- logical error rates read off the plots of ``notebooks/rotated_surface_code.ipynb``,
  ``unrotated_surface_code.ipynb`` and ``steane.ipynb`` (10^5 shots / 1000 errors per point, +-3 % read-off error).

These are the only outputs of the reference's own hot path that exist anywhere, so they are what
pins the circuit -> DEM -> matching-graph semantics that cannot be checked against the wheels: Stim's error decomposition
(graphqec_b200/dem.py) and PyMatching's graph conventions.  The full comparison (114 plot points, 20 toric rows, three
candidate decomposition rules) is in profiles/r02_reference_stats_*.log; here a subset sized for the GPU suite.
"""
import json
import math
import os
import sys

import numpy as np
import pytest

import graphqec_b200 as gq
from graphqec_b200 import backend as be

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_reference_circuits import materialise  # noqa: E402

pytestmark = pytest.mark.gpu
_GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "notebook_ler_readoffs.json")))
SEED = 20261019


def run_point(code_name, spec, p, logic_check, shots):
    code = getattr(gq, code_name)(**materialise(spec), depolarize1_rate=p, depolarize2_rate=p)
    code.build_memory_circuit(number_of_rounds=code.distance, logic_check=logic_check)
    try:                                   # sinter's rule
        dem = code.memory_circuit.detector_error_model(decompose_errors=True)
    except ValueError:
        dem = code.memory_circuit.detector_error_model()
    demh = be.DemHandle(dem, device=0)
    dech = be.DecoderHandle(demh, be.MATCHING)
    n = -(-int(shots) // demh.tile_shots) * demh.tile_shots
    got = dech.collect(SEED, 0, n, 0)
    dech.close(); demh.close()
    return got[0], got[1]


def test_toric_error_counts_match_the_reference_taskstats():
    """20 rows at matched shot counts: z-scores of the two binomial counts against each other.  With the round-1
    decomposer (fewest known pieces) these rows sat 1.6 % low (mean z -1.8); with a pairs-first variant 3.9 % high
    (mean z +4.7); with the Stim-style rule: mean z +0.1, rms 1.0."""
    zs = []
    for row in _GOLD["toric_taskstats"]:
        if row["L"] > 12 or (row["L"] == 12 and row["p"] > 0.0105):      # the heavier rows: profiles/r02_reference_stats_stim_style.log
            continue
        n, e = run_point("CssCode", {"toric": row["L"], "name": "Toric"}, row["p"], "Z", row["shots"])
        ref, ours = row["errors"] / row["shots"], e / n
        z = math.log(ours / ref) / math.sqrt(1.0 / row["errors"] + 1.0 / e)
        zs.append(z)
        assert abs(z) < 4.0, (row["L"], row["p"], row["errors"], row["shots"], e, n)
    zs = np.array(zs)
    assert len(zs) == 15
    assert abs(zs.mean()) < 1.0 and math.sqrt((zs ** 2).mean()) < 1.6, zs


def _z(pt, e, n):
    rel = math.sqrt(1.0 / max(pt["approx_errors"], 1.0) + 0.03 ** 2 + 1.0 / max(e, 1))
    return math.log(max(e / n, 1e-12) / pt["ler"]) / rel


def test_steane_matches_the_reference_plot_through_the_undecomposed_model():
    """notebooks/steane.ipynb (Z memory): LER 2.9e-2 at p = 2.2e-3.  Stim cannot decompose this model, sinter falls back
    to the undecomposed one and PyMatching drops its hyperedges; any decomposed model gives 2.0 .. 2.2e-2 (z = -6 .. -8)."""
    pts = [pt for pt in _GOLD["points"] if pt["code"] == "CssCode"]
    assert len(pts) == 9
    zs = []
    for pt in pts:
        n, e = run_point(pt["code"], pt["spec"], pt["p"], pt["logic_check"], max(1 << 18, int(30000 / pt["ler"])))
        zs.append(_z(pt, e, n))
        assert abs(zs[-1]) < 3.5, (pt["p"], pt["ler"], e / n)
    assert abs(np.mean(zs)) < 2.0


def test_rotated_surface_code_matches_the_reference_plots():
    """notebooks/rotated_surface_code.ipynb, Z and X memory, d = 3, 5, 7, 11 (older revision of the reference: names count data
    qubits only).  d = 11: 4.2e-4 at p = 4.4e-3 and 4.9e-3 at p = 6.7e-3 (the read-offs VERDICT r01 quotes)."""
    pts = [pt for pt in _GOLD["points"] if pt["code"] == "RotatedSurfaceCode" and pt["p"] < 0.0125]
    zs, ratios = [], []
    for pt in pts:
        n, e = run_point(pt["code"], pt["spec"], pt["p"], pt["logic_check"], min(1 << 22, max(1 << 18, int(20000 / pt["ler"]))))
        zs.append(_z(pt, e, n))
        ratios.append(e / n / pt["ler"])
        assert abs(zs[-1]) < 3.5, (pt["spec"], pt["logic_check"], pt["p"], pt["ler"], e / n)
    assert len(pts) >= 20
    # all points together: no systematic offset beyond the read-off accuracy (measured: mean ratio 0.96, mean z -1.1)
    assert 0.90 < np.mean(ratios) < 1.10 and abs(np.mean(zs)) < 1.8


@pytest.mark.xfail(strict=False, reason="notebooks/unrotated_surface_code.ipynb was recorded with another revision of the unrotated "
                   "schedule: d = 3 sits 20 % above this circuit's LER at every p (ratio 0.76 .. 0.93, z down to -4.7), d >= 5 agrees "
                   "(mean ratio 0.97); numbers in profiles/r02_reference_stats_stim_style.log")
def test_unrotated_surface_code_against_the_reference_plots():
    pts = [pt for pt in _GOLD["points"] if pt["code"] == "UnrotatedSurfaceCode" and pt["p"] < 0.0095 and pt["spec"]["distance"] <= 7]
    for pt in pts:
        n, e = run_point(pt["code"], pt["spec"], pt["p"], pt["logic_check"], min(1 << 22, max(1 << 18, int(20000 / pt["ler"]))))
        assert abs(_z(pt, e, n)) < 3.5, (pt["spec"], pt["logic_check"], pt["p"], pt["ler"], e / n)


def test_unrotated_d5_d7_agree_with_the_reference_plots():
    pts = [pt for pt in _GOLD["points"] if pt["code"] == "UnrotatedSurfaceCode" and pt["p"] < 0.0095 and pt["spec"]["distance"] in (5, 7)]
    ratios = []
    for pt in pts:
        n, e = run_point(pt["code"], pt["spec"], pt["p"], pt["logic_check"], min(1 << 22, max(1 << 18, int(20000 / pt["ler"]))))
        if pt["approx_errors"] >= 100:
            ratios.append(e / n / pt["ler"])
        assert abs(_z(pt, e, n)) < 4.0, (pt["spec"], pt["logic_check"], pt["p"], pt["ler"], e / n)
    assert 0.85 < np.mean(ratios) < 1.10
