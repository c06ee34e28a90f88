"""`from graphqec import ...` -- the reference's import name (README.md:62,75 there) -- resolves to this implementation."""
import importlib

import numpy as np


def test_reference_import_lines_work_unchanged():
    from graphqec import RepetitionCode, ThresholdLAB              # reference README.md:75
    from graphqec import CssCode, BivariateBicycleCode, RotatedSurfaceCode, UnrotatedSurfaceCode   # its notebooks
    import graphqec
    import graphqec_b200

    assert ThresholdLAB is graphqec_b200.ThresholdLAB and RepetitionCode is graphqec_b200.RepetitionCode
    # the reference's package exports: codes, Measurement, X_check / Z_check, ThresholdLAB, the GF(2) helpers
    for name in ("BaseCode", "Measurement", "X_check", "Z_check", "commutation_test", "compute_kernel", "find_pivots"):
        assert getattr(graphqec, name) is getattr(graphqec_b200, name)
    rep = RepetitionCode(distance=5, depolarize1_rate=0.05, depolarize2_rate=0.05)
    rep.build_memory_circuit(number_of_rounds=2, logic_check="Z")   # reference README.md:62-71
    assert rep.name == "Repetition [[9,1,5]]" and rep.memory_circuit.num_detectors == 12
    th = ThresholdLAB(configurations=[{"distance": d} for d in (3, 5)], code=RepetitionCode,
                      error_rates=np.linspace(0, 0.2, 3), decoder="pymatching")
    assert len(list(th.generate_sinter_tasks())) == 6
    assert all(c is not None for c in (CssCode, BivariateBicycleCode, RotatedSurfaceCode, UnrotatedSurfaceCode))


def test_reference_submodule_paths_are_the_same_module_objects():
    for path in ("codes", "codes.base_code", "codes.css_code", "codes.rotated_surface_code", "lab", "lab.threshold",
                 "lab.threshold.threshold_lab", "math", "measurement", "stab"):
        assert importlib.import_module("graphqec." + path) is importlib.import_module("graphqec_b200." + path)
    from graphqec.lab.threshold.threshold_lab import ThresholdLAB
    from graphqec.codes.base_code import BaseCode
    from graphqec.stab import X_check
    import graphqec_b200

    assert ThresholdLAB is graphqec_b200.ThresholdLAB and BaseCode is graphqec_b200.BaseCode and X_check is graphqec_b200.X_check
