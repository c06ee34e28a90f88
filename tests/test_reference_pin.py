"""The product's circuit builders against the REFERENCE's own builders, instruction by instruction.

``tests/golden/reference_circuits.json`` was produced by running ``/root/reference/graphqec`` itself
(``base_code.py:184-425``, the five ``build_graph``s, ``threshold_lab.py:118-151``) under a recording
``stim`` shim -- ``tests/golden/make_reference_circuits.py``.  Every BASELINE.json configuration is in
it: Repetition d in {3,5,7,11} with r = 2 and r = d, Rotated d in {3,...,11} (+ even d = 4, 20 and the
example's d = 23), Unrotated d = 15, Toric L = 8 / 12 / 15 and Steane through ``CssCode``, BB-72 / BB-144,
Z and X memory.  The circuit is the whole input of the hot path, so identical text == identical DEM.
"""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

import graphqec_b200 as gq
from graphqec_b200 import ThresholdLAB

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_reference_circuits import materialise  # noqa: E402  (spec -> constructor kwargs)

_GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_circuits.json")))


def _logicals(code):
    out = {}
    for k, v in code.logic_check.items():
        v = list(v)
        out[k] = [[int(x) for x in row] for row in v] if v and hasattr(v[0], "__iter__") else [int(x) for x in v]
    return out


@pytest.mark.parametrize("ent", _GOLD["circuits"], ids=[e["id"] for e in _GOLD["circuits"]])
def test_memory_circuit_identical_to_the_reference_builder(ent):
    kw = materialise(ent["spec"])
    code = getattr(gq, ent["code"])(**kw, depolarize1_rate=ent["p"], depolarize2_rate=ent["p"])
    assert code.name == ent["name"]
    assert int(code.distance) == ent["distance"]
    assert int(code.num_physical_qubits) == ent["num_physical_qubits"]
    assert int(code.num_logical_qubits) == ent["num_logical_qubits"]
    assert _logicals(code) == ent["logic_check_ops"]
    rounds = code.distance if ent["rounds_arg"] is None else ent["rounds_arg"]
    assert rounds == ent["rounds"]
    code.build_memory_circuit(number_of_rounds=rounds, logic_check=ent["logic_check"])
    circ = code.memory_circuit
    text = str(circ)
    if "text" in ent:
        assert text == ent["text"]
    assert circ.num_measurements == ent["num_measurements"]
    assert circ.num_detectors == ent["num_detectors"]
    assert circ.num_observables == ent["num_observables"]
    assert circ.count("CX") == ent["num_cx"]
    assert circ.count("DEPOLARIZE1") == ent["num_depolarize1"]
    assert circ.count("DEPOLARIZE2") == ent["num_depolarize2"]
    assert text.count("\n") + 1 == ent["num_lines"]
    assert hashlib.sha256(text.encode()).hexdigest() == ent["sha256"]


@pytest.mark.parametrize("lab", _GOLD["tasks"], ids=[t["id"] for t in _GOLD["tasks"]])
def test_generate_sinter_tasks_identical_to_the_reference(lab):
    """Task order, metadata (name, error, n, k, d, r) and circuits of threshold_lab.py:118-151."""
    th = ThresholdLAB(code=getattr(gq, lab["code"]), configurations=[materialise(s) for s in lab["specs"]],
                      error_rates=lab["error_rates"], decoder="pymatching")
    tasks = list(th.generate_sinter_tasks(logic_check=lab["logic_check"]))
    assert len(tasks) == len(lab["rows"])
    for task, row in zip(tasks, lab["rows"]):
        md = {k: (v.item() if isinstance(v, np.generic) else v) for k, v in task.json_metadata.items()}
        assert md == row["json_metadata"]
        text = str(task.circuit)
        assert text.count("\n") + 1 == row["num_lines"]
        assert hashlib.sha256(text.encode()).hexdigest() == row["sha256"]


def test_fixture_covers_every_baseline_config():
    ids = {e["id"] for e in _GOLD["circuits"]}
    need = {"rep3_r2", "rep5_r2", "rep7_r2", "rep11_r2", "rep3_rd", "rep11_rd", "rot5_Z", "rot7_Z", "rot9_Z", "rot11_Z",
            "unrot15_Z", "toric15_Z", "toric8_Z", "steane_Z", "bb144_Z", "rot11_X", "rot20_Z"}
    assert need <= ids
    by = {e["id"]: e for e in _GOLD["circuits"]}
    # SURVEY 8a sizes, now read from the reference's own output
    assert by["rot11_Z"]["num_detectors"] == 1320 and by["rot11_Z"]["num_cx"] == 4840
    assert by["toric15_Z"]["num_detectors"] == 6750 and by["toric15_Z"]["num_cx"] == 27000
    assert by["unrot15_Z"]["num_detectors"] == 6300 and by["bb144_Z"]["num_detectors"] == 1728
