"""Circuit builder parity: names, sizes, notebook known-answer vectors, schedule quirks."""
import json
import os

import numpy as np
import pytest

import graphqec_b200 as gq
from graphqec_b200 import (BivariateBicycleCode, Circuit, CssCode, RepetitionCode,
                           RotatedSurfaceCode, UnrotatedSurfaceCode)
from graphqec_b200.math import commutation_test


def _detector_sets(circ):
    """absolute measurement-record sets of every detector / observable 0"""
    seen, dets, obs = 0, [], set()
    for op in circ:
        if op.kind in ("measure", "measure_reset"):
            seen += len(op.targets)
        elif op.name == "DETECTOR":
            dets.append(sorted(seen + t.value for t in op.targets))
        elif op.name == "OBSERVABLE_INCLUDE":
            obs ^= {seen + t.value for t in op.targets}
    return dets, sorted(obs)


def test_names():
    # reference tests/codes/test_repetition_code.py:33
    assert RepetitionCode(distance=5).name == "Repetition [[9,1,5]]"
    assert RotatedSurfaceCode(distance=11).name == "Rotated Surface [[241,1,11]]"
    assert UnrotatedSurfaceCode(distance=3).name == "Unrotated Surface [[25,1,3]]"


def test_memory_circuit_type_and_target_rec():
    # reference tests/codes/test_rotated_surface_code.py:29-39
    code = RotatedSurfaceCode(distance=3, depolarize1_rate=0.01, depolarize2_rate=0)
    code.build_memory_circuit(number_of_rounds=2)
    assert type(code.memory_circuit) is Circuit
    assert isinstance(code.measurement, gq.Measurement)
    assert isinstance(code.register_count, int)
    fresh = RotatedSurfaceCode(distance=3, depolarize1_rate=0.01, depolarize2_rate=0)
    assert fresh.get_outcome(qubit=0, round=0) is None
    fresh.add_outcome(outcome="0", qubit=0, round=0, type="check")
    assert isinstance(fresh.get_target_rec(qubit=0, round=0), int)
    assert fresh.get_target_rec(qubit=1000, round=0) is None


def test_bad_logic_check_raises():
    with pytest.raises(ValueError):
        RepetitionCode(distance=3).build_memory_circuit(number_of_rounds=2, logic_check="X")


def test_notebook_known_answer_circuits(golden_dir):
    """Detector and observable record sets recorded in the reference's notebooks."""
    kats = json.load(open(os.path.join(golden_dir, "notebook_circuits.json")))
    assert len(kats) == 4
    for kat in kats:
        code = getattr(gq, kat["code"])(**kat["kwargs"])
        code.build_memory_circuit(number_of_rounds=kat["rounds"], logic_check=kat["logic_check"])
        dets, obs = _detector_sets(code.memory_circuit)
        assert code.memory_circuit.num_measurements == kat["num_measurements"], kat["source"]
        assert dets == kat["detectors"], kat["source"]
        assert obs == kat["observable0"], kat["source"]


def test_rebuilding_on_the_same_object_keeps_offsets_right(golden_dir):
    # the rotated notebook builds r=3 then r=2 on one object (cells 3, 4)
    kats = json.load(open(os.path.join(golden_dir, "notebook_circuits.json")))
    code = RotatedSurfaceCode(distance=3, depolarize1_rate=0.01, depolarize2_rate=0.01)
    for kat in kats[1:3]:
        code.build_memory_circuit(number_of_rounds=kat["rounds"], logic_check="X")
        dets, obs = _detector_sets(code.memory_circuit)
        assert dets == kat["detectors"] and obs == kat["observable0"]


def test_appendix_c_instruction_stream():
    """repetition d=3, 2 rounds -- exact instruction stream (fused like Stim)."""
    rep = RepetitionCode(distance=3, depolarize1_rate=0.05, depolarize2_rate=0.05)
    rep.build_memory_circuit(number_of_rounds=2, logic_check="Z")
    round_body = [
        "CX 0 3", "DEPOLARIZE2(0.05) 0 3", "CX 1 4", "DEPOLARIZE2(0.05) 1 4",
        "CX 1 3", "DEPOLARIZE2(0.05) 1 3", "CX 2 4", "DEPOLARIZE2(0.05) 2 4",
        "DEPOLARIZE1(0.05) 3 4", "MR 3 4",
    ]
    want = (["R 0 1 2 3 4", "DEPOLARIZE1(0.05) 0 1 2 3 4"] + round_body
            + ["DETECTOR rec[-2]", "DETECTOR rec[-1]", "DEPOLARIZE1(0.05) 3 4"] + round_body
            + ["DETECTOR rec[-4] rec[-2]", "DETECTOR rec[-3] rec[-1]", "DEPOLARIZE1(0.05) 0 1 2", "M 0 1 2",
               "DETECTOR rec[-3] rec[-2] rec[-5]", "DETECTOR rec[-2] rec[-1] rec[-4]", "OBSERVABLE_INCLUDE(0) rec[-3]"])
    assert str(rep.memory_circuit).splitlines() == want


@pytest.mark.parametrize("d,nq,ops_cx,D", [(5, 49, 400, 120), (7, 97, 1176, 336), (11, 241, 4840, 1320)])
def test_rotated_sizes(d, nq, ops_cx, D):
    """SURVEY 8a size table (PROBE values measured on the reference builder)."""
    code = RotatedSurfaceCode(distance=d, depolarize1_rate=0.005, depolarize2_rate=0.005)
    code.build_memory_circuit(number_of_rounds=d, logic_check="Z")
    c = code.memory_circuit
    assert code.num_physical_qubits == nq
    assert c.count("CX") == ops_cx and c.count("DEPOLARIZE2") == ops_cx
    assert c.num_detectors == D and c.num_observables == 1
    if d == 11:
        assert c.num_measurements == 1441 and c.count("DEPOLARIZE1") == 4202


def test_detector_layout_formula():
    """D = 2 n_B + (r-1) C; X checks come first in every MR (Q4)."""
    code = RotatedSurfaceCode(distance=5, depolarize1_rate=0.001, depolarize2_rate=0.001)
    code.build_memory_circuit(number_of_rounds=3, logic_check="Z")
    data, groups, anc = code._qubit_groups()
    assert list(groups) == ["X-check", "Z-check"]
    assert anc[: len(groups["X-check"])] == groups["X-check"]
    nz, C = len(groups["Z-check"]), len(anc)
    assert code.memory_circuit.num_detectors == 2 * nz + 2 * C


def test_even_distance_is_accepted():
    # even d is not rejected and yields an undetectable logical mechanism
    code = RotatedSurfaceCode(distance=4, depolarize1_rate=0.005, depolarize2_rate=0.005)
    code.build_memory_circuit(number_of_rounds=4, logic_check="Z")
    dem = code.memory_circuit.detector_error_model()
    sizes = np.diff(dem.det_indptr.astype(int))
    assert (sizes == 0).sum() == 1


def test_css_steane():
    H = np.array([[1, 1, 1, 1, 0, 0, 0], [0, 1, 1, 0, 1, 1, 0], [0, 0, 1, 1, 0, 1, 1]])
    st = CssCode(Hx=H, Hz=H)
    assert st.name == "CSS [[13,1,3]]"
    assert st.num_logical_qubits == 1 and st.distance == 3
    # current reference algorithm gives L_z = [[1,1,0,0,1,0,0]]
    assert st.L_z.tolist() == [[1, 1, 0, 0, 1, 0, 0]]
    assert not (H @ st.L_z.T % 2).any() and not (H @ st.L_x.T % 2).any()
    assert st._length_sym_measurement == 16  # 16 CNOT steps per round
    st.build_memory_circuit(number_of_rounds=3)
    assert st.memory_circuit.num_detectors == 18 and st.memory_circuit.count("CX") == 72


def test_css_rejects_non_commuting():
    with pytest.raises(ValueError):
        CssCode(Hx=np.array([[1, 0]]), Hz=np.array([[1, 0]]))


def _toric(L):
    n = 2 * L * L
    Hx = np.zeros((L * L, n), dtype=int)
    Hz = np.zeros((L * L, n), dtype=int)
    h = lambda x, y: (x % L) * L + (y % L)          # horizontal edge
    v = lambda x, y: L * L + (x % L) * L + (y % L)  # vertical edge
    for x in range(L):
        for y in range(L):
            Hx[x * L + y, [h(x, y), h(x, y - 1), v(x, y), v(x - 1, y)]] = 1
            Hz[x * L + y, [h(x, y), h(x + 1, y), v(x, y), v(x, y + 1)]] = 1
    return Hx, Hz


def test_css_toric_parameters():
    Hx, Hz = _toric(4)
    assert commutation_test(Hx, Hz)
    t = CssCode(Hx=Hx, Hz=Hz, name="Toric")
    assert t.num_logical_qubits == 2
    assert t.name.startswith("Toric [[64,2,")
    t.build_memory_circuit(number_of_rounds=2)
    # all k logicals are XORed into observable 0
    assert t.memory_circuit.num_observables == 1


def test_bivariate_bicycle_small_and_144():
    bb = BivariateBicycleCode(L1=6, L2=6, a1=3, a2=1, a3=2, b1=3, b2=1, b3=2)
    assert bb.num_data_qubits == 72 and bb.num_logical_qubits == 12   # [[72,12,6]]
    assert bb._length_sym_measurement == 7
    assert commutation_test(bb.Hx, bb.Hz)
    degs = {bb.graph.degree(n) for n, a in bb.graph.nodes(data=True) if a["type"] == "check"}
    assert degs == {6}
    big = BivariateBicycleCode()  # defaults: the [[144,12,12]] gross code
    assert big.num_data_qubits == 144 and big.num_logical_qubits == 12
    assert big.name.startswith("Bivariate Bicycle [[144,12,")
    assert big.num_physical_qubits == 288
