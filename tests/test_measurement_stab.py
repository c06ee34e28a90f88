"""Mirrors /root/reference/tests/test_measurement.py and tests/test_stab.py."""
from graphqec_b200 import Circuit, Measurement, X_check, Z_check


def test_measurement_init():
    m = Measurement()
    assert m.data == {}
    assert m.register_count == 0


def test_measurement_add_get():
    m = Measurement()
    m.add_outcome(outcome="0", qubit=0, round=1, type="check")
    assert m.get_outcome(qubit=0, round=1) == "0"
    assert m.get_outcome(qubit=1, round=1) is None
    assert m.register_count == 1
    # an unknown type is accepted silently, as in the reference (test_measurement.py:34)
    m.add_outcome(outcome="0", qubit=0, round=2, type="cheeeeeeck")
    m.add_outcome(outcome="0", qubit=1, round=1, type="check")
    assert m.register_count == 3


def test_measurement_register_id():
    m = Measurement()
    m.add_outcome(outcome="0", qubit=0, round=1, type="check")
    assert m.get_register_id(qubit=0, round=1) == 0
    assert m.get_register_id(qubit=1, round=1) is None


def test_x_check_direction():
    circ = Circuit()
    X_check(circ=circ, data_qubit=0, check_qubit=1)
    assert str(circ) == "CX 1 0"  # reference tests/test_stab.py:25


def test_z_check_direction():
    circ = Circuit()
    Z_check(circ=circ, data_qubit=0, check_qubit=1)
    assert str(circ) == "CX 0 1"  # reference tests/test_stab.py:30


def test_circuit_text_roundtrip_and_fusion():
    c = Circuit()
    c.append("R", [0, 1, 2])
    c.append("R", [3, 4])
    c.append("DEPOLARIZE1", [0, 1], 0.05)
    c.append("DEPOLARIZE1", [], 0.05)
    c.append("MR", [3, 4])
    c.append_from_stim_program_text("DETECTOR rec[-2]\nOBSERVABLE_INCLUDE(0) rec[-1]")
    assert str(c).splitlines()[0] == "R 0 1 2 3 4"
    assert c.num_measurements == 2 and c.num_detectors == 1 and c.num_observables == 1
    assert Circuit(str(c)) == c
