"""Circuit -> DEM analysis: known answers, independent forward-propagation oracle, decomposition."""
import numpy as np
import pytest

from graphqec_b200 import (CssCode, DetectorErrorModel, RepetitionCode, RotatedSurfaceCode,
                           UnrotatedSurfaceCode)
from graphqec_b200.dem import NO_DET, depolarize1_component, depolarize2_component
from oracle.dem_forward import forward_fault_table

APPENDIX_C = """
error(0.045948281621) D0 L0      error(0.045948281621) D0 D1      error(0.087674074074) D0 D2
error(0.026666666667) D0 D3      error(0.045948281621) D1         error(0.087674074074) D1 D3
error(0.026666666667) D2 L0      error(0.026666666667) D2 D3      error(0.087674074074) D2 D4
error(0.026666666667) D2 D5      error(0.026666666667) D3         error(0.087674074074) D3 D5
error(0.045948281621) D4 L0      error(0.045948281621) D4 D5      error(0.045948281621) D5
"""


def _as_table(dem):
    return {(dem.mechanism(i)[1], dem.mechanism(i)[2]): dem.mechanism(i)[0] for i in range(dem.num_errors)}


def test_component_probabilities():
    assert abs(depolarize1_component(0.05) - 0.016954108460) < 1e-11
    assert abs(depolarize2_component(0.05) - 0.003413807381) < 1e-11


def test_appendix_c_dem():
    """the smallest known-answer DEM (repetition d=3, 2 rounds, p=0.05)."""
    rep = RepetitionCode(distance=3, depolarize1_rate=0.05, depolarize2_rate=0.05)
    rep.build_memory_circuit(number_of_rounds=2, logic_check="Z")
    dem = rep.memory_circuit.detector_error_model()
    want = DetectorErrorModel.from_text(APPENDIX_C.replace("      ", "\n"))
    got, ref = _as_table(dem), _as_table(want)
    assert set(got) == set(ref) and len(got) == 15
    assert max(abs(got[k] - ref[k]) for k in got) < 1e-11
    assert dem.num_detectors == 6 and dem.num_observables == 1


def test_north_star_dem_sizes():
    """D=1320 O=1 M=24483 nnz(H)=79956 nnz(L)=835 sum p=36.94 min/max p."""
    code = RotatedSurfaceCode(distance=11, depolarize1_rate=0.005, depolarize2_rate=0.005)
    code.build_memory_circuit(number_of_rounds=11, logic_check="Z")
    s = code.memory_circuit.detector_error_model().summary()
    assert (s["D"], s["O"], s["M"], s["nnz_H"], s["nnz_L"]) == (1320, 1, 24483, 79956, 835)
    assert abs(s["sum_p"] - 36.94) < 0.005
    assert abs(s["min_p"] - 3.341e-4) < 1e-7 and abs(s["max_p"] - 2.2227e-2) < 1e-6
    assert s["dets_per_mechanism"] == {1: 360, 2: 6570, 3: 3756, 4: 13797}


@pytest.mark.parametrize("code,rounds,lc", [
    (RepetitionCode(distance=5, depolarize1_rate=0.1, depolarize2_rate=0.1), 5, "Z"),
    (RotatedSurfaceCode(distance=3, depolarize1_rate=0.01, depolarize2_rate=0.02), 3, "Z"),
    (RotatedSurfaceCode(distance=5, depolarize1_rate=0.005, depolarize2_rate=0.005), 4, "X"),
    (UnrotatedSurfaceCode(distance=3, depolarize1_rate=0.003, depolarize2_rate=0.007), 3, "Z"),
    (CssCode(Hx=np.array([[1, 1, 1, 1, 0, 0, 0], [0, 1, 1, 0, 1, 1, 0], [0, 0, 1, 1, 0, 1, 1]]),
             Hz=np.array([[1, 1, 1, 1, 0, 0, 0], [0, 1, 1, 0, 1, 1, 0], [0, 0, 1, 1, 0, 1, 1]]),
             depolarize1_rate=0.002, depolarize2_rate=0.002), 3, "Z"),
])
def test_backward_analysis_equals_forward_fault_propagation(code, rounds, lc):
    code.build_memory_circuit(number_of_rounds=rounds, logic_check=lc)
    got = _as_table(code.memory_circuit.detector_error_model())
    ref = forward_fault_table(code.memory_circuit)
    assert set(got) == set(ref)
    assert max(abs(got[k] - ref[k]) for k in got) < 1e-14


def test_steane_dem_sizes():
    H = np.array([[1, 1, 1, 1, 0, 0, 0], [0, 1, 1, 0, 1, 1, 0], [0, 0, 1, 1, 0, 1, 1]])
    st = CssCode(Hx=H, Hz=H, depolarize1_rate=0.005, depolarize2_rate=0.005)
    st.build_memory_circuit(number_of_rounds=st.distance)
    s = st.memory_circuit.detector_error_model().summary()
    # Steane D=18 M=276 nnz=851 nnz_L=74
    assert (s["D"], s["M"], s["nnz_H"], s["nnz_L"]) == (18, 276, 851, 74)
    assert s["dets_per_mechanism"] == {1: 24, 2: 72, 3: 84, 4: 59, 5: 27, 6: 10}


def test_decomposition_surface_code_splits_cleanly():
    """Stim-style decomposition (dem.py): every piece is graphlike, the pieces XOR to the symptom, every distinct
    decomposition is its own error, and the symptom-merged probabilities equal the undecomposed model's."""
    code = RotatedSurfaceCode(distance=5, depolarize1_rate=0.005, depolarize2_rate=0.005)
    code.build_memory_circuit(number_of_rounds=5, logic_check="Z")
    dem = code.memory_circuit.detector_error_model(decompose_errors=True)
    plain = _as_table(code.memory_circuit.detector_error_model())
    merged, seen = {}, set()
    for i in range(dem.num_errors):
        p, dets, obs = dem.mechanism(i)
        pieces = dem.mechanism_components(i)
        assert tuple(pieces) not in seen
        seen.add(tuple(pieces))
        x, om = set(), 0
        for a, b, o in pieces:
            assert a != NO_DET
            x ^= {a}
            if b != NO_DET:
                x ^= {b}
            om ^= o
        assert sorted(x) == list(dets)
        assert om == sum(1 << k for k in obs)
        assert (len(pieces) == 1) == (len(dets) <= 2)
        merged[(dets, obs)] = merged.get((dets, obs), 0.0) * (1 - p) + p * (1 - merged.get((dets, obs), 0.0))
    assert merged.keys() == plain.keys()
    assert max(abs(merged[k] - plain[k]) for k in plain) < 1e-15
    # the same symptom reached through different decompositions (Stim keys errors by decomposition): 1910 vs 1677
    assert dem.num_errors == 1910 and len(plain) == 1677


def test_in_channel_stage_prefers_single_detector_pieces():
    """Stage 1 of Stim's decomposition: when the single-detector errors of a channel cover a composite error of the
    same channel, they are its decomposition -- not a two-detector edge that happens to exist elsewhere."""
    from graphqec_b200.dem import _combos, _in_channel_keys
    nd = 8
    X_a, Z_a, X_b, Z_b = 0b0001, 0b0110, 0b1000, 0
    sym = _combos((X_a, Z_a, X_b, Z_b))
    keys = _in_channel_keys(sym, (1 << nd) - 1)
    assert keys[0b0111] == (Z_a, X_a, X_b)      # D1 D2 ^ D0 ^ D3: the irreducible pair first, then the singles by channel index
    assert keys[0b0101] == (X_a ^ X_b,)         # two detectors: graphlike as it is
    # a basis error that is itself a hyperedge stays whole for the global pass
    sym = _combos((0b0111, 0b1000))
    assert _in_channel_keys(sym, 255)[1] == (0b0111,) and _in_channel_keys(sym, 255)[3] == (0b1111,)


def test_global_pass_greedy_pairs_remnants_and_failures():
    from graphqec_b200.dem import _global_decomposition_pass
    nd = 10
    L0 = 1 << nd
    base = {(0b0011,): 0.1, (0b1100 | L0,): 0.2, (0b10000,): 0.05}
    # known pairs, observables reproduced: D0 D1 ^ D2 D3 L0
    out = _global_decomposition_pass({**base, **{(0b1111 | L0,): 0.01}}, nd)
    assert out[(0b0011, 0b1100 | L0)] == 0.01
    # D0 D1 known, D5 D6 not: remnant edge D5 D6
    out = _global_decomposition_pass({**base, **{(0b1100011,): 0.03}}, nd)
    assert out[(0b0011, 0b1100000)] == 0.03
    # D0 D1 ^ D2 D3 L0 ^ D4 reproduces the detectors but leaves L0 over with no detector to carry it: failure
    with pytest.raises(ValueError, match="D0, D1, D2, D3, D4"):
        _global_decomposition_pass({**base, **{(0b11111,): 0.02}}, nd)
    # three detectors none of which is part of a known piece: failure
    with pytest.raises(ValueError, match="D5, D6, D7"):
        _global_decomposition_pass({**base, **{(0b11100000,): 0.04}}, nd)
    # two errors that end up with the same decomposition merge
    out = _global_decomposition_pass({**base, **{(0b1111 | L0,): 0.01, (0b0011, 0b1100 | L0): 0.5}}, nd)
    assert abs(out[(0b0011, 0b1100 | L0)] - (0.01 * 0.5 + 0.5 * 0.99)) < 1e-15


def test_decomposition_failure_raises_like_stim():
    """Steane through CssCode: Hx = Hz gives edges that exist with and without the observable, the greedy pass
    cannot reproduce the observables from known pieces and the model falls back to the undecomposed one -- which is
    what reproduces the reference's recorded Steane error rates (tests/test_gpu_reference_stats.py)."""
    H = np.array([[1, 1, 1, 1, 0, 0, 0], [0, 1, 1, 0, 1, 1, 0], [0, 0, 1, 1, 0, 1, 1]])
    st = CssCode(Hx=H, Hz=H, depolarize1_rate=0.005, depolarize2_rate=0.005)
    st.build_memory_circuit(number_of_rounds=3)
    with pytest.raises(ValueError, match="Failed to decompose errors into graphlike components"):
        st.memory_circuit.detector_error_model(decompose_errors=True)
    assert st.memory_circuit.detector_error_model().num_errors == 276


def test_toric_through_csscode_decomposes():
    """BASELINE configs[2]: the hook errors of the serial CssCode schedule (up to 6 detectors) decompose."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_reference_circuits import toric_checks
    hx, hz = toric_checks(4, 4)
    tc = CssCode(Hx=hx, Hz=hz, name="Toric", depolarize1_rate=0.006, depolarize2_rate=0.006)
    tc.build_memory_circuit(number_of_rounds=tc.distance)
    dem = tc.memory_circuit.detector_error_model(decompose_errors=True)
    assert dem.decomposed and dem.num_errors == 2832
    assert tc.memory_circuit.detector_error_model().num_errors == 2368
    a, b, om, pp = dem.graphlike_components()
    assert len(a) == dem.comp_a.shape[0]          # every piece is graphlike


def test_p_zero_gives_empty_dem():
    rep = RepetitionCode(distance=3, depolarize1_rate=0.0, depolarize2_rate=0.0)
    rep.build_memory_circuit(number_of_rounds=3)
    dem = rep.memory_circuit.detector_error_model(decompose_errors=True)
    assert dem.num_errors == 0 and dem.num_detectors == 8 and dem.num_observables == 1


def test_dem_text_roundtrip():
    code = RotatedSurfaceCode(distance=3, depolarize1_rate=0.01, depolarize2_rate=0.01)
    code.build_memory_circuit(number_of_rounds=3)
    for dec in (False, True):
        dem = code.memory_circuit.detector_error_model(decompose_errors=dec)
        back = DetectorErrorModel.from_text(str(dem))
        assert _as_table(back).keys() == _as_table(dem).keys()
        assert back.num_detectors == dem.num_detectors and back.num_observables == dem.num_observables
        assert np.allclose(back.probs, dem.probs, rtol=0, atol=0)
        if dec:
            assert (back.comp_a == dem.comp_a).all() and (back.comp_b == dem.comp_b).all()


def test_dem_text_repeat_and_shift():
    txt = "error(0.1) D0\nrepeat 2 {\n error(0.2) D0 D1\n shift_detectors 1\n}\nerror(0.3) D0 L0\n"
    dem = DetectorErrorModel.from_text(txt)
    assert dem.num_errors == 4 and dem.num_detectors == 3
    assert [dem.mechanism(i)[1] for i in range(4)] == [(0,), (0, 1), (1, 2), (2,)]


# ------------------------------------------------------------------------------------------
# native builder (gq_circuit_to_dem, csrc/circuit_dem.cu) == Python builder, array for array
# ------------------------------------------------------------------------------------------

def _native_cases():
    from graphqec_b200 import BivariateBicycleCode
    steane = np.array([[1, 1, 1, 1, 0, 0, 0], [0, 1, 1, 0, 1, 1, 0], [0, 0, 1, 1, 0, 1, 1]])
    return [
        ("rep5", RepetitionCode(distance=5, depolarize1_rate=0.1, depolarize2_rate=0.1), 5, "Z"),
        ("rot5z", RotatedSurfaceCode(distance=5, depolarize1_rate=0.005, depolarize2_rate=0.005), 5, "Z"),
        ("rot5x", RotatedSurfaceCode(distance=5, depolarize1_rate=0.005, depolarize2_rate=0.003), 4, "X"),
        ("rot4_even", RotatedSurfaceCode(distance=4, depolarize1_rate=0.01, depolarize2_rate=0.01), 4, "Z"),
        ("unrot3", UnrotatedSurfaceCode(distance=3, depolarize1_rate=0.003, depolarize2_rate=0.007), 3, "Z"),
        ("steane", CssCode(Hx=steane, Hz=steane, depolarize1_rate=0.002, depolarize2_rate=0.002), 3, "Z"),
        ("bb72", BivariateBicycleCode(L1=6, L2=6, a1=3, a2=1, a3=2, b1=3, b2=1, b3=2,
                                      depolarize1_rate=0.001, depolarize2_rate=0.001), 2, "Z"),
    ]


@pytest.mark.parametrize("case", _native_cases(), ids=lambda c: c[0])
def test_native_builder_equals_python_builder(case):
    from graphqec_b200.dem import circuit_to_dem, circuit_to_dem_native
    _, code, rounds, lc = case
    code.build_memory_circuit(number_of_rounds=rounds, logic_check=lc)
    circ = code.memory_circuit
    for decompose in (False, True):
        try:
            want = circuit_to_dem(circ, decompose)
        except ValueError as exc:
            with pytest.raises(ValueError) as got:
                circuit_to_dem_native(circ, decompose)
            assert str(got.value) == str(exc)      # same first undecomposable mechanism, same wording
            continue
        got = circuit_to_dem_native(circ, decompose)
        assert (got.num_detectors, got.num_observables, got.decomposed) == (want.num_detectors, want.num_observables, want.decomposed)
        for f in ("probs", "det_indptr", "det_idx", "obs_indptr", "obs_idx", "comp_indptr", "comp_a", "comp_b", "comp_obs"):
            a, b = getattr(want, f), getattr(got, f)
            assert a.dtype == b.dtype and a.shape == b.shape and (a == b).all(), f    # probabilities bit for bit
        assert str(got) == str(want)


def test_native_builder_text_front_end():
    from graphqec_b200 import backend
    from graphqec_b200.circuit import Circuit
    from graphqec_b200.dem import circuit_to_dem
    body = "CNOT 0 1 1 2\nDEPOLARIZE2(0.01) 0 1 1 2  # hook\nMR(0.002) 1\nDETECTOR rec[-1] rec[-2]\n"
    flat = "R 0 1 2\nMR 1\n" + body * 3 + "M 0 2\nOBSERVABLE_INCLUDE(0) rec[-1]\n"
    looped = "R 0 1 2\nMR 1\nREPEAT 3 {\n" + "".join("    " + l + "\n" for l in body.splitlines()) + "}\nM 0 2\nOBSERVABLE_INCLUDE(0) rec[-1]\n"
    want = circuit_to_dem(Circuit(flat), False)
    for text in (flat, looped):
        D, O, p, dp, di, op_, oi, comps = backend.circuit_to_dem_arrays(text, False)
        assert (D, O) == (3, 1) and comps is None
        assert (p == want.probs).all() and (di == want.det_idx).all() and (oi == want.obs_idx).all()
    for bad in ("FOO 1", "CX 0", "DETECTOR rec[-1]", "M 0\nDETECTOR 0", "REPEAT 2 {\nM 0\n"):
        with pytest.raises(ValueError):
            backend.circuit_to_dem_arrays(bad, False)
