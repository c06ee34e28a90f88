"""End-to-end through the reference-facing API: ThresholdLAB.collect_stats on the GPU."""
import numpy as np
import pytest

from graphqec_b200 import RepetitionCode, RotatedSurfaceCode, ThresholdLAB, wilson_interval
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def test_readme_repetition_sweep():
    """README.md:75-91 of the reference: repetition code threshold sweep (BASELINE configs[0])."""
    th = ThresholdLAB(configurations=[{"distance": d} for d in [3, 5, 7, 11]], code=RepetitionCode,
                      error_rates=np.linspace(0, 0.2, 10), decoder="pymatching")
    assert th.collect_stats(num_workers=4, max_shots=10**5, max_errors=1000, logic_check="Z") is None
    s = th.samples
    assert len(s) == 40
    assert s[0].json_metadata == {"name": "Repetition [[5,1,3]]", "error": 0.0, "n": 5, "k": 1, "d": 3, "r": 3}
    for st in s:
        if st.json_metadata["error"] == 0:
            assert (st.shots, st.errors) == (10**5, 0)     # p = 0 runs to max_shots with no failure
        else:
            assert st.shots == 10**5 or st.errors >= 1000
            assert st.discards == 0 and st.seconds > 0
    # below threshold (p ~ 0.022) larger distance is better; far above (p = 0.2) it is not
    ler = {(st.json_metadata["d"], round(st.json_metadata["error"], 3)): st.errors / st.shots for st in s}
    assert ler[(11, 0.022)] < ler[(3, 0.022)]
    assert ler[(11, 0.2)] > 0.3
    curves = th.error_rate_curves()
    assert len(curves) == 4


def test_rotated_point_inside_oracle_interval():
    th = ThresholdLAB(configurations=[{"distance": 5}], code=RotatedSurfaceCode, error_rates=[0.008])
    th.collect_stats(max_shots=200_000, max_errors=10**9)
    st = th.samples[0]
    assert st.shots == 200_000
    code = RotatedSurfaceCode(distance=5, depolarize1_rate=0.008, depolarize2_rate=0.008)
    code.build_memory_circuit(number_of_rounds=5)
    dem = code.memory_circuit.detector_error_model(decompose_errors=True)
    n, e = po.collect(dem, seed=5, nshots=100_000, nthreads=8)
    lo, hi = po.wilson(e, n)
    glo, ghi = wilson_interval(st.errors, st.shots)
    assert glo <= hi and lo <= ghi


def test_unsupported_decoder_raises():
    th = ThresholdLAB(configurations=[{"distance": 3}], code=RepetitionCode, error_rates=[0.1], decoder="mwpf")
    with pytest.raises(NotImplementedError):
        th.collect_stats(max_shots=100)


def test_sinter_style_adapters():
    """The plug-in seam of the reference (custom_decoders): compile_decoder_for_dem / decode_shots_bit_packed."""
    from graphqec_b200.sinter_compat import B200Decoder, B200Sampler
    from graphqec_b200.stats import Task

    code = RotatedSurfaceCode(distance=3, depolarize1_rate=0.01, depolarize2_rate=0.01)
    code.build_memory_circuit(number_of_rounds=3)
    dem = code.memory_circuit.detector_error_model(decompose_errors=True)
    compiled = B200Decoder().compile_decoder_for_dem(dem=dem)
    dets, obs = po.sample(dem, 1, 4000)
    pred = compiled.decode_shots_bit_packed(bit_packed_detection_event_data=dets)
    assert pred.shape == (4000, 1) and pred.dtype == np.uint8
    ref = po.MwpmOracle(dem).decode_batch(dets, nthreads=4)
    assert (pred[:, 0] == ref).mean() > 0.97           # equal up to ties between minimum-weight matchings
    # DEM given as Stim text works too
    from graphqec_b200 import DetectorErrorModel
    again = B200Decoder().compile_decoder_for_dem(dem=DetectorErrorModel.from_text(str(dem)))
    assert (again.decode_shots_bit_packed(bit_packed_detection_event_data=dets) == pred).all()
    st = B200Sampler().compiled_sampler_for_task(Task(circuit=code.memory_circuit)).sample(10000)
    assert st.shots >= 10000 and 0 < st.errors < st.shots // 10


def test_decode_via_files(tmp_path):
    """sinter's file form of the decoder plug-in: .dem text + b8 detection events in, b8 predictions out."""
    from graphqec_b200.sinter_compat import B200Decoder

    code = RotatedSurfaceCode(distance=5, depolarize1_rate=0.006, depolarize2_rate=0.006)
    code.build_memory_circuit(number_of_rounds=5)
    dem = code.memory_circuit.detector_error_model(decompose_errors=True)
    dets, obs = po.sample(dem, 3, 3001)
    (tmp_path / "m.dem").write_text(str(dem))
    dets.tofile(tmp_path / "dets.b8")
    B200Decoder().decode_via_files(num_shots=3001, num_dets=dem.num_detectors, num_obs=dem.num_observables,
                                   dem_path=tmp_path / "m.dem", dets_b8_in_path=tmp_path / "dets.b8",
                                   obs_predictions_b8_out_path=tmp_path / "pred.b8", tmp_dir=tmp_path)
    pred = np.fromfile(tmp_path / "pred.b8", dtype=np.uint8).reshape(3001, 1)
    direct = B200Decoder().compile_decoder_for_dem(dem=dem).decode_shots_bit_packed(bit_packed_detection_event_data=dets)
    assert (pred == direct).all()
    with pytest.raises(ValueError):
        B200Decoder().decode_via_files(num_shots=3000, num_dets=dem.num_detectors, num_obs=1, dem_path=tmp_path / "m.dem",
                                       dets_b8_in_path=tmp_path / "dets.b8", obs_predictions_b8_out_path=tmp_path / "p2.b8")


def test_bposd_name_runs_min_sum_bp():
    th = ThresholdLAB(configurations=[{"distance": 3}], code=RotatedSurfaceCode, error_rates=[0.002], decoder="bposd")
    th.collect_stats(max_shots=20000, max_errors=10**9)
    st = th.samples[0]
    assert st.shots == 20000 and st.errors < 2000


def test_x_memory_point_inside_oracle_interval():
    """logic_check="X" (three of the reference's notebooks run X-memory experiments)."""
    th = ThresholdLAB(configurations=[{"distance": 5}], code=RotatedSurfaceCode, error_rates=[0.008], logic_check="X")
    th.collect_stats(max_shots=200_000, max_errors=10**9, logic_check="X")
    st = th.samples[0]
    assert st.shots == 200_000 and st.json_metadata["d"] == 5
    code = RotatedSurfaceCode(distance=5, depolarize1_rate=0.008, depolarize2_rate=0.008)
    code.build_memory_circuit(number_of_rounds=5, logic_check="X")
    dem = code.memory_circuit.detector_error_model(decompose_errors=True)
    n, e = po.collect(dem, seed=6, nshots=100_000, nthreads=8)
    lo, hi = po.wilson(e, n)
    glo, ghi = wilson_interval(st.errors, st.shots)
    assert glo <= hi and lo <= ghi


def test_collect_stats_resume_on_the_device(tmp_path):
    """save_resume_filepath through the real backend: a second call with the same budget does no work, a bigger
    budget adds exactly the difference, and the CSV totals equal the samples."""
    from graphqec_b200.stats import read_stats_from_csv
    path = str(tmp_path / "sweep.csv")
    th = ThresholdLAB(configurations=[{"distance": 3}], code=RotatedSurfaceCode, error_rates=[0.004, 0.008])
    th.collect_stats(max_shots=50_000, max_errors=10**9, save_resume_filepath=path)
    first = [(s.shots, s.errors) for s in th.samples]
    th.collect_stats(max_shots=50_000, max_errors=10**9, save_resume_filepath=path)
    assert [(s.shots, s.errors) for s in th.samples] == first
    th.collect_stats(max_shots=120_000, max_errors=10**9, save_resume_filepath=path)
    assert [s.shots for s in th.samples] == [120_000, 120_000]
    rows = {s.strong_id: s for s in read_stats_from_csv(path)}
    for s in th.samples:
        assert (rows[s.strong_id].shots, rows[s.strong_id].errors) == (s.shots, s.errors)
        assert s.errors > first[0][1] // 2
