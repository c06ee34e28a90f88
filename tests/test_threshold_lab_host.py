"""Host logic of ThresholdLAB: task generation, stop rule, rank sharding (gloo, world size 2)."""
import os
import socket
import sys

import numpy as np
import pytest

from graphqec_b200 import RepetitionCode, RotatedSurfaceCode, ThresholdLAB, wilson_interval
from graphqec_b200.lab.threshold import threshold_lab as tl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_init_roundtrip():
    # reference tests/lab/threshold/test_threshold_lab.py:30-34 (minus the stale attribute)
    th = ThresholdLAB(configurations=[{"distance": d} for d in [3, 5]], code=RepetitionCode,
                      error_rates=np.linspace(0, 0.1, 10))
    assert th.configurations == [{"distance": d} for d in [3, 5]]
    assert th.code == RepetitionCode
    assert (th.error_rates == np.linspace(0, 0.1, 10)).all()
    assert th.decoder == "pymatching"
    assert th.logic_check() == "Z"       # a method in the reference


def test_unknown_decoder_is_not_rejected_at_construction():
    th = ThresholdLAB(configurations=[], code=RepetitionCode, error_rates=[], decoder="nope")
    with pytest.raises(AttributeError):
        th.decoder


def test_tasks_metadata_and_rounds():
    th = ThresholdLAB(configurations=[{"distance": 3}, {"distance": 5}], code=RotatedSurfaceCode,
                      error_rates=[0.001, 0.002])
    tasks = list(th.generate_sinter_tasks())
    assert len(tasks) == 4
    md = tasks[2].json_metadata
    assert md == {"name": "Rotated Surface [[49,1,5]]", "error": 0.001, "n": 49, "k": 1, "d": 5, "r": 5}
    assert tasks[2].circuit.num_detectors == 2 * 12 + 4 * 24
    with pytest.raises(ValueError):
        list(ThresholdLAB(configurations=[{"distance": 3}], code=RepetitionCode, error_rates=[0.1]).generate_sinter_tasks(logic_check="X"))


def test_check_is_family_sense():
    class S:  # minimal stand-in for TaskStats
        def __init__(self, n): self.json_metadata = {"name": n}
    assert ThresholdLAB.check_is_family([S("a"), S("a")]) is False
    assert ThresholdLAB.check_is_family([S("a"), S("b")]) is True


def test_split_batch_covers_range_exactly():
    for world in (1, 2, 3, 8):
        for n in (1, 255, 256, 257, 1000, 16384, 100000):
            parts = [tl._split_batch(512, n, 256, r, world) for r in range(world)]
            assert sum(p[1] for p in parts) == n
            pos = 512
            for s0, cnt in parts:
                if cnt:
                    assert s0 == pos and s0 % 256 == 0
                    pos += cnt


class _FakeDecoder:
    """Counts one 'error' per shot whose global index is divisible by 1000."""

    def collect(self, seed, shot0, nshots, max_errors):
        first = -(-shot0 // 1000) * 1000
        errs = len(range(first, shot0 + nshots, 1000))
        return nshots, errs, 0


def test_stop_rule_single_process():
    shots, errors, _ = tl._run_task(_FakeDecoder(), 256, 1, 100000, 0, 0, None, 0, 1, 0)
    assert (shots, errors) == (100000, 100)
    shots, errors, _ = tl._run_task(_FakeDecoder(), 256, 1, 10**9, 50, 0, None, 0, 1, 0)
    assert errors >= 50 and shots < 10**6    # stopped early, overshoot allowed like sinter


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from graphqec_b200.lab.threshold import threshold_lab as tl2
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    d, r, w = tl2._dist_info()
    out = tl2._run_task(_FakeDecoder(), 256, 1, 300000, 0, 0, d, r, w, 0)
    q.put((rank, out))
    dist.destroy_process_group()


def test_rank_sharding_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs: p.join(timeout=60)
    # both ranks agree, and the totals equal the single-process run: independent of world size
    assert res[0] == res[1] == (300000, 300, 0)


def test_wilson_interval():
    lo, hi = wilson_interval(10, 1000)
    assert 0.0048 < lo < 0.0055 and 0.0182 < hi < 0.0185
    assert wilson_interval(0, 0) == (0.0, 1.0)


def test_curves_without_matplotlib():
    from graphqec_b200.stats import TaskStats
    th = ThresholdLAB(configurations=[{"distance": 3}], code=RepetitionCode, error_rates=[0.1])
    th._samples = [TaskStats("x", "pymatching", {"name": "A", "error": 0.1, "r": 3, "k": 1}, shots=1000, errors=100)]
    (name, xs, ys, lo, hi), = th.error_rate_curves("per_round")
    assert name == "A" and xs == [0.1]
    assert abs(ys[0] - (0.5 - 0.5 * (1 - 0.2) ** (1 / 3))) < 1e-12 and lo[0] < ys[0] < hi[0]


def test_csv_roundtrip_in_sinter_schema(tmp_path):
    from graphqec_b200.stats import CSV_HEADER, TaskStats, append_stats_to_csv, read_stats_from_csv
    path = str(tmp_path / "stats.csv")
    a = TaskStats("id-a", "pymatching", {"name": 'Rot "x"', "error": 0.001, "d": 5, "r": 5}, shots=1000, errors=7, seconds=1.5)
    b = TaskStats("id-b", "pymatching", {"name": "B", "error": 0.002}, shots=10, errors=1)
    append_stats_to_csv(path, a)
    append_stats_to_csv(path, b)
    append_stats_to_csv(path, TaskStats("id-a", "pymatching", a.json_metadata, shots=500, errors=3, discards=2, seconds=0.5))
    lines = open(path).read().splitlines()
    assert lines[0] == CSV_HEADER and len(lines) == 4
    got = {s.strong_id: s for s in read_stats_from_csv(path)}
    assert (got["id-a"].shots, got["id-a"].errors, got["id-a"].discards) == (1500, 10, 2)
    assert got["id-a"].json_metadata == a.json_metadata and got["id-b"].shots == 10
    assert read_stats_from_csv(str(tmp_path / "missing.csv")) == []


def test_collect_stats_resumes_from_csv(tmp_path, monkeypatch):
    """save_resume_filepath: recorded rows count towards the stop rule, new shots continue the RNG stream."""
    from graphqec_b200 import backend as be
    calls = []

    class FakeDem:
        tile_shots = 256
        def __init__(self, dem, device=0): pass
        def close(self): pass

    class FakeDec(_FakeDecoder):
        def __init__(self, demh, kind, **kw): pass
        def collect(self, seed, shot0, nshots, max_errors):
            calls.append((shot0, nshots))
            return super().collect(seed, shot0, nshots, max_errors)
        def close(self): pass

    monkeypatch.setattr(be, "DemHandle", FakeDem)
    monkeypatch.setattr(be, "DecoderHandle", FakeDec)
    path = str(tmp_path / "resume.csv")
    th = ThresholdLAB(configurations=[{"distance": 3}], code=RepetitionCode, error_rates=[0.05])
    th.collect_stats(max_shots=20000, max_errors=0, save_resume_filepath=path)
    assert th.samples[0].shots == 20000 and th.samples[0].errors == 20
    first = list(calls)
    calls.clear()
    # same budget again: nothing left to do, the totals come from the file
    th.collect_stats(max_shots=20000, max_errors=0, save_resume_filepath=path)
    assert calls == [] and th.samples[0].shots == 20000
    # a bigger budget continues behind the recorded shots (next whole tile)
    th.collect_stats(max_shots=50000, max_errors=0, decoder_params={"save_resume_filepath": path})
    assert calls[0][0] == -(-20000 // 256) * 256 and sum(n for _, n in calls) == 30000
    assert th.samples[0].shots == 50000
    assert len(open(path).read().splitlines()) == 3      # header + one row per run that did work
    assert first[0][0] == 0
    # budgets that are not tile multiples -- run 1 used [0, 20000), run 2 [20224, 50224); a third run must
    # start behind 50224 (next_shot recorded per row), not at ceil(50000 / 256) * 256 = 50176, which run 2 already drew
    calls.clear()
    th.collect_stats(max_shots=60000, max_errors=0, save_resume_filepath=path)
    run2_end = -(-20000 // 256) * 256 + -(-30000 // 256) * 256
    assert calls[0][0] == run2_end and calls[0][0] >= 20224 + 30000
    assert th.samples[0].shots == 60000 and th.samples[0].custom_counts["next_shot"] == run2_end + -(-10000 // 256) * 256
    from graphqec_b200.stats import read_stats_from_csv
    merged = read_stats_from_csv(path)[0]
    assert merged.shots == 60000 and merged.custom_counts["next_shot"] == th.samples[0].custom_counts["next_shot"]
