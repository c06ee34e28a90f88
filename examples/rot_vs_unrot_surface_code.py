"""Rotated against unrotated surface code: logical error rate per round over sqrt(number of qubits), with a
log-linear fit per physical error rate -- the workload of the reference's
``examples/rot_vs_unrot_surface_code.py`` (``main`` ``:24-204``, ``fit`` ``:207-228``, grid ``:233-237``) on the B200 path.

Same entry points and arguments as the reference's script (``main(configs, errors, max_shots, max_errors, logic_check)``,
``fit(collected_stats, error)``); what differs is what is underneath: ``ThresholdLAB.collect_stats`` runs on the GPU,
the per-round conversion and the line fit are a few lines of numpy (no sinter / scipy needed), the JSON is always
written and the PNG only when matplotlib is installed.

  python examples/rot_vs_unrot_surface_code.py                       # the reference's grid: d in 9..23, p in 1e-3..1e-2, 1e6 shots
  python examples/rot_vs_unrot_surface_code.py --distances 3 5 7 --errors 0.002 0.004 0.006 --max-shots 200000
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from collections import namedtuple

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from graphqec import RotatedSurfaceCode, ThresholdLAB, UnrotatedSurfaceCode  # noqa: E402

LineFit = namedtuple("LineFit", "slope intercept rvalue")


def shot_error_rate_to_piece_error_rate(shot_error_rate: float, pieces: int) -> float:
    """sinter.shot_error_rate_to_piece_error_rate: the per-round rate whose XOR over ``pieces`` rounds gives the per-shot rate."""
    if shot_error_rate > 0.5:
        return 1 - shot_error_rate_to_piece_error_rate(1 - shot_error_rate, pieces)
    return 0.5 - 0.5 * (1 - 2 * shot_error_rate) ** (1.0 / pieces)


def fit(collected_stats, error: float):
    """Line fit of log(per-round logical error rate) against sqrt(n) for one physical error rate (reference ``:207-228``)."""
    xs, ys, log_ys = [], [], []
    for stats in collected_stats:
        if stats.errors == 0 or stats.json_metadata["error"] != error:
            continue
        per_shot = stats.errors / stats.shots
        per_round = shot_error_rate_to_piece_error_rate(per_shot, pieces=stats.json_metadata["r"])
        xs.append(float(np.sqrt(stats.json_metadata["n"])))
        ys.append(per_round)
        log_ys.append(float(np.log(per_round)))
    if len(xs) < 2:
        return LineFit(float("nan"), float("nan"), float("nan")), xs, ys
    slope, intercept = np.polyfit(xs, log_ys, 1)
    r = float(np.corrcoef(xs, log_ys)[0, 1]) if len(xs) > 2 else 1.0
    return LineFit(float(slope), float(intercept), r), xs, ys


def main(configs, errors: np.ndarray, max_shots: int, max_errors: int, logic_check: str, out_prefix: str = "rot_vs_unrot_surface_code"):
    labs = {}
    for kind, code in (("Rotated", RotatedSurfaceCode), ("Unrotated", UnrotatedSurfaceCode)):
        th = ThresholdLAB(configurations=configs, code=code, error_rates=errors, decoder="pymatching")
        th.collect_stats(max_shots=max_shots, max_errors=max_errors, logic_check=logic_check)
        labs[kind] = th
    ns = [np.sqrt(s.json_metadata["n"]) for th in labs.values() for s in th.samples]
    max_n, min_n = float(max(ns)), float(min(ns))
    data = []
    for error in errors:
        if error == 0:
            continue
        for kind in ("Unrotated", "Rotated"):
            f, xs, ys = fit(labs[kind].samples, error)
            data.append({
                "n": [int(round(n ** 2)) for n in xs],
                "logical_error": [float(y) for y in ys],
                "type": kind,
                "error_rate": float(error),
                "fit": {"x": [0, max_n], "y": [float(np.exp(f.intercept)), float(np.exp(f.intercept + f.slope * max_n))],
                        "slope": f.slope, "intercept": f.intercept},
            })
    with open(out_prefix + ".json", "w") as fh:
        json.dump(data, fh, indent=4)
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        return data
    cmap = matplotlib.colormaps["tab10"]
    fig, ax = plt.subplots(1, 1)
    for row in data:
        colour = cmap(list(errors).index(row["error_rate"]) % 10)
        ax.scatter(np.sqrt(row["n"]), row["logical_error"], marker="^" if row["type"] == "Unrotated" else "o", color=colour, s=50, zorder=2)
        ax.plot(row["fit"]["x"], row["fit"]["y"], linestyle="--", color=colour, dashes=(2, 2), alpha=0.8, linewidth=1.0, zorder=2)
    ax.semilogy()
    ax.set_xlim(min_n - 2, max_n + 2)
    ax.set_xlabel(r"$\sqrt{\mathrm{Number\ of\ Physical\ Qubits}}$")
    ax.set_ylabel("Logical Error Rate per Round")
    ax.grid(which="major", zorder=0)
    ax.grid(which="minor", zorder=0)
    fig.savefig(out_prefix + ".png", bbox_inches="tight")
    return data


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--distances", type=int, nargs="+", default=[9, 11, 13, 17, 20, 23])          # reference :236
    ap.add_argument("--errors", type=float, nargs="+", default=list(np.linspace(0.001, 0.01, 10)))  # reference :237
    ap.add_argument("--max-shots", type=int, default=1_000_000)
    ap.add_argument("--max-errors", type=int, default=1000)
    ap.add_argument("--logic-check", default="Z")
    ap.add_argument("--out", default="rot_vs_unrot_surface_code")
    a = ap.parse_args()
    rows = main(configs=[{"distance": d} for d in a.distances], errors=np.array(a.errors), max_shots=a.max_shots,
                max_errors=a.max_errors, logic_check=a.logic_check, out_prefix=a.out)
    for row in rows:
        print(row["type"], "p=%.4f" % row["error_rate"], "n", row["n"], "per-round LER", ["%.3e" % y for y in row["logical_error"]],
              "slope %.3f" % row["fit"]["slope"])
