"""Logical error rates the REFERENCE recorded, read off the plots stored in its notebooks (run HERE).

The reference holds no numeric LER for the surface / Steane sweeps -- only the PNGs that
``ThresholdLAB.plot_stats`` (``threshold_lab.py:199-259`` -> ``sinter.plot_error_rate``) left in the notebook
outputs.  This script decodes those PNGs and locates the plot markers by colour (matplotlib's default cycle),
then converts pixel rows into error rates with the axis limits passed to ``plot_stats`` in the same cell
(or, for the autoscaled Steane figure, with the major tick marks).  x positions are matched to the sweep's
``np.linspace`` grid, so only the y read-off carries pixel error (+-1 px = +-3 %).

Each point was produced by sinter + PyMatching on the reference's circuits with ``max_shots = 10^5`` and
``max_errors = 1000`` (the cells' arguments), so it also carries its own binomial error, recorded as
``approx_errors`` (= LER x shots, shots = min(10^5, 1000 / LER)).  Output: notebook_ler_readoffs.json;
consumed by tests/test_gpu_reference_stats.py.  The exact sinter TaskStats of the toric notebook are copied
verbatim into ``toric_taskstats`` (``notebooks/toric_code.ipynb`` output of ``th.samples``).
"""
import base64
import io
import json
import os
import re

import numpy as np
from PIL import Image
from scipy import ndimage

REF = "/root/reference/notebooks"
HERE = os.path.dirname(os.path.abspath(__file__))
CYCLE = [(31, 119, 180), (255, 127, 14), (44, 160, 44), (214, 39, 40)]   # matplotlib tab:blue, orange, green, red

# notebook, cell index, code, logic_check, series (legend order = colour order) as constructor kwargs, sweep grid,
# axis limits (x_min, x_max, y_min, y_max) given to plot_stats, or None = calibrate y on the major ticks
FIGURES = [
    ("rotated_surface_code", 5, "RotatedSurfaceCode", "Z", [{"distance": 11}, {"distance": 5}, {"distance": 7}, {"distance": 3}],
     (0.0, 0.02, 10), (2e-3, 2.2e-2, 1e-5, 6e-1)),
    ("rotated_surface_code", 6, "RotatedSurfaceCode", "X", [{"distance": 11}, {"distance": 5}, {"distance": 7}, {"distance": 3}],
     (0.0, 0.02, 10), (2e-3, 2.2e-2, 1e-5, 6e-1)),
    ("unrotated_surface_code", 4, "UnrotatedSurfaceCode", "Z", [{"distance": 7}, {"distance": 3}, {"distance": 11}, {"distance": 5}],
     (0.0, 0.02, 10), (2e-3, 2.2e-2, 1e-5, 6e-1)),
    ("unrotated_surface_code", 5, "UnrotatedSurfaceCode", "X", [{"distance": 7}, {"distance": 3}, {"distance": 11}, {"distance": 5}],
     (0.0, 0.02, 10), (2e-3, 2.2e-2, 1e-5, 6e-1)),
    ("steane", 3, "CssCode", "Z", [{"steane": True, "name": "Steane"}], (0.0, 0.02, 10), None),
]
LEGEND_BOX = {"steane": (175, 35)}     # legend extent right / below the axes' top-left corner, px (default: 4-entry legend)


def png_of(nb, cell):
    d = json.load(open(f"{REF}/{nb}.ipynb"))
    for o in d["cells"][cell].get("outputs", []):
        if "data" in o and "image/png" in o["data"]:
            return np.array(Image.open(io.BytesIO(base64.b64decode(o["data"]["image/png"]))).convert("RGB")).astype(int)
    raise KeyError((nb, cell))


def frame(im):
    dark = im.sum(axis=2) < 120
    cols = np.where(dark.sum(axis=0) > im.shape[0] * 0.5)[0]
    rows = np.where(dark.sum(axis=1) > im.shape[1] * 0.5)[0]
    return cols.min(), cols.max(), rows.min(), rows.max()


def markers(im, colour, legend_box):
    m = np.abs(im - np.array(colour)).sum(axis=2) < 30
    er = ndimage.binary_erosion(m, iterations=2)          # removes the 1.5 px line, keeps the 6 px markers
    lab, n = ndimage.label(er)
    out = []
    for (y, x), s in zip(ndimage.center_of_mass(er, lab, range(1, n + 1)), ndimage.sum(er, lab, range(1, n + 1))):
        if s >= 4 and not (x < legend_box[0] and y < legend_box[1]):
            out.append((float(x), float(y)))
    return sorted(out)


def major_tick_rows(im, x0):
    dark = im.sum(axis=2) < 150
    return np.where(dark[:, x0 - 4] & dark[:, x0 - 3])[0].tolist()      # major ticks are 3.5 pt long, minor ones 2 pt


def main():
    out = {"generator": "tests/golden/make_notebook_ler_readoffs.py", "points": [], "toric_taskstats": []}
    for nb, cell, code, lc, series, (g0, g1, gn), lims in FIGURES:
        im = png_of(nb, cell)
        x0, x1, y0, y1 = frame(im)
        grid = np.linspace(g0, g1, gn)
        if lims is None:
            ticks = major_tick_rows(im, x0)                  # decades, top to bottom
            assert len(ticks) >= 2
            per_decade = (ticks[-1] - ticks[0]) / (len(ticks) - 1)
            top_decade = -1                                  # steane: labels 10^-1 ... 10^-4 (checked by eye)
            y_of = lambda r: 10 ** (top_decade - (r - ticks[0]) / per_decade)
            # x: autoscaled too -- match markers to the grid by order (p = 0 is not drawable on a log axis)
            x_of = None
        else:
            xmin, xmax, ymin, ymax = lims
            y_of = lambda r: ymin * (ymax / ymin) ** ((y1 - r) / (y1 - y0))
            x_of = lambda c: xmin * (xmax / xmin) ** ((c - x0) / (x1 - x0))
        for colour, kw in zip(CYCLE, series):
            lb = LEGEND_BOX.get(nb, (290, 100))
            pts = markers(im, colour, legend_box=(x0 + lb[0], y0 + lb[1]))
            assert x_of is not None or len(pts) == gn - 1, (nb, cell, len(pts))
            for i, (c, r) in enumerate(pts):
                if x_of is not None:
                    p_px = x_of(c)
                    k = int(np.argmin(np.abs(np.log(grid[1:]) - np.log(p_px)))) + 1
                    assert abs(np.log(grid[k] / p_px)) < 0.03, (nb, cell, kw, p_px, grid[k])
                else:
                    k = i + 1                                # every p > 0 of the grid is on the autoscaled figure
                if r >= y1 - 4:                              # clipped at the bottom of the axes
                    continue
                ler = float(y_of(r))
                shots = min(1e5, 1000 / ler)
                out["points"].append({"notebook": f"notebooks/{nb}.ipynb", "cell": cell, "code": code, "spec": kw, "logic_check": lc,
                                      "p": float(grid[k]), "ler": ler, "approx_errors": round(ler * shots, 1),
                                      "max_shots": 100000, "max_errors": 1000})
    # exact numbers: the toric notebook's th.samples
    d = json.load(open(f"{REF}/toric_code.ipynb"))
    for c in d["cells"]:
        for o in c.get("outputs", []):
            if o.get("output_type") == "execute_result":
                txt = "".join(o["data"]["text/plain"])
                for name, err, shots, errors, secs in re.findall(
                        r"json_metadata=\{'name': '([^']+)', 'error': ([0-9.e-]+)\}, shots=(\d+), errors=(\d+), seconds=([0-9.]+)", txt):
                    L = int(re.search(r",2,(\d+)\]\]", name).group(1))
                    out["toric_taskstats"].append({"name": name, "L": L, "p": float(err), "shots": int(shots), "errors": int(errors),
                                                   "seconds": float(secs), "max_shots": 10**6, "max_errors": 10**5})
    path = os.path.join(HERE, "notebook_ler_readoffs.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, len(out["points"]), "plot points,", len(out["toric_taskstats"]), "toric rows")
    for pt in out["points"]:
        print(pt["code"], pt["spec"], pt["logic_check"], "p=%.5f LER=%.3e (~%d errors)" % (pt["p"], pt["ler"], pt["approx_errors"]))


if __name__ == "__main__":
    main()
