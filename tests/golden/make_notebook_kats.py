"""Extract the circuit known-answer vectors the reference recorded in its notebooks.

Run HERE (needs /root/reference); writes tests/golden/notebook_circuits.json.  Each entry holds the
constructor call and, read off the Stim timeline diagram stored in the notebook output, the
measurement count, every detector's record set (absolute record indices) and the observable's.
Sources: notebooks/repetition_code.ipynb:41-58, notebooks/rotated_surface_code.ipynb:90-123,178-211,
notebooks/unrotated_surface_code.ipynb:106-155.
"""
import json, re, os

REF = "/root/reference/notebooks"
SPECS = [
    ("repetition_code", 1, {"code": "RepetitionCode", "kwargs": {"distance": 5, "depolarize1_rate": 0.05, "depolarize2_rate": 0.05}, "rounds": 2, "logic_check": "Z"}),
    ("rotated_surface_code", 3, {"code": "RotatedSurfaceCode", "kwargs": {"distance": 3, "depolarize1_rate": 0.01, "depolarize2_rate": 0.01}, "rounds": 3, "logic_check": "X"}),
    ("rotated_surface_code", 4, {"code": "RotatedSurfaceCode", "kwargs": {"distance": 3, "depolarize1_rate": 0.01, "depolarize2_rate": 0.01}, "rounds": 2, "logic_check": "X"}),
    ("unrotated_surface_code", 3, {"code": "UnrotatedSurfaceCode", "kwargs": {"distance": 3, "depolarize1_rate": 0.01, "depolarize2_rate": 0.01}, "rounds": 2, "logic_check": "Z"}),
]
out = []
for nb, cell, spec in SPECS:
    d = json.load(open(f"{REF}/{nb}.ipynb"))
    c = d["cells"][cell]
    txt = ""
    for o in c.get("outputs", []):
        if "text" in o:
            txt += "".join(o["text"])
        for k, v in o.get("data", {}).items():
            if k.startswith("text"):
                txt += "".join(v)
    dets = {int(a): sorted(int(x) for x in re.findall(r"rec\[(\d+)\]", b)) for a, b in re.findall(r"DETECTOR:D(\d+)=([rec\[\]\d\*]+)", txt)}
    obs = {int(a): sorted(int(x) for x in re.findall(r"rec\[(\d+)\]", b)) for a, b in re.findall(r"OBSERVABLE_INCLUDE:L(\d+)\*=([rec\[\]\d\*]+)", txt)}
    meas = sorted({int(x) for x in re.findall(r"M[RX]?:rec\[(\d+)\]", txt)})
    entry = dict(spec)
    entry.update({"source": f"notebooks/{nb}.ipynb cell {cell}", "num_measurements": len(meas),
                  "detectors": [dets[i] for i in range(len(dets))], "observable0": obs[0]})
    out.append(entry)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "notebook_circuits.json")
json.dump(out, open(path, "w"), indent=1)
print("wrote", path, [len(e["detectors"]) for e in out])
