"""Golden circuits produced by the REFERENCE's own builders (run HERE; needs /root/reference).

The reference (``/root/reference/graphqec``) is pure Python whose only hard dependency on the build
side is ``stim.Circuit`` / ``stim.target_rec``.  This script imports the reference package itself with

* ``sys.modules["stim"]``  = a recorder: ``Circuit.append`` / ``append_from_stim_program_text`` calls are
  forwarded to ``graphqec_b200.circuit.Circuit`` (the text printer) AND logged verbatim, un-fused, so
  the fixture also holds the number of raw ``append`` calls the reference made;
* inert stubs for ``sinter`` (``Task`` only), ``matplotlib``, ``stimbposd``, ``beliefmatching``, ``mwpf``

and runs ``<Code>(**kwargs).build_memory_circuit(...)`` (``base_code.py:184-327``) and
``ThresholdLAB.generate_sinter_tasks`` (``threshold_lab.py:118-151``) for every BASELINE.json config.
For each case it writes: sha256 of the circuit text, instruction / measurement / detector counts, the
number of raw append calls, ``name`` / ``distance`` / qubit counts, the logical operators, and the full
text for small circuits.  ``tests/test_reference_pin.py`` asserts that ``graphqec_b200``'s builders print
the identical text.  Output: ``tests/golden/reference_circuits.json``.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from graphqec_b200 import circuit as pc  # noqa: E402  (text printer only)

FULL_TEXT_LIMIT = 6000  # characters; bigger circuits are pinned by sha256 + counts


# ----------------------------------------------------------------------------- shims
class RecCircuit(pc.Circuit):
    """Recorder standing in for stim.Circuit while the reference builder runs."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.raw_appends = 0

    def append(self, name, targets=(), arg=None):
        self.raw_appends += 1
        return super().append(name, targets, arg)


def install_shims():
    stim = types.ModuleType("stim")
    stim.Circuit = RecCircuit
    stim.target_rec = pc.target_rec
    sys.modules["stim"] = stim

    sinter = types.ModuleType("sinter")

    class Task:
        def __init__(self, circuit=None, json_metadata=None, **_):
            self.circuit = circuit
            self.json_metadata = json_metadata

    sinter.Task = Task
    sinter.TaskGenerator = object
    sinter.TaskStats = object
    sys.modules["sinter"] = sinter

    mpl = types.ModuleType("matplotlib")
    for sub in ("pyplot", "patches"):
        m = types.ModuleType(f"matplotlib.{sub}")
        setattr(mpl, sub, m)
        sys.modules[f"matplotlib.{sub}"] = m
    sys.modules["matplotlib"] = mpl
    for mod, attr in (("stimbposd", "SinterDecoder_BPOSD"), ("beliefmatching", "BeliefMatchingSinterDecoder"),
                      ("mwpf", "SinterMWPFDecoder")):
        m = types.ModuleType(mod)
        setattr(m, attr, type(attr, (), {}))
        sys.modules[mod] = m
    sys.path.insert(0, REF)


# ----------------------------------------------------------------------------- cases
def toric_checks(L1, L2):
    """The hypergraph-product construction of notebooks/toric_code.ipynb (cell 2), in numpy."""
    import numpy as np

    def rep(n):
        h = np.zeros((n, n), dtype=np.uint8)
        for i in range(n):
            h[i, i] = 1
            h[i, (i + 1) % n] = 1
        return h

    h1, h2 = rep(L1), rep(L2)
    hx = np.hstack([np.kron(h1, np.eye(L2, dtype=np.uint8)), np.kron(np.eye(L1, dtype=np.uint8), h2.T)]) % 2
    hz = np.hstack([np.kron(np.eye(L1, dtype=np.uint8), h2), np.kron(h1.T, np.eye(L2, dtype=np.uint8))]) % 2
    return hx.astype(int), hz.astype(int)


H_STEANE = [[1, 1, 1, 1, 0, 0, 0], [0, 1, 1, 0, 1, 1, 0], [0, 0, 1, 1, 0, 1, 1]]
BB144 = {"L1": 12, "L2": 6, "a1": 3, "a2": 1, "a3": 2, "b1": 3, "b2": 1, "b3": 2}
BB72 = {"L1": 6, "L2": 6, "a1": 3, "a2": 1, "a3": 2, "b1": 3, "b2": 1, "b3": 2}


def cases():
    """(id, code class name, kwargs-spec, rounds or None (= distance), logic_check, p)."""
    out = []
    for d in (3, 5, 7, 11):                                  # BASELINE configs[0]
        out.append((f"rep{d}_r2", "RepetitionCode", {"distance": d}, 2, "Z", 0.1))
        out.append((f"rep{d}_rd", "RepetitionCode", {"distance": d}, None, "Z", 0.1))
    out.append(("rep5_p0", "RepetitionCode", {"distance": 5}, None, "Z", 0.0))
    for d in (3, 5, 7, 9, 11):                               # configs[1] (+ d = 3)
        out.append((f"rot{d}_Z", "RotatedSurfaceCode", {"distance": d}, None, "Z", 0.005))
    out.append(("rot5_X", "RotatedSurfaceCode", {"distance": 5}, None, "X", 0.005))
    out.append(("rot11_X", "RotatedSurfaceCode", {"distance": 11}, None, "X", 0.005))
    out.append(("rot4_Z", "RotatedSurfaceCode", {"distance": 4}, None, "Z", 0.005))        # even d
    out.append(("rot20_Z", "RotatedSurfaceCode", {"distance": 20}, None, "Z", 0.005))
    out.append(("rot23_Z", "RotatedSurfaceCode", {"distance": 23}, None, "Z", 0.005))      # examples/...:236
    for d in (3, 5, 15):                                     # configs[2]
        out.append((f"unrot{d}_Z", "UnrotatedSurfaceCode", {"distance": d}, None, "Z", 0.005))
    out.append(("unrot5_X", "UnrotatedSurfaceCode", {"distance": 5}, None, "X", 0.005))
    out.append(("unrot4_Z", "UnrotatedSurfaceCode", {"distance": 4}, None, "Z", 0.005))
    for L in (3, 4, 8, 12, 15):
        out.append((f"toric{L}_Z", "CssCode", {"toric": L, "name": "Toric"}, None, "Z", 0.006))
    out.append(("toric4_X", "CssCode", {"toric": 4, "name": "Toric"}, None, "X", 0.006))
    out.append(("steane_Z", "CssCode", {"steane": True, "name": "Steane"}, None, "Z", 0.005))  # configs[3]
    out.append(("steane_X", "CssCode", {"steane": True, "name": "Steane"}, None, "X", 0.005))
    out.append(("steane_noname_Z", "CssCode", {"steane": True}, None, "Z", 0.002))
    out.append(("bb72_Z", "BivariateBicycleCode", dict(BB72), None, "Z", 0.001))
    out.append(("bb144_Z", "BivariateBicycleCode", dict(BB144), None, "Z", 0.005))            # configs[4]
    out.append(("bb144_X", "BivariateBicycleCode", dict(BB144), None, "X", 0.005))
    return out


def materialise(spec: dict) -> dict:
    """Turn the JSON-able kwargs spec into constructor kwargs (check matrices as int ndarrays)."""
    import numpy as np

    kw = {k: v for k, v in spec.items() if k not in ("toric", "steane")}
    if "toric" in spec:
        hx, hz = toric_checks(spec["toric"], spec["toric"])
        kw.update(Hx=hx, Hz=hz)
    if spec.get("steane"):
        kw.update(Hx=np.array(H_STEANE), Hz=np.array(H_STEANE))
    return kw


def describe(code, circ, rounds) -> dict:
    text = str(circ)
    lc = {k: [[int(x) for x in row] for row in v] if v and hasattr(v[0], "__iter__") else [int(x) for x in v]
          for k, v in code.logic_check.items()}
    ent = {
        "name": code.name,
        "distance": int(code.distance),
        "num_physical_qubits": int(code.num_physical_qubits),
        "num_logical_qubits": int(code.num_logical_qubits),
        "rounds": int(rounds),
        "logic_check_ops": lc,
        "sha256": hashlib.sha256(text.encode()).hexdigest(),
        "num_lines": text.count("\n") + 1,
        "num_measurements": circ.num_measurements,
        "num_detectors": circ.num_detectors,
        "num_observables": circ.num_observables,
        "num_cx": circ.count("CX"),
        "num_depolarize1": circ.count("DEPOLARIZE1"),
        "num_depolarize2": circ.count("DEPOLARIZE2"),
        "raw_appends": getattr(circ, "raw_appends", None),
    }
    if len(text) <= FULL_TEXT_LIMIT:
        ent["text"] = text
    return ent


def _plain(o):
    """numpy scalars in the reference's metadata -> JSON numbers."""
    if hasattr(o, "item"):
        return o.item()
    raise TypeError(type(o))


def main():
    install_shims()
    import graphqec as ref  # the reference package itself
    from graphqec.lab.threshold.threshold_lab import ThresholdLAB as RefLab

    assert ref.__file__.startswith(REF), ref.__file__
    out = {"generator": "tests/golden/make_reference_circuits.py", "reference": "adelshb/graphqec @ /root/reference",
           "circuits": [], "tasks": []}
    for cid, cls, spec, rounds, lc, p in cases():
        kw = materialise(spec)
        code = getattr(ref, cls)(**kw, depolarize1_rate=p, depolarize2_rate=p)
        r = code.distance if rounds is None else rounds
        code.build_memory_circuit(number_of_rounds=r, logic_check=lc)
        ent = {"id": cid, "code": cls, "spec": spec, "logic_check": lc, "p": p, "rounds_arg": rounds}
        ent.update(describe(code, code.memory_circuit, r))
        out["circuits"].append(ent)
        print(f"{cid:18s} {ent['name']:34s} r={r:2d} lines={ent['num_lines']:6d} D={ent['num_detectors']:6d} "
              f"meas={ent['num_measurements']:6d} {ent['sha256'][:12]}")

    # ThresholdLAB.generate_sinter_tasks of the reference: metadata + circuit per (configuration, p)
    lab_cases = [
        ("lab_rep", "RepetitionCode", [{"distance": d} for d in (3, 5, 7, 11)], [0.0, 0.1, 0.2], "Z"),
        ("lab_rot", "RotatedSurfaceCode", [{"distance": d} for d in (5, 7)], [0.001, 0.01], "Z"),
        ("lab_rot_X", "RotatedSurfaceCode", [{"distance": 3}], [0.004], "X"),
        ("lab_steane", "CssCode", [{"steane": True, "name": "Steane"}], [0.0, 0.02], "Z"),
        ("lab_toric", "CssCode", [{"toric": 4, "name": "Toric"}], [0.006], "Z"),
        ("lab_bb", "BivariateBicycleCode", [dict(BB72)], [0.001], "Z"),
    ]
    for lid, cls, specs, rates, lc in lab_cases:
        lab = RefLab(code=getattr(ref, cls), configurations=[materialise(s) for s in specs], error_rates=rates,
                     decoder="pymatching")
        rows = []
        for task in lab.generate_sinter_tasks(logic_check=lc):
            text = str(task.circuit)
            rows.append({"json_metadata": task.json_metadata, "sha256": hashlib.sha256(text.encode()).hexdigest(),
                         "num_lines": text.count("\n") + 1})
        out["tasks"].append({"id": lid, "code": cls, "specs": specs, "error_rates": rates, "logic_check": lc,
                             "rows": rows})
        print(lid, [r["json_metadata"]["name"] for r in rows][:2], len(rows))
    path = os.path.join(HERE, "reference_circuits.json")
    json.dump(out, open(path, "w"), indent=1, default=_plain)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
