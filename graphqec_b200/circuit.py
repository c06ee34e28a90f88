"""Stim-free circuit IR for the memory experiments graphqec compiles.

The reference builds a ``stim.Circuit`` (``/root/reference/graphqec/codes/base_code.py:19,215-327``,
``/root/reference/graphqec/stab.py:22,29``).  Stim is not available on the B200 boxes, so this module
provides the small subset of ``stim.Circuit`` the reference touches -- ``append``,
``append_from_stim_program_text``, ``target_rec``, ``str()`` in Stim's text syntax (adjacent identical
instructions fused, ``CNOT`` canonicalised to ``CX`` as pinned by
``/root/reference/tests/test_stab.py:25,30``) -- plus the counters the DEM builder needs.

The circuit is the *input* of the hot path; the hot path itself starts at
``Circuit.detector_error_model`` (``graphqec_b200/dem.py``).
"""

from __future__ import annotations

import re
from dataclasses import dataclass
from typing import Iterable, Iterator, Sequence

__all__ = ["Circuit", "CircuitInstruction", "GateTarget", "target_rec"]


@dataclass(frozen=True)
class GateTarget:
    """A measurement-record target ``rec[-k]`` (the only non-qubit target kind graphqec emits)."""

    value: int  # negative look-back offset

    @property
    def is_measurement_record_target(self) -> bool:
        return True

    def __repr__(self) -> str:
        return f"stim.target_rec({self.value})"

    def __str__(self) -> str:
        return f"rec[{self.value}]"


def target_rec(lookback_index: int) -> GateTarget:
    """Mirror of ``stim.target_rec``: ``lookback_index`` must be negative."""
    lookback_index = int(lookback_index)
    if lookback_index >= 0:
        raise ValueError("rec targets look backwards: the offset must be negative")
    return GateTarget(lookback_index)


# canonical name -> (kind, takes_arg)
#   kind: "reset", "measure", "measure_reset", "clifford1", "clifford2", "noise1", "noise2",
#         "annotation"
_GATES = {
    "R": "reset",
    "RX": "reset",
    "M": "measure",
    "MX": "measure",
    "MR": "measure_reset",
    "MRX": "measure_reset",
    "H": "clifford1",
    "X": "clifford1",
    "Y": "clifford1",
    "Z": "clifford1",
    "I": "clifford1",
    "S": "clifford1",
    "S_DAG": "clifford1",
    "CX": "clifford2",
    "CZ": "clifford2",
    "SWAP": "clifford2",
    "DEPOLARIZE1": "noise1",
    "X_ERROR": "noise1",
    "Y_ERROR": "noise1",
    "Z_ERROR": "noise1",
    "DEPOLARIZE2": "noise2",
    "DETECTOR": "annotation",
    "OBSERVABLE_INCLUDE": "annotation",
    "TICK": "annotation",
    "QUBIT_COORDS": "annotation",
    "SHIFT_COORDS": "annotation",
}
_ALIASES = {"CNOT": "CX", "ZCX": "CX", "ZCZ": "CZ", "RZ": "R", "MZ": "M", "MRZ": "MR", "H_XZ": "H"}
_FUSABLE_KINDS = {"reset", "measure", "measure_reset", "clifford1", "clifford2", "noise1", "noise2"}


def _canonical(name: str) -> str:
    name = name.upper()
    name = _ALIASES.get(name, name)
    if name not in _GATES:
        raise ValueError(f"Gate not supported by the graphqec_b200 circuit IR: {name!r}")
    return name


def _fmt_arg(x: float) -> str:
    if float(x) == int(x) and abs(x) < 1e15:
        return str(int(x))
    return repr(float(x))


@dataclass
class CircuitInstruction:
    name: str
    targets: list
    gate_args: list

    @property
    def kind(self) -> str:
        return _GATES[self.name]

    def targets_copy(self) -> list:
        return list(self.targets)

    def gate_args_copy(self) -> list:
        return list(self.gate_args)

    def __str__(self) -> str:
        head = self.name
        if self.gate_args:
            head += "(" + ", ".join(_fmt_arg(a) for a in self.gate_args) + ")"
        if self.targets:
            head += " " + " ".join(str(t) for t in self.targets)
        return head


_LINE_RE = re.compile(r"^\s*([A-Za-z_][A-Za-z0-9_]*)\s*(?:\(([^)]*)\))?\s*(.*)$")
_REC_RE = re.compile(r"^rec\[(-\d+)\]$")


class Circuit:
    """A flat (loop-free) stabilizer circuit with noise channels and annotations."""

    def __init__(self, program_text: str | None = None) -> None:
        self._ops: list[CircuitInstruction] = []
        self._num_measurements = 0
        self._num_detectors = 0
        self._num_observables = 0
        self._num_qubits = 0
        if program_text:
            self.append_from_stim_program_text(program_text)

    # ------------------------------------------------------------------ building
    def append(self, name: str, targets: Iterable | int | GateTarget = (), arg=None) -> None:
        """Same call shape as ``stim.Circuit.append(name, targets, arg)``.

        An instruction with the same gate and argument as the previous one is fused into it, as Stim
        does (this is what makes ``R data`` + ``R checks`` print as a single line).
        """
        cname = _canonical(name)
        if isinstance(targets, (int, GateTarget)):
            targets = [targets]
        tlist = []
        for t in targets:
            if isinstance(t, GateTarget):
                tlist.append(t)
            else:
                ti = int(t)
                if ti < 0:
                    raise ValueError("qubit targets must be non-negative")
                tlist.append(ti)
        if arg is None:
            args = []
        elif isinstance(arg, (list, tuple)):
            args = [float(a) for a in arg]
        else:
            args = [float(arg)]
        kind = _GATES[cname]
        if kind in ("clifford2", "noise2") and len(tlist) % 2:
            raise ValueError(f"{cname} needs an even number of targets")
        if kind.startswith("noise") and args and not (0.0 <= args[0] <= 1.0):
            raise ValueError(f"{cname} probability out of range: {args[0]}")
        if kind not in ("annotation",):
            for t in tlist:
                if isinstance(t, GateTarget):
                    raise ValueError(f"{cname} does not take rec targets")
                self._num_qubits = max(self._num_qubits, t + 1)
        if kind in ("measure", "measure_reset"):
            self._num_measurements += len(tlist)
        if cname == "DETECTOR":
            self._check_recs(tlist)
            self._num_detectors += 1
        elif cname == "OBSERVABLE_INCLUDE":
            self._check_recs(tlist)
            if not args:
                raise ValueError("OBSERVABLE_INCLUDE needs an observable index argument")
            self._num_observables = max(self._num_observables, int(args[0]) + 1)
        if kind in _FUSABLE_KINDS and not tlist:
            # graphqec emits target-less DEPOLARIZE1 (base_code.py:379-380 with every data qubit
            # busy); it acts on nothing, so it is not recorded.
            return
        prev = self._ops[-1] if self._ops else None
        if (
            prev is not None
            and kind in _FUSABLE_KINDS
            and prev.name == cname
            and prev.gate_args == args
        ):
            prev.targets.extend(tlist)
        else:
            self._ops.append(CircuitInstruction(cname, tlist, args))

    append_operation = append

    def _check_recs(self, tlist: Sequence) -> None:
        for t in tlist:
            if not isinstance(t, GateTarget):
                raise ValueError("DETECTOR / OBSERVABLE_INCLUDE take rec[-k] targets only")
            if -t.value > self._num_measurements:
                raise ValueError(f"{t} refers to a measurement before the start of the circuit")

    def append_from_stim_program_text(self, text: str) -> None:
        for raw in text.splitlines():
            line = raw.split("#", 1)[0].strip()
            if not line:
                continue
            if line.startswith("REPEAT") or line in ("{", "}"):
                raise ValueError("REPEAT blocks are not supported by the graphqec_b200 circuit IR")
            m = _LINE_RE.match(line)
            if not m:
                raise ValueError(f"cannot parse circuit line: {raw!r}")
            name, argtxt, rest = m.group(1), m.group(2), m.group(3)
            args = [float(a) for a in argtxt.split(",")] if argtxt and argtxt.strip() else None
            targets = []
            for tok in rest.split():
                r = _REC_RE.match(tok)
                targets.append(GateTarget(int(r.group(1))) if r else int(tok))
            self.append(name, targets, args)

    # ------------------------------------------------------------------ inspection
    @property
    def num_qubits(self) -> int:
        return self._num_qubits

    @property
    def num_measurements(self) -> int:
        return self._num_measurements

    @property
    def num_detectors(self) -> int:
        return self._num_detectors

    @property
    def num_observables(self) -> int:
        return self._num_observables

    def __len__(self) -> int:
        return len(self._ops)

    def __iter__(self) -> Iterator[CircuitInstruction]:
        return iter(self._ops)

    def __getitem__(self, i):
        return self._ops[i]

    def __eq__(self, other) -> bool:
        return isinstance(other, Circuit) and str(self) == str(other)

    def __str__(self) -> str:
        return "\n".join(str(op) for op in self._ops)

    def __repr__(self) -> str:
        return f"graphqec_b200.Circuit('''\n{self}\n''')"

    def copy(self) -> "Circuit":
        c = Circuit()
        c._ops = [CircuitInstruction(o.name, list(o.targets), list(o.gate_args)) for o in self._ops]
        c._num_measurements = self._num_measurements
        c._num_detectors = self._num_detectors
        c._num_observables = self._num_observables
        c._num_qubits = self._num_qubits
        return c

    def count(self, name: str) -> int:
        """Number of *targets groups* of a gate (pairs for 2-qubit gates) -- e.g. noise sites."""
        cname = _canonical(name)
        step = 2 if _GATES[cname] in ("clifford2", "noise2") else 1
        return sum(len(o.targets) // step for o in self._ops if o.name == cname)

    def diagram(self, *_a, **_k) -> str:
        """Stim renders a timeline here; without Stim the program text is the diagram."""
        return str(self)

    # ------------------------------------------------------------------ hot-path entry points
    def detector_error_model(
        self,
        decompose_errors: bool = False,
        approximate_disjoint_errors: bool = True,
        **_ignored,
    ):
        """Circuit -> DEM, the restatement of Stim's error analysis.

        Built by the library's native builder (``gq_circuit_to_dem``) when ``libgqec_b200.so`` is present,
        else by the equivalent Python one (``dem.circuit_to_dem``); both are host-side set-up and give
        identical arrays (``tests/test_dem.py``).  ``GQEC_PY_DEM=1`` forces the Python builder."""
        import os

        from graphqec_b200 import backend
        from graphqec_b200.dem import circuit_to_dem, circuit_to_dem_native

        if os.environ.get("GQEC_PY_DEM") != "1" and os.path.exists(backend.library_path()):
            return circuit_to_dem_native(self, decompose_errors=decompose_errors)
        return circuit_to_dem(self, decompose_errors=decompose_errors)

    def to_stim(self):
        """Convert to a real ``stim.Circuit`` when Stim is importable (cross-check only)."""
        import stim  # noqa: F401  (optional dependency, absent on the B200 image)

        return stim.Circuit(str(self))
