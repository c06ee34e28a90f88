"""Detector error model (DEM): container, circuit -> DEM analysis, graphlike decomposition, text IO.

This is the host-side (setup-time) stage in front of the CUDA hot path.  It restates what sinter
asks Stim for when ``ThresholdLAB`` hands it a ``Task`` without a DEM
(``/root/reference/graphqec/lab/threshold/threshold_lab.py:151,182-197``):
``circuit.detector_error_model(decompose_errors=True, approximate_disjoint_errors=True)`` with a
fallback to the undecomposed model when decomposition fails.

Analysis method (restated, not Stim's code): walk the circuit backwards keeping, for every qubit,
the set of detectors/observables an X or a Z inserted at that point would flip; every Pauli
component of every noise channel then reads its symptom set off that table.  Sets are Python ints
used as bit masks (bit d = detector d, bit D+o = observable o).
"""

from __future__ import annotations

import itertools
import math
import re
from typing import Iterable, Sequence

import numpy as np

from graphqec_b200.circuit import Circuit, GateTarget

__all__ = ["DetectorErrorModel", "circuit_to_dem", "circuit_to_dem_native", "depolarize1_component", "depolarize2_component"]

NO_DET = 0xFFFFFFFF


def depolarize1_component(p: float) -> float:
    """Probability of each of the 3 independent Pauli components equivalent to DEPOLARIZE1(p)."""
    if p > 0.75:
        raise ValueError("DEPOLARIZE1 probability above 3/4 has no independent-component form")
    return 0.5 - 0.5 * math.sqrt(1.0 - 4.0 * p / 3.0)


def depolarize2_component(p: float) -> float:
    """Probability of each of the 15 independent Pauli components equivalent to DEPOLARIZE2(p)."""
    if p > 15.0 / 16.0:
        raise ValueError("DEPOLARIZE2 probability above 15/16 has no independent-component form")
    return 0.5 - 0.5 * (1.0 - 16.0 * p / 15.0) ** 0.125


def _xor_prob(p1: float, p2: float) -> float:
    return p1 * (1.0 - p2) + p2 * (1.0 - p1)


def _bits(mask: int) -> list[int]:
    out = []
    while mask:
        low = mask & -mask
        out.append(low.bit_length() - 1)
        mask ^= low
    return out


class DetectorErrorModel:
    """Independent error mechanisms ``(probability, detector set, observable set)`` in CSR form.

    ``components`` (optional) lists, per mechanism, its graphlike pieces as
    ``(det_a, det_b or NO_DET, observable mask)`` -- what Stim prints as ``D.. ^ D..`` -- and is what
    the matching decoder builds its graph from (PyMatching ignores pieces with more than two
    detectors).
    """

    def __init__(
        self,
        num_detectors: int,
        num_observables: int,
        probs: Sequence[float],
        dets: Sequence[Sequence[int]],
        obs: Sequence[Sequence[int]],
        components: Sequence[Sequence[tuple[int, int, int]]] | None = None,
    ) -> None:
        self.num_detectors = int(num_detectors)
        self.num_observables = int(num_observables)
        self.probs = np.asarray(probs, dtype=np.float64).reshape(-1)
        m = self.probs.shape[0]
        if len(dets) != m or len(obs) != m:
            raise ValueError("probs / dets / obs must have one entry per mechanism")
        self.det_indptr = np.zeros(m + 1, dtype=np.uint32)
        self.obs_indptr = np.zeros(m + 1, dtype=np.uint32)
        if m:
            self.det_indptr[1:] = np.cumsum([len(d) for d in dets])
            self.obs_indptr[1:] = np.cumsum([len(o) for o in obs])
        self.det_idx = np.fromiter(
            itertools.chain.from_iterable(dets), dtype=np.uint32, count=int(self.det_indptr[-1])
        )
        self.obs_idx = np.fromiter(
            itertools.chain.from_iterable(obs), dtype=np.uint32, count=int(self.obs_indptr[-1])
        )
        if self.det_idx.size and int(self.det_idx.max()) >= self.num_detectors:
            raise ValueError("detector index out of range")
        if self.obs_idx.size and int(self.obs_idx.max()) >= self.num_observables:
            raise ValueError("observable index out of range")
        if np.any(self.probs < 0) or np.any(self.probs > 1):
            raise ValueError("mechanism probability out of [0, 1]")
        self.decomposed = components is not None
        if components is None:
            components = [
                [self._self_component(i)] for i in range(m)
            ]
        self.comp_indptr = np.zeros(m + 1, dtype=np.uint32)
        if m:
            self.comp_indptr[1:] = np.cumsum([len(c) for c in components])
        flat = list(itertools.chain.from_iterable(components))
        self.comp_a = np.array([c[0] for c in flat], dtype=np.uint32)
        self.comp_b = np.array([c[1] for c in flat], dtype=np.uint32)
        self.comp_obs = np.array([c[2] for c in flat], dtype=np.uint64)

    @classmethod
    def from_arrays(cls, num_detectors, num_observables, probs, det_indptr, det_idx, obs_indptr, obs_idx,
                    comps=None) -> "DetectorErrorModel":
        """Adopt CSR arrays as they are (the native builder's output, ``gq_circuit_to_dem``)."""
        self = cls.__new__(cls)
        self.num_detectors = int(num_detectors)
        self.num_observables = int(num_observables)
        self.probs = np.ascontiguousarray(probs, dtype=np.float64)
        self.det_indptr = np.ascontiguousarray(det_indptr, dtype=np.uint32)
        self.det_idx = np.ascontiguousarray(det_idx, dtype=np.uint32)
        self.obs_indptr = np.ascontiguousarray(obs_indptr, dtype=np.uint32)
        self.obs_idx = np.ascontiguousarray(obs_idx, dtype=np.uint32)
        self.decomposed = comps is not None
        if comps is None:
            # every mechanism is its own piece when it is graphlike (<= 2 detectors), else none
            m = self.probs.shape[0]
            sizes = np.diff(self.det_indptr.astype(np.int64))
            first = self.det_idx[np.minimum(self.det_indptr[:-1], max(self.det_idx.size - 1, 0))] if self.det_idx.size else np.zeros(m, np.uint32)
            second = self.det_idx[np.minimum(self.det_indptr[:-1] + 1, max(self.det_idx.size - 1, 0))] if self.det_idx.size else np.zeros(m, np.uint32)
            a = np.where((sizes == 1) | (sizes == 2), first, NO_DET).astype(np.uint32)
            b = np.where(sizes == 2, second, NO_DET).astype(np.uint32)
            omask = np.zeros(m, np.uint64)
            rows = np.repeat(np.arange(m), np.diff(self.obs_indptr.astype(np.int64)))
            np.bitwise_xor.at(omask, rows, np.uint64(1) << self.obs_idx.astype(np.uint64))
            comps = (np.arange(m + 1, dtype=np.uint32), a, b, omask)
        self.comp_indptr, self.comp_a, self.comp_b, self.comp_obs = (
            np.ascontiguousarray(comps[0], np.uint32), np.ascontiguousarray(comps[1], np.uint32),
            np.ascontiguousarray(comps[2], np.uint32), np.ascontiguousarray(comps[3], np.uint64))
        return self

    # ------------------------------------------------------------------ helpers
    def _self_component(self, i: int) -> tuple[int, int, int]:
        d = self.det_idx[self.det_indptr[i] : self.det_indptr[i + 1]]
        omask = 0
        for o in self.obs_idx[self.obs_indptr[i] : self.obs_indptr[i + 1]]:
            omask |= 1 << int(o)
        if len(d) == 1:
            return (int(d[0]), NO_DET, omask)
        if len(d) == 2:
            return (int(d[0]), int(d[1]), omask)
        # 0 detectors or a hyperedge: not graphlike -> marked with both ends absent (ignored by
        # the matching-graph builder, sampled all the same).
        return (NO_DET, NO_DET, omask)

    @property
    def num_errors(self) -> int:
        return int(self.probs.shape[0])

    @property
    def nnz(self) -> int:
        return int(self.det_idx.shape[0])

    def mechanism(self, i: int) -> tuple[float, tuple[int, ...], tuple[int, ...]]:
        d = self.det_idx[self.det_indptr[i] : self.det_indptr[i + 1]]
        o = self.obs_idx[self.obs_indptr[i] : self.obs_indptr[i + 1]]
        return float(self.probs[i]), tuple(int(x) for x in d), tuple(int(x) for x in o)

    def mechanism_components(self, i: int) -> list[tuple[int, int, int]]:
        s, e = int(self.comp_indptr[i]), int(self.comp_indptr[i + 1])
        return [(int(self.comp_a[k]), int(self.comp_b[k]), int(self.comp_obs[k])) for k in range(s, e)]

    def graphlike_components(self) -> tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        """Flat list of graphlike pieces ``(a, b|NO_DET, obs mask, probability of the parent)``."""
        reps = np.diff(self.comp_indptr.astype(np.int64))
        p = np.repeat(self.probs, reps)
        keep = self.comp_a != NO_DET
        return self.comp_a[keep], self.comp_b[keep], self.comp_obs[keep], p[keep]

    def dense_check_matrices(self) -> tuple[np.ndarray, np.ndarray]:
        """Dense uint8 H (D x M) and L (O x M); only for small models / tests."""
        h = np.zeros((self.num_detectors, self.num_errors), dtype=np.uint8)
        l = np.zeros((self.num_observables, self.num_errors), dtype=np.uint8)
        for i in range(self.num_errors):
            h[self.det_idx[self.det_indptr[i] : self.det_indptr[i + 1]], i] = 1
            l[self.obs_idx[self.obs_indptr[i] : self.obs_indptr[i + 1]], i] = 1
        return h, l

    def summary(self) -> dict:
        sizes = np.diff(self.det_indptr.astype(np.int64))
        hist = {int(k): int(v) for k, v in zip(*np.unique(sizes, return_counts=True))}
        return {
            "D": self.num_detectors,
            "O": self.num_observables,
            "M": self.num_errors,
            "nnz_H": self.nnz,
            "nnz_L": int(self.obs_idx.shape[0]),
            "sum_p": float(self.probs.sum()),
            "min_p": float(self.probs.min()) if self.num_errors else 0.0,
            "max_p": float(self.probs.max()) if self.num_errors else 0.0,
            "distinct_p": int(np.unique(self.probs).shape[0]),
            "dets_per_mechanism": hist,
        }

    # ------------------------------------------------------------------ text IO (Stim .dem syntax)
    def __str__(self) -> str:
        lines = []
        for i in range(self.num_errors):
            p, d, o = self.mechanism(i)
            if self.decomposed:
                parts = []
                for a, b, om in self.mechanism_components(i):
                    toks = [f"D{x}" for x in (a, b) if x != NO_DET] + [f"L{x}" for x in _bits(om)]
                    parts.append(" ".join(toks))
                if parts == [""] or any(a == NO_DET and b == NO_DET for a, b, _ in self.mechanism_components(i)):
                    parts = [" ".join([f"D{x}" for x in d] + [f"L{x}" for x in o])]
                body = " ^ ".join(parts)
            else:
                body = " ".join([f"D{x}" for x in d] + [f"L{x}" for x in o])
            lines.append(f"error({p!r}) {body}".rstrip())
        if self.num_detectors:
            covered = int(self.det_idx.max()) + 1 if self.det_idx.size else 0
            if covered < self.num_detectors:
                lines.append(f"detector D{self.num_detectors - 1}")
        if self.num_observables:
            covered = int(self.obs_idx.max()) + 1 if self.obs_idx.size else 0
            if covered < self.num_observables:
                lines.append(f"logical_observable L{self.num_observables - 1}")
        return "\n".join(lines)

    @staticmethod
    def from_text(text: str) -> "DetectorErrorModel":
        """Parse Stim's .dem syntax: error / detector / logical_observable / shift_detectors / repeat."""
        lines = _flatten_dem_lines(text)
        probs, dets, obs, comps = [], [], [], []
        any_sep = False
        nd = no = 0
        shift = 0
        for kind, arg, toks in lines:
            if kind == "shift_detectors":
                shift += int(toks[0]) if toks else 0
                continue
            if kind == "detector":
                for t in toks:
                    if t.startswith("D"):
                        nd = max(nd, int(t[1:]) + shift + 1)
                continue
            if kind == "logical_observable":
                for t in toks:
                    if t.startswith("L"):
                        no = max(no, int(t[1:]) + 1)
                continue
            if kind != "error":
                raise ValueError(f"unsupported DEM instruction: {kind}")
            dset: set[int] = set()
            oset: set[int] = set()
            pieces = []
            cur_d: list[int] = []
            cur_o = 0
            for t in toks + ["^"]:
                if t == "^":
                    a = cur_d[0] if len(cur_d) >= 1 else NO_DET
                    b = cur_d[1] if len(cur_d) == 2 else NO_DET
                    if len(cur_d) > 2:
                        a = b = NO_DET
                    pieces.append((a, b, cur_o))
                    cur_d, cur_o = [], 0
                    continue
                if t.startswith("D"):
                    k = int(t[1:]) + shift
                    nd = max(nd, k + 1)
                    dset ^= {k}
                    cur_d.append(k)
                elif t.startswith("L"):
                    k = int(t[1:])
                    no = max(no, k + 1)
                    oset ^= {k}
                    cur_o ^= 1 << k
                else:
                    raise ValueError(f"bad DEM target {t!r}")
            if len(pieces) > 1:
                any_sep = True
            probs.append(float(arg))
            dets.append(sorted(dset))
            obs.append(sorted(oset))
            comps.append(pieces)
        return DetectorErrorModel(nd, no, probs, dets, obs, comps if any_sep else None)


_DEM_LINE = re.compile(r"^([a-z_]+)\s*(?:\(([^)]*)\))?\s*(.*)$")


def _flatten_dem_lines(text: str) -> list[tuple[str, str | None, list[str]]]:
    """Expand ``repeat N { ... }`` blocks (detector shifts are applied by the caller in order)."""
    raw = [ln.split("#", 1)[0].strip() for ln in text.splitlines()]
    raw = [ln for ln in raw if ln]

    def parse(block: list[str], pos: int) -> tuple[list, int]:
        out = []
        while pos < len(block):
            ln = block[pos]
            if ln == "}":
                return out, pos + 1
            if ln.startswith("repeat"):
                count = int(ln.split()[1])
                body, pos = parse(block, pos + 1)
                for _ in range(count):
                    out.extend(body)
                continue
            m = _DEM_LINE.match(ln)
            if not m:
                raise ValueError(f"cannot parse DEM line {ln!r}")
            out.append((m.group(1), m.group(2), m.group(3).split()))
            pos += 1
        return out, pos

    return parse(raw, 0)[0]


# --------------------------------------------------------------------------------------------
# circuit -> DEM
# --------------------------------------------------------------------------------------------

def _measurement_masks(circuit: Circuit) -> list[int]:
    """mask[m] = detectors/observables whose record list holds measurement m an odd number of times."""
    nmeas = circuit.num_measurements
    nd = circuit.num_detectors
    masks = [0] * nmeas
    seen = 0
    det = 0
    for op in circuit:
        kind = op.kind
        if kind in ("measure", "measure_reset"):
            seen += len(op.targets)
        elif op.name == "DETECTOR":
            for t in op.targets:
                masks[seen + t.value] ^= 1 << det
            det += 1
        elif op.name == "OBSERVABLE_INCLUDE":
            bit = 1 << (nd + int(op.gate_args[0]))
            for t in op.targets:
                masks[seen + t.value] ^= bit
    return masks


def _sweep_channels(circuit: Circuit, on_channel) -> None:
    """Backward sensitivity sweep; calls ``on_channel(basis, p)`` for every noise channel instance with the
    symptom masks of its basis Paulis -- (X, Z) for DEPOLARIZE1, (X_a, Z_a, X_b, Z_b) for DEPOLARIZE2, one mask for
    X/Y/Z_ERROR and measurement flips -- and the probability of each of its independent components."""
    masks = _measurement_masks(circuit)
    nq = circuit.num_qubits
    xs = [0] * nq  # symptom of an X inserted here
    zs = [0] * nq  # symptom of a Z inserted here
    m = circuit.num_measurements
    for op in reversed(circuit._ops):
        name, kind, t = op.name, op.kind, op.targets
        if kind == "annotation":
            continue
        if kind == "noise1":
            p = op.gate_args[0] if op.gate_args else 0.0
            if name == "DEPOLARIZE1":
                q1 = depolarize1_component(p)
                for q in t:
                    on_channel((xs[q], zs[q]), q1)
            elif name == "X_ERROR":
                for q in t:
                    on_channel((xs[q],), p)
            elif name == "Z_ERROR":
                for q in t:
                    on_channel((zs[q],), p)
            else:  # Y_ERROR
                for q in t:
                    on_channel((xs[q] ^ zs[q],), p)
        elif kind == "noise2":
            q2 = depolarize2_component(op.gate_args[0] if op.gate_args else 0.0)
            for k in range(0, len(t), 2):
                a, b = t[k], t[k + 1]
                on_channel((xs[a], zs[a], xs[b], zs[b]), q2)
        elif kind == "measure":
            flip = op.gate_args[0] if op.gate_args else 0.0
            for q in reversed(t):
                m -= 1
                if name == "M":
                    xs[q] ^= masks[m]
                else:
                    zs[q] ^= masks[m]
                on_channel((masks[m],), flip)
        elif kind == "measure_reset":
            flip = op.gate_args[0] if op.gate_args else 0.0
            for q in reversed(t):
                m -= 1
                if name == "MR":
                    xs[q], zs[q] = masks[m], 0
                else:
                    xs[q], zs[q] = 0, masks[m]
                on_channel((masks[m],), flip)
        elif kind == "reset":
            for q in t:
                xs[q] = 0
                zs[q] = 0
        elif name == "H":
            for q in t:
                xs[q], zs[q] = zs[q], xs[q]
        elif name in ("S", "S_DAG"):
            for q in t:
                xs[q] ^= zs[q]
        elif name == "CX":
            for k in range(len(t) - 2, -1, -2):
                c, x = t[k], t[k + 1]
                xs[c] ^= xs[x]
                zs[x] ^= zs[c]
        elif name == "CZ":
            for k in range(len(t) - 2, -1, -2):
                a, b = t[k], t[k + 1]
                xs[a] ^= zs[b]
                xs[b] ^= zs[a]
        elif name == "SWAP":
            for k in range(len(t) - 2, -1, -2):
                a, b = t[k], t[k + 1]
                xs[a], xs[b] = xs[b], xs[a]
                zs[a], zs[b] = zs[b], zs[a]
        # X, Y, Z, I: Pauli gates do not move sensitivities
    if m != 0:
        raise AssertionError("measurement bookkeeping out of sync")


def _combos(basis: tuple[int, ...]) -> list[int]:
    """Symptoms of all 2^s - 1 Pauli products of a channel: index k = subset of the basis (bit b = basis[b])."""
    sym = [0] * (1 << len(basis))
    for k in range(1, len(sym)):
        low = k & -k
        sym[k] = sym[k ^ low] ^ basis[low.bit_length() - 1]
    return sym


def analyze_circuit_errors(circuit: Circuit) -> dict[int, float]:
    """Symptom mask -> probability for every independent error mechanism of the circuit."""
    out: dict[int, float] = {}

    def on_channel(basis, p: float) -> None:
        if p <= 0.0:
            return
        for mask in _combos(basis)[1:]:
            if mask:
                old = out.get(mask)
                out[mask] = p if old is None else _xor_prob(old, p)

    _sweep_channels(circuit, on_channel)
    return out


# ---------------------------------------------------------------------------------------------------
# decompose_errors=True, restated after Stim's ErrorAnalyzer (UPSTREAM, stim >= 1.10): two stages.
#
# Stage 1, per noise channel instance ("in-channel"): a composite Pauli of the channel that flips more than two
# detectors is rewritten as a product of OTHER Paulis of the SAME channel -- its single-detector errors and its
# "irreducible pairs" (two-detector errors not covered by the single-detector ones): singles alone when they
# suffice, else one pair (+ singles), else two disjoint pairs (+ singles).  What cannot be written that way
# (e.g. a basis Pauli that is itself a hyperedge: hook errors) stays whole for stage 2.  Errors are keyed by
# their decomposition, as in Stim: the same symptom reached through two different decompositions is two errors.
#
# Stage 2, global: every remaining component with more than two detectors is matched greedily against the
# graphlike components known anywhere in the model -- for each detector in ascending order the first later
# detector that forms a known pair with it, then known single detectors; up to two unmatched detectors may stay
# as a "remnant edge".  More than two unmatched detectors is a decomposition failure (ValueError; sinter then
# falls back to the undecomposed model).  So is -- INFERRED, not recalled from the source -- a decomposition whose
# known pieces reproduce all detectors but leave an observable over: the reference's own Steane sweep
# (notebooks/steane.ipynb, LER 2.9e-2 at p = 2.2e-3) is reproduced by the undecomposed model (2.9e-2) and not by
# any decomposed one (2.0 .. 2.2e-2), while its toric sweeps (notebooks/toric_code.ipynb) are reproduced by this
# decomposition to 0.3 % and neither by the undecomposed model nor by a fewest-pieces search (-1.3 %);
# evidence: profiles/r02_reference_stats.log, DESIGN.md section 2.
# ---------------------------------------------------------------------------------------------------
_MAX_CHANNEL_DETECTORS = 15


def _in_channel_keys(sym: list[int], dmask: int) -> list[tuple[int, ...]]:
    """Stage 1: key (tuple of component masks) of every Pauli product k >= 1 of one channel."""
    n = len(sym)
    keys: list[tuple[int, ...]] = [(x,) for x in sym]
    dm = [x & dmask for x in sym]
    cnt = [bin(x).count("1") for x in dm]
    if max(cnt) <= 2:
        return keys
    union = 0
    for x in dm:
        union |= x
    if bin(union).count("1") > _MAX_CHANNEL_DETECTORS:
        raise ValueError("An error involves too many detectors (>15) to find reducible errors.")
    singles = 0
    for k in range(1, n):
        if cnt[k] == 1:
            singles |= dm[k]
    irreducible = [k for k in range(1, n) if cnt[k] == 2 and (dm[k] & ~singles)]
    for g in range(1, n):
        if cnt[g] <= 2:
            continue
        goal = dm[g]
        parts: list[int] = []
        rem = goal
        if goal & ~singles:
            found = False
            for k in irreducible:
                mk = dm[k]
                if (goal & mk) == mk and (goal & ~(singles | mk)) == 0:
                    parts, rem, found = [k], goal & ~mk, True
                    break
            if not found:
                for i1 in range(len(irreducible)):
                    m1 = dm[irreducible[i1]]
                    if (goal & m1) != m1:
                        continue
                    for i2 in range(i1 + 1, len(irreducible)):
                        m2 = dm[irreducible[i2]]
                        if (m1 & m2) == 0 and (goal & m2) == m2 and (goal & ~(singles | m1 | m2)) == 0:
                            parts, rem, found = [irreducible[i1], irreducible[i2]], goal & ~(m1 | m2), True
                            break
                    if found:
                        break
            if not found:
                continue
        for k in range(1, n):
            if rem == 0:
                break
            if cnt[k] == 1 and (dm[k] & ~rem) == 0:
                rem &= ~dm[k]
                parts.append(k)
        if rem:
            continue
        x = 0
        for k in parts:
            x ^= sym[k]
        if x == sym[g]:          # the pieces must reproduce the observables too, else stage 2 deals with it
            keys[g] = tuple(sym[k] for k in parts)
    return keys


def _stim_order(key: tuple[int, ...], nd: int) -> tuple[int, ...]:
    """Stim's ordering of error keys: token sequences compared lexicographically, D < L < separator."""
    dmask = (1 << nd) - 1
    toks: list[int] = []
    for i, c in enumerate(key):
        if i:
            toks.append(1 << 62)
        toks.extend(_bits(c & dmask))
        toks.extend((1 << 40) + o for o in _bits(c >> nd))
    return tuple(toks)


def _global_decomposition_pass(table: dict[tuple[int, ...], float], nd: int) -> dict[tuple[int, ...], float]:
    """Stage 2 over all errors in Stim order; merges errors that end up with the same decomposition."""
    dmask = (1 << nd) - 1
    items = sorted(table.items(), key=lambda kv: _stim_order(kv[0], nd))
    known: dict[tuple[int, ...], int] = {}
    for key, _ in items:
        for c in key:
            d = tuple(_bits(c & dmask))
            if 1 <= len(d) <= 2:
                known.setdefault(d, c)
    out: dict[tuple[int, ...], float] = {}
    for key, p in items:
        new: list[int] = []
        for c in key:
            d = _bits(c & dmask)
            if len(d) <= 2:
                new.append(c)
                continue
            done = [False] * len(d)
            left = c
            for i in range(len(d)):
                if done[i]:
                    continue
                for j in range(i + 1, len(d)):
                    if not done[j]:
                        kc = known.get((d[i], d[j]))
                        if kc is not None:
                            done[i] = done[j] = True
                            new.append(kc)
                            left ^= kc
                            break
            missed = 0
            for i in range(len(d)):
                if not done[i]:
                    kc = known.get((d[i],))
                    if kc is not None:
                        done[i] = True
                        new.append(kc)
                        left ^= kc
                    else:
                        missed += 1
            if missed > 2 or (left and not (left & dmask)):
                raise ValueError(
                    "Failed to decompose errors into graphlike components with at most two symptoms. "
                    "The error component that failed to decompose is '"
                    + ", ".join([f"D{x}" for x in d] + [f"L{x}" for x in _bits(c >> nd)]) + "'."
                )
            if left:
                new.append(left)       # remnant edge: the unmatched detectors (+ whatever observables are left over)
        nk = tuple(new)
        old = out.get(nk)
        out[nk] = p if old is None else _xor_prob(old, p)
    return out


def analyze_circuit_errors_decomposed(circuit: Circuit) -> dict[tuple[int, ...], float]:
    """Decomposition (tuple of graphlike component masks) -> probability; raises ValueError when Stim would."""
    nd = circuit.num_detectors
    dmask = (1 << nd) - 1
    table: dict[tuple[int, ...], float] = {}

    def on_channel(basis, p: float) -> None:
        if p <= 0.0:
            return
        sym = _combos(basis)
        keys = _in_channel_keys(sym, dmask)
        for k in range(1, len(sym)):
            if sym[k]:
                old = table.get(keys[k])
                table[keys[k]] = p if old is None else _xor_prob(old, p)

    _sweep_channels(circuit, on_channel)
    return _global_decomposition_pass(table, nd)


def circuit_to_dem_native(circuit: Circuit, decompose_errors: bool = False) -> DetectorErrorModel:
    """The same model as ``circuit_to_dem`` from the library's native builder (``gq_circuit_to_dem``,
    csrc/circuit_dem.cu: host C++, ~50x faster on a d = 11 memory circuit); bit-identical arrays."""
    from graphqec_b200 import backend

    return DetectorErrorModel.from_arrays(*backend.circuit_to_dem_arrays(str(circuit), decompose_errors))


def circuit_to_dem(circuit: Circuit, decompose_errors: bool = False) -> DetectorErrorModel:
    """Restatement of ``stim.Circuit.detector_error_model`` for graphqec's gate set.

    With ``decompose_errors=True`` a mechanism flipping more than two detectors is rewritten as graphlike
    pieces the way Stim does it (in-channel stage, then a greedy global pass; see above) and every distinct
    decomposition is its own error; ``ValueError`` is raised when Stim's decomposition would fail (sinter then
    falls back to the undecomposed model -- callers mirror that).
    """
    nd, no = circuit.num_detectors, circuit.num_observables
    dmask = (1 << nd) - 1
    if not decompose_errors:
        rows = [(tuple(_bits(mask & dmask)), mask >> nd, p) for mask, p in analyze_circuit_errors(circuit).items()]
        rows.sort(key=lambda r: (r[0], r[1]))
        return DetectorErrorModel(nd, no, [r[2] for r in rows], [r[0] for r in rows], [_bits(r[1]) for r in rows], None)
    table = analyze_circuit_errors_decomposed(circuit)
    probs, dets, obs, comps = [], [], [], []
    for key, p in sorted(table.items(), key=lambda kv: _stim_order(kv[0], nd)):
        whole = 0
        pieces = []
        for c in key:
            whole ^= c
            d = _bits(c & dmask)
            pieces.append((d[0] if d else NO_DET, d[1] if len(d) == 2 else NO_DET, c >> nd))
        probs.append(p)
        dets.append(tuple(_bits(whole & dmask)))
        obs.append(_bits(whole >> nd))
        comps.append(pieces)
    return DetectorErrorModel(nd, no, probs, dets, obs, comps)
