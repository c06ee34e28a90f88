"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by graphqec_b200).

Independent check of the circuit -> DEM analysis: instead of the
product's backward sensitivity sweep, every single Pauli fault of every noise channel is injected
and propagated *forwards* through the rest of the circuit as a Pauli frame (bit-packed, all faults
at once), recording which measurements -- hence detectors/observables -- it flips.  Identical
symptom sets are merged with p1(1-p2)+p2(1-p1).  PARITY UNPINNED against Stim itself (not
installed); pinned by SURVEY Appendix C and the notebook circuits.
"""

from __future__ import annotations

import numpy as np

from graphqec_b200.dem import depolarize1_component, depolarize2_component


def forward_fault_table(circuit) -> dict[tuple[tuple[int, ...], tuple[int, ...]], float]:
    ops = list(circuit)
    # ---- enumerate faults
    faults = []  # (op index, [(qubit, 'X'|'Z'|'Y'), ...], prob)
    for i, op in enumerate(ops):
        if op.name == "DEPOLARIZE1":
            q1 = depolarize1_component(op.gate_args[0])
            for q in op.targets:
                for pl in "XYZ":
                    faults.append((i, [(q, pl)], q1))
        elif op.name == "DEPOLARIZE2":
            q2 = depolarize2_component(op.gate_args[0])
            t = op.targets
            for k in range(0, len(t), 2):
                for pa in "IXYZ":
                    for pb in "IXYZ":
                        if pa == pb == "I":
                            continue
                        f = []
                        if pa != "I":
                            f.append((t[k], pa))
                        if pb != "I":
                            f.append((t[k + 1], pb))
                        faults.append((i, f, q2))
        elif op.name in ("X_ERROR", "Y_ERROR", "Z_ERROR"):
            for q in op.targets:
                faults.append((i, [(q, op.name[0])], op.gate_args[0]))
    F = len(faults)
    FW = (F + 63) // 64
    nq = circuit.num_qubits
    X = np.zeros((nq, FW), dtype=np.uint64)
    Z = np.zeros((nq, FW), dtype=np.uint64)
    nd, no = circuit.num_detectors, circuit.num_observables
    flips = np.zeros((circuit.num_measurements, FW), dtype=np.uint64)
    rows = np.zeros((nd + no, FW), dtype=np.uint64)
    by_op: dict[int, list[int]] = {}
    for f, (i, _, _) in enumerate(faults):
        by_op.setdefault(i, []).append(f)
    m = 0
    det = 0
    one = np.uint64(1)
    for i, op in enumerate(ops):
        name, t = op.name, op.targets
        if i in by_op:
            for f in by_op[i]:
                w, b = f >> 6, one << np.uint64(f & 63)
                for q, pl in faults[f][1]:
                    if pl in "XY":
                        X[q, w] ^= b
                    if pl in "ZY":
                        Z[q, w] ^= b
            continue
        if name in ("R", "RX"):
            X[t] = 0
            Z[t] = 0
        elif name == "H":
            tmp = X[t].copy()
            X[t] = Z[t]
            Z[t] = tmp
        elif name == "CX":
            for k in range(0, len(t), 2):
                c, x = t[k], t[k + 1]
                X[x] ^= X[c]
                Z[c] ^= Z[x]
        elif name in ("M", "MR"):
            for q in t:
                flips[m] = X[q]
                m += 1
                if name == "MR":
                    X[q] = 0
                    Z[q] = 0
        elif name in ("MX", "MRX"):
            for q in t:
                flips[m] = Z[q]
                m += 1
                if name == "MRX":
                    X[q] = 0
                    Z[q] = 0
        elif name == "DETECTOR":
            for r in t:
                rows[det] ^= flips[m + r.value]
            det += 1
        elif name == "OBSERVABLE_INCLUDE":
            k = nd + int(op.gate_args[0])
            for r in t:
                rows[k] ^= flips[m + r.value]
        elif name in ("TICK", "QUBIT_COORDS", "SHIFT_COORDS", "X", "Y", "Z", "I"):
            pass
        else:
            raise ValueError(f"forward oracle does not know gate {name}")
    bits = np.unpackbits(rows.view(np.uint8), axis=1, bitorder="little")[:, :F]
    rr, ff = np.nonzero(bits)
    order = np.argsort(ff, kind="stable")
    rr, ff = rr[order], ff[order]
    starts = np.searchsorted(ff, np.arange(F + 1))
    table: dict = {}
    for f in range(F):
        sym = rr[starts[f] : starts[f + 1]]
        if sym.size == 0:
            continue
        key = (tuple(int(x) for x in sym[sym < nd]), tuple(int(x) - nd for x in sym[sym >= nd]))
        p = faults[f][2]
        if p <= 0:
            continue
        old = table.get(key)
        table[key] = p if old is None else old * (1 - p) + p * (1 - old)
    return table
