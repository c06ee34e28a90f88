"""Generic CSS code from a pair of parity-check matrices (Hx, Hz).

Mirror of ``/root/reference/graphqec/codes/css_code.py:23-261``: logical operators come from
GF(2) algebra (``compute_logicals`` ``:234-261``), the CNOT schedule is the reference's greedy
serial one (``:155-231``: an edge gets step ``max(next_free[data], next_free[check])``, then the
data qubit is free at that step + 1 and the ancilla at that step + 2), and ``distance`` is the
minimum row weight of the computed logicals (``:104-113``) -- which ``ThresholdLAB`` also uses as
the number of rounds.
"""

from __future__ import annotations

import numpy as np

from graphqec_b200.codes.base_code import BaseCode
from graphqec_b200.math import commutation_test, compute_kernel, find_pivots

__all__ = ["CssCode", "compute_logicals"]


def compute_logicals(Hx: np.ndarray, Hz: np.ndarray, logical: str = "Z") -> np.ndarray:
    """Logical operators of one type: kernel of the opposite checks modulo the own checks.

    Rows of ``ker(H_other)`` that are picked as pivots *after* all rows of ``H_same`` in the
    stacked matrix are independent of the stabilizers, hence logical.
    """
    Hx = np.asarray(Hx)
    Hz = np.asarray(Hz)
    same, other = (Hz, Hx) if logical == "Z" else (Hx, Hz)
    stacked = np.vstack([same, compute_kernel(other)])
    chosen = [i for i in find_pivots(stacked) if i >= len(same)]
    return np.array([stacked[i, :] for i in chosen])


class CssCode(BaseCode):
    def __init__(self, Hx, Hz, name: str | None = None, *args, **kwargs) -> None:
        Hx = np.asarray(Hx)
        Hz = np.asarray(Hz)
        if not commutation_test(Hz, Hx):
            raise ValueError("The input parity check operators do not commute with each other.")
        self._Hx = Hx
        self._Hz = Hz
        self._num_data_qubits = len(Hz[0])
        self._L_z = compute_logicals(Hx=Hx, Hz=Hz, logical="Z")
        self._L_x = compute_logicals(Hx=Hx, Hz=Hz, logical="X")
        self._num_logical_qubits = len(self._L_z)
        self._logic_check = {
            "Z": [np.flatnonzero(row == 1).tolist() for row in self._L_z],
            "X": [np.flatnonzero(row == 1).tolist() for row in self._L_x],
        }
        super().__init__(*args, **kwargs)
        self._distance = self.distance_upper_bound()
        label = name if name is not None else "CSS"
        self._name = (
            f"{label} [[{self.num_physical_qubits},{self.num_logical_qubits},{self.distance}]]"
        )

    @property
    def Hx(self) -> np.ndarray:
        return self._Hx

    @property
    def Hz(self) -> np.ndarray:
        return self._Hz

    @property
    def L_z(self) -> np.ndarray:
        return self._L_z

    @property
    def L_x(self) -> np.ndarray:
        return self._L_x

    def distance_upper_bound(self) -> int:
        """Smallest weight among the computed logical operators (an upper bound on d)."""
        return int(min(min(int(r.sum()) for r in self._L_x), min(int(r.sum()) for r in self._L_z)))

    def build_graph(self) -> None:
        n = self._num_data_qubits
        nx_checks, nz_checks = len(self._Hx), len(self._Hz)
        g = self._graph
        for q in range(n):
            g.add_node(q, type="data", label=None, label_id=q)
        for q in range(n, n + nx_checks):
            g.add_node(q, type="check", label="X", label_id=q)
        for q in range(n + nx_checks, n + nx_checks + nz_checks):
            g.add_node(q, type="check", label="Z", label_id=q)

        next_free = [1] * (n + nx_checks + nz_checks)
        for base, matrix in ((n, self._Hx), (n + nx_checks, self._Hz)):
            for r, row in enumerate(matrix):
                anc = base + r
                for q in np.flatnonzero(row):
                    q = int(q)
                    step = max(next_free[q], next_free[anc])
                    g.add_edge(q, anc, weight=step)
                    next_free[q] = step + 1
                    next_free[anc] = step + 2
