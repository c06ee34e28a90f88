"""Bivariate bicycle (BB) codes of Bravyi et al. (arXiv:2308.07915).

Mirror of ``/root/reference/graphqec/codes/bivariate_bicycle_code.py:26-284``:
A = x^a1 + y^a2 + y^a3, B = y^b1 + x^b2 + x^b3 with x = S_L1 (x) I, y = I (x) S_L2 (``:132-156``),
Hx = [A | B], Hz = [B^T | A^T] (``:163-164``), the depth-7 CNOT order of the paper's Table 5
(``:76-93``), data qubits split into an "L" and an "R" block, logicals from ``compute_logicals``.
The name carries the *data* qubit count (``:99``), unlike the other codes.
"""

from __future__ import annotations

import numpy as np

from graphqec_b200.codes.base_code import BaseCode
from graphqec_b200.codes.css_code import compute_logicals

__all__ = ["BivariateBicycleCode"]


class BivariateBicycleCode(BaseCode):
    def __init__(
        self,
        L1: int = 12,
        L2: int = 6,
        a1: int = 3,
        a2: int = 1,
        a3: int = 2,
        b1: int = 3,
        b2: int = 1,
        b3: int = 2,
        *args,
        **kwargs,
    ) -> None:
        self._name = "Bivariate Bicycle"
        self._logic_check = []
        self.L1, self.L2 = L1, L2
        self.a1, self.a2, self.a3 = a1, a2, a3
        self.b1, self.b2, self.b3 = b1, b2, b3
        self.compute_check_matrices()
        # (monomial, data block) coupled at CNOT steps 1..7; None = idle step  (Table 5)
        self.sequence_X = [
            (None, None), (self.A2, "L"), (self.B2, "R"), (self.B1, "R"),
            (self.B3, "R"), (self.A1, "L"), (self.A3, "L"),
        ]
        self.sequence_Z = [
            (self.A1.T, "R"), (self.A3.T, "R"), (self.B1.T, "L"), (self.B2.T, "L"),
            (self.B3.T, "L"), (self.A2.T, "R"), (None, None),
        ]
        super().__init__(*args, **kwargs)
        self.get_parameters()
        self._name = (
            f"Bivariate Bicycle [[{self.num_data_qubits},{self.num_logical_qubits},{self.distance}]]"
        )
        self._logic_check = {
            "Z": [np.flatnonzero(row == 1).tolist() for row in self._L_z],
            "X": [np.flatnonzero(row == 1).tolist() for row in self._L_x],
        }

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

    def compute_check_matrices(self) -> None:
        eye1 = np.identity(self.L1, dtype=int)
        eye2 = np.identity(self.L2, dtype=int)
        self.x = {i: np.kron(np.roll(eye1, i, axis=1), eye2) for i in range(self.L1)}
        self.y = {i: np.kron(eye1, np.roll(eye2, i, axis=1)) for i in range(self.L2)}
        self.A1, self.A2, self.A3 = self.x[self.a1], self.y[self.a2], self.y[self.a3]
        self.B1, self.B2, self.B3 = self.y[self.b1], self.x[self.b2], self.x[self.b3]
        self.A = (self.A1 + self.A2 + self.A3) % 2
        self.B = (self.B1 + self.B2 + self.B3) % 2

    def get_parameters(self) -> None:
        """Fill Hx, Hz, the logicals and [[n, k, d]] (d = min logical weight found)."""
        self._Hx = np.hstack((self.A, self.B))
        self._Hz = np.hstack((self.B.T, self.A.T))
        self._L_z = compute_logicals(Hx=self._Hx, Hz=self._Hz, logical="Z")
        self._L_x = compute_logicals(Hx=self._Hx, Hz=self._Hz, logical="X")
        self._num_logical_qubits = len(self._L_z)
        self._num_data_qubits = len(self._Hz[0])
        self._distance = int(
            min(min(int(r.sum()) for r in self._L_x), min(int(r.sum()) for r in self._L_z))
        )

    def build_graph(self) -> None:
        half = self.L1 * self.L2
        g = self._graph
        block_base = {"L": 0, "R": half}
        for blk in ("L", "R"):
            for i in range(half):
                g.add_node(block_base[blk] + i, type="data", label=blk, label_id=i)
        for label, base, seq in (("X", 2 * half, self.sequence_X), ("Z", 3 * half, self.sequence_Z)):
            for i in range(half):
                g.add_node(base + i, type="check", label=label, label_id=i)
            for i in range(half):
                for step, (mono, blk) in enumerate(seq, start=1):
                    if mono is None:
                        continue
                    partner = int(np.flatnonzero(mono[i, :])[0])
                    g.add_edge(block_base[blk] + partner, base + i, weight=step)

    def get_neighbor_qubits(self, node) -> list:
        """Ordered data partners (``None`` for the idle step) of a check node ``(id, attrs)``."""
        attrs = node[1]
        seq = self.sequence_X if attrs["label"] == "X" else self.sequence_Z
        half = self.L1 * self.L2
        out = []
        for mono, blk in seq:
            if mono is None:
                out.append(None)
            else:
                partner = int(np.flatnonzero(mono[attrs["label_id"], :])[0])
                out.append(partner + (0 if blk == "L" else half))
        return out
