"""Repetition code: d data qubits in a line, d-1 Z checks between neighbours.

Mirror of ``/root/reference/graphqec/codes/repetition_code.py:20-70``: check i couples to data i at
step 1 and to data i+1 at step 2; only a Z memory is defined, read out on data qubit 0.
"""

from __future__ import annotations

from graphqec_b200.codes.base_code import BaseCode
from graphqec_b200.codes._lattice import wire_lattice

__all__ = ["RepetitionCode"]


class RepetitionCode(BaseCode):
    def __init__(self, distance: int, *args, **kwargs) -> None:
        self._distance = distance
        self._num_data_qubits = distance
        self._num_logical_qubits = 1
        self._name = "Repetition"
        self._logic_check = {"Z": [0]}
        super().__init__(*args, **kwargs)
        self._name = (
            f"Repetition [[{self.num_physical_qubits},{self.num_logical_qubits},{self.distance}]]"
        )

    def build_graph(self) -> None:
        d = self._distance
        wire_lattice(
            self._graph,
            [(i, i) for i in range(d)],
            [("Z", [(i + 0.5, i + 0.5) for i in range(d - 1)], [(-0.5, -0.5), (0.5, 0.5)])],
        )
