"""Rotated surface code on a d x d data lattice.

Mirror of ``/root/reference/graphqec/codes/rotated_surface_code.py:20-156``.  X ancillas sit on the
half-integer sites of rows 0.5 ... d+0.5, Z ancillas on rows 1.5 ... d-0.5; the CNOT orders
(reference ``:86`` X = [1,3,0,2], ``:115`` Z = [1,0,3,2] over the offsets
(-,-),(-,+),(+,-),(+,+)) are the hook-error-safe "N"/"Z" patterns.  Logical Z is read on the first
row of data qubits, logical X on the first column (reference ``:40-43``).  Even distances are not
rejected.
"""

from __future__ import annotations

from graphqec_b200.codes.base_code import BaseCode
from graphqec_b200.codes._lattice import wire_lattice

__all__ = ["RotatedSurfaceCode"]

_H = 0.5
_X_STEPS = [(-_H, +_H), (+_H, +_H), (-_H, -_H), (+_H, -_H)]
_Z_STEPS = [(-_H, +_H), (-_H, -_H), (+_H, +_H), (+_H, -_H)]


class RotatedSurfaceCode(BaseCode):
    def __init__(self, distance: int, *args, **kwargs) -> None:
        self._distance = distance
        self._num_data_qubits = distance**2
        self._num_logical_qubits = 1
        self._logic_check = {
            "Z": list(range(distance)),
            "X": [k * distance for k in range(distance)],
        }
        super().__init__(*args, **kwargs)
        self._name = (
            f"Rotated Surface [[{self.num_physical_qubits},{self.num_logical_qubits},{self.distance}]]"
        )

    def build_graph(self) -> None:
        d = self._distance
        data = [(col, row) for row in range(1, d + 1) for col in range(1, d + 1)]
        x_sites = [
            (col + (_H if row % 2 else -_H), row - _H)
            for row in range(1, d + 2)
            for col in range(2, d, 2)
        ]
        z_sites = [
            (col + (-_H if row % 2 else _H), row + _H)
            for row in range(1, d)
            for col in range(1, d + 1, 2)
        ]
        wire_lattice(self._graph, data, [("X", x_sites, _X_STEPS), ("Z", z_sites, _Z_STEPS)])

    def get_neighbor_qubits(self, coord, index_order=None) -> list:
        """Diagonal data neighbours of a site, in (-,-),(-,+),(+,-),(+,+) order or re-indexed."""
        where = {
            a["coords"]: n for n, a in self._graph.nodes(data=True) if a.get("type") == "data"
        }
        cx, cy = coord
        found = [where.get((cx + dx, cy + dy)) for dx in (-_H, _H) for dy in (-_H, _H)]
        return found if index_order is None else [found[i] for i in index_order]
