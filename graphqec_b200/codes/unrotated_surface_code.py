"""Unrotated (planar) surface code on a (2d-1) x (2d-1) grid.

Mirror of ``/root/reference/graphqec/codes/unrotated_surface_code.py:20-164``.  Data qubits occupy
the sites with even coordinate sum; the reference stores ancilla coordinates as (row, col) with X
ancillas on (odd, even) and Z ancillas on (even, odd) (``:66-69,99-102``) and looks neighbours up at
(0,-1),(-1,0),(+1,0),(0,+1) (``:144-149``) in order [0,1,2,3] for X and [0,2,1,3] for Z
(``:90,120``).  ``_num_data_qubits`` keeps the reference's formula (2d-1)^2 (``:38``).
"""

from __future__ import annotations

from graphqec_b200.codes.base_code import BaseCode
from graphqec_b200.codes._lattice import wire_lattice

__all__ = ["UnrotatedSurfaceCode"]

_CROSS = [(0, -1), (-1, 0), (1, 0), (0, 1)]


class UnrotatedSurfaceCode(BaseCode):
    def __init__(self, distance: int, *args, **kwargs) -> None:
        self._distance = distance
        self._num_data_qubits = (2 * distance - 1) ** 2
        self._num_logical_qubits = 1
        self._logic_check = {
            "Z": list(range(distance)),
            "X": [k * (2 * distance - 1) for k in range(distance)],
        }
        super().__init__(*args, **kwargs)
        self._name = (
            f"Unrotated Surface [[{self.num_physical_qubits},{self.num_logical_qubits},{self.distance}]]"
        )

    def build_graph(self) -> None:
        side = 2 * self._distance - 1
        grid = [(row, col) for row in range(side) for col in range(side)]
        data = [(col, row) for row, col in grid if (row + col) % 2 == 0]
        x_sites = [(row, col) for row, col in grid if row % 2 == 1 and col % 2 == 0]
        z_sites = [(row, col) for row, col in grid if row % 2 == 0 and col % 2 == 1]
        wire_lattice(
            self._graph,
            data,
            [
                ("X", x_sites, [_CROSS[i] for i in (0, 1, 2, 3)]),
                ("Z", z_sites, [_CROSS[i] for i in (0, 2, 1, 3)]),
            ],
        )

    def get_neighbor_qubits(self, coord, index_order=None) -> list:
        """Data neighbours of a site at (0,-1),(-1,0),(+1,0),(0,+1), optionally re-indexed."""
        where = {
            a["coords"]: n for n, a in self._graph.nodes(data=True) if a.get("type") == "data"
        }
        cx, cy = coord
        found = [where.get((cx + dx, cy + dy)) for dx, dy in _CROSS]
        return found if index_order is None else [found[i] for i in index_order]
