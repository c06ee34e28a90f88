"""Shared helper for the lattice codes: place qubits by coordinate, wire checks by offset order."""

from __future__ import annotations

import networkx as nx


def wire_lattice(
    graph: nx.Graph,
    data_coords: list[tuple],
    groups: list[tuple[str, list[tuple], list[tuple]]],
) -> None:
    """Fill ``graph`` with data qubits and check ancillas.

    ``groups`` holds, per check kind in insertion order, ``(label, ancilla coords, neighbour
    offsets)``; the k-th offset (1-based) that lands on a data qubit becomes an edge of weight k
    -- i.e. the CNOT of time step k.  Offsets that fall off the lattice still consume their step.
    Node ids: data first, then each group's ancillas, in the order given.
    """
    where = {}
    for i, xy in enumerate(data_coords):
        graph.add_node(i, type="data", label=None, coords=xy)
        where[xy] = i
    next_id = len(data_coords)
    for label, coords, offsets in groups:
        ids = range(next_id, next_id + len(coords))
        for q, xy in zip(ids, coords):
            graph.add_node(q, type="check", label=label, coords=xy)
        for q, (cx, cy) in zip(ids, coords):
            for step, (dx, dy) in enumerate(offsets, start=1):
                nb = where.get((cx + dx, cy + dy))
                if nb is not None:
                    graph.add_edge(nb, q, weight=step)
        next_id += len(coords)
