"""GF(2) linear algebra used at set-up time by ``CssCode`` / ``BivariateBicycleCode``.

Public names follow ``/root/reference/graphqec/math.py:17-27``.  ``compute_kernel`` (``:46-67``) and
``find_pivots`` (``:70-87``) decide *which* logical operators a CSS code gets -- and therefore its
``distance`` upper bound, round count and observable -- so their elimination order is restated
exactly (column sweep in index order, pivot = first set row, every other column holding the pivot
row is cleared), but on boolean arrays with whole-matrix XOR updates instead of per-row Python loops.
"""

from __future__ import annotations

import numpy as np

__all__ = [
    "add_rows",
    "commutation_test",
    "compute_kernel",
    "find_pivots",
    "prep_matrix",
    "row_extended_check",
    "transpose_check",
    "find_intersection_point",
    "rankF2",
]


def add_rows(row1: np.ndarray, row2: np.ndarray) -> np.ndarray:
    """Sum of two equal-length binary vectors over GF(2)."""
    row1 = np.asarray(row1)
    row2 = np.asarray(row2)
    if row1.shape != row2.shape:
        raise AssertionError("rows must have the same shape")
    return (row1 + row2) % 2


def commutation_test(Hx: np.ndarray, Hz: np.ndarray) -> bool:
    """True when every X check commutes with every Z check (Hx Hz^T = 0 over GF(2))."""
    Hx = np.asarray(Hx)
    Hz = np.asarray(Hz)
    if Hx.shape[1] != Hz.shape[1]:
        raise AssertionError("check matrices act on different numbers of qubits")
    return bool(np.all((Hx @ Hz.T) % 2 == 0))


def row_extended_check(check_matrix: np.ndarray) -> np.ndarray:
    """The check matrix stacked on top of an identity of matching width."""
    check_matrix = np.asarray(check_matrix)
    return np.vstack([check_matrix, np.eye(check_matrix.shape[1], dtype=int)])


def transpose_check(check_matrix: np.ndarray) -> np.ndarray:
    return np.asarray(check_matrix).T.copy()


def prep_matrix(check_matrix: np.ndarray) -> np.ndarray:
    """``[H ; I]`` transposed: one row per column of H, carrying its identity tag."""
    return transpose_check(row_extended_check(check_matrix))


def _sweep(work: np.ndarray, width: int) -> list[int]:
    """Column-order elimination on a boolean matrix whose rows are the vectors being combined.

    For each row in index order whose first ``width`` entries are not all zero, take its first set
    position as pivot and XOR the row into every *other* row that has that position set.  Returns
    the pivot positions in the order they were chosen.
    """
    pivots = []
    for r in range(work.shape[0]):
        head = work[r, :width]
        if not head.any():
            continue
        mark = int(np.argmax(head))
        pivots.append(mark)
        hit = work[:, mark].copy()
        hit[r] = False
        work[hit] ^= work[r]
    return pivots


def compute_kernel(check_matrix: np.ndarray) -> np.ndarray:
    """Basis of the GF(2) null space of ``check_matrix`` (rows = kernel vectors)."""
    check_matrix = np.asarray(check_matrix)
    nrows = check_matrix.shape[0]
    work = prep_matrix(check_matrix).astype(bool)
    _sweep(work, nrows)
    dead = ~work[:, :nrows].any(axis=1)
    return work[dead][:, nrows:].astype(int)


def find_pivots(check_matrix: np.ndarray) -> list[int]:
    """Indices of rows of ``check_matrix`` forming a maximal independent set (sweep order)."""
    work = transpose_check(check_matrix).astype(bool)
    return _sweep(work, work.shape[1])


def rankF2(A: np.ndarray) -> int:
    """Rank of a binary matrix over GF(2)."""
    work = np.asarray(A).astype(bool).copy()
    rank = 0
    rows, cols = work.shape
    for c in range(cols):
        cand = np.nonzero(work[rank:, c])[0]
        if cand.size == 0:
            continue
        p = rank + int(cand[0])
        if p != rank:
            work[[rank, p]] = work[[p, rank]]
        hit = work[:, c].copy()
        hit[rank] = False
        work[hit] ^= work[rank]
        rank += 1
        if rank == rows:
            break
    return rank


def find_intersection_point(dict_lines) -> tuple[float, float]:
    """Least-squares common point of straight-line fits (one ``{x: y}`` dict per line)."""
    slopes, intercepts = [], []
    for pts in dict_lines:
        x = np.array(list(pts.keys()), dtype=float)
        y = np.array(list(pts.values()), dtype=float)
        m, b = np.polyfit(x, y, 1)
        slopes.append(m)
        intercepts.append(b)
    lhs = np.column_stack([np.array(slopes), -np.ones(len(slopes))])
    rhs = -np.array(intercepts)
    sol, *_ = np.linalg.lstsq(lhs, rhs, rcond=None)
    return tuple(sol)
