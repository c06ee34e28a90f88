"""One CNOT of a stabilizer measurement (``/root/reference/graphqec/stab.py:18-29``).

X-type check: the ancilla controls (``CX check data``); Z-type check: the data qubit controls
(``CX data check``).  Direction is pinned by ``/root/reference/tests/test_stab.py:25,30``.
"""

from __future__ import annotations

__all__ = ["X_check", "Z_check"]


def X_check(circ, data_qubit, check_qubit) -> None:
    circ.append("CNOT", [check_qubit, data_qubit])


def Z_check(circ, data_qubit, check_qubit) -> None:
    circ.append("CNOT", [data_qubit, check_qubit])
