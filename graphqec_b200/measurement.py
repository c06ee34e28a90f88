"""Measurement-record bookkeeping: (round, qubit) -> position in the circuit's record.

Mirrors the public surface of ``/root/reference/graphqec/measurement.py:18-102`` (``Measurement``
with ``data``, ``register_count``, ``get_outcome``, ``get_register_id``, ``add_outcome``), including
its leniency: an unknown ``type`` string is accepted silently (reference ``:90-91`` builds the
``ValueError`` without raising it; ``/root/reference/tests/test_measurement.py:34`` relies on that).
"""

from __future__ import annotations

__all__ = ["Measurement"]


class Measurement:
    """Append-only log of measurement outcomes, indexed by round then qubit."""

    __slots__ = ("_data", "_register_count")

    def __init__(self) -> None:
        self._data: dict = {}
        self._register_count = 0

    @property
    def data(self) -> dict:
        return self._data

    @property
    def register_count(self) -> int:
        return self._register_count

    def _entry(self, qubit, round):
        return self._data.get(round, {}).get(qubit)

    def get_outcome(self, qubit, round):
        entry = self._entry(qubit, round)
        return None if entry is None else entry["outcome"]

    def get_register_id(self, qubit, round):
        entry = self._entry(qubit, round)
        return None if entry is None else entry["register_id"]

    def add_outcome(self, outcome, qubit, round: int, type: str | None) -> None:
        # the reference does not reject unknown types; neither do we
        slot = self._data.setdefault(round, {})
        slot[qubit] = {"outcome": outcome, "type": type, "register_id": self._register_count}
        self._register_count += 1
