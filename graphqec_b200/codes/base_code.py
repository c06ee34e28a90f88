"""Tanner graph -> noisy memory-experiment circuit.

Host-side mirror of ``BaseCode`` in ``/root/reference/graphqec/codes/base_code.py`` (public
properties ``:76-165``, ``build_memory_circuit`` ``:184-327``, ``append_stab_circuit`` ``:329-397``,
``append_stab_element`` ``:399-425``, ``get_target_rec`` ``:455-469``).  The circuit defines the
detector error model the CUDA hot path samples and decodes, so the emission order and the noise
placement are reproduced exactly -- including the quirks listed in SURVEY.md section 8a (Q1, Q4, Q5,
Q6) -- but the implementation is organised differently: the CNOT layers are resolved once per graph
into a flat schedule instead of being re-derived by scanning the graph every round.
"""

from __future__ import annotations

from abc import ABC, abstractmethod

import networkx as nx

from graphqec_b200.circuit import Circuit, target_rec
from graphqec_b200.measurement import Measurement
from graphqec_b200.stab import X_check, Z_check

__all__ = ["BaseCode"]


class BaseCode(ABC):
    """A QEC code described by a Tanner graph whose edge weight is the CNOT time step."""

    def __init__(self, depolarize1_rate: float = 0, depolarize2_rate: float = 0) -> None:
        self._depolarize1_rate = depolarize1_rate
        self._depolarize2_rate = depolarize2_rate
        self._memory_circuit: Circuit | None = None
        self._measurement = Measurement()
        self._graph = nx.Graph()
        self.build_graph()
        self._num_physical_qubits = self._graph.number_of_nodes()
        self.get_check_label()
        self._length_sym_measurement = max(
            w for _, _, w in self._graph.edges(data="weight")
        )
        self._layers_cache = None

    # ------------------------------------------------------------------ read-only views
    @property
    def name(self) -> str:
        return self._name

    @property
    def num_physical_qubits(self) -> int:
        return self._num_physical_qubits

    @property
    def num_data_qubits(self) -> int:
        return self._num_data_qubits

    @property
    def num_logical_qubits(self) -> int:
        return self._num_logical_qubits

    @property
    def distance(self) -> int:
        return self._distance

    @property
    def memory_circuit(self) -> Circuit:
        return self._memory_circuit

    @property
    def depolarize1_rate(self) -> float:
        return self._depolarize1_rate

    @property
    def depolarize2_rate(self) -> float:
        return self._depolarize2_rate

    @property
    def measurement(self) -> Measurement:
        return self._measurement

    @property
    def register_count(self) -> int:
        return self._measurement.register_count

    @property
    def graph(self) -> nx.Graph:
        return self._graph

    @property
    def checks(self) -> list[str]:
        return self._checks

    @property
    def logic_check(self) -> dict:
        return self._logic_check

    # ------------------------------------------------------------------ graph-derived tables
    def get_check_label(self) -> None:
        """Sorted distinct check kinds present in the graph, e.g. ``["X-check", "Z-check"]``."""
        kinds = {
            f"{attrs['label']}-check"
            for _, attrs in self._graph.nodes(data=True)
            if attrs["type"] == "check"
        }
        self._checks = sorted(kinds)

    @abstractmethod
    def build_graph(self) -> None:
        """Populate ``self._graph`` (nodes: type/label/coords; edges: weight = CNOT step)."""

    def _qubit_groups(self) -> tuple[list, dict[str, list], list]:
        """(data qubits, {check kind: ancillas}, all ancillas) in graph insertion order."""
        data = [n for n, a in self._graph.nodes(data=True) if a.get("type") == "data"]
        groups: dict[str, list] = {kind: [] for kind in self._checks}
        for n, a in self._graph.nodes(data=True):
            key = f"{a.get('label')}-{a.get('type')}"
            if key in groups:
                groups[key].append(n)
        flat = [q for kind in groups for q in groups[kind]]
        return data, groups, flat

    def _cnot_layers(self, groups: dict[str, list]) -> list[list[tuple[str, int, int]]]:
        """``layers[s-1]`` = ordered ``(check kind, ancilla, data)`` CNOTs of time step ``s``.

        A check that has two edges with the same step is skipped for that step, as the reference
        does (its ``len(data) == 1`` test, ``base_code.py:368``).
        """
        if self._layers_cache is not None:
            return self._layers_cache
        per_check: dict[int, dict[int, list[int]]] = {}
        for kind in groups:
            for q in groups[kind]:
                slots: dict[int, list[int]] = {}
                for nb, attrs in self._graph[q].items():
                    slots.setdefault(attrs.get("weight"), []).append(nb)
                per_check[q] = slots
        layers = []
        for step in range(1, self._length_sym_measurement + 1):
            layer = []
            for kind in groups:
                for q in groups[kind]:
                    partners = per_check[q].get(step, ())
                    if len(partners) == 1:
                        layer.append((kind, q, partners[0]))
            layers.append(layer)
        self._layers_cache = layers
        return layers

    # ------------------------------------------------------------------ circuit emission
    def build_memory_circuit(self, number_of_rounds: int, logic_check: str = "Z") -> None:
        """Emit the memory experiment: ``number_of_rounds`` stabilizer rounds, then data readout."""
        if logic_check not in self.logic_check.keys():
            raise ValueError(f"Logic check must be one of {self.logic_check.keys()}")

        data, groups, ancillas = self._qubit_groups()
        everyone = list(self._graph.nodes())
        p1 = self._depolarize1_rate
        memory_kind = f"{logic_check}-check"
        circ = self._memory_circuit = Circuit()

        # preparation
        circ.append("R" if logic_check == "Z" else "RX", data)
        circ.append("R", ancillas)
        circ.append("DEPOLARIZE1", everyone, p1)

        # first round: only the memory-basis checks are deterministic
        self.append_stab_circuit(round=0, data_qubits=data, check_qubits=groups)
        for q in groups[memory_kind]:
            circ.append("DETECTOR", [target_rec(self.get_target_rec(qubit=q, round=0))])

        # bulk rounds: every check compared with its previous outcome
        for rnd in range(1, number_of_rounds):
            self.append_stab_circuit(round=rnd, data_qubits=data, check_qubits=groups)
            for q in ancillas:
                before = self.get_target_rec(qubit=q, round=rnd - 1)
                now = self.get_target_rec(qubit=q, round=rnd)
                circ.append("DETECTOR", [target_rec(before), target_rec(now)])

        # data readout in the memory basis
        circ.append("DEPOLARIZE1", data, p1)
        circ.append("M" if logic_check == "Z" else "MX", data)
        for i, q in enumerate(data):
            self.add_outcome(
                outcome=target_rec(-1 - i), qubit=q, round=number_of_rounds, type="data"
            )
        for q in groups[memory_kind]:
            recs = [
                self.get_target_rec(qubit=nb, round=number_of_rounds)
                for nb in self._graph.neighbors(q)
            ]
            recs.append(self.get_target_rec(qubit=q, round=number_of_rounds - 1))
            circ.append("DETECTOR", [target_rec(r) for r in recs])

        # logical observable(s): every logical operator lands in observable 0
        support = self.logic_check[logic_check]
        operators = support if isinstance(support[0], list) else [support]
        for qubits in operators:
            recs = [self.get_target_rec(qubit=q, round=number_of_rounds) for q in qubits]
            circ.append_from_stim_program_text(
                "OBSERVABLE_INCLUDE(0) " + " ".join(f"rec[{r}]" for r in recs)
            )

    def append_stab_circuit(
        self, round: int, data_qubits: list[int], check_qubits: dict[str, list[int]]
    ) -> None:
        """One stabilizer round: ancilla idle noise, H on X ancillas, CNOT layers, H, measure+reset."""
        circ = self._memory_circuit
        p1, p2 = self._depolarize1_rate, self._depolarize2_rate
        ancillas = [q for kind in check_qubits for q in check_qubits[kind]]
        x_ancillas = list(check_qubits["X-check"]) if "X-check" in self.checks else None

        if round > 0:
            circ.append("DEPOLARIZE1", ancillas, p1)
        if x_ancillas is not None:
            circ.append("H", x_ancillas)
            circ.append("DEPOLARIZE1", x_ancillas, p1)

        touched = set()
        for layer in self._cnot_layers(check_qubits):
            for kind, anc, dat in layer:
                self.append_stab_element(data_qubit=dat, check_qubit=anc, check=kind)
                circ.append("DEPOLARIZE2", [dat, anc], p2)
                touched.add(dat)

        # data qubits that sat out the whole round pick up idle noise
        circ.append("DEPOLARIZE1", [q for q in data_qubits if q not in touched], p1)

        if x_ancillas is not None:
            circ.append("H", x_ancillas)
            circ.append("DEPOLARIZE1", x_ancillas, p1)
        circ.append("DEPOLARIZE1", ancillas, p1)
        circ.append("MR", ancillas)
        for i, q in enumerate(ancillas):
            self.add_outcome(outcome=target_rec(-1 - i), qubit=q, round=round, type="check")

    def append_stab_element(self, data_qubit, check_qubit, check: str) -> None:
        """Emit the CNOT coupling one data qubit to one ancilla; unknown kinds are ignored."""
        if check == "Z-check":
            Z_check(circ=self._memory_circuit, data_qubit=data_qubit, check_qubit=check_qubit)
        elif check == "X-check":
            X_check(circ=self._memory_circuit, data_qubit=data_qubit, check_qubit=check_qubit)

    # ------------------------------------------------------------------ record bookkeeping
    def get_outcome(self, qubit, round: int):
        return self._measurement.get_outcome(qubit=qubit, round=round)

    def add_outcome(self, outcome, qubit, round: int, type: str | None) -> None:
        self._measurement.add_outcome(outcome=outcome, qubit=qubit, round=round, type=type)

    def get_target_rec(self, qubit, round: int) -> int | None:
        """Look-back offset (negative) of a recorded measurement, or ``None`` if never recorded."""
        rid = self._measurement.get_register_id(qubit=qubit, round=round)
        if rid is None:
            return None
        return rid - self._measurement.register_count

    # ------------------------------------------------------------------ drawing (host only)
    def draw_graph(self) -> None:
        """Plot the Tanner graph with matplotlib (optional dependency; not on the hot path)."""
        try:
            import matplotlib.pyplot as plt
            import matplotlib.patches as mpatches
        except ImportError as exc:  # matplotlib is absent on the B200 image
            raise RuntimeError("draw_graph needs matplotlib, which is not installed") from exc

        palette = {
            "data": "#D3D3D3",
            "L-data": "#FFCC99",
            "R-data": "#FFB6C1",
            "Z-check": "#d62728",
            "X-check": "#1f77b4",
            "Y-check": "#2ca02c",
        }
        category = {}
        for node, attrs in self._graph.nodes(data=True):
            lbl = attrs["label"]
            category[node] = attrs["type"] if lbl is None else f"{lbl}-{attrs['type']}"
        colours = [palette.get(category[n], "#808080") for n in self._graph.nodes()]
        try:
            layout = {int(n): tuple(a["coords"][:2]) for n, a in self._graph.nodes(data=True)}
        except KeyError:
            layout = nx.spring_layout(self._graph)

        plt.figure(figsize=(6, 6))
        nx.draw(
            self._graph, layout, with_labels=True, node_size=400, node_color=colours,
            font_size=8, font_weight="bold", edge_color="gray", width=1,
        )
        nx.draw_networkx_edge_labels(
            self._graph, layout, edge_labels=nx.get_edge_attributes(self._graph, "weight"),
            font_size=8, font_weight="bold",
        )
        handles = [
            mpatches.Patch(color=palette[c], label=f"{c} qubit") for c in sorted(set(category.values()))
        ]
        plt.legend(handles=handles, loc="upper left", bbox_to_anchor=(1, 1), title="")
        plt.title("")
        plt.show()
