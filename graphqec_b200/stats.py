"""Task / TaskStats records and the Wilson score interval.

``sinter.Task`` / ``sinter.TaskStats`` are what the reference passes to and receives from
``sinter.collect`` (``/root/reference/graphqec/lab/threshold/threshold_lab.py:151,182-197``); its
consumers read ``.json_metadata``, ``.shots``, ``.errors``, ``.discards``, ``.seconds``
(``threshold_lab.py:220-246``, ``/root/reference/examples/rot_vs_unrot_surface_code.py:216-222``).
sinter is not installed on the B200 image, so equivalent plain dataclasses live here.
"""

from __future__ import annotations

import hashlib
import json
import math
from dataclasses import dataclass, field

__all__ = ["Task", "TaskStats", "AnonTaskStats", "wilson_interval", "CSV_HEADER",
           "read_stats_from_csv", "append_stats_to_csv"]

CSV_HEADER = "     shots,    errors,  discards, seconds,decoder,strong_id,json_metadata,custom_counts"


@dataclass
class Task:
    """One (circuit, metadata) point of a sweep."""

    circuit: object
    json_metadata: dict = field(default_factory=dict)
    decoder: str | None = None
    detector_error_model: object | None = None

    def strong_id(self, decoder: str | None = None) -> str:
        blob = json.dumps(
            {
                "circuit": str(self.circuit),
                "decoder": decoder or self.decoder,
                "json_metadata": self.json_metadata,
            },
            sort_keys=True,
            default=float,
        )
        return hashlib.sha256(blob.encode()).hexdigest()


@dataclass
class AnonTaskStats:
    shots: int = 0
    errors: int = 0
    discards: int = 0
    seconds: float = 0.0

    def __add__(self, other: "AnonTaskStats") -> "AnonTaskStats":
        return AnonTaskStats(
            self.shots + other.shots,
            self.errors + other.errors,
            self.discards + other.discards,
            self.seconds + other.seconds,
        )


@dataclass
class TaskStats:
    strong_id: str
    decoder: str
    json_metadata: dict
    shots: int = 0
    errors: int = 0
    discards: int = 0
    seconds: float = 0.0
    custom_counts: dict = field(default_factory=dict)   # sinter's last CSV column; "next_shot" = RNG stream position

    def to_anon_stats(self) -> AnonTaskStats:
        return AnonTaskStats(self.shots, self.errors, self.discards, self.seconds)

    def to_csv_line(self) -> str:
        """One row in sinter's CSV schema (see ``CSV_HEADER``)."""
        meta = json.dumps(self.json_metadata, sort_keys=True, default=float).replace('"', '""')
        custom = ""
        if self.custom_counts:
            custom = '"' + json.dumps(self.custom_counts, sort_keys=True).replace('"', '""') + '"'
        return (
            f"{self.shots:>10},{self.errors:>10},{self.discards:>10},{self.seconds:>8.3g},"
            f"{self.decoder},{self.strong_id},\"{meta}\",{custom}"
        )

    @property
    def error_rate(self) -> float:
        kept = self.shots - self.discards
        return self.errors / kept if kept else float("nan")


def wilson_interval(errors: int, shots: int, z: float = 1.959963984540054) -> tuple[float, float]:
    """Wilson score interval for a binomial proportion (default 95 %)."""
    if shots <= 0:
        return (0.0, 1.0)
    phat = errors / shots
    denom = 1.0 + z * z / shots
    centre = (phat + z * z / (2 * shots)) / denom
    half = z * math.sqrt(phat * (1 - phat) / shots + z * z / (4.0 * shots * shots)) / denom
    return (max(0.0, centre - half), min(1.0, centre + half))


def read_stats_from_csv(path: str) -> list[TaskStats]:
    """Rows of a CSV in sinter's schema, rows with one ``strong_id`` summed (what
    ``sinter.read_stats_from_csv_files`` returns and what ``save_resume_filepath`` resumes from)."""
    import csv
    import os

    merged: dict[str, TaskStats] = {}
    if not os.path.exists(path):
        return []
    with open(path, newline="") as f:
        rows = list(csv.reader(f, skipinitialspace=True))
    if not rows:
        return []
    header = [h.strip() for h in rows[0]]
    col = {name: header.index(name) for name in ("shots", "errors", "discards", "seconds", "decoder", "strong_id", "json_metadata")}
    for row in rows[1:]:
        if not row or len(row) <= col["json_metadata"]:
            continue
        sid = row[col["strong_id"]].strip()
        meta_txt = row[col["json_metadata"]].strip()
        inc = TaskStats(
            strong_id=sid,
            decoder=row[col["decoder"]].strip(),
            json_metadata=json.loads(meta_txt) if meta_txt else None,
            shots=int(row[col["shots"]]),
            errors=int(row[col["errors"]]),
            discards=int(row[col["discards"]]),
            seconds=float(row[col["seconds"]]),
        )
        if "custom_counts" in header and len(row) > header.index("custom_counts"):
            txt = row[header.index("custom_counts")].strip()
            if txt:
                inc.custom_counts = {str(k): int(v) for k, v in json.loads(txt).items()}
        if sid in merged:
            m = merged[sid]
            m.shots += inc.shots
            m.errors += inc.errors
            m.discards += inc.discards
            m.seconds += inc.seconds
            for k, v in inc.custom_counts.items():
                # "next_shot" is a position in the task's RNG stream (the furthest any recorded run got), not a count
                m.custom_counts[k] = max(m.custom_counts.get(k, 0), v) if k == "next_shot" else m.custom_counts.get(k, 0) + v
        else:
            merged[sid] = inc
    return list(merged.values())


def append_stats_to_csv(path: str, stats: TaskStats) -> None:
    """Append one row (writing the header first when the file is new or empty)."""
    import os

    new = not os.path.exists(path) or os.path.getsize(path) == 0
    with open(path, "a") as f:
        if new:
            f.write(CSV_HEADER + "\n")
        f.write(stats.to_csv_line() + "\n")
