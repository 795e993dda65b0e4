"""Subgrid water levels: piecewise-linear h(h) relations from the Basin level to the level of
each subgrid element (`Subgrid(db, config, basin)` core/src/read.jl:1568-1682,
`update_subgrid_level!` core/src/callback.jl:788-813, validation.jl:403-430).

Host-side and vectorised over the ensemble: the Basin levels come from the device path."""

from __future__ import annotations

import numpy as np

from .tables import ModelTables


def _check(subgrid_id: int, node_id: int, basin_of: dict[int, int], basin_level, subgrid_level) -> None:
    if node_id not in basin_of:
        raise ValueError(f"The node_id of the Basin / subgrid does not exist: node_id {node_id}, "
                         f"subgrid_id {subgrid_id}")
    if len(set(basin_level)) != len(basin_level):
        raise ValueError(f"Basin / subgrid subgrid_id {subgrid_id} has repeated basin levels, "
                         "this cannot be interpolated.")
    if len(set(subgrid_level)) != len(subgrid_level):
        raise ValueError(f"Basin / subgrid subgrid_id {subgrid_id} has repeated element levels, "
                         "this cannot be interpolated.")


class _Relation:
    """LinearInterpolation(subgrid_level, basin_level), constant to the left, linear to the right."""

    def __init__(self, basin_level, subgrid_level):
        self.t = np.asarray(basin_level, dtype=float)
        self.u = np.asarray(subgrid_level, dtype=float)

    def __call__(self, x: np.ndarray) -> np.ndarray:
        t, u = self.t, self.u
        if len(t) == 1:
            return np.full_like(x, u[0], dtype=float)
        y = np.interp(x, t, u)  # constant outside [t0, t_end]
        slope = (u[-1] - u[-2]) / (t[-1] - t[-2])
        return np.where(x > t[-1], u[-1] + slope * (x - t[-1]), y)

    def same(self, other: "_Relation") -> bool:
        return np.array_equal(self.t, other.t) and np.array_equal(self.u, other.u)


class Subgrid:
    """`subgrid_id` sorted ascending; `levels(basin_level, t)` -> `[n_subgrid][n_members]`."""

    def __init__(self, m: ModelTables, basin_node_ids: list[int]):
        basin_of = {int(v): k for k, v in enumerate(basin_node_ids)}
        cyclic = {r["node_id"]: bool(r.get("cyclic_time")) for r in m.node}
        entries: list[tuple[int, int, list[float], list[_Relation], bool]] = []
        static = m.table("Basin / subgrid") if "Basin / subgrid" in m.tables else []
        for sid in sorted({r["subgrid_id"] for r in static}):
            rows = [r for r in static if r["subgrid_id"] == sid]
            node_id = rows[0]["node_id"]
            bl, sl = [r["basin_level"] for r in rows], [r["subgrid_level"] for r in rows]
            _check(sid, node_id, basin_of, bl, sl)
            entries.append((sid, basin_of[node_id], [0.0], [_Relation(bl, sl)], False))
        timed = m.table("Basin / subgrid_time") if "Basin / subgrid_time" in m.tables else []
        for sid in sorted({r["subgrid_id"] for r in timed}):
            rows = [r for r in timed if r["subgrid_id"] == sid]
            node_id = rows[0]["node_id"]
            times, rels = [], []
            for t in sorted({m.seconds_since(r["time"]) for r in rows}):
                grp = [r for r in rows if m.seconds_since(r["time"]) == t]
                bl, sl = [r["basin_level"] for r in grp], [r["subgrid_level"] for r in grp]
                _check(sid, node_id, basin_of, bl, sl)
                times.append(t)
                rels.append(_Relation(bl, sl))
            per = cyclic.get(node_id, False)
            if per and not rels[0].same(rels[-1]):
                raise ValueError(f"For Basin #{node_id} with cyclic_time the first and last h(h) relations "
                                 f"for subgrid_id {sid} are not equal.")
            entries.append((sid, basin_of[node_id], times, rels, per))
        entries.sort(key=lambda e: e[0])
        ids = [e[0] for e in entries]
        if len(set(ids)) != len(ids):
            raise ValueError("Basin / subgrid: a subgrid_id occurs in both the static and the time table")
        self.entries = entries
        self.subgrid_id = np.array(ids, dtype=np.int32)
        self.basin_idx = np.array([e[1] for e in entries], dtype=np.int64)

    def __len__(self) -> int:
        return len(self.entries)

    @staticmethod
    def _current(times: list[float], periodic: bool, t: float) -> int:
        # ConstantInterpolation of the relation index: forward fill, constant (or periodic) outside
        if periodic and len(times) > 1 and (t < times[0] or t > times[-1]):
            t = (t - times[0]) % (times[-1] - times[0]) + times[0]
        k = int(np.searchsorted(times, t, side="right")) - 1
        k = min(max(k, 0), len(times) - 1)
        return 0 if periodic and k == len(times) - 1 else k

    def levels(self, basin_level: np.ndarray, t: float) -> np.ndarray:
        basin_level = np.asarray(basin_level, dtype=float)
        if basin_level.ndim == 1:
            basin_level = basin_level[:, None]
        out = np.empty((len(self.entries), basin_level.shape[1]))
        for k, (_, b, times, rels, per) in enumerate(self.entries):
            out[k] = rels[self._current(times, per, t)](basin_level[b])
        return out
