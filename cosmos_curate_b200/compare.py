"""Task comparison with the semantics of the reference's own stage-output comparator
(cosmos_curate/core/utils/misc/stage_compare.py:173-392 `_compare_values` and friends; the SplitPipeTask comparator checks
`session_id`, `videos`, `errors` and ignores `stage_perf`, pipelines/video/utils/data_model_compare.py:28-53).  It is how the
reference decides that a changed stage still produces "the same" tasks (stage replay: `run_stage_compare`), so it is the yardstick a
drop-in stage should be held to:

  * types must be identical (no int / float, list / tuple, ndarray / list leniency);
  * numeric arrays: equal shapes, then |golden - candidate| <= atol element-wise (rtol = 0), NaN equal to NaN; the reported
    maximum difference is computed in float64 (no unsigned wrap-around) over the elements that are not NaN in both;
  * non-numeric arrays: exact equality;
  * attrs instances: field by field; mappings: equal key sets, then value by value in repr-sorted key order; sequences (not str /
    bytes): equal lengths, then element by element; everything else: `==`.

`compare_values` returns the list of differences (empty = same); `compare_tasks` applies it to lists of tasks over the semantic
SplitPipeTask fields.  Used by the stage tests and usable against pickled reference outputs.
"""

from __future__ import annotations

from collections.abc import Mapping, Sequence

import attrs
import numpy as np

SPLIT_TASK_FIELDS = ("session_id", "videos", "errors")


@attrs.define(frozen=True)
class FieldDiff:
    field: str
    detail: str
    max_diff_observed: float | None = None
    shape_mismatch: bool = False


def _arrays(path: str, g: np.ndarray, c: np.ndarray, atol: float) -> list[FieldDiff]:
    if g.shape != c.shape:
        return [FieldDiff(path, f"shape mismatch golden={g.shape} new={c.shape}", shape_mismatch=True)]
    if np.issubdtype(g.dtype, np.number) and np.issubdtype(c.dtype, np.number):
        if np.allclose(g, c, atol=atol, rtol=0.0, equal_nan=True):
            return []
        g64, c64 = g.astype(np.float64), c.astype(np.float64)
        d = np.abs(g64 - c64)[~(np.isnan(g64) & np.isnan(c64))]
        worst = float(np.nanmax(d)) if d.size > 0 else 0.0
        return [FieldDiff(path, f"max diff {worst}", max_diff_observed=worst)]
    return [] if np.array_equal(g, c) else [FieldDiff(path, "array values differ")]


def compare_values(path: str, golden: object, candidate: object, *, atol: float) -> list[FieldDiff]:
    if type(golden) is not type(candidate):
        return [FieldDiff(path, f"type mismatch golden={type(golden).__name__} new={type(candidate).__name__}")]
    if isinstance(golden, np.ndarray):
        return _arrays(path, golden, candidate, atol)
    out: list[FieldDiff] = []
    if attrs.has(golden.__class__):
        for f in attrs.fields(golden.__class__):
            out += compare_values(f"{path}.{f.name}" if path else f.name, getattr(golden, f.name), getattr(candidate, f.name), atol=atol)
        return out
    if isinstance(golden, Mapping):
        gk, ck = set(golden.keys()), set(candidate.keys())
        if gk != ck:
            return [FieldDiff(path, f"dict key mismatch golden={sorted(gk, key=repr)!r} new={sorted(ck, key=repr)!r}")]
        for k in sorted(gk, key=repr):
            out += compare_values(f"{path}.{k}" if path else str(k), golden[k], candidate[k], atol=atol)
        return out
    if isinstance(golden, Sequence) and not isinstance(golden, (str, bytes, bytearray)):
        if len(golden) != len(candidate):
            return [FieldDiff(path, f"length mismatch golden={len(golden)} new={len(candidate)}")]
        for i, (a, b) in enumerate(zip(golden, candidate, strict=True)):
            out += compare_values(f"{path}[{i}]", a, b, atol=atol)
        return out
    return [] if golden == candidate else [FieldDiff(path, f"value mismatch golden={golden!r} new={candidate!r}")]


def compare_tasks(golden_tasks, candidate_tasks, *, atol: float, field_names: Sequence[str] | None = SPLIT_TASK_FIELDS) -> list[tuple[int, FieldDiff]]:
    """[(task index, difference)] over the named task fields (None: every field of the task, the reference's generic comparator)."""
    if len(golden_tasks) != len(candidate_tasks):
        return [(0, FieldDiff("tasks", f"task count mismatch golden={len(golden_tasks)} new={len(candidate_tasks)}"))]
    out = []
    for i, (g, c) in enumerate(zip(golden_tasks, candidate_tasks, strict=True)):
        if field_names is None:
            diffs = compare_values("", g, c, atol=atol)
        else:
            diffs = [d for name in field_names for d in compare_values(name, getattr(g, name), getattr(c, name), atol=atol)]
        out += [(i, d) for d in diffs]
    return out
