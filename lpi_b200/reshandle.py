"""Post-processing of the continual-retrieval results written by SPrompts.incremental_train (`./res/<datetime>.json`), the
counterpart of the reference's retrieval/res_handle/reshandle.py:52-121: average Recall@1/5/10 over tasks and sessions and the
forgetting measure (last-session recall minus the best earlier recall, averaged over tasks).

Result schema (unchanged from the reference, sprompt.py:638-646): {session: {dataset: {side: {task: [R@1, R@5, R@10]}}}} with
dataset = 'mscoco' and side in {'i2t', 't2i'}; JSON turns the integer keys into strings, both are accepted here.
"""
from __future__ import annotations

import json
from typing import Dict, List, Optional, Sequence


def load_results(path: str) -> dict:
    with open(path) as f:
        return json.load(f)


def _get(d: dict, k):
    return d[k] if k in d else d[str(k)]


class TaskHistory:
    """Recall triples of ONE task over the sessions in which it was evaluated (reshandle.py `Eval`)."""

    def __init__(self):
        self.history: List[List[float]] = []

    def insert(self, triple: Sequence[float]):
        self.history.append([float(x) for x in triple])

    @property
    def cnt(self) -> int:
        return len(self.history)

    def mean(self) -> List[float]:
        return [sum(h[k] for h in self.history) / self.cnt for k in range(3)]

    def last(self) -> List[float]:
        return self.history[-1]

    def forgetting(self) -> List[float]:
        """last - max(earlier sessions); 0 for a task seen in a single session (reshandle.py:39-49)."""
        if self.cnt < 2:
            return [0.0, 0.0, 0.0]
        return [self.history[-1][k] - max(h[k] for h in self.history[:-1]) for k in range(3)]


def summarize(results: dict, dataset: str = "mscoco", side: str = "i2t", task_sizes: Optional[Sequence[int]] = None) -> Dict:
    """-> {'avg_recall': [R@1, R@5, R@10] (unweighted mean over tasks of each task's mean over sessions, reshandle.py 'org average'),
           'weighted_recall': same weighted by task_sizes (reshandle.py:57 num_list) or None,
           'forgetting': per-K mean over tasks 0..n-2 divided by (n - 1) as reshandle.py:97-99, 'avg_forgetting': their mean,
           'final': {task: triple of the last session}, 'per_task': {task: {'mean', 'last', 'forgetting', 'sessions'}}}"""
    sessions = sorted(int(k) for k in results.keys())
    tasks: Dict[int, TaskHistory] = {}
    for s in sessions:
        table = _get(_get(_get(results, s), dataset), side)
        for t in sorted(int(k) for k in table.keys()):
            tasks.setdefault(t, TaskHistory()).insert(_get(table, t))
    n = len(tasks)
    if n == 0:
        raise ValueError("no tasks in the results")
    order = sorted(tasks)
    avg = [sum(tasks[t].mean()[k] for t in order) / n for k in range(3)]
    weighted = None
    if task_sizes is not None:
        if len(task_sizes) < n:
            raise ValueError("task_sizes shorter than the number of tasks")
        tot = float(sum(task_sizes[:n]))
        weighted = [sum(tasks[t].mean()[k] * task_sizes[i] for i, t in enumerate(order)) / tot for k in range(3)]
    forget = [sum(tasks[t].forgetting()[k] for t in order) / max(1, n - 1) for k in range(3)]
    return {"avg_recall": avg, "weighted_recall": weighted, "forgetting": forget, "avg_forgetting": sum(forget) / 3.0,
            "final": {t: tasks[t].last() for t in order},
            "per_task": {t: {"mean": tasks[t].mean(), "last": tasks[t].last(), "forgetting": tasks[t].forgetting(),
                             "sessions": tasks[t].cnt} for t in order}}


def get_res(filename: str, task_name: str = "mscoco", task_type: str = "i2t", n: Optional[int] = None) -> Dict:
    """Same entry point as reshandle.py:113-118 (prints the summary lines and returns the dict)."""
    out = summarize(load_results(filename), task_name, task_type)
    a, f = out["avg_recall"], out["forgetting"]
    print(f"org average precision: P@1: {a[0]}, P@5: {a[1]}, P@10: {a[2]}")
    print(f"average forget: P1: {f[0]}, P5: {f[1]}, P10:{f[2]}, avg forget: {out['avg_forgetting']}")
    for t, d in out["per_task"].items():
        print(f"task {t}: mean {d['mean']} last {d['last']} forgetting {d['forgetting']}")
    return out
