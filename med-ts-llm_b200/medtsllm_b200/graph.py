"""CUDA-graph replay of an inference path made of libmtsb200 launches.

A steady stream of same-shaped batches (the reference Trainer's val / test loops, tasks/forecasting.py:55-78)
replays ONE captured graph instead of re-issuing every launch: the windows are copied into the graph's input
buffer, the graph runs, the predictions are copied out.  The caller supplies a `key` that must change whenever
anything the captured launches depend on changes (input shape, prompt table, `training` flag, `_version` and
address of every parameter); a key seen for the first time runs kernel by kernel — which also refreshes the
caller's weight caches — and the second sighting in a row captures.
"""
from __future__ import annotations

import torch

from ._lib import launch_count, note_replay


class GraphReplay:

    def __init__(self):
        self.entry = None     # {"key", "graph", "x", "out", "launches", "hold"}
        self.seen = None      # (key, hold) of the previous kernel-by-kernel call

    def run(self, key, x: torch.Tensor, fn, hold=None) -> torch.Tensor:
        """`fn(x) -> Tensor` issues the launches on the current stream; `hold` is kept alive with the key (objects
        whose id() is part of it)."""
        g = self.entry
        if g is not None and g["key"] == key:
            g["x"].copy_(x)
            g["graph"].replay()
            note_replay(g["launches"])
            return g["out"].clone()
        if self.seen is None or self.seen[0] != key:
            self.seen = (key, hold)
            return fn(x)
        self.entry = None                      # frees the previous graph's memory pool before capturing
        static_x = x.clone()
        graph = torch.cuda.CUDAGraph()
        n0 = launch_count()
        # thread_local: the reference's DataLoaders run a pin_memory thread in this process (tasks/base.py:175-197);
        # its cudaHostAlloc calls during our capture must not be treated as capture violations
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            out = fn(static_x)
        # (capturing records the launches without running them; the launch counter then follows the replays)
        self.entry = {"key": key, "graph": graph, "x": static_x, "out": out, "launches": launch_count() - n0, "hold": hold}
        graph.replay()
        return out.clone()

    @property
    def captured(self) -> bool:
        return self.entry is not None
