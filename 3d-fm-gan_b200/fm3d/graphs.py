"""CUDA-graph replay of an engine plan.

A plan's forward is a fixed sequence of ~50-100 libfm3d launches on persistent buffers, many of
them a few microseconds long; issuing them from Python costs more than running them.  After two
eager warm-up calls the sequence is captured once per input signature (shapes + which optional
inputs are present) and replayed: inputs are copied into static buffers, outputs are cloned out.
``FM3D_GRAPH=0`` disables capture.  A plan invalidates its graphs when derived weights change.
"""
import os

import torch

_ENABLED = os.environ.get("FM3D_GRAPH", "1") != "0"


class GraphRunner:
    def __init__(self, fn, warmup=2):
        self.fn = fn                    # fn(*tensors_or_None) -> tensor | list[tensor]
        self.warmup = warmup
        self.entries = {}

    def invalidate(self):
        self.entries = {}

    @staticmethod
    def _sig(args):
        return tuple(None if a is None else (tuple(a.shape), a.dtype, a.device.index) for a in args)

    def __call__(self, *args):
        from . import ops
        with ops.splitk_scope(self):
            return self._call(*args)

    def _call(self, *args):
        from . import _lib, ops
        if not _ENABLED or ops.PROFILE is not None or torch.cuda.is_current_stream_capturing():
            return self.fn(*args)
        sig = self._sig(args)
        ent = self.entries.get(sig)
        if ent is None:
            ent = self.entries[sig] = {"calls": 0, "graph": None}
        if ent["graph"] is None:
            ent["calls"] += 1
            if ent["calls"] <= self.warmup:
                return self.fn(*args)
            static_in = [None if a is None else a.detach().clone() for a in args]
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(graph):
                out = self.fn(*static_in)
            ent.update(graph=graph, static_in=static_in, out=out, launches=_lib.launch_count() - n0)
            # the capture pass does not execute: fall through to a replay with the real inputs
        for s, a in zip(ent["static_in"], args):
            if s is not None:
                s.copy_(a, non_blocking=True)
        ent["graph"].replay()
        _lib.lib().fm_add_launches(ent["launches"])      # replayed kernels bypass the launch wrappers
        out = ent["out"]
        if isinstance(out, (list, tuple)):
            return [o.clone() for o in out]
        return out.clone()
