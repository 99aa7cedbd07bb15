"""CUDA-graph replay of a generator call.

The DETR / LXMERT / ViT / VisualBERT generators are host-side tapes: one ctypes call per kernel, ~1-3 thousand launches
per call (forward, dgrad, rules).  At DETR-R50 / LXMERT-base sizes the kernels are a few microseconds each, so the call is
bound by launch latency, not by the GPU.  ``Graphed`` runs the call once under stream capture and replays
the recorded graph afterwards: the same kernels, the same arithmetic, one ``cudaGraphLaunch`` per call.

Shapes, flags and the engine are frozen at capture time; only the CONTENTS of the tensor arguments change between
replays (they are copied into the static input buffers, device-to-device or from pinned host memory).  The returned
tensors are static too: a replay overwrites them.  Host-side asserts of the generators (``assert diag(R - I) >= 0``,
DETR/modules/ExplanationGenerator.py:50) cannot run inside a capture; the generator leaves its operand on the device
(``generator.min_diag``) and :meth:`Graphed.check` performs the assert after a replay.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch

from ._lib import MmxError


def _flatten(out):
    if isinstance(out, torch.Tensor):
        return [out]
    if isinstance(out, (tuple, list)):
        return [t for o in out for t in _flatten(o)]
    if isinstance(out, dict):
        return [t for o in out.values() for t in _flatten(o)]
    return []


class Graphed:
    """``g = Graphed(fn, example_args, device)``; ``out = g(*args)`` replays.  ``fn`` must enqueue all its work on
    torch's current stream, allocate through torch, and not synchronise (no ``.item()``, no host reads)."""

    def __init__(self, fn: Callable, example_args: Sequence, device, warmup: int = 2, deferred_checks: Sequence[Callable] = ()):
        if not torch.cuda.is_available():
            raise MmxError("mmx_b200 needs a CUDA (sm_100) device; there is no CPU fallback")
        self.device = torch.device(device)
        self.fn = fn
        self.checks = list(deferred_checks)
        with torch.cuda.device(self.device):
            self.static_in = [a.to(self.device).clone() if isinstance(a, torch.Tensor) else a for a in example_args]
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):          # lazy one-time setup (function attributes, LRP weight splits) happens here
                    fn(*self.static_in)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(self.device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.static_out = fn(*self.static_in)
        self.replays = 0

    def __call__(self, *args):
        if len(args) != len(self.static_in):
            raise MmxError("argument count differs from the captured call")
        with torch.cuda.device(self.device):
            for s, a in zip(self.static_in, args):
                if isinstance(s, torch.Tensor):
                    if not isinstance(a, torch.Tensor) or tuple(a.shape) != tuple(s.shape):
                        raise MmxError(f"argument shape {getattr(a, 'shape', None)} differs from the captured {tuple(s.shape)}")
                    if a.data_ptr() != s.data_ptr():
                        s.copy_(a, non_blocking=True)
                elif a != s:
                    raise MmxError("non-tensor arguments are frozen at capture time")
            self.graph.replay()
        self.replays += 1
        return self.static_out

    def outputs(self):
        return _flatten(self.static_out)

    def check(self):
        """Runs the host-side asserts deferred by the capture (one synchronisation)."""
        for c in self.checks:
            c()
