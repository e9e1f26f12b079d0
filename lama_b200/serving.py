"""Throughput-oriented public entry point: a copy/compute pipeline around the generator program.

``bin/predict.py`` feeds one image at a time from host memory (predict.py:67-94).  For batched serving the
host<->device copies of one batch (134 MB in, 100 MB out at bs32 512x512) would add ~10% to every step if they
ran serially with the kernels, so :class:`GeneratorPipeline` keeps ``depth`` batches in flight on three CUDA
streams: H2D of batch i+1 and D2H of batch i-1 overlap the CUDA-graph replay of batch i.

    pipe = GeneratorPipeline(generator, batch=32, height=512, width=512)
    t = pipe.submit(x_pinned)        # (B,4,H,W) float32, pinned host memory; returns immediately
    y = pipe.result(t)               # (B,3,H,W) float32 pinned host tensor (valid until the slot is reused)
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import engine as E


class GeneratorPipeline:
    def __init__(self, generator, batch: int, height: int, width: int, device: Optional[torch.device] = None,
                 depth: int = 2, math: Optional[int] = None):
        dev = device if device is not None else next(generator.parameters()).device
        assert dev.type == "cuda", "GeneratorPipeline needs the generator on a CUDA device"
        self.device, self.depth = dev, depth
        cin = generator.model[1].ffc.convl2l.in_channels
        probe = torch.empty(batch, cin, height, width, device=dev)
        if not E.generator_supported(generator, probe):
            raise ValueError("generator / shape is outside the native path")
        self.ex = E.get_executor(generator, "generator", (probe,), math=math)
        del probe
        self.graph = E.GraphedProgram(self.ex)
        self.in_shape = tuple(self.ex.prog.inputs["x0"])
        self.out_shape = tuple(self.ex.prog.outputs["y0"])
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        mk = lambda shape: [torch.empty(shape, device=dev) for _ in range(depth)]  # noqa: E731
        self.dev_in, self.dev_out = mk(self.in_shape), mk(self.out_shape)
        self.host_out = [torch.empty(self.out_shape).pin_memory() for _ in range(depth)]
        ev = lambda: [torch.cuda.Event() for _ in range(depth)]  # noqa: E731
        self.ev_in, self.ev_run, self.ev_out, self.ev_free = ev(), ev(), ev(), ev()
        self._n = 0
        self._pending: List[int] = []

    def submit(self, x_host: torch.Tensor) -> int:
        """Enqueue one batch (pinned host float32, shape (B,4,H,W)); returns a ticket for :meth:`result`."""
        assert tuple(x_host.shape) == self.in_shape and x_host.dtype == torch.float32
        n, slot = self._n, self._n % self.depth
        if n >= self.depth:
            self.ev_out[slot].synchronize()          # the host buffer of this slot must have been drained
        with torch.cuda.stream(self.s_in):
            if n >= self.depth:
                self.s_in.wait_event(self.ev_free[slot])     # its device input must have been consumed
            self.dev_in[slot].copy_(x_host, non_blocking=True)
            self.ev_in[slot].record(self.s_in)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(self.ev_in[slot])
            if n >= self.depth:
                self.s_run.wait_event(self.ev_out[slot])     # D2H of the previous occupant has read dev_out[slot]
            self.graph.static_in["x0"].copy_(self.dev_in[slot], non_blocking=True)
            self.ev_free[slot].record(self.s_run)
            self.graph.graph.replay()
            self.dev_out[slot].copy_(self.ex.outputs["y0"], non_blocking=True)
            self.ev_run[slot].record(self.s_run)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_run[slot])
            self.host_out[slot].copy_(self.dev_out[slot], non_blocking=True)
            self.ev_out[slot].record(self.s_out)
        self._n += 1
        return n

    def result(self, ticket: int) -> torch.Tensor:
        """Block until batch ``ticket`` is on the host; the tensor is reused ``depth`` submissions later."""
        assert self._n - self.depth <= ticket < self._n, "result no longer (or not yet) available"
        slot = ticket % self.depth
        self.ev_out[slot].synchronize()
        return self.host_out[slot]

    def drain(self):
        for s in (self.s_in, self.s_run, self.s_out):
            s.synchronize()

    @property
    def launches_per_batch(self) -> int:
        return self.ex.launches_per_run
