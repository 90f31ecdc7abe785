"""Host -> device staging for the render step: a two-deep prefetcher on a side stream.

LASR's trainer copies each batch with blocking ``.cuda()`` calls at the top of the step
(nnutils/train_utils.py:179-180,269), so PCIe time adds to the step.  At B200 render speed the
per-step inputs (target images: 4 planes x N x S^2 floats) take about as long to cross PCIe Gen5 as
they take to render, so the copy of step i+1 is overlapped with the compute of step i: pinned host
buffers, ``cudaMemcpyAsync`` on a copy stream, events in both directions (the copy waits until the
buffer's previous consumer has finished; the compute waits until the copy has landed)."""
from __future__ import annotations

import torch


class HostPrefetcher:
    """``stage(host_tensors)`` starts the async copy of one step's inputs; ``get()`` returns the
    device tensors of the oldest staged step (and makes the current stream wait for them);
    ``release()`` marks them free for reuse once the current stream has consumed them."""

    def __init__(self, device, depth: int = 2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.depth = depth
        self.slots = [dict(bufs=None, ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(depth)]
        self.head = 0      # next slot to stage into
        self.tail = 0      # next slot to hand out
        self.staged = 0
        for s in self.slots:
            s["free"].record(torch.cuda.current_stream(self.device))

    def stage(self, host_tensors):
        assert self.staged < self.depth, "prefetcher full: call get()/release() first"
        slot = self.slots[self.head]
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(slot["free"])           # previous consumer of this slot is done
            if slot["bufs"] is None:
                slot["bufs"] = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in host_tensors]
            for d, h in zip(slot["bufs"], host_tensors):
                d.copy_(h, non_blocking=True)
            slot["ready"].record(self.stream)
        self.head = (self.head + 1) % self.depth
        self.staged += 1

    def get(self):
        assert self.staged > 0, "nothing staged"
        slot = self.slots[self.tail]
        torch.cuda.current_stream(self.device).wait_event(slot["ready"])
        return slot["bufs"]

    def release(self):
        slot = self.slots[self.tail]
        slot["free"].record(torch.cuda.current_stream(self.device))
        self.tail = (self.tail + 1) % self.depth
        self.staged -= 1


class GraphedStep:
    """A whole render step (forward, loss, backward, gradient reduction) captured once into a CUDA graph
    and replayed, so the ~40 kernel launches of a step cost one launch.  Shapes are static in LASR's
    optimisation loop (fixed batch, mesh and image size), which is exactly the case CUDA graphs serve.

    ``fn(*static_inputs)`` must do everything on the current stream, allocate only through PyTorch, read
    its inputs from the given device tensors (the same storage at every replay) and return a tuple of
    tensors; ``GraphedStep(...)()`` replays and returns those (static) output tensors."""

    def __init__(self, fn, static_inputs, warmup: int = 3):
        self.static_inputs = list(static_inputs)
        dev = self.static_inputs[0].device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                     # warm-up off the capture stream (PyTorch's recipe)
            for _ in range(warmup):
                fn(*self.static_inputs)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = fn(*self.static_inputs)

    def __call__(self):
        self.graph.replay()
        return self.outputs
