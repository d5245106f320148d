"""CUDA-graph replay of the inference paths (small-batch / latency use: the ov-* scripts and cliptoolsoptimized.py run the
towers at batch 1-8, where a forward is ~130 kernel launches, each with its ctypes crossing and 3-5 host-encoded tensor
maps — the launch path, not the GPU, sets the latency).

`GraphedCall` records one call of a no-grad function of fixed-shape CUDA tensors into a `torch.cuda.CUDAGraph` (stream
capture sees the libovk launches because they are enqueued on torch's current stream) and replays it: one `cudaGraphLaunch`
per forward.  Tensor maps are encoded at capture time, so the captured kernels keep reading the static input buffer and the
parameter copies that were live then: rebuild the object after changing weights (a version check raises otherwise).
No tracing compiler is involved; the kernels are exactly the ones the eager path launches.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch

from ._lib import OvkError


class GraphedCall:
    def __init__(self, fn: Callable, example_inputs: Sequence[torch.Tensor], params: Sequence[torch.Tensor] = (), warmup: int = 2):
        if not example_inputs or any(not t.is_cuda for t in example_inputs):
            raise OvkError("GraphedCall: inputs must be CUDA tensors (there is no CPU path)")
        self._fn = fn
        self._params = list(params)
        self._versions = [(p.data_ptr(), p._version) for p in self._params]
        self.static_inputs = [t.detach().clone() for t in example_inputs]
        side = torch.cuda.Stream(device=self.static_inputs[0].device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(max(1, warmup)):     # fills the packed-weight caches, latches kernel attributes, allocator warm-up
                fn(*self.static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_output = fn(*self.static_inputs)

    def __call__(self, *inputs: torch.Tensor):
        """copies `inputs` into the captured buffers, replays, returns the STATIC output tensor(s) (overwritten by the next
        call: clone to keep)"""
        if [(p.data_ptr(), p._version) for p in self._params] != self._versions:
            raise OvkError("GraphedCall: a parameter changed since capture (the graph reads the packed copies made then); "
                           "build a new GraphedCall")
        if len(inputs) != len(self.static_inputs):
            raise OvkError("GraphedCall: wrong number of inputs")
        for dst, src in zip(self.static_inputs, inputs):
            if dst.shape != src.shape:
                raise OvkError(f"GraphedCall: captured for shape {tuple(dst.shape)}, got {tuple(src.shape)}")
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_output


def graphed_encode_image(model, example_images: torch.Tensor, normalize: bool = True) -> GraphedCall:
    """`CLIP.encode_image(images, normalize)` (model.py:265-267) as one graph launch for the batch shape of `example_images`."""
    return GraphedCall(lambda x: model.encode_image(x, normalize=normalize), [example_images], list(model.visual.parameters()))


def graphed_encode_text(model, example_tokens: torch.Tensor, normalize: bool = True) -> GraphedCall:
    """`CLIP.encode_text(tokens, normalize)` (model.py:269-284) as one graph launch."""
    params = [p for n, p in model.named_parameters() if not n.startswith("visual.")]
    return GraphedCall(lambda t: model.encode_text(t, normalize=normalize), [example_tokens], params)
