"""One training step as ONE CUDA graph.

The hot path (src/models/backbone.py:147-222 + the loss block and optimizer step of tools/ssl_train.py:441-474) is ~33
kernel launches whose device time (2.5 ms at batch 256) is a sixth of what Python, autograd and the launch path spend
issuing them.  Nothing in the step syncs with the host -- upstream gradient, learning rates, step count and GradScaler state
are device scalars, tensor maps are encoded at capture time, the optimizer's pointer table comes from pinned memory -- so the
whole step can be captured once and replayed.  Single GPU: the cross-rank exchanges carry host-side sequence numbers.
"""
from __future__ import annotations

from typing import Callable, Dict, Sequence, Union

import torch

Tensors = Union[torch.Tensor, Sequence[torch.Tensor]]


class GraphedStep:
    """``GraphedStep(step_fn, static_inputs, optimizer)`` runs ``step_fn(static_inputs)`` -- forward, ``zero_grad``,
    backward, ``optimizer.step()``, returning the loss -- ``warmup`` times on a side stream (PyTorch's whole-network capture
    recipe: lazy initialisations, cuDNN autotuning and the optimizer state must exist before the capture), then captures it.
    Calling the object copies ``inputs`` (same structure: a dict of tensors or lists of tensors) into the captured tensors
    and replays; it returns the static loss tensor, which the next replay overwrites.  A replay reproduces the eager step
    bit for bit (tests/test_cuda_graph_gpu.py)."""

    def __init__(self, step_fn: Callable[[Dict[str, Tensors]], torch.Tensor], static_inputs: Dict[str, Tensors],
                 optimizer: torch.optim.Optimizer, warmup: int = 3):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            raise RuntimeError("GraphedStep is single-GPU: the cross-rank batch-norm exchanges carry host-side sequence numbers")
        from . import _lib
        self.static_inputs = static_inputs
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                step_fn(static_inputs)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)  # the gradients of the replayed step live in the graph's memory pool
        before = _lib.launch_count
        with torch.cuda.graph(self.graph):
            self.loss = step_fn(static_inputs)
        self.launches_per_replay = _lib.launch_count - before  # C-ABI kernel launches recorded in the graph

    def __call__(self, inputs: Dict[str, Tensors] | None = None) -> torch.Tensor:
        from . import _lib
        if inputs is not None and inputs is not self.static_inputs:
            for k, v in inputs.items():
                dst = self.static_inputs[k]
                for d, s in zip(dst if isinstance(dst, (list, tuple)) else [dst], v if isinstance(v, (list, tuple)) else [v]):
                    d.copy_(s, non_blocking=True)
        self.graph.replay()
        _lib.launch_count += self.launches_per_replay
        return self.loss
