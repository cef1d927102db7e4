"""One training iteration as one CUDA graph.

The reference's iteration (main.py:170-207: zero_grad -> model(X) -> MSELoss -> accuracy -> backward -> Adam) is
~480 kernel launches of 5-300 us each; enqueueing them one by one from Python costs as much host time as the
GPU needs to run them.  ``GraphedStep`` runs the caller's own step function a few times eagerly (so every grow-only
workspace and every lazily created optimizer state exists), captures it once on a side stream, and afterwards an
iteration is: copy the new batch into the captured input buffers, ``cudaGraphLaunch``, read the captured outputs.

Everything the step enqueues must be capturable: the C-ABI kernels are (they only launch on the caller's stream
and never synchronise), NCCL collectives are, ``torch.optim.Adam`` is with ``capturable=True``.  A host read inside
the step (``.item()``, ``float(loss)``) is not -- read the returned tensors after the call instead, as the
reference does at main.py:207.
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional, Sequence

import torch

from . import _lib


class GraphedStep:
    """``step = GraphedStep(fn, (X, Z), model=model, optimizer=optimizer)``; then ``loss = step(X, Z)``.

    fn(*inputs) must return a tensor or a tuple of tensors.  ``inputs`` passed at call time may live on the device
    or in (pinned) host memory; they are copied into the captured input buffers on the current stream.
    If ``model`` / ``optimizer`` are given their state is put back, in place, to what it was before the warm-up
    iterations, so capturing does not advance training.
    """

    def __init__(self, fn: Callable, example_inputs: Sequence[torch.Tensor], device=None, warmup: int = 3,
                 model: Optional[torch.nn.Module] = None, optimizer: Optional[torch.optim.Optimizer] = None,
                 capture_error_mode: str = "global"):
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedStep needs a CUDA device -- this package has no CPU path")
        _lib.load()
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.fn = fn
        self.static_inputs = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in example_inputs]
        for s, t in zip(self.static_inputs, example_inputs):
            s.copy_(t, non_blocking=True)

        saved = self._snapshot(model, optimizer)
        cur = torch.cuda.current_stream(dev)
        # Warm up and capture on a HIGH-priority stream: streams the step forks internally (the weight-gradient stream
        # of engine.py) keep the default, lower priority, so the block scheduler serves the dependent chain first and
        # fills the gaps with the independent weight-gradient CTAs.
        side = torch.cuda.Stream(device=dev, priority=-1)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                fn(*self.static_inputs)
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)

        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph, stream=side, capture_error_mode=capture_error_mode):
            self.static_outputs = fn(*self.static_inputs)
        #: kernels of libzsv_b200.so inside one replay (the host-side launch counter does not see replays)
        self.launches_per_replay = _lib.launch_count() - n0
        self._restore(saved, model, optimizer)
        self.replays = 0
        # double buffering of the inputs: prefetch() uploads the NEXT batch on a copy stream while the current replay runs
        self._staging = None
        self._copy_stream = None
        self._staged = None          # event: staging buffers hold a complete batch
        self._staging_free = None    # event: the replay that consumed the staging buffers has copied them out
        self._has_prefetch = False

    # -- state kept out of the capture's side effects ------------------------------------------------------------
    @staticmethod
    def _state_tensors(optimizer) -> Iterable[torch.Tensor]:
        for st in optimizer.state.values():
            for v in st.values():
                if torch.is_tensor(v):
                    yield v

    def _snapshot(self, model, optimizer):
        saved = {}
        if model is not None:
            for t in list(model.parameters()) + list(model.buffers()):
                saved[id(t)] = t.detach().clone()
        if optimizer is not None:
            for t in self._state_tensors(optimizer):
                saved[id(t)] = t.detach().clone()
        return saved

    def _restore(self, saved, model, optimizer):
        with torch.no_grad():
            if model is not None:
                for t in list(model.parameters()) + list(model.buffers()):
                    t.copy_(saved[id(t)])
            if optimizer is not None:
                for t in self._state_tensors(optimizer):
                    if id(t) in saved:
                        t.copy_(saved[id(t)])
                    else:
                        t.zero_()      # state created lazily by the warm-up iterations (Adam: step, exp_avg, exp_avg_sq)

    # -- input prefetch ------------------------------------------------------------------------------------------
    def prefetch(self, *inputs: torch.Tensor) -> None:
        """Start copying the next batch (pinned host or device tensors of the captured shapes) into staging buffers on
        a side stream; the next ``step()`` call without arguments consumes it.  This is the loader-side overlap the
        reference gets from DataLoader workers + ``.cuda()`` (main.py:166-167): the H2D copy of batch i+1 runs while
        the graph of batch i executes."""
        if len(inputs) != len(self.static_inputs):
            raise RuntimeError(f"GraphedStep.prefetch: expected {len(self.static_inputs)} inputs, got {len(inputs)}")
        for s, t in zip(self.static_inputs, inputs):
            if tuple(t.shape) != tuple(s.shape) or t.dtype != s.dtype:
                raise RuntimeError("GraphedStep.prefetch: shape/dtype differs from the captured batch; pass ragged "
                                   "batches to __call__ directly")
        dev = self.static_inputs[0].device
        if self._staging is None:
            self._staging = [torch.empty_like(s) for s in self.static_inputs]
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._staged = torch.cuda.Event()
            self._staging_free = torch.cuda.Event()
            self._staging_free.record(torch.cuda.current_stream(dev))
        self._copy_stream.wait_event(self._staging_free)
        with torch.cuda.stream(self._copy_stream):
            for st, t in zip(self._staging, inputs):
                st.copy_(t, non_blocking=True)
            self._staged.record(self._copy_stream)
        self._has_prefetch = True

    # -- one iteration -------------------------------------------------------------------------------------------
    def __call__(self, *inputs: torch.Tensor):
        if not inputs and self._has_prefetch:
            cur = torch.cuda.current_stream(self.static_inputs[0].device)
            cur.wait_event(self._staged)
            for s, st in zip(self.static_inputs, self._staging):
                s.copy_(st, non_blocking=True)
            self._staging_free.record(cur)
            self._has_prefetch = False
            self._replay()
            return self.static_outputs
        if len(inputs) != len(self.static_inputs):
            raise RuntimeError(f"GraphedStep: expected {len(self.static_inputs)} inputs, got {len(inputs)}")
        if any(tuple(t.shape) != tuple(s.shape) or t.dtype != s.dtype for s, t in zip(self.static_inputs, inputs)):
            # ragged batch (the reference filters broken samples, main.py:156-158, and keeps the short last batch,
            # dataset.py:28 drop_last=False): same step function, enqueued kernel by kernel
            dev = self.static_inputs[0].device
            return self.fn(*(t.to(dev, non_blocking=True) for t in inputs))
        for s, t in zip(self.static_inputs, inputs):
            if t is not s:
                s.copy_(t, non_blocking=True)
        self._replay()
        return self.static_outputs

    def _replay(self) -> None:
        from . import engine, optim
        engine.note_weights_changed()    # a replay rewrites parameters / running statistics without any Python running
        engine.ensure_packed_fresh()     # weight images the captured forward reads must match the current masters
        optim.sync_all_lr()              # learning rates live in device scalars: a schedule reaches the replays
        self.graph.replay()
        self.replays += 1
