"""Batches of an evaluation sweep decoded concurrently on one GPU.

The reference's evaluation loop (test.py:27-40: ``for img, label in loader: pred = model.generate(max_len=.., src=img)``)
decodes one batch at a time.  On a B200 one batch-512 generate call is bound by the *latency* of its dependent kernel
chain (47 small launches per decode step): HBM is ~40 % busy and most SMs idle.  Batches are
independent, so this module keeps several of them in flight: ``in_flight`` native engine handles (replicas of the same
weights, each with its own KV cache, workspaces, streams and CUDA graphs), one host thread and one CUDA stream per handle
(the ctypes call drops the GIL).  Every batch still goes through the same ``texocr_generate`` call with the same kernels,
so the tokens are bit-identical to ``model.generate`` (tests/test_gpu_parity.py).

    pipe = GeneratePipeline(model, in_flight=6)
    for tokens in pipe.generate_batches(batches, max_len=256):     # yields in submission order
        ...
"""
from __future__ import annotations

import queue
import threading
from typing import Iterable, Iterator, List, Optional, Sequence

import torch

from ._lib import Engine


class GeneratePipeline:
    """``in_flight`` engine replicas of ``model`` on its device, fed from a queue.

    ``branches``: decode branches per replica (rows per kernel launch = batch / branches).  With several batches in flight
    the device sees ``in_flight * branches`` independent kernel chains; fewer, larger launches per chain are better then
    (measured on B200 at batch 512 with the projected-K/V attention: in_flight 6 x 1 branch 7.9 k eq/s, 8 x 1: 7.9 k, 4 x 2: 7.5 k,
    4 x 4: 7.4 k, 3 x 6: 6.8 k, one batch at a time with 6 branches 5.5 k; with the absorbed attention 6 x 1: 11.0-11.1 k, 8 x 1 about the same,
    6 x 2: 5 % less).
    """

    def __init__(self, model, in_flight: int = 6, branches: Optional[int] = 1, engines: Optional[Sequence] = None):
        if in_flight < 1:
            raise ValueError("in_flight must be >= 1")
        self.model = model
        if engines is not None:
            # host-logic tests inject stand-in engines (objects with generate(images, max_len, out=)); no device is touched
            self.device = None
            self.engines = list(engines)
        else:
            if model.device.type != "cuda":
                raise RuntimeError("texocr_b200 runs on a CUDA (sm_100a) device only; there is no CPU fallback")
            self.device = model.device
            idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
            sd = model.state_dict()
            self.engines = [Engine(model.dims, sd, model.precision, idx) for _ in range(in_flight)]
            if branches:
                for e in self.engines:
                    e.set_option("decode_branches", int(branches))
        in_flight = len(self.engines)
        self._q: "queue.Queue" = queue.Queue()
        self._threads = [threading.Thread(target=self._worker, args=(i,), daemon=True) for i in range(in_flight)]
        for t in self._threads:
            t.start()

    # ------------------------------------------------------------------ workers
    def _worker(self, i: int):
        eng = self.engines[i]
        stream = None
        if self.device is not None:
            torch.cuda.set_device(self.device)
            stream = torch.cuda.Stream(device=self.device)
        while True:
            job = self._q.get()
            if job is None:
                return
            images, max_len, out, ready, slot, results, done = job
            try:
                if stream is None:
                    results[slot] = eng.generate(images, max_len, out=out)
                    continue
                with torch.cuda.stream(stream):
                    if ready is not None:
                        stream.wait_event(ready)              # inputs were produced on the caller's stream
                    results[slot] = eng.generate(images, max_len, out=out)     # returns after its stream has drained
            except BaseException as ex:                        # surfaced by the caller
                results[slot] = ex
            finally:
                done.release()

    # ------------------------------------------------------------------ API
    def generate_batches(self, batches: Iterable, max_len: int, outs: Optional[Sequence[torch.Tensor]] = None) -> Iterator[torch.Tensor]:
        """Greedy token ids (B_i, n_steps_i) int64 for every batch, yielded in order.  A batch is whatever
        ``model.generate`` takes as ``src`` (host or device tensors, or a ragged list); ``outs[i]`` may name the result
        buffer of batch i (e.g. pinned host memory)."""
        batches = list(batches)
        n = len(batches)
        results: List = [None] * n
        done = [threading.Semaphore(0) for _ in range(n)]
        ready = None
        def on_device(b):
            if isinstance(b, torch.Tensor):
                return b.is_cuda
            return any(isinstance(t, torch.Tensor) and t.is_cuda for t in b)        # ragged batch: a list of image tensors

        if self.device is not None and any(on_device(b) for b in batches):
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(self.device))
        for i, b in enumerate(batches):
            self._q.put((b, int(max_len), None if outs is None else outs[i], ready, i, results, done[i]))
        for i in range(n):
            done[i].acquire()
            if isinstance(results[i], BaseException):
                raise results[i]
            yield results[i]

    def warm_up(self, batch, max_len: int):
        """One generate call on every replica from the calling thread (allocations, CUDA-graph capture), so that no timed
        batch pays for them.  Must not overlap generate_batches."""
        for e in self.engines:
            e.generate(batch, int(max_len))
        if self.device is not None:
            torch.cuda.synchronize(self.device)

    def kernel_launches(self) -> int:
        return sum(e.kernel_launches() for e in self.engines)

    def close(self):
        for _ in self._threads:
            self._q.put(None)
        for t in self._threads:
            t.join(timeout=10)
        for e in self.engines:
            if hasattr(e, "close"):
                e.close()
        self.engines = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
