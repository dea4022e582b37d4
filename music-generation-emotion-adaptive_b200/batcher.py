"""Request coalescing in front of ``Generator.generate``.

The reference service handles one prompt per HTTP request; the endpoint is a sync ``def``, so Starlette runs concurrent
requests on threadpool threads that all call ``sample_kvcache`` on the module-global model (api_cache.py:186-204).  The
engine serialises calls (one mutex per engine), which turns N concurrent requests into N batch-1 runs.  ``RequestBatcher``
puts a queue in between: requests that arrive within ``max_wait_ms`` of each other (or until ``max_batch`` are waiting) and
share the sampling settings are decoded in ONE batched engine call -- row b of a batched call equals a batch-1 run on
prompt b, so nothing changes for the caller except the latency / throughput trade-off.

    batcher = RequestBatcher(model.engine, max_batch=64, max_wait_ms=2.0)
    tokens = batcher.generate(prompt_ids, max_new_tokens, temperature, top_k, eos_id)     # blocking, thread-safe
    ...
    batcher.close()

This is coalescing, not continuous batching: a batch runs to completion (the persistent decode kernel owns its sequences for
the whole generation; a cluster stops early when all of its sequences have hit EOS) before the next one starts.
"""
from __future__ import annotations

import itertools
import os
import threading
import time
from concurrent.futures import Future
from typing import Dict, List, Optional, Sequence, Tuple


class _Request:
    __slots__ = ("prompt", "max_new", "key", "future", "t_submit")

    def __init__(self, prompt, max_new, key, future):
        self.prompt, self.max_new, self.key, self.future = prompt, max_new, key, future
        self.t_submit = time.perf_counter()


class RequestBatcher:
    """Thread-safe front end: ``submit`` returns a Future, ``generate`` blocks.  One worker thread drives the engine."""

    def __init__(self, engine, max_batch: int = 64, max_wait_ms: float = 2.0, seed: Optional[int] = None):
        """``seed=None`` (default): a fresh 64-bit Philox key per batcher, so outputs do not repeat after a process
        restart (the reference samples from torch's global RNG); requests of one batcher never share a stream
        (sequence index = a running request counter)."""
        if max_batch < 1:
            raise ValueError("max_batch must be positive")
        self.engine, self.max_batch, self.max_wait = engine, int(max_batch), float(max_wait_ms) * 1e-3
        self._seed = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed)
        self._cv = threading.Condition()
        self._queue: List[_Request] = []
        self._closed = False
        self._seq = itertools.count()            # Philox sequence index base: no two requests ever share a stream
        self.batches: List[int] = []             # sizes of the engine calls made so far (observability / tests)
        self._worker = threading.Thread(target=self._run, name="mgea-batcher", daemon=True)
        self._worker.start()

    # -- caller side ----------------------------------------------------------------------------------
    def submit(self, prompt_ids: Sequence[int], max_new_tokens: int, temperature: float = 1.0, top_k: Optional[int] = 50,
               eos_id: int = -1) -> "Future[List[int]]":
        fut: "Future[List[int]]" = Future()
        req = _Request(list(prompt_ids), int(max_new_tokens), (float(temperature), top_k, int(eos_id)), fut)
        with self._cv:
            if self._closed:
                raise RuntimeError("RequestBatcher is closed")
            self._queue.append(req)
            self._cv.notify_all()
        return fut

    def generate(self, prompt_ids: Sequence[int], max_new_tokens: int, temperature: float = 1.0,
                 top_k: Optional[int] = 50, eos_id: int = -1) -> List[int]:
        """Blocking call with the engine's exceptions re-raised in the calling thread."""
        return self.submit(prompt_ids, max_new_tokens, temperature, top_k, eos_id).result()

    def close(self) -> None:
        with self._cv:
            self._closed = True
            self._cv.notify_all()
        self._worker.join()

    # -- worker -----------------------------------------------------------------------------------------
    def _take_batch(self) -> Optional[Tuple[Tuple, List[_Request]]]:
        """Waits for work, then for up to ``max_wait`` (counted from the oldest request) for more of the same kind."""
        with self._cv:
            while not self._queue and not self._closed:
                self._cv.wait()
            if not self._queue:
                return None
            key = self._queue[0].key
            deadline = self._queue[0].t_submit + self.max_wait
            while not self._closed:
                same = sum(1 for r in self._queue if r.key == key)
                left = deadline - time.perf_counter()
                if same >= self.max_batch or left <= 0:
                    break
                self._cv.wait(timeout=left)
            batch = [r for r in self._queue if r.key == key][:self.max_batch]
            taken = set(map(id, batch))
            self._queue = [r for r in self._queue if id(r) not in taken]
            return key, batch

    def _run(self) -> None:
        while True:
            item = self._take_batch()
            if item is None:
                return
            (temperature, top_k, eos_id), batch = item
            base = next(self._seq) * self.max_batch
            try:
                out = self.engine.generate([r.prompt for r in batch], [r.max_new for r in batch], temperature, top_k,
                                           eos_id=eos_id, seed=self._seed, seq_index_base=base)
                self.batches.append(len(batch))
                for r, o in zip(batch, out):
                    r.future.set_result(o)
            except (ValueError, RuntimeError, KeyError) as e:   # argument / prompt validation of the engine call
                if len(batch) == 1 or _is_device_failure(e):
                    for r in batch:                          # a CUDA error or OOM fails the whole batch: no serial re-runs
                        r.future.set_exception(e)
                else:
                    # one bad request (e.g. a prompt longer than the position table) must not fail its neighbours;
                    # request i keeps the Philox sequence index base + i it would have had inside the batch
                    for i, r in enumerate(batch):
                        try:
                            o = self.engine.generate([r.prompt], [r.max_new], temperature, top_k, eos_id=eos_id,
                                                     seed=self._seed, seq_index_base=base + i)[0]
                            r.future.set_result(o)
                        except (ValueError, RuntimeError, KeyError) as e1:
                            r.future.set_exception(e1)
                    self.batches.extend([1] * len(batch))
            except BaseException as e:                       # MemoryError, KeyboardInterrupt, ...: propagate to every caller
                for r in batch:
                    r.future.set_exception(e)
                if not isinstance(e, Exception):
                    raise


def _is_device_failure(e: BaseException) -> bool:
    """MG_E_CUDA (-4) / MG_E_OOM (-5) / MG_E_STATE (-6) as re-raised by engine._check: not a property of one request."""
    msg = str(e)
    return isinstance(e, MemoryError) or any(f"[mg status {c}]" in msg for c in (-4, -5, -6))


def stats(batcher: RequestBatcher) -> Dict[str, float]:
    n = len(batcher.batches)
    return {"engine_calls": n, "requests": sum(batcher.batches), "mean_batch": (sum(batcher.batches) / n) if n else 0.0}


class ContinuousBatcher:
    """Continuous batching in front of one engine (second half): a slot session of ``n_slots`` sequences.

    One worker thread alternates between (1) admitting waiting requests into free slots -- their prompts are prefilled into
    the K/V rows of the slot while everything in flight stays where it is --, (2) a chunk of ``chunk_steps`` decode steps for
    all slots in flight (the persistent cluster kernel, or the step graph for geometries it does not take), and (3) retiring
    the slots that hit ``[END_SEQUENCE]`` or their budget.  A request therefore waits at most one chunk (~ chunk_steps x 70 us)
    for a free slot instead of a whole batch generation, and the GPU never idles on a half-empty batch while requests wait.
    Sampling settings are per batcher (they are fixed for a slot session); every request has its own Philox stream
    (``seed``, running request index), so its tokens do not depend on the slot or the moment it was admitted.

        cb = ContinuousBatcher(model.engine, n_slots=64, max_len=1024, temperature=1.0, top_k=50, eos_id=eos)
        tokens = cb.generate(prompt_ids, max_new_tokens)          # blocking, thread-safe (the /generate endpoint's call)
        cb.close()
    """

    def __init__(self, engine, n_slots: int = 64, max_len: int = 1024, temperature: float = 1.0, top_k: Optional[int] = 50,
                 eos_id: int = -1, chunk_steps: int = 32, seed: Optional[int] = None, first_seq_index: int = 0):
        if n_slots < 1 or chunk_steps < 1:
            raise ValueError("n_slots and chunk_steps must be positive")
        self.engine, self.n_slots, self.max_len, self.chunk = engine, int(n_slots), int(max_len), int(chunk_steps)
        self._seed = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed)
        engine.slots_begin(self.n_slots, self.max_len, temperature, top_k, eos_id, self._seed)
        self._cv = threading.Condition()
        self._queue: List[Tuple[List[int], int, int, Future]] = []
        self._closed = False
        self._seq = itertools.count(int(first_seq_index))
        self._slot_req: List[Optional[Future]] = [None] * self.n_slots
        self.admissions: List[Tuple[int, int]] = []      # (chunk number, requests admitted) -- observability / tests
        self.chunks = 0
        self._worker = threading.Thread(target=self._run, name="mgea-continuous", daemon=True)
        self._worker.start()

    def submit(self, prompt_ids: Sequence[int], max_new_tokens: int) -> "Future[List[int]]":
        fut: "Future[List[int]]" = Future()
        with self._cv:
            if self._closed:
                raise RuntimeError("ContinuousBatcher is closed")
            self._queue.append((list(prompt_ids), max(int(max_new_tokens), 0), next(self._seq), fut))
            self._cv.notify()
        return fut

    def generate(self, prompt_ids: Sequence[int], max_new_tokens: int) -> List[int]:
        return self.submit(prompt_ids, max_new_tokens).result()

    def close(self) -> None:
        with self._cv:
            self._closed = True
            self._cv.notify()
        self._worker.join()
        self.engine.slots_end()

    # -- worker -----------------------------------------------------------------------------------------
    def _admit(self) -> None:
        free = [b for b in range(self.n_slots) if self._slot_req[b] is None]
        with self._cv:
            take, self._queue = self._queue[:len(free)], self._queue[len(free):]
        ok: List[Tuple[int, Tuple]] = []
        for req in take:
            prompt, max_new, _, fut = req
            if max_new == 0:                                    # api_cache.py:166: empty range, the prompt comes back unchanged
                fut.set_result(list(prompt))
            elif len(prompt) + max_new > self.max_len:
                fut.set_exception(ValueError(f"prompt + max_new_tokens = {len(prompt) + max_new} exceeds the batcher's max_len "
                                             f"{self.max_len}"))
            else:
                ok.append((free[len(ok)], req))
        if not ok:
            return

        def admit(items):
            self.engine.slots_admit([s for s, _ in items], [r[0] for _, r in items], [r[1] for _, r in items],
                                    [r[2] for _, r in items])
            for slot, r in items:
                self._slot_req[slot] = r[3]

        n_in = 0
        try:
            admit(ok)                                            # admission validates every prompt before it touches a slot
            n_in = len(ok)
        except Exception as e:                                   # noqa: BLE001
            for item in ok:                                      # a bad prompt fails its own request only
                if _is_device_failure(e):
                    item[1][3].set_exception(e)
                    continue
                try:
                    admit([item])
                    n_in += 1
                except Exception as e1:                          # noqa: BLE001
                    item[1][3].set_exception(e1)
        if n_in:
            self.admissions.append((self.chunks, n_in))

    def _run(self) -> None:
        while True:
            with self._cv:
                while not self._queue and not self._closed and all(f is None for f in self._slot_req):
                    self._cv.wait()
                if self._closed and not self._queue and all(f is None for f in self._slot_req):
                    return
            self._admit()
            if all(f is None for f in self._slot_req):
                continue
            try:
                fin, _ = self.engine.slots_step(self.chunk)
                self.chunks += 1
                done = [b for b in range(self.n_slots) if fin[b] and self._slot_req[b] is not None]
                if done:
                    fetch_many = getattr(self.engine, "slots_fetch_many", None)
                    rows = fetch_many(done) if fetch_many else [self.engine.slots_fetch(b, self.max_len + 8) for b in done]
                    for b, row in zip(done, rows):
                        self._slot_req[b].set_result(row)
                        self._slot_req[b] = None
            except Exception as e:                               # noqa: BLE001 -- device failure: nothing in flight survives
                for b in range(self.n_slots):
                    if self._slot_req[b] is not None:
                        self._slot_req[b].set_exception(e)
                        self._slot_req[b] = None
