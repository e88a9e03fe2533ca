"""Pinned host-memory pools for the drop-in API (interpolator.py / physics.py).

``cudaHostAlloc`` of a result grid costs seconds at 1024^3 (12.9 GB), so the buffers the NumPy-facing
functions need are kept and reused:

* ``stage_to_device``: pageable NumPy array -> cached pinned staging chunk(s) -> device tensor.  The host
  copy into the pinned chunk runs on torch's CPU thread pool and overlaps the DMA of the previous chunk
  (two chunks in flight).
* ``ResultPool``: pinned result buffers keyed by (shape, dtype).  A buffer is handed out as a NumPy array;
  it returns to the pool when the caller has dropped every view of it (weakref finaliser), so results a
  caller keeps are never overwritten -- the next call then simply allocates another buffer.
"""
from __future__ import annotations

import threading
import warnings
import weakref

import numpy as np
import torch

_CHUNK_BYTES = 64 << 20
_lock = threading.Lock()
_staging = {}  # device index -> [pinned uint8 tensor, pinned uint8 tensor, events]


def _staging_for(dev):
    key = dev.index
    with _lock:
        if key not in _staging:
            bufs = [torch.empty(_CHUNK_BYTES, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
            evs = [torch.cuda.Event(), torch.cuda.Event()]
            _staging[key] = (bufs, evs)
        return _staging[key]


def stage_to_device(arr, dev, out=None, stream=None):
    """Copy a C-contiguous NumPy array to a (new or given) device tensor of the same shape/dtype through
    the pinned staging chunks; the DMAs run on ``stream`` (default: the current stream).  Returns when the
    host array has been fully read and the last DMA has completed."""
    arr = np.ascontiguousarray(arr)
    with warnings.catch_warnings():  # read-only inputs (broadcast views, memory maps) are only read here
        warnings.simplefilter("ignore", UserWarning)
        t_host = torch.from_numpy(arr.view(np.uint8).reshape(-1)) if arr.dtype == np.bool_ else \
            torch.from_numpy(arr.reshape(-1).view(np.uint8))
    nbytes = t_host.numel()
    tdt = torch.from_numpy(np.empty(0, dtype=np.uint8 if arr.dtype == np.bool_ else arr.dtype)).dtype
    if out is None:
        out = torch.empty(arr.shape, dtype=tdt, device=dev)
    dflat = out.reshape(-1).view(torch.uint8)
    if nbytes == 0:
        return out
    bufs, evs = _staging_for(dev)
    if stream is None:
        stream = torch.cuda.current_stream(dev)
    used = [False, False]
    for i, off in enumerate(range(0, nbytes, _CHUNK_BYTES)):
        b = i & 1
        n = min(_CHUNK_BYTES, nbytes - off)
        if used[b]:
            evs[b].synchronize()  # the DMA that last read this chunk has finished
        bufs[b][:n].copy_(t_host[off:off + n])  # host -> pinned (torch's CPU thread pool)
        with torch.cuda.stream(stream):
            dflat[off:off + n].copy_(bufs[b][:n], non_blocking=True)
        evs[b].record(stream)
        used[b] = True
    for b in range(2):  # the chunks are shared by later calls on other streams
        if used[b]:
            evs[b].synchronize()
    return out


class ResultPool:
    """Pinned result buffers that come back when the caller lets go of them."""

    def __init__(self):
        self._free = {}
        self._lock = threading.Lock()

    def take(self, shape, dtype):
        key = (tuple(shape), dtype)
        with self._lock:
            lst = self._free.get(key)
            t = lst.pop() if lst else None
        if t is None:
            t = torch.empty(shape, dtype=dtype, pin_memory=True)
        return t

    def as_numpy(self, t):
        """NumPy view of a pool tensor; the tensor returns to the pool when the view (and every array
        derived from it) has been garbage-collected."""
        host = t.numpy()
        weakref.finalize(host, self._give_back, (tuple(t.shape), t.dtype), t)
        return host

    def _give_back(self, key, t):
        with self._lock:
            lst = self._free.setdefault(key, [])
            if len(lst) < 2:  # do not hoard pinned memory
                lst.append(t)

    def clear(self):
        with self._lock:
            self._free.clear()


results = ResultPool()
