"""Pinned host-memory pools for the drop-in API (interpolator.py / physics.py).

``cudaHostAlloc`` of a result grid costs seconds at 1024^3 (12.9 GB), so the buffers the NumPy-facing
functions need are kept and reused:

* ``stage_to_device``: pageable NumPy array -> cached pinned staging chunk(s) -> device tensor.  The host
  copy into the pinned chunk is split over a few threads and overlaps the DMA of the previous chunk
  (two chunks in flight).
* ``ResultPool``: pinned result buffers keyed by (shape, dtype).  A buffer is handed out as a NumPy array;
  it returns to the pool when the caller has dropped every view of it (weakref finaliser), so results a
  caller keeps are never overwritten -- the next call then simply allocates another buffer.
"""
from __future__ import annotations

import os
import threading
import warnings
import weakref

import numpy as np
import torch

_CHUNK_BYTES = 32 << 20
_lock = threading.Lock()
_pool = None


def _copy_pool():
    """A few host threads for the pageable -> pinned copies (np.copyto releases the GIL).  Independent of
    OMP_NUM_THREADS, which torchrun pins to 1; sized so that the ranks of one node share the cores."""
    global _pool
    if _pool is None:
        from concurrent.futures import ThreadPoolExecutor
        ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
        _pool = ThreadPoolExecutor(max_workers=max(1, min(8, (os.cpu_count() or 1) // ranks)))
    return _pool


def _parallel_copy(dst, src):
    """dst[:] = src for two 1-D uint8 NumPy arrays, split over the copy threads."""
    n = dst.shape[0]
    pool = _copy_pool()
    parts = pool._max_workers
    if parts == 1 or n < (4 << 20):
        np.copyto(dst, src)
        return
    step = -(-n // parts)
    list(pool.map(lambda a: np.copyto(dst[a:a + step], src[a:a + step]), range(0, n, step)))
_staging = {}  # device index -> [pinned uint8 tensor, pinned uint8 tensor, events]


def _staging_for(dev):
    key = dev.index
    with _lock:
        if key not in _staging:
            bufs = [torch.empty(_CHUNK_BYTES, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
            evs = [torch.cuda.Event(), torch.cuda.Event()]
            _staging[key] = (bufs, evs)
        return _staging[key]


def _as_bytes(arr):
    arr = np.ascontiguousarray(arr)
    flat = arr.view(np.uint8).reshape(-1) if arr.dtype == np.bool_ else arr.reshape(-1).view(np.uint8)
    return arr, flat


def stage_many_to_device(pairs, dev, stream=None):
    """``pairs``: (C-contiguous NumPy array, device tensor of the same byte size) -- all of them go through the two
    pinned staging chunks as ONE stream of chunks: the host copy of a chunk overlaps the DMA of the previous one
    across array boundaries, and the host waits for the DMAs only once, at the end (six particle columns of 80 MB:
    27 ms one by one, each with its own tail -> about half)."""
    bufs, evs = _staging_for(dev)
    if stream is None:
        stream = torch.cuda.current_stream(dev)
    used = [False, False]
    i = 0
    for arr, out in pairs:
        _, h_flat = _as_bytes(arr)
        nbytes = h_flat.shape[0]
        dflat = out.reshape(-1).view(torch.uint8)
        if dflat.numel() != nbytes:
            raise ValueError("stage_many_to_device: size mismatch")
        for off in range(0, nbytes, _CHUNK_BYTES):
            b = i & 1
            i += 1
            n = min(_CHUNK_BYTES, nbytes - off)
            if used[b]:
                evs[b].synchronize()  # the DMA that last read this chunk has finished
            _parallel_copy(bufs[b].numpy()[:n], h_flat[off:off + n])  # host -> pinned, a few threads
            with torch.cuda.stream(stream):
                dflat[off:off + n].copy_(bufs[b][:n], non_blocking=True)
            evs[b].record(stream)
            used[b] = True
    for b in range(2):  # the chunks are shared by later calls on other streams
        if used[b]:
            evs[b].synchronize()


def stage_to_device(arr, dev, out=None, stream=None):
    """Copy a C-contiguous NumPy array to a (new or given) device tensor of the same shape/dtype through
    the pinned staging chunks; the DMAs run on ``stream`` (default: the current stream).  Returns when the
    host array has been fully read and the last DMA has completed."""
    arr, h_flat = _as_bytes(arr)
    tdt = torch.from_numpy(np.empty(0, dtype=np.uint8 if arr.dtype == np.bool_ else arr.dtype)).dtype
    if out is None:
        out = torch.empty(arr.shape, dtype=tdt, device=dev)
    if h_flat.shape[0] == 0:
        return out
    stage_many_to_device([(arr, out)], dev, stream=stream)
    return out


def bind_host_to_device(device_index):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off (sysfs ``local_cpulist`` of its PCI
    device), so that the pinned buffers allocated afterwards are node-local and host<->device DMA does not
    cross the socket interconnect.  One process per GPU; call before the first pinned allocation.  Returns a
    dict describing what was done ({} if the topology cannot be read -- then nothing changes)."""
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        base = "/sys/bus/pci/devices/" + bus
        with open(base + "/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return {"pci": bus, "cpus": len(allowed), "changed": False}
        os.sched_setaffinity(0, cpus)
        node = None
        try:
            with open(base + "/numa_node") as f:
                node = int(f.read().strip())
        except Exception:
            pass
        return {"pci": bus, "numa_node": node, "cpus": len(cpus), "changed": True}
    except Exception:
        return {}


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
