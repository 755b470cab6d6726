"""Buffer plumbing: torch owns every device buffer; kernels get raw pointers + a stream."""
from __future__ import annotations

import ctypes as C
import warnings
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("pyperiod_b200 needs a CUDA device (sm_100a); there is no CPU path")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("pyperiod_b200 only computes on CUDA devices")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


@dataclass
class Windows:
    """A batch of windows resident on the device: window b = base[b*ldx : b*ldx + n]."""
    tensor: torch.Tensor   # keeps the storage alive; data_ptr() is window 0
    b: int
    n: int
    ldx: int
    was_1d: bool
    from_host: bool        # results go back to numpy
    # pipelined upload (stage_windows(..., pipeline=True)): generator factory of (first window, end window, event)
    plan: object = None

    def launch_plan(self):
        """Yields (b0, b1, event-or-None), one entry per kernel launch: windows [b0, b1) are complete on the
        device once `event` has fired.  With a pipelined upload the copy of the NEXT piece is issued just
        before a chunk is yielded, so it overlaps the kernel the caller launches for this chunk (also for
        pageable host memory, whose copies block the host thread)."""
        if self.plan is None:
            yield (0, self.b, None)
        else:
            yield from self.plan()

    @property
    def ptr(self) -> int:
        return self.tensor.data_ptr()

    @property
    def device(self) -> torch.device:
        return self.tensor.device


PIPELINE_MIN_WINDOWS = 16384   # below this a single upload + launch is cheaper than the extra launches
PIPELINE_PIECES = 8
_copy_streams: dict = {}


def pipeline_bounds(b: int, pieces: int = None) -> list:
    """[(first window, end window)] of the upload pieces of a b-window batch: a small first piece (the first kernel
    starts after ~1 ms of PCIe traffic), then `pieces` equal ones."""
    pieces = PIPELINE_PIECES if pieces is None else pieces
    per = -(-b // pieces)
    first = max(1, per // 4)
    return [(0, min(b, first))] + [(b0, min(b, b0 + per)) for b0 in range(first, b, per)]


def _pipelined_upload(host_flat: torch.Tensor, b: int, n: int, hop: int, dev: torch.device) -> Windows:
    """Upload a host sample stream in pieces on a side stream; window chunk k may be launched as soon as
    its piece has landed, so the PCIe copy of piece k+1 overlaps the kernel on piece k."""
    span = host_flat.numel()
    dflat = torch.empty((span,), dtype=torch.float64, device=dev)
    cur = torch.cuda.current_stream(dev)
    cs = _copy_streams.get(str(dev))
    if cs is None:
        cs = _copy_streams[str(dev)] = torch.cuda.Stream(dev)
    cs.wait_stream(cur)              # the allocation point of dflat is on the current stream
    dflat.record_stream(cs)
    bounds = pipeline_bounds(b)

    def copy_piece(k):
        b0, b1 = bounds[k]
        s0 = 0 if k == 0 else (bounds[k - 1][1] - 1) * hop + n
        s1 = (b1 - 1) * hop + n          # one past the last sample window b1-1 reads
        with torch.cuda.stream(cs):
            dflat[s0:s1].copy_(host_flat[s0:s1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        return ev

    def plan():
        ev = copy_piece(0)
        for k, (b0, b1) in enumerate(bounds):
            yield (b0, b1, ev)           # the caller launches the kernel for this chunk ...
            if k + 1 < len(bounds):
                ev = copy_piece(k + 1)   # ... and the next copy is issued behind it

    view = torch.as_strided(dflat, (b, n), (hop, 1))
    return Windows(view, b, n, hop, False, True, plan)


def stage_windows(data, device=None, pipeline: bool = False) -> Windows:
    """Accept a 1-D signal or a (B, N) batch (numpy / torch, host / device) and put it on the GPU.

    Overlapping row-strided views (stride0 < N, e.g. hop-512 frames of one stream) are uploaded
    once as the underlying stream and read in place by the kernels (ldx = hop).  With
    `pipeline=True` a large host batch is uploaded in pieces on a side stream and the returned
    Windows carries a launch plan (see Windows.launch_plan).
    """
    if pipeline:
        flat = None
        if isinstance(data, torch.Tensor) and data.device.type != "cuda" and data.dtype == torch.float64 \
                and data.dim() == 2 and data.shape[0] >= PIPELINE_MIN_WINDOWS and data.stride(1) == 1 \
                and 0 < data.stride(0) <= data.shape[1]:
            b, n = data.shape
            hop = data.stride(0)
            flat = torch.as_strided(data, ((b - 1) * hop + n,), (1,))
        elif isinstance(data, np.ndarray) and data.dtype == np.float64 and data.ndim == 2 \
                and data.shape[0] >= PIPELINE_MIN_WINDOWS and data.strides[1] == 8 and data.strides[0] % 8 == 0 \
                and 0 < data.strides[0] <= data.shape[1] * 8:
            b, n = data.shape
            hop = data.strides[0] // 8
            flat_np = np.lib.stride_tricks.as_strided(data, shape=((b - 1) * hop + n,), strides=(8,), writeable=False)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")   # read-only view: we only read from it
                flat = torch.from_numpy(flat_np)
        if flat is not None:
            return _pipelined_upload(flat, b, n, hop, require_cuda(device))
    if isinstance(data, torch.Tensor):
        t = data
        if t.dtype != torch.float64:
            t = t.to(torch.float64)
        was_1d = t.dim() == 1
        if was_1d:
            t = t.unsqueeze(0)
        if t.dim() != 2:
            raise ValueError("data must be 1-D or (B, N)")
        from_host = t.device.type != "cuda"
        dev = require_cuda(device if from_host else t.device)
        b, n = t.shape
        if b > 1 and t.stride(1) == 1 and 0 < t.stride(0) < n:       # overlapping frames
            span = (b - 1) * t.stride(0) + n
            flat = torch.as_strided(t, (span,), (1,))
            flat = flat.to(dev, non_blocking=True) if from_host else flat
            view = torch.as_strided(flat, (b, n), (t.stride(0), 1))
            return Windows(view, b, n, t.stride(0), was_1d, from_host)
        if t.stride(1) != 1 or (b > 1 and t.stride(0) < n):
            t = t.contiguous()
        if from_host:
            t = t.to(dev, non_blocking=True)
        return Windows(t, b, n, t.stride(0) if b > 1 else n, was_1d, from_host)

    arr = np.asarray(data)
    if arr.dtype != np.float64:
        arr = arr.astype(np.float64)
    was_1d = arr.ndim == 1
    if was_1d:
        arr = arr[None, :]
    if arr.ndim != 2:
        raise ValueError("data must be 1-D or (B, N)")
    dev = require_cuda(device)
    b, n = arr.shape
    s0, s1 = arr.strides
    if b > 1 and s1 == 8 and 0 < s0 < n * 8 and s0 % 8 == 0:          # overlapping frames
        hop = s0 // 8
        span = (b - 1) * hop + n
        flat = np.lib.stride_tricks.as_strided(arr, shape=(span,), strides=(8,), writeable=False)
        dflat = torch.from_numpy(np.ascontiguousarray(flat)).to(dev)
        view = torch.as_strided(dflat, (b, n), (hop, 1))
        return Windows(view, b, n, hop, was_1d, True)
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
    return Windows(t, b, n, n, was_1d, True)


class Workspace:
    """Growable scratch buffers handed to the library (caller-owned workspace), one per (device, stream, thread):
    a launch owns its workspace until the stream has run it, so calls issued on different CUDA streams or from
    different Python threads never share window counters, job tables or Cholesky factors."""
    _bufs: dict = {}
    # Buffers up to this size stay with their (device, stream, thread) key.  The factor storage of a large-dictionary
    # QO launch is 20 GB (296 CTAs x 68 MB): handing it back after every call made the next call pay for a fresh
    # allocation and a 20 GB memset -- 80 to 320 ms for the same 80 ms of kernels.  Workspace.release() frees them.
    KEEP_BYTES = 48 << 30

    @classmethod
    def get(cls, device: torch.device, nbytes: int) -> torch.Tensor:
        import threading
        stream = torch.cuda.current_stream(device)
        key = (str(device), int(stream.cuda_stream), threading.get_ident())
        buf = cls._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            cls._bufs.pop(key, None)   # give the smaller buffer back before asking for the larger one
            buf = None
            size = max(int(nbytes), 1 << 20)
            with torch.cuda.device(device):
                buf = torch.zeros(size, dtype=torch.uint8, device=device)   # once per buffer: it stays cached
            if buf.numel() <= cls.KEEP_BYTES:
                cls._bufs[key] = buf
        return buf

    @classmethod
    def release(cls):
        """Drop every cached workspace (the memory goes back to torch's caching allocator)."""
        cls._bufs.clear()


def stream_ptr(device: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def workspace_for(device: torch.device, size_fn, *args) -> torch.Tensor:
    """Workspace sized by one of the library's *_workspace_bytes queries (which read the current device)."""
    with torch.cuda.device(device):
        nbytes = int(size_fn(*args))
    return Workspace.get(device, nbytes)


def call(fn, what: str, device: torch.device, *args):
    """Run one C-ABI entry point with `device` current: the library sizes its grids, sets kernel attributes and
    launches on the CUDA runtime's current device, which must be the one that owns the pointers and the stream."""
    with torch.cuda.device(device):
        _lib.check(fn(*args), what)


def ptr(t) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


PINNED_D2H_MIN_BYTES = 1 << 20
STAGE_BYTES = 64 << 20      # size of each of the two page-locked staging buffers of a pipelined download
COPY_THREADS = 4            # host threads that move a staged chunk into the caller's array (first touch of fresh pages)
COPY_THREADS_MIN_BYTES = 8 << 20
_stage_bufs: dict = {}
_down_streams: dict = {}
_stage_locks: dict = {}
_copy_pool = None


def _host_copy(dst: np.ndarray, src: np.ndarray):
    """dst[:] = src on COPY_THREADS threads (numpy releases the GIL while it copies; one thread moves ~5 GB/s into
    freshly allocated pages, which made the host side of a 2.6 GB download longer than the kernels it hides behind)."""
    global _copy_pool
    n = dst.shape[0]
    if n < COPY_THREADS_MIN_BYTES or COPY_THREADS < 2:
        np.copyto(dst, src)
        return
    if _copy_pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _copy_pool = ThreadPoolExecutor(COPY_THREADS)
    step = -(-n // COPY_THREADS)
    step = (step + 4095) & ~4095
    futs = [_copy_pool.submit(np.copyto, dst[o:o + step], src[o:o + step]) for o in range(0, n, step)]
    for f in futs:
        f.result()


class RowDownloader:
    """Pipelined device -> host copy of the rows of large result arrays, piece by piece while the next piece is still
    being computed.  Rows go through two page-locked staging buffers (allocated once per device: a fresh page-locked
    allocation of a 2 GB result costs 0.9 s, its copy 40 ms) into ordinary numpy arrays the caller owns.

        d = RowDownloader(device, {"res": (dev_tensor_2d, host_array_2d), ...})
        ... launch the kernel for rows [b0, b1) ...;  d.mark(b0, b1)       # records an event behind the launch
        ... launch the next piece ...;                d.drain()            # copies every marked piece but the last
        d.finish()                                                         # copies what is left
    """

    def __init__(self, device: torch.device, pairs: dict):
        self.dev = device
        self.pairs = {k: v for k, v in pairs.items() if v[0] is not None}
        key = str(device)
        import threading
        if key not in _stage_bufs:
            _stage_locks.setdefault(key, threading.Lock())
            with _stage_locks[key]:
                if key not in _stage_bufs:
                    _down_streams[key] = torch.cuda.Stream(device)
                    _stage_bufs[key] = [torch.empty(STAGE_BYTES, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        self.stage = _stage_bufs[key]
        self.cs = _down_streams[key]
        self.lock = _stage_locks[key]   # the staging buffers of a device are shared by every thread that uses it
        self.pending = []   # (b0, b1, event)

    def mark(self, b0: int, b1: int):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        self.pending.append((b0, b1, ev))

    def _copy(self, b0: int, b1: int, ev):
        with self.lock:
            self._copy_locked(b0, b1, ev)

    def _copy_locked(self, b0: int, b1: int, ev):
        self.cs.wait_event(ev)
        for src, dst in self.pairs.values():
            s = src[b0:b1].reshape(-1).view(torch.uint8)
            d = dst[b0:b1].reshape(-1).view(np.uint8)
            n = s.numel()
            inflight = [None, None]
            k = 0
            for o in range(0, n, STAGE_BYTES):
                m = min(STAGE_BYTES, n - o)
                b = k & 1
                if inflight[b] is not None:   # this staging buffer still holds the chunk before last: move it out
                    e2, po, pm = inflight[b]
                    e2.synchronize()
                    _host_copy(d[po:po + pm], self.stage[b][:pm].numpy())
                with torch.cuda.stream(self.cs):
                    self.stage[b][:m].copy_(s[o:o + m], non_blocking=True)
                    e2 = torch.cuda.Event()
                    e2.record(self.cs)
                inflight[b] = (e2, o, m)
                k += 1
            for b in ((k & 1), ((k + 1) & 1)):   # oldest first
                if inflight[b] is not None:
                    e2, po, pm = inflight[b]
                    e2.synchronize()
                    _host_copy(d[po:po + pm], self.stage[b][:pm].numpy())

    def drain(self):
        """Copy every marked piece except the most recent one (whose kernel is still running)."""
        while len(self.pending) > 1:
            self._copy(*self.pending.pop(0))

    def finish(self):
        while self.pending:
            self._copy(*self.pending.pop(0))


STAGED_D2H_MIN_BYTES = 256 << 20


def to_host(t: torch.Tensor) -> np.ndarray:
    """Device tensor -> numpy.  Mid-sized results travel through a page-locked buffer of their own (torch's caching
    host allocator reuses the blocks; a pageable `.cpu()` runs at a quarter of the PCIe rate); very large ones (the
    2 GB of residuals of a 65,536-window batch) through the two staging buffers of RowDownloader into an ordinary
    numpy array: a fresh page-locked allocation of that size costs 0.9 s, more than the copy and the kernels."""
    nbytes = t.numel() * t.element_size()
    if t.device.type != "cuda" or nbytes < PINNED_D2H_MIN_BYTES:
        return t.cpu().numpy()
    src = t if t.is_contiguous() else t.contiguous()
    if nbytes >= STAGED_D2H_MIN_BYTES:
        dst = np.empty(tuple(src.shape), dtype=np.dtype(str(src.dtype).replace("torch.", "")))
        d = RowDownloader(src.device, {"t": (src.reshape(1, -1), dst.reshape(1, -1))})
        d.mark(0, 1)
        d.finish()
        return dst
    host = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
    host.copy_(src, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return host.numpy()
