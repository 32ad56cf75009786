"""DeviceIndex -- one row shard resident in the HBM of one B200, driven through the C ABI.

Host buffers are numpy arrays; device buffers are torch CUDA tensors passed by ``data_ptr()``
(torch is plumbing for device memory and streams only -- every kernel is in
``libvecsearch_b200.so``).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _native as N

_DTYPES = {"f32": N.VS_F32, "fp32": N.VS_F32, "float32": N.VS_F32,
           "bf16": N.VS_BF16, "bfloat16": N.VS_BF16}
_MODES = {"auto": N.VS_Q_AUTO, "scan": N.VS_Q_SCAN, "tensor": N.VS_Q_TENSOR}


def _mode(mode: str, pipelined: bool = False) -> int:
    """``pipelined``: the caller vouches that the query buffer was complete before the previous launch
    on the stream (VS_Q_PIPELINED, see include/vecsearch_b200.h)."""
    return _MODES[mode] | (N.VS_Q_PIPELINED if pipelined else 0)


def bits_to_words(bits_lists: Sequence[Sequence[int]]) -> np.ndarray:
    """Per-row filter-bit index lists -> uint64 [n, MASK_WORDS]."""
    words = np.zeros((len(bits_lists), N.MASK_WORDS), dtype=np.uint64)
    for j, bits in enumerate(bits_lists):
        for b in bits:
            if not 0 <= b < 64 * N.MASK_WORDS:
                raise ValueError(f"filter bit {b} out of range [0,{64 * N.MASK_WORDS})")
            words[j, b // 64] |= np.uint64(1 << (b % 64))
    return words


def _bits_array(bits: Optional[Sequence[int]]):
    """iterable of filter-bit indices -> (uint64 * 4) or None."""
    if bits is None:
        return None
    words = [0] * N.MASK_WORDS
    for b in bits:
        if not 0 <= b < 64 * N.MASK_WORDS:
            raise ValueError(f"filter bit {b} out of range [0,{64 * N.MASK_WORDS})")
        words[b // 64] |= 1 << (b % 64)
    return (C.c_uint64 * N.MASK_WORDS)(*words)


def _ptr(t) -> int:
    return 0 if t is None else int(t.data_ptr())


_CUDA_STREAM_LEGACY = 1   # cudaStreamLegacy: the C ABI reserves NULL for "the index's own stream"


def _stream_ptr(stream) -> int:
    """torch stream / raw handle -> cudaStream_t value for the C ABI.  torch's default stream
    has handle 0, which the ABI would read as "use the index's own stream"; pass the explicit
    legacy-default-stream handle instead so kernels really run on the caller's stream."""
    if stream is None:
        return 0
    h = int(getattr(stream, "cuda_stream", stream))
    return h if h != 0 else _CUDA_STREAM_LEGACY


class DeviceIndex:
    def __init__(self, dim: int, dtype: str = "f32", device: int = 0, capacity: int = 0, row_base: int = 0,
                 row_stride: int = 1, _handle=None):
        self._lib = N.load()
        if dtype not in _DTYPES:
            raise ValueError(f"dtype must be one of {sorted(_DTYPES)}")
        self.dim, self.dtype, self.device = int(dim), "bf16" if _DTYPES[dtype] else "f32", int(device)
        self._owned = _handle is None
        if _handle is not None:                       # a shard of a vs_group_t: the group owns it
            self._h = C.c_void_p(_handle)
            self.device = int(self._lib.vs_device(self._h))
            return
        h = C.c_void_p()
        N.check(self._lib.vs_create(self.device, self.dim, _DTYPES[dtype], int(capacity), C.byref(h)))
        self._h = h
        if row_base or row_stride != 1:
            self.set_row_map(row_base, row_stride)

    # -- lifecycle ---------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            if self._owned:
                self._lib.vs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(self._lib.vs_count(self._h))

    count = __len__

    @property
    def sm_count(self) -> int:
        return int(self._lib.vs_device_sm_count(self._h))

    @property
    def last_query_path(self) -> str:
        return {N.VS_Q_SCAN: "scan", N.VS_Q_TENSOR: "tensor"}.get(int(self._lib.vs_last_query_path(self._h)), "none")

    def set_row_base(self, row_base: int):
        N.check(self._lib.vs_set_row_base(self._h, int(row_base)))

    def set_row_map(self, row_base: int, row_stride: int = 1):
        """Reported row = row_base + local row * row_stride (striped shards: (shard, G))."""
        N.check(self._lib.vs_set_row_map(self._h, int(row_base), int(row_stride)))

    # -- ingest / maintenance ------------------------------------------------------------
    def add(self, rows) -> int:
        """Append rows ([n, dim] float32; numpy -> host path, torch CUDA tensor -> device path).
        Returns the local row number of the first appended row."""
        first = C.c_int64(-1)
        if isinstance(rows, np.ndarray) or not hasattr(rows, "data_ptr"):
            a = np.ascontiguousarray(rows, dtype=np.float32)
            if a.ndim == 1:
                a = a[None]
            if a.ndim != 2 or a.shape[1] != self.dim:
                raise ValueError(f"expected [n,{self.dim}] rows, got {a.shape}")
            N.check(self._lib.vs_add_host(self._h, a.ctypes.data, a.shape[0], C.byref(first)))
        else:
            import torch
            t = rows
            if not t.is_cuda or t.dtype != torch.float32 or t.dim() != 2 or t.shape[1] != self.dim:
                raise ValueError("device rows must be a CUDA float32 [n, dim] tensor")
            t = t.contiguous()
            st = torch.cuda.current_stream(t.device)
            N.check(self._lib.vs_add_dev(self._h, _ptr(t), t.shape[0], C.byref(first), _stream_ptr(st)))
        return int(first.value)

    def remove(self, row: int) -> int:
        """Remove a row (last row is moved into its place).  Returns the moved row or -1."""
        moved = C.c_int64(-1)
        N.check(self._lib.vs_remove(self._h, int(row), C.byref(moved)))
        return int(moved.value)

    def remove_rows(self, rows) -> Tuple[np.ndarray, np.ndarray]:
        """Remove many rows with ONE compaction kernel (vs_remove_rows).  Returns (moved_src, moved_dst):
        surviving row ``moved_src[i]`` now lives at ``moved_dst[i]``; the new count is ``count - len(rows)``."""
        r = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1)
        m = int(r.shape[0])
        src = np.empty(max(m, 1), dtype=np.int64)
        dst = np.empty(max(m, 1), dtype=np.int64)
        nm = C.c_int64(0)
        N.check(self._lib.vs_remove_rows(self._h, r.ctypes.data, m, src.ctypes.data, dst.ctypes.data, C.byref(nm)))
        return src[:nm.value].copy(), dst[:nm.value].copy()

    def truncate(self, new_count: int):
        N.check(self._lib.vs_truncate(self._h, int(new_count)))

    def reserve(self, capacity: int):
        N.check(self._lib.vs_reserve(self._h, int(capacity)))

    def copy_row_from(self, dst_row: int, src: "DeviceIndex", src_row: int):
        """dst row <- src row (vector, inverse norm, filter bits); the two shards may sit on different GPUs."""
        N.check(self._lib.vs_copy_row(self._h, int(dst_row), src._h, int(src_row)))

    def move_rows_from(self, src: "DeviceIndex", src_rows, dst_rows):
        """Batched ``copy_row_from``: one kernel on this index's GPU (vs_move_rows)."""
        a = np.ascontiguousarray(src_rows, dtype=np.int64).reshape(-1)
        b = np.ascontiguousarray(dst_rows, dtype=np.int64).reshape(-1)
        if a.shape != b.shape:
            raise ValueError("src_rows and dst_rows differ in length")
        if a.shape[0]:
            N.check(self._lib.vs_move_rows(self._h, src._h, a.ctypes.data, b.ctypes.data, int(a.shape[0])))

    def replicate_from(self, src: "DeviceIndex", dst_first: int, dst_stride: int):
        """Row ``dst_first + l*dst_stride`` of this index <- row l of ``src`` for all its rows (NVLink peer reads)."""
        N.check(self._lib.vs_replicate_from(self._h, src._h, int(dst_first), int(dst_stride)))

    @property
    def storage_dtype(self):
        return np.dtype(np.uint16) if self.dtype == "bf16" else np.dtype(np.float32)

    def get_raw(self, first: int, n: int) -> np.ndarray:
        """Stored rows [first, first+n) in the storage dtype (bf16 as uint16 bit patterns): the persistence slab."""
        out = np.empty((n, self.dim), dtype=self.storage_dtype)
        N.check(self._lib.vs_get_raw_host(self._h, int(first), int(n), out.ctypes.data))
        return out

    def add_raw(self, stored_rows: np.ndarray) -> int:
        """Append rows that are already in the storage dtype (chunked pinned upload, norms recomputed on device)."""
        a = np.ascontiguousarray(stored_rows)
        if a.dtype != self.storage_dtype or a.ndim != 2 or a.shape[1] != self.dim:
            raise ValueError(f"expected [n,{self.dim}] rows of dtype {self.storage_dtype}, got {a.dtype} {a.shape}")
        first = C.c_int64(-1)
        N.check(self._lib.vs_add_raw_host(self._h, a.ctypes.data, a.shape[0], C.byref(first)))
        return int(first.value)

    def set_row(self, row: int, vec):
        """Overwrite the vector of an existing row in place."""
        a = np.ascontiguousarray(vec, dtype=np.float32).reshape(-1)
        if a.shape[0] != self.dim:
            raise ValueError(f"expected a [{self.dim}] vector")
        N.check(self._lib.vs_set_row_host(self._h, int(row), a.ctypes.data))

    def clear(self):
        N.check(self._lib.vs_clear(self._h))

    def set_filter_bits(self, row: int, bits: Sequence[int]):
        N.check(self._lib.vs_set_mask_bits(self._h, int(row), _bits_array(bits)))

    def set_filter_bits_range(self, first: int, bits_lists: Sequence[Sequence[int]]):
        """Bulk form: filter-bit indices of rows [first, first + len(bits_lists)) in one copy."""
        n = len(bits_lists)
        if n == 0:
            return
        self.set_filter_words_range(first, bits_to_words(bits_lists))

    def set_filter_words_range(self, first: int, words: np.ndarray):
        """Raw form: uint64 [n, MASK_WORDS] bit words of rows [first, first + n)."""
        w = np.ascontiguousarray(words, dtype=np.uint64).reshape(-1, N.MASK_WORDS)
        if w.shape[0]:
            N.check(self._lib.vs_set_mask_bits_range(self._h, int(first), int(w.shape[0]), w.ctypes.data))

    def get_filter_words_range(self, first: int, n: int) -> np.ndarray:
        out = np.zeros((n, N.MASK_WORDS), dtype=np.uint64)
        if n:
            N.check(self._lib.vs_get_mask_bits_range(self._h, int(first), int(n), out.ctypes.data))
        return out

    def apply_sweep_bits_dev(self, words_dev, bit: int, stream=None):
        """Filter bit ``bit`` of every row := the row's bit in one filter's sweep output (stays on the GPU)."""
        import torch
        st = stream if stream is not None else torch.cuda.current_stream(words_dev.device)
        N.check(self._lib.vs_apply_sweep_bits_dev(self._h, _ptr(words_dev), int(bit), _stream_ptr(st)))

    def get_filter_bits(self, row: int):
        w = (C.c_uint64 * N.MASK_WORDS)()
        N.check(self._lib.vs_get_mask_bits(self._h, int(row), w))
        return [b for b in range(64 * N.MASK_WORDS) if (w[b // 64] >> (b % 64)) & 1]

    def get_rows(self, first: int, n: int) -> np.ndarray:
        out = np.empty((n, self.dim), dtype=np.float32)
        N.check(self._lib.vs_get_rows_host(self._h, int(first), int(n), out.ctypes.data))
        return out

    def get_rows_dev(self, first: int, n: int, out=None, stream=None):
        """Stored rows [first, first+n) as a float32 CUDA tensor [n, dim] (exact for bf16 storage)."""
        import torch
        dev = torch.device("cuda", self.device)
        if out is None:
            out = torch.empty((n, self.dim), dtype=torch.float32, device=dev)
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        N.check(self._lib.vs_get_rows_dev(self._h, int(first), int(n), _ptr(out), _stream_ptr(st)))
        return out

    # -- query ----------------------------------------------------------------------------
    def query(self, q, k: int, require_bits: Optional[Sequence[int]] = None, mode: str = "auto"
              ) -> Tuple[np.ndarray, np.ndarray]:
        """HOST-buffer query (H2D + kernels + D2H inside): q [B, dim] float32 numpy ->
        (scores [B,k] float32, rows [B,k] int64); empty slots are (-inf, -1)."""
        a = np.ascontiguousarray(q, dtype=np.float32)
        if a.ndim == 1:
            a = a[None]
        if a.ndim != 2 or a.shape[1] != self.dim:
            raise ValueError(f"expected [B,{self.dim}] queries, got {a.shape}")
        B = a.shape[0]
        s = np.empty((B, k), dtype=np.float32)
        r = np.empty((B, k), dtype=np.int64)
        N.check(self._lib.vs_query_topk_host(self._h, a.ctypes.data, B, int(k), _bits_array(require_bits),
                                             _MODES[mode], s.ctypes.data, r.ctypes.data))
        return s, r

    def query_dev(self, q, k: int, out_scores=None, out_rows=None, require_bits: Optional[Sequence[int]] = None,
                  mode: str = "auto", stream=None, pipelined: bool = False):
        """DEVICE-buffer query, asynchronous on ``stream`` (default: torch's current stream)."""
        import torch
        if q.dim() == 1:
            q = q[None]
        if not q.is_cuda or q.dtype != torch.float32 or q.shape[1] != self.dim:
            raise ValueError("queries must be a CUDA float32 [B, dim] tensor")
        q = q.contiguous()
        B = q.shape[0]
        if out_scores is None:
            out_scores = torch.empty((B, k), dtype=torch.float32, device=q.device)
        if out_rows is None:
            out_rows = torch.empty((B, k), dtype=torch.int64, device=q.device)
        st = stream if stream is not None else torch.cuda.current_stream(q.device)
        N.check(self._lib.vs_query_topk_dev(self._h, _ptr(q), B, int(k), _bits_array(require_bits),
                                            _mode(mode, pipelined), _ptr(out_scores), _ptr(out_rows), _stream_ptr(st)))
        return out_scores, out_rows

    def query_multimodal(self, img, txt, w, k: int, require_bits: Optional[Sequence[int]] = None,
                         mode: str = "auto") -> Tuple[np.ndarray, np.ndarray]:
        """HOST-buffer multimodal query: blend (backend/app/main.py:850-860) on the device, then
        top-k.  img/txt [B, dim] float32, w [B] (python floats / float64)."""
        a = np.ascontiguousarray(img, dtype=np.float32)
        t = np.ascontiguousarray(txt, dtype=np.float32)
        if a.ndim == 1:
            a, t = a[None], t[None]
        ww = np.ascontiguousarray(np.broadcast_to(np.asarray(w, dtype=np.float64), (a.shape[0],)))
        if a.shape != t.shape or a.shape[1] != self.dim:
            raise ValueError("img/txt must both be [B, dim]")
        B = a.shape[0]
        s = np.empty((B, k), dtype=np.float32)
        r = np.empty((B, k), dtype=np.int64)
        N.check(self._lib.vs_query_multimodal_host(self._h, a.ctypes.data, t.ctypes.data, ww.ctypes.data, B, int(k),
                                                   _bits_array(require_bits), _MODES[mode], s.ctypes.data,
                                                   r.ctypes.data))
        return s, r

    def blend_dev(self, img, txt, w, out=None, stream=None):
        """search_multimodal's blend (backend/app/main.py:850-860) for B (img, txt, w) triples on
        the device; ``w`` is a float64 CUDA tensor [B]."""
        import torch
        B = img.shape[0]
        if out is None:
            out = torch.empty((B, self.dim), dtype=torch.float32, device=img.device)
        st = stream if stream is not None else torch.cuda.current_stream(img.device)
        N.check(self._lib.vs_blend_dev(self._h, _ptr(img.contiguous()), _ptr(txt.contiguous()),
                                       _ptr(w.to(torch.float64).contiguous()), B, _ptr(out), _stream_ptr(st)))
        return out

    def merge_dev(self, cand_scores, cand_rows, out_scores=None, out_rows=None, stream=None):
        """[G,B,k] gathered candidates -> [B,k] (runs on the candidates' device)."""
        import torch
        G, B, k = cand_scores.shape
        if out_scores is None:
            out_scores = torch.empty((B, k), dtype=torch.float32, device=cand_scores.device)
        if out_rows is None:
            out_rows = torch.empty((B, k), dtype=torch.int64, device=cand_scores.device)
        st = stream if stream is not None else torch.cuda.current_stream(cand_scores.device)
        N.check(self._lib.vs_merge_topk_dev(self._h, _ptr(cand_scores.contiguous()), _ptr(cand_rows.contiguous()),
                                            G, B, k, _ptr(out_scores), _ptr(out_rows), _stream_ptr(st)))
        return out_scores, out_rows

    # -- peer exchange (row-sharded collection) --------------------------
    def exchange_create(self, world_size: int, rank: int, b_max: int = 1024, k_max: int = 32) -> bytes:
        """Allocate this shard's exchange buffer; returns the 64-byte CUDA IPC handle to publish."""
        N.check(self._lib.vs_exchange_create(self._h, int(world_size), int(rank), int(b_max), int(k_max)))
        h = (C.c_ubyte * 64)()
        N.check(self._lib.vs_exchange_ipc_handle(self._h, h))
        return bytes(h)

    def exchange_local_ptr(self) -> int:
        return int(self._lib.vs_exchange_local_ptr(self._h) or 0)

    def exchange_attach(self, ipc_handles: Optional[Sequence[bytes]] = None,
                        peer_ptrs: Optional[Sequence[int]] = None):
        """Map the peers' buffers: ``ipc_handles[g]`` (other processes) or ``peer_ptrs[g]`` (same process)."""
        hb = None
        if ipc_handles is not None:
            blob = b"".join(bytes(h) if h else b"\0" * 64 for h in ipc_handles)
            hb = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        pp = None
        if peer_ptrs is not None:
            pp = (C.c_void_p * len(peer_ptrs))(*[C.c_void_p(int(p)) if p else C.c_void_p(None) for p in peer_ptrs])
        N.check(self._lib.vs_exchange_attach(self._h, hb, pp))

    def query_sharded_dev(self, q, k: int, out_scores=None, out_rows=None,
                          require_bits: Optional[Sequence[int]] = None, mode: str = "auto", stream=None):
        """Like ``query_dev`` but returns the GLOBAL top-k over all shards on every rank: the exchange
        over NVLink peer memory is fused into the query kernel (vs_query_topk_sharded_dev)."""
        import torch
        if q.dim() == 1:
            q = q[None]
        if not q.is_cuda or q.dtype != torch.float32 or q.shape[1] != self.dim:
            raise ValueError("queries must be a CUDA float32 [B, dim] tensor")
        q = q.contiguous()
        B = q.shape[0]
        if out_scores is None:
            out_scores = torch.empty((B, k), dtype=torch.float32, device=q.device)
        if out_rows is None:
            out_rows = torch.empty((B, k), dtype=torch.int64, device=q.device)
        st = stream if stream is not None else torch.cuda.current_stream(q.device)
        N.check(self._lib.vs_query_topk_sharded_dev(self._h, _ptr(q), B, int(k), _bits_array(require_bits),
                                                    _MODES[mode], _ptr(out_scores), _ptr(out_rows), _stream_ptr(st)))
        return out_scores, out_rows

    def query_sharded(self, q, k: int, require_bits: Optional[Sequence[int]] = None, mode: str = "auto"
                      ) -> Tuple[np.ndarray, np.ndarray]:
        """HOST-buffer sharded query (vs_query_topk_sharded_host): H2D, kernel + fused exchange, D2H."""
        a = np.ascontiguousarray(q, dtype=np.float32)
        if a.ndim == 1:
            a = a[None]
        if a.ndim != 2 or a.shape[1] != self.dim:
            raise ValueError(f"expected [B,{self.dim}] queries, got {a.shape}")
        B = a.shape[0]
        s = np.empty((B, k), dtype=np.float32)
        r = np.empty((B, k), dtype=np.int64)
        N.check(self._lib.vs_query_topk_sharded_host(self._h, a.ctypes.data, B, int(k), _bits_array(require_bits),
                                                     _MODES[mode], s.ctypes.data, r.ctypes.data))
        return s, r

    def exchange_merge_dev(self, cand_scores, cand_rows, out_scores=None, out_rows=None, stream=None):
        """The exchange kernel alone (vs_exchange_merge_dev): this rank's [B,k] candidates (global
        rows) are pushed into every peer's buffer, flags are exchanged, and the G lists are merged."""
        import torch
        B, k = cand_scores.shape
        if out_scores is None:
            out_scores = torch.empty((B, k), dtype=torch.float32, device=cand_scores.device)
        if out_rows is None:
            out_rows = torch.empty((B, k), dtype=torch.int64, device=cand_scores.device)
        st = stream if stream is not None else torch.cuda.current_stream(cand_scores.device)
        N.check(self._lib.vs_exchange_merge_dev(self._h, _ptr(cand_scores.contiguous()), _ptr(cand_rows.contiguous()),
                                                B, k, _ptr(out_scores), _ptr(out_rows), _stream_ptr(st)))
        return out_scores, out_rows

    def exchange_begin(self):
        """Open a new deferred exchange (see ``query_push_dev`` / ``exchange_collect_dev``)."""
        N.check(self._lib.vs_exchange_begin(self._h))

    def query_push_dev(self, q, k: int, slot0: int, require_bits: Optional[Sequence[int]] = None,
                       mode: str = "auto", stream=None, pipelined: bool = False):
        """Local query + push of its candidates into slots [slot0, slot0+B) of every peer; no waiting."""
        import torch
        if q.dim() == 1:
            q = q[None]
        if not q.is_cuda or q.dtype != torch.float32 or q.shape[1] != self.dim:
            raise ValueError("queries must be a CUDA float32 [B, dim] tensor")
        q = q.contiguous()
        st = stream if stream is not None else torch.cuda.current_stream(q.device)
        N.check(self._lib.vs_query_topk_push_dev(self._h, _ptr(q), q.shape[0], int(k), _bits_array(require_bits),
                                                 _mode(mode, pipelined), int(slot0), _stream_ptr(st)))

    def exchange_collect_dev(self, B: int, k: int, out_scores=None, out_rows=None, stream=None):
        """Wait for and merge slots [0, B) of the open exchange -> global (scores [B,k], rows [B,k])."""
        import torch
        dev = torch.device("cuda", self.device)
        if out_scores is None:
            out_scores = torch.empty((B, k), dtype=torch.float32, device=dev)
        if out_rows is None:
            out_rows = torch.empty((B, k), dtype=torch.int64, device=dev)
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        N.check(self._lib.vs_exchange_collect_dev(self._h, int(B), int(k), _ptr(out_scores), _ptr(out_rows),
                                                  _stream_ptr(st)))
        return out_scores, out_rows

    def exchange_error(self) -> int:
        """Non-zero after a peer failed to deliver within ~3 s (the affected results came back EMPTY).
        Reads a host-mapped word: no CUDA call, but only meaningful after the stream was synchronised."""
        return int(self._lib.vs_exchange_error(self._h))

    def exchange_clear_error(self):
        N.check(self._lib.vs_exchange_clear_error(self._h))

    # -- filter sweep / dedup -----------------------------------------------------------------
    def filter_words(self) -> int:
        return int(self._lib.vs_filter_words(self._h))

    def filter_sweep(self, prompts, tau: float) -> np.ndarray:
        """prompts [F, dim] float32 numpy -> uint32 bit mask [F, filter_words()]."""
        a = np.ascontiguousarray(prompts, dtype=np.float32)
        if a.ndim == 1:
            a = a[None]
        out = np.zeros((a.shape[0], self.filter_words()), dtype=np.uint32)
        N.check(self._lib.vs_filter_sweep_host(self._h, a.ctypes.data, a.shape[0], float(tau), out.ctypes.data))
        return out

    def apply_filter_sweep(self, prompt, tau: float, bit: int) -> int:
        """One prompt: sweep all rows (K3) and store the outcome as filter bit ``bit`` of every row, all on
        the GPU (the answers the reference writes one Moondream call at a time, backend/app/main.py:1010-1033).
        Returns the number of rows that passed."""
        import torch
        if len(self) == 0:
            return 0
        dev = torch.device("cuda", self.device)
        p = torch.from_numpy(np.ascontiguousarray(prompt, dtype=np.float32).reshape(1, self.dim)).to(dev)
        with torch.cuda.device(dev):
            words = self.filter_sweep_dev(p, tau)
            self.apply_sweep_bits_dev(words, bit)
            w = words.cpu().numpy()
        return int(np.unpackbits(w.view(np.uint8)).sum())

    def filter_sweep_dev(self, prompts, tau: float, out_bits=None, stream=None):
        import torch
        F = prompts.shape[0]
        if out_bits is None:
            out_bits = torch.zeros((F, self.filter_words()), dtype=torch.int32, device=prompts.device)
        st = stream if stream is not None else torch.cuda.current_stream(prompts.device)
        N.check(self._lib.vs_filter_sweep_dev(self._h, _ptr(prompts.contiguous()), F, float(tau), _ptr(out_bits),
                                              _stream_ptr(st)))
        return out_bits

    def dedup(self, tau: float, row_lo: int = 0, row_hi: Optional[int] = None, capacity: int = 1 << 20):
        """All pairs (i<j) with cos >= tau for i in [row_lo,row_hi): (i, j, score) sorted by (i,j)."""
        if row_hi is None:
            row_hi = len(self)
        while True:
            oi = np.empty(capacity, dtype=np.int64)
            oj = np.empty(capacity, dtype=np.int64)
            os_ = np.empty(capacity, dtype=np.float32)
            cnt = C.c_int64(0)
            rc = self._lib.vs_dedup_host(self._h, int(row_lo), int(row_hi), float(tau), capacity,
                                         oi.ctypes.data, oj.ctypes.data, os_.ctypes.data, C.byref(cnt))
            if rc == N.VS_ERR_OVERFLOW:
                capacity = int(cnt.value) + 1024
                continue
            N.check(rc)
            m = int(cnt.value)
            order = np.lexsort((oj[:m], oi[:m]))
            return oi[:m][order], oj[:m][order], os_[:m][order]

    def dedup_dev(self, tau: float, row_lo: int, row_hi: int, out_i, out_j, out_score, out_count, stream=None):
        import torch
        st = stream if stream is not None else torch.cuda.current_stream(out_i.device)
        N.check(self._lib.vs_dedup_dev(self._h, int(row_lo), int(row_hi), float(tau), out_i.numel(), _ptr(out_i),
                                       _ptr(out_j), _ptr(out_score), _ptr(out_count), _stream_ptr(st)))
