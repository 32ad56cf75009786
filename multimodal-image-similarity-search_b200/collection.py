"""Drop-in replacement for the chromadb objects the reference uses.

The reference builds its vector store in ``init_chromadb()``
(``/root/reference/backend/app/utils.py:104-138``)::

    client = chromadb.PersistentClient(path=CHROMA_PERSIST_DIR)
    if COLLECTION_NAME in client.list_collections(): collection = client.get_collection(name=...)
    else: collection = client.create_collection(name=..., metadata={"hnsw:space": "cosine"})

and then only ever calls ``collection.add / get / query / update / delete / count``
(call sites listed per method below).  ``PersistentClient`` and ``Collection`` here honour
exactly that surface -- same keyword names, same return shapes (lists-per-query for ``query``,
flat lists for ``get``), cosine *distance* (``1 - cos``) ascending -- with the arithmetic done
by the sm_100a kernels behind ``libvecsearch_b200.so``.  Exact search: there is no HNSW graph,
so recall is 1 by construction.

Host-side state (ids, metadata dicts, documents) lives in Python, as it lives in SQLite inside
chromadb; vectors, inverse norms and filter bits live in HBM.  The backing index is a
``DeviceIndex`` (one GPU), a ``GroupIndex`` (one process, all GPUs of the box) or a ``ShardedIndex``
(one process per GPU); the collection only uses their common surface.

Persistence (``path=``; the directory the reference reopens at ``utils.py:109-123`` and walks at
``main.py:522-579``) is a SNAPSHOT + OPERATION LOG:

* snapshot of generation g -- ``rows.<g>.bin`` (raw row slab in the storage dtype, bf16 or f32),
  ``ids.<g>.txt``, ``meta.<g>.jsonl``, ``docs.<g>.jsonl``, ``masks.<g>.u64`` (filter bits) -- is loaded by
  chunked pinned host->device copies (``vs_add_raw_host``; a GroupIndex stripes it over its GPUs);
  metadata lines stay unparsed until somebody reads them;
* ``oplog.<g>.jsonl`` + ``oplog.<g>.vec`` hold what happened since: every record carries the byte offset of
  its vectors, a torn tail is ignored and truncated, the files are flushed as one unit per call;
* a checkpoint (``persist()``, ``close()``, or automatically when the log outgrows the snapshot) writes
  generation g+1, switches ``collection.json`` atomically, and drops the old files -- deleted rows vanish
  from disk at that point (compaction).
"""
from __future__ import annotations

import base64
import json
import os
import threading
from typing import Any, Dict, Iterable, List, Optional, Sequence

import numpy as np

from .index import DeviceIndex

_SCALAR = (str, int, float, bool)
FILTER_JSON_KEY = "filter_results_json"   # reference: backend/app/main.py:731, 1024
MAX_FILTERS = 256
FORMAT_VERSION = 2


def _is_cuda_rows(x) -> bool:
    """A float32 CUDA torch tensor: taken as is (encoder output never leaves the GPU)."""
    return hasattr(x, "is_cuda") and bool(x.is_cuda)


def _as_cuda_rows(x, dim: Optional[int]):
    import torch
    t = x.detach()
    if t.dim() == 1:
        t = t[None]
    if t.dim() != 2:
        raise ValueError(f"embeddings must be [n, dim], got shape {tuple(t.shape)}")
    if dim is not None and t.shape[1] != dim:
        raise ValueError(f"Embedding dimension {t.shape[1]} does not match collection dimensionality {dim}")
    return t.to(torch.float32).contiguous()


def _as_rows(embeddings, dim: Optional[int]) -> np.ndarray:
    if hasattr(embeddings, "detach"):          # torch tensor
        embeddings = embeddings.detach().cpu().numpy()
    a = np.asarray(embeddings, dtype=np.float32)
    if a.ndim == 1:
        a = a[None]
    if a.ndim != 2:
        raise ValueError(f"embeddings must be [n, dim], got shape {a.shape}")
    if dim is not None and a.shape[1] != dim:
        raise ValueError(f"Embedding dimension {a.shape[1]} does not match collection dimensionality {dim}")
    return np.ascontiguousarray(a)


def _yes(answer: Any) -> bool:
    """The predicate of backend/app/main.py:215."""
    return isinstance(answer, str) and answer.lower().strip() == "yes"


def _answers_of(meta: Optional[Dict[str, Any]]) -> Dict[str, Any]:
    """``json.loads(meta["filter_results_json"])`` with the reference's failure mode (main.py:207-213): {}."""
    if not meta or FILTER_JSON_KEY not in meta:
        return {}
    try:
        answers = json.loads(meta[FILTER_JSON_KEY])
    except (json.JSONDecodeError, TypeError):
        return {}
    return answers if isinstance(answers, dict) else {}


def _fsync_dir(path: str):
    try:
        fd = os.open(path, os.O_RDONLY)
        try:
            os.fsync(fd)
        finally:
            os.close(fd)
    except OSError:
        pass


class Collection:
    """Duck-type of ``chromadb.api.models.Collection`` restricted to what the reference calls."""

    def __init__(self, name: str, metadata: Optional[Dict[str, Any]] = None, *, device: int = 0,
                 dtype: str = "f32", path: Optional[str] = None, row_base: int = 0, index=None,
                 index_factory=None, durable: bool = False):
        metadata = dict(metadata or {})
        space = metadata.get("hnsw:space", "cosine")
        if space != "cosine":
            raise ValueError(f"only the cosine space is implemented (got hnsw:space={space!r})")
        self.name = name
        self.metadata = metadata
        self._device, self._dtype, self._row_base = device, dtype, row_base
        # created on first add (dimension fixed then) by index_factory(dim) / DeviceIndex, or injected
        self._index = index
        self._index_factory = index_factory
        self._ids: List[str] = []
        self._row_of: Dict[str, int] = {}
        # metadata / document of row r; a ``bytes`` entry is a raw JSON line of the snapshot, parsed on first use
        self._metas: List[Any] = []
        self._docs: List[Any] = []
        self._filters: List[str] = []                 # filter name -> bit index
        self._swept: set = set()                      # filters whose answers live in the device bits (apply_filter_sweep)
        self.filter_progress: Dict[str, Dict[str, Any]] = {}   # shape of main.py:964-986, served by /api/filter-progress
        self._lock = threading.RLock()                # update() runs on a worker thread (main.py:410)
        self._path = path
        self._durable = durable
        self._gen = 0
        self._log = self._vec = None
        self._log_ops = 0
        self._log_rows = 0
        self._snap_rows = 0
        if path is not None:
            self._open_store()

    # ------------------------------------------------------------------ internals
    @property
    def dim(self) -> Optional[int]:
        return None if self._index is None else self._index.dim

    @property
    def index(self):
        return self._index

    def _ensure_index(self, dim: int):
        if self._index is None:
            if self._index_factory is not None:
                self._index = self._index_factory(dim)
            else:
                self._index = DeviceIndex(dim, self._dtype, self._device, row_base=self._row_base)
        return self._index

    def _filter_bit(self, name: str, create: bool) -> Optional[int]:
        try:
            return self._filters.index(name)
        except ValueError:
            if not create:
                return None
            if len(self._filters) >= MAX_FILTERS:
                raise ValueError(f"more than {MAX_FILTERS} distinct filters")
            self._filters.append(name)
            return len(self._filters) - 1

    def _meta(self, row: int) -> Optional[Dict[str, Any]]:
        m = self._metas[row]
        if isinstance(m, bytes):
            m = json.loads(m) if m and m != b"null" else None
            self._metas[row] = m
        return m

    def _doc(self, row: int) -> Optional[str]:
        d = self._docs[row]
        if isinstance(d, bytes):
            d = json.loads(d) if d and d != b"null" else None
            self._docs[row] = d
        return d

    def _swept_answers(self, row: int) -> Dict[str, str]:
        """Answers of swept filters for one row, read from its device bits."""
        if not self._swept:
            return {}
        have = set(self._index.get_filter_bits(row))
        return {f: ("yes" if self._filters.index(f) in have else "no") for f in self._filters if f in self._swept}

    def _meta_out(self, row: int) -> Optional[Dict[str, Any]]:
        """The metadata a caller sees: stored dict + (lazily) the answers of swept filters merged into
        ``filter_results_json``, so the reference's post-filter and UI keep working unchanged."""
        m = self._meta(row)
        if not self._swept:
            return dict(m) if m is not None else None
        out = dict(m or {})
        answers = _answers_of(m)
        answers.update(self._swept_answers(row))
        out[FILTER_JSON_KEY] = json.dumps(answers)
        return out

    def _json_bits_of(self, meta: Optional[Dict[str, Any]]):
        """(bits set by "yes" answers, bits mentioned at all) of a row's stored filter_results_json."""
        yes, named = [], []
        for fname, ans in _answers_of(meta).items():
            b = self._filter_bit(fname, create=True)
            named.append(b)
            if _yes(ans):
                yes.append(b)
        return yes, named

    def _sync_filter_bits(self, row: int):
        """Mirror the row's ``filter_results_json`` "yes" answers into its device filter bits; bits of swept
        filters the JSON does not mention are kept."""
        yes, named = self._json_bits_of(self._meta(row))
        if self._swept:
            keep = {self._filters.index(f) for f in self._swept} - set(named)
            yes = sorted(set(yes) | (set(self._index.get_filter_bits(row)) & keep))
        self._index.set_filter_bits(row, yes)

    @staticmethod
    def _check_meta(m: Optional[Dict[str, Any]]):
        if m is None:
            return
        if not isinstance(m, dict):
            raise ValueError("metadata must be a dict")
        for k, v in m.items():
            if not isinstance(k, str) or not isinstance(v, _SCALAR):
                raise ValueError(f"metadata values must be str, int, float or bool (key {k!r})")

    # ------------------------------------------------------------------ persistence: snapshot + operation log
    def _f(self, stem: str, gen: Optional[int] = None) -> str:
        g = self._gen if gen is None else gen
        base, ext = stem.split(".")
        return os.path.join(self._path, f"{base}.{g}.{ext}")

    def _manifest_path(self) -> str:
        return os.path.join(self._path, "collection.json")

    def _read_manifest(self) -> Dict[str, Any]:
        try:
            with open(self._manifest_path(), "r", encoding="utf-8") as f:
                return json.load(f)
        except FileNotFoundError:
            return {}

    def _write_manifest(self, **extra):
        info = self._read_manifest()
        info.update({"name": self.name, "metadata": self.metadata, "dtype": self._dtype, "format": FORMAT_VERSION})
        info.update(extra)
        tmp = self._manifest_path() + ".tmp"
        with open(tmp, "w", encoding="utf-8") as f:
            json.dump(info, f)
            f.flush()
            os.fsync(f.fileno())
        os.replace(tmp, self._manifest_path())
        _fsync_dir(self._path)

    def _open_store(self):
        os.makedirs(self._path, exist_ok=True)
        info = self._read_manifest()
        if info.get("format", FORMAT_VERSION) != FORMAT_VERSION and "gen" in info:
            raise ValueError(f"{self._path}: unknown collection format {info.get('format')}")
        self._gen = int(info.get("gen", 0))
        if info.get("count", 0) > 0:
            self._load_snapshot(info)
        self._replay_log()
        self._log = open(self._f("oplog.jsonl"), "ab")
        self._vec = open(self._f("oplog.vec"), "ab")
        # stale generations (a crash between the manifest switch and the clean-up)
        for fn in os.listdir(self._path):
            parts = fn.split(".")
            if len(parts) == 3 and parts[1].isdigit() and int(parts[1]) != self._gen and \
                    parts[0] in ("rows", "ids", "meta", "docs", "masks", "oplog"):
                try:
                    os.remove(os.path.join(self._path, fn))
                except OSError:
                    pass

    def _load_snapshot(self, info: Dict[str, Any]):
        n, dim = int(info["count"]), int(info["dim"])
        self._filters = list(info.get("filters", []))
        self._swept = set(info.get("swept", []))
        ix = self._ensure_index(dim)
        if hasattr(ix, "reserve"):
            ix.reserve(n)
        if hasattr(ix, "add_raw"):
            slab = np.memmap(self._f("rows.bin"), dtype=ix.storage_dtype, mode="r", shape=(n, dim))
            ix.add_raw(slab)                            # chunked pinned host -> device, per GPU of a group
            del slab
        else:                                           # an index without the raw path (tests' fake): f32 rows
            ix.add(np.fromfile(self._f("rows.bin"), dtype=np.float32).reshape(n, dim))
        with open(self._f("ids.txt"), "rb") as f:
            raw = f.read()
        if info.get("ids_format") == "json":
            self._ids = [json.loads(line) for line in raw.split(b"\n")[:n]]
        else:
            self._ids = raw.decode("utf-8").split("\n")[:n]
        if len(self._ids) != n:
            raise ValueError(f"{self._path}: snapshot holds {len(self._ids)} ids for {n} rows")
        self._row_of = dict(zip(self._ids, range(n)))
        with open(self._f("meta.jsonl"), "rb") as f:
            self._metas = f.read().split(b"\n")[:n]     # parsed lazily (_meta)
        with open(self._f("docs.jsonl"), "rb") as f:
            self._docs = f.read().split(b"\n")[:n]
        if len(self._metas) != n or len(self._docs) != n:
            raise ValueError(f"{self._path}: snapshot metadata/documents do not match {n} rows")
        mp = self._f("masks.u64")
        if os.path.exists(mp):
            words = np.fromfile(mp, dtype=np.uint64).reshape(n, -1)
            if hasattr(ix, "set_filter_words_range"):
                ix.set_filter_words_range(0, words)
            else:
                ix.set_filter_bits_range(0, [[b for b in range(64 * words.shape[1]) if (int(w[b // 64]) >> (b % 64)) & 1]
                                             for w in words])
        self._snap_rows = n

    def _replay_log(self):
        log, vec = self._f("oplog.jsonl"), self._f("oplog.vec")
        if not os.path.exists(log):
            return
        good_log = good_vec = 0
        with open(log, "rb") as f, open(vec, "rb") if os.path.exists(vec) else open(os.devnull, "rb") as v:
            vec_size = os.path.getsize(vec) if os.path.exists(vec) else 0
            for line in f:
                if not line.endswith(b"\n"):
                    break                               # torn tail: the record was never completed
                try:
                    op = json.loads(line)
                except json.JSONDecodeError:
                    break
                kind = op.get("op")
                if kind == "add":
                    nbytes = int(op["nbytes"])
                    if op["off"] + nbytes > vec_size:
                        break                           # vectors of this record never reached the disk
                    v.seek(op["off"])
                    dt = np.dtype(op["vdtype"])
                    rows = np.frombuffer(v.read(nbytes), dtype=dt).reshape(len(op["ids"]), op["dim"])
                    self._add_rows(op["ids"], rows, op.get("metadatas") or [None] * len(op["ids"]),
                                   op.get("documents") or [None] * len(op["ids"]), log=False, raw=(op["vdtype"] != "float32"))
                    good_vec = max(good_vec, op["off"] + nbytes)
                elif kind == "update":
                    self._update_one(op["id"], op.get("metadata"), op.get("document"), log=False)
                elif kind == "delete":
                    self._delete_ids(op["ids"], log=False)
                elif kind == "clear":
                    self._delete_ids(list(self._ids), log=False)
                elif kind == "sweep":
                    self._replay_sweep(op)
                good_log = f.tell()
                self._log_ops += 1
        # drop a torn tail / orphan vectors so the next append starts at a clean boundary
        if good_log != os.path.getsize(log):
            with open(log, "r+b") as f:
                f.truncate(good_log)
        if os.path.exists(vec) and good_vec != os.path.getsize(vec):
            with open(vec, "r+b") as f:
                f.truncate(good_vec)

    def _log_op(self, op: Dict[str, Any], rows: Optional[np.ndarray] = None):
        if self._log is None:
            return
        if rows is not None:
            a = np.ascontiguousarray(rows)
            op = dict(op, off=self._vec.tell(), nbytes=int(a.nbytes), vdtype=str(a.dtype), dim=int(a.shape[1]))
            self._vec.write(a.tobytes())
            self._log_rows += int(a.shape[0])
        self._log.write(json.dumps(op, separators=(",", ":")).encode("utf-8") + b"\n")
        self._log_ops += 1

    def _log_flush(self):
        """Make the call durable as ONE unit: vectors first, then the records that point at them."""
        if self._log is None:
            return
        self._vec.flush()
        if self._durable:
            os.fsync(self._vec.fileno())
        self._log.flush()
        if self._durable:
            os.fsync(self._log.fileno())
        # the log is replayed op by op on reopen: fold it into a snapshot once it outgrows the snapshot
        if self._log_ops > 1000 and (self._log_ops + self._log_rows) > max(self._snap_rows, 1) // 2 + 1000:
            self.persist()

    def persist(self):
        """Checkpoint: write generation g+1 (compacted -- deleted rows are gone), switch the manifest
        atomically, drop generation g and its log."""
        with self._lock:
            if self._path is None:
                return
            n = len(self._ids)
            old, new = self._gen, self._gen + 1
            extra: Dict[str, Any] = {"gen": new, "count": n, "filters": list(self._filters), "swept": sorted(self._swept)}
            if n > 0:
                ix = self._index
                extra["dim"] = int(ix.dim)
                with open(self._f("rows.bin", new), "wb") as f:
                    step = max(1, (256 << 20) // (ix.dim * 4))
                    for r0 in range(0, n, step):
                        m = min(step, n - r0)
                        rows = ix.get_raw(r0, m) if hasattr(ix, "get_raw") else ix.get_rows(r0, m)
                        f.write(np.ascontiguousarray(rows).tobytes())
                    f.flush()
                    os.fsync(f.fileno())
                plain = not any("\n" in i for i in self._ids)
                extra["ids_format"] = "plain" if plain else "json"
                with open(self._f("ids.txt", new), "wb") as f:
                    f.write(("\n".join(self._ids) if plain else "\n".join(json.dumps(i) for i in self._ids)).encode("utf-8"))
                    f.write(b"\n")
                for stem, items in (("meta.jsonl", self._metas), ("docs.jsonl", self._docs)):
                    with open(self._f(stem, new), "wb") as f:
                        f.write(b"\n".join(x if isinstance(x, bytes) else
                                           (b"null" if x is None else json.dumps(x, separators=(",", ":")).encode("utf-8"))
                                           for x in items))
                        f.write(b"\n")
                if self._filters:
                    if hasattr(ix, "get_filter_words_range"):
                        words = ix.get_filter_words_range(0, n)
                    else:
                        from .index import bits_to_words
                        words = bits_to_words([ix.get_filter_bits(r) for r in range(n)])
                    if words.any():
                        words.tofile(self._f("masks.u64", new))
            if self._log is not None:
                self._log.close()
                self._vec.close()
            self._write_manifest(**extra)               # the switch
            self._gen = new
            for stem in ("rows.bin", "ids.txt", "meta.jsonl", "docs.jsonl", "masks.u64", "oplog.jsonl", "oplog.vec"):
                try:
                    os.remove(self._f(stem, old))
                except OSError:
                    pass
            self._log = open(self._f("oplog.jsonl"), "ab")
            self._vec = open(self._f("oplog.vec"), "ab")
            self._log_ops = self._log_rows = 0
            self._snap_rows = n

    # ------------------------------------------------------------------ add
    def add(self, ids, embeddings, metadatas=None, documents=None):
        """``collection.add(ids=[id], embeddings=[vec], metadatas=[meta], documents=[desc])``
        (backend/app/main.py:735-740).  Ids that already exist are skipped, as in chromadb."""
        if isinstance(ids, str):
            ids = [ids]
        ids = list(ids)
        # a CUDA tensor (e.g. a batch of CLIP image embeddings) is ingested device-to-device by K6
        rows = _as_cuda_rows(embeddings, self.dim) if _is_cuda_rows(embeddings) else _as_rows(embeddings, self.dim)
        n = len(ids)
        if rows.shape[0] != n:
            raise ValueError(f"{n} ids but {rows.shape[0]} embeddings")
        metadatas = [None] * n if metadatas is None else ([metadatas] if isinstance(metadatas, dict) else list(metadatas))
        documents = [None] * n if documents is None else ([documents] if isinstance(documents, str) else list(documents))
        if len(metadatas) != n or len(documents) != n:
            raise ValueError("ids, embeddings, metadatas and documents must have the same length")
        for m in metadatas:
            self._check_meta(m)
        with self._lock:
            keep, seen = [], set()
            for i, id_ in enumerate(ids):
                if not isinstance(id_, str):
                    raise ValueError("ids must be strings")
                if id_ in self._row_of or id_ in seen:
                    continue
                seen.add(id_)
                keep.append(i)
            if not keep:
                return
            if len(keep) != n:
                rows = rows[keep]
                ids, metadatas, documents = [ids[i] for i in keep], [metadatas[i] for i in keep], [documents[i] for i in keep]
            self._add_rows(ids, rows, metadatas, documents, log=True)
            self._log_flush()

    def _add_rows(self, ids, rows, metadatas, documents, log: bool, raw: bool = False):
        ix = self._ensure_index(int(rows.shape[1]))
        first = ix.add_raw(rows) if raw else ix.add(rows)
        assert first == len(self._ids), "host/device row bookkeeping diverged"
        n = len(ids)
        self._ids.extend(ids)
        self._row_of.update(zip(ids, range(first, first + n)))
        self._metas.extend(dict(m) if m is not None else None for m in metadatas)
        self._docs.extend(documents)
        if any(m and FILTER_JSON_KEY in m for m in metadatas):
            # one bulk copy (one call per shard) instead of one call per row
            self._index.set_filter_bits_range(first, [self._json_bits_of(m)[0] for m in metadatas])
        if log and self._log is not None:
            # the log keeps the rows exactly as stored (bf16 stays bf16): a replay re-creates the same bits
            stored = ix.get_raw(first, n) if hasattr(ix, "get_raw") else \
                (rows if isinstance(rows, np.ndarray) else rows.cpu().numpy())
            self._log_op({"op": "add", "ids": list(ids),
                          "metadatas": None if all(m is None for m in metadatas) else list(metadatas),
                          "documents": None if all(d is None for d in documents) else list(documents)}, stored)

    # ------------------------------------------------------------------ query
    def query(self, query_embeddings=None, query_texts=None, n_results: int = 10, where=None,
              where_document=None, include: Sequence[str] = ("metadatas", "documents", "distances"),
              where_filters: Optional[Sequence[str]] = None, filter_mode: str = "post", mode: str = "auto"):
        """``collection.query(query_embeddings=[emb.tolist()], n_results=k,
        include=["metadatas","distances"])`` (backend/app/main.py:761-765; legacy app.py:310-314).

        Returns chromadb's dict-of-lists-per-query; ``distances`` are cosine distances
        ``1 - cos`` in ascending order.  ``n_results`` is clamped to ``count()``.
        ``query_texts`` raises: the collection has no embedding function, and the legacy
        caller relies on that to fall back to CLIP (app.py:343-372).

        Extensions (north_star): ``where_filters=[filter names]`` with ``filter_mode="pre"``
        restricts the search to rows whose stored answers are all "yes" *inside the kernel*
        (bit test fused with the top-k insert); ``"post"`` reproduces the reference's order of
        operations (top-k first, then the predicate of main.py:215), so fewer than n_results
        rows may come back."""
        if query_texts is not None:
            raise ValueError("query_texts is not supported: this collection has no embedding function; "
                             "pass query_embeddings")
        if query_embeddings is None:
            raise ValueError("query_embeddings is required")
        if where is not None or where_document is not None:
            raise NotImplementedError("where / where_document are not used by the reference and not implemented")
        if filter_mode not in ("post", "pre"):
            raise ValueError("filter_mode must be 'post' or 'pre'")
        if n_results <= 0:
            raise ValueError("n_results must be positive")
        with self._lock:
            if _is_cuda_rows(query_embeddings):       # device-resident queries: no H2D, only the [B,k] result comes back
                qd = _as_cuda_rows(query_embeddings, self.dim)

                def run_dev(k, require):
                    s_, r_ = self._index.query_dev(qd, k, require_bits=require, mode=mode)
                    return s_.cpu().numpy(), r_.cpu().numpy()
                return self._run_query(int(qd.shape[0]), n_results, list(include), where_filters, filter_mode, run_dev)
            q = _as_rows(query_embeddings, self.dim)
            return self._run_query(q.shape[0], n_results, list(include), where_filters, filter_mode,
                                   lambda k, require: self._index.query(q, k, require_bits=require, mode=mode))

    def query_multimodal(self, image_embeddings, text_embeddings, weight_image=0.5, n_results: int = 10,
                         include: Sequence[str] = ("metadatas", "documents", "distances"),
                         where_filters: Optional[Sequence[str]] = None, filter_mode: str = "post", mode: str = "auto"):
        """``search_multimodal`` (backend/app/main.py:829-867) as one call: the blend
        ``c = w*i^ + (1-w)*t^; c /= ||c||`` (main.py:850-860) runs on the device and feeds the same
        top-k kernels as :meth:`query`.  ``weight_image`` is a float or one float per query."""
        if filter_mode not in ("post", "pre"):
            raise ValueError("filter_mode must be 'post' or 'pre'")
        if n_results <= 0:
            raise ValueError("n_results must be positive")
        with self._lock:
            qi = _as_rows(image_embeddings, self.dim)
            qt = _as_rows(text_embeddings, self.dim)
            if qi.shape != qt.shape:
                raise ValueError("image and text embeddings must have the same shape")
            return self._run_query(qi.shape[0], n_results, list(include), where_filters, filter_mode,
                                   lambda k, require: self._index.query_multimodal(qi, qt, weight_image, k,
                                                                                   require_bits=require, mode=mode))

    def _run_query(self, B, n_results, include, where_filters, filter_mode, run):
        count = len(self._ids)
        out: Dict[str, Any] = {"ids": [[] for _ in range(B)], "embeddings": None, "documents": None,
                               "metadatas": None, "distances": None, "uris": None, "data": None,
                               "included": include}
        for key in ("metadatas", "documents", "distances", "embeddings"):
            if key in include:
                out[key] = [[] for _ in range(B)]
        if count == 0:
            return out
        k = min(int(n_results), count)
        require = None
        if where_filters and filter_mode == "pre":
            require = []
            for f in where_filters:
                b = self._filter_bit(f, create=False)
                if b is None:
                    return out          # nobody answered this filter: nothing can match
                require.append(b)
        scores, rows = run(k, require)
        for b in range(B):
            for s, r in zip(scores[b].tolist(), rows[b].tolist()):
                if r < 0:
                    continue
                r -= self._row_base
                if where_filters and filter_mode == "post" and not self._passes(r, where_filters):
                    continue
                out["ids"][b].append(self._ids[r])
                if out["distances"] is not None:
                    out["distances"][b].append(1.0 - s)
                if out["metadatas"] is not None:
                    out["metadatas"][b].append(self._meta_out(r))
                if out["documents"] is not None:
                    out["documents"][b].append(self._doc(r))
                if out["embeddings"] is not None:
                    out["embeddings"][b].append(self._index.get_rows(r, 1)[0].tolist())
        return out

    def _passes(self, row: int, filters: Iterable[str]) -> bool:
        answers = _answers_of(self._meta(row))
        if self._swept:
            answers.update(self._swept_answers(row))
        return all(_yes(answers.get(f, "")) for f in filters)

    # ------------------------------------------------------------------ get / count
    def get(self, ids=None, where=None, limit: Optional[int] = None, offset: Optional[int] = None,
            where_document=None, include: Sequence[str] = ("metadatas", "documents")):
        """``collection.get(include=[])`` -> ``{"ids": [...]}`` (backend/app/main.py:533,556,1065);
        ``collection.get(ids=[id], include=["metadatas"])`` (main.py:563-566, 631-634): flat
        lists, unknown ids are simply absent (the duplicate check at main.py:636-640 relies on
        ``existing_check["ids"]`` being empty for a new image)."""
        if where is not None or where_document is not None:
            raise NotImplementedError("where / where_document are not used by the reference and not implemented")
        include = list(include)
        with self._lock:
            if ids is None:
                rows = range(len(self._ids))
            else:
                if isinstance(ids, str):
                    ids = [ids]
                rows = [self._row_of[i] for i in ids if i in self._row_of]
            if offset:
                rows = rows[offset:]
            if limit is not None:
                rows = rows[:limit]
            all_rows = ids is None and not offset and limit is None
            out: Dict[str, Any] = {"ids": list(self._ids) if all_rows else [self._ids[r] for r in rows],
                                   "embeddings": None, "documents": None,
                                   "metadatas": None, "uris": None, "data": None, "included": include}
            if "metadatas" in include:
                out["metadatas"] = [self._meta_out(r) for r in rows]
            if "documents" in include:
                out["documents"] = [self._doc(r) for r in rows]
            if "embeddings" in include:
                out["embeddings"] = [self._index.get_rows(r, 1)[0] for r in rows]
            return out

    def count(self) -> int:
        """``collection.count()`` (init_db.py:58)."""
        with self._lock:
            return len(self._ids)

    def peek(self, limit: int = 10):
        return self.get(limit=limit)

    # ------------------------------------------------------------------ update / delete
    def update(self, ids, embeddings=None, metadatas=None, documents=None):
        """``collection.update(ids=[id], metadatas=[partial], documents=[...])``
        (backend/app/main.py:503-510, 1030-1033; app.py:2404-2408).  Metadata keys are MERGED into
        the stored dict (chromadb semantics -- main.py:503 passes a partial dict).  Unknown ids
        are ignored.  Thread-safe against concurrent ``query`` (the filter worker thread)."""
        if isinstance(ids, str):
            ids = [ids]
        ids = list(ids)
        n = len(ids)
        metadatas = [None] * n if metadatas is None else ([metadatas] if isinstance(metadatas, dict) else list(metadatas))
        documents = [None] * n if documents is None else ([documents] if isinstance(documents, str) else list(documents))
        if len(metadatas) != n or len(documents) != n:
            raise ValueError("ids, metadatas and documents must have the same length")
        for m in metadatas:
            self._check_meta(m)
        rows = None if embeddings is None else _as_rows(embeddings, self.dim)
        with self._lock:
            for j, id_ in enumerate(ids):
                if id_ not in self._row_of:
                    continue
                if rows is not None:
                    # re-embed: delete + add keeps the slab dense (row number changes, id does not)
                    r = self._row_of[id_]
                    meta, doc = self._meta(r), self._doc(r)
                    self._delete_ids([id_], log=True)
                    self._add_rows([id_], rows[j:j + 1], [meta], [doc], log=True)
                self._update_one(id_, metadatas[j], documents[j], log=True)
            self._log_flush()

    def _update_one(self, id_: str, metadata, document, log: bool):
        r = self._row_of.get(id_)
        if r is None:
            return
        if metadata is not None:
            merged = dict(self._meta(r) or {})
            merged.update(metadata)
            self._metas[r] = merged
            if FILTER_JSON_KEY in metadata:
                self._sync_filter_bits(r)
        if document is not None:
            self._docs[r] = document
        if log and (metadata is not None or document is not None):
            self._log_op({"op": "update", "id": id_, "metadata": metadata, "document": document})

    def delete(self, ids=None, where=None, where_document=None):
        """``collection.delete(ids=all_ids)`` (backend/app/main.py:1069): any number of ids costs ONE
        compaction kernel and one synchronise per shard; deleting everything just drops the row count."""
        if where is not None or where_document is not None:
            raise NotImplementedError("where / where_document are not used by the reference and not implemented")
        with self._lock:
            if ids is None:
                ids = list(self._ids)
            elif isinstance(ids, str):
                ids = [ids]
            self._delete_ids(list(ids), log=True)
            self._log_flush()

    def _delete_ids(self, ids: List[str], log: bool):
        n = len(self._ids)
        if n and len(ids) == n and ids == self._ids:     # reset_system (main.py:1065-1069): every id, in store order
            self._index.clear()
            self._ids, self._metas, self._docs, self._row_of = [], [], [], {}
            if log:
                self._log_op({"op": "clear"})
            return
        rows = sorted({self._row_of[i] for i in ids if i in self._row_of})
        if not rows:
            return
        gone = [self._ids[r] for r in rows]
        if len(rows) == n:
            self._index.clear()
            self._ids, self._metas, self._docs, self._row_of = [], [], [], {}
        elif len(rows) == 1 or not hasattr(self._index, "remove_rows"):
            for r in reversed(rows):                     # descending: earlier moves never touch later holes
                moved = self._index.remove(r)
                last = len(self._ids) - 1
                assert moved in (-1, last)
                del self._row_of[self._ids[r]]
                if r != last:
                    self._ids[r], self._metas[r], self._docs[r] = self._ids[last], self._metas[last], self._docs[last]
                    self._row_of[self._ids[r]] = r
                self._ids.pop(); self._metas.pop(); self._docs.pop()
        else:
            src, dst = self._index.remove_rows(np.asarray(rows, dtype=np.int64))
            for i in gone:
                del self._row_of[i]
            for s, d in zip(src.tolist(), dst.tolist()):
                self._ids[d], self._metas[d], self._docs[d] = self._ids[s], self._metas[s], self._docs[s]
                self._row_of[self._ids[d]] = d
            new_n = n - len(rows)
            del self._ids[new_n:], self._metas[new_n:], self._docs[new_n:]
        if log:
            self._log_op({"op": "delete", "ids": gone})

    # ------------------------------------------------------------------ north_star extensions
    def filter_names(self) -> List[str]:
        with self._lock:
            return list(self._filters)

    def filter_sweep(self, prompt_embeddings, tau: float) -> np.ndarray:
        """[F, dim] prompt embeddings -> bool [F, count]: cos(prompt, row) >= tau, on tcgen05
        (BASELINE config 4).  bf16 collections only."""
        with self._lock:
            n = len(self._ids)
            p = _as_rows(prompt_embeddings, self.dim)
            if n == 0:
                return np.zeros((p.shape[0], 0), dtype=bool)
            bits = self._index.filter_sweep(p, tau)
            return np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")[:, :n].astype(bool)

    def apply_filter_sweep(self, filter_name: str, prompt_embedding, tau: float) -> int:
        """The CLIP-side counterpart of ``process_filter_on_all_images`` (backend/app/main.py:939-1056):
        one prompt embedding is swept over ALL rows on tensor cores and the outcome becomes the rows'
        answers for ``filter_name``.  The answers are written where the query kernels test them -- filter
        bit of every row, set by one kernel per GPU, nothing per row on the host -- and are materialised
        as ``filter_results_json[filter_name] = "yes" | "no"`` whenever metadata is read (get / query /
        the post-filter), so the reference's filter pass and UI see what the Moondream loop would have
        stored.  ``filter_progress[filter_name]`` follows the reference's progress contract
        (main.py:964-986, 1040-1056).  Returns the number of "yes" rows."""
        with self._lock:
            total = len(self._ids)
            self.filter_progress[filter_name] = {"status": "processing", "progress": 0, "current_image": "",
                                                 "processed": 0, "total": total}
            try:
                bit = self._filter_bit(filter_name, create=True)
                n_yes = 0
                if total:
                    p = _as_rows(prompt_embedding, self.dim)[0]
                    if hasattr(self._index, "apply_filter_sweep"):
                        n_yes = int(self._index.apply_filter_sweep(p, float(tau), bit))
                    else:                                 # an index without the device-side write: bits via the host
                        mask = self.filter_sweep(p[None], tau)[0]
                        from .index import bits_to_words
                        cur = [set(self._index.get_filter_bits(r)) for r in range(total)]
                        self._index.set_filter_bits_range(0, [sorted((c - {bit}) | ({bit} if h else set()))
                                                              for c, h in zip(cur, mask.tolist())])
                        n_yes = int(mask.sum())
                self._swept.add(filter_name)
                if self._log is not None:
                    words = self._column_bits(bit, total)
                    self._log_op({"op": "sweep", "name": filter_name, "bit": bit, "n": total,
                                  "bits": base64.b64encode(words.tobytes()).decode("ascii")})
                    self._log_flush()
                self.filter_progress[filter_name] = {"status": "completed", "progress": 100, "processed": total,
                                                     "total": total, "matched": n_yes}
                return n_yes
            except Exception as e:
                self.filter_progress[filter_name] = {"status": "error", "message": str(e), "progress": 0}
                raise

    def _column_bits(self, bit: int, n: int) -> np.ndarray:
        """Packed (uint8, little bit order) column ``bit`` of the rows' filter words."""
        ix = self._index
        if hasattr(ix, "get_filter_words_range"):
            col = (ix.get_filter_words_range(0, n)[:, bit // 64] >> np.uint64(bit % 64)) & np.uint64(1)
            return np.packbits(col.astype(np.uint8), bitorder="little")
        return np.packbits(np.array([bit in ix.get_filter_bits(r) for r in range(n)], dtype=np.uint8), bitorder="little")

    def _replay_sweep(self, op: Dict[str, Any]):
        n, bit = int(op["n"]), int(op["bit"])
        while len(self._filters) <= bit:
            self._filters.append(op["name"] if len(self._filters) == bit else f"__unused_{len(self._filters)}")
        self._swept.add(op["name"])
        if n == 0 or n != len(self._ids):
            return
        yes = np.unpackbits(np.frombuffer(base64.b64decode(op["bits"]), dtype=np.uint8), bitorder="little")[:n].astype(bool)
        ix = self._index
        if hasattr(ix, "get_filter_words_range"):
            words = ix.get_filter_words_range(0, n)
            m = np.uint64(1 << (bit % 64))
            words[:, bit // 64] = np.where(yes, words[:, bit // 64] | m, words[:, bit // 64] & ~m)
            ix.set_filter_words_range(0, words)
        else:
            ix.set_filter_bits_range(0, [sorted((set(ix.get_filter_bits(r)) - {bit}) | ({bit} if y else set()))
                                         for r, y in enumerate(yes.tolist())])

    def find_duplicates(self, threshold: float = 0.95):
        """All pairs of stored embeddings with cosine >= threshold (BASELINE config 5):
        list of (id_i, id_j, score) with row(i) < row(j).  bf16 collections only."""
        with self._lock:
            if len(self._ids) < 2:
                return []
            i, j, s = self._index.dedup(threshold)
            return [(self._ids[a], self._ids[b], float(c)) for a, b, c in zip(i.tolist(), j.tolist(), s.tolist())]

    def close(self):
        with self._lock:
            if self._log is not None:
                if self._log_ops > 0:
                    self.persist()                      # fold the log: the next open is one slab upload
                self._log.close(); self._vec.close()
                self._log = None
            if self._index is not None:
                self._index.close()
                self._index = None


class PersistentClient:
    """``chromadb.PersistentClient(path=...)`` as used at backend/app/utils.py:113 and
    init_db.py:36: collections persist under ``path/<name>/`` (snapshot + operation log, see the module
    docstring) and are uploaded to the GPU(s) when reopened.  ``devices=[...]`` makes every collection a
    single-process multi-GPU one (``GroupIndex``)."""

    def __init__(self, path: str = "./chroma", *, device: int = 0, dtype: str = "f32",
                 devices: Optional[Sequence[int]] = None, durable: bool = False):
        self.path = path
        self._device, self._dtype, self._devices, self._durable = device, dtype, devices, durable
        os.makedirs(path, exist_ok=True)
        self._open: Dict[str, Collection] = {}

    def _dir(self, name: str) -> str:
        return os.path.join(self.path, name)

    def _factory(self, dtype: str):
        if self._devices is None:
            return None
        devices = list(self._devices)

        def make(dim: int):
            from .group_index import GroupIndex
            return GroupIndex(dim, dtype, devices=devices)
        return make

    def list_collections(self) -> List[str]:
        """Names only -- the reference does ``COLLECTION_NAME in client.list_collections()``
        (utils.py:119-121, the chromadb >= 0.6 behaviour)."""
        on_disk = {d for d in os.listdir(self.path) if os.path.exists(os.path.join(self.path, d, "collection.json"))}
        return sorted(on_disk | set(self._open))

    def create_collection(self, name: str, metadata: Optional[Dict[str, Any]] = None, get_or_create: bool = False,
                          **_unused) -> Collection:
        if name in self.list_collections():
            if get_or_create:
                return self.get_collection(name)
            raise ValueError(f"Collection {name} already exists")
        space = (metadata or {}).get("hnsw:space", "cosine")
        if space != "cosine":
            raise ValueError(f"only the cosine space is implemented (got hnsw:space={space!r})")
        d = self._dir(name)
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "collection.json"), "w", encoding="utf-8") as f:
            json.dump({"name": name, "metadata": metadata or {}, "dtype": self._dtype, "format": FORMAT_VERSION,
                       "gen": 0, "count": 0}, f)
        c = Collection(name, metadata, device=self._device, dtype=self._dtype, path=d,
                       index_factory=self._factory(self._dtype), durable=self._durable)
        self._open[name] = c
        return c

    def get_collection(self, name: str, **_unused) -> Collection:
        if name in self._open:
            return self._open[name]
        d = self._dir(name)
        cfg = os.path.join(d, "collection.json")
        if not os.path.exists(cfg):
            raise ValueError(f"Collection {name} does not exist.")
        with open(cfg, "r", encoding="utf-8") as f:
            info = json.load(f)
        dtype = info.get("dtype", self._dtype)
        c = Collection(name, info.get("metadata"), device=self._device, dtype=dtype, path=d,
                       index_factory=self._factory(dtype), durable=self._durable)
        self._open[name] = c
        return c

    def get_or_create_collection(self, name: str, metadata: Optional[Dict[str, Any]] = None, **kw) -> Collection:
        return self.create_collection(name, metadata, get_or_create=True, **kw)

    def delete_collection(self, name: str):
        import shutil
        c = self._open.pop(name, None)
        if c is not None:
            c.close()
        d = self._dir(name)
        if not os.path.exists(os.path.join(d, "collection.json")):
            raise ValueError(f"Collection {name} does not exist.")
        shutil.rmtree(d)

    def heartbeat(self) -> int:
        import time
        return time.time_ns()


def Client(*_a, **kw) -> PersistentClient:   # in-memory flavour used by tests
    import tempfile
    return PersistentClient(tempfile.mkdtemp(prefix="vecsearch_b200_"), **kw)
