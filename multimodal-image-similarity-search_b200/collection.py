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
chromadb; vectors, inverse norms and filter bits live in HBM.
"""
from __future__ import annotations

import json
import os
import threading
from typing import Any, Dict, Iterable, List, Optional, Sequence

import numpy as np

from .index import DeviceIndex

_SCALAR = (str, int, float, bool)
FILTER_JSON_KEY = "filter_results_json"   # reference: backend/app/main.py:731, 1024
MAX_FILTERS = 256


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


class _NoRows:
    """Placeholder for device rows that need no host copy (in-memory collection: nothing is logged)."""

    def __init__(self, dim: int):
        self.shape = (0, dim)

    def __getitem__(self, j):
        return None


def _yes(answer: Any) -> bool:
    """The predicate of backend/app/main.py:215."""
    return isinstance(answer, str) and answer.lower().strip() == "yes"


class Collection:
    """Duck-type of ``chromadb.api.models.Collection`` restricted to what the reference calls."""

    def __init__(self, name: str, metadata: Optional[Dict[str, Any]] = None, *, device: int = 0,
                 dtype: str = "f32", path: Optional[str] = None, row_base: int = 0, index=None):
        metadata = dict(metadata or {})
        space = metadata.get("hnsw:space", "cosine")
        if space != "cosine":
            raise ValueError(f"only the cosine space is implemented (got hnsw:space={space!r})")
        self.name = name
        self.metadata = metadata
        self._device, self._dtype, self._row_base = device, dtype, row_base
        # created on first add (dimension fixed then), or injected: a ShardedIndex spanning several GPUs
        self._index: Optional[DeviceIndex] = index
        self._ids: List[str] = []
        self._row_of: Dict[str, int] = {}
        self._metas: List[Optional[Dict[str, Any]]] = []
        self._docs: List[Optional[str]] = []
        self._filters: List[str] = []                 # filter name -> bit index
        self._lock = threading.RLock()                # update() runs on a worker thread (main.py:410)
        self._path = path
        self._log = None
        if path is not None:
            self._open_log()

    # ------------------------------------------------------------------ internals
    @property
    def dim(self) -> Optional[int]:
        return None if self._index is None else self._index.dim

    @property
    def index(self) -> Optional[DeviceIndex]:
        return self._index

    def _ensure_index(self, dim: int) -> DeviceIndex:
        if self._index is None:
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

    def _sync_filter_bits(self, row: int):
        """Mirror the row's ``filter_results_json`` "yes" answers into its device filter bits."""
        self._index.set_filter_bits(row, self._filter_bits_of(row))

    def _filter_bits_of(self, row: int) -> List[int]:
        meta = self._metas[row] or {}
        bits = []
        raw = meta.get(FILTER_JSON_KEY)
        if raw is not None:
            try:
                answers = json.loads(raw)
            except (json.JSONDecodeError, TypeError):
                answers = {}
            if isinstance(answers, dict):
                for fname, ans in answers.items():
                    b = self._filter_bit(fname, create=True)
                    if _yes(ans):
                        bits.append(b)
        return bits

    @staticmethod
    def _check_meta(m: Optional[Dict[str, Any]]):
        if m is None:
            return
        if not isinstance(m, dict):
            raise ValueError("metadata must be a dict")
        for k, v in m.items():
            if not isinstance(k, str) or not isinstance(v, _SCALAR):
                raise ValueError(f"metadata values must be str, int, float or bool (key {k!r})")

    # ------------------------------------------------------------------ persistence (append-only log)
    def _open_log(self):
        os.makedirs(self._path, exist_ok=True)
        log = os.path.join(self._path, "oplog.jsonl")
        vec = os.path.join(self._path, "vectors.f32")
        if os.path.exists(log):
            self._replay(log, vec)
        self._log = open(log, "a", encoding="utf-8")
        self._vec = open(vec, "ab")

    def _replay(self, log: str, vec: str):
        with open(log, "r", encoding="utf-8") as f, open(vec, "rb") as v:
            pend_ids, pend_rows, pend_meta, pend_doc = [], [], [], []

            def flush():
                if pend_ids:
                    self._add_rows(pend_ids, np.stack(pend_rows), pend_meta, pend_doc, log=False)
                    pend_ids.clear(); pend_rows.clear(); pend_meta.clear(); pend_doc.clear()

            for line in f:
                op = json.loads(line)
                if op["op"] == "add":
                    row = np.frombuffer(v.read(4 * op["dim"]), dtype=np.float32)
                    pend_ids.append(op["id"]); pend_rows.append(row)
                    pend_meta.append(op.get("metadata")); pend_doc.append(op.get("document"))
                    if len(pend_ids) >= 65536:
                        flush()
                else:
                    flush()
                    if op["op"] == "update":
                        self._update_one(op["id"], op.get("metadata"), op.get("document"), log=False)
                    elif op["op"] == "delete":
                        self._delete_ids([op["id"]], log=False)
            flush()

    def _log_op(self, op: Dict[str, Any], row: Optional[np.ndarray] = None):
        if self._log is None:
            return
        if row is not None:
            self._vec.write(np.ascontiguousarray(row, dtype=np.float32).tobytes())
        self._log.write(json.dumps(op, separators=(",", ":")) + "\n")

    def _log_flush(self):
        if self._log is not None:
            self._vec.flush()
            self._log.flush()

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
            self._add_rows([ids[i] for i in keep], rows, [metadatas[i] for i in keep],
                           [documents[i] for i in keep], log=True)

    def _add_rows(self, ids, rows, metadatas, documents, log: bool):
        ix = self._ensure_index(int(rows.shape[1]))
        first = ix.add(rows)
        assert first == len(self._ids), "host/device row bookkeeping diverged"
        if log and self._log is not None and not isinstance(rows, np.ndarray):
            rows = rows.cpu().numpy()                 # the persistence log stores the f32 rows
        elif not isinstance(rows, np.ndarray):
            rows = _NoRows(int(rows.shape[1]))
        with_bits = False
        for j, id_ in enumerate(ids):
            self._row_of[id_] = first + j
            self._ids.append(id_)
            self._metas.append(dict(metadatas[j]) if metadatas[j] is not None else None)
            self._docs.append(documents[j])
            if metadatas[j] and FILTER_JSON_KEY in metadatas[j]:
                with_bits = True
            if log:
                self._log_op({"op": "add", "id": id_, "dim": int(rows.shape[1]), "metadata": metadatas[j],
                              "document": documents[j]}, rows[j])
        if with_bits:      # one bulk copy (one collective on a ShardedIndex) instead of one call per row
            self._index.set_filter_bits_range(first, [self._filter_bits_of(first + j) for j in range(len(ids))])
        if log:
            self._log_flush()

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
        if True:
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
                        m = self._metas[r]
                        out["metadatas"][b].append(dict(m) if m is not None else None)
                    if out["documents"] is not None:
                        out["documents"][b].append(self._docs[r])
                    if out["embeddings"] is not None:
                        out["embeddings"][b].append(self._index.get_rows(r, 1)[0].tolist())
            return out

    def _passes(self, row: int, filters: Iterable[str]) -> bool:
        meta = self._metas[row] or {}
        try:
            answers = json.loads(meta[FILTER_JSON_KEY]) if FILTER_JSON_KEY in meta else {}
        except (json.JSONDecodeError, TypeError):
            answers = {}
        if not isinstance(answers, dict):
            answers = {}
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
                rows = list(range(len(self._ids)))
            else:
                if isinstance(ids, str):
                    ids = [ids]
                rows = [self._row_of[i] for i in ids if i in self._row_of]
            if offset:
                rows = rows[offset:]
            if limit is not None:
                rows = rows[:limit]
            out: Dict[str, Any] = {"ids": [self._ids[r] for r in rows], "embeddings": None, "documents": None,
                                   "metadatas": None, "uris": None, "data": None, "included": include}
            if "metadatas" in include:
                out["metadatas"] = [dict(self._metas[r]) if self._metas[r] is not None else None for r in rows]
            if "documents" in include:
                out["documents"] = [self._docs[r] for r in rows]
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
                    meta, doc = self._metas[r], self._docs[r]
                    self._delete_ids([id_], log=True)
                    self._add_rows([id_], rows[j:j + 1], [meta], [doc], log=True)
                self._update_one(id_, metadatas[j], documents[j], log=True)
            self._log_flush()

    def _update_one(self, id_: str, metadata, document, log: bool):
        r = self._row_of.get(id_)
        if r is None:
            return
        if metadata is not None:
            merged = dict(self._metas[r] or {})
            merged.update(metadata)
            self._metas[r] = merged
            if FILTER_JSON_KEY in metadata:
                self._sync_filter_bits(r)
        if document is not None:
            self._docs[r] = document
        if log and (metadata is not None or document is not None):
            self._log_op({"op": "update", "id": id_, "metadata": metadata, "document": document})

    def delete(self, ids=None, where=None, where_document=None):
        """``collection.delete(ids=all_ids)`` (backend/app/main.py:1069)."""
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
        for id_ in ids:
            r = self._row_of.pop(id_, None)
            if r is None:
                continue
            moved = self._index.remove(r)
            last = len(self._ids) - 1
            assert moved in (-1, last)
            if r != last:
                self._ids[r], self._metas[r], self._docs[r] = self._ids[last], self._metas[last], self._docs[last]
                self._row_of[self._ids[r]] = r
            self._ids.pop(); self._metas.pop(); self._docs.pop()
            if log:
                self._log_op({"op": "delete", "id": id_})

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
        """Run the sweep for one prompt and store the outcome the way the reference stores
        Moondream's answers (main.py:1010-1033): ``filter_results_json[filter_name] = "yes"|"no"``
        for every row, plus the device filter bit.  Returns the number of "yes" rows."""
        with self._lock:
            mask = self.filter_sweep(prompt_embedding, tau)[0]
            for r, hit in enumerate(mask.tolist()):
                meta = dict(self._metas[r] or {})
                try:
                    answers = json.loads(meta.get(FILTER_JSON_KEY, "{}"))
                except (json.JSONDecodeError, TypeError):
                    answers = {}
                answers[filter_name] = "yes" if hit else "no"
                self._update_one(self._ids[r], {FILTER_JSON_KEY: json.dumps(answers)}, None, log=True)
            self._log_flush()
            return int(mask.sum())

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
                self._log_flush()
                self._log.close(); self._vec.close()
                self._log = None
            if self._index is not None:
                self._index.close()
                self._index = None


class PersistentClient:
    """``chromadb.PersistentClient(path=...)`` as used at backend/app/utils.py:113 and
    init_db.py:36: collections persist under ``path/<name>/`` as an append-only operation log +
    raw float32 vectors and are re-ingested onto the GPU when reopened."""

    def __init__(self, path: str = "./chroma", *, device: int = 0, dtype: str = "f32"):
        self.path = path
        self._device, self._dtype = device, dtype
        os.makedirs(path, exist_ok=True)
        self._open: Dict[str, Collection] = {}

    def _dir(self, name: str) -> str:
        return os.path.join(self.path, name)

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
        d = self._dir(name)
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "collection.json"), "w", encoding="utf-8") as f:
            json.dump({"name": name, "metadata": metadata or {}, "dtype": self._dtype}, f)
        c = Collection(name, metadata, device=self._device, dtype=self._dtype, path=d)
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
        c = Collection(name, info.get("metadata"), device=self._device, dtype=info.get("dtype", self._dtype), path=d)
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
