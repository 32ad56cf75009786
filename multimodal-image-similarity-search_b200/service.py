"""Host-side mirror of the reference's service functions that sit directly on the hot path.

Same names, argument meaning and error behaviour as ``/root/reference/backend/app/main.py``
(``search_similar`` :748-805, ``search_by_text`` :807-827, ``search_multimodal`` :829-867, the
filter-application pass :201-222, the duplicate check of ``process_image`` :627-640), so the
parity tests read like tests of the reference.  The CLIP encoder is NOT part of this package
(north_star keeps it on the reference's PyTorch model): it is injected as a callable with the
signature of ``generate_clip_embedding`` (backend/app/utils.py:59-102).
"""
from __future__ import annotations

import json
import logging
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

logger = logging.getLogger("mmiss_b200")

Encoder = Callable[..., Dict[str, np.ndarray]]   # (image=None, text=None) -> {"image": [1,D], "text": [1,D]}


def similarity_from_distance(distances: Sequence[float], legacy: bool = False) -> List[float]:
    """``1 - d/2`` (backend/app/main.py:782) or the legacy ``1 - d`` (app.py:326)."""
    if legacy:
        return [1.0 - distance for distance in distances]
    return [1 - (distance / 2) for distance in distances]


def apply_filters(results: List[Dict[str, Any]], filters: Optional[Sequence[str]]) -> List[Dict[str, Any]]:
    """The filter-application pass shared by the three search routes
    (backend/app/main.py:201-222 == :257-278 == :319-340): a *post*-filter over ranked results."""
    if not filters:
        return results
    filtered_results = []
    for r in results:
        filter_results = {}
        if "filter_results_json" in r:
            try:
                filter_results = json.loads(r["filter_results_json"])
            except (json.JSONDecodeError, TypeError):
                logger.warning("Error parsing filter_results_json for image %s", r.get("id"))
        if all(filter_results.get(f, "").lower().strip() == "yes" for f in filters):
            filtered_results.append(r)
    return filtered_results


class SearchService:
    """``collection`` is a :class:`mmiss_b200.Collection` (or anything chromadb-shaped)."""

    def __init__(self, collection, encoder: Optional[Encoder] = None, legacy_scores: bool = False):
        self.collection = collection
        self.encoder = encoder
        self.legacy_scores = legacy_scores

    # -- backend/app/main.py:748-805 -------------------------------------------------------
    def search_similar(self, embedding: np.ndarray, limit: int = 10) -> List[Dict]:
        try:
            actual_limit = 1000 if limit <= 0 else limit          # main.py:757 ("All" option)
            results = self.collection.query(
                query_embeddings=[np.asarray(embedding).tolist()],
                n_results=actual_limit,
                include=["metadatas", "distances"],
            )
            return self._assemble(results)
        except Exception as e:  # the reference swallows and returns [] (main.py:803-805)
            logger.error("Error searching for similar images: %s", e)
            return []

    def _assemble(self, results) -> List[Dict]:
        """Result assembly of search_similar (main.py:767-801)."""
        if not results or "ids" not in results or not results["ids"]:
            return []
        result_ids = results["ids"][0]
        result_metadatas = results["metadatas"][0]
        result_distances = results["distances"][0]
        similarities = similarity_from_distance(result_distances, self.legacy_scores)
        similar_images = []
        for i, img_id in enumerate(result_ids):
            result_metadata = (result_metadatas[i] or {}).copy()
            result_metadata["similarity_score"] = similarities[i]
            if "url" not in result_metadata:
                result_metadata["url"] = f"/static/processed/{img_id}.png"
            if "thumbnail_url" not in result_metadata:
                result_metadata["thumbnail_url"] = f"/static/processed/{img_id}.png"
            similar_images.append(result_metadata)
        return similar_images

    # -- backend/app/main.py:807-827 -------------------------------------------------------
    def search_by_text(self, query_text: str, limit: int = 10) -> List[Dict]:
        try:
            text_embedding = self.encoder(text=query_text)["text"][0]
            return self.search_similar(embedding=text_embedding, limit=limit)
        except Exception as e:
            logger.error("Error in text search: %s", e)
            return []

    def search_by_image(self, image, limit: int = 10) -> List[Dict]:
        """Body of the /api/search/image route (backend/app/main.py:193-199)."""
        try:
            image_embedding = self.encoder(image=image)["image"][0]
            return self.search_similar(embedding=image_embedding, limit=limit)
        except Exception as e:
            logger.error("Error in image search: %s", e)
            return []

    # -- backend/app/main.py:829-867 -------------------------------------------------------
    def search_multimodal(self, image, query_text: str, weight_image: float = 0.5, limit: int = 10) -> List[Dict]:
        try:
            image_embedding = self.encoder(image=image)["image"][0]
            text_embedding = self.encoder(text=query_text)["text"][0]
            # blend (main.py:850-860) + query run on the device in one call
            actual_limit = 1000 if limit <= 0 else limit
            results = self.collection.query_multimodal(
                image_embeddings=[np.asarray(image_embedding).tolist()],
                text_embeddings=[np.asarray(text_embedding).tolist()],
                weight_image=weight_image, n_results=actual_limit, include=["metadatas", "distances"])
            return self._assemble(results)
        except Exception as e:
            logger.error("Error in multimodal search: %s", e)
            return []

    # -- the three routes' bodies: search + filter pass --------------------------------------
    def route_search_text(self, query: str, filters: Optional[Sequence[str]] = None, limit: int = 10) -> Dict:
        results = apply_filters(self.search_by_text(query, limit), filters)
        return {"results": results}

    def route_search_image(self, image, filters: Optional[Sequence[str]] = None, limit: int = 10) -> Dict:
        return {"results": apply_filters(self.search_by_image(image, limit), filters)}

    def route_search_multimodal(self, image, query: str, weight_image: float = 0.5,
                                filters: Optional[Sequence[str]] = None, limit: int = 10) -> Dict:
        return {"results": apply_filters(self.search_multimodal(image, query, weight_image, limit), filters)}

    # -- duplicate check + ingest: process_image (backend/app/main.py:627-640, 686-744) ------------
    def add_embedding(self, image_id: str, embedding: np.ndarray, metadata: Dict[str, Any],
                      description: Optional[str] = None) -> Tuple[Dict[str, Any], bool]:
        """Returns ``(metadata, is_new)``; an id that already exists returns the stored metadata and
        ``False`` (the route turns that into HTTP 409, main.py:157-168)."""
        existing_check = self.collection.get(ids=[image_id], include=["metadatas"])
        if existing_check and existing_check["ids"]:
            metadata_idx = existing_check["ids"].index(image_id)
            return existing_check["metadatas"][metadata_idx], False
        self.collection.add(ids=[image_id], embeddings=[np.asarray(embedding).tolist()], metadatas=[metadata],
                            documents=[description])
        return metadata, True


    # -- filter sweep: process_filter_on_all_images (backend/app/main.py:939-1056) ------------------
    def process_filter_on_all_images(self, filter_query: str, tau: float = 0.25, prompt_embedding=None) -> None:
        """Same name, progress contract and error behaviour as the reference's background task
        (main.py:939-1056; started from POST /api/filters at :410): the reference asks Moondream one
        image at a time; here the filter text is CLIP-encoded ONCE and swept over every stored image
        embedding on tensor cores (K3); rows with ``cos >= tau`` answer "yes".  The outcome lands where the
        post-filter (main.py:215) and the pre-filter kernels read it (``Collection.apply_filter_sweep``).
        Progress: ``get_filter_progress`` / ``collection.filter_progress`` (GET /api/filter-progress)."""
        try:
            if prompt_embedding is None:
                if self.encoder is None:
                    logger.error("CLIP encoder not available, cannot process filter")
                    self.collection.filter_progress[filter_query] = {"status": "error", "message": "Model not available",
                                                                     "progress": 0}
                    return
                prompt_embedding = self.encoder(text=filter_query)["text"][0]
            self.collection.apply_filter_sweep(filter_query, prompt_embedding, tau)
        except Exception as e:       # the reference records the error and returns (main.py:1049-1056)
            logger.error("Error processing filter on all images: %s", e)
            self.collection.filter_progress[filter_query] = {"status": "error", "message": str(e), "progress": 0}

    def get_filter_progress(self, filter_query: str) -> Dict[str, Any]:
        """Body of GET /api/filter-progress (main.py:1100-1108)."""
        progress = getattr(self.collection, "filter_progress", {})
        if filter_query not in progress:
            return {"status": "not_found"}
        return progress[filter_query]

    # -- reset_system (backend/app/main.py:1058-1098): delete every id in ONE call ----------------------
    def reset_system(self) -> bool:
        try:
            all_ids = self.collection.get(include=[])["ids"]
            if all_ids:
                self.collection.delete(ids=all_ids)
            return True
        except Exception as e:
            logger.error("Error during system reset: %s", e)
            return False


class MicroBatcher:
    """Server-side micro-batching of concurrent single queries: the
    reference answers one request at a time (`search_similar`, backend/app/main.py:748-805, called
    inline from the async routes); under load the requests that arrive within ``max_wait_ms`` of each
    other are stacked into ONE ``collection.query`` so the batched tcgen05 kernel (K2) serves them in
    a single pass over the corpus.  ``mode="scan"`` (default) keeps every caller's answer bit-identical to
    its own ``collection.query(query_embeddings=[e], n_results=n, include=...)``; ``mode="auto"`` lets
    batches of >= 16 queries on a bf16 collection take the tensor path, whose queries are rounded to
    bf16: scores then agree within the bf16 tolerance (2e-3) and near-ties may order differently."""

    def __init__(self, collection, max_batch: int = 64, max_wait_ms: float = 2.0,
                 include: Sequence[str] = ("metadatas", "distances"), mode: str = "scan"):
        import threading
        self.collection, self.max_batch, self.max_wait = collection, int(max_batch), max_wait_ms / 1e3
        self.mode = mode
        self.include = list(include)
        self._cv = threading.Condition()
        self._pending: List[Dict[str, Any]] = []
        self._closed = False
        self.batches: List[int] = []                      # sizes of the batches run so far (introspection)
        self._worker = threading.Thread(target=self._run, daemon=True)
        self._worker.start()

    def query(self, embedding, n_results: int = 10) -> Dict[str, Any]:
        """Blocking; thread-safe.  Returns the chromadb-shaped dict for this one query."""
        import threading
        req = {"e": np.asarray(embedding, dtype=np.float32).reshape(-1), "n": int(n_results),
               "done": threading.Event(), "out": None, "err": None}
        with self._cv:
            if self._closed:
                raise RuntimeError("MicroBatcher is closed")
            self._pending.append(req)
            self._cv.notify_all()
        req["done"].wait()
        if req["err"] is not None:
            raise req["err"]
        return req["out"]

    def _run(self):
        import time
        while True:
            with self._cv:
                while not self._pending and not self._closed:
                    self._cv.wait()
                if self._closed and not self._pending:
                    return
                deadline = time.monotonic() + self.max_wait
                while len(self._pending) < self.max_batch and not self._closed:
                    left = deadline - time.monotonic()
                    if left <= 0:
                        break
                    self._cv.wait(left)
                batch, self._pending = self._pending[:self.max_batch], self._pending[self.max_batch:]
            try:
                n_max = max(r["n"] for r in batch)
                res = self.collection.query(query_embeddings=np.stack([r["e"] for r in batch]), n_results=n_max,
                                            include=self.include, mode=self.mode)
                self.batches.append(len(batch))
                for b, r in enumerate(batch):
                    one = dict(res)
                    for key in ("ids", "distances", "metadatas", "documents", "embeddings"):
                        if res.get(key) is not None:
                            one[key] = [res[key][b][:r["n"]]]
                    r["out"] = one
            except Exception as e:                        # noqa: BLE001 -- every waiter gets the error
                for r in batch:
                    r["err"] = e
            for r in batch:
                r["done"].set()

    def close(self):
        with self._cv:
            self._closed = True
            self._cv.notify_all()
        self._worker.join(timeout=5)
