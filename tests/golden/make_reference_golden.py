"""Generate tests/golden/reference_golden.json by running the REAL reference code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_reference_golden.py

What is pinned.  The reference's hot path ends in ``chromadb.Collection.query`` (third-party,
not installable here), so that call is replaced by a *recording stub* that returns canned
distances.  Everything around it is the reference's own, unmodified source, imported from
``/root/reference/backend/app/main.py``:

* ``search_multimodal`` (main.py:829-867): the blend arithmetic -- we record the exact
  float32 vector it hands to ``collection.query``.
* ``search_similar`` (main.py:748-805): limit rule (:757), the kwargs it passes to
  ``collection.query`` (:761-765), the ``1 - d/2`` map (:782), result assembly (:785-798).
* the filter-application pass of ``search_by_text_route`` (main.py:257-278), driven through
  the real FastAPI route function.
* ``process_image``'s duplicate check (main.py:627-640) via ``collection.get(ids=[id])``.

Third-party modules that are absent (chromadb, imagehash, rembg, moondream, pillow_avif) are
stubbed in ``sys.modules`` before import; the CLIP encoder is replaced by a deterministic
seeded stub (no weights offline).  Nothing from /root/reference is copied into this repo.
"""
from __future__ import annotations

import asyncio
import json
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.json")
D = 768  # LongCLIP ViT-L/14 width used by the reference (backend/app/utils.py:16)


def _stub_modules():
    for name in ("imagehash", "rembg", "chromadb", "moondream", "pillow_avif"):
        m = types.ModuleType(name)
        sys.modules[name] = m
    sys.modules["rembg"].remove = lambda img: img
    sys.modules["imagehash"].phash = lambda img: "00ff00ff00ff00ff"
    sys.modules["chromadb"].PersistentClient = lambda path: None


class RecordingCollection:
    """Stands where chromadb's Collection stands; records calls, returns canned results."""

    def __init__(self):
        self.calls = []
        self.canned = None
        self.rows = {}

    def query(self, **kw):
        self.calls.append(("query", {k: v for k, v in kw.items()}))
        return self.canned

    def get(self, ids=None, include=None):
        self.calls.append(("get", {"ids": ids, "include": include}))
        hit = [i for i in (ids or list(self.rows)) if i in self.rows]
        out = {"ids": hit}
        if include and "metadatas" in include:
            out["metadatas"] = [self.rows[i] for i in hit]
        return out


def f32_list(a):
    return [float(np.float32(x)) for x in np.asarray(a).ravel()]


def bits_list(a):
    return np.asarray(a, dtype=np.float32).view(np.uint32).ravel().tolist()


def main():
    _stub_modules()
    work = tempfile.mkdtemp(prefix="refgold_")
    os.chdir(work)  # the reference creates static/ and chroma_data/ relative to CWD
    sys.path.insert(0, REF)
    import backend.app.main as ref  # the real reference module

    rng = np.random.default_rng(20261018)
    golden = {"dim": D, "blend": [], "search_similar": [], "filter_pass": [], "duplicate_check": []}

    # ---- search_multimodal: record the embedding it sends to collection.query ----------
    coll = RecordingCollection()
    ref.collection = coll
    ref.load_clip_model = lambda: (None, None)
    for case, w in enumerate([0.5, 0.0, 1.0, 0.25, 0.9, 1.5, -0.25]):
        img = rng.standard_normal(D).astype(np.float32) * np.float32(rng.uniform(0.1, 7.0))
        txt = rng.standard_normal(D).astype(np.float32) * np.float32(rng.uniform(0.1, 7.0))

        def fake_embed(image=None, text=None, model=None, processor=None, _i=img, _t=txt):
            out = {}
            if image is not None:
                out["image"] = _i[None].copy()
            if text is not None:
                out["text"] = _t[None].copy()
            return out

        ref.generate_clip_embedding = fake_embed
        coll.calls.clear()
        coll.canned = {"ids": [[]], "metadatas": [[]], "distances": [[]]}
        ref.search_multimodal(image=object(), query_text="q", weight_image=w, limit=7)
        (name, kw), = coll.calls
        assert name == "query"
        sent = np.asarray(kw["query_embeddings"][0], dtype=np.float64)
        golden["blend"].append({
            "weight_image": w,
            "image_bits": bits_list(img), "text_bits": bits_list(txt),
            # the reference calls ndarray.tolist(): f32 values widened to python floats
            "sent_is_f32_exact": bool(np.all(sent == sent.astype(np.float32))),
            "sent_bits": bits_list(sent.astype(np.float32)),
            "n_results": kw["n_results"], "include": kw["include"],
        })

    # ---- search_similar: limit rule, kwargs, score map, result assembly ----------------
    for limit in (10, 5, 0, -3, 1000):
        n = 4
        ids = [f"img_{i:016x}" for i in range(n)]
        metas = [{"id": ids[0], "filename": "a.jpg", "url": "/static/uploads/a.jpg",
                  "thumbnail_url": "/static/processed/a.png"},
                 {"id": ids[1], "filename": "b.jpg"},
                 {"id": ids[2], "filename": "c.jpg", "url": "/u/c"},
                 {"id": ids[3], "filename": "d.jpg", "filter_results_json": "{\"x\": \"yes\"}"}]
        dists = [float(np.float32(x)) for x in np.sort(rng.uniform(0.0, 2.0, n)).astype(np.float32)]
        coll.calls.clear()
        coll.canned = {"ids": [ids], "metadatas": [metas], "distances": [dists]}
        emb = rng.standard_normal(D).astype(np.float32)
        res = ref.search_similar(embedding=emb, limit=limit)
        (name, kw), = coll.calls
        golden["search_similar"].append({
            "limit": limit, "n_results": kw["n_results"], "include": kw["include"],
            "query_kwargs": sorted(kw.keys()),
            "ids": ids, "metadatas": metas, "distances": dists, "results": res,
        })
    # empty collection behaviour
    coll.canned = {"ids": [], "metadatas": [], "distances": []}
    golden["search_similar_empty"] = ref.search_similar(embedding=np.zeros(D, np.float32), limit=3)

    # ---- filter-application pass through the real route function -----------------------
    fr = [json.dumps({"is red?": "yes", "has cord?": "no"}),
          json.dumps({"is red?": " Yes ", "has cord?": "YES"}),
          "not json", None,
          json.dumps({"is red?": "yes"}),
          json.dumps({"is red?": "no", "has cord?": "yes"})]
    n = len(fr)
    ids = [f"img_{i:016x}" for i in range(n)]
    metas = []
    for i, f in enumerate(fr):
        m = {"id": ids[i], "filename": f"{i}.jpg", "url": f"/static/uploads/{i}.jpg",
             "thumbnail_url": f"/static/processed/{i}.png"}
        if f is not None:
            m["filter_results_json"] = f
        metas.append(m)
    dists = [0.1 * (i + 1) for i in range(n)]
    ref.generate_clip_embedding = lambda image=None, text=None, model=None, processor=None: {
        "text": np.ones((1, D), np.float32) / np.float32(np.sqrt(D))}
    for filters in ([], ["is red?"], ["is red?", "has cord?"], ["unknown"], ["has cord?"]):
        coll.calls.clear()
        coll.canned = {"ids": [ids], "metadatas": [metas], "distances": [dists]}
        out = asyncio.run(ref.search_by_text_route(query="drill", filters=filters, limit=10))
        golden["filter_pass"].append({"filters": filters, "input_metadatas": metas,
                                      "distances": dists, "kept_ids": [r["id"] for r in out["results"]],
                                      "results": out["results"]})

    # ---- duplicate check (id lookup) ----------------------------------------------------
    ref.generate_image_hash = lambda image: "img_00ff00ff00ff00ff"
    coll.rows = {"img_00ff00ff00ff00ff": {"id": "img_00ff00ff00ff00ff", "filename": "dup.jpg"}}
    coll.calls.clear()
    meta, is_new = ref.process_image(image=object(), filename="again.jpg")
    golden["duplicate_check"].append({"existing_id": "img_00ff00ff00ff00ff", "returned": meta,
                                      "is_new": is_new, "calls": [c[0] for c in coll.calls],
                                      "get_kwargs": coll.calls[0][1]})

    with open(OUT, "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
