"""CPU, build container only (skipped where /root/reference is absent, e.g. on the GPU box):
the UNMODIFIED reference backend (`/root/reference/backend/app/main.py`) runs on top of
mmiss_b200.Collection exactly as INTEGRATION.md describes -- `import mmiss_b200 as chromadb` in its
utils module, nothing else changed.  The device index is the numpy fake (no GPU here); what is
under test is the drop-in surface: keyword names, return shapes, error behaviour."""
import asyncio
import json
import os
import sys
import types

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "backend", "app")),
                                reason="reference tree not present")
D = 768


@pytest.fixture(scope="module")
def ref(tmp_path_factory):
    import mmiss_b200
    from mmiss_b200 import collection as C
    from tests.fake_index import FakeIndex
    saved = C.DeviceIndex
    C.DeviceIndex = FakeIndex
    for name in ("imagehash", "rembg", "moondream", "pillow_avif"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["rembg"].remove = lambda img: img
    sys.modules["imagehash"].phash = lambda img: "0f" * 8
    sys.modules["chromadb"] = mmiss_b200                      # <- the integration: engine in chromadb's place
    work = tmp_path_factory.mktemp("refapp")
    cwd = os.getcwd()
    os.chdir(work)
    os.environ["CHROMA_PERSIST_DIR"] = str(work / "chroma_data")
    sys.path.insert(0, REF)
    for m in [m for m in sys.modules if m.startswith("backend")]:
        del sys.modules[m]
    import backend.app.main as main
    import backend.app.utils as utils
    main.collection = utils.init_chromadb()                   # the reference's own factory, unmodified
    yield main
    os.chdir(cwd)
    sys.path.remove(REF)
    sys.modules.pop("chromadb", None)
    C.DeviceIndex = saved


def _unit(rng):
    v = rng.standard_normal(D).astype(np.float32)
    return v / np.linalg.norm(v)


def test_reference_backend_runs_on_the_engine(ref):
    import mmiss_b200
    assert isinstance(ref.collection, mmiss_b200.Collection) and ref.collection.name == ref.COLLECTION_NAME
    rng = np.random.default_rng(0)
    vecs = {f"img_{i:016x}": _unit(rng) for i in range(12)}
    # ingest the way process_image does (main.py:735-744)
    for i, (image_id, v) in enumerate(vecs.items()):
        meta = {"id": image_id, "filename": f"{i}.jpg", "description": f"drill {i}", "custom_metadata": "",
                "url": f"/static/uploads/{i}.jpg", "thumbnail_url": f"/static/processed/{image_id}.png",
                "processed_url": f"/static/processed/{image_id}.png", "created_at": f"2024-01-{i + 1:02d}T00:00:00",
                "filter_results_json": json.dumps({"is red?": "yes" if i % 3 == 0 else "no"})}
        ref.collection.add(ids=[image_id], embeddings=[v.tolist()], metadatas=[meta], documents=[meta["description"]])
    assert ref.collection.count() == 12
    # startup path (main.py:550-579)
    ref.image_metadata.clear()
    loaded = ref.load_metadata_from_chromadb()
    assert set(loaded) == set(vecs)
    # search_similar (main.py:748-805), exact ranking + score map
    target = "img_%016x" % 5
    res = ref.search_similar(embedding=vecs[target], limit=3)
    assert res[0]["id"] == target and abs(res[0]["similarity_score"] - 1.0) < 1e-6 and len(res) == 3
    assert [r["similarity_score"] for r in res] == sorted((r["similarity_score"] for r in res), reverse=True)
    assert len(ref.search_similar(embedding=vecs[target], limit=0)) == 12          # "All" -> n_results 1000, clamped
    # text route with the filter pass (main.py:234-293)
    q = _unit(rng)
    ref.load_clip_model = lambda: (None, None)
    ref.generate_clip_embedding = lambda image=None, text=None, model=None, processor=None: {
        "text": q[None], "image": vecs[target][None]}
    out = asyncio.run(ref.search_by_text_route(query="red drill", filters=["is red?"], limit=10))
    ids = [r["id"] for r in out["results"]]
    assert ids and all(int(i[4:], 16) % 3 == 0 for i in ids)
    # multimodal service function (main.py:829-867)
    mm = ref.search_multimodal(image=object(), query_text="red", weight_image=0.25, limit=4)
    c = 0.25 * vecs[target] + 0.75 * q
    c /= np.linalg.norm(c)
    want = sorted(vecs, key=lambda i: -float(vecs[i] @ c))[:4]
    assert [r["id"] for r in mm] == want
    # duplicate check of process_image (main.py:627-640): existing id -> (metadata, False)
    ref.generate_image_hash = lambda image: target
    meta, is_new = ref.process_image(image=object(), filename="again.jpg")
    assert is_new is False and meta["filename"] == "5.jpg"
    # metadata update (main.py:503-510 passes a partial dict) and reset (main.py:1058-1098)
    ref.collection.update(ids=[target], metadatas=[{"description": "edited"}], documents=["edited"])
    assert ref.collection.get(ids=[target], include=["metadatas"])["metadatas"][0]["filename"] == "5.jpg"
    ref.reset_system()
    assert ref.collection.count() == 0 and ref.search_similar(embedding=q, limit=5) == []
