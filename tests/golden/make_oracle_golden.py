"""Small seeded known-answer vectors produced by oracle/cosine_oracle.py (committed as
tests/golden/oracle_golden.npz).  The reference pins nothing for the cosine top-k (its arithmetic
is inside the un-installable chromadb wheel), so these are OUR known answers: each is
cross-checked against an independent float64 computation before being written.

    python tests/golden/make_oracle_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cosine_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.npz")


def main():
    rng = np.random.default_rng(42)
    N, D, B, K = 700, 96, 5, 10
    X = rng.standard_normal((N, D)).astype(np.float32) * rng.uniform(0.2, 5.0, (N, 1)).astype(np.float32)
    X[17] = X[3]                 # exact duplicate rows -> exact score ties (row-order tie-break)
    X[400] = 2.5 * X[3]          # same direction, different norm
    X[55] = 0.0                  # zero row -> score 0
    Q = rng.standard_normal((B, D)).astype(np.float32)
    Q[1] = X[3] * 0.5            # query parallel to the tied rows
    out = {"X": X, "Q": Q, "k": np.int64(K)}
    for tag, cd, rq in (("f32", "f32", False), ("bf16", "bf16", False), ("bf16q", "bf16", True)):
        s, r = O.cosine_topk(Q, X, K, corpus_dtype=cd, round_queries=rq)
        full64 = O.cosine_scores(Q, X, cd, rq, accumulate="f64")
        for b in range(B):
            ok, why = O.topk_matches(s[b], r[b], full64[b], K, 2e-6)
            assert ok, (tag, b, why)
        out[f"scores_{tag}"], out[f"rows_{tag}"] = s, r
    # blend known answers (float64 check)
    img, txt = rng.standard_normal((4, D)).astype(np.float32), rng.standard_normal((4, D)).astype(np.float32)
    w = np.array([0.5, 0.0, 1.0, 0.3])
    bl = np.stack([O.blend(img[i], txt[i], float(w[i])) for i in range(4)])
    i64, t64 = img.astype(np.float64), txt.astype(np.float64)
    c = w[:, None] * i64 / np.linalg.norm(i64, axis=1, keepdims=True) + \
        (1 - w)[:, None] * t64 / np.linalg.norm(t64, axis=1, keepdims=True)
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    assert np.abs(bl - c).max() < 1e-6
    out.update(blend_img=img, blend_txt=txt, blend_w=w, blend_out=bl.astype(np.float32))
    # filter sweep + dedup (bf16, rounded prompts)
    Xd = rng.standard_normal((300, D)).astype(np.float32)
    Xd[100:110] = Xd[0:10] + 0.05 * rng.standard_normal((10, D)).astype(np.float32)
    F = rng.standard_normal((6, D)).astype(np.float32)
    out["Xd"], out["F"] = Xd, F
    out["filter_tau"] = np.float32(0.1)
    out["filter_mask"] = O.filter_mask(F, Xd, 0.1)
    di, dj, ds = O.dedup_pairs(Xd, 0.9)
    assert len(di) >= 10
    out.update(dedup_tau=np.float32(0.9), dedup_i=di, dedup_j=dj, dedup_s=ds)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
