"""Diagnostic run of the tcgen05 path on a GPU box (prints error statistics, asserts nothing)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmiss_b200 as M
from oracle import cosine_oracle as O

def run(n, d, B, k):
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((B, d)).astype(np.float32)
    ix = M.DeviceIndex(d, "bf16")
    ix.add(X)
    s, r = ix.query(Q, k, mode="tensor")
    full = O.cosine_scores(Q, X, "bf16", True)
    es, er = O.cosine_topk(Q, X, k, "bf16", True)
    bad = 0
    for b in range(B):
        ok, why = O.topk_matches(s[b][r[b] >= 0], r[b][r[b] >= 0], full[b], min(k, n), 2e-3)
        if not ok:
            bad += 1
            if bad <= 3:
                print("  q", b, why, "\n   got", r[b][:6], s[b][:6], "\n   exp", er[b][:6], es[b][:6])
    print(f"n={n} d={d} B={B} k={k}: {B - bad}/{B} queries ok")
    ix.close()

if __name__ == "__main__":
    for cfg in [(256, 64, 16, 10), (1000, 512, 16, 10), (70000, 512, 130, 10), (3000, 768, 20, 5)]:
        run(*cfg)
