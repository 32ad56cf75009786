"""Independent third-party pin of the cosine-space definition the reference selects with
``metadata={"hnsw:space": "cosine"}`` (/root/reference/backend/app/utils.py:129).

chromadb / hnswlib themselves cannot be installed here (no wheel, no network), so the top-k of
``collection.query`` stays "chromadb unpinned".  What CAN be pinned is the definition it implements --
cosine distance ``d = 1 - <q, x> / (|q| |x|)``, neighbours in ascending distance -- against two
independent, widely used implementations that ARE installed:

* ``scipy.spatial.distance.cdist(Q, X, "cosine")``                       (float64 distances)
* ``sklearn.neighbors.NearestNeighbors(metric="cosine", algorithm="brute")``  (distances + neighbour ids)

This script runs both on the committed golden corpus (tests/golden/oracle_golden.npz: rows of very
different norms, an exact duplicate, a scaled duplicate, a zero row is EXCLUDED here because both libraries
define 0/0 differently) and writes their outputs to tests/golden/thirdparty_golden.npz;
tests/test_oracle.py checks oracle.cosine_topk / cosine_scores against them.

    python tests/golden/make_thirdparty_golden.py
"""
import os

import numpy as np
import scipy
import sklearn
from scipy.spatial.distance import cdist
from sklearn.neighbors import NearestNeighbors

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    g = np.load(os.path.join(HERE, "oracle_golden.npz"))
    X, Q, k = g["X"], g["Q"], int(g["k"])
    keep = np.flatnonzero(np.linalg.norm(X, axis=1) > 0)            # drop the zero row (cosine undefined there)
    Xk = X[keep].astype(np.float64)
    Qd = Q.astype(np.float64)
    d_scipy = cdist(Qd, Xk, metric="cosine")                         # [B, n] float64
    nn = NearestNeighbors(n_neighbors=k, metric="cosine", algorithm="brute").fit(Xk)
    d_sk, i_sk = nn.kneighbors(Qd, return_distance=True)
    np.savez_compressed(os.path.join(HERE, "thirdparty_golden.npz"), keep=keep, cdist_cosine=d_scipy,
                        sk_dist=d_sk, sk_rows=keep[i_sk], versions=np.array([scipy.__version__, sklearn.__version__]))
    print("scipy", scipy.__version__, "sklearn", sklearn.__version__, "->", d_scipy.shape, i_sk.shape)


if __name__ == "__main__":
    main()
