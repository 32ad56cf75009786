"""CPU: the parts of bench.py a reader has to trust without a GPU -- the planted-row parity check it runs at every N, the
ncu traffic table lookup, the reference (CPU) arm's JSON line.  The planted check is exercised against the ORACLE's exact
top-k on a small corpus: it must accept the right answer and name the first problem of a wrong one."""
import importlib.util
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import cosine_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("dtype,tol", [("f32", 1e-5), ("bf16", 2e-3)])
def test_planted_rows_are_what_the_oracle_ranks_first(bench, dtype, tol):
    n, d, k = 6000, 512, 10
    rng = np.random.default_rng(5)
    X = rng.standard_normal((n, d)).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    q, plants = bench.parity_plan(n, d)
    for order in plants:
        rows = [g for g, _ in order]
        assert len(set(rows)) == len(rows) == len(bench.PLANT_EPS) + 2 and all(0 <= g < n for g in rows)
        for g, v in order:
            X[g] = v
    # the planted rows are spread over the whole corpus (every shard of a row-sharded collection holds some)
    all_rows = sorted(g for order in plants for g, _ in order)
    assert all_rows[0] < n // 4 and all_rows[-1] > 3 * n // 4
    want = bench.expected_scores(q, plants, dtype)
    s, r = O.cosine_topk(q, X, k, corpus_dtype=dtype)
    ok, why = bench.check_planted(np.asarray(r), np.asarray(s), plants, want, tol)
    assert ok, why
    # the exact tie is ordered by ascending global row
    for j, order in enumerate(plants):
        tie = [g for g, _ in order if g in (1000 + j, n - 1000 - j)]
        assert tie == sorted(tie) and len(tie) == 2
    # a wrong answer is caught: swap two ranks / perturb one score
    r_bad = np.asarray(r).copy()
    r_bad[3][[0, 1]] = r_bad[3][[1, 0]]
    ok, why = bench.check_planted(r_bad, np.asarray(s), plants, want, tol)
    assert not ok and why.startswith("query 3: rows")
    s_bad = np.asarray(s).copy()
    s_bad[6][2] += 10 * tol
    ok, why = bench.check_planted(np.asarray(r), s_bad, plants, want, tol)
    assert not ok and why.startswith("query 6: score error")


def test_traffic_table_lookup(bench):
    cold = bench.ncu_traffic_bytes(10_000_000, 512, "bf16")
    warm = bench.ncu_traffic_bytes(10_000_000, 512, "bf16", "back_to_back_read")
    algorithmic = 10_000_000 * (512 * 2 + 4)
    assert 1.0 <= cold / algorithmic < 1.01            # no wasted re-reads
    assert 0.99 < warm / algorithmic < 1.0             # the L2-resident slice never reaches HBM
    assert bench.ncu_traffic_bytes(123, 512, "bf16") is None


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--rows", "20000", "--steps", "1",
                          "--warmup", "0", "--hnsw-rows", "0"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "queries/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["config"]["workload"].startswith("20000x512")
