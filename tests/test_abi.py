"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares;
without a GPU every compute entry point fails LOUDLY (no CPU / PyTorch fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import mmiss_b200
from mmiss_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vecsearch_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vs_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_in_tree_and_loads():
    assert os.path.dirname(N.LIB_PATH) == os.path.join(ROOT, "multimodal-image-similarity-search_b200")
    lib = mmiss_b200.load_native()
    assert lib.vs_abi_version() == 2
    assert lib.vs_launch_count() >= 0


def test_every_header_symbol_is_exported_and_bound():
    declared = _declared_symbols()
    assert len(declared) >= 25
    lib = ctypes.CDLL(N.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(N.SYMBOLS) == declared, "python binding table and header disagree"


def test_sass_is_sm100_native():
    """tcgen05 / TMA really are in the binary: UTC*MMA, LDTM, UTMALDG, UBLKCP (B200_PROFILING.md table)."""
    out = subprocess.run(["cuobjdump", "-sass", N.LIB_PATH], capture_output=True, text=True, timeout=600).stdout
    assert "sm_100a" in out
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "SYNCS"):
        assert mnemonic in out, mnemonic


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="GPU present")
def test_no_gpu_means_loud_failure_not_fallback():
    with pytest.raises(mmiss_b200.VecSearchError) as ei:
        mmiss_b200.DeviceIndex(8)
    assert "no CPU fallback" in str(ei.value)
    col = mmiss_b200.Collection("c")
    with pytest.raises(mmiss_b200.VecSearchError):
        col.add(ids=["a"], embeddings=np.zeros((1, 8), np.float32))
    assert col.count() == 0


def test_bad_arguments_are_rejected_before_touching_the_device():
    lib = mmiss_b200.load_native()
    h = ctypes.c_void_p()
    assert lib.vs_create(0, 0, 0, 0, ctypes.byref(h)) == -1            # dim = 0
    assert b"dim" in lib.vs_last_error()
    assert lib.vs_create(0, 8, 7, 0, ctypes.byref(h)) == -1            # bad dtype
    assert lib.vs_count(None) == 0 and lib.vs_destroy(None) == 0
    assert lib.vs_merge_topk_dev(None, None, None, 1, 1, 1, None, None, None) == -1
