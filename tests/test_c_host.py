"""The C ABI is usable from plain C (no Python, no torch): examples/c_host.c is compiled with gcc as
C99 against include/vecsearch_b200.h and linked with the in-tree library.  Without a GPU it must fail
loudly (no CPU fallback); on a B200 its rankings must equal its own brute-force loop."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "multimodal-image-similarity-search_b200")


def _build(tmp_path):
    import mmiss_b200
    mmiss_b200.load_native()                                     # the library must exist
    exe = str(tmp_path / "c_host")
    cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_host.c"), "-o", exe, "-L" + LIBDIR, "-lvecsearch_b200", "-lm",
           "-Wl,-rpath," + LIBDIR]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_is_valid_c_and_host_fails_loudly_without_gpu(tmp_path):
    exe = _build(tmp_path)
    if _has_gpu():
        pytest.skip("GPU present: covered by the gpu test")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 3 and "no CPU fallback" in out.stderr


@pytest.mark.gpu
def test_c_host_rankings_match_bruteforce(gpu, tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "rankings identical to brute force" in out.stdout
