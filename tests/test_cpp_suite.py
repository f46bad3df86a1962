"""The reference's own tests restated in C++ (tests/cpp/reference_suite.cpp) against the C++ host mirror of its API
(include/hnsw_rs.hpp) over the C ABI.  CPU: the suite builds, links and refuses to compute; GPU: it passes."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "reference_suite")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "-s"])


def test_cpp_suite_builds_and_lists():
    _build()
    out = subprocess.run([BIN, GOLDEN, "--list"], capture_output=True, text=True, check=True).stdout.split()
    assert "hnsw::hnsw_glove_build_eval" in out and "vectors::quant::distance" in out and "hnsw(FullVec)::hnsw_serialize" in out and len(out) == 14


def test_cpp_suite_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _build()
    r = subprocess.run([BIN, GOLDEN], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_cpp_reference_suite_passes():
    _build()
    r = subprocess.run([BIN, GOLDEN], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "14 passed; 0 failed" in r.stdout


def test_host_edge_store_matches_set_semantics():
    """hostgraph.h (the build's host-side Graph: add_edge / remove_edge / replace_neighbors, graph.rs:37-148) against
    std::set / std::map at degrees below, at and far above the row width; host only."""
    _build()
    r = subprocess.run([os.path.join(ROOT, "tests", "cpp", "hostgraph_test")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "all checks passed" in r.stdout, r.stdout[-2000:]


def test_record_layouts():
    """layout.h: the lane-sliced QuantVec record and the padded f32 (FullVec) record, dims 1..520; host only."""
    _build()
    r = subprocess.run([os.path.join(ROOT, "tests", "cpp", "layout_test")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "all checks passed" in r.stdout, r.stdout[-2000:]


def test_visited_set_geometry_names_every_id():
    """vis_geometry.h (shared by the launcher, the kernel's VisB4 and this test): for every id below 2^B the pair
    (home bucket, 15-bit entry) is unique -- exhaustively, B = 10..21, every bucket count the launcher can produce -- so the
    bucketed visited set of search_kernel_fast has no false positives (results.rs:101-103 stays exact); host only."""
    _build()
    r = subprocess.run([os.path.join(ROOT, "tests", "cpp", "visgeom_test")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "visgeom ok" in r.stdout, r.stdout[-2000:]
