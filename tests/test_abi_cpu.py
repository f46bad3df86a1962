"""CPU-only checks of the boundary: libhnsw_b200.so loads and exports every symbol include/hnsw_b200.h declares,
the entry points that need no device behave like the reference's, every compute entry point fails loudly without
a GPU (there is no CPU fallback), and the host-side mirror of the reference's data types (graph / params / byte
formats) follows the reference's own unit tests."""
import ctypes as C
import os
import re
import struct

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def H():
    import hnsw_rs_b200
    hnsw_rs_b200.lib()
    return hnsw_rs_b200


def _have_gpu():
    import torch
    return torch.cuda.is_available()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hnsw_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hnswb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(H):
    from hnsw_rs_b200 import _ffi
    names = declared_symbols()
    assert len(names) >= 40
    L = C.CDLL(_ffi.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} is declared in include/hnsw_b200.h but not exported"
    # the ctypes signature table covers exactly the header
    assert sorted(_ffi.SIGNATURES) == names


def test_rust_binding_declares_every_symbol():
    src = open(os.path.join(ROOT, "bindings", "rust", "hnsw-b200-sys", "src", "lib.rs")).read()
    bound = set(re.findall(r"pub fn (hnswb200_[a-z0-9_]+)\s*\(", src))
    assert bound == set(declared_symbols())


def test_rust_crates_resolve_the_reference_callers_imports():
    """The `use` paths of the reference's own caller (eval_glove/src/main.rs:8-15) and the items its code touches
    (:37-41 HNSW::new / insert_bulk / insert_vec / ann_by_vector, :104-105 get_point / get_vals) exist in the drop-in
    crates under bindings/rust/ with the reference's crate names and module paths.  (Source-level check: there is no Rust
    toolchain in this image.)"""
    rs = os.path.join(ROOT, "bindings", "rust")

    def src(*p):
        return open(os.path.join(rs, *p)).read()

    def crate_name(d):
        return re.search(r'^name\s*=\s*"([^"]+)"', src(d, "Cargo.toml"), re.M).group(1)

    for d in ("hnsw", "vectors", "points", "graph"):
        assert crate_name(d) == d
    # use hnsw::helpers::args::parse_args_eval; use hnsw::helpers::glove::load_glove_array; use hnsw::template::HNSW;
    assert re.search(r"pub mod helpers;", src("hnsw", "src", "lib.rs")) and re.search(r"pub mod template;", src("hnsw", "src", "lib.rs"))
    assert re.search(r"pub mod params;", src("hnsw", "src", "lib.rs"))
    assert re.search(r"pub mod args;", src("hnsw", "src", "helpers.rs")) and re.search(r"pub mod glove;", src("hnsw", "src", "helpers.rs"))
    assert re.search(r"pub fn parse_args_eval\(", src("hnsw", "src", "helpers", "args.rs"))
    glove = src("hnsw", "src", "helpers", "glove.rs")
    assert re.search(r"pub fn load_glove_array\(", glove) and re.search(r"pub fn brute_force_nns\(", glove)
    tpl = src("hnsw", "src", "template.rs")
    assert re.search(r"pub struct HNSW\b", tpl) and "pub params: Params" in tpl
    for fn in ("new", "insert_bulk", "insert_vec", "ann_by_vector", "save", "load", "len", "distance", "get_point", "get_layer",
               "layer_degrees", "assert_param_compliance"):
        assert re.search(r"pub fn %s\(" % fn, tpl), fn
    assert "unsafe impl Sync" not in tpl                      # the context sits behind a Mutex instead
    assert "Mutex<Handles>" in tpl and "q.len() != dim" in tpl  # ... and ann_batch validates every query length
    prm = src("hnsw", "src", "params.rs")
    assert re.search(r"pub struct Params\b", prm) and all(re.search(r"pub fn %s\(" % f, prm) for f in ("from_m", "from_m_efcons", "from"))
    # use vectors::VecBase;  (+ the items of SURVEY 8b)
    vec = src("vectors", "src", "lib.rs")
    assert re.search(r"pub trait VecBase\b", vec) and re.search(r"pub struct QuantVec\b", vec) and re.search(r"pub struct FullVec\b", vec)
    for fn in ("new", "dim", "iter_vals", "distance", "dist2other", "dist2many", "get_vals"):
        assert re.search(r"fn %s\b" % fn, vec), fn
    assert re.search(r"pub trait Serializer\b", src("vectors", "src", "serializer.rs"))
    pts = src("points", "src", "points.rs")
    assert re.search(r"pub trait Points\b", pts) and re.search(r"pub struct SimplePoints\b", pts)
    assert re.search(r"pub struct Point\b", src("points", "src", "point.rs"))
    g = src("graph", "src", "lib.rs")
    assert "pub type NodeID = u32" in g and re.search(r"pub struct Dist\b", g) and re.search(r"pub enum GraphError\b", g)
    assert re.search(r"pub struct Graph\b", g) and re.search(r"pub fn neighbors_vec\(", g)


def test_version_and_params_default(H):  # hnsw/src/params.rs:15-44
    from hnsw_rs_b200 import _ffi
    L = H.lib()
    assert L.hnswb200_version() == 100
    p = _ffi.Params()
    L.hnswb200_params_default(12, -1, 50, C.byref(p))
    assert (p.m, p.mmax, p.mmax0, p.ef_cons, p.dim, p.ep) == (12, 12, 24, 24, 50, 0)
    assert np.float32(p.ml) == np.float32(1.0) / np.log(np.float32(12.0))
    L.hnswb200_params_default(16, 200, 100, C.byref(p))
    assert (p.m, p.mmax0, p.ef_cons) == (16, 32, 200)


def test_load_glove_matches_oracle_parser(H, oracle):  # helpers/glove.rs:14-71
    words, emb = H.load_glove_array(0, os.path.join(GOLDEN, "store.txt"))
    ref = oracle.load_glove(os.path.join(GOLDEN, "store.txt"))
    assert emb.shape == (1000, 50) and len(words) == 1000
    assert np.array_equal(emb.view(np.uint32), ref.view(np.uint32))
    words, emb = H.load_glove_array(7, os.path.join(GOLDEN, "queries.txt"))
    assert emb.shape == (7, 50)
    with pytest.raises(H.HnswB200Error):
        H.load_glove_array(0, os.path.join(GOLDEN, "no_such_file.txt"))


def test_no_cpu_fallback(H):
    """Without a CUDA device the context cannot be created and nothing computes."""
    if _have_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(H.HnswB200Error) as e:
        H.Context(0)
    assert e.value.code == -2 and e.value.msg
    with pytest.raises(H.HnswB200Error):
        H.quantise_rows(np.zeros((2, 8), np.float32))
    with pytest.raises(H.HnswB200Error):
        H.HNSW.new(12, None, 8).insert_bulk(np.zeros((4, 8), np.float32))
    # NULL handles are rejected, not dereferenced
    L = H.lib()
    assert L.hnswb200_search(None, None, None, 1, 8, 1, 1, None, None, None, None) == -1
    assert L.hnswb200_bruteforce_topk(None, None, None, 1, 1, 0, None, None) == -1
    assert L.hnswb200_index_len(None) == 0 and L.hnswb200_points_dim(None) == 0


def test_missing_library_is_an_import_error(monkeypatch):
    from hnsw_rs_b200 import _ffi
    monkeypatch.setattr(_ffi, "_lib", None)
    monkeypatch.setattr(_ffi, "LIB_PATH", "/nonexistent/libhnsw_b200.so")
    with pytest.raises(ImportError):
        _ffi.lib()


# ---- host-side mirror of the graph crate (graph/src/graph.rs:299-486, dist.rs, layers.rs) ------------------
def _simple_graph(H):  # graph.rs:278-290
    g = H.Graph(0, 8)
    for i in range(6):
        g.add_node(i)
    for a, b in [(0, 1), (0, 2), (1, 2), (2, 3), (3, 4), (4, 5)]:
        g.add_edge(a, b)
    return g


def test_graph_edges_are_symmetric_and_checked(H):
    g = _simple_graph(H)
    for a in g.iter_nodes():
        for b in g.neighbors(a):
            assert a in g.neighbors(b)
    with pytest.raises(H.GraphError) as e:
        g.add_edge(1, 1)
    assert e.value.kind == H.GraphError.SelfConnection
    with pytest.raises(H.GraphError) as e:
        g.add_edge(1, 99)
    assert e.value.kind == H.GraphError.NodeNotInGraph and e.value.node == 99
    with pytest.raises(H.GraphError):
        g.neighbors(42)
    g.remove_edge(0, 1)
    assert 1 not in g.neighbors(0) and 0 not in g.neighbors(1)


def test_graph_replace_neighbors_keeps_degree_one_neighbours(H):  # graph.rs:85-94,128-137,385-432
    g = _simple_graph(H)
    g.replace_neighbors(4, [0, 1])
    # 5 had degree 1: isolate_node refuses to cut it off
    assert g.neighbors(4) == {0, 1, 5}
    assert 4 in g.neighbors(0) and 4 in g.neighbors(1) and 4 not in g.neighbors(3)


def test_graph_serialisation_round_trip(H):  # graph.rs:201-252,440-460
    g = _simple_graph(H)
    data = g.serialize()
    assert len(data) == g.size() == 7 + 6 * 4 * 9
    assert data[0] == 0 and struct.unpack(">I", data[1:5])[0] == 6 and struct.unpack(">H", data[5:7])[0] == 8
    h = H.Graph.deserialize(data)
    assert h.level == 0 and h.m == 8 and {n: h.neighbors(n) for n in h.iter_nodes()} == {n: g.neighbors(n) for n in g.iter_nodes()}


def test_layers_caps_and_levels(H):  # layers.rs:48-70
    ly = H.Layers(12)
    ly.add_node(3, 2)
    assert len(ly) == 3 and [l.m for l in ly.iter_layers()] == [24, 12, 12]
    assert all(l.contains(3) for l in ly.iter_layers())
    ly.add_node(4, 0)
    assert ly.get_layer(0).contains(4) and not ly.get_layer(1).contains(4)
    with pytest.raises(IndexError):
        ly.get_layer(7)


def test_dist_order_breaks_ties_by_id(H):  # graph/src/dist.rs:16-37, results.rs:223-231
    a, b, c = H.Dist(1, 0.5), H.Dist(2, 0.5), H.Dist(0, 0.75)
    assert a < b < c and a != b and a == H.Dist(1, 0.5)
    assert len({a, b, H.Dist(1, 0.5)}) == 2
    with pytest.raises(ValueError):
        H.Dist(0, float("nan")) < a  # the reference panics in partial_cmp().unwrap()


def test_params_and_fullvec_byte_formats(H):  # params.rs:64-114, full.rs:44-70 (big-endian)
    p = H.Params.from_m(12, 50)
    assert (p.m, p.mmax, p.mmax0, p.ef_cons, p.dim) == (12, 12, 24, 24, 50)
    raw = p.serialize()
    assert len(raw) == 52 and struct.unpack(">Q", raw[:8])[0] == 12
    q = H.Params.deserialize(raw)
    assert (q.m, q.mmax0, q.ef_cons, q.dim, q.ep) == (p.m, p.mmax0, p.ef_cons, p.dim, p.ep)
    v = H.FullVec([1.5, -2.0, 0.25])
    assert v.serialize() == struct.pack(">3f", 1.5, -2.0, 0.25) and v.size() == 12
    assert list(H.FullVec.deserialize(v.serialize()).get_vals()) == [1.5, -2.0, 0.25]


def test_new_layer_distribution(H):  # points.rs:148-160: floor(-ln(u) * ml), u in (0,1)
    class Seq:
        def __init__(self, vals):
            self.vals = list(vals)

        def random(self):
            return self.vals.pop(0)
    ml = H.get_default_ml(12)
    assert H.new_layer(ml, Seq([0.0, 1.0, 0.5])) == int(np.floor(-np.log(np.float32(0.5)) * np.float32(ml)))  # 0 and 1 are redrawn
    assert H.new_layer(ml, Seq([0.9])) == 0
    assert H.new_layer(ml, Seq([1e-6])) == 5


def test_public_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/hnsw_b200.h must compile as C99 (no C++-isms, no torch types), and link
    against the shared library from a C translation unit."""
    import subprocess
    src = tmp_path / "use.c"
    src.write_text('#include "hnsw_b200.h"\n'
                   'int main(void) { hnswb200_ctx* c = 0; int rc = hnswb200_ctx_create(0, &c);\n'
                   '  if (rc == 0) { hnswb200_ctx_set_vec_type(c, HNSWB200_VEC_FULL); hnswb200_ctx_destroy(c); }\n'
                   '  return hnswb200_version() > 0 ? 0 : 1; }\n')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, "hnsw_rs_b200")
    exe = tmp_path / "use"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                           str(src), "-o", str(exe), "-L", lib_dir, "-lhnsw_b200", "-Wl,-rpath," + lib_dir])
    assert subprocess.run([str(exe)]).returncode == 0   # runs with or without a device: ctx_create just fails without one
