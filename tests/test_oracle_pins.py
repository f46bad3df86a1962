"""Pins the CPU oracle to the reference's own known-answer tests, invariants and fixture test.

Each test names the reference test it restates (paths relative to the reference repo).
"""
import os

import numpy as np
import pytest

SQRT2 = np.sqrt(np.float32(2.0))


def _q(o, v):
    return o.quantise(np.asarray(v, np.float32))


# vectors/src/quant.rs:143-202  mod tests::distance (exact equalities on the quantised path)
def test_quantvec_distance_kats(oracle):
    o = oracle
    assert o.dist_quant(_q(o, [0.5]), _q(o, [0.25])) == np.float32(0.25)
    assert o.dist_quant(_q(o, [0.75]), _q(o, [0.25])) == np.float32(0.5)
    assert o.dist_quant(_q(o, [0.0, 0.0]), _q(o, [0.0, 1.0])) == np.float32(1.0)
    assert o.dist_quant(_q(o, [1.0, 0.0]), _q(o, [0.0, 1.0])) == SQRT2
    assert o.dist_quant(_q(o, [-1.0, 0.0]), _q(o, [0.0, 1.0])) == SQRT2
    assert o.dist_quant(_q(o, [1.0, 0.0]), _q(o, [0.0, -1.0])) == SQRT2
    rng = np.random.default_rng(0)
    a = _q(o, rng.random(128, dtype=np.float32))
    assert o.dist_quant(a, a) == np.float32(0.0)
    for _ in range(100):
        b = _q(o, rng.random(128, dtype=np.float32))
        assert o.dist_quant(a, b) >= 0.0
        assert o.dist_quant(a, b) == o.dist_quant(b, a)  # (x-y)^2 symmetry, SURVEY App. A


# constant vector: delta = 0 -> 0/0 = NaN -> code 0, dequantises back to `min` (quant.rs:154-166)
def test_quantiser_constant_vector(oracle):
    codes, mn, dl = oracle.quantise(np.array([0.5], np.float32))
    assert codes.tolist() == [0] and mn == np.float32(0.5) and dl == np.float32(0.0)
    assert oracle.dequantise(codes, mn, dl)[0] == np.float32(0.5)


def test_quantiser_codes_and_bounds(oracle):
    v = np.array([0.0, 1.0, 0.5, 0.25, 1.0 / 255.0], np.float32)
    codes, mn, dl = oracle.quantise(v)
    assert mn == np.float32(0.0) and dl == np.float32(1.0) / np.float32(255.0)
    assert codes[0] == 0 and codes[1] == 255
    # numpy restatement of quant.rs:41-66 op by op
    t = ((v - mn) / dl + np.float32(0.5)).astype(np.float32)
    assert codes.tolist() == np.floor(t).astype(np.uint8).tolist()
    with pytest.raises(oracle.OracleError):
        oracle.quantise(np.array([0.0, np.nan], np.float32))  # partial_cmp().unwrap() panics


# vectors/src/full.rs:88-147
def test_fullvec_distance_kats(oracle):
    o = oracle
    f = lambda a, b: o.dist_full(np.asarray(a, np.float32), np.asarray(b, np.float32))
    assert f([0.5], [0.25]) == np.float32(0.25)
    assert f([0.75], [0.25]) == np.float32(0.5)
    assert f([0.0, 0.0], [0.0, 1.0]) == np.float32(1.0)
    assert f([1.0, 0.0], [0.0, 1.0]) == SQRT2
    assert f([-1.0, 0.0], [0.0, 1.0]) == SQRT2
    assert f([1.0, 0.0], [0.0, -1.0]) == SQRT2
    a = np.random.default_rng(1).random(128, dtype=np.float32)
    assert f(a, a) == np.float32(0.0)


# numpy restatement of distance_unrolled (quant.rs:14-37) on random data, all remainder sizes
@pytest.mark.parametrize("dim", [1, 2, 7, 8, 9, 15, 16, 50, 96, 100, 128, 300])
def test_distance_unrolled_vs_numpy(oracle, dim):
    rng = np.random.default_rng(dim)
    for _ in range(20):
        a = _q(oracle, rng.normal(size=dim).astype(np.float32))
        b = _q(oracle, rng.normal(size=dim).astype(np.float32))
        x = (a[0].astype(np.float32) * a[2] + a[1]).astype(np.float32)
        y = (b[0].astype(np.float32) * b[2] + b[1]).astype(np.float32)
        sq = ((x - y).astype(np.float32) ** 2).astype(np.float32)
        acc = np.zeros(8, np.float32)
        nfull = dim // 8
        for k in range(nfull):
            acc = (acc + sq[8 * k:8 * k + 8]).astype(np.float32)
        for i in range(8 * nfull, dim):
            acc[0] = np.float32(acc[0] + sq[i])
        s = np.float32(0)
        for j in range(8):
            s = np.float32(s + acc[j])
        assert oracle.dist_quant(a, b) == np.sqrt(s)


# vectors/tests/full_lvq_tests.rs:3-27
def test_dist_err_lt_one_percent(oracle):
    rng = np.random.default_rng(7)
    for _ in range(1000):
        ra, rb = rng.random(128, dtype=np.float32), rng.random(128, dtype=np.float32)
        full = oracle.dist_full(ra, rb)
        qa, qb = _q(oracle, ra), _q(oracle, rb)
        q2f = oracle.dist_full(oracle.dequantise(*qa), rb)
        q2q = oracle.dist_quant(qa, qb, generic=True)
        assert abs(full - q2f) / full < 0.01
        assert abs(full - q2q) / full < 0.01
        assert abs(full - oracle.dist_quant(qa, qb)) / full < 0.01


# graph/src/dist.rs:16-37 and hnsw/src/template/results.rs:223-231
def test_dist_order_and_tie_rule(oracle):
    c = oracle.dist_cmp
    assert c(0, 0.5, 1, 0.6) == -1 and c(1, 0.6, 0, 0.5) == 1
    assert c(0, 0.5, 1, 0.5) == -1 and c(1, 0.5, 0, 0.5) == 1  # tie broken by id
    assert c(3, 0.5, 3, 0.5) == 0
    keys = {(0, 0.5), (1, 0.5), (2, 0.0), (4, 0.0)}
    assert len(keys) == 4  # equal distances with different ids stay distinct members


def _simple_graph(o):  # graph/src/graph.rs:278-290
    g = o.Graph(1, 12)
    for i in range(5):
        g.add_node(i)
    for a, b in [(0, 1), (0, 2), (1, 2), (2, 3), (3, 4), (4, 1)]:
        assert g.add_edge(a, b) == 0
    return g


# graph/src/graph.rs:305-340
def test_graph_symmetry_selfloops_missing(oracle):
    rng = np.random.default_rng(3)
    g = oracle.Graph(0, 12)
    for i in range(100):
        g.add_node(i)
    for i in range(100):
        for n in rng.choice(100, 8, replace=False):
            g.add_edge(i, int(n))
    for i in range(100):
        for n in g.neighbors(i):
            assert i in g.neighbors(n)
            assert n != i
    g = oracle.Graph(0, 12)
    assert not g.contains(42)
    g.add_node(42)
    assert g.contains(42) and g.nb_nodes() == 1 and g.degree(42) == 0
    g.add_node(1)
    assert g.add_edge(1, 2) != 0 and g.add_edge(999, 1000) != 0
    assert g.add_edge(42, 42) != 0
    assert g.degree(999) == -1


# graph/src/graph.rs:342-383
def test_graph_remove_and_neighbors(oracle):
    g = _simple_graph(oracle)
    assert g.remove_edge(0, 1) == 0
    assert 1 not in g.neighbors(0) and 0 not in g.neighbors(1)
    assert g.remove_edge(0, 999) != 0
    g = _simple_graph(oracle)
    assert g.neighbors(1) == {0, 2, 4}


# graph/src/graph.rs:385-432
def test_graph_replace_neighbors(oracle):
    g = oracle.Graph(0, 12)
    for i in range(6):
        g.add_node(i)
    for a, b in [(0, 1), (0, 2), (0, 3), (1, 3), (1, 2)]:
        g.add_edge(a, b)
    assert g.degree(0) == 3
    assert g.replace_neighbors(0, [4, 5]) == 0
    assert g.neighbors(0) == {4, 5}
    assert 0 not in g.neighbors(1) and 0 not in g.neighbors(2)
    assert 0 in g.neighbors(4) and 0 in g.neighbors(5)
    g = oracle.Graph(0, 12)
    g.add_node(100)
    for i in range(200, 205):
        g.add_node(i)
    assert g.replace_neighbors(100, range(200, 205)) == 0
    assert g.degree(100) == 5
    for i in range(200, 205):
        assert 100 in g.neighbors(i)


# graph.rs:85-94: isolate_node never cuts the edge to a degree-1 neighbour
def test_graph_isolate_keeps_degree_one_neighbours(oracle):
    g = oracle.Graph(0, 2)
    for i in range(5):
        g.add_node(i)
    g.add_edge(0, 1)  # 1 has degree 1
    g.add_edge(0, 2)
    g.add_edge(2, 3)
    assert g.replace_neighbors(0, [4]) == 0
    assert g.neighbors(0) == {1, 4}  # edge to 1 survives, edge to 2 is cut


# hnsw/src/template.rs:518-572 hnsw_glove_build_eval (the reference's end-to-end test)
def test_glove_build_eval(oracle, glove, glove_index):
    store, queries = glove
    ix = glove_index
    gt, _ = ix.bruteforce(queries, 10)
    ids, _, counts, hops, evals = ix.search_batch(queries, 10, 100)
    hits = sum(len(set(gt[i].tolist()) & set(ids[i, :counts[i]].tolist())) for i in range(len(queries)))
    assert hits / (len(queries) * 10) > 0.99
    for l in range(ix.nb_layers):
        if oracle.lib().oracle_layer_nb_nodes(ix.h, l) <= 1:
            continue
        mn, mx = ix.layer_degree_range(l)
        assert mn > 0
        # assert_param_compliance (template.rs:341-370) tolerates ceil(1.1 * cap)
        assert mx <= int(np.ceil(np.float32(ix.layer_cap(l)) * np.float32(1.1)))


# hnsw/src/template.rs:465-516
def test_hnsw_build_and_inserts(oracle):
    rng = np.random.default_rng(5)
    ix = oracle.Index(12, None, 10).insert_bulk(rng.random((100, 10), dtype=np.float32))
    assert len(ix) == 100
    ix.insert_vec(rng.random(10, dtype=np.float32))
    assert len(ix) == 101
    ix.insert_bulk(rng.random((100, 10), dtype=np.float32))
    assert len(ix) == 201
    with pytest.raises(oracle.OracleError):  # can_not_add_different_dim (reference panics)
        ix.insert_bulk(rng.random((10, 12), dtype=np.float32))


# hnsw/src/template.rs:574-611 hnsw_serialize + byte layout of SURVEY App. B
def test_save_load_round_trip(oracle, tmp_path):
    rng = np.random.default_rng(11)
    for it in range(5):
        ix = oracle.Index(12, None, 10).insert_bulk(rng.random((100, 10), dtype=np.float32))
        d = tmp_path / f"ix{it}"
        ix.save(d)
        assert os.path.getsize(d / "params") == 52
        assert os.path.getsize(d / "points") == 16 + 100 * (9 + 10)
        ld = oracle.Index.load(d)
        assert len(ld) == 100 and ld.ep == ix.ep and ld.nb_layers == ix.nb_layers
        for a, b in zip(ix.export_points(), ld.export_points()):
            assert np.array_equal(a, b)
        for l in range(ix.nb_layers):
            for a, b in zip(ix.export_layer(l), ld.export_layer(l)):
                assert np.array_equal(a, b)
        assert ld.params() == ix.params()


# params.rs:15-30
def test_params_defaults(oracle):
    p = oracle.Index(12, None, 50).params()
    assert (p["m"], p["mmax"], p["mmax0"], p["ef_cons"], p["dim"]) == (12, 12, 24, 24, 50)
    assert p["ml"] == np.float32(1.0) / np.log(np.float32(12.0))
    assert oracle.Index(16, 200, 8).params()["ef_cons"] == 200


# ChaCha core against the RFC 7539 section 2.3.2 block (20 rounds); the 12-round StdRng stream
# itself is UNPINNED (no reference vector) - see DESIGN.md.
def test_chacha_core_rfc7539_zero_nonce(oracle):
    # RFC 7539 A.1 test vector #1: all-zero key/nonce, counter 0
    out = oracle.chacha_block(np.zeros(8, np.uint32), 0, 20)
    expect = bytes.fromhex(
        "76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7"
        "da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586")
    assert out.astype("<u4").tobytes() == expect
    lv = oracle.levels(12, 100000)
    # level law: P(level >= 1) = 1/m for ml = 1/ln m
    assert abs((lv >= 1).mean() - 1 / 12) < 0.005


def test_search_order_independence_and_counts(oracle, glove, glove_index):
    """selected == ef smallest evaluated Dists; ef < n returns ef ids."""
    _, queries = glove
    ids, dists, counts, hops, evals = glove_index.search_batch(queries, 10, 5)
    assert (counts == 5).all()
    assert (ids[:, 5:] == 0xFFFFFFFF).all()
    ids, dists, counts, _, _ = glove_index.search_batch(queries, 10, 100)
    for i in range(len(queries)):
        k = [(dists[i, j], ids[i, j]) for j in range(counts[i])]
        assert k == sorted(k)


# ---- the oracle's FullVec mode (`type VecType = FullVec;`, points/src/point.rs:4): the same reference tests ----
def test_fullvec_mode_glove_build_eval(oracle, glove):  # template.rs:518-572 with the alias flipped
    store, queries = glove
    ix = oracle.Index(12, None, store.shape[1], full=True).insert_bulk(store)
    assert ix.full and len(ix) == 1000
    gt, gd = ix.bruteforce(queries, 10)
    # the ground truth is the f32 metric itself: FullVec::distance, one sequential sum (full.rs:23-29)
    for i in (0, 17, 99):
        s = np.float32(0)
        for a, b in zip(queries[i], store[gt[i, 0]]):
            t = np.float32(a) - np.float32(b)
            s = np.float32(s + np.float32(t * t))
        assert np.sqrt(s) == gd[i, 0]
    ids, _, counts, hops, evals = ix.search_batch(queries, 10, 100)
    hits = sum(len(set(gt[i].tolist()) & set(ids[i, :counts[i]].tolist())) for i in range(len(queries)))
    assert hits / (len(queries) * 10) > 0.99
    for l in range(ix.nb_layers):
        if oracle.lib().oracle_layer_nb_nodes(ix.h, l) <= 1:
            continue
        mn, mx = ix.layer_degree_range(l)
        assert mn > 0 and mx <= int(np.ceil(np.float32(ix.layer_cap(l)) * np.float32(1.1)))
    with pytest.raises(oracle.OracleError):  # a NaN distance makes the reference panic (graph/src/dist.rs:32)
        bad = queries[:1].copy()
        bad[0, 3] = np.nan
        ix.search_batch(bad, 10, 100)


def test_fullvec_mode_save_load_round_trip(oracle, tmp_path):  # template.rs:574-611; point size 1 + 4*dim (full.rs:45-61)
    rng = np.random.default_rng(12)
    rows = rng.random((100, 10), dtype=np.float32)
    ix = oracle.Index(12, None, 10, full=True).insert_bulk(rows)
    ix.save(tmp_path / "ix")
    blob = (tmp_path / "ix" / "points").read_bytes()
    assert len(blob) == 16 + 100 * (1 + 4 * 10)
    assert np.array_equal(np.frombuffer(blob[16 + 1:16 + 41], ">f4").astype(np.float32), rows[0])
    ld = oracle.Index.load(tmp_path / "ix")
    assert ld.full and len(ld) == 100 and ld.ep == ix.ep and ld.params() == ix.params()
    assert np.array_equal(ld.export_values()[0], rows)
    for l in range(ix.nb_layers):
        for a, b in zip(ix.export_layer(l), ld.export_layer(l)):
            assert np.array_equal(a, b)
    for i in range(99):
        assert ix.distance(i, i + 1) == ld.distance(i, i + 1)
    # a QuantVec directory still loads as one
    q = oracle.Index(12, None, 10).insert_bulk(rows)
    q.save(tmp_path / "q")
    assert not oracle.Index.load(tmp_path / "q").full


# ---- committed expected outputs (tests/golden/expected_*.npz): the oracle rebuilt from source must reproduce them ----
@pytest.mark.parametrize("name,full", [("quant", False), ("full", True)])
def test_oracle_reproduces_committed_expected_outputs(oracle, glove, name, full):
    from golden_check import check_against_golden
    store, queries = glove
    ix = oracle.Index(12, None, store.shape[1], full=full).insert_bulk(store)
    check_against_golden(name, {"ep": ix.ep, "layers": ix.export_layers(),
                                "search": lambda q, n, ef: ix.search_batch(q, n, ef),
                                "bruteforce": lambda q, k: ix.bruteforce(q, k)}, queries)
