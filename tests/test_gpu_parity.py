"""GPU parity tests: every call goes through the C ABI (hnsw_rs_b200 -> libhnsw_b200.so) and is
compared bit for bit with the CPU oracle on the same seeded inputs.  Integer / index results
must be identical; f32 distances must be bit-identical (the north star allows 1e-5 relative,
the arithmetic contract gives equality)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SQRT2 = np.sqrt(np.float32(2.0))
DIMS = [1, 2, 7, 8, 9, 15, 16, 33, 50, 63, 96, 100, 128, 300]


@pytest.fixture(scope="module")
def H():
    import hnsw_rs_b200
    hnsw_rs_b200.Context.default()  # raises if there is no CUDA device: no CPU fallback
    return hnsw_rs_b200


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def synth(n, dim, ncent, seed, sigma=0.35, normalise=True):
    rc = np.random.default_rng(1234)
    cent = rc.standard_normal((ncent, dim), dtype=np.float32)
    r = np.random.default_rng(seed)
    x = cent[r.integers(0, ncent, n)] + np.float32(sigma) * r.standard_normal((n, dim), dtype=np.float32)
    if normalise:
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32)


def to_gpu(H, orc):
    """Import an oracle-built index into the device engine through flat arrays."""
    codes, mins, deltas, levels = orc.export_points()
    p = orc.params()
    prm = H.Params(p["ep"], p["m"], p["mmax"], p["mmax0"], p["ml"], p["ef_cons"], p["dim"])
    caps = [orc.layer_cap(l) for l in range(orc.nb_layers)]
    return H.HNSW.from_parts(prm, codes, mins, deltas, levels, orc.export_layers(), caps)


def to_oracle(oracle, ix):
    codes, mins, deltas, levels = ix._points().download()
    p = ix.params
    layers = [ix.export_layer(l) for l in range(ix.nb_layers())]
    return oracle.Index.from_parts(p.m, p.ef_cons, p.dim, p.ep, codes, mins, deltas, levels, layers)


def assert_same_graph(a_layers, b_layers):
    assert len(a_layers) == len(b_layers)
    for (ai, ao, an), (bi, bo, bn) in zip(a_layers, b_layers):
        assert np.array_equal(ai, bi)
        assert np.array_equal(ao, bo)
        assert np.array_equal(an, bn)


# ---- K1 quantiser ------------------------------------------------------------------------
@pytest.mark.parametrize("dim", DIMS)
def test_quantise_matches_oracle(H, oracle, dim):
    rng = np.random.default_rng(dim)
    rows = rng.normal(size=(257, dim)).astype(np.float32)
    rows[0] = 0.5                      # constant vector: delta = 0 -> NaN -> code 0
    rows[1, :] = np.float32(-0.0)      # signed zeros
    if dim > 1:
        rows[2, 0], rows[2, 1:] = np.float32(0.0), np.float32(-0.0)
        rows[3] = np.abs(rows[3]) * np.float32(1e-40)  # denormals
        rows[4] = rows[4] * np.float32(1e30)
    codes, mins, deltas = H.quantise_rows(rows)
    oc, om, od = oracle.quantise_rows(rows)
    assert np.array_equal(codes, oc)
    assert np.array_equal(bits(mins), bits(om))
    assert np.array_equal(bits(deltas), bits(od))


def test_quantise_nan_is_an_error(H):
    rows = np.zeros((4, 10), np.float32)
    rows[2, 3] = np.nan
    with pytest.raises(H.HnswB200Error):
        H.quantise_rows(rows)


# ---- K2 distances --------------------------------------------------------------------------
def test_quantvec_kats(H):  # vectors/src/quant.rs:143-202 through the device
    q = H.QuantVec.new
    assert q([0.5]).dist2other(q([0.25])) == np.float32(0.25)
    assert q([0.75]).dist2other(q([0.25])) == np.float32(0.5)
    assert q([0.0, 0.0]).dist2other(q([0.0, 1.0])) == np.float32(1.0)
    assert q([1.0, 0.0]).dist2other(q([0.0, 1.0])) == SQRT2
    assert q([-1.0, 0.0]).dist2other(q([0.0, 1.0])) == SQRT2
    assert q([1.0, 0.0]).dist2other(q([0.0, -1.0])) == SQRT2
    a = q(np.random.default_rng(0).random(128, dtype=np.float32))
    assert a.dist2other(a) == np.float32(0.0)


def test_fullvec_kats(H):  # vectors/src/full.rs:88-147
    f = H.FullVec.new
    assert f([0.5]).distance(f([0.25])) == np.float32(0.25)
    assert f([0.75]).distance(f([0.25])) == np.float32(0.5)
    assert f([0.0, 0.0]).distance(f([0.0, 1.0])) == np.float32(1.0)
    assert f([1.0, 0.0]).distance(f([0.0, 1.0])) == SQRT2
    assert f([-1.0, 0.0]).distance(f([0.0, 1.0])) == SQRT2
    assert f([1.0, 0.0]).distance(f([0.0, -1.0])) == SQRT2


def test_full_and_generic_distance_match_oracle(H, oracle):
    rng = np.random.default_rng(3)
    for dim in (1, 5, 128, 300):
        x = rng.normal(size=(64, dim)).astype(np.float32)
        y = rng.normal(size=(64, dim)).astype(np.float32)
        from hnsw_rs_b200.vectors import _dist_full_rows
        got = _dist_full_rows(x, y)
        want = np.array([oracle.dist_full(x[i], y[i]) for i in range(64)], np.float32)
        assert np.array_equal(bits(got), bits(want))


def test_dist_err_lt_one_percent(H):  # vectors/tests/full_lvq_tests.rs:3-27
    rng = np.random.default_rng(7)
    for _ in range(50):
        ra, rb = rng.random(128, dtype=np.float32), rng.random(128, dtype=np.float32)
        full = H.FullVec.new(ra).distance(H.FullVec.new(rb))
        qa, qb = H.QuantVec.new(ra), H.QuantVec.new(rb)
        assert abs(full - qa.distance(H.FullVec.new(rb))) / full < 0.01
        assert abs(full - qa.distance(qb)) / full < 0.01
        assert abs(full - qa.dist2other(qb)) / full < 0.01


@pytest.mark.parametrize("dim", DIMS)
def test_distances_match_oracle(H, oracle, dim):
    rng = np.random.default_rng(100 + dim)
    n = 300
    rows = (rng.normal(size=(n, dim)) * rng.uniform(0.1, 5.0, size=(n, 1))).astype(np.float32)
    rows[0] = 1.25  # constant (delta = 0) vector
    pts = H.SimplePoints.new(rows)
    codes, mins, deltas, _ = pts.download()
    oc, om, od = oracle.quantise_rows(rows)
    assert np.array_equal(codes, oc) and np.array_equal(bits(mins), bits(om)) and np.array_equal(bits(deltas), bits(od))
    a = rng.integers(0, n, 1000).astype(np.uint32)
    b = rng.integers(0, n, 1000).astype(np.uint32)
    got = pts.distances(a, b)
    want = np.array([oracle.dist_quant((oc[i], om[i], od[i]), (oc[j], om[j], od[j])) for i, j in zip(a, b)], np.float32)
    assert np.array_equal(bits(got), bits(want))
    assert np.array_equal(bits(got), bits(pts.distances(b, a)))  # exact symmetry
    # distance2point / dist2many: the f32 query is quantised first
    qv = rng.normal(size=dim).astype(np.float32)
    ids = rng.integers(0, n, 97).astype(np.uint32)
    got = pts.dist_query_many(qv, ids)
    qq = oracle.quantise(qv)
    want = np.array([oracle.dist_quant(qq, (oc[j], om[j], od[j])) for j in ids], np.float32)
    assert np.array_equal(bits(got), bits(want))
    assert pts.distance(0, n) is None and pts.distance2point(qv, n) is None


# ---- K3 search --------------------------------------------------------------------------------
def check_search(H, oracle, orc, queries, n, ef):
    ix = to_gpu(H, orc)
    ids, dists, counts, st = ix.ann_batch(queries, n, ef, with_stats=True)
    oids, odists, ocounts, ohops, oevals = orc.search_batch(queries, n, ef)
    assert np.array_equal(ids, oids)
    assert np.array_equal(bits(dists), bits(odists))
    assert np.array_equal(counts, ocounts)
    assert (st["flags"] == 0).all()
    assert np.array_equal(st["hops"], ohops)
    assert np.array_equal(st["evals"], oevals)
    return ix


@pytest.mark.parametrize("ef", [1, 5, 10, 37, 100, 300])
def test_search_glove_fixture(H, oracle, glove, glove_index, ef):
    _, queries = glove
    check_search(H, oracle, glove_index, queries, 10, ef)


def test_ann_by_vector_and_recall(H, oracle, glove, glove_index):  # template.rs:518-554
    store, queries = glove
    ix = to_gpu(H, glove_index)
    gt, gd = H.bruteforce_topk(ix._points(), queries, 10)
    ogt, ogd = glove_index.bruteforce(queries, 10)
    assert np.array_equal(gt, ogt) and np.array_equal(bits(gd), bits(ogd))
    hits = 0
    for i, q in enumerate(queries):
        ann = ix.ann_by_vector(q, 10, 100)
        assert ann == glove_index.ann_by_vector(q, 10, 100)
        hits += len(set(ann) & set(gt[i].tolist()))
    assert hits / (len(queries) * 10) > 0.99
    assert len(ix.ann_by_vector(queries[0], 10, 5)) == 5  # ef < n returns ef ids (results.rs:59-61)


@pytest.mark.parametrize("dim,n,m,efc", [(100, 6000, 16, 40), (128, 3000, 12, None), (96, 3000, 8, 32), (33, 2000, 5, None)])
def test_search_synthetic(H, oracle, dim, n, m, efc):
    base = synth(n, dim, 64, 1, normalise=(dim != 128))
    queries = synth(200, dim, 64, 2, normalise=(dim != 128))
    orc = oracle.Index(m, efc, dim).insert_bulk(base)
    for ef in (1, 10, 64, 150):
        check_search(H, oracle, orc, queries, 10, ef)
    check_search(H, oracle, orc, queries, 100, 100)
    check_search(H, oracle, orc, queries, 3, 50)


# {} = the default path (ef <= 128, quantised records of dimension 50/96/100/128: csrc/search_fast.cuh);
# NO_FAST = the round-1 register-list kernel (ef <= 64: 3584-entry visited table); VIS_POW2 = its 4096-entry table
PATHS = [{}, {"HNSWB200_NO_FAST": "1"}, {"HNSWB200_VIS_POW2": "1"}, {"HNSWB200_GENERAL_PATH": "1"}, {"HNSWB200_VIS32": "1"},
         {"HNSWB200_GENERAL_PATH": "1", "HNSWB200_VIS32": "1"}]


@pytest.mark.parametrize("env", PATHS)
def test_search_kernel_variants_agree(H, oracle, glove, glove_index, monkeypatch, env):
    """Variants of the search kernel (register-resident list with 2/4/8 keys per lane vs the shared-memory
    list of runtime width; 16-bit vs 32-bit visited entries) must give the same answers and counters as
    the oracle."""
    _, queries = glove
    ix = to_gpu(H, glove_index)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for ef in (1, 64, 100, 128, 129, 256, 257, 300):
        a = ix.ann_batch(queries, 10, ef, with_stats=True)
        o = glove_index.search_batch(queries, 10, ef)
        assert np.array_equal(a[0], o[0]) and np.array_equal(bits(a[1]), bits(o[1])), ef
        assert np.array_equal(a[3]["hops"], o[3]) and np.array_equal(a[3]["evals"], o[4]), ef
    for n in (1, 33, 100, 300):  # more results than ef
        a = ix.ann_batch(queries[:20], n, 50)
        o = glove_index.search_batch(queries[:20], n, 50)
        assert np.array_equal(a[0], o[0]) and np.array_equal(a[2], o[2])


@pytest.mark.parametrize("env", PATHS)
def test_search_visited_overflow_all_paths(H, oracle, glove, glove_index, monkeypatch, env):
    _, queries = glove
    ix = to_gpu(H, glove_index)
    monkeypatch.setenv("HNSWB200_VIS_SLOTS", "64")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    ids, dists, counts, st = ix.ann_batch(queries, 10, 100, with_stats=True)
    oids, odists, ocounts, _, oevals = glove_index.search_batch(queries, 10, 100)
    assert (st["flags"] & 2).any()
    assert np.array_equal(ids, oids) and np.array_equal(bits(dists), bits(odists))
    assert (st["evals"] >= oevals).all()


def test_build_with_32bit_visited(H, oracle, glove, glove_index, monkeypatch):
    store, _ = glove
    monkeypatch.setenv("HNSWB200_VIS32", "1")
    ix = H.HNSW.new(12, None, 50).insert_bulk(store, batch=1)
    assert_same_graph([ix.export_layer(l) for l in range(ix.nb_layers())], glove_index.export_layers())


def test_search_visited_overflow_keeps_results_exact(H, oracle, glove, glove_index, monkeypatch):
    _, queries = glove
    ix = to_gpu(H, glove_index)
    monkeypatch.setenv("HNSWB200_VIS_SLOTS", "64")
    ids, dists, counts, st = ix.ann_batch(queries, 10, 100, with_stats=True)
    monkeypatch.delenv("HNSWB200_VIS_SLOTS")
    oids, odists, ocounts, _, oevals = glove_index.search_batch(queries, 10, 100)
    assert (st["flags"] & 2).any()
    assert np.array_equal(ids, oids) and np.array_equal(bits(dists), bits(odists))
    assert (st["evals"] >= oevals).all()


def test_search_errors(H, oracle, glove, glove_index):
    _, queries = glove
    ix = to_gpu(H, glove_index)
    with pytest.raises(H.HnswB200Error):
        ix.ann_batch(np.zeros((2, 49), np.float32), 10, 10)  # dimension mismatch
    bad = queries[:3].copy()
    bad[1, 7] = np.nan
    with pytest.raises(H.HnswB200Error):
        ix.ann_batch(bad, 10, 10)  # NaN: the reference panics
    with pytest.raises(H.HnswB200Error):
        ix.ann_batch(queries, 10, 0)
    ids, dists, counts = ix.ann_batch(np.zeros((0, 50), np.float32), 10, 10)
    assert ids.shape == (0, 10)


# ---- K5 brute force / K6 merge -----------------------------------------------------------------
@pytest.mark.parametrize("dim,n,k", [(100, 20000, 100), (128, 5000, 10), (50, 1000, 1), (33, 3000, 7)])
def test_bruteforce_matches_oracle(H, oracle, dim, n, k):
    base = synth(n, dim, 32, 5)
    base[n // 2] = base[n // 3]  # exact duplicates: ties broken by id
    queries = synth(64, dim, 32, 6)
    queries[0] = base[n // 3]
    orc = oracle.Index(4, None, dim)
    codes, mins, deltas = oracle.quantise_rows(base)
    orc2 = oracle.Index.from_parts(4, 8, dim, 0, codes, mins, deltas, np.zeros(n, np.uint8),
                                   [(np.arange(n, dtype=np.uint32), np.zeros(n + 1, np.uint64), np.zeros(0, np.uint32))])
    pts = H.SimplePoints.new(base)
    ids, dists = H.bruteforce_topk(pts, queries, k)
    oids, odists = orc2.bruteforce(queries, k, threads=8)
    assert np.array_equal(ids, oids)
    assert np.array_equal(bits(dists), bits(odists))
    # base-sharded + merge == unsharded (configs 4 / 5): two shards with global ids
    half = n // 2
    s0 = H.SimplePoints.new(base[:half])
    s1 = H.SimplePoints.new(base[half:])
    i0, d0 = H.bruteforce_topk(s0, queries, k, 0)
    i1, d1 = H.bruteforce_topk(s1, queries, k, half)
    mi, md = H.topk_merge(np.stack([i0, i1]), np.stack([d0, d1]))
    assert np.array_equal(mi, oids) and np.array_equal(bits(md), bits(odists))


def test_bruteforce_adversarial_order(H, oracle):
    """Base sorted by decreasing distance to the query: every row beats the running threshold, the
    per-query buffer overflows and the chunk is redone in safe pieces."""
    dim, n = 16, 12000
    base = np.zeros((n, dim), np.float32)
    base[:, 0] = np.linspace(10.0, 1.0, n, dtype=np.float32)
    base[:, 1] = 1.0
    q = np.zeros((3, dim), np.float32)
    q[:, 1] = 1.0
    q[1, 0] = 0.5
    codes, mins, deltas = oracle.quantise_rows(base)
    orc = oracle.Index.from_parts(4, 8, dim, 0, codes, mins, deltas, np.zeros(n, np.uint8),
                                  [(np.arange(n, dtype=np.uint32), np.zeros(n + 1, np.uint64), np.zeros(0, np.uint32))])
    ids, dists = H.bruteforce_topk(H.SimplePoints.new(base), q, 10)
    oids, odists = orc.bruteforce(q, 10)
    assert np.array_equal(ids, oids) and np.array_equal(bits(dists), bits(odists))


def _flat_oracle(oracle, base, dim):
    n = len(base)
    codes, mins, deltas = oracle.quantise_rows(base)
    return oracle.Index.from_parts(4, 8, dim, 0, codes, mins, deltas, np.zeros(n, np.uint8),
                                   [(np.arange(n, dtype=np.uint32), np.zeros(n + 1, np.uint64), np.zeros(0, np.uint32))])


@pytest.mark.parametrize("dim", [100, 96, 128])
def test_bruteforce_tensor_core_adversarial_order(H, oracle, dim, monkeypatch):
    """The tcgen05 filter path (records of 128 bytes) with the base sorted by decreasing distance: candidate lists
    overflow, the chunk is redone by the exact kernel; and the same data through the CUDA-core path."""
    n = 30000
    r = np.random.default_rng(3)
    base = r.standard_normal((n, dim)).astype(np.float32) * 0.05
    base[:, 0] += np.linspace(10.0, 1.0, n, dtype=np.float32)
    q = r.standard_normal((5, dim)).astype(np.float32) * 0.05
    orc = _flat_oracle(oracle, base, dim)
    oids, odists = orc.bruteforce(q, 10, threads=8)
    pts = H.SimplePoints.new(base)
    ids, dists = H.bruteforce_topk(pts, q, 10)
    assert np.array_equal(ids, oids) and np.array_equal(bits(dists), bits(odists))
    monkeypatch.setenv("HNSWB200_BF_NO_TC", "1")
    ids, dists = H.bruteforce_topk(pts, q, 10)
    assert np.array_equal(ids, oids) and np.array_equal(bits(dists), bits(odists))


def test_bruteforce_tensor_core_extreme_records(H, oracle):
    """Records the algebraic filter cannot rank (constant vectors: delta = 0; huge offsets; huge ranges; exact
    duplicates) must still come out exactly: the filter only ever lets too many through, never too few."""
    dim, n = 100, 20000
    r = np.random.default_rng(11)
    base = r.standard_normal((n, dim)).astype(np.float32)
    base[100] = 0.25                      # constant vector: delta = 0, every code 0 (quant.rs:154-166)
    base[101] = base[7]                   # exact duplicate: tie broken by id
    base[200:300] += np.float32(1.0e4)    # large offset: cancellation in Sum x^2 + Sum y^2 - 2 Sum xy
    base[300:400] *= np.float32(1.0e3)    # large range
    base[400:420] *= np.float32(1.0e-6)   # tiny range around zero
    q = np.concatenate([base[[7, 100, 250, 350, 410]], r.standard_normal((27, dim)).astype(np.float32)])
    q[6] += np.float32(1.0e4)
    orc = _flat_oracle(oracle, base, dim)
    for k in (1, 10, 100):
        oids, odists = orc.bruteforce(q, k, threads=8)
        ids, dists = H.bruteforce_topk(H.SimplePoints.new(base), q, k)
        assert np.array_equal(ids, oids), k
        assert np.array_equal(bits(dists), bits(odists)), k


def test_bruteforce_tensor_core_large_matches_cuda_core_path(H, monkeypatch):
    """Size-independent property at a scale the oracle does not reach in seconds: both device paths agree bit for bit
    (300,000 x 100 base, 2,000 queries, top-100), results are sorted by (dist, id) and ids are unique per query."""
    base = synth(300000, 100, 512, 21)
    q = synth(2000, 100, 512, 22)
    pts = H.SimplePoints.new(base)
    ids, dists = H.bruteforce_topk(pts, q, 100)
    monkeypatch.setenv("HNSWB200_BF_NO_TC", "1")
    ids2, dists2 = H.bruteforce_topk(pts, q, 100)
    assert np.array_equal(ids, ids2) and np.array_equal(bits(dists), bits(dists2))
    key = (bits(dists).astype(np.uint64) << np.uint64(32)) | ids.astype(np.uint64)
    assert (np.diff(key.astype(np.int64), axis=1) > 0).all()


# ---- save / load (hnsw/src/template.rs:43-131, 574-611) ---------------------------------------------
def test_save_load_cross_with_oracle(H, oracle, glove, glove_index, tmp_path):
    _, queries = glove
    ix = to_gpu(H, glove_index)
    ix.save(tmp_path / "gpu")
    back = oracle.Index.load(tmp_path / "gpu")          # the oracle reads what the engine wrote
    assert back.params() == glove_index.params()
    for a, b in zip(back.export_points(), glove_index.export_points()):
        assert np.array_equal(a, b)
    assert_same_graph(back.export_layers(), glove_index.export_layers())
    glove_index.save(tmp_path / "cpu")
    for name in ("points", "params"):
        assert (tmp_path / "gpu" / name).read_bytes() == (tmp_path / "cpu" / name).read_bytes()
    ld = H.HNSW.load(tmp_path / "cpu")                   # the engine reads what the oracle wrote
    assert ld.len() == 1000 and ld.params.ep == glove_index.ep
    assert_same_graph([ld.export_layer(l) for l in range(ld.nb_layers())], glove_index.export_layers())
    a = ld.ann_batch(queries, 10, 100)
    b = glove_index.search_batch(queries, 10, 100)
    assert np.array_equal(a[0], b[0]) and np.array_equal(bits(a[1]), bits(b[1]))
    with pytest.raises(H.HnswB200Error):
        H.HNSW.load(tmp_path / "missing")


# ---- build (hnsw/src/template.rs:177-293, 388-444) ------------------------------------------------
def test_build_batch1_reproduces_oracle_graph(H, oracle, glove, glove_index):
    store, queries = glove
    ix = H.HNSW.new(12, None, 50).insert_bulk(store, batch=1)
    assert ix.params.ep == glove_index.ep
    codes, mins, deltas, levels = ix._points().download()
    oc, om, od, ol = glove_index.export_points()
    assert np.array_equal(levels, ol) and np.array_equal(codes, oc)
    assert_same_graph([ix.export_layer(l) for l in range(ix.nb_layers())], glove_index.export_layers())


@pytest.mark.parametrize("dim,n,m,efc", [(100, 1500, 16, 60), (128, 800, 6, None), (33, 700, 4, 9)])
def test_build_batch1_synthetic(H, oracle, dim, n, m, efc):
    base = synth(n, dim, 16, 11)
    orc = oracle.Index(m, efc, dim).insert_bulk(base)
    ix = H.HNSW.new(m, efc, dim).insert_bulk(base, batch=1)
    assert ix.params.ep == orc.ep
    assert_same_graph([ix.export_layer(l) for l in range(ix.nb_layers())], orc.export_layers())


def test_build_batched_quality_and_search_parity(H, oracle, glove):
    store, queries = glove
    ix = H.HNSW.new(12, None, 50).insert_bulk(store)  # default batching
    assert ix.len() == 1000
    assert ix.assert_param_compliance()               # template.rs:341-370
    for l in range(ix.nb_layers()):                   # template.rs:556-571: min degree > 0
        ids, off, _ = ix.export_layer(l)
        if len(ids) > 1:
            assert np.diff(off.astype(np.int64)).min() > 0
    gt, _ = H.bruteforce_topk(ix._points(), queries, 10)
    ids, dists, counts, st = ix.ann_batch(queries, 10, 100, with_stats=True)
    hits = sum(len(set(gt[i].tolist()) & set(ids[i].tolist())) for i in range(len(queries)))
    assert hits / 1000 > 0.99
    orc = to_oracle(oracle, ix)                       # oracle search over the device-built graph
    oids, odists, ocounts, ohops, oevals = orc.search_batch(queries, 10, 100)
    assert np.array_equal(ids, oids) and np.array_equal(bits(dists), bits(odists))
    assert np.array_equal(st["hops"], ohops) and np.array_equal(st["evals"], oevals)
    # graph invariants of graph.rs:305-340: symmetric, no self loops
    for l in range(ix.nb_layers()):
        g = ix.get_layer(l)
        for node in g.iter_nodes():
            for nb in g.neighbors(node):
                assert nb != node and node in g.neighbors(nb)


def test_build_larger_batched(H, oracle):
    base = synth(30000, 100, 256, 21)
    queries = synth(500, 100, 256, 22)
    ix = H.HNSW.new(16, 100, 100).insert_bulk(base)
    gt, _ = H.bruteforce_topk(ix._points(), queries, 10)
    ids, dists, counts, st = ix.ann_batch(queries, 10, 64, with_stats=True)
    hits = sum(len(set(gt[i].tolist()) & set(ids[i].tolist())) for i in range(len(queries)))
    assert hits / (10 * len(queries)) > 0.99
    orc = to_oracle(oracle, ix)
    oids, odists, _, ohops, oevals = orc.search_batch(queries, 10, 64, threads=8)
    assert np.array_equal(ids, oids) and np.array_equal(bits(dists), bits(odists))
    assert np.array_equal(st["hops"], ohops) and np.array_equal(st["evals"], oevals)
    assert ix.assert_param_compliance()


def test_insert_after_build(H, oracle):  # template.rs:479-504
    rng = np.random.default_rng(5)
    a = rng.random((100, 10), dtype=np.float32)
    b = rng.random((100, 10), dtype=np.float32)
    v = rng.random(10, dtype=np.float32)
    orc = oracle.Index(12, None, 10).insert_bulk(a)
    ix = H.HNSW.new(12, None, 10).insert_bulk(a, batch=1)
    assert orc.insert_vec(v) == ix.insert_vec(v) == 100
    orc.insert_bulk(b)
    ix.insert_bulk(b, batch=1)
    assert ix.len() == 201 and ix.params.ep == orc.ep
    assert_same_graph([ix.export_layer(l) for l in range(ix.nb_layers())], orc.export_layers())
    with pytest.raises(H.HnswB200Error):  # can_not_add_different_dim (reference panics)
        ix.insert_bulk(rng.random((10, 12), dtype=np.float32))


def test_extend_loaded_index(H, oracle, tmp_path):
    """Edge lengths of an imported graph are recomputed on the device before it is extended."""
    rng = np.random.default_rng(9)
    a = rng.random((300, 24), dtype=np.float32)
    b = rng.random((50, 24), dtype=np.float32)
    orc = oracle.Index(8, None, 24).insert_bulk(a)
    orc.save(tmp_path / "ix")
    ix = H.HNSW.load(tmp_path / "ix")
    ix.insert_bulk(b, batch=1)
    orc.insert_bulk(b)
    assert_same_graph([ix.export_layer(l) for l in range(ix.nb_layers())], orc.export_layers())


# ---- BASELINE configs[1] at full size: size-independent properties ---------------------------------------
def test_full_size_c2_properties(H, oracle):
    """1,183,514 x 100 (the bench workload), built on the device: results are sorted by (dist, id), ids are unique,
    every returned distance equals the batched distance kernel's value bit for bit, results do not depend on how the
    queries are batched or on n, recall@10 >= 0.99 at ef = 64 against the exact ground truth, and a sample agrees with
    the oracle searching the exported graph (ids, distances, hop and evaluation counters)."""
    base = synth(1183514, 100, 2048, 1)
    queries = synth(2000, 100, 2048, 2)
    ix = H.HNSW.new(16, 200, 100).insert_bulk(base)
    assert ix.len() == 1183514
    ids, dists, counts, st = ix.ann_batch(queries, 10, 64, with_stats=True)
    assert (counts == 10).all()
    key = (bits(dists).astype(np.uint64) << np.uint64(32)) | ids.astype(np.uint64)
    assert (key[:, 1:] > key[:, :-1]).all()                      # strictly ascending (dist, id): sorted and unique
    for q in (0, 1, 777, 1999):                                  # distances are the kernel's own exact values
        d = ix._points().dist_query_many(queries[q], ids[q])
        assert np.array_equal(bits(d), bits(dists[q]))
    ids2, dists2, _ = ix.ann_batch(queries, 10, 64)              # idempotent
    assert np.array_equal(ids, ids2) and np.array_equal(bits(dists), bits(dists2))
    sub = np.arange(0, 2000, 7)                                  # independent of the batch composition
    ids3, _, _ = ix.ann_batch(queries[sub], 10, 64)
    assert np.array_equal(ids3, ids[sub])
    ids4, _, c4 = ix.ann_batch(queries[:300], 40, 64)            # n only truncates the list
    assert (c4 == 40).all() and np.array_equal(ids4[:, :10], ids[:300])
    gt, gd = H.bruteforce_topk(ix._points(), queries, 10)
    hits = sum(len(set(gt[i].tolist()) & set(ids[i].tolist())) for i in range(len(queries)))
    assert hits / gt.size >= 0.99
    gkey = (bits(gd).astype(np.uint64) << np.uint64(32)) | gt.astype(np.uint64)
    assert (gkey[:, 1:] > gkey[:, :-1]).all() and (gkey[:, 0] <= key[:, 0]).all()   # nothing beats the exact nearest
    orc = to_oracle(oracle, ix)
    o = orc.search_batch(queries[:200], 10, 64, threads=8)
    ok = st["flags"][:200] == 0
    assert np.array_equal(ids[:200], o[0]) and np.array_equal(bits(dists[:200]), bits(o[1]))
    assert np.array_equal(st["hops"][:200], o[3]) and np.array_equal(st["evals"][:200][ok], o[4][ok])


def test_search_pinned_host_buffers_are_used_in_place(H, oracle, glove, glove_index, monkeypatch):
    """hnswb200_search reads page-locked query buffers and writes page-locked result buffers in place (zero-copy over
    PCIe); pageable buffers are staged.  Same answers either way, and a NaN query is still reported."""
    import ctypes as C
    import torch
    from hnsw_rs_b200 import _ffi
    _, queries = glove
    ix = to_gpu(H, glove_index)
    nq, dim = queries.shape
    ref = ix.ann_batch(queries, 10, 60, with_stats=True)  # pageable numpy buffers: staged
    hq = torch.from_numpy(queries.copy()).pin_memory()
    hid = torch.zeros((nq, 10), dtype=torch.int32).pin_memory()
    hd = torch.zeros((nq, 10), dtype=torch.float32).pin_memory()
    hc = torch.zeros(nq, dtype=torch.int32).pin_memory()
    hh = torch.zeros(nq, dtype=torch.int32).pin_memory()
    he = torch.zeros(nq, dtype=torch.int32)  # one pageable statistic among pinned buffers
    st = _ffi.SearchStats(C.cast(hh.data_ptr(), _ffi.u32p), C.cast(he.data_ptr(), _ffi.u32p), None, None)
    lib = _ffi.lib()

    def call():
        _ffi.check(lib.hnswb200_search(ix.ctx.h, ix.h, C.cast(hq.data_ptr(), _ffi.f32p), nq, dim, 10, 60,
                                       C.cast(hid.data_ptr(), _ffi.u32p), C.cast(hd.data_ptr(), _ffi.f32p),
                                       C.cast(hc.data_ptr(), _ffi.u32p), C.byref(st)))
    call()
    assert np.array_equal(hid.numpy().view(np.uint32), ref[0]) and np.array_equal(bits(hd.numpy()), bits(ref[1]))
    assert np.array_equal(hc.numpy().view(np.uint32), ref[2])
    assert np.array_equal(hh.numpy().view(np.uint32), ref[3]["hops"]) and np.array_equal(he.numpy().view(np.uint32), ref[3]["evals"])
    hid.zero_()
    monkeypatch.setenv("HNSWB200_NO_ZERO_COPY", "1")
    call()
    monkeypatch.delenv("HNSWB200_NO_ZERO_COPY")
    assert np.array_equal(hid.numpy().view(np.uint32), ref[0])
    hq[3, 5] = float("nan")
    with pytest.raises(H.HnswB200Error):
        call()
    hq[3, 5] = 0.0
    call()  # the context recovers


# ---- cosine (an addition: the reference has L2 only) ---------------------------------------------
def test_cosine_index_is_l2_over_unit_rows(H, oracle):
    """metric="cosine": rows and queries are L2-normalised on the device before they are quantised, then the reference's
    L2 path runs unchanged.  (1) the normalisation matches numpy within 1e-6 relative; (2) a cosine index over raw rows
    returns exactly what an L2 index returns over the device-normalised rows and queries, and that equals the oracle;
    (3) the ranking is the cosine ranking: recall@10 against exact float cosine top-10 is high."""
    r = np.random.default_rng(5)
    base = (synth(5000, 100, 64, 31, normalise=False) * r.uniform(0.2, 5.0, (5000, 1))).astype(np.float32)
    queries = (synth(200, 100, 64, 32, normalise=False) * r.uniform(0.2, 5.0, (200, 1))).astype(np.float32)
    nb, nq = H.normalise_rows(base), H.normalise_rows(queries)
    ref = base / np.linalg.norm(base.astype(np.float64), axis=1, keepdims=True)
    assert np.allclose(nb, ref, rtol=1e-6, atol=1e-7)
    assert np.array_equal(H.normalise_rows(np.zeros((2, 7), np.float32)), np.zeros((2, 7), np.float32))
    cx = H.HNSW.new(16, 60, 100, metric="cosine").insert_bulk(base, batch=1)
    lx = H.HNSW.new(16, 60, 100).insert_bulk(nb, batch=1)
    assert cx.metric == "cosine" and lx.metric == "l2"
    a = cx.ann_batch(queries, 10, 80, with_stats=True)
    b = lx.ann_batch(nq, 10, 80, with_stats=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(bits(a[1]), bits(b[1])) and np.array_equal(a[3]["evals"], b[3]["evals"])
    orc = oracle.Index(16, 60, 100).insert_bulk(nb)
    o = orc.search_batch(nq, 10, 80)
    assert np.array_equal(a[0], o[0]) and np.array_equal(bits(a[1]), bits(o[1]))
    gt, _ = H.bruteforce_topk(cx._points(), queries, 10)          # normalises the queries itself
    gt2, _ = H.bruteforce_topk(lx._points(), nq, 10)
    assert np.array_equal(gt, gt2)
    cos = (queries.astype(np.float64) / np.linalg.norm(queries, axis=1, keepdims=True)) @ ref.T
    true10 = np.argsort(-cos, axis=1)[:, :10]
    hits = sum(len(set(true10[i].tolist()) & set(a[0][i].tolist())) for i in range(len(queries)))
    assert hits / true10.size > 0.9
    d = cx._points().dist_query_many(queries[0], a[0][0])          # distances are L2 on the unit vectors: d^2 = 2 - 2 cos
    assert np.allclose(d.astype(np.float64) ** 2, 2 - 2 * cos[0, a[0][0]], atol=2e-2)
    with pytest.raises(H.HnswB200Error):
        cx.set_metric("l2")                                          # rows are already quantised as unit vectors


@pytest.mark.parametrize("dim,n,m,efc", [(100, 4000, 24, 60), (96, 3000, 40, 90), (50, 2500, 33, None)])
def test_search_wide_rows(H, oracle, dim, n, m, efc):
    """M > 16: adjacency rows wider than one 32-slot batch (layer 0: 2M slots, upper layers: M slots), on both list
    implementations and through the device build."""
    base = synth(n, dim, 48, 41)
    queries = synth(150, dim, 48, 42)
    orc = oracle.Index(m, efc, dim).insert_bulk(base)
    for ef in (1, 30, 64, 100, 200, 300):
        check_search(H, oracle, orc, queries, 10, ef)
    ix = H.HNSW.new(m, efc, dim).insert_bulk(base, batch=1)
    assert_same_graph([ix.export_layer(l) for l in range(ix.nb_layers())], orc.export_layers())


def test_search_randomised_against_oracle(H, oracle):
    """A small randomised sweep over dimension, M, ef_cons, ef and n (fixed seed): ids, distances, counts and counters
    equal the oracle's on the oracle's graph, for every combination."""
    r = np.random.default_rng(2024)
    for trial in range(12):
        dim = int(r.choice([3, 8, 17, 50, 64, 96, 100, 128, 130]))
        m = int(r.integers(3, 21))
        n = int(r.integers(300, 2500))
        efc = None if r.random() < 0.3 else int(r.integers(m, 3 * m + 8))
        base = synth(n, dim, 16, 100 + trial, normalise=bool(r.integers(0, 2)))
        queries = synth(64, dim, 16, 200 + trial, normalise=False)
        orc = oracle.Index(m, efc, dim).insert_bulk(base)
        for ef in sorted(set(int(x) for x in r.integers(1, 280, 3))):
            nres = int(r.choice([1, 10, 33]))
            check_search(H, oracle, orc, queries, nres, ef)


def test_search_async_host_buffers(H, oracle, glove, glove_index):
    """hnswb200_search_async: page-locked buffers, several batches in flight, one hnswb200_ctx_sync; pageable buffers are
    refused; a NaN query surfaces at the sync."""
    import ctypes as C
    import torch
    from hnsw_rs_b200 import _ffi
    _, queries = glove
    ix = to_gpu(H, glove_index)
    nq, dim = queries.shape
    ref = ix.ann_batch(queries, 10, 60)
    lib = _ffi.lib()
    hq = torch.from_numpy(queries.copy()).pin_memory()
    outs = [(torch.zeros((nq, 10), dtype=torch.int32).pin_memory(), torch.zeros((nq, 10), dtype=torch.float32).pin_memory(),
             torch.zeros(nq, dtype=torch.int32).pin_memory()) for _ in range(3)]
    for i, d, c in outs:
        _ffi.check(lib.hnswb200_search_async(ix.ctx.h, ix.h, C.cast(hq.data_ptr(), _ffi.f32p), nq, dim, 10, 60,
                                             C.cast(i.data_ptr(), _ffi.u32p), C.cast(d.data_ptr(), _ffi.f32p),
                                             C.cast(c.data_ptr(), _ffi.u32p)))
    ix.ctx.sync()
    for i, d, c in outs:
        assert np.array_equal(i.numpy().view(np.uint32), ref[0]) and np.array_equal(bits(d.numpy()), bits(ref[1]))
        assert np.array_equal(c.numpy().view(np.uint32), ref[2])
    pageable = np.zeros((nq, 10), np.uint32)
    assert lib.hnswb200_search_async(ix.ctx.h, ix.h, C.cast(hq.data_ptr(), _ffi.f32p), nq, dim, 10, 60,
                                     pageable.ctypes.data_as(_ffi.u32p), None, None) == -1
    hq[1, 0] = float("nan")
    _ffi.check(lib.hnswb200_search_async(ix.ctx.h, ix.h, C.cast(hq.data_ptr(), _ffi.f32p), nq, dim, 10, 60,
                                         C.cast(outs[0][0].data_ptr(), _ffi.u32p), None, None))
    with pytest.raises(H.HnswB200Error):
        ix.ctx.sync()
    ix.ctx.sync()  # the status is cleared by the failing sync


def test_two_contexts_search_one_index_concurrently(H, oracle, glove, glove_index):
    """One context per host thread: two threads with their own contexts (streams, workspaces) search the
    same device-resident index at the same time; both get the oracle's answers."""
    import ctypes as C
    import threading
    from hnsw_rs_b200 import _ffi
    _, queries = glove
    ix = to_gpu(H, glove_index)
    ref = glove_index.search_batch(queries, 10, 50)
    lib = _ffi.lib()
    results, errors = {}, []

    def worker(name):
        try:
            ctx = H.Context(0)
            q = np.ascontiguousarray(queries, np.float32)
            for rep in range(20):
                ids = np.zeros((len(q), 10), np.uint32)
                d = np.zeros((len(q), 10), np.float32)
                c = np.zeros(len(q), np.uint32)
                _ffi.check(lib.hnswb200_search(ctx.h, ix.h, q.ctypes.data_as(_ffi.f32p), len(q), q.shape[1], 10, 50,
                                               ids.ctypes.data_as(_ffi.u32p), d.ctypes.data_as(_ffi.f32p),
                                               c.ctypes.data_as(_ffi.u32p), None))
                if not (np.array_equal(ids, ref[0]) and np.array_equal(bits(d), bits(ref[1])) and np.array_equal(c, ref[2])):
                    errors.append((name, rep))
            ctx.close()
        except Exception as e:  # noqa: BLE001
            errors.append((name, repr(e)))

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


def test_search_small_visited_table_under_heavy_load(H, oracle, monkeypatch):
    """Wide rows (M = 32) at ef = 64 fill the visited table far beyond its design load.  Default kernel (search_fast.cuh):
    full buckets spill to the following ones and, after 8 of them, to the exact spill list, so ids, distances, hops AND
    the evaluation counter equal the oracle's for EVERY query (searcher.rs:61-72 evaluates an id once), also with the
    smallest table the kernel accepts.  Round-1 kernel (NO_FAST): answers identical, the counter may over-count where the
    overflow flag is raised."""
    from hnsw_rs_b200 import _ffi
    base = synth(20000, 100, 64, 31)
    queries = synth(300, 100, 64, 32)
    orc = oracle.Index(32, 64, 100).insert_bulk(base)
    ix = to_gpu(H, orc)
    oids, odists, ocounts, ohops, oevals = orc.search_batch(queries, 10, 64, threads=8)
    spilled = 0
    for nb in (None, "514", "300"):  # 300 <= 512: the entry keeps 13 bits of the hash and 2 of displacement
        if nb:
            monkeypatch.setenv("HNSWB200_FAST_NB", nb)
        ids, dists, counts, st = ix.ann_batch(queries, 10, 64, with_stats=True)
        assert "search_kernel_fast" in _ffi.last_search_variant()
        assert np.array_equal(ids, oids) and np.array_equal(bits(dists), bits(odists)) and np.array_equal(counts, ocounts)
        assert np.array_equal(st["hops"], ohops)
        assert not (st["flags"] & 2).any()
        assert np.array_equal(st["evals"], oevals)            # including the queries that used the spill list
        spilled += int(((st["flags"] & 4) != 0).sum())
    monkeypatch.delenv("HNSWB200_FAST_NB")
    monkeypatch.setenv("HNSWB200_NO_FAST", "1")
    ids1, dists1, counts1, st1 = ix.ann_batch(queries, 10, 64, with_stats=True)
    assert "search_kernel_reg" in _ffi.last_search_variant()
    assert np.array_equal(ids1, oids) and np.array_equal(bits(dists1), bits(odists)) and np.array_equal(st1["hops"], ohops)
    clean = (st1["flags"] & 2) == 0
    assert np.array_equal(st1["evals"][clean], oevals[clean]) and (st1["evals"] >= oevals).all()
    print("queries that used the spill list:", spilled)


@pytest.mark.parametrize("dim,m", [(100, 16), (128, 16), (96, 12), (50, 12)])
def test_fast_kernel_matches_oracle_and_round1_kernel(H, oracle, monkeypatch, dim, m):
    """The second-generation search kernel (csrc/search_fast.cuh; ef <= 128, dimensions 50 / 96 / 100 / 128) against the
    oracle and against the round-1 kernel on the same index: ids, distance bits, counts, hops and evaluations."""
    from hnsw_rs_b200 import _ffi
    norm = dim != 128
    base = synth(8000, dim, 64, 51, normalise=norm)
    queries = synth(257, dim, 64, 52, normalise=norm)
    orc = oracle.Index(m, 3 * m, dim).insert_bulk(base)
    ix = to_gpu(H, orc)
    for ef, n in ((1, 1), (7, 10), (57, 10), (64, 64), (65, 10), (100, 100), (128, 130)):
        o = orc.search_batch(queries, n, ef, threads=8)
        a = ix.ann_batch(queries, n, ef, with_stats=True)
        assert "search_kernel_fast" in _ffi.last_search_variant(), _ffi.last_search_variant()
        monkeypatch.setenv("HNSWB200_NO_FAST", "1")
        b = ix.ann_batch(queries, n, ef, with_stats=True)
        assert "search_kernel_fast" not in _ffi.last_search_variant()
        monkeypatch.delenv("HNSWB200_NO_FAST")
        for r in (a, b):
            assert np.array_equal(r[0], o[0]) and np.array_equal(bits(r[1]), bits(o[1])) and np.array_equal(r[2], o[2]), ef
            assert np.array_equal(r[3]["hops"], o[3]) and np.array_equal(r[3]["evals"], o[4]), ef
        assert (a[3]["flags"] == 0).all()
        c = ix.ann_batch(queries, n, ef)                      # the variant without counters (the one bench.py times)
        assert np.array_equal(c[0], o[0]) and np.array_equal(bits(c[1]), bits(o[1])) and np.array_equal(c[2], o[2]), ef


def test_fast_kernel_extreme_deltas(H, oracle):
    """Records whose delta is >= 2^100 (or whose range overflows f32) take the separately-rounded multiply path of the
    fast kernel's distance (FastQuery::partial): still bit-identical to the oracle."""
    r = np.random.default_rng(77)
    base = synth(3000, 100, 16, 61, normalise=False)
    base[::7] *= np.float32(3.0e32)      # delta ~ 1e30..1e31 >= 2^100
    base[5::11] *= np.float32(1.0e-30)   # tiny deltas
    queries = synth(100, 100, 16, 62, normalise=False)
    queries[::3] *= np.float32(3.0e32)
    orc = oracle.Index(12, 36, 100).insert_bulk(base)
    for ef in (10, 64, 100):
        check_search(H, oracle, orc, queries, 10, ef)
