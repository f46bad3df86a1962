"""GPU parity tests of the f32 (FullVec) index mode: the index the reference builds when its vector type alias is
flipped (`type VecType = FullVec;`, points/src/point.rs:4).  Points keep their f32 values (vectors/src/full.rs:3-6) and
every distance is FullVec::distance (full.rs:23-29): one strictly sequential f32 sum of (x - y)^2, then sqrt.
Everything goes through the C ABI and is compared bit for bit with the CPU oracle in the same mode."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DIMS = [1, 2, 7, 15, 16, 17, 33, 50, 96, 100, 128, 300]


@pytest.fixture(scope="module")
def H():
    import hnsw_rs_b200
    hnsw_rs_b200.Context.default()  # raises if there is no CUDA device: no CPU fallback
    return hnsw_rs_b200


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def synth(n, dim, ncent, seed, sigma=0.35):
    rc = np.random.default_rng(4321)
    cent = rc.standard_normal((ncent, dim), dtype=np.float32)
    r = np.random.default_rng(seed)
    x = cent[r.integers(0, ncent, n)] + np.float32(sigma) * r.standard_normal((n, dim), dtype=np.float32)
    return x.astype(np.float32)


def to_gpu(H, orc):
    vals, levels = orc.export_values()
    p = orc.params()
    prm = H.Params(p["ep"], p["m"], p["mmax"], p["mmax0"], p["ml"], p["ef_cons"], p["dim"])
    caps = [orc.layer_cap(l) for l in range(orc.nb_layers)]
    ix = H.HNSW.from_parts(prm, vals, None, None, levels, orc.export_layers(), caps)
    assert ix.vec_type == "full"
    return ix


def to_oracle(oracle, ix):
    vals, levels = ix._points().values()
    p = ix.params
    layers = [ix.export_layer(l) for l in range(ix.nb_layers())]
    return oracle.Index.from_parts(p.m, p.ef_cons, p.dim, p.ep, vals, None, None, levels, layers)


def layers_of(ix):
    return [ix.export_layer(l) for l in range(ix.nb_layers())]


def assert_same_graph(a_layers, b_layers):
    assert len(a_layers) == len(b_layers)
    for a, b in zip(a_layers, b_layers):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def seq_dist(x, y):
    """FullVec::distance written out in numpy f32, one rounding per operation"""
    s = np.float32(0)
    for a, b in zip(x, y):
        t = np.float32(a) - np.float32(b)
        s = np.float32(s + np.float32(t * t))
    return np.sqrt(s)


# ---- points and distances ----------------------------------------------------------------
@pytest.mark.parametrize("dim", DIMS)
def test_full_points_and_distances(H, oracle, dim):
    rng = np.random.default_rng(dim)
    rows = (rng.standard_normal((257, dim)) * 3).astype(np.float32)
    pts = H.SimplePoints.new(rows, vec_type="full")
    assert pts.vec_type == "full" and pts.len() == 257 and pts.dim() == dim
    back, _ = pts.values()
    assert np.array_equal(bits(back), bits(rows))
    a = rng.integers(0, 257, 500)
    b = rng.integers(0, 257, 500)
    d = pts.distances(a, b)                                    # Points::distance (points.rs:86-93)
    want = np.array([oracle.dist_full(rows[i], rows[j]) for i, j in zip(a, b)], np.float32)
    assert np.array_equal(bits(d), bits(want))
    for i in range(5):
        assert bits(d[i:i + 1])[0] == bits(np.array([seq_dist(rows[a[i]], rows[b[i]])]))[0]
    q = (rng.standard_normal(dim) * 3).astype(np.float32)     # distance2point / dist2many: the query is NOT quantised
    ids = rng.integers(0, 257, 300)
    dq = pts.dist_query_many(q, ids)
    want = np.array([oracle.dist_full(q, rows[i]) for i in ids], np.float32)
    assert np.array_equal(bits(dq), bits(want))
    p = pts.get_point(3)
    assert np.array_equal(p.get_vals(), rows[3]) and p.vector.size() == 4 * dim
    with pytest.raises(H.HnswB200Error):                       # codes / min / delta do not exist for FullVec points
        pts.download()


def test_full_points_refuse_non_finite(H):
    rows = np.ones((4, 9), np.float32)
    rows[2, 5] = np.nan
    with pytest.raises(H.HnswB200Error):
        H.SimplePoints.new(rows, vec_type="full")
    rows[2, 5] = np.inf
    with pytest.raises(H.HnswB200Error):
        H.SimplePoints.new(rows, vec_type="full")


def test_values_of_quantised_points_are_the_dequantised_values(H, oracle):
    rng = np.random.default_rng(3)
    rows = rng.standard_normal((64, 37)).astype(np.float32)
    pts = H.SimplePoints.new(rows)
    assert pts.vec_type == "quant"
    vals, _ = pts.values()
    codes, mins, deltas, _ = pts.download()
    for i in range(64):
        assert np.array_equal(bits(vals[i]), bits(oracle.dequantise(codes[i], mins[i], deltas[i])))


# ---- search ------------------------------------------------------------------------------
def check_search(H, oracle, orc, queries, n, ef, ix=None):
    ix = ix or to_gpu(H, orc)
    ids, dists, counts, st = ix.ann_batch(queries, n, ef, with_stats=True)
    oids, odists, ocounts, ohops, oevals = orc.search_batch(queries, n, ef)
    assert np.array_equal(ids, oids)
    assert np.array_equal(bits(dists), bits(odists))
    assert np.array_equal(counts, ocounts)
    assert np.array_equal(st["hops"], ohops)
    assert np.array_equal(st["evals"], oevals)
    return ix


@pytest.fixture(scope="module")
def glove_full(oracle, glove):
    store, _ = glove
    return oracle.Index(12, None, store.shape[1], full=True).insert_bulk(store)


@pytest.mark.parametrize("ef", [1, 10, 37, 100, 300])
def test_full_search_glove_fixture(H, oracle, glove, glove_full, ef):
    _, queries = glove
    check_search(H, oracle, glove_full, queries, 10, ef)


def test_full_recall_on_the_fixture(H, oracle, glove, glove_full):  # template.rs:518-554 with VecType = FullVec
    _, queries = glove
    ix = to_gpu(H, glove_full)
    gt, gd = H.bruteforce_topk(ix._points(), queries, 10)
    ogt, ogd = glove_full.bruteforce(queries, 10)
    assert np.array_equal(gt, ogt) and np.array_equal(bits(gd), bits(ogd))
    hits = 0
    for i, q in enumerate(queries):
        ann = ix.ann_by_vector(q, 10, 100)
        assert ann == glove_full.ann_by_vector(q, 10, 100)
        hits += len(set(ann) & set(gt[i].tolist()))
    assert hits / (len(queries) * 10) > 0.99


@pytest.mark.parametrize("dim,n,m,efc", [(100, 5000, 16, 40), (128, 2500, 12, None), (33, 2000, 5, None), (17, 1500, 20, 50)])
def test_full_search_synthetic(H, oracle, dim, n, m, efc):
    base = synth(n, dim, 64, 1)
    queries = synth(150, dim, 64, 2)
    orc = oracle.Index(m, efc, dim, full=True).insert_bulk(base)
    ix = to_gpu(H, orc)
    for n_, ef in ((10, 1), (10, 10), (10, 64), (10, 150), (100, 100), (3, 50), (10, 400)):
        check_search(H, oracle, orc, queries, n_, ef, ix)


@pytest.mark.parametrize("env", [{"HNSWB200_GENERAL_PATH": "1"}, {"HNSWB200_VIS32": "1"}, {"HNSWB200_VIS_SLOTS": "64"}])
def test_full_search_kernel_variants(H, oracle, glove, glove_full, monkeypatch, env):
    _, queries = glove
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    ix = to_gpu(H, glove_full)
    ids, dists, counts = ix.ann_batch(queries, 10, 100)
    oids, odists, ocounts, _, _ = glove_full.search_batch(queries, 10, 100)
    assert np.array_equal(ids, oids) and np.array_equal(bits(dists), bits(odists)) and np.array_equal(counts, ocounts)


def test_full_search_refuses_non_finite_query(H, oracle, glove, glove_full):
    _, queries = glove
    ix = to_gpu(H, glove_full)
    q = queries[:4].copy()
    q[1, 7] = np.inf
    with pytest.raises(H.HnswB200Error):
        ix.ann_batch(q, 10, 50)
    ids, _, _ = ix.ann_batch(queries[:4], 10, 50)            # and the context is usable afterwards
    assert np.array_equal(ids, glove_full.search_batch(queries[:4], 10, 50)[0])


# ---- brute force ---------------------------------------------------------------------------
@pytest.mark.parametrize("dim,n,k", [(50, 1000, 10), (100, 20000, 100), (128, 7000, 10), (7, 300, 500)])
def test_full_bruteforce_matches_oracle(H, oracle, dim, n, k):
    base = synth(n, dim, 32, 5)
    queries = synth(64, dim, 32, 6)
    pts = H.SimplePoints.new(base, vec_type="full")
    orc = oracle.Index.from_parts(8, 16, dim, 0, base, None, None, np.zeros(n, np.uint8),
                                  [(np.arange(n, dtype=np.uint32), np.zeros(n + 1, np.uint64), np.zeros(0, np.uint32))])
    gt, gd = H.bruteforce_topk(pts, queries, k)
    ogt, ogd = orc.bruteforce(queries, k, threads=8)
    kk = min(k, n)
    assert np.array_equal(gt[:, :kk], ogt[:, :kk]) and np.array_equal(bits(gd[:, :kk]), bits(ogd[:, :kk]))


# ---- build -----------------------------------------------------------------------------------
def test_full_build_batch1_reproduces_oracle_graph(H, oracle, glove, glove_full):
    store, _ = glove
    ix = H.HNSW.new(12, None, 50, vec_type="full").insert_bulk(store, batch=1)
    assert ix.vec_type == "full" and ix.params.ep == glove_full.ep
    vals, levels = ix._points().values()
    ov, ol = glove_full.export_values()
    assert np.array_equal(levels, ol) and np.array_equal(bits(vals), bits(ov))
    assert_same_graph(layers_of(ix), glove_full.export_layers())
    # the context's vector type is left as it was: the next index is a QuantVec one again
    assert H.HNSW.new(12, None, 50).insert_bulk(store[:50], batch=1).vec_type == "quant"


@pytest.mark.parametrize("dim,n,m,efc", [(100, 1200, 16, 60), (128, 700, 6, None), (33, 700, 4, 9)])
def test_full_build_batch1_synthetic(H, oracle, dim, n, m, efc):
    base = synth(n, dim, 16, 11)
    orc = oracle.Index(m, efc, dim, full=True).insert_bulk(base)
    ix = H.HNSW.new(m, efc, dim, vec_type="full").insert_bulk(base, batch=1)
    assert ix.params.ep == orc.ep
    assert_same_graph(layers_of(ix), orc.export_layers())


def test_full_build_batched_quality_and_search_parity(H, oracle):
    base = synth(20000, 100, 256, 21)
    queries = synth(400, 100, 256, 22)
    ix = H.HNSW.new(16, 100, 100, vec_type="full").insert_bulk(base)
    assert ix.assert_param_compliance()
    gt, _ = H.bruteforce_topk(ix._points(), queries, 10)
    ids, dists, counts, st = ix.ann_batch(queries, 10, 64, with_stats=True)
    hits = sum(len(set(gt[i].tolist()) & set(ids[i].tolist())) for i in range(len(queries)))
    assert hits / (10 * len(queries)) > 0.99
    orc = to_oracle(oracle, ix)
    oids, odists, _, ohops, oevals = orc.search_batch(queries, 10, 64, threads=8)
    assert np.array_equal(ids, oids) and np.array_equal(bits(dists), bits(odists))
    assert np.array_equal(st["hops"], ohops) and np.array_equal(st["evals"], oevals)


def test_full_insert_after_build(H, oracle):  # template.rs:479-504
    rng = np.random.default_rng(5)
    a = rng.random((100, 10), dtype=np.float32)
    b = rng.random((100, 10), dtype=np.float32)
    v = rng.random(10, dtype=np.float32)
    orc = oracle.Index(12, None, 10, full=True).insert_bulk(a)
    ix = H.HNSW.new(12, None, 10, vec_type="full").insert_bulk(a, batch=1)
    assert orc.insert_vec(v) == ix.insert_vec(v) == 100
    orc.insert_bulk(b)
    ix.insert_bulk(b, batch=1)
    assert ix.len() == 201 and ix.params.ep == orc.ep and ix.vec_type == "full"
    assert_same_graph(layers_of(ix), orc.export_layers())


# ---- save / load -------------------------------------------------------------------------------
def test_full_save_load_cross_with_oracle(H, oracle, glove, glove_full, tmp_path):
    _, queries = glove
    ix = to_gpu(H, glove_full)
    ix.save(tmp_path / "gpu")
    glove_full.save(tmp_path / "cpu")
    for name in ("points", "params"):                            # point size 1 + 4*dim (point.rs:55-61, full.rs:45-61)
        assert (tmp_path / "gpu" / name).read_bytes() == (tmp_path / "cpu" / name).read_bytes()
    assert len((tmp_path / "gpu" / "points").read_bytes()) == 16 + 1000 * (1 + 4 * 50)
    back = oracle.Index.load(tmp_path / "gpu")                  # the oracle reads what the engine wrote
    assert back.full and back.params() == glove_full.params()
    assert_same_graph(back.export_layers(), glove_full.export_layers())
    ld = H.HNSW.load(tmp_path / "cpu")                           # the engine reads what the oracle wrote
    assert ld.vec_type == "full" and ld.len() == 1000
    a = ld.ann_batch(queries, 10, 100)
    b = glove_full.search_batch(queries, 10, 100)
    assert np.array_equal(a[0], b[0]) and np.array_equal(bits(a[1]), bits(b[1]))
    # SimplePoints serializer in this mode (points.rs:119-146)
    blob = ld._points().serialize()
    assert blob == (tmp_path / "cpu" / "points").read_bytes()
    again = H.SimplePoints.deserialize(blob, vec_type="full")
    assert np.array_equal(bits(again.values()[0]), bits(glove_full.export_values()[0]))


# ---- committed expected outputs (tests/golden/expected_*.npz, written by the oracle): the device engine end to end ----
@pytest.mark.parametrize("name,vec_type", [("quant", "quant"), ("full", "full")])
def test_device_reproduces_committed_expected_outputs(H, glove, name, vec_type):
    """HNSW::new(12, None, 50).insert_bulk(store) in the single-thread order, ann_by_vector at ef 10 / 100 and brute_force_nns
    through the C ABI against the committed golden outputs: graph, ids, distance bits, counters."""
    from golden_check import check_against_golden
    store, queries = glove
    ix = H.HNSW.new(12, None, store.shape[1], vec_type=vec_type).insert_bulk(store, batch=1)

    def search(q, n, ef):
        ids, dists, counts, st = ix.ann_batch(q, n, ef, with_stats=True)
        return ids, dists, counts, st["hops"], st["evals"]

    check_against_golden(name, {"ep": ix.params.ep, "layers": layers_of(ix), "search": search,
                                "bruteforce": lambda q, k: H.bruteforce_topk(ix._points(), q, k)}, queries)
