"""Comparison of an engine's answers on the reference's fixture with the committed expected outputs
(tests/golden/expected_{quant,full}.npz, written by tests/golden/make_expected.py).  `engine` is a dict of callables so that the
CPU suite (the oracle rebuilt from source) and the GPU suite (the device engine through the C ABI) run the very same checks."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_expected(name):
    return dict(np.load(os.path.join(GOLDEN, f"expected_{name}.npz")))


def check_against_golden(name, engine, queries):
    """engine: {"ep": int, "layers": [(ids, off, nbrs), ...], "search": f(queries, n, ef) -> (ids, dists, counts, hops, evals),
    "bruteforce": f(queries, k) -> (ids, dists)}"""
    exp = load_expected(name)
    assert int(engine["ep"]) == int(exp["ep"])
    assert len(engine["layers"]) == int(exp["nb_layers"])
    for l, (ids, off, nb) in enumerate(engine["layers"]):
        assert np.array_equal(ids, exp[f"layer{l}_ids"]), f"layer {l}: node ids"
        assert np.array_equal(off, exp[f"layer{l}_off"]), f"layer {l}: degrees"
        assert np.array_equal(nb, exp[f"layer{l}_nbrs"]), f"layer {l}: neighbours"
    for ef in (10, 100):
        ids, dists, counts, hops, evals = engine["search"](queries, 10, ef)
        assert np.array_equal(ids, exp[f"ef{ef}_ids"]), f"ef={ef}: neighbour ids"
        assert np.array_equal(np.ascontiguousarray(dists, np.float32).view(np.uint32), exp[f"ef{ef}_dist_bits"]), f"ef={ef}: distances"
        assert np.array_equal(counts, exp[f"ef{ef}_counts"]), f"ef={ef}: counts"
        assert np.array_equal(hops, exp[f"ef{ef}_hops"]), f"ef={ef}: hops"
        assert np.array_equal(evals, exp[f"ef{ef}_evals"]), f"ef={ef}: evaluations"
    gt, gd = engine["bruteforce"](queries, 10)
    assert np.array_equal(gt, exp["bf_ids"]), "brute force: ids"
    assert np.array_equal(np.ascontiguousarray(gd, np.float32).view(np.uint32), exp["bf_dist_bits"]), "brute force: distances"
