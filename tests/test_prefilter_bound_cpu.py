"""Host-side check of the search kernel's integer pre-filter (csrc/search_fast.cuh: FastQuery::prefilter / prefilter2).

The filter may only REJECT a candidate whose exact key is above the admission bound (the reference evaluates such a
candidate and drops it, hnsw/src/template.rs search_layer via searcher.rs:74-94), so that skipping the exact arithmetic
cannot change a result.  This test restates the filter's float32 expression operation by operation in numpy (every
numpy float32 operation rounds to nearest like the kernel's __f*_rn intrinsics), computes the exact distance with the
oracle (vectors/src/quant.rs:14-37), and checks, for the LARGEST bound the filter would still reject at (and a few ulps
around it):        rejected  ==>  exact distance > bound   (strictly: an equal distance can still win on the id).
Data regimes: unit-norm clustered rows (C2), non-negative heavy-tailed rows (C3), tiny and huge magnitudes,
near-duplicates (the cancellation case: d^2 << Sum x^2 + Sum y^2), constant rows (delta = 0).
"""
import numpy as np
import pytest

F = np.float32


def _lane_sum(vals):
    """Sum as quantise_kernel / init_filter do it: lane i adds elements i, i + 32, ...; then a xor butterfly."""
    acc = np.zeros(32, F)
    for i, v in enumerate(vals):
        acc[i % 32] = F(acc[i % 32] + F(v))
    for o in (16, 8, 4, 2, 1):
        acc = (acc + acc[np.arange(32) ^ o]).astype(F)
    return acc[0]


def _aux(codes, mn, dl):
    y = (codes.astype(F) * F(dl) + F(mn)).astype(F)          # __fadd_rn(__fmul_rn(c, dl), mn)
    return _lane_sum(y), _lane_sum((y * y).astype(F))


def _estimate(qc, qmn, qdl, qq, bc_codes, bmn, bdl, sy, sy2):
    """est and nrm of FastQuery::prefilter for one pair (float32 throughout, integer dot exact)."""
    dot = int(np.dot(qc.astype(np.int64), bc_codes.astype(np.int64)))
    sc = int(qc.astype(np.int64).sum())
    ax, ay, az, aw = F(qdl), F(F(qdl) * F(sc)), F(qmn), F(qq)
    t = F(F(F(F(ax * F(bdl)) * F(dot)) + F(ay * F(bmn))) + F(az * F(sy)))
    nrm = F(aw + F(sy2))
    est = F(nrm + F(F(-2.0) * t))
    return est, nrm


def _rejects(est, nrm, wd):
    T = F(F(F(wd) * F(wd)) * F(1.0001))
    return bool(est > F(T + F(F(1e-5) * nrm)))


def _regimes(rng, dim):
    n = 96
    centres = rng.standard_normal((8, dim)).astype(F)
    a = centres[rng.integers(0, 8, n)] + 0.3 * rng.standard_normal((n, dim)).astype(F)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    yield "unit-norm clustered", a.astype(F)
    yield "non-negative heavy-tailed", np.minimum(rng.gamma(0.6, 30.0, (n, dim)), 255.0).astype(F)
    yield "tiny", (1e-4 * rng.standard_normal((n, dim))).astype(F)
    yield "huge", (1e4 * rng.standard_normal((n, dim))).astype(F)
    base = rng.standard_normal((1, dim)).astype(F)
    yield "near-duplicates", (base + 1e-3 * rng.standard_normal((n, dim))).astype(F)
    mixed = rng.standard_normal((n, dim)).astype(F)
    mixed[::4] = F(0.25)  # constant rows: delta = 0, every code 0
    yield "with constant rows", mixed


@pytest.mark.parametrize("dim", [100, 128, 96])
def test_prefilter_only_rejects_candidates_above_the_bound(oracle, dim):
    rng = np.random.default_rng(100 + dim)
    checked = rejected_somewhere = 0
    for name, rows in _regimes(rng, dim):
        quant = [oracle.quantise(r) for r in rows]
        aux = [_aux(c, mn, dl) for c, mn, dl in quant]
        deq = [oracle.dequantise(c, mn, dl) for c, mn, dl in quant]
        nq = 12
        for qi in range(nq):
            qc, qmn, qdl = quant[qi]
            qq = _lane_sum((deq[qi] * deq[qi]).astype(F))
            for bi in range(nq, len(rows)):
                bc, bmn, bdl = quant[bi]
                sy, sy2 = aux[bi]
                est, nrm = _estimate(qc, qmn, qdl, qq, bc, bmn, bdl, sy, sy2)
                d = F(oracle.dist_quant(quant[qi], quant[bi]))
                # the largest bound that is still rejected lies near sqrt((est - 1e-5 nrm) / 1.0001)
                room = F(est - F(F(1e-5) * nrm))
                cands = [d, np.nextafter(d, F(0)), np.nextafter(d, F(np.inf))]
                if room > 0:
                    w = F(np.sqrt(F(room / F(1.0001))))
                    for _ in range(6):
                        w = np.nextafter(w, F(0))
                    for _ in range(13):
                        cands.append(w)
                        w = np.nextafter(w, F(np.inf))
                for wd in cands:
                    if not np.isfinite(wd) or wd < 0:
                        continue
                    checked += 1
                    if _rejects(est, nrm, wd):
                        rejected_somewhere += 1
                        assert d > wd, (name, dim, qi, bi, float(d), float(wd), float(est), float(nrm))
    assert checked > 30000 and rejected_somewhere > 3000  # the test exercised the rejecting side of the predicate


def test_prefilter_passes_nan_and_inf_on_to_the_exact_arithmetic():
    # "everything else (NaN / inf included) goes to the exact arithmetic": the predicate is !(est > ...)
    for est, nrm in ((F(np.nan), F(1.0)), (F(1.0), F(np.nan)), (F(np.inf), F(np.inf)), (F(-np.inf), F(1.0))):
        assert not _rejects(est, nrm, F(0.5))
    assert not _rejects(F(1.0), F(1.0), F(np.nan))
