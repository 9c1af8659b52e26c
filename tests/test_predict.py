"""Batched predict + top-k (BASELINE config 5; predict.cu:17-29,49-63 for all users at once)."""
import numpy as np
import pytest

import cu2rec_b200 as cu
import oracle as O

pytestmark = pytest.mark.gpu


def _model(rng, U, I, k):
    P = (rng.standard_normal((U, k)) / np.sqrt(k)).astype(np.float32)
    Q = (rng.standard_normal((I, k)) / np.sqrt(k)).astype(np.float32)
    return P, Q, (rng.standard_normal(U) * 0.3).astype(np.float32), (rng.standard_normal(I) * 0.3).astype(np.float32)


@pytest.mark.parametrize("U,I,k,topk", [(300, 500, 32, 5), (1000, 1777, 64, 10), (777, 2000, 128, 10), (130, 127, 96, 3),
                                        (257, 4000, 128, 16),
                                        # the reference's own factor counts (config.h:27 default 50; experiments/cu2rec.sh:10
                                        # {50, 300}) and odd ones: rows are zero-padded to whole swizzle rows, k > 128 streams
                                        (500, 900, 50, 10), (300, 1500, 300, 10), (200, 700, 1, 4), (200, 700, 7, 8),
                                        (260, 1300, 130, 12), (140, 600, 512, 6),
                                        # more than 16 per user: further passes over the catalogue, final order kernel
                                        (300, 2500, 64, 20), (260, 3000, 128, 64), (150, 2100, 300, 40), (90, 1000, 50, 128)])
def test_topk_matches_brute_force(U, I, k, topk):
    """Items and their order equal the CPU brute force (exact fp32 scores in the reference's op
    order, ties by item id); scores are bit-identical."""
    rng = np.random.RandomState(U + k)
    P, Q, ub, ib = _model(rng, U, I, k)
    tr, _ = cu.synth_ratings(U, I, min(U * I // 4, 40 * U), rank=4, noise=0.3, seed=U)
    ex = cu.createSparseMatrix(tr, U, I)
    items, scores, ms = cu.predict_topk(P, Q, ub, ib, 3.5, topk, exclude=ex)
    want_i, want_s = O.predict_topk(P, Q, ub, ib, 3.5, topk, exclude=(ex.indptr, ex.indices))
    assert np.array_equal(items, want_i)
    assert np.array_equal(scores.view(np.uint32), want_s.view(np.uint32))
    # nothing the user already rated is recommended
    rated = set(zip(tr["user"].tolist(), tr["item"].tolist()))
    assert not any((u, int(i)) in rated for u in range(U) for i in items[u])
    assert ms["candidates_ms"] > 0


def test_topk_without_exclusion_and_short_catalogue():
    rng = np.random.RandomState(3)
    P, Q, ub, ib = _model(rng, 200, 300, 64)
    items, scores, _ = cu.predict_topk(P, Q, ub, ib, 3.0, 10)
    want_i, want_s = O.predict_topk(P, Q, ub, ib, 3.0, 10)
    assert np.array_equal(items, want_i) and np.array_equal(scores.view(np.uint32), want_s.view(np.uint32))
    # a user who rated all but 4 items gets 4 recommendations, the rest is -1 / NaN
    U, I = 40, 200
    P, Q, ub, ib = _model(rng, U, I, 32)
    rows = [(0, i, 3.0) for i in range(I - 4)] + [(u, u, 4.0) for u in range(1, U)]
    ex = cu.createSparseMatrix(np.array(rows, dtype=cu.RATING_DTYPE), U, I)
    items, scores, _ = cu.predict_topk(P, Q, ub, ib, 3.0, 10, exclude=ex)
    assert sorted(items[0][:4].tolist()) == [196, 197, 198, 199] and (items[0][4:] == -1).all() and np.isnan(scores[0][4:]).all()
    want_i, _ = O.predict_topk(P, Q, ub, ib, 3.0, 10, exclude=(ex.indptr, ex.indices))
    assert np.array_equal(items, want_i)


def test_topk_rejects_unsupported_shapes():
    rng = np.random.RandomState(4)
    P, Q, ub, ib = _model(rng, 10, 10, 516)
    with pytest.raises(cu._lib.Cu2bError):
        cu.predict_topk(P, Q, ub, ib, 3.0, 5)
    P, Q, ub, ib = _model(rng, 10, 10, 64)
    with pytest.raises(cu._lib.Cu2bError):
        cu.predict_topk(P, Q, ub, ib, 3.0, 129)


def test_topk_longer_than_the_catalogue_is_padded():
    rng = np.random.RandomState(5)
    P, Q, ub, ib = _model(rng, 70, 30, 50)
    items, scores, _ = cu.predict_topk(P, Q, ub, ib, 3.0, 40)
    want_i, want_s = O.predict_topk(P, Q, ub, ib, 3.0, 40)
    assert np.array_equal(items, want_i) and (items[:, 30:] == -1).all() and np.isnan(scores[:, 30:]).all()
    assert np.array_equal(scores[:, :30].view(np.uint32), want_s[:, :30].view(np.uint32))
