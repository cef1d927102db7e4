"""GPU nearest-class search vs the CPU oracle: indices must be bit-identical (north_star)."""
import numpy as np
import pytest
import torch

from oracle import nearest_oracle as no

pytestmark = pytest.mark.gpu


def _unit(rng, n, d=300):
    x = rng.standard_normal((n, d)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


@pytest.mark.parametrize("C", [51, 101, 200, 664])
def test_indices_and_distances_bit_exact(C):
    from zeroshotvideoclassification_b200 import ops
    rng = np.random.default_rng(C)
    n = 10000 if C != 664 else 22
    emb, cls = _unit(rng, n), _unit(rng, C)
    dist = no.cosine_distance_table(emb, cls)
    ref = no.topk_lowest_index(dist, 5)
    idx, d = ops.nearest_class(torch.from_numpy(emb).cuda(), torch.from_numpy(cls).cuda(), k=5, return_dist=True)
    idx, d = idx.cpu().numpy(), d.cpu().numpy()
    assert np.array_equal(idx, ref)
    assert np.array_equal(d, np.take_along_axis(dist, ref, axis=1))      # fp64 distances, bit for bit
    top1 = ops.nearest_class(torch.from_numpy(emb).cuda(), torch.from_numpy(cls).cuda(), k=1).cpu().numpy()[:, 0]
    assert np.array_equal(top1, dist.argmin(1))


def test_edge_cases():
    """duplicates of class rows (distance exactly 0), duplicated class rows (lowest index wins), near-ties,
    zero-norm embedding (all-NaN row -> indices 0..k-1 like numpy argsort), odd D, ragged N."""
    from zeroshotvideoclassification_b200 import ops
    rng = np.random.default_rng(7)
    cls = _unit(rng, 101)
    cls[57] = cls[12]                      # duplicated class vector
    emb = np.concatenate([cls[:30], cls[:30] + 1e-4 * _unit(rng, 30), np.zeros((1, 300), np.float32), _unit(rng, 3)])
    dist = no.cosine_distance_table(emb, cls)
    ref = no.topk_lowest_index(dist, 5)
    idx = ops.nearest_class(torch.from_numpy(emb).cuda(), torch.from_numpy(cls).cuda(), k=5).cpu().numpy()
    assert np.array_equal(idx, ref)
    assert idx[12, 0] == 12 and idx[12, 1] == 57
    assert list(idx[60]) == [0, 1, 2, 3, 4]
    # odd embedding width exercises the tail term of the even/odd accumulation
    e2, c2 = rng.standard_normal((17, 301)).astype(np.float32), rng.standard_normal((9, 301)).astype(np.float32)
    got = ops.nearest_class(torch.from_numpy(e2).cuda(), torch.from_numpy(c2).cuda(), k=3, return_dist=True)
    d2 = no.cosine_distance_table(e2, c2)
    assert np.array_equal(got[0].cpu().numpy(), no.topk_lowest_index(d2, 3))
    assert np.array_equal(got[1].cpu().numpy(), np.sort(d2, axis=1)[:, :3])
    # empty batch
    assert ops.nearest_class(torch.zeros((0, 300)).cuda(), torch.from_numpy(cls).cuda(), k=1).shape == (0, 1)


@pytest.mark.parametrize("N,C,D", [(37, 1500, 301), (1100, 700, 300), (25, 130, 64), (3000, 9, 302), (1, 5, 3)],
                         ids=["cluster8_two_rounds_odd_D", "many_rows_wide_table", "second_row_block", "tiny_table", "one_row"])
def test_kernel_variants_bit_exact(N, C, D):
    """Every launch shape of zsv_nearest_class: the cluster kernel (class tiles over up to 8 CTAs, several rounds,
    8-row fallback for wide tables, more than one row block) and the per-block kernel with a ragged last block."""
    from zeroshotvideoclassification_b200 import ops
    rng = np.random.default_rng(N + C)
    emb, cls = _unit(rng, N, D), _unit(rng, C, D)
    dist = no.cosine_distance_table(emb, cls)
    k = min(5, C)
    ref = no.topk_lowest_index(dist, k)
    idx, d = ops.nearest_class(torch.from_numpy(emb).cuda(), torch.from_numpy(cls).cuda(), k=k, return_dist=True)
    assert np.array_equal(idx.cpu().numpy(), ref)
    assert np.array_equal(d.cpu().numpy(), np.take_along_axis(dist, ref, axis=1))


def test_class_overlap_filter_matches_scipy():
    """filter_overlapping_classes (auxiliary/auxiliary_dataset.py:141-144): the kept-class mask is identical to scipy's."""
    cdist = pytest.importorskip("scipy.spatial.distance").cdist
    from zeroshotvideoclassification_b200 import class_overlap_mask
    rng = np.random.default_rng(5)
    train, test = _unit(rng, 664), _unit(rng, 101)
    train[:20] = test[:20] + 0.02 * _unit(rng, 20)        # near-duplicates of test classes
    train[20:25] = test[20:25]                             # exact overlaps (distance 0)
    for tau in (0.0, 0.05, 0.5, 0.9):
        ref = cdist(train, test, "cosine").min(1) > tau
        got = class_overlap_mask(train, test, tau).cpu().numpy()
        assert np.array_equal(got, ref), tau


def test_sharded_accuracy_equals_compute_accuracy():
    """dist.compute_accuracy_sharded (single process: one shard) and accuracy.count_correct agree with
    compute_accuracy and with the oracle's top-1 / top-5 (main.py:316-325)."""
    from zeroshotvideoclassification_b200 import compute_accuracy
    from zeroshotvideoclassification_b200 import accuracy, dist as zdist
    rng = np.random.default_rng(11)
    cls = _unit(rng, 101)
    labels = rng.integers(0, 101, 3000)
    true = cls[labels]
    pred = true + 0.8 * _unit(rng, 3000)
    ref = compute_accuracy(pred, cls, true)
    assert zdist.compute_accuracy_sharded(pred, cls, true) == ref
    c = accuracy.count_correct(pred[:0], cls, true[:0]).cpu().tolist()
    assert c == [0, 0, 0]
    parts = sum(accuracy.count_correct(pred[lo:hi], cls, true[lo:hi]) for lo, hi in
                (zdist.shard_rows(3000, r, 7) for r in range(7))).cpu().tolist()
    assert parts[2] == 3000 and (100.0 * parts[0] / 3000, 100.0 * parts[1] / 3000) == pytest.approx(ref, abs=1e-4)
    dist = no.cosine_distance_table(pred.astype(np.float32), cls)
    top5 = no.topk_lowest_index(dist, 5)
    assert ref[0] == pytest.approx(100.0 * float((top5[:, 0] == labels).mean()), abs=1e-4)
    assert ref[1] == pytest.approx(100.0 * float((top5 == labels[:, None]).any(1).mean()), abs=1e-4)
