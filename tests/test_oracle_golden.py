"""CPU: pin the oracle (and the drop-in module tree's init) against fixtures generated from the UNMODIFIED
reference by oracle/make_golden.py.  The reference has no tests/golden vectors of its own,
so these fixtures -- outputs of the reference itself -- are what makes the oracle trustworthy."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import nearest_oracle as no
from oracle import video_oracle as vo
from oracle.make_golden import checksum, synthetic_batch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return json.load(open(os.path.join(GOLD, name)))


def _close(a, b, rtol, atol=0.0):
    return abs(a - b) <= atol + rtol * max(abs(a), abs(b))


def _check_sum(got: torch.Tensor, gold: dict, rtol, what):
    c = checksum(got)
    scale = gold["abs"] / max(1, got.numel())                 # mean |value|
    assert _close(c["abs"], gold["abs"], rtol), (what, c["abs"], gold["abs"])
    assert _close(c["sq"], gold["sq"], 2 * rtol), (what, c["sq"], gold["sq"])
    for s, g in zip(c["samples"], gold["samples"]):
        assert abs(s - g) <= 20 * rtol * max(scale, abs(g)) + 1e-12, (what, s, g)


@pytest.fixture(scope="module")
def seeded_model():
    from zeroshotvideoclassification_b200 import video_models as vm
    torch.manual_seed(0)
    return vm.get_network(vm.default_opt("r2plus1d_18"))


def test_module_tree_matches_reference_state_dict(seeded_model):
    gold = _load("r2plus1d_state_dict_keys.json")
    got = {k: list(v.shape) for k, v in seeded_model.state_dict().items()}
    assert list(got.keys()) == list(gold.keys())          # same names in the same order
    assert got == gold
    n_backbone = sum(v.numel() for k, v in seeded_model.model.named_parameters())
    assert n_backbone == 31_505_325                       # torchvision meta for r2plus1d_18


def test_init_is_bitwise_the_reference_init(seeded_model):
    """Same construction order + same init calls => the same RNG stream => identical weights for a given seed
    (resnet.py:226-236, network.py:500-530)."""
    gold = _load("r2plus1d_init_seed0.json")
    for k, v in seeded_model.state_dict().items():
        if not v.is_floating_point():
            continue
        c = checksum(v)
        assert c["sum"] == gold[k]["sum"] and c["abs"] == gold[k]["abs"] and c["samples"] == gold[k]["samples"], k


@pytest.mark.parametrize("fixture", ["r2plus1d_step_small.json", "r2plus1d_step_bs2_16x112.json"])
def test_oracle_reproduces_reference_step(seeded_model, fixture):
    """Oracle forward+backward (fp32 CPU) == the reference's own forward+backward on the same weights/clips."""
    gold = _load(fixture)
    cfg = gold["config"]
    sd = {k: v.detach().clone() for k, v in seeded_model.state_dict().items()}
    x, z, _ = synthetic_batch(cfg["B"], cfg["T"], cfg["H"], cfg["W"], cfg["seed"] + 100)
    trace = {}
    emb, loss, grads = vo.train_step_grads(sd, x, z, trace=trace)
    assert torch.allclose(emb, torch.tensor(gold["emb"]), atol=2e-5, rtol=1e-4)
    assert _close(float(loss), gold["loss"], 1e-5)
    # per-layer activations
    for name, g in gold["acts"].items():
        _check_sum(trace[name], g, 2e-4, name)
    # gradients of every live parameter; dead ones (network.py:500-517) get none
    assert sorted(k for k in gold["grads"]) == sorted(grads.keys())
    for name, g in gold["grads"].items():
        # BN gamma/beta gradients are fp32 sums over ~1e6 positions with heavy cancellation: two fp32 CPU
        # implementations already differ by ~3e-3 there
        _check_sum(grads[name], g, 1e-2 if grads[name].dim() == 1 else 2e-3, name)
    for name in gold["dead"]:
        assert name not in grads
    # running statistics after one training-mode forward
    for name, g in gold["bn_after"].items():
        _check_sum(sd[name], g, 1e-4, name)


def test_nearest_oracle_bit_exact_vs_scipy_fixture():
    z = np.load(os.path.join(GOLD, "nearest_scipy.npz"))
    for C in (51, 101, 200):
        emb, cls = z[f"emb_{C}"], z[f"cls_{C}"]
        d = no.cosine_distance_table(emb, cls)
        assert np.array_equal(d, z[f"dist_{C}"])                       # fp64, bit for bit
        assert np.array_equal(no.nearest_class(emb, cls, 1)[:, 0], z[f"argmin_{C}"])
        assert np.array_equal(no.nearest_class(emb, cls, 5), z[f"top5_{C}"])
        assert np.all(d[:10, :10].diagonal() < 1e-15)                  # exact hits: 0 up to sqrt rounding


def test_nearest_oracle_vs_live_scipy():
    cdist = pytest.importorskip("scipy.spatial.distance").cdist
    rng = np.random.default_rng(11)
    for D in (300, 301, 7, 1):
        e = rng.standard_normal((64, D)).astype(np.float32)
        c = rng.standard_normal((33, D)).astype(np.float32)
        assert np.array_equal(no.cosine_distance_table(e, c), cdist(e, c, "cosine")), D
    # zero-norm row -> NaN distances; numpy argmin gives 0 and argsort the identity
    e = np.zeros((1, 300), np.float32)
    c = rng.standard_normal((5, 300)).astype(np.float32)
    assert np.isnan(no.cosine_distance_table(e, c)).all()
    assert list(no.nearest_class(e, c, 5)[0]) == [0, 1, 2, 3, 4]


def test_compute_accuracy_restatement():
    """main.py:316-325 on a constructed case with known answer."""
    rng = np.random.default_rng(3)
    cls = rng.standard_normal((20, 300)).astype(np.float32)
    labels = rng.integers(0, 20, 200)
    true = cls[labels]
    pred = true + 0.01 * rng.standard_normal(true.shape).astype(np.float32)
    pred[:50] = cls[(labels[:50] + 1) % 20]                 # 25 % wrong on purpose
    top1, top5 = no.compute_accuracy(pred, cls, true)
    assert top1 == 75.0 and top5 >= top1


# ---- r3d_18: the third backbone network.get_network can select (network.py:28-30, resnet.py:293-314) ----
@pytest.fixture(scope="module")
def seeded_r3d():
    from zeroshotvideoclassification_b200 import video_models as vm
    torch.manual_seed(0)
    return vm.get_network(vm.default_opt("r3d_18"))


def test_r3d_module_tree_and_init_match_reference(seeded_r3d):
    gold = _load("r3d_state_dict_keys.json")
    got = {k: list(v.shape) for k, v in seeded_r3d.state_dict().items()}
    assert list(got.keys()) == list(gold.keys()) and got == gold
    assert sum(v.numel() for v in seeded_r3d.model.parameters()) == 33_371_472   # torchvision meta for r3d_18
    init = _load("r3d_init_seed0.json")
    for k, v in seeded_r3d.state_dict().items():
        if v.is_floating_point():
            c = checksum(v)
            assert c["sum"] == init[k]["sum"] and c["samples"] == init[k]["samples"], k


def test_oracle_reproduces_reference_r3d_step(seeded_r3d):
    gold = _load("r3d_step_small.json")
    cfg = gold["config"]
    sd = {k: v.detach().clone() for k, v in seeded_r3d.state_dict().items()}
    x, z, _ = synthetic_batch(cfg["B"], cfg["T"], cfg["H"], cfg["W"], cfg["seed"] + 100)
    trace = {}
    emb, loss, grads = vo.train_step_grads(sd, x, z, trace=trace, arch="r3d_18")
    assert torch.allclose(emb, torch.tensor(gold["emb"]), atol=2e-5, rtol=1e-4)
    assert _close(float(loss), gold["loss"], 1e-5)
    for name, g in gold["acts"].items():
        _check_sum(trace[name], g, 2e-4, name)
    assert sorted(gold["grads"]) == sorted(grads.keys())
    for name, g in gold["grads"].items():
        _check_sum(grads[name], g, 1e-2 if grads[name].dim() == 1 else 2e-3, name)
    for name in gold["dead"]:
        assert name not in grads


# ---- C3D (network.py:95-180): pinned against the reference's own module on the full 16x112x112 clip ----
@pytest.fixture(scope="module")
def seeded_c3d():
    from zeroshotvideoclassification_b200 import video_models as vm
    torch.manual_seed(0)
    return vm.get_network(vm.default_opt("c3d"))


def test_c3d_module_tree_and_init_match_reference(seeded_c3d):
    gold = _load("c3d_state_dict_keys.json")
    got = {k: list(v.shape) for k, v in seeded_c3d.state_dict().items()}
    assert list(got.keys()) == list(gold.keys()) and got == gold        # conv1 .. conv5b, fc6, fc7, fc8, regressor
    assert sum(v.numel() for v in seeded_c3d.parameters()) == 81_220_115
    init = _load("c3d_init_seed0.json")
    for k, v in seeded_c3d.state_dict().items():
        c = checksum(v)
        assert c["sum"] == init[k]["sum"] and c["abs"] == init[k]["abs"] and c["samples"] == init[k]["samples"], k


def test_oracle_reproduces_reference_c3d_step(seeded_c3d):
    """oracle.c3d_forward + autograd == network.C3D forward + backward (Dropout p = 0) on the same weights / clips:
    every conv, pool and linear output, the embedding, the loss and every live gradient."""
    gold = _load("c3d_step_bs2_16x112.json")
    cfg = gold["config"]
    sd = {k: v.detach().clone() for k, v in seeded_c3d.state_dict().items()}
    x, z, _ = synthetic_batch(cfg["B"], cfg["T"], cfg["H"], cfg["W"], cfg["seed"] + 100)
    trace = {}
    emb, loss, grads = vo.c3d_train_step_grads(sd, x, z, trace=trace)
    assert torch.allclose(emb, torch.tensor(gold["emb"]), atol=2e-5, rtol=1e-4)
    assert _close(float(loss), gold["loss"], 1e-5)
    assert len(gold["acts"]) == 15
    for name, g in gold["acts"].items():
        _check_sum(trace[name], g, 2e-4, name)
    assert sorted(gold["grads"]) == sorted(grads.keys())
    for name, g in gold["grads"].items():
        _check_sum(grads[name], g, 2e-3, name)
    assert gold["dead"] == ["fc7.bias", "fc7.weight", "fc8.bias", "fc8.weight"]
    for name in gold["dead"]:
        assert name not in grads


def test_transform_oracle_matches_reference_transform():
    """oracle/transform_oracle.py == the reference's own transform functions (fixture from auxiliary/transforms.py)."""
    from oracle import transform_oracle as to
    z = np.load(os.path.join(GOLD, "clip_transform.npz"))
    f = torch.from_numpy(z["landscape_frames"])
    hr, wr = 128, int(np.floor(90 * (128.0 / 68)))
    got = to.clip_transform(f, to.center_crop_origin(hr, wr), flip=False)
    assert np.array_equal(got.numpy(), z["landscape_val"])                  # same torch ops in the same order
    f = torch.from_numpy(z["portrait_frames"])
    got = to.clip_transform(f, (5, 9), flip=True)
    assert np.array_equal(got.numpy(), z["portrait_train_5_9_flip"])
    assert tuple(z["portrait_resized_hw"]) == (int(np.floor(200 * (128.0 / 150))), 128)
