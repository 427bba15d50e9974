"""Activation-trace producer (torch forward hooks) and the reference's badge layout on disk
(activation_persistor.py:13-72, handler_model.py:175-206) — CPU tests."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from simple_tip_b200.core import activations as A  # noqa: E402


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv2d(1, 3, 3), torch.nn.ReLU(), torch.nn.Flatten(), torch.nn.Linear(3 * 6 * 6, 5),
                               torch.nn.Softmax(dim=1))


def test_transparent_model_returns_the_selected_layer_outputs():
    model = _model()
    x = torch.randn(23, 1, 8, 8)
    tm = A.TransparentModel(model, [1, 3], include_last_layer=True)
    acts = tm.get_activations(x)
    with torch.no_grad():
        h1 = model[1](model[0](x))
        h3 = model[3](model[2](h1))
        out = model[4](h3)
    assert len(acts) == 3
    assert torch.equal(acts[0], h1) and torch.equal(acts[1], h3) and torch.equal(acts[2], out)
    assert acts[0].shape == (23, 3, 6, 6)
    # by name, without the output
    by_name = A.TransparentModel(model, ["3"], include_last_layer=False).get_activations(x)
    assert len(by_name) == 1 and torch.equal(by_name[0], h3)
    # collected in badges == one pass; walk yields badge by badge
    coll = tm.collect(x, batch_size=10)
    assert all(torch.allclose(c, a, atol=1e-6) for c, a in zip(coll, acts)) and coll[0].shape[0] == 23
    badges = list(tm.walk_activations([x[:10], x[10:20], x[20:]]))
    assert [b[0].shape[0] for b in badges] == [10, 10, 3]
    # training mode is restored and hooks are gone
    model.train()
    tm.get_activations(x[:2])
    assert model.training and all(len(m._forward_hooks) == 0 for m in model.children())
    with pytest.raises(ValueError, match="No activation layers specified"):
        A.TransparentModel(model, None)
    with pytest.raises(IndexError):
        A.TransparentModel(model, [9])


def test_badges_on_disk_use_the_reference_layout(tmp_path):
    model = _model()
    x, y = torch.randn(250, 1, 8, 8), torch.arange(250) % 5
    tm = A.TransparentModel(model, [1, 3], include_last_layer=False)
    n = A.persist_badges(str(tmp_path), "mnist", 7, "test_nominal", tm, x, y)
    assert n == 3
    base = os.path.join(str(tmp_path), "activations", "mnist", "model_7", "test_nominal")
    assert sorted(os.listdir(base)) == ["labels", "layer_0", "layer_1"]
    assert sorted(os.listdir(os.path.join(base, "layer_0"))) == ["badge_0.npy", "badge_1.npy", "badge_2.npy"]
    assert np.load(os.path.join(base, "layer_1", "badge_2.npy")).shape == (50, 5)
    assert np.array_equal(np.load(os.path.join(base, "labels", "badge_1.npy")), y[100:200].numpy())
    acts, labels = A.load_badges(str(tmp_path), "mnist", 7, "test_nominal")
    want = tm.collect(x, batch_size=100)
    assert np.array_equal(labels, y.numpy())
    assert all(np.allclose(a, w.numpy(), atol=1e-6) for a, w in zip(acts, want))
    # the flattened form the scorers use (surprise.py:62-66) has one row per input
    from simple_tip_b200.core.surprise import _flatten_layers

    assert _flatten_layers(acts).shape == (250, 3 * 6 * 6 + 5)


def test_result_files_use_the_reference_names_and_reload_as_apfd(tmp_path):
    """eval_prioritization.py:22-58 writes, eval_apfd_table.py:43-108 reads: scores are ranked by np.argsort(-scores),
    CAM orders are taken as stored, names carry metric / parameter / model id."""
    from simple_tip_b200.core import results as R
    from simple_tip_b200.core.apfd import apfd_from_order

    out = str(tmp_path)
    rng = np.random.default_rng(3)
    n = 50
    mis = rng.random(n) < 0.3
    mis[0] = True
    R.persist(out, "mnist", "nominal", "is_misclassified", 0, mis)
    dsa = rng.random(n)
    cam_order = rng.permutation(n)
    R.persist_tip(out, "mnist", "nominal", 0, "dsa", dsa, cam_order, times=[1.0, 2.0, 3.0, 4.0])
    nac = rng.integers(0, 40, size=n)
    R.persist_tip(out, "mnist", "nominal", 0, "NAC_0.75", nac, cam_order[::-1].copy())
    R.persist_tip(out, "mnist", "nominal", 0, "pc-lsa", -dsa)
    R.persist(out, "mnist", "nominal", "uncertainty_deep_gini", 0, dsa * 2)
    R.persist_tip(out, "mnist", "ood", 0, "dsa", dsa)                       # another dataset: must not be picked up
    assert sorted(os.listdir(os.path.join(out, "priorities")))[:3] == [
        "mnist_nominal_0_NAC_0.75_cam_order.npy", "mnist_nominal_0_NAC_0.75_scores.npy", "mnist_nominal_0_dsa_cam_order.npy"]
    assert np.array_equal(R.load(out, "mnist", "nominal", "dsa_scores", 0), dsa)
    assert R.load_times(out, "mnist", "nominal", 0, "dsa") == [1.0, 2.0, 3.0, 4.0]
    apfd = R.load_apfd_values(out, "mnist", "nominal")
    assert set(apfd) == {"dsa", "dsa-cam", "NAC_0.75", "NAC_0.75-cam", "pc-lsa", "deep_gini"}
    assert apfd["dsa"][0] == apfd_from_order(mis, np.argsort(-dsa))
    assert apfd["dsa-cam"][0] == apfd_from_order(mis, cam_order)
    assert apfd["NAC_0.75-cam"][0] == apfd_from_order(mis, cam_order[::-1])
    assert apfd["deep_gini"][0] == apfd["dsa"][0]
