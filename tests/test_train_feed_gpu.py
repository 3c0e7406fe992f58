"""GPU parity: consumers of device-resident targets -- the TVF value loss with its gradient
(rl/tvf.py:32-77) against the reference's own autograd results, and the minibatch row gather (rl/rollout.py:2349-2372)
bit for bit against NumPy indexing."""
import types

import numpy as np
import pytest
import torch

import oracle
from ppo_b200 import rollout, tvf
from util import TOL, rel_err

pytestmark = pytest.mark.gpu

LOSS_CASES = [("all", "off", None, None), ("hweighted", "h_weighted", None, None), ("single", "h_weighted", [5], 5),
              ("prefix", "off", list(range(8)), -7)]


def _module(g, weighting, dropout=0.0):
    log = types.SimpleNamespace(watch_mean=lambda *a, **k: None)
    runner = types.SimpleNamespace(ext_rewards=np.zeros([2, 3], np.float32), value_heads=["ext", "int"],
                                   tvf_horizons=list(g["loss_in_horizons"]), tvf_weights=g["loss_in_weights"], log=log)
    args = types.SimpleNamespace(tvf=types.SimpleNamespace(coef=0.7, head_weighting=weighting, max_horizon=1000,
                                                           horizon_dropout=dropout),
                                 tvf_return_n_step=20, value_opt=types.SimpleNamespace(epochs=2))
    return tvf.TVFRunnerModule(runner, args)


@pytest.mark.parametrize("tag,weighting,required,single", LOSS_CASES)
def test_value_loss_golden(golden_norm, tag, weighting, required, single):
    """on_train_value_minibatch through the module mirror, forward and backward (rl/rollout.py:1553-1555)."""
    g = golden_norm
    sel = slice(None) if required is None else required
    p = torch.tensor(g["loss_in_predictions"][:, sel, :], device="cuda", requires_grad=True)
    loss = _module(g, weighting).on_train_value_minibatch(
        {"tvf_value": p}, {"tvf_returns": torch.from_numpy(g["loss_in_targets"]).cuda()},
        required_tvf_heads=required, single_value_head=single)
    assert loss.shape == (p.shape[0],) and loss.is_cuda
    (loss * 0.25).mean().backward()
    assert rel_err(loss.detach().cpu().numpy(), g[f"loss_{tag}"]) <= 1e-6
    assert rel_err(p.grad.cpu().numpy(), g[f"loss_{tag}_grad"]) <= 1e-6
    assert not p.grad[:, :, 1].any()  # only the "ext" head receives gradient (:43)


@pytest.mark.parametrize("rows,heads", [(1, 1), (7, 33), (4096, 128), (1000, 200)])
def test_value_loss_vs_oracle(rows, heads):
    rng = np.random.RandomState(rows + heads)
    hz = np.sort(rng.choice(30000, heads, replace=False))
    t = rng.normal(0, 3, [rows, heads]).astype(np.float32)
    q = rng.normal(0, 3, [rows, heads]).astype(np.float32)
    w = rng.randint(1, 4, heads).astype(np.float32)
    mask = (rng.random_sample([rows, heads]) < 0.8).astype(np.float32) / 0.8
    ref_loss, ref_grad = oracle.tvf_value_loss(t, q, w, hz, 0.5, head_weighting="h_weighted", max_horizon=30000,
                                               n_step=20, mask=mask, loss_scale=1.0)
    weights = tvf.tvf_head_loss_weights(w, hz, None, None, "h_weighted", 30000, 20)
    p = torch.tensor(q, device="cuda", requires_grad=True)
    loss = tvf.tvf_value_loss(torch.from_numpy(t).cuda(), p, weights, 0.5, mask=torch.from_numpy(mask).cuda())
    loss.mean().backward()
    assert rel_err(loss.detach().cpu().numpy(), ref_loss) <= TOL and rel_err(p.grad.cpu().numpy(), ref_grad) <= TOL
    # no gradient wanted: same loss, nothing saved
    again = tvf.tvf_value_loss(torch.from_numpy(t).cuda(), torch.from_numpy(q).cuda(), weights, 0.5,
                               mask=torch.from_numpy(mask).cuda())
    assert torch.equal(again, loss.detach())


def test_value_loss_dropout_and_errors(golden_norm):
    g = golden_norm
    torch.manual_seed(3)
    p = torch.tensor(g["loss_in_predictions"], device="cuda", requires_grad=True)
    loss = _module(g, "off", dropout=0.5).on_train_value_minibatch(
        {"tvf_value": p}, {"tvf_returns": torch.from_numpy(g["loss_in_targets"]).cuda()},
        required_tvf_heads=None, single_value_head=None)
    assert loss.shape == (p.shape[0],) and torch.isfinite(loss).all() and not torch.equal(
        loss.detach().cpu(), torch.from_numpy(g["loss_all"]))
    with pytest.raises(AssertionError):
        _module(g, "off").on_train_value_minibatch({"tvf_value": p}, {}, required_tvf_heads=None, single_value_head=None)
    with pytest.raises(ValueError):
        tvf.tvf_value_loss(torch.zeros(3, 5).cuda(), torch.zeros(3, 4).cuda(), np.ones(4, np.float32))
    with pytest.raises(IndexError):
        tvf.tvf_value_loss(torch.zeros(3, 5).cuda(), torch.zeros(3, 2).cuda(), np.ones(2, np.float32),
                           required_tvf_heads=[1, 7])


@pytest.mark.parametrize("shape,dtype", [((1000, 128), torch.float32), ((513,), torch.float32), ((300, 7, 3), torch.float32),
                                         ((64, 5), torch.uint8), ((200, 84, 84), torch.uint8), ((100, 3), torch.float64)])
def test_gather_rows(shape, dtype):
    rng = np.random.RandomState(len(shape) + shape[0])
    src = rng.randint(0, 255, size=shape).astype(np.float64)
    buf = torch.from_numpy(src).to(dtype).cuda()
    sample = rng.permutation(shape[0])[: max(shape[0] // 3, 1)]
    sample[0] = -1  # NumPy semantics for negative indices
    got = rollout.gather_rows(buf, sample)
    assert got.dtype == dtype and torch.equal(got.cpu(), buf.cpu()[torch.from_numpy(sample)])
    # a strided row axis (every second row of a bigger buffer)
    view = buf[::2]
    idx = torch.from_numpy(rng.randint(0, view.shape[0], size=17)).cuda()
    assert torch.equal(rollout.gather_rows(view, idx), view[idx])
    with pytest.raises(IndexError):
        rollout.gather_rows(buf, [0, shape[0]])
    assert rollout.gather_rows(buf, np.zeros(0, np.int64)).shape == (0,) + tuple(shape[1:])
