"""CPU oracle (NumPy): the consumers of device-resident targets.  TEST INFRASTRUCTURE ONLY.

* ``tvf_value_loss``: ``TVFRunnerModule.on_train_value_minibatch`` (``rl/tvf.py:32-77``) in float32 with the
  reference's order of operations, plus the analytic gradient of ``mean(loss * loss_scale)`` with respect to the
  predictions (what ``mean_loss.backward()`` leaves in ``.grad``, ``rl/rollout.py:1553-1555``).
* the minibatch selection ``var_value[sample]`` (``rl/rollout.py:2349-2372``) is plain NumPy indexing and needs no
  restatement.

Pinned against the reference's own method run under torch autograd on the CPU (``tests/golden/make_golden.py``,
``loss_*`` arrays of ``tests/golden/normalization_vectors.npz``): the loss to 2 float32 ulp (torch and NumPy sum the
horizon axis in different orders), the gradient to 2 ulp (autograd chains float32 products).
"""
import numpy as np


def tvf_value_loss(targets, predictions, tvf_weights, horizons, coef, required_tvf_heads=None, single_value_head=None,
                   head_weighting="off", max_horizon=None, n_step=None, mask=None, loss_scale=1.0):
    """Returns (loss [B] float32, d mean(loss * loss_scale) / d predictions [B, K] float32)."""
    targets = np.asarray(targets, dtype=np.float32)
    predictions = np.asarray(predictions, dtype=np.float32)
    if required_tvf_heads is not None:
        targets = targets[:, required_tvf_heads]  # :45-46
    diff = targets - predictions
    loss = np.float32(coef * 0.5) * np.square(diff)  # :49
    head_filter = required_tvf_heads if required_tvf_heads is not None else slice(None, None)
    weights = np.asarray(tvf_weights, dtype=np.float32)[head_filter][None, :]  # :52-53
    loss = loss * weights
    factor = weights.astype(np.float64)
    if head_weighting == "h_weighted" and single_value_head is None:  # :56-62
        weighting = np.asarray([1 + ((max_horizon - h) / n_step) for h in horizons], dtype=np.float32)[None, :]
        adjustment = 2 / (np.min(weighting) + np.max(weighting))
        loss = loss * weighting * np.float32(adjustment)
        factor = factor * weighting * float(adjustment)
    if mask is not None:  # :64-69
        loss = loss * np.asarray(mask, dtype=np.float32)
        factor = factor * np.asarray(mask, dtype=np.float64)
    heads = loss.shape[-1]
    per_sample = np.float32(heads ** 0.5) * loss.mean(axis=-1, dtype=np.float32)  # :73
    grad = (loss_scale / loss.shape[0]) * (heads ** 0.5 / heads) * coef * (-diff.astype(np.float64)) * factor
    return per_sample.astype(np.float32), grad.astype(np.float32)
