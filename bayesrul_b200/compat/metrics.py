"""Online metrics logged every step (reference: bayesrul/results/metrics.py:210-297), as free functions with the
reference's signatures (torch on whatever device the tensors live on): the function-level mirror and the comparison
target of the tests.  The wrappers (compat.BNN / compat.HNN) do NOT call these per step: they log from
Engine.step_metrics (brl_step_metrics: nll, mse, sharpness, rmsce, mace from one fused pass + a 100-bin histogram)."""
import torch
from torch import Tensor


def sharpness(sigma_hat: Tensor) -> Tensor:
    sharp = torch.sqrt(torch.square(sigma_hat).mean())
    assert sharp.numel() == 1, f"Sharpness calculated is of shape {sharp.shape}"
    return sharp


def get_proportion_lists(y_pred: Tensor, y_std: Tensor, y_true: Tensor, num_bins: int, prop_type: str = "interval"):
    exp_p = torch.linspace(0, 1, num_bins, device=y_true.device)
    z = ((y_pred - y_true).flatten() / y_std.flatten()).reshape(-1, 1)
    if prop_type == "interval":
        # |z| <= icdf(0.5 + p/2)  <=>  erf(|z| / sqrt 2) <= p
        q = torch.erf(z.abs() * 0.7071067811865476)
        obs = (q <= exp_p).sum(0).flatten() / z.shape[0]
    elif prop_type == "quantile":
        q = 0.5 * (1 + torch.erf(z * 0.7071067811865476))
        obs = (q <= exp_p).sum(0).flatten() / z.shape[0]
    else:
        raise AssertionError(prop_type)
    return exp_p, obs


def rms_calibration_error(y_pred: Tensor, y_std: Tensor, y_true: Tensor, num_bins: int = 100,
                          prop_type: str = "interval") -> Tensor:
    assert y_pred.shape == y_std.shape == y_true.shape
    assert y_std.min() >= 0, "Not all values are positive"
    exp_p, obs_p = get_proportion_lists(y_pred, y_std, y_true, num_bins, prop_type)
    return torch.sqrt(torch.mean(torch.square(exp_p - obs_p)))


def mean_absolute_calibration_error(y_pred, y_std, y_true, num_bins: int = 100, prop_type: str = "interval"):
    assert y_pred.shape == y_std.shape == y_true.shape
    exp_p, obs_p = get_proportion_lists(y_pred, y_std, y_true, num_bins, prop_type)
    return torch.mean(torch.abs(exp_p - obs_p))
