"""Reference-facing mirrors: the classes / functions bayesrul's tasks and configs name on the hot path.

    bayesrul.models.bayesian.BNN            -> compat.BNN
    bayesrul.models.frequentist.HNN         -> compat.HNN
    bayesrul.models.nets.{Inception,Conv,Linear}
    bayesrul.models.guides.radial.AutoRadial
    bayesrul.models.deepens.deep_ensemble (the mixture-moment formula; the generator stays in the reference)
    the parquet writing loop of tasks/predict.py -> compat.predictions
    tyxe.* / pyro.* names used by bayesian.py  -> compat.tyxe_shim / compat.pyro_shim
"""
import sys
import types

from . import pyro_shim, tyxe_shim
from .bayesian import BNN, param_store_to, remove_dict_entry_startswith
from .deepens import deep_ensemble, mixture_moments
from .predictions import predictions_to_frame, write_predictions
from .frequentist import HNN
from .metrics import rms_calibration_error, sharpness
from .nets import Conv, Inception, Linear, enable_dropout, weights_init
from .radial import AutoRadial, Radial


def install_shims(force: bool = False) -> None:
    """Register the shims as `pyro` / `tyxe` modules when the real packages are not importable, so that
    code written against `import pyro, tyxe` (bayesrul/models/bayesian.py:5-13) keeps running."""
    for name, shim in (("pyro", pyro_shim), ("tyxe", tyxe_shim)):
        if not force:
            try:
                __import__(name)
                continue
            except Exception:  # noqa: BLE001
                pass
        mod = types.ModuleType(name)
        for k in dir(shim):
            if not k.startswith("__"):
                setattr(mod, k, getattr(shim, k))
        sys.modules[name] = mod
    pyro = sys.modules["pyro"]
    if not hasattr(pyro, "infer"):
        infer = types.ModuleType("pyro.infer")
        infer.SVI, infer.Trace_ELBO, infer.TraceMeanField_ELBO = pyro_shim.SVI, pyro_shim.Trace_ELBO, pyro_shim.TraceMeanField_ELBO
        optim = types.ModuleType("pyro.optim")
        optim.ClippedAdam = pyro_shim.ClippedAdam
        dist = types.ModuleType("pyro.distributions")
        dist.Normal = tyxe_shim.Normal
        pyro.infer, pyro.optim, pyro.distributions = infer, optim, dist
        sys.modules.update({"pyro.infer": infer, "pyro.optim": optim, "pyro.distributions": dist})


__all__ = ["BNN", "HNN", "Inception", "Conv", "Linear", "AutoRadial", "Radial", "deep_ensemble", "mixture_moments", "predictions_to_frame", "write_predictions",
           "weights_init", "enable_dropout", "rms_calibration_error", "sharpness", "install_shims", "pyro_shim", "tyxe_shim",
           "param_store_to", "remove_dict_entry_startswith"]
