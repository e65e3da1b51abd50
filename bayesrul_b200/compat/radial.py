"""Radial guide (reference: bayesrul/models/guides/radial.py:31-41 RadialNormal.rsample, :44-144 AutoRadial).
w = loc + scale * (eps / ||eps||_2) * r with ||.|| over the whole site tensor and r ~ N(0,1) one scalar per
site per draw; the sampler is the CUDA kernel pair radial_norm / radial_apply (csrc/brl_kernels.cu)."""
from __future__ import annotations

from .tyxe_shim import AutoNormal


class AutoRadial(AutoNormal):
    family = "radial"

    def __init__(self, module, init_loc_fn=None, init_scale=1e-1, train_loc=True, train_scale=True,
                 max_guide_scale=None, prior=None):
        if not (train_loc and train_scale) or max_guide_scale is not None:
            raise NotImplementedError("bayesrul uses the default trainable loc/scale (bayesian.py:78-91)")
        super().__init__(module, init_scale=init_scale, init_loc_fn=init_loc_fn, prior=prior)


class Radial:
    def __init__(self):
        self.name = "Radial"

    def guide(self):
        return AutoRadial
