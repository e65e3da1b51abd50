"""Parameter containers mirroring the reference nets' constructor, attributes and state_dict keys.

Reference: bayesrul/models/nets/inception.py:142-217 (`Inception`), conv.py:14-79 (`Conv`),
linear.py:10-72 (`Linear`).  The arithmetic does not live here: `forward` hands the flat parameter
buffer to the CUDA engine (deterministic / MC-dropout rule).  All per-site parameters are views into
ONE flat fp32 buffer `theta` (the layout the C ABI consumes), so `state_dict()` / `load_state_dict()`
keep the reference's key names (`layers.0.conv1.0.weight`, ...) while kernels see a single pointer.
"""
from __future__ import annotations

from typing import List, Optional

import math
import torch
import torch.nn as nn

from ..engine import Engine, Noise, net_info

_INCEPTION_SITES = ["layers.0.conv1.0", "layers.0.conv3.0", "layers.0.conv5.0", "layers.0.convpool.1",
                    "layers.1.branch1.0", "layers.1.branch2.0", "layers.1.branch2.2", "layers.1.branch3.0",
                    "layers.1.branch3.2", "layers.1.branch4.1", "layers.3", "last"]


def site_prefixes(kind: str, dropout: bool) -> List[str]:
    """Module paths of the parameterised layers, i.e. named_parameters() order of the reference nets
    (Sequential indices shift when Dropout modules are inserted: conv.py:31-45, linear.py:27-42)."""
    if kind == "inception":
        return list(_INCEPTION_SITES)
    if kind == "conv":
        return [f"layers.{i}" for i in ((0, 3, 7) if dropout else (0, 2, 5))] + ["last"]
    if kind == "linear":
        return [f"layers.{i}" for i in ((1, 4, 7, 10) if dropout else (1, 3, 5, 7))] + ["last"]
    raise ValueError(kind)


class _Holder(nn.Module):
    """Empty container used to reproduce dotted parameter paths."""


class TableNet(nn.Module):
    kind = "inception"

    def __init__(self, win_length: int, n_features: int, activation: str = "relu", dropout: float = 0, bias=True,
                 out_size: int = 2):
        super().__init__()
        if activation != "relu":
            raise ValueError("bayesrul_b200 implements the ReLU nets shipped in bayesrul's configs "
                             "(conf/model/inception.yaml:5); got activation=%r" % (activation,))
        if (win_length, n_features) != (30, 18):
            raise ValueError("bayesrul_b200 kernels are specialised for N-CMAPSS windows 30x18")
        if out_size != 2:
            raise ValueError("heteroscedastic head needs out_size=2")
        self.win_length, self.n_features, self.out_size = win_length, n_features, out_size
        self.dropout = dropout
        info = net_info(self.kind)
        self.P = info["P"]
        self._sites = info["sites"]
        self._layers = info["layers"]
        names = []
        for pre in site_prefixes(self.kind, dropout > 0):
            names += [pre + ".weight", pre + ".bias"]
        self.site_names = names
        self.register_buffer("theta", torch.zeros(self.P), persistent=False)
        self._params: List[nn.Parameter] = []
        for name, (off, shape) in zip(names, self._sites):
            mod = self
            parts = name.split(".")
            for part in parts[:-1]:
                if not hasattr(mod, part):
                    mod.add_module(part, _Holder())
                mod = getattr(mod, part)
            p = nn.Parameter(torch.empty(0))
            mod.register_parameter(parts[-1], p)
            self._params.append(p)
        self._rebind()
        self.reset_parameters()
        self._engine: Optional[Engine] = None
        self._calls = 0
        self.mc_seed = 0

    # every parameter is a view into the flat buffer -------------------------------------------
    def _rebind(self):
        for p, (off, shape) in zip(self._params, self._sites):
            n = 1
            for s in shape:
                n *= s
            p.data = self.theta[off: off + n].view(shape)

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        self._rebind()
        self._engine = None
        return self

    def reset_parameters(self):
        """torch defaults of nn.Conv*/nn.Linear (kaiming_uniform(a=sqrt 5) weight, uniform bias)."""
        with torch.no_grad():
            for i, p in enumerate(self._params):
                co, ci, _, _ = self._layers[i // 2]
                fan_in = p.numel() // co if i % 2 == 0 else self._params[i - 1].numel() // co
                bound = 1.0 / math.sqrt(fan_in)
                p.uniform_(-bound, bound)

    def flat(self) -> torch.Tensor:
        return self.theta

    def engine(self) -> Engine:
        if self._engine is None or self._engine.device != self.theta.device:
            self._engine = Engine(self.kind, self.theta.device)
        return self._engine

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[B,30,18] -> [B,2] (loc, scale): softplus + Threshold(1e-9) head (inception.py:211-215).
        Dropout modules follow nn.Module train/eval state (enable_dropout sets `mc_dropout`)."""
        p = float(self.dropout) if (self.training or getattr(self, "mc_dropout", False)) else 0.0
        self._calls += 1
        out = self.engine().forward(x.contiguous(), "det", theta=self.theta, p_dropout=p,
                                    noise=Noise(seed=self.mc_seed, sample0=self._calls))
        return out[0]

    def save(self, path: str) -> None:
        torch.save(self.state_dict(), path)

    def load(self, path: str, map_location=torch.device("cpu")):
        self.load_state_dict(torch.load(path, map_location=map_location))


class Inception(TableNet):
    kind = "inception"

    def __init__(self, win_length, n_features, activation="relu", bias="True", dropout=0, out_size=2):
        assert n_features == 18, "TODO, Generalize Inception model for other than 18 features"
        super().__init__(win_length, n_features, activation, dropout, bias, out_size)


class Conv(TableNet):
    kind = "conv"

    def __init__(self, win_length, n_features, activation="relu", dropout=0, bias=True, out_size=2):
        super().__init__(win_length, n_features, activation, dropout, bias, out_size)


class Linear(TableNet):
    kind = "linear"

    def __init__(self, win_length, n_features, activation="relu", dropout=0, bias=True, out_size=2):
        super().__init__(win_length, n_features, activation, dropout, bias, out_size)


def weights_init(m):
    """utils/miscellaneous.py:53-63 -- xavier_normal_ for conv weights, kaiming_normal_ for linear weights."""
    if isinstance(m, TableNet):
        with torch.no_grad():
            for p in m._params[0::2]:
                if p.dim() > 2:
                    nn.init.xavier_normal_(p)
                else:
                    nn.init.kaiming_normal_(p)


def enable_dropout(model):
    """utils/miscellaneous.py:66-70 -- keep the dropout sites stochastic at eval time."""
    for m in model.modules():
        if isinstance(m, TableNet):
            m.mc_dropout = True


def init_flat_params(kind: str = "inception", seed: int = 12345) -> torch.Tensor:
    """Flat [P] parameter vector of a freshly built net after `weights_init` under `torch.manual_seed(seed)`
    (the pre-training initialisation of tasks/train.py + utils/miscellaneous.py:53-63); synthetic-weight helper."""
    cls = {"inception": Inception, "conv": Conv, "linear": Linear}[kind]
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed)
        net = cls(30, 18)
        weights_init(net)
        return net.flat().detach().clone()
