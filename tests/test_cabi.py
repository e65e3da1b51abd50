"""CPU-side checks of the C ABI: the library loads and exports every symbol include/*.h declares
(no compute calls: there is no GPU in the dev container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from bayesrul_b200.build import build
    build()
    from bayesrul_b200 import _lib
    return _lib.load()


def _declared():
    src = open(os.path.join(ROOT, "include", "bayesrul_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(brl_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from bayesrul_b200 import _lib
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.SIGNATURES) == names


def test_static_net_description_matches_oracle(lib):
    from bayesrul_b200 import net_info
    from oracle import bnn_oracle as O
    for net in ("inception", "conv", "linear"):
        info = net_info(net)
        assert info["P"] == O.num_params(net)
        assert [(o, tuple(s)) for o, s in info["sites"]] == [(off, tuple(shape)) for _, shape, off in O.site_table(net)]
        assert info["flops_fwd"] == O.F_FWD[net]
        for i, (co, ci, oe, df) in enumerate(info["layers"]):
            assert co == O.net_layers(net)[i].cout and ci == O.layer_in_channels(net, i)
            assert oe == int(__import__("numpy").prod(O.layer_out_shape(net, i)))
            assert abs(df - O.DROPOUT_SITES[net].get(i, 0.0)) < 1e-7


def test_errors_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from bayesrul_b200 import Engine
    with pytest.raises(RuntimeError):
        Engine("inception")
    assert lib.brl_net_num_params(7) < 0
    ctx = ctypes.c_void_p()
    assert lib.brl_create(ctypes.byref(ctx), 0, 0) != 0  # no device -> CUDA error code, not a crash
    assert b"bayesrul_b200" in lib.brl_last_error()
