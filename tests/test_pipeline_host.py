"""CPU: the packaged entropy passes mirror the reference models' module layout (state_dict names a reference
checkpoint carries) and refuse to run without the GPU — there is no CPU fallback on the product path."""
import pytest
import torch

from reslic_tcm_b200 import _cabi
from reslic_tcm_b200.pipeline import TcmEntropyPath, TcmStanhEntropyPath

CFG = dict(beta=10.0, num_sigmoids=0, extrema=20, trainable=True, removing_mean=True, symmetry=False)


def test_module_layout_matches_the_reference_models():
    # tcm.py:416-417: one bottleneck, one Gaussian conditional
    plain = TcmEntropyPath()
    keys = set(plain.state_dict())
    assert {f"entropy_bottleneck._matrix{i}" for i in range(5)} <= keys and "entropy_bottleneck.quantiles" in keys
    assert any(k.startswith("gaussian_conditional.") for k in keys)
    # tcm_stanh.py:331-336: a ModuleList of GaussianConditionalStanh, one per lambda, each owning its STanH (w, b)
    st = TcmStanhEntropyPath(CFG, z_channels=8, channels=8, levels=3)
    keys = set(st.state_dict())
    for lv in range(3):
        assert f"gaussian_conditional.{lv}.stanh.w" in keys and f"gaussian_conditional.{lv}.stanh.b" in keys
    assert {f"entropy_bottleneck._bias{i}" for i in range(5)} <= keys
    assert len(st.gaussian_conditional) == 3 and st.num_slices == 5


def test_passes_fail_loudly_on_cpu_tensors():
    y = torch.zeros(1, 20, 4, 4)
    z = torch.zeros(1, 8, 1, 1)
    st = TcmStanhEntropyPath(CFG, z_channels=8, channels=8)
    with pytest.raises((_cabi.ReslicError, ValueError, RuntimeError, AssertionError)):
        st(y, y, y + 1, z)
    plain = TcmEntropyPath(z_channels=8)
    with pytest.raises((_cabi.ReslicError, ValueError, RuntimeError, AssertionError)):
        plain(y, y, y + 1, z)


def test_stanh_pass_rejects_channel_counts_that_do_not_slice():
    st = TcmStanhEntropyPath(CFG, z_channels=8, channels=8)
    assert st.num_slices == 5
    with pytest.raises(Exception):
        st(torch.zeros(1, 7, 4, 4), torch.zeros(1, 7, 4, 4), torch.ones(1, 7, 4, 4), torch.zeros(1, 8, 1, 1))


def test_stanh_pass_rejects_options_it_would_otherwise_ignore():
    st = TcmStanhEntropyPath(CFG, z_channels=8, channels=8)
    y, z = torch.zeros(1, 20, 4, 4), torch.zeros(1, 8, 1, 1)
    with pytest.raises(ValueError, match="publish"):
        st(y, y, y + 1, z, exchange=object())
    with pytest.raises(ValueError, match="symbols"):
        st(y, y, y + 1, z, with_indexes=True)
