"""CPU: beta schedules (host-side boundary)."""
import random

import pytest

from reslic_tcm_b200.annealings import StanhAnnealings


def test_gap_rule_accumulates_factor_times_gap():
    a = StanhAnnealings(beta=1, factor=25, type="gap")
    a.step(0.02, 0, 0.0)
    a.step(0.01, 0, 0.0)
    assert a.beta == pytest.approx(1 + 25 * 0.02 + 25 * 0.01)


def test_gap_stoc_draws_below_the_running_maximum():
    a = StanhAnnealings(beta=1, factor=25, type="gap_stoc", max_beta=5, rng=random.Random(0))
    for _ in range(200):
        a.step(0.05, 0, 0.0)
        assert 1 <= a.beta <= 5
    assert a.beta_max > 5


def test_linear_and_decreasing():
    a = StanhAnnealings(iteration=10, beta=1, factor=5, type="linear", decreasing=True, dec_epoch=2, decreasing_factor=5)
    a.step(None, 0, 0.0)
    assert a.beta == pytest.approx(1.5)
    a.step(None, 3, 0.0)
    assert a.beta == pytest.approx(1.0)


def test_plateau_multiplies_after_patience():
    a = StanhAnnealings(beta=2, factor=3, type="AugmentBetaOnPlateau", patience=1, threshold=0.0)
    for loss in (1.0, 1.0, 1.0):
        a.step(None, 0, loss, plat=True)
    assert a.beta == 6 and a.beta_list == [2, 6]


def test_constant_and_bad_type():
    a = StanhAnnealings(beta=7, type="constant")
    assert a.step(0.3, 1, 0.2) == 7
    with pytest.raises(AssertionError):
        StanhAnnealings(type="triangle")
