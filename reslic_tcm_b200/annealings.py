"""Host-side beta schedules for the STanH quantizer.

Boundary row a4 of SURVEY.md §8: the schedules stay scalar Python; the kernels consume ``beta`` and
``stanh.compute_gap`` supplies the soft/hard gap they are driven by.  Mirrors the public surface of
``StanhAnnealings`` (src/annealings/functions.py:7-141: constructor arguments, ``beta``,
``step(gap, epoch, lss, plat=False)``) so that the trainer's hook (src/training/step.py:46-54) works
unchanged.  Reference defects not replicated: the ``dec_epoc`` attribute typo (:88) and the
unreachable "triangle" branch (:135-139).  Not mirrored (host-side schedules the trainer does not instantiate for
the TCM path, src/annealings/functions.py:144-346): ``RandomAnnealings``, ``Annealing_triangle``,
``AugmentBetaOnPlateau``; this class is plain Python, not an ``nn.Module`` (it holds no tensors).
"""
from __future__ import annotations

import math
import random
from typing import Callable, Dict, List, Optional

KINDS = ("linear_stoc", "linear", "gap", "constant", "loss", "AugmentBetaOnPlateau", "gap_stoc")


class StanhAnnealings:
    def __init__(self, iteration: int = 1500, beta: float = 1, factor: float = 50, type: str = "gap",
                 decreasing: bool = False, dec_epoch: int = -1, decreasing_factor: float = 0, threshold: float = 0.02,
                 mode: str = "min", threshold_mode: str = "abs", patience: int = 10, max_beta: float = 1000,
                 starting_epochs: int = -1, rng: Optional[random.Random] = None):
        if type not in KINDS:
            raise AssertionError(f"unknown annealing type {type!r}")
        self.iteration, self.beta, self.factor, self.type = iteration, beta, factor, type
        self.decreasing, self.dec_epoch, self.decreasing_factor = decreasing, dec_epoch, decreasing_factor
        self.threshold, self.mode, self.threshold_mode = threshold, mode, threshold_mode
        self.patience, self.max_beta, self.starting_epochs = patience, max_beta, starting_epochs
        self.gap = 0.0
        self.loss: List[float] = []
        self.num_bad_epochs: Optional[int] = None
        self.best = 1e2
        self.list_epoch: List[int] = []
        self.current_epoch = 0
        self.counter = 0
        self.beta_list = [self.beta]
        self.beta_max = self.beta
        # the stochastic schedules draw from torch's global generator, as the reference does (random.uniform over torch
        # ops in functions.py:103-113), so torch.manual_seed makes beta reproducible; `rng` overrides (tests)
        self._rng = rng
        self._rules: Dict[str, Callable] = {
            "linear": self._linear, "linear_stoc": self._linear_stoc, "gap": self._gap, "gap_stoc": self._gap_stoc,
            "loss": self._loss, "AugmentBetaOnPlateau": self._plateau, "constant": lambda *a: None,
        }

    # ---- bookkeeping kept for API compatibility
    def update_gap(self, gp):
        self.gap = float(gp)

    def update_loss(self, cl):
        self.loss.append(float(cl))

    def update_max_beta(self):
        self.max_beta = 1
        self.beta = 1

    def is_better(self, a, best):
        if self.threshold_mode == "abs":
            return a < best - self.threshold if self.mode == "min" else a > best - self.threshold
        eps = 1.0 - self.threshold
        return a < best * eps if self.mode == "min" else a > best * eps

    def _uniform(self, lo: float, hi: float) -> float:
        if self._rng is not None:
            return self._rng.uniform(lo, hi)
        import torch

        return float(torch.empty(1).uniform_(float(lo), float(max(hi, lo))).item())

    # ---- the schedules
    def _linear(self, gap, epoch, lss, plat):
        if self.beta >= 50000:
            self.beta = self.beta / 2
        elif not self.decreasing or self.dec_epoch > epoch:
            self.beta += self.factor / self.iteration
        else:
            self.beta -= self.decreasing_factor / self.iteration

    def _linear_stoc(self, gap, epoch, lss, plat):
        self.max_beta += self.factor / self.iteration
        self.beta = self._uniform(1, self.beta_max)

    def _gap(self, gap, epoch, lss, plat):
        self.update_gap(gap)
        self.beta = self.beta + self.factor * self.gap

    def _gap_stoc(self, gap, epoch, lss, plat):
        self.update_gap(gap)
        self.beta_max = self.beta_max + self.factor * self.gap
        self.beta = self._uniform(1, min(self.beta_max, self.max_beta))

    def _loss(self, gap, epoch, lss, plat):
        self.update_loss(lss)
        if len(self.loss) >= 2:
            d = math.fabs(self.loss[-1] - self.loss[-2])
            if 0 < d <= self.threshold:
                self.beta = self.beta + self.factor * (1 / d)
            self.loss = self.loss[-2:]

    def _plateau(self, gap, epoch, lss, plat):
        if not plat:
            return
        self.current_epoch = epoch
        current = float(lss)
        if self.num_bad_epochs is None:
            self.num_bad_epochs = 0
        if self.is_better(current, self.best):
            self.best, self.num_bad_epochs = current, 0
        else:
            self.num_bad_epochs += 1
        if self.num_bad_epochs > self.patience and self.beta_list[-1] < self.max_beta:
            self.beta = self.beta * self.factor
            self.num_bad_epochs = 0
            self.beta_list.append(self.beta)
            self.list_epoch.append(epoch)

    def step(self, gap, epoch, lss, plat: bool = False):
        """One scheduler step; ``gap`` may be a float or a 0-d tensor (one host read)."""
        g = float(gap) if gap is not None else 0.0
        self._rules[self.type](g, epoch, lss, plat)
        return self.beta
