"""Minimal `pytorch_lightning` stand-in (TEST INFRASTRUCTURE): just enough of LightningModule for the reference's own
train scripts (creste/train_pefree.py, train_ssc.py, train_traversability.py) to be imported and for their
`training_step` / `validation_step` / `configure_optimizers` to be DRIVEN by a test without a Trainer.

  LightningModule   nn.Module + save_hyperparameters / log / log_dict (recorded in `.logged`), `optimizers()` (the first
                    optimiser of configure_optimizers(), built once), `manual_backward`, `automatic_optimization`,
                    `current_epoch`, `device`, `loggers` (mocks)
  Trainer, callbacks, loggers, strategies   MagicMock (never exercised)
"""
import sys
import types
from unittest import mock

import torch
from torch import nn


class LightningModule(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self.automatic_optimization = True
        self.current_epoch = 0
        self.logged = {}
        self._shim_opt = None
        self.loggers = [mock.MagicMock()]

    def save_hyperparameters(self, *a, **k):
        return None

    def log(self, name, value, *a, **k):
        self.logged[name] = value

    def log_dict(self, d, *a, **k):
        self.logged.update(d)

    def optimizers(self):
        if self._shim_opt is None:
            opts = self.configure_optimizers()
            opts = opts[0] if isinstance(opts, (tuple, list)) else opts
            self._shim_opt = opts[0] if isinstance(opts, (tuple, list)) else opts
        return self._shim_opt

    def manual_backward(self, loss, *a, **k):
        loss.backward(*a, **k)

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")


def install():
    pl = types.ModuleType("pytorch_lightning")
    pl.LightningModule = LightningModule
    pl.LightningDataModule = type("LightningDataModule", (), {"__init__": lambda self, *a, **k: None})
    pl.Trainer = mock.MagicMock(name="Trainer")
    pl.seed_everything = lambda *a, **k: None
    pl.__path__ = []
    sys.modules["pytorch_lightning"] = pl
    for sub in ("loggers", "strategies", "callbacks", "utilities", "utilities.rank_zero"):
        m = mock.MagicMock(name="pytorch_lightning." + sub)
        m.__path__ = []
        sys.modules["pytorch_lightning." + sub] = m
        if "." not in sub:
            setattr(pl, sub, m)
    return pl
