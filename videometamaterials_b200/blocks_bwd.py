"""Backward passes of the Unet3D blocks and the training loss (autograd.Functions over the C ABI)."""
from __future__ import annotations

import torch


def training_loss(model, x0, noise, qcoef, t, cond, null_mask, l2=False):
    raise NotImplementedError("training path under construction")
