"""Stub of the `lpips` package: the reference instantiates lpips.LPIPS(net="vgg")
at import time (source/training/core/base_losses.py:139); it is never evaluated on
the hot path.  Golden generation only."""
import torch


class LPIPS(torch.nn.Module):
    def __init__(self, net="vgg", **kw):
        super().__init__()

    def forward(self, a, b, **kw):
        return torch.zeros(())
