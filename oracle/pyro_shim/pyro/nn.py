"""PyroModule is an nn.Module subclass; the hot path only needs nn.Module behaviour."""
import torch


class PyroModule(torch.nn.Module):
    pass
