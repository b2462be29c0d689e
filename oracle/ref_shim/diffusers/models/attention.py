import torch


class AttentionBlock(torch.nn.Module):  # only reachable through UNetMidBlock2D, never built for SD-1.5
    def __init__(self, *a, **k):
        raise NotImplementedError
