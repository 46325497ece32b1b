"""Stand-in for the `adamp` package (not installed here): AdamP's constructor signature on top of AdamW."""
import torch


class AdamP(torch.optim.AdamW):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, delta=0.1, wd_ratio=0.1, nesterov=False):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
