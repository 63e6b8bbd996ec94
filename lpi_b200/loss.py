"""ClipLoss / nt_bxent_loss with the reference's signatures (retrieval/loss/loss.py:6-33, 38-87) on the loss kernels."""
from __future__ import annotations

import torch
from torch import nn

from .autograd import ClipLossFn, TaskLossFn


def nt_bxent_loss(x, target, temperature=1.0):
    """loss.py:6-33 incl. the double sigmoid (BCE-with-logits of an already sigmoided value, loss.py:21).  The gradient
    flows to the LAST row of x only -- in LPI every other row belongs to a frozen earlier task (slinet.py:176-180)."""
    assert len(x.size()) == 2
    return TaskLossFn.apply(x, target, temperature)


class ClipLoss(nn.Module):
    def __init__(self, local_loss=False, gather_with_grad=False, cache_labels=True, rank=0, world_size=1, use_horovod=False):
        super().__init__()
        self.local_loss = local_loss
        self.gather_with_grad = gather_with_grad
        self.cache_labels = cache_labels
        self.rank = rank
        self.world_size = world_size
        self.use_horovod = use_horovod
        self.prev_num_logits = 0
        self.labels = {}

    def get_ground_truth(self, device, num_logits) -> torch.Tensor:
        """loss.py:62-73 (the kernel builds the diagonal target implicitly; kept for API compatibility)."""
        if self.prev_num_logits != num_logits or device not in self.labels:
            labels = torch.arange(num_logits, device=device, dtype=torch.long)
            if self.world_size > 1 and self.local_loss:
                labels = labels + num_logits * self.rank
            if self.cache_labels:
                self.labels[device] = labels
                self.prev_num_logits = num_logits
        else:
            labels = self.labels[device]
        return labels

    def forward(self, logits):
        """1/2 [CE(logits, arange) + CE(logits^T, arange)], mean reduction (loss.py:75-87)."""
        return ClipLossFn.apply(logits)
