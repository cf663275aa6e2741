"""SDF loss driving the backward of the hot path (reference network/losses.py:6-38); stays PyTorch."""
import torch
import torch.nn as nn


class SDFLoss(nn.Module):
    def __init__(self, sdf_scale):
        super().__init__()
        self.sdf_scale = sdf_scale

    def forward(self, outputs, targets):
        sdf_loss = torch.mean(((targets * self.sdf_scale - outputs) ** 2).sum(-1))
        real = torch.mean((targets - outputs / self.sdf_scale) ** 2) * 10000
        acc = torch.mean(torch.eq(torch.gt(targets, 0.5), torch.gt(outputs, 0.5)).float())
        return {"sdf_loss": sdf_loss, "ignore_sdf_loss_realvalue": real, "ignore_sdf_accuracy": acc}
