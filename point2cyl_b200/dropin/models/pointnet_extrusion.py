"""Drop-in for the reference's models/pointnet_extrusion.py: `backbone` with the same constructor,
forward signature and state_dict keys (123 entries for output_sizes=[3,16]), running on libp2c.so."""
import torch
import torch.nn as nn

from point2cyl_b200 import autograd, pipeline
from point2cyl_b200.dropin.models.pointnet_util import (PointNetFeaturePropagation,  # noqa: F401
                                                        PointNetSetAbstraction,
                                                        PointNetSetAbstractionMsg)


class backbone(nn.Module):
    """reference models/pointnet_extrusion.py:8-66.  forward(x (B,N,3[+3])) -> [ (B,N,o_i) ]."""

    def __init__(self, normal_channel=False, output_sizes=[3]):
        super().__init__()
        additional_channel = 3 if normal_channel else 0
        self.normal_channel = normal_channel
        self.dim_pos = 3
        self.sa1 = PointNetSetAbstraction(npoint=512, radius=0.2, nsample=64,
                                          in_channel=3 + additional_channel, mlp=[64, 64, 128], group_all=False)
        self.sa2 = PointNetSetAbstraction(npoint=128, radius=0.4, nsample=64, in_channel=128 + 3,
                                          mlp=[128, 128, 256], group_all=False)
        self.sa3 = PointNetSetAbstraction(npoint=None, radius=None, nsample=None, in_channel=256 + 3,
                                          mlp=[256, 512, 1024], group_all=True)
        self.fp3 = PointNetFeaturePropagation(in_channel=1024 + 256, mlp=[256, 256])
        self.fp2 = PointNetFeaturePropagation(in_channel=256 + 128, mlp=[256, 128])
        self.fp1 = PointNetFeaturePropagation(in_channel=128 + additional_channel, mlp=[128, 128, 128])
        self.fc1 = nn.Conv1d(128, 128, 1)
        self.bn1 = nn.BatchNorm1d(128)
        self.fc2 = nn.ModuleList(nn.Conv1d(128, o, 1) for o in output_sizes)

    def forward(self, x, fps_start=None):
        """fps_start: optional ((B,) int64, (B,) int64) first FPS centroids for sa1/sa2; default draws
        them from the CPU generator in the reference's order."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return autograd.backbone_apply(self, x, fps_start)        # backward = csrc/backward.cu kernels
        return pipeline.backbone_forward(self, x, fps_start)
