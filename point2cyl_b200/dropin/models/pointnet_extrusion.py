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

    # (attribute, npoint, radius, nsample, feature channels in, widths); sa3 groups all points (reference :21-23)
    _LEVELS = (("sa1", 512, 0.2, 64, 0, (64, 64, 128)),
               ("sa2", 128, 0.4, 64, 128, (128, 128, 256)),
               ("sa3", None, None, None, 256, (256, 512, 1024)))
    # (attribute, skip channels, coarse channels, widths) (reference :25-27); fp1's skip input is the raw extra channels
    _UPS = (("fp3", 256, 1024, (256, 256)), ("fp2", 128, 256, (256, 128)), ("fp1", 0, 128, (128, 128, 128)))

    def __init__(self, normal_channel=False, output_sizes=[3]):
        super().__init__()
        extra = 3 if normal_channel else 0
        self.normal_channel, self.dim_pos = normal_channel, 3
        for name, npoint, radius, nsample, c_in, widths in self._LEVELS:
            c_feat = extra if name == "sa1" else c_in
            setattr(self, name, PointNetSetAbstraction(npoint, radius, nsample, self.dim_pos + c_feat, list(widths),
                                                       group_all=npoint is None))
        for name, c_skip, c_coarse, widths in self._UPS:
            c_skip = extra if name == "fp1" else c_skip
            setattr(self, name, PointNetFeaturePropagation(c_coarse + c_skip, list(widths)))
        width = self._UPS[-1][3][-1]
        self.fc1, self.bn1 = nn.Conv1d(width, width, 1), nn.BatchNorm1d(width)
        self.fc2 = nn.ModuleList([nn.Conv1d(width, int(o), 1) for o in output_sizes])

    def forward(self, x, fps_start=None):
        """fps_start: optional ((B,) int64, (B,) int64) first FPS centroids for sa1/sa2; default draws
        them from the CPU generator in the reference's order."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return autograd.backbone_apply(self, x, fps_start)        # backward = csrc/backward.cu kernels
        return pipeline.backbone_forward(self, x, fps_start)
