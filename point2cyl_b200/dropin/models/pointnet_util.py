"""Drop-in for the reference's models/pointnet_util.py: same names, signatures, state_dict keys.

Every function and module here runs on the libp2c.so CUDA kernels (point2cyl_b200.ops /
.pipeline); there is no torch-eager or CPU fallback.  Shapes and semantics are the reference's:
function level is point-major (B,N,C); nn.Module level is channel-first (B,C,N).
"""
import torch
import torch.nn as nn

from point2cyl_b200 import autograd, ops, pipeline


def square_distance(src, dst):
    """(B,S,C=3),(B,N,3) -> (B,S,N); reference models/pointnet_util.py:19-40."""
    return ops.square_distance(src, dst)


def index_points(points, idx):
    """points (B,N,C), idx (B,S) or (B,S,ns) -> (B,S[,ns],C); reference :43-60."""
    return ops.gather_rows(points, idx)


def farthest_point_sample(xyz, npoint):
    """xyz (B,N,3) -> (B,npoint) int64; reference :63-84.  The first centroid is drawn from the CPU
    generator exactly like the reference (:75), so a shared torch.manual_seed gives equal samples."""
    B, N, _ = xyz.shape
    start = pipeline.draw_fps_start(B, N, xyz.device)
    return ops.fps(xyz, npoint, start)[0]


def query_ball_point(radius, nsample, xyz, new_xyz):
    """-> (B,S,nsample) int64; reference :87-107."""
    return ops.ball_query(radius, nsample, xyz, new_xyz)


def sample_and_group(npoint, radius, nsample, xyz, points, returnfps=False):
    """reference :110-143 -> new_xyz (B,S,3), new_points (B,S,nsample,3+D)."""
    B, N, _ = xyz.shape
    start = pipeline.draw_fps_start(B, N, xyz.device)
    fps_idx, new_xyz = ops.fps(xyz, npoint, start)
    idx = ops.ball_query(radius, nsample, xyz, new_xyz)
    D = 0 if points is None else points.shape[-1]
    feats = None if points is None else points.contiguous().float().reshape(B * N, D)
    rows = ops.group(xyz, feats, new_xyz, idx, ldo=3 + D)
    new_points = rows.reshape(B, npoint, nsample, 3 + D)
    if returnfps:
        grouped_xyz = ops.gather_rows(xyz.contiguous().float(), idx)
        return new_xyz, new_points, grouped_xyz, fps_idx
    return new_xyz, new_points


def sample_and_group_all(xyz, points):
    """reference :146-163 -> new_xyz (B,1,3) zeros, new_points (B,1,N,3+D)."""
    B, N, _ = xyz.shape
    D = 0 if points is None else points.shape[-1]
    feats = None if points is None else points.contiguous().float().reshape(B * N, D)
    rows = ops.group(xyz, feats, None, None, ldo=3 + D)
    return torch.zeros(B, 1, 3, dtype=torch.float32, device=xyz.device), rows.reshape(B, 1, N, 3 + D)


def _to_rows(cf):
    """(B,C,N) channel-first -> (B*N, C) rows (free when cf is a permuted view of row storage)."""
    if cf is None:
        return None
    B, C, N = cf.shape
    return cf.permute(0, 2, 1).contiguous().float().reshape(B * N, C)


def _to_cf(rows, B):
    return rows.reshape(B, -1, rows.shape[1]).permute(0, 2, 1)


class PointNetSetAbstraction(nn.Module):
    """reference :166-207.  forward(xyz (B,3,N), points (B,D,N)|None) -> ((B,3,S), (B,D',S))."""

    def __init__(self, npoint, radius, nsample, in_channel, mlp, group_all):
        super().__init__()
        self.npoint, self.radius, self.nsample, self.group_all = npoint, radius, nsample, group_all
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv2d(last, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm2d(out_channel))
            last = out_channel

    def forward(self, xyz, points):
        B = xyz.shape[0]
        xyz_pm = xyz.permute(0, 2, 1).contiguous().float()
        start = None if self.group_all else pipeline.draw_fps_start(B, xyz_pm.shape[1], xyz.device)
        rows = _to_rows(points)
        if torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters())
                                        or (rows is not None and rows.requires_grad)):
            # backward = csrc/backward.cu kernels (gradients w.r.t. the features and the parameters)
            new_xyz, feats = autograd.SetAbstractionFn.apply(self, xyz_pm, rows, start, *list(self.parameters()))
        else:
            new_xyz, feats = pipeline.set_abstraction(self, xyz_pm, rows, start)
        return new_xyz.permute(0, 2, 1), _to_cf(feats, B)


class PointNetSetAbstractionMsg(nn.Module):
    """reference :210-267 (multi-scale grouping; defined upstream but never instantiated there)."""

    def __init__(self, npoint, radius_list, nsample_list, in_channel, mlp_list):
        super().__init__()
        self.npoint, self.radius_list, self.nsample_list = npoint, radius_list, nsample_list
        self.conv_blocks = nn.ModuleList()
        self.bn_blocks = nn.ModuleList()
        for mlp in mlp_list:
            convs, bns = nn.ModuleList(), nn.ModuleList()
            last = in_channel + 3
            for out_channel in mlp:
                convs.append(nn.Conv2d(last, out_channel, 1))
                bns.append(nn.BatchNorm2d(out_channel))
                last = out_channel
            self.conv_blocks.append(convs)
            self.bn_blocks.append(bns)

    def forward(self, xyz, points):
        B = xyz.shape[0]
        xyz_pm = xyz.permute(0, 2, 1).contiguous().float()
        N = xyz_pm.shape[1]
        feats = _to_rows(points)
        D = 0 if feats is None else feats.shape[1]
        start = pipeline.draw_fps_start(B, N, xyz.device)
        if torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters())
                                        or (feats is not None and feats.requires_grad)):
            new_xyz, out = autograd.SetAbstractionMsgFn.apply(self, xyz_pm, feats, start, *list(self.parameters()))
            return new_xyz.permute(0, 2, 1), _to_cf(out, B)
        _, new_xyz = ops.fps(xyz_pm, self.npoint, start)
        outs = []
        for i, radius in enumerate(self.radius_list):
            idx = ops.ball_query(radius, self.nsample_list[i], xyz_pm, new_xyz)
            # the Msg variant concatenates [features, centred xyz] (:246-249): gather, then swap blocks
            rows = ops.group(xyz_pm, feats, new_xyz, idx, ldo=3 + D)
            if D:
                rows = torch.cat([rows[:, 3:], rows[:, :3]], dim=1).contiguous()
            outs.append(pipeline.mlp_stack(rows, 3 + D, self.conv_blocks[i], self.bn_blocks[i],
                                           self.training, pool_group=self.nsample_list[i]))
        return new_xyz.permute(0, 2, 1), _to_cf(torch.cat(outs, dim=1), B)


class PointNetFeaturePropagation(nn.Module):
    """reference :270-320.  forward(xyz1 (B,3,N), xyz2 (B,3,S), points1 (B,D1,N)|None, points2 (B,D2,S))
    -> (B,D',N)."""

    def __init__(self, in_channel, mlp):
        super().__init__()
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv1d(last, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm1d(out_channel))
            last = out_channel

    def forward(self, xyz1, xyz2, points1, points2):
        B = xyz1.shape[0]
        x1 = xyz1.permute(0, 2, 1).contiguous().float()
        x2 = xyz2.permute(0, 2, 1).contiguous().float()
        f1, f2 = _to_rows(points1), _to_rows(points2)
        if torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or f2.requires_grad
                                        or (f1 is not None and f1.requires_grad)):
            out = autograd.FeaturePropagationFn.apply(self, x1, x2, f1, f2, *list(self.parameters()))
        else:
            out = pipeline.feature_propagation(self, x1, x2, f1, f2)
        return _to_cf(out, B)
