"""Drop-in for the reference's IGR/sampler.py: the off-surface sampler of the implicit loss (train_Point2Cyl.py:609).
It consumes torch's generator exactly like the reference (randn_like, then rand on the input's device), so a shared
seed gives the same samples; the draws are torch's RNG kernels - plumbing, there is nothing to accelerate here."""
import torch


class NormalPerPoint:
    """reference IGR/sampler.py:19-37."""

    def __init__(self, global_sigma, local_sigma=0.01):
        self.global_sigma = global_sigma
        self.local_sigma = local_sigma

    def get_points(self, pc_input, local_sigma=None):
        batch_size, sample_size, dim = pc_input.shape
        if local_sigma is not None:
            sample_local = pc_input + torch.randn_like(pc_input) * local_sigma.unsqueeze(-1)
        else:
            sample_local = pc_input + torch.randn_like(pc_input) * self.local_sigma
        sample_global = torch.rand(batch_size, sample_size // 8, dim, device=pc_input.device) * (self.global_sigma * 2) \
            - self.global_sigma
        return torch.cat([sample_local, sample_global], dim=1)
