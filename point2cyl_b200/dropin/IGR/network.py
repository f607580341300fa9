"""Drop-in for the reference's IGR/network.py (the implicit sketch network of train_Point2Cyl.py): same names,
constructor signatures and state_dict keys; forward passes run on the libp2c.so kernels (point2cyl_b200.igr).

    implicit_net = ImplicitNet(d_in=D_IN + LATENT_SIZE, dims=[512] * 8, skip_in=[4], geometric_init=True, radius_init=1, beta=100)
    sk_pnts.requires_grad_()                      # train_Point2Cyl.py:617
    sk_pred = implicit_net(sk_pnts)               # forward sweep; keeps softplus'(z) because the input asks for a gradient
    mnfld_grad = gradient(sk_pnts, sk_pred)       # closed-form reverse sweep, last two columns (reference :8-17)

With grad mode on, `sk_pred`, `mnfld_grad` and the latent codes are autograd nodes whose backward runs the closed-form
sweeps of point2cyl_b200.igr (implicit_backward, the encoder's layer backward): `im_loss.backward()` of the unmodified
training loop fills the gradients of ImplicitNet and PointNetEncoder (train_Point2Cyl.py:608-672, :686-690).
"""
import numpy as np
import torch
import torch.nn as nn

from point2cyl_b200 import igr


def gradient(inputs, outputs):
    """reference :8-17."""
    return igr.gradient(inputs, outputs)


class ImplicitNet(nn.Module):
    """reference :20-92."""

    def __init__(self, d_in, dims, skip_in=(), geometric_init=True, radius_init=1, beta=100):
        super().__init__()
        dims = [d_in] + list(dims) + [1]
        self.num_layers = len(dims)
        self.skip_in = skip_in
        for layer in range(0, self.num_layers - 1):
            out_dim = dims[layer + 1] - d_in if layer + 1 in skip_in else dims[layer + 1]
            lin = nn.Linear(dims[layer], out_dim)
            if geometric_init:
                if layer == self.num_layers - 2:
                    torch.nn.init.normal_(lin.weight, mean=np.sqrt(np.pi) / np.sqrt(dims[layer]), std=0.00001)
                    torch.nn.init.constant_(lin.bias, -radius_init)
                else:
                    torch.nn.init.constant_(lin.bias, 0.0)
                    torch.nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))
            setattr(self, "lin" + str(layer), lin)
        self.activation = nn.Softplus(beta=beta) if beta > 0 else nn.ReLU()

    def forward(self, input):
        return igr.implicit_net_forward(self, input)


class PointNetEncoder(nn.Module):
    """reference :132-174."""

    def __init__(self, embedding_size, input_channels=2, with_normals=False):
        super().__init__()
        self.input_channels = input_channels * 2 if with_normals else input_channels
        self.mlp1 = nn.Sequential(nn.Conv1d(self.input_channels, 64, 1), nn.BatchNorm1d(64), nn.ReLU(),
                                  nn.Conv1d(64, 64, 1), nn.BatchNorm1d(64), nn.ReLU())
        self.mlp2 = nn.Sequential(nn.Conv1d(64, 64, 1), nn.BatchNorm1d(64), nn.ReLU(),
                                  nn.Conv1d(64, 128, 1), nn.BatchNorm1d(128), nn.ReLU(),
                                  nn.Conv1d(128, 1024, 1), nn.BatchNorm1d(1024), nn.ReLU())
        self.fc = nn.Linear(1024, embedding_size)

    def forward(self, x):
        return igr.encoder_forward(self, x)


def add_latent(points, latent_codes):
    """reference :200-206."""
    return igr.add_latent(points, latent_codes)
