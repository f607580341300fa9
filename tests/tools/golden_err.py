import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import p2c_oracle as orc
from point2cyl_b200 import pipeline, synthetic
from point2cyl_b200.dropin.models.pointnet_extrusion import backbone
pipeline.dropout_mask_fn = lambda x, p=0.5, **kw: x
for name in ("backbone_b2_n1024_k4.npz", "backbone_b1_n1024_k4.npz"):
    g = np.load("tests/golden/" + name)
    B, N, K, seed = (int(v) for v in g["meta"])
    data = synthetic.s_cyl(B, N, K, seed)
    for mode in ("train", "eval"):
        for prec in ("3xtf32", "fp32"):
            net = backbone(output_sizes=[3, 2 * K]); net.load_state_dict(orc.init_state_dict((3, 2 * K), seed=seed)); net = net.cuda().train(mode == "train")
            starts = (torch.from_numpy(g[f"{mode}_s1"]).cuda(), torch.from_numpy(g[f"{mode}_s2"]).cuda())
            tr = {}
            with torch.no_grad():
                X, W = pipeline.backbone_forward(net, data["pcs"].cuda(), starts, precision=prec, trace=tr)
            e = lambda a, b: float((a.cpu().double() - torch.from_numpy(b).double()).abs().max() / np.abs(b).max())
            print(name, mode, prec, "X %.2e W %.2e" % (e(X, g[f"{mode}_X"]), e(W, g[f"{mode}_W"])))
