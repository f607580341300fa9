"""How much head-room do the end-to-end gradient bars have?  Worst relative-L2 / median error per run (3 runs each)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import p2c_oracle as orc
from point2cyl_b200 import pipeline, synthetic
from point2cyl_b200.dropin.models.pointnet_extrusion import backbone
from tests.test_oracle_golden import grad_errors, live_keys
for name, training in (("train_bneval_b2_n1024_k4.npz", False), ("train_b2_n1024_k4.npz", True)):
    g = np.load("tests/golden/" + name)
    B, N, K, seed = (int(v) for v in g["meta"])
    data = {k: v.cuda() for k, v in synthetic.s_cyl(B, N, K, seed).items()}
    mask = ((torch.rand(B, 128, N, generator=torch.Generator().manual_seed(seed + 3)) > 0.5).float() * 2.0).cuda()
    pipeline.dropout_mask_fn = lambda x, p=0.5, **kw: mask
    starts = (torch.from_numpy(g["s1"]).cuda(), torch.from_numpy(g["s2"]).cuda())
    for rep in range(3):
        net = backbone(output_sizes=[3, 2 * K]); net.load_state_dict(orc.init_state_dict((3, 2 * K), seed=seed)); net = net.cuda().train(training)
        X_raw, W_raw = net(data["pcs"], fps_start=starts)
        out = pipeline.loss_forward(data["pcs"], X_raw, W_raw, data["normals"], data["inst"], data["bb"], data["axes"], data["centers"])
        out["total"].backward()
        named = dict(net.named_parameters())
        errs = [(k,) + grad_errors(named[k].grad, g, k) for k in live_keys(g, training)]
        wl2 = max(errs, key=lambda e: e[1]); wmed = max(errs, key=lambda e: e[2]); wn = max(errs, key=lambda e: e[3])
        print(name, "rep", rep, "worst l2 %.2e (%s)  worst median %.2e (%s)  worst norm err %.2e (%s)" % (wl2[1], wl2[0], wmed[2], wmed[0], wn[3], wn[0]),
              "heads l2 %.2e" % max(e[1] for e in errs if e[0].startswith("fc2")))
