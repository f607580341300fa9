"""Worst relative error of every gradient of the with-sketch training step (im_loss.backward() on the kernels) against
the reference's gradient goldens tests/golden/igr_*.npz - the numbers the 1e-4 bar of tests/test_gpu_igr.py
test_sketch_training_step_golden rests on.  GPU box: python tests/tools/igr_grad_err.py"""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import test_gpu_igr as T
from point2cyl_b200 import igr
for name in T.CASES:
    g = np.load(os.path.join(os.path.join(ROOT, 'tests', 'golden'), name))
    B, K, S, seed, is_l2 = (int(v) for v in g["meta"])
    net, enc, enc_gt = T.nets(seed)
    sk = torch.from_numpy(g["gt_sketches"]).reshape(B * K, S, 4).cuda()
    mask_gt = torch.from_numpy(g["mask_gt"]).cuda()
    lat = enc(torch.from_numpy(g["global_pc"]).cuda()); lat.retain_grad()
    sp, sn = sk[:, :, :2].contiguous(), sk[:, :, 2:].contiguous()
    lat_gt = enc_gt(torch.cat((sp, sn), dim=-1))
    off = torch.from_numpy(g["nonmnfld_pnts"]).reshape(B * K, S + S // 8, 2).cuda()
    out = igr.sketch_loss_block(net, lat, lat_gt, sp, sn, off, mask_gt, bool(is_l2))
    out["im_loss"].backward()
    print(name, 'd_latent', T.rel_l2(lat.grad, g["d_latent"]))
    for prefix, mod in (("net", net), ("enc", enc), ("encgt", enc_gt)):
        top = max(float(g[f"gradnorm_{prefix}.{n}"]) for n, _ in mod.named_parameters())
        worst = (0, '')
        for pname, p in mod.named_parameters():
            want, nrm = g[f"grad_{prefix}.{pname}"], float(g[f"gradnorm_{prefix}.{pname}"])
            if nrm <= 1e-5 * top: continue
            got = p.grad.detach().reshape(-1)
            e_n = abs(float(got.double().norm()) - nrm) / nrm
            e_s = T.rel_l2(got[:want.size], want) if float(np.linalg.norm(want)) > 1e-3 * nrm else 0.0
            if max(e_n, e_s) > worst[0]: worst = (max(e_n, e_s), pname)
        print('  ', prefix, worst)
