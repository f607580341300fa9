"""The reference arm: the UPSTREAM reference's own code on the forward+loss path, timed (bench / test infrastructure).

`ReferenceStep` is the loop body of the reference's training script run verbatim:
  * `model` is models/pointnet_extrusion.py `backbone` (stock nn.Module, stock PointNet++ utilities),
  * the statements from `X, W_raw = model(pcs)` to `total_loss += total_center_loss`
    (train_Point2Cyl_without_sketch.py:244-353) are READ FROM THE REFERENCE FILE, dedented and exec'd with the script's
    globals (all five --pred_* branches on, multipliers 1) - losses.compute_all_losses, the inline base/barrel loss,
    data_utils.estimate_extrusion_axis (dense (B,N,N) diag_embed route) / estimate_extrusion_centers are the reference's.
The files come from baseline/_ref/ (staged by baseline/make_ref.py, git-ignored, travels to the GPU box) or, in the
authoring container, /root/reference; oracle/ref_shim.py supplies the import stubs.  Nothing of point2cyl_b200's
kernels or engine is on this path; `module_bindings` lets tests/test_gpu_dropin_literal.py run the SAME lines with the
drop-in modules bound instead (the literal drop-in proof).
"""
from __future__ import annotations

import os
import sys
import textwrap
import time
from typing import Dict, Optional

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from baseline import make_ref  # noqa: E402

FIRST_STMT = "X, W_raw = model(pcs)"
LAST_STMT = "total_loss += total_center_loss"


def reference_root() -> Optional[str]:
    return make_ref.root()


def loop_body_source(root: str) -> str:
    """The forward+loss statements of the training loop, as the reference wrote them."""
    path = os.path.join(root, "train_Point2Cyl_without_sketch.py")
    lines = open(path).read().split("\n")
    i0 = next(i for i, l in enumerate(lines) if l.strip() == FIRST_STMT)
    i1 = next(i for i, l in enumerate(lines) if l.strip().startswith(LAST_STMT) and i > i0)
    block = [l.replace("\t", "    ") for l in lines[i0:i1 + 1]]
    src = textwrap.dedent("\n".join(l if l.strip() else "" for l in block))
    assert "compute_all_losses(" in src and "estimate_extrusion_axis(" in src and "cross_entropy" in src
    return src


def load_modules(root: str):
    """The reference's modules through the import shim (stubs for unused imports, torch.symeig)."""
    from oracle import ref_shim
    ref_shim.REF_ROOT = root
    return ref_shim.load()


def script_globals(bindings: Dict[str, object], K: int, N: int) -> Dict[str, object]:
    """What the loop body reads from the script's module scope (train_Point2Cyl_without_sketch.py:24-120)."""
    g = dict(torch=torch, F=F, K=K, NUM_POINT=N, PRED_NORMAL=True, PRED_SEG=True, PRED_BB=True, PRED_EXT=True,
             PRED_CENTER=True, NORM_EIG=False, normal_loss_multiplier=1.0, miou_loss_multiplier=1.0,
             bb_loss_multiplier=1.0, extrusion_loss_multiplier=1.0, center_loss_multiplier=1.0)
    g.update(bindings)
    return g


def reference_bindings(ref) -> Dict[str, object]:
    return dict(compute_all_losses=ref.losses.compute_all_losses, get_mask_gt=ref.losses.get_mask_gt,
                compute_normal_loss=ref.losses.compute_normal_loss,
                reduce_mean_masked_instance=ref.losses.reduce_mean_masked_instance,
                estimate_extrusion_axis=ref.data_utils.estimate_extrusion_axis,
                estimate_extrusion_centers=ref.data_utils.estimate_extrusion_centers)


class LoopBody:
    """The exec'd loop body bound to a model and a set of loss functions."""

    def __init__(self, root: str, model, bindings: Dict[str, object], K: int, N: int):
        self.code = compile(loop_body_source(root), os.path.join(root, "train_Point2Cyl_without_sketch.py"), "exec")
        self.globals = script_globals(bindings, K, N)
        self.globals["model"] = model

    def __call__(self, batch: Dict[str, torch.Tensor]) -> Dict[str, object]:
        ns = dict(self.globals)
        ns.update(pcs=batch["pcs"], sampled_pcs=batch["pcs"], batch_size=batch["pcs"].shape[0],
                  gt_normals=batch["normals"], gt_extrusion_instances=batch["inst"], gt_bb_labels=batch["bb"],
                  gt_extrusion_axes=batch["axes"], gt_extrusion_centers=batch["centers"])
        exec(self.code, ns)
        return ns


class ReferenceStep:
    """forward+loss of the unmodified reference on `device` (CPU: the reference arm; CUDA: the eager-GPU bar)."""

    def __init__(self, K: int, N: int, device="cpu", seed: int = 0, root: Optional[str] = None):
        self.root = root or reference_root()
        if self.root is None:
            raise FileNotFoundError("no reference files: run `python baseline/make_ref.py` where /root/reference exists")
        self.ref = load_modules(self.root)
        torch.manual_seed(seed)
        self.model = self.ref.net.backbone(output_sizes=[3, 2 * K]).to(device).train()
        self.body = LoopBody(self.root, self.model, reference_bindings(self.ref), K, N)
        self.device = device

    def __call__(self, batch, grad: bool = False):
        with torch.set_grad_enabled(grad):
            ns = self.body(batch)
        return ns["total_loss"]


def time_reference(B: int, N: int, K: int, steps: int, warmup: int, device="cpu", grad: bool = False,
                   seed: int = 1234) -> Dict[str, object]:
    """seconds per step (mean over `steps` after `warmup`) of the reference's forward+loss on B synthetic clouds."""
    from point2cyl_b200 import synthetic
    step = ReferenceStep(K, N, device)
    batch = {k: v.to(device) for k, v in synthetic.s_cyl(B, N, K, seed=seed).items()}
    times = []
    for i in range(warmup + steps):
        if str(device).startswith("cuda"):
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss = step(batch, grad=grad)
        loss.item()                                   # the script logs total_loss.item() every step (:373)
        if str(device).startswith("cuda"):
            torch.cuda.synchronize()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"sec_per_step": sec, "clouds_per_s": B / sec, "B": B, "steps": len(times), "grad": grad,
            "root": os.path.relpath(step.root, ROOT) if step.root.startswith(ROOT) else step.root}
