"""The reference-side binding of INTEGRATION.md, exercised on CPU: with the drop-in directories first on sys.path the
reference's own import statements (train_Point2Cyl_without_sketch.py:14-23,180) resolve to this repo's modules, every
name the scripts use from them exists, and - when the reference checkout is present - each function / constructor has
the reference's parameter names, order and defaults."""
import inspect
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

LOSSES = ["compute_all_losses", "hungarian_matching", "compute_miou_loss", "compute_normal_loss", "get_mask_gt",
          "reduce_mean_masked_instance", "hard_W_encoding", "compute_segmentation_iou", "compute_normal_difference",
          "sequence_mask", "acos_safe"]
DATA_UTILS = ["estimate_extrusion_axis", "estimate_extrusion_centers", "add_noise", "sketch_implicit_projection",
              "sketch_implicit_projection2", "sketch_implicit_projection3", "get_extrusion_extents"]
UTIL = ["PointNetSetAbstraction", "PointNetSetAbstractionMsg", "PointNetFeaturePropagation", "farthest_point_sample",
        "query_ball_point", "index_points", "square_distance", "sample_and_group", "sample_and_group_all"]


def test_reference_import_statements_resolve_to_the_dropin():
    code = textwrap.dedent(f"""
        import sys, importlib
        sys.path.insert(0, {ROOT!r})
        sys.path.insert(0, {os.path.join(ROOT, 'point2cyl_b200', 'dropin')!r})
        sys.path.insert(0, {os.path.join(ROOT, 'point2cyl_b200', 'dropin', 'models')!r})
        MODEL_IMPORTED = importlib.import_module('pointnet_extrusion')      # train_...without_sketch.py:180
        from losses import *                                                # :23
        from data_utils import *                                            # :21
        from models.pointnet_util import PointNetSetAbstractionMsg, PointNetSetAbstraction, PointNetFeaturePropagation
        import losses, data_utils, models.pointnet_util as pu
        for m in (MODEL_IMPORTED, losses, data_utils, pu):
            assert 'point2cyl_b200' in m.__file__, m.__file__
        for n in {LOSSES!r}: assert callable(globals()[n]), n
        for n in {DATA_UTILS!r}: assert callable(globals()[n]), n
        for n in {UTIL!r}: assert hasattr(pu, n), n
        model = MODEL_IMPORTED.backbone(output_sizes=[3, 16])               # :197
        assert len(model.state_dict()) == 123
        print('OK', g_zero_tol)
    """)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=200)
    assert res.returncode == 0 and res.stdout.startswith("OK"), res.stderr[-2000:]


def _params(fn):
    return [(p.name, p.default if p.default is not inspect.Parameter.empty else "<required>")
            for p in inspect.signature(fn).parameters.values() if p.name != "self"]


def test_signatures_equal_the_reference():
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("the reference checkout is not here")
    ref = ref_shim.load()
    from point2cyl_b200.dropin import data_utils as du, losses as ls
    from point2cyl_b200.dropin.models import pointnet_extrusion as pe, pointnet_util as pu
    checked = 0
    for names, ours, theirs in ((LOSSES, ls, ref.losses), (DATA_UTILS, du, ref.data_utils), (UTIL, pu, ref.util)):
        for n in names:
            a, b = getattr(ours, n), getattr(theirs, n)
            if inspect.isclass(a):
                assert _params(a.__init__) == _params(b.__init__), n
                assert [p for p, _ in _params(a.forward)] == [p for p, _ in _params(b.forward)], n
            else:
                pa, pb = _params(a), _params(b)
                assert [p for p, _ in pa] == [p for p, _ in pb], (n, pa, pb)
                assert [d for _, d in pa] == [d for _, d in pb], (n, pa, pb)
            checked += 1
    assert _params(pe.backbone.__init__) == _params(ref.net.backbone.__init__)
    # our forward adds one OPTIONAL trailing argument (explicit FPS start indices); the reference call forward(x) is unchanged
    ours_fwd = _params(pe.backbone.forward)
    assert ours_fwd[0] == ("x", "<required>") and all(d != "<required>" for _, d in ours_fwd[1:])
    assert checked == len(LOSSES) + len(DATA_UTILS) + len(UTIL)


def test_add_noise_equals_the_reference():
    """Host-side augmentation (data_utils.py:84-96): same numpy random stream, same values and dtype."""
    import numpy as np
    import torch
    from oracle import ref_shim
    from point2cyl_b200.dropin import data_utils as du
    g = torch.Generator().manual_seed(0)
    pcs = torch.rand(3, 50, 3, generator=g)
    nrm = torch.nn.functional.normalize(torch.randn(3, 50, 3, generator=g), dim=-1)
    np.random.seed(7)
    out = du.add_noise(pcs, nrm, sigma=0.02)
    assert out.shape == pcs.shape and out.dtype == torch.float64
    off = out - pcs.double()
    along = (off * nrm.double()).sum(-1)
    assert float((off - along[..., None] * nrm.double()).abs().max()) <= 1e-7     # displaced along the normal only
    assert 0.005 < float(along.std()) < 0.04
    if not ref_shim.available():
        return
    np.random.seed(7)
    ref = ref_shim.load().data_utils.add_noise(pcs, nrm, sigma=0.02)
    assert ref.dtype == out.dtype and torch.equal(ref, out)


def test_mask_helpers_equal_the_reference():
    """The torch-only helpers of losses.py (:70-88, :123-124) - no kernel behind them - give the reference's results,
    including clouds with a single instance and the all-masked row."""
    import torch
    from oracle import ref_shim
    from point2cyl_b200.dropin import losses as ls
    g = torch.Generator().manual_seed(2)
    K = 6
    I_gt = torch.stack([torch.randint(0, n, (40,), generator=g) for n in (1, 3, 6, 2)])
    I_gt[:, 0] = torch.tensor([0, 2, 5, 1])                              # make the maximum label present
    mask = ls.get_mask_gt(I_gt, K)
    assert mask.dtype == torch.bool and mask.sum(dim=1).tolist() == [1, 3, 6, 2]
    loss = torch.rand(4, K, generator=g)
    red = ls.reduce_mean_masked_instance(loss, mask)
    assert torch.allclose(red, torch.stack([loss[b, :n].mean() for b, n in enumerate((1, 3, 6, 2))]))
    none = torch.zeros(2, K, dtype=torch.bool)
    assert torch.equal(ls.reduce_mean_masked_instance(loss[:2], none), torch.zeros(2))
    x = torch.tensor([-2.0, -1.0, 0.0, 0.3, 1.0, 2.0])
    assert torch.isfinite(ls.acos_safe(x)).all()
    if not ref_shim.available():
        return
    rl = ref_shim.load().losses
    assert torch.equal(rl.get_mask_gt(I_gt, K), mask)
    assert torch.equal(rl.sequence_mask(torch.tensor([0, 2, 5]), maxlen=5), ls.sequence_mask(torch.tensor([0, 2, 5]), maxlen=5))
    assert torch.equal(rl.sequence_mask(torch.tensor([1, 4])), ls.sequence_mask(torch.tensor([1, 4])))
    assert torch.allclose(rl.reduce_mean_masked_instance(loss, mask), red, atol=1e-7)
    assert torch.equal(rl.acos_safe(x), ls.acos_safe(x))
    n1 = torch.nn.functional.normalize(torch.randn(2, 30, 3, generator=g), dim=-1)
    n2 = torch.nn.functional.normalize(torch.randn(2, 30, 3, generator=g), dim=-1)
    for angle_diff in (False, True):
        for collapse in (False, True):
            assert torch.allclose(rl.compute_normal_loss(n1, n2, angle_diff, collapse),
                                  ls.compute_normal_loss(n1, n2, angle_diff, collapse), atol=1e-7)


def test_module_state_dicts_equal_the_reference():
    """Stand-alone modules of models/pointnet_util.py: same parameter / buffer names and shapes as the reference for
    the single-scale, multi-scale (defined upstream, never instantiated) and feature-propagation classes."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("the reference checkout is not here")
    ru = ref_shim.load().util
    from point2cyl_b200.dropin.models import pointnet_util as pu
    cases = [("PointNetSetAbstraction", dict(npoint=128, radius=0.4, nsample=64, in_channel=131, mlp=[128, 128, 256], group_all=False)),
             ("PointNetSetAbstraction", dict(npoint=None, radius=None, nsample=None, in_channel=259, mlp=[256, 512, 1024], group_all=True)),
             ("PointNetSetAbstractionMsg", dict(npoint=512, radius_list=[0.1, 0.2, 0.4], nsample_list=[32, 64, 128], in_channel=3,
                                                mlp_list=[[32, 32, 64], [64, 64, 128], [64, 96, 128]])),
             ("PointNetFeaturePropagation", dict(in_channel=384, mlp=[256, 128]))]
    for cls, kw in cases:
        a = {k: tuple(v.shape) for k, v in getattr(pu, cls)(**kw).state_dict().items()}
        b = {k: tuple(v.shape) for k, v in getattr(ru, cls)(**kw).state_dict().items()}
        assert a == b, cls


def test_every_function_the_training_scripts_use_is_provided():
    """Parse the reference scripts: every name they take from `losses` / `data_utils` (star imports) exists in the
    drop-in for both training scripts; eval.py additionally needs only the reference's visualisation helpers, which stay
    with the reference (INTEGRATION.md shows the import order for that)."""
    import ast
    ref_root = "/root/reference"
    if not os.path.isfile(os.path.join(ref_root, "train_Point2Cyl.py")):
        pytest.skip("the reference checkout is not here")

    def tree(path):
        return ast.parse(open(path).read().replace("\t", "    "))

    def defined(path):
        return {n.name for n in tree(path).body if isinstance(n, (ast.FunctionDef, ast.ClassDef))}

    def used(path):
        return {n.id for n in ast.walk(tree(path)) if isinstance(n, ast.Name)}

    dropin = os.path.join(ROOT, "point2cyl_b200", "dropin")
    missing = {}
    for script in ("train_Point2Cyl_without_sketch.py", "train_Point2Cyl.py", "eval.py"):
        u = used(os.path.join(ref_root, script))
        for mod in ("losses", "data_utils"):
            need = u & defined(os.path.join(ref_root, mod + ".py"))
            missing[(script, mod)] = sorted(need - defined(os.path.join(dropin, mod + ".py")))
    for script in ("train_Point2Cyl_without_sketch.py", "train_Point2Cyl.py"):
        assert missing[(script, "losses")] == [] and missing[(script, "data_utils")] == [], missing
    assert missing[("eval.py", "losses")] == []
    assert all(n.startswith("visualize_") for n in missing[("eval.py", "data_utils")]), missing
