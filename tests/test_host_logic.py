"""Host-side logic that needs no GPU: the drop-in's module surface (state_dict keys, BN naming the training script's
update_momentum relies on), the synthetic workload generator, and bench.py's algorithmic-work formulas."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import p2c_oracle as orc
from point2cyl_b200 import synthetic
from point2cyl_b200.dropin.models.pointnet_extrusion import backbone

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CKPT = "/root/reference/results/Point2Cyl_without_sketch/model.pth"


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_state_dict_surface_matches_the_restated_reference():
    """Same parameter / buffer names and shapes as models/pointnet_extrusion.py:8-35 (what load_state_dict at
    eval.py:207 needs): 123 entries for heads [3, 16], 1,404,243 trainable floats."""
    K = 8
    net = backbone(output_sizes=[3, 2 * K])
    sd, ref = net.state_dict(), orc.init_state_dict((3, 2 * K), seed=0)
    assert set(sd.keys()) == set(ref.keys()) and len(sd) == 123
    for k in sd:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    assert sum(p.numel() for p in net.parameters()) == 1_404_243
    net.load_state_dict(ref, strict=True)


@pytest.mark.skipif(not os.path.isfile(REF_CKPT), reason="the reference checkout (with its shipped checkpoint) is not here")
def test_shipped_checkpoint_loads_strict():
    ck = torch.load(REF_CKPT, map_location="cpu", weights_only=False)
    sd = ck["model"] if "model" in ck else ck
    n_out = [v.shape[0] for k, v in sd.items() if k.startswith("fc2.") and k.endswith(".weight")]
    net = backbone(output_sizes=n_out)
    net.load_state_dict(sd, strict=True)


def test_update_momentum_reaches_every_batchnorm():
    """train_Point2Cyl_without_sketch.py:153-156 sets `.momentum` on every module whose qualified name contains 'bn';
    the kernels read `bn.momentum`, so every BatchNorm must be reachable that way and nothing else may be hit."""
    net = backbone(output_sizes=[3, 16])
    bns = {n for n, m in net.named_modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)}
    named = {n for n, m in net.named_modules() if "bn" in n and hasattr(m, "momentum")}
    assert bns and bns == named
    for n, m in net.named_modules():
        if "bn" in n:
            m.momentum = 0.25
    assert all(m.momentum == 0.25 for m in net.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm))


def test_forward_without_cuda_raises_instead_of_falling_back():
    from point2cyl_b200._lib import P2CError
    net = backbone(output_sizes=[3, 8])
    with pytest.raises(P2CError):
        net(torch.rand(1, 256, 3))


@pytest.mark.parametrize("B,N,K,seed", [(3, 1024, 4, 0), (2, 4096, 8, 1234), (1, 777, 16, 5)])
def test_synthetic_s_cyl_contract(B, N, K, seed):
    """SURVEY.md 8(d) S-cyl: gap-free labels 0..n_inst-1 (losses.py:34 needs them), unit normals, clouds centred and
    scaled into the unit sphere, barrel normals perpendicular to the instance axis and cap normals along it, padded
    gt axes / centres, deterministic in the seed."""
    d = synthetic.s_cyl(B, N, K, seed)
    assert d["pcs"].shape == (B, N, 3) and d["normals"].shape == (B, N, 3)
    assert d["inst"].dtype == torch.long and d["bb"].dtype == torch.long
    assert d["axes"].shape == (B, K, 3) and d["centers"].shape == (B, K, 3)
    for b in range(B):
        n_inst = int(d["inst"][b].max()) + 1
        assert 1 <= n_inst <= K
        assert sorted(d["inst"][b].unique().tolist()) == list(range(n_inst))
        assert bool((d["axes"][b, n_inst:] == 0).all()) and bool((d["centers"][b, n_inst:] == 0).all())
        assert torch.allclose(d["axes"][b, :n_inst].norm(dim=-1), torch.ones(n_inst), atol=1e-5)
        ax = d["axes"][b][d["inst"][b]]                                   # (N,3) axis of each point's instance
        dots = (ax * d["normals"][b]).sum(-1).abs()
        assert float(dots[d["bb"][b] == 0].max()) <= 1e-4                 # barrel: normal perpendicular to the axis
        assert float((dots[d["bb"][b] == 1] - 1).abs().max()) <= 1e-4     # base: normal = +-axis
    assert torch.allclose(d["normals"].norm(dim=-1), torch.ones(B, N), atol=1e-5)
    assert float(d["pcs"].norm(dim=-1).max()) <= 1.0 + 1e-5
    assert set(d["bb"].unique().tolist()) <= {0, 1}
    d2 = synthetic.s_cyl(B, N, K, seed)
    assert all(torch.equal(d[k], d2[k]) for k in d)


def test_algorithmic_work_matches_survey_table():
    """The bytes / flops bench.py divides by the measured kernel times are SURVEY.md 8(d)'s per-cloud figures."""
    bm = _bench()
    assert (bm.N_POINTS, bm.K_INST, bm.B_PER_GPU) == (8192, 8, 32)
    B = 1
    assert bm.algorithmic_work("p2c_fps", "sa1", B) == ("hbm", 102_400)
    assert bm.algorithmic_work("p2c_fps", "sa2", B) == ("hbm", 7_168)
    assert bm.algorithmic_work("p2c_ball_query", "sa1", B) == ("hbm", 366_592)
    assert bm.algorithmic_work("p2c_ball_query", "sa2", B) == ("hbm", 73_216)
    assert bm.algorithmic_work("p2c_three_nn_interp", "fp1", B) == ("hbm", 4_560_896)
    # the MLP stack: the per-layer flops add up to the table's 3.431 GFLOP per cloud.  sa1.0 / sa2.0 run as the fused
    # gather + first conv (xyz part) plus the once-per-source-point feature half `.q` (conv linearity), so their flops
    # are counted from the table's layer dims instead
    layers = {"sa1": 3, "sa2": 3, "sa3": 3, "fp3": 2, "fp2": 2, "fp1": 3, "fc1": 1, "fc2": 1}
    flops = 0.0
    for stage, n in layers.items():
        for i in range(n):
            kind, (f, by) = bm.algorithmic_work("p2c_linear", f"{stage}.{i}", B)
            assert kind == "linear" and f > 0 and by > 0
            flops += f
    assert abs(flops / 3.431e9 - 1.0) < 2e-3
    # a pooled last layer writes max AND min per group instead of the rows: 8 bytes * rows / group * width
    _, (f, by) = bm.algorithmic_work("p2c_linear", "sa1.2", B)
    rows = 512 * 64
    assert by == 4.0 * rows * 64 + 8.0 * rows / 64 * 128


def test_bn_decay_schedule_values():
    """get_batch_norm_decay (train_Point2Cyl_without_sketch.py:143-151) gives 0.5 at step 0 and clips at 0.01: the
    kernels must accept the whole range as `momentum` (checked on the GPU in test_gpu_parity; here the host values)."""
    def decay(step, bs=32, every=200000):
        return max(0.5 * 0.5 ** int(np.floor(step * bs / every)), 0.01)
    assert decay(0) == 0.5 and decay(6250) == 0.25 and decay(10 ** 7) == 0.01


@pytest.mark.timeout(300)
def test_reference_arm_under_torchrun_prints_one_line():
    """`bench.py --impl reference` launched like the driver launches it for N=2: rank 0 alone times the CPU reference and
    prints ONE JSON line with the contract's keys; the other rank exits 0 without work."""
    import json
    import socket
    import subprocess
    import sys
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "0"]
    res = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=280)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["steps"] == 1 and d["warmup"] == 0
    assert d["metric"].startswith("point-clouds/sec") and d["unit"] == "clouds/s" and d["higher_is_better"] is True
    from baseline import make_ref
    kind = "reference" if make_ref.root() else "port"      # the reference's own files when staged (baseline/_ref)
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == kind
    assert abs(d["ms_per_step"] * 1e-3 * d["value"] - d["clouds_per_step"]) < 1e-6      # measured time of the sample
    assert res.stdout.count("\n") == 1                     # nothing but the line on stdout (import-time prints)
    assert d["cpu_baseline"]["cores"] >= 1 and "clouds" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("forward+loss") and d["gpu_launches"] == 0


def test_clock_sampler_degrades_without_a_gpu():
    """bench.py's clock sampler (in-process NVML, nvidia-smi child as the fallback) must never take the bench down: on a
    machine with neither it reports that instead of raising."""
    spec = importlib.util.spec_from_file_location("bench_mod_clk", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    c = bench.ClockSampler(0)
    c.start()
    out = c.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    if not torch.cuda.is_available():
        assert out["sm_mhz"] is None or isinstance(out["sm_mhz"], float)


def test_pipelined_trainer_surface():
    """train.PipelinedTrainer keeps Trainer's interface (step / forward_backward / flat buffers) and adds prime / join."""
    from point2cyl_b200 import train
    assert issubclass(train.PipelinedTrainer, train.Trainer) and issubclass(train.GraphedTrainer, train.Trainer)
    for name in ("prime", "step", "join", "rebuild_if_stale", "stale"):
        assert callable(getattr(train.PipelinedTrainer, name)), name
