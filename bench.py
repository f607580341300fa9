#!/usr/bin/env python
"""bench.py — clouds/s of Point2Cyl forward+loss at B=32 x N=8192, K=8 per GPU (BASELINE.json config 2).

  python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--precision fp32|3xtf32|bf16]

A "step" is one forward+loss pass over one synthetic batch (S-cyl clouds; random-init
weights of the reference architecture; train-mode BatchNorm; dropout on, as the reference has it).
One JSON line is printed by rank 0:
  value      clouds/s over all ranks, inputs resident in HBM, device-timed (CUDA events per step, L2
             flushed between steps, max over ranks)
  e2e        same through point2cyl_b200.forward_loss_host: pinned HOST batch -> H2D -> forward+loss
             -> six loss scalars D2H, all inside the timed region
  roofline   the dominant kernel, timed live with CUDA events on the launching stream
  cpu_baseline   the CPU oracle (restated reference algorithm, torch CPU, all host threads) on a
             bounded sample, rank 0, N=1 only
  --impl reference   times that CPU path alone (same metric / config / unit)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

B_PER_GPU, N_POINTS, K_INST = 32, 8192, 8
METRIC = "point-clouds/sec forward+loss at B=32 N=8192"
UNIT = "clouds/s"
CPU_SAMPLE_B = 2  # clouds per CPU-baseline step (B=32 needs ~137 GB on the reference's dense axis fit)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        pk = dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops_sustained"], bf16_burst=p["bf16_tflops"], src="measured")
    else:
        pk = dict(hbm=6650.0, bf16=1400.0, bf16_burst=1590.0, src="fallback")
    # MEASURED_PEAKS has no TF32 figure: tools/tf32_peak_probe.py measures cuBLAS TF32 8192^3 on the box
    # (profiles/tf32_peak.json); without it the bf16 peak / 2 (the nominal tf32 : bf16 ratio) stands in, and says so
    tpath = os.path.join(ROOT, "profiles", "tf32_peak.json")
    if os.path.isfile(tpath):
        pk.update(tf32=json.load(open(tpath))["tf32_tflops_sustained"], tf32_src="measured (profiles/tf32_peak.json)")
    else:
        pk.update(tf32=pk["bf16"] / 2, tf32_src="bf16 sustained / 2 (nominal ratio)")
    return pk


def kernel_source_sha():
    """sha256 over the CUDA sources: ties profiles/linear_traffic.json (an ncu pass) to the kernels it measured."""
    import glob
    import hashlib
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(ROOT, "point2cyl_b200", "csrc", "*.cu*"))):
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md recipe).  In-process NVML (nvidia_ml_py) on a
    polling thread: one `nvidia-smi -lms` child per rank attaches to every GPU of the box when it starts - inside the
    timed region of an 8-rank run - and its start-up cost grows with the GPU count; the NVML handle is opened before
    the warm-up instead.  Falls back to the nvidia-smi child when NVML is not importable."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.nvml = self.handle = None
        self._stop = threading.Event()
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = getattr(torch.cuda.get_device_properties(index), "uuid", None)
            if uuid is not None:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(f"GPU-{uuid}")
            else:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = self.handle = None

    def _sample_nvml(self):
        n = self.nvml
        bits = {"hw_slowdown": n.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": n.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": n.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": n.nvmlClocksEventReasonSwPowerCap}
        while True:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                self.rows.append([str(mhz), str(self.max_mhz)] + ["Active" if mask & bits[k] else "Not Active" for k in self.NAMES])
            except Exception:
                pass
            if self._stop.wait(0.02):
                return

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._sample_nvml, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=1.0)
        elif self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            time.sleep(0.15)
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(self.NAMES, r[2:6]) if v == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvml" if self.thread is not None else "nvidia-smi"}


def make_net(device):
    from point2cyl_b200.dropin.models.pointnet_extrusion import backbone
    torch.manual_seed(0)
    net = backbone(output_sizes=[3, 2 * K_INST])  # default nn init = "random-init weights of that architecture"
    return net.to(device).train()


def _quiet_reference_timing(**kw):
    """baseline.ref_arm.time_reference with the reference's import-time prints kept off stdout (one JSON line only)."""
    import contextlib
    from baseline import ref_arm
    with contextlib.redirect_stdout(sys.stderr):
        return ref_arm.time_reference(**kw)


def reference_available():
    from baseline import make_ref
    return make_ref.root() is not None


def run_reference(args, rank, world):
    """--impl reference: the UNMODIFIED reference (baseline/_ref, staged by baseline/make_ref.py: stock backbone
    module + the training script's own loop-body lines exec'd, dense (B,N,N) axis fit) on the host cores, all
    threads, no_grad, a bounded sample of the same workload per step.  Falls back to the oracle port only when the
    reference files did not travel.  Rank 0 only."""
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    extra = {}
    if reference_available():
        r = _quiet_reference_timing(B=CPU_SAMPLE_B, N=N_POINTS, K=K_INST, steps=args.steps, warmup=args.warmup)
        sec, kind = r["sec_per_step"], "reference"
        sample = (f"{CPU_SAMPLE_B} clouds x N={N_POINTS} K={K_INST} per step, {r['steps']} steps after {args.warmup} "
                  f"warm-up, the reference's own modules and loop body ({r['root']}), torch CPU, no_grad")
        try:   # the reference's default batch (4) with autograd recording, as its training loop runs forward+loss
            d = _quiet_reference_timing(B=4, N=N_POINTS, K=K_INST, steps=1, warmup=0, grad=True)
            extra["reference_default_b4_autograd"] = {"value": d["clouds_per_s"], "unit": UNIT,
                                                      "ms_per_step": d["sec_per_step"] * 1e3, "clouds_per_step": 4}
        except (RuntimeError, MemoryError) as e:
            extra["reference_default_b4_autograd"] = {"value": None, "error": str(e)[:120]}
    else:
        from oracle import p2c_oracle as orc
        from point2cyl_b200 import synthetic
        data = synthetic.s_cyl(CPU_SAMPLE_B, N_POINTS, K_INST, seed=1234)
        sd = orc.init_state_dict((3, 2 * K_INST), seed=0)
        starts = (torch.zeros(CPU_SAMPLE_B, dtype=torch.long), torch.zeros(CPU_SAMPLE_B, dtype=torch.long))
        mask = torch.nn.functional.dropout(torch.ones(CPU_SAMPLE_B, 128, N_POINTS), p=0.5)
        times = []
        with torch.no_grad():
            for i in range(args.warmup + args.steps):
                t0 = time.perf_counter()
                orc.forward_loss(sd, data, training=True, fps_start=starts, dropout_mask=mask, dense_axis=True)
                if i >= args.warmup:
                    times.append(time.perf_counter() - t0)
        sec, kind = sum(times) / len(times), "port"
        sample = (f"{CPU_SAMPLE_B} clouds x N={N_POINTS} K={K_INST} per step, {len(times)} steps, oracle port "
                  "(reference files absent), torch CPU no_grad")
    v = CPU_SAMPLE_B / sec
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "ms_per_step_extrapolated_to_b32": sec * 1e3 * B_PER_GPU / CPU_SAMPLE_B, "clouds_per_step": CPU_SAMPLE_B,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": workload_config(args.gpus, args.precision),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, **extra}), flush=True)


def workload_config(n_gpus, precision):
    return {"workload": f"forward+loss, S-cyl synthetic clouds, B={B_PER_GPU}/GPU x N={N_POINTS}, K={K_INST}, "
                        "train-mode BN, all five loss terms (BASELINE.json configs[1])",
            "global_batch": B_PER_GPU * n_gpus, "points": N_POINTS, "K": K_INST, "precision": precision,
            "parallelism": f"dp{n_gpus} (clouds sharded, no data-path collective)",
            "l2": "flushed between timed steps (512 MiB write)"}


def cpu_baseline_sample():
    """rank 0, N=1: the reference's own forward+loss on the host cores (bounded sample); oracle port if absent."""
    torch.set_num_threads(os.cpu_count() or 1)
    if reference_available():
        r = _quiet_reference_timing(B=CPU_SAMPLE_B, N=N_POINTS, K=K_INST, steps=2, warmup=1)
        return {"value": r["clouds_per_s"], "unit": UNIT, "cores": torch.get_num_threads(), "kind": "reference",
                "sample": f"{CPU_SAMPLE_B} clouds x N={N_POINTS} K={K_INST}, mean of 2 steps after 1 warm-up, the reference's "
                          f"own modules and loop body ({r['root']}), torch CPU no_grad"}
    from oracle import p2c_oracle as orc
    from point2cyl_b200 import synthetic
    data = synthetic.s_cyl(CPU_SAMPLE_B, N_POINTS, K_INST, seed=1234)
    sd = orc.init_state_dict((3, 2 * K_INST), seed=0)
    starts = (torch.zeros(CPU_SAMPLE_B, dtype=torch.long), torch.zeros(CPU_SAMPLE_B, dtype=torch.long))
    ts = []
    with torch.no_grad():
        for i in range(3):
            t0 = time.perf_counter()
            orc.forward_loss(sd, data, training=True, fps_start=starts, dense_axis=True)
            ts.append(time.perf_counter() - t0)
    sec = min(ts[1:])
    return {"value": CPU_SAMPLE_B / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{CPU_SAMPLE_B} clouds x N={N_POINTS} K={K_INST}, best of 2 after 1 warm-up, oracle port (reference "
                      "files absent), torch CPU no_grad, dense (B,N,N) axis fit as the reference"}


def gpu_eager_reference(dev):
    """The reference's own modules as eager torch-CUDA ops on this GPU at the headline batch (the only 'GPU path' the
    reference has): forward+loss, no_grad, B=32.  Context for the speed-up, not the product."""
    if not reference_available():
        return None
    try:
        r = _quiet_reference_timing(B=B_PER_GPU, N=N_POINTS, K=K_INST, steps=3, warmup=1, device=str(dev))
    except RuntimeError as e:
        return {"value": None, "error": str(e)[:160]}
    finally:
        torch.cuda.empty_cache()
    return {"value": r["clouds_per_s"], "unit": UNIT, "ms_per_step": r["sec_per_step"] * 1e3, "clouds_per_step": r["B"],
            "what": f"unmodified reference ({r['root']}) on torch-CUDA eager ops, no_grad, wall clock incl. its host syncs"}


# algorithmic work per launch for the roofline (SURVEY.md 8d formulas; B clouds)
def algorithmic_work(name, tag, B):
    N = N_POINTS
    mlp = {  # rows, K, N per layer
        "sa1": (B * 512 * 64, [(3, 64), (64, 64), (64, 128)]),
        "sa2": (B * 128 * 64, [(131, 128), (128, 128), (128, 256)]),
        "sa3": (B * 128, [(259, 256), (256, 512), (512, 1024)]),
        "fp3": (B * 128, [(1280, 256), (256, 256)]),
        "fp2": (B * 512, [(384, 256), (256, 128)]),
        "fp1": (B * N, [(128, 128), (128, 128), (128, 128)]),
        "fc1": (B * N, [(128, 128)]),
        "fc2": (B * N, [(128, 3 + 2 * K_INST)]),
    }
    if name == "p2c_sa_xyz_linear":
        # sa1's second layer with the first (3 -> 64) recomputed in its operand transform: reads the neighbour indices,
        # writes its raw output rows; the first layer's rows never cross HBM
        rows, layers = mlp["sa1"]
        (k0, n0), (k1, n1) = layers[0], layers[1]
        return "linear", (2.0 * rows * (k0 * n0 + k1 * n1), 8.0 * rows + 4.0 * rows * n1)
    if name == "p2c_linear_group_bias":
        # fp3's first layer through its linearity: K = 256 skip features per row, the 1024 pooled channels act as a
        # per-cloud bias (p2c_linear_small, 32 rows)
        rows, layers = mlp["fp3"]
        n = layers[0][1]
        return "linear", (2.0 * rows * 256 * n, 4.0 * rows * (256 + n))
    if name == "p2c_head_masked":
        rows, layers = mlp["fc2"]
        k, n = layers[0]
        return "linear", (2.0 * rows * k * n, 4.0 * rows * (k + n))
    if name == "p2c_linear":
        # (flops, compulsory bytes): a layer reads its raw input rows once and writes its raw output rows once
        # (pooled last layers write rows/nsample instead); weights are negligible
        stage, _, li = tag.partition(".")
        rows, layers = mlp[stage]
        if li == "q":   # feature half of a level's first conv, once per source point (conv linearity)
            k, n = layers[0]
            r = rows // 64
            return "linear", (2.0 * r * (k - 3) * n, 4.0 * r * ((k - 3) + n))
        i = int(li) if li else 0
        k, n = layers[i]
        pooled = stage in ("sa1", "sa2", "sa3") and i == len(layers) - 1
        return "linear", (2.0 * rows * k * n, 4.0 * rows * k + (8.0 * rows / (128 if stage == "sa3" else 64) * n if pooled else 4.0 * rows * n))
    if name == "p2c_sa_first_layer":   # fused gather + first conv: HBM bound on its output rows (+ 8-byte index)
        rows, layers = mlp[tag.partition(".")[0]]
        return "hbm", rows * (4.0 * layers[0][1] + 8.0)
    if name == "p2c_fps":
        return "hbm", B * (12 * N + 8 * 512) if tag == "sa1" else B * (12 * 512 + 8 * 128)
    if name == "p2c_ball_query":
        return "hbm", B * (12 * N + 12 * 512 + 8 * 512 * 64) if tag == "sa1" else B * (12 * 512 + 12 * 128 + 8 * 128 * 64)
    if name == "p2c_three_nn_interp" and tag == "fp1":
        return "hbm", B * (12 * N + 12 * 512 + 4 * 512 * 128 + 4 * N * 128)
    return "hbm", None


def train_step_numbers(net, batch, host, flush, timed, pd, dev, world, steps, warmup):
    """forward + loss + backward + gradient all-reduce (NCCL) + Adam through point2cyl_b200.train.Trainer:
    device-resident and end-to-end (pinned host batch -> H2D -> step -> loss scalar D2H) clouds/s over all ranks."""
    import point2cyl_b200
    from point2cyl_b200 import _lib
    from point2cyl_b200.train import GraphedTrainer, PipelinedTrainer, Trainer
    graphed = not getattr(train_step_numbers, "no_graph", False)
    ms_seq = None
    if graphed:
        # one batch at a time (one CUDA graph pair per step): the per-batch latency figure, reported beside the pipelined one
        seq = GraphedTrainer(net, batch, lr=1e-3)
        for _ in range(warmup):
            seq.step(None)
        pd.barrier()
        ms_seq = timed(lambda: seq.step(None), steps)
        del seq
        tr = PipelinedTrainer(net, batch, lr=1e-3)       # re-flattens the parameters: `seq` must not be used after this
        tr.prime(None)
    else:
        tr = Trainer(net, lr=1e-3)

    loss_host = None

    def step_resident():
        if graphed:
            out = tr.step(None)                           # both slots already hold `batch`; coordinate stage of the next
            tr.join()                                     # one runs beside this step and ends inside the timed region
            return out
        return tr.step(batch)

    def step_e2e():
        if graphed:
            out = tr.step(host)                           # H2D of the NEXT batch's six tensors beside this step, replay
        else:
            out = tr.step({k: host[k].to(dev, non_blocking=True) for k in point2cyl_b200.BATCH_KEYS})
        nonlocal loss_host
        if loss_host is None:
            loss_host = torch.empty(out["losses"].shape, dtype=out["losses"].dtype).pin_memory()
        loss_host.copy_(out["losses"], non_blocking=True)     # D2H of the loss scalars, stream-ordered, inside the event pair
        out["loss_host"] = loss_host
        if graphed:
            tr.join()
        return out

    for _ in range(warmup):
        step_resident()
    step_e2e()
    pd.barrier()
    ms = timed(step_resident, steps)
    l0 = _lib.launch_count
    Trainer.forward_backward(tr, batch)                 # launches per step, counted on one eager pass
    launches = _lib.launch_count - l0 + 1               # + the Adam kernel
    pd.barrier()
    ms_e2e = timed(step_e2e, steps)
    pd.barrier()
    tot, tot_e2e = pd.reduce_max(sum(ms), dev), pd.reduce_max(sum(ms_e2e), dev)
    tot_seq = None if ms_seq is None else pd.reduce_max(sum(ms_seq), dev)
    clouds = B_PER_GPU * world * steps
    n_param = tr.flat_param.numel()
    return {"metric": "point-clouds/sec training step (forward+loss+backward+grad all-reduce+Adam), B=32/GPU N=8192 K=8",
            "value": clouds / (tot / 1e3), "unit": UNIT, "ms_per_step": tot / steps,
            "e2e": {"value": clouds / (tot_e2e / 1e3), "unit": UNIT, "ms_per_step": tot_e2e / steps,
                    "h2d_bytes_per_step": point2cyl_b200.h2d_bytes(host), "d2h_bytes_per_step": 24},
            "gpu_launches": launches, "steps": steps, "warmup": warmup,
            "launch_mode": ("cuda_graphs, two-stage pipeline over batches: coordinate stage of batch i+1 (second stream) | "
                            "layers + loss + backward of batch i, then all-reduce + Adam (main stream); one batch of every "
                            "kind of work per step, side stream joined inside the timed region") if graphed else "eager",
            "sequential": None if ms_seq is None else {
                "value": clouds / (tot_seq / 1e3), "unit": UNIT, "ms_per_step": tot_seq / steps,
                "what": "one batch at a time (train.GraphedTrainer): per-batch latency"},
            "collective": None if world == 1 else f"one NCCL sum all-reduce of the flat fp32 gradient ({n_param} floats)",
            "parameters": n_param, "global_batch": B_PER_GPU * world}


def run_train(args, rank, world, dev, pd, net, host, batch, flush, timed):
    """--workload train: BASELINE.json configs[3] (data-parallel training step, 32 clouds per GPU)."""
    import torch.distributed as dist
    clocks = ClockSampler(dev.index or 0)
    clocks.start()
    t = train_step_numbers(net, batch, host, flush, timed, pd, dev, world, steps=args.steps, warmup=args.warmup)
    clk = clocks.stop()
    if rank == 0:
        cfg = workload_config(world, args.precision)
        cfg["workload"] = ("training step: forward+loss+backward+NCCL gradient all-reduce+Adam, S-cyl synthetic clouds, "
                           f"B={B_PER_GPU}/GPU x N={N_POINTS}, K={K_INST} (BASELINE.json configs[3])")
        cfg["parallelism"] = f"dp{world} (clouds sharded; one flat-gradient all-reduce per step)"
        print(json.dumps({"metric": t["metric"], "value": t["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": t["ms_per_step"], "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "config": cfg,
                          "e2e": t["e2e"], "gpu_launches": t["gpu_launches"], "launch_mode": t["launch_mode"], "sequential": t.get("sequential"),
                          "clocks": clk, "collective": t["collective"]}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_stress(args, rank, world, dev, pd):
    """--workload stress: BASELINE.json configs[4] - FPS + ball query of both set-abstraction levels on
    B=128 x N=32768 uniform clouds (sparse balls: the padding path), clouds sharded over the ranks; achieved HBM GB/s
    on the algorithmic bytes of SURVEY.md 8d (1,139,200 B per cloud)."""
    import torch.distributed as dist
    from point2cyl_b200 import ops, synthetic
    B_total, N = 128, 32768
    lo, hi = pd.shard_range(B_total, rank, world)
    xyz = synthetic.s_uniform(hi - lo, N, seed=77 + rank).to(dev)
    s1 = torch.zeros(hi - lo, dtype=torch.long, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def step():
        _, c1 = ops.fps(xyz, 512, s1)
        g1 = ops.ball_query(0.2, 64, xyz, c1)
        _, c2 = ops.fps(c1, 128, s1)
        g2 = ops.ball_query(0.4, 64, c1, c2)
        return g1, g2

    for _ in range(args.warmup):
        step()
    pd.barrier()
    clocks = ClockSampler(dev.index or 0)
    clocks.start()
    ms = []
    for _ in range(args.steps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        step()
        e.record()
        e.synchronize()
        ms.append(s.elapsed_time(e))
    pd.barrier()
    clk = clocks.stop()
    tot = pd.reduce_max(sum(ms), dev)
    if rank == 0:
        pk = peaks()
        per_cloud = 12 * N + 8 * 512 + 12 * 512 + 8 * 128 + (12 * N + 12 * 512 + 8 * 512 * 64) + (12 * 512 + 12 * 128 + 8 * 128 * 64)
        clouds = B_total * args.steps
        gbs = per_cloud * clouds / (tot / 1e3) / 1e9
        print(json.dumps({"metric": "FPS+ball-query clouds/sec at B=128 N=32768 (stress sweep)", "value": clouds / (tot / 1e3),
                          "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": tot / args.steps, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
                          "config": {"workload": "FPS (N->512->128) + ball query (r=.2/.4, nsample 64) on S-uniform clouds, "
                                                 "B=128 total x N=32768 (BASELINE.json configs[4])",
                                     "global_batch": B_total, "points": N, "parallelism": f"dp{world} (clouds sharded)",
                                     "l2": "flushed between timed steps (512 MiB write)"},
                          "roofline": {"kernel": "p2c_fps + p2c_ball_query", "bound": "hbm", "achieved": gbs,
                                       "peak": pk["hbm"] * world, "unit": "GB/s", "frac": gbs / (pk["hbm"] * world),
                                       "traffic": None,
                                       "note": f"algorithmic bytes {per_cloud} per cloud (SURVEY.md 8d); FPS is bound by "
                                               "its 640 dependent arg-max rounds, not by HBM"},
                          "gpu_launches": 4, "clocks": clk}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_igr(args, rank, world, dev, pd):
    """--workload igr: the implicit sketch network of the with-sketch trainer at its training shape -
    B=32 clouds x K=8 sketch instances x (1024 on-surface + 1152 off-surface) points = 557,056 rows per GPU through the
    8 x 512 softplus network: forward sweep, closed-form input gradient, the four loss terms, PointNetEncoder latents
    (train_Point2Cyl.py:598-672).  Tensor bound: reported against the TF32 peak (3 tf32 MMA passes per product)."""
    import torch.distributed as dist
    from point2cyl_b200 import igr
    from point2cyl_b200.dropin.IGR import network as dnet
    from point2cyl_b200.dropin.IGR.sampler import NormalPerPoint
    B, K, S = B_PER_GPU, K_INST, 1024
    I = B * K
    torch.manual_seed(0)
    net = dnet.ImplicitNet(d_in=258, dims=[512] * 8, skip_in=[4], geometric_init=True, radius_init=1, beta=100).to(dev)
    enc = dnet.PointNetEncoder(256, 2, with_normals=True).to(dev).train()
    enc_gt = dnet.PointNetEncoder(256, 2, with_normals=True).to(dev).train()
    g = torch.Generator().manual_seed(1 + rank)
    ang = torch.rand(I, S, generator=g) * 6.2831853
    rad = 0.5 + 0.3 * torch.rand(I, 1, generator=g)
    pts = torch.stack([rad * torch.cos(ang), rad * torch.sin(ang)], -1)
    nrm = torch.stack([torch.cos(ang), torch.sin(ang)], -1)
    sk = torch.cat([pts, nrm], -1).to(dev)                       # (I, S, 4) circles: synthetic sketches
    mask_gt = torch.ones(B, K, dtype=torch.bool, device=dev)
    sampler = NormalPerPoint(1.8, 0.01)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def step():
        latent = enc(sk)
        latent_gt = enc_gt(sk)
        off = sampler.get_points(sk[:, :, :2])
        return igr.sketch_loss_block(net, latent, latent_gt, sk[:, :, :2], sk[:, :, 2:], off, mask_gt)

    with torch.no_grad():
        for _ in range(args.warmup):
            out = step()
    pd.barrier()
    clocks = ClockSampler(dev.index or 0)
    clocks.start()
    ms = []
    with torch.no_grad():
        for _ in range(args.steps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = step()
            e.record()
            e.synchronize()
            ms.append(s.elapsed_time(e))
        # per entry point
        _lib_mod = __import__("point2cyl_b200._lib", fromlist=["x"])
        _lib_mod.profile_start()
        step()
        prof = _lib_mod.profile_stop()
    # the training step of the same block: im_loss.backward() through the closed-form second-order sweeps
    # (igr.implicit_backward) and the encoders' layer backward, torch.optim.Adam as in the script (train_Point2Cyl.py:686-690)
    params = [p for m in (net, enc, enc_gt) for p in m.parameters()]
    opt = torch.optim.Adam(params, lr=1e-4)

    def train_step():
        opt.zero_grad(set_to_none=True)
        o = step()
        o["im_loss"].backward()
        opt.step()
        return o

    tms = []
    for it in range(2 + max(3, args.steps // 2)):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        train_step()
        e.record()
        e.synchronize()
        if it >= 2:
            tms.append(s.elapsed_time(e))
    _lib_mod.profile_start()
    train_step()
    tprof = _lib_mod.profile_stop()
    pd.barrier()
    clk = clocks.stop()
    tot = pd.reduce_max(sum(ms), dev)
    ttot = pd.reduce_max(sum(tms), dev)
    if rank == 0:
        pk = peaks()
        R = I * (S + S + S // 8)
        bwd_ms = {}
        for n, tag, t in tprof:
            if "bwd" in tag:
                bwd_ms[n] = round(bwd_ms.get(n, 0.0) + t, 3)
        fwd = 2.0 * (258 * 512 + 2 * 512 * 512 + 512 * 254 + 4 * 512 * 512 + 512)
        rev = 2.0 * (5 * 512 * 512 + 2 * 512 * 254)
        flops = R * (fwd + rev)
        act_ms = sum(t for n, tag, t in prof if n == "p2c_linear_act")
        n_act = sum(1 for n, tag, t in prof if n == "p2c_linear_act")
        tf = flops / (act_ms / 1e3) / 1e12
        stages = {}
        for n, tag, t in prof:
            d = stages.setdefault(n, {"ms": 0.0, "launches": 0})
            d["ms"] = round(d["ms"] + t, 4)
            d["launches"] += 1
        print(json.dumps({
            "metric": "sketch instances/sec, implicit network forward + input gradient + loss (with-sketch trainer block)",
            "value": I * world * args.steps / (tot / 1e3), "unit": "instances/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": tot / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32 (3xTF32 on tcgen05)", "data": "synthetic",
            "config": {"workload": f"IGR block: B={B}/GPU x K={K} instances x ({S} + {S + S // 8}) points = {R} rows through "
                                   "ImplicitNet 8x512 softplus(100) skip@4, PointNetEncoder latents, 4 loss terms "
                                   "(train_Point2Cyl.py:598-672), forward values", "rows": R,
                       "parallelism": f"dp{world} (instances sharded)", "l2": "flushed between timed steps (512 MiB write)"},
            "roofline": {"kernel": "p2c_linear_act", "bound": "tensor", "achieved": 3.0 * tf, "peak": pk["tf32"],
                         "unit": "TFLOP/s", "frac": 3.0 * tf / pk["tf32"], "traffic": None,
                         "algorithmic_tflops": tf, "launches": n_act, "ms": act_ms, "tf32_peak_source": pk["tf32_src"],
                         "note": "achieved = 3 x algorithmic FLOP (three tf32 MMA passes per fp32-faithful product: hi*hi + "
                                 "lo*hi + hi*lo) / summed CUDA-event time of the 15 layer launches; algorithmic work per row "
                                 f"{fwd + rev:.0f} FLOP (forward {fwd:.0f}, input-gradient sweep {rev:.0f})"},
            "losses": {k: float(out[k]) for k in ("im_loss", "mnfld_loss", "grad_loss", "normals_loss", "latent_loss")},
            "train_step": {"metric": "sketch instances/sec, the same block + im_loss.backward() + Adam",
                           "value": I * world * len(tms) / (ttot / 1e3), "unit": "instances/s",
                           "ms_per_step": ttot / len(tms), "steps": len(tms), "gpu_launches": len(tprof),
                           "backward_ms_by_entry_point": bwd_ms},
            "gpu_launches": len(prof), "stages": stages, "clocks": clk}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("P2C_PRECISION", "3xtf32"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernel by kernel instead of replaying a CUDA graph")
    ap.add_argument("--mode", default="deep", choices=["deep", "pipelined", "sequential"],
                    help="deep (default): graph.DeepPipelinedForwardLoss - per step the coordinate-only stage of batch i+1 "
                         "(second stream), the layers of batch i (main stream) and the loss block of batch i-1 (third "
                         "stream); pipelined: graph.PipelinedForwardLoss - two stages, the loss behind the layers on the "
                         "main stream; sequential: one batch at a time (graph.GraphedForwardLoss)")
    ap.add_argument("--workload", default="forward_loss", choices=["forward_loss", "train", "stress", "igr"],
                    help="forward_loss = BASELINE.json configs[1] (the headline metric, default); train = configs[3], "
                         "the data-parallel training step (32 clouds per GPU); stress = configs[4], FPS + ball query "
                         "at B=128 x N=32768")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    import point2cyl_b200
    from point2cyl_b200 import _lib, pipeline, synthetic
    from point2cyl_b200 import dist as pd

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pd.init("nccl")          # one process per GPU; NCCL is used for the timing barrier / max-reduce only
    pipeline.set_precision(args.precision)
    _lib.load()

    if args.workload == "stress":
        run_stress(args, rank, world, dev, pd)
        return
    if args.workload == "igr":
        run_igr(args, rank, world, dev, pd)
        return

    net = make_net(dev)
    host = point2cyl_b200.pin_batch(synthetic.s_cyl(B_PER_GPU, N_POINTS, K_INST, seed=1234 + rank))
    batch = {k: v.to(dev) for k, v in host.items()}
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    barrier = pd.barrier

    def timed(fn, steps):
        """per-step CUDA-event time on the current stream, L2 flushed (untimed, in stream order) between steps.  The host
        does not wait between steps - it synchronises once after the last one - so a step's events bracket device work
        only, never the host's launch latency of the next step (the caller's barrier + synchronize bracket the K steps)."""
        evs = []
        for _ in range(steps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        return [s.elapsed_time(e) for s, e in evs]

    graphed = pipe = None
    if not args.no_graph:
        from point2cyl_b200.graph import GraphedForwardLoss, PipelinedForwardLoss
        graphed = GraphedForwardLoss(net, batch, precision=args.precision)
        if args.mode in ("pipelined", "deep") and args.workload == "forward_loss":
            gsm = os.environ.get("P2C_GEO_SMS")          # tools: sweep of the SMs left to the coordinate stage
            if args.mode == "deep":
                from point2cyl_b200.graph import DeepPipelinedForwardLoss
                pipe = DeepPipelinedForwardLoss(net, batch, precision=args.precision, geometry_sms=int(gsm) if gsm else None,
                                                partition=os.environ.get("P2C_PARTITION", "soft"))
                pipe.prime(None)
                pipe.step(None)                          # fill the third stage: every later step returns a loss
            else:
                pipe = PipelinedForwardLoss(net, batch, precision=args.precision, geometry_sms=int(gsm) if gsm else None)
                pipe.prime(None)

    def step_eager():
        with torch.no_grad():
            return pipeline.forward_loss(net, batch)

    def step_sequential():
        return graphed() if graphed is not None else step_eager()     # static inputs already hold `batch`

    def step_resident():
        if pipe is not None:
            # loss of the resident current batch + the coordinate stage of the (resident) next one; join(): the timed
            # region ends only when the side-stream work launched in it has finished too
            out = pipe.step(None)
            pipe.join()
            return out
        return step_sequential()

    losses_host = torch.empty(6, dtype=torch.float32).pin_memory()

    def step_e2e():
        # the six loss scalars of the step go device -> pinned host inside the step's event pair (stream-ordered copy;
        # the host reads them after the synchronize that ends the timed region)
        if pipe is not None:
            out = pipe.step(host)                 # H2D of the NEXT batch's six tensors + its coordinate stage, beside
            losses_host.copy_(out["losses"], non_blocking=True)   # the layers + loss of the current one; D2H of its losses
            out["losses_host"] = losses_host
            pipe.join()
            return out
        if graphed is not None:
            out = graphed(host)                   # H2D of the six batch tensors, replay
            losses_host.copy_(out["losses"], non_blocking=True)
            out["losses_host"] = losses_host
            return out
        with torch.no_grad():
            return point2cyl_b200.forward_loss_host(net, host, device=dev)

    train_step_numbers.no_graph = args.no_graph
    if args.workload == "train":
        run_train(args, rank, world, dev, pd, net, host, batch, flush, timed)
        return

    clocks = ClockSampler(local_rank)               # NVML handle opened here, before the warm-up and the barrier
    with torch.no_grad():
        for _ in range(args.warmup):
            step_resident()
        for _ in range(2):
            step_e2e()
    barrier()
    clocks.start()
    ms = timed(step_resident, args.steps)
    ms_seq = timed(step_sequential, args.steps) if pipe is not None else ms
    l0 = _lib.launch_count
    step_eager()                                  # launches per step, counted on one eager pass (a replay re-issues them)
    launches = _lib.launch_count - l0
    barrier()
    ms_e2e = timed(step_e2e, args.steps)
    barrier()
    clk = clocks.stop()

    total_ms = pd.reduce_max(sum(ms), dev)
    total_ms_e2e = pd.reduce_max(sum(ms_e2e), dev)
    total_ms_seq = pd.reduce_max(sum(ms_seq), dev)
    clouds = B_PER_GPU * world * args.steps
    value = clouds / (total_ms / 1e3)
    e2e = clouds / (total_ms_e2e / 1e3)

    # ---- per-kernel live timing for the roofline (rank 0) --------------------------------------
    roof, stages = None, None
    if rank == 0:
        pk = peaks()
        agg = {}
        reps = 3
        for _ in range(reps):
            flush.zero_()
            # keep the GPU busy while the host enqueues the whole eager pass (~45 launches): with an idle GPU the
            # CUDA-event pair around a launch also times the host's own work between the two records (tensor-map
            # encoding, ctypes), 5-20 us per launch; queued behind a spin kernel the pairs time the device only
            torch.cuda._sleep(int(1.2e7))
            _lib.profile_start()
            step_eager()
            for name, tag, t in _lib.profile_stop():
                a = agg.setdefault((name, tag), [0.0, 0])
                a[0] += t
                a[1] += 1
        per_kernel = {}
        # every launch of the tcgen05 layer kernels counts as "p2c_linear": the plain layers, sa1's second layer with the
        # first one recomputed inside it, fp3's first layer with its per-cloud bias, and the output heads
        layer_group = {"p2c_sa_xyz_linear": "p2c_linear", "p2c_linear_group_bias": "p2c_linear",
                       "p2c_head_masked": "p2c_linear"}
        layer_tag = {"p2c_sa_xyz_linear": "sa1.1", "p2c_linear_group_bias": "fp3.0", "p2c_head_masked": "fc2"}
        for (name, tag), (t, n) in agg.items():
            d = per_kernel.setdefault(layer_group.get(name, name), {"ms": 0.0, "bytes": 0.0, "flops": 0.0, "launches": 0})
            bound, work = algorithmic_work(name, tag, B_PER_GPU)
            d["ms"] += t / reps
            d["launches"] += n // reps
            if bound == "linear":
                d["flops"] += work[0]
                d["bytes"] += work[1]
            elif work:
                d["bytes"] += work
        stages = {}
        for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1]["ms"]):
            e = {"ms": round(v["ms"], 4), "launches": v["launches"]}
            if v["bytes"]:
                e["hbm_gbs"] = round(v["bytes"] / (v["ms"] / 1e3) / 1e9, 1)
                e["hbm_frac"] = round(e["hbm_gbs"] / pk["hbm"], 4)
            if v["flops"]:
                e["tflops"] = round(v["flops"] / (v["ms"] / 1e3) / 1e12, 1)
            stages[k] = e
        # per MLP layer: both floors and which one binds (3xTF32 = three tf32 MMA passes per product)
        layers, floor_us, meas_us = {}, 0.0, 0.0
        for (name, tag), (t, n) in sorted(agg.items(), key=lambda kv: kv[0][1]):
            if layer_group.get(name, name) != "p2c_linear":
                continue
            _, (fl, by) = algorithmic_work(name, tag, B_PER_GPU)
            tag = layer_tag.get(name, tag)
            us = t / reps * 1e3
            hbm_floor = by / (pk["hbm"] * 1e9) * 1e6
            tc_floor = 3.0 * fl / (pk["tf32"] * 1e12) * 1e6
            fl_us = max(hbm_floor, tc_floor)
            layers[tag] = {"us": round(us, 1), "hbm_floor_us": round(hbm_floor, 1), "tensor_floor_us": round(tc_floor, 1),
                           "bound": "hbm" if hbm_floor >= tc_floor else "tensor", "frac_of_floor": round(fl_us / us, 3)}
            floor_us += fl_us
            meas_us += us
        stages["p2c_linear"]["layers"] = layers
        top = max(per_kernel, key=lambda k: per_kernel[k]["ms"])
        d = per_kernel[top]
        traffic, traffic_note, ncu_us = None, None, None
        tpath = os.path.join(ROOT, "profiles", "linear_traffic.json")   # ncu dram bytes of the same launches
        if top == "p2c_linear" and os.path.isfile(tpath):
            tj = json.load(open(tpath))
            if tj.get("kernel_source_sha") == kernel_source_sha():
                traffic = tj.get("dram_bytes_per_step")
                traffic_note = tj.get("source")
                ncu_us = tj.get("ncu_duration_us_per_step")
            else:
                traffic_note = "profiles/linear_traffic.json was measured on other kernel sources: not reported"
        ach = d["bytes"] / (d["ms"] / 1e3) / 1e9 if d["bytes"] else 0.0
        roof = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                "frac": ach / pk["hbm"], "traffic": traffic, "traffic_source": traffic_note,
                "note": f"the {d['launches']} tcgen05 layer launches of one step taken together (p2c_linear, "
                        "p2c_sa_xyz_linear, p2c_linear_group_bias, p2c_head_masked): compulsory bytes "
                        f"(4*rows*(K+N) per layer) / summed CUDA-event time, peak = {pk['src']} copy bandwidth.  Per "
                        "layer (stages.p2c_linear.layers) the binding floor is HBM for the K, N <= 128 layers on the big "
                        "row counts and the tensor pipe (3 tf32 passes) for the pooled 64->128 / 128->256 layers and the "
                        "coarse levels"}
        if ncu_us:
            # the same launches' gpu__time_duration under ncu (committed launch list, same kernel sources; cold cache and
            # serialised): a CUDA-event pair around ONE launch also spans the launch gap on either side of the kernel
            # (5-8 us per launch here, 17 launches), which ncu's per-kernel duration does not
            roof["ncu_duration_us"] = ncu_us
            roof["frac_by_ncu_duration"] = d["bytes"] / (ncu_us * 1e-6) / 1e9 / pk["hbm"] if d["bytes"] else None
        if d["flops"]:
            tf = d["flops"] / (d["ms"] / 1e3) / 1e12
            roof["tensor_tflops"] = tf
            roof["tensor_frac_of_bf16_peak"] = tf / pk["bf16"]
            roof["tf32_mma_tflops"] = 3.0 * tf
            roof["tf32_mma_frac_of_tf32_peak"] = 3.0 * tf / pk["tf32"]
            roof["tf32_peak"] = {"tflops": pk["tf32"], "source": pk["tf32_src"]}
            roof["frac_of_per_layer_floor"] = floor_us / meas_us if meas_us else None
    # ---- inference mode (eval.py: BatchNorm on running statistics): sa1 then runs as ONE kernel (p2c_sa_stack_fused) ----
    eval_leg = None
    if rank == 0 and graphed is not None:
        net.eval()
        try:
            ge = GraphedForwardLoss(net, batch, precision=args.precision)
            for _ in range(3):
                ge()
            ems = timed(ge, 10)
            _lib.profile_start()
            with torch.no_grad():
                pipeline.forward_loss(net, batch)
            names = {}
            for n_, tg, t in _lib.profile_stop():
                names[n_] = round(names.get(n_, 0.0) + t * 1e3, 1)
            eval_leg = {"metric": "point-clouds/sec forward+loss, eval-mode BatchNorm (one batch at a time, CUDA graph)",
                        "value": B_PER_GPU * 10 / (sum(ems) / 1e3), "unit": UNIT, "ms_per_step": sum(ems) / 10,
                        "train_mode_same_launch_mode_ms": total_ms_seq / args.steps,
                        "sa1_one_kernel_us": names.get("p2c_sa_stack_fused"),
                        "what": "sa1 (gather, 3->64->64->128 conv/BN/ReLU, max-pool) is ONE tcgen05 kernel here "
                                "(csrc/sa_stack_tc.cu): no activation of the level crosses HBM"}
            del ge
        finally:
            net.train()
    # ---- the training step of configs[3] on the same batch (all ranks: it contains the gradient all-reduce) ----
    train = train_step_numbers(net, batch, host, flush, timed, pd, dev, world, steps=10, warmup=3)
    cpu = eager = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_sample()
        eager = gpu_eager_reference(dev)

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": workload_config(world, args.precision),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": point2cyl_b200.h2d_bytes(host),
                    "d2h_bytes_per_step": 24 + (0 if graphed is not None else B_PER_GPU * K_INST * K_INST * 4 + B_PER_GPU * 4), "ms_per_step": total_ms_e2e / args.steps},
            "gpu_launches": launches,
            "launch_mode": "eager" if graphed is None else ("cuda_graph" if pipe is None else
                           (f"cuda_graphs, three-stage pipeline over batches: coordinate stage of batch i+1 (second stream, "
                            f"{pipe.geometry_sms} SMs left to it) | layers of batch i (main stream) | loss block of batch i-1 "
                            "(third stream); one batch of every kind of work per step, launched and completed inside the "
                            "timed region (side streams joined); the step returns the loss of batch i-1"
                            if args.mode == "deep" else
                            f"cuda_graphs, two-stage pipeline over batches (coordinate stage of batch i+1 on a second stream, "
                            f"{pipe.geometry_sms} SMs left to it, beside the layers + loss of batch i; one batch of every kind "
                            "of work per step, side stream joined inside the timed region)")),
            "sequential": None if pipe is None else {"value": B_PER_GPU * world * args.steps / (total_ms_seq / 1e3),
                                                     "unit": UNIT, "ms_per_step": total_ms_seq / args.steps,
                                                     "what": "one batch at a time (graph.GraphedForwardLoss): per-batch latency"},
            "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
            "gpu_eager_reference": eager, "stages": stages, "eval_forward": eval_leg, "train_step": train}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
