"""GPU box: cuBLAS TF32 throughput (torch.matmul, allow_tf32) on 8192^3 - burst (best of 10) and sustained (back to back
for 3 s), the way MEASURED_PEAKS.json measures bf16.  Writes profiles/tf32_peak.json (the denominator of bench.py's
per-layer tensor floors) and gpurun_out/tf32_peak.json."""
import json, os, sys, time
import torch

torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
for _ in range(3):
    a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); a @ b; e.record(); e.synchronize()
    best = min(best, s.elapsed_time(e))
t0 = time.perf_counter(); it = 0
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
while time.perf_counter() - t0 < 3.0:
    for _ in range(20):
        a @ b
    it += 20
    torch.cuda.synchronize()
e.record(); e.synchronize()
out = {"tf32_tflops": 2 * n ** 3 / (best / 1e3) / 1e12, "tf32_tflops_sustained": 2 * n ** 3 * it / (s.elapsed_time(e) / 1e3) / 1e12,
       "how": "torch.matmul fp32 8192^3 with allow_tf32 (cuBLAS TF32): best of 10 (burst), back to back for 3 s (sustained)",
       "gpu": torch.cuda.get_device_name()}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ("profiles", "gpurun_out"):
    os.makedirs(os.path.join(root, d), exist_ok=True)
    json.dump(out, open(os.path.join(root, d, "tf32_peak.json"), "w"), indent=1)
print(json.dumps(out))
