"""GPU-vs-GPU bar of SURVEY.md 8(d): the oracle's restatement of the reference algorithm run as EAGER torch-CUDA ops
on one B200 (what the unmodified reference does on a GPU: (B,S,N) distance tensors, full sorts, Python FPS loop,
host-side scipy assignment, dense (B,N,N) axis fit), same workload as bench.py's headline (B=32, N=8192, K=8,
forward+loss, no_grad).  This is a measurement tool: it executes oracle/ only as the thing timed BESIDE the product."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import p2c_oracle as orc
from point2cyl_b200 import synthetic

B, N, K = 32, 8192, 8
steps, warmup = 5, 2
data = {k: v.cuda() for k, v in synthetic.s_cyl(B, N, K, seed=1234).items()}      # CPU generators first ...
sd = {k: v.cuda() for k, v in orc.init_state_dict((3, 2 * K), seed=0).items()}
torch.set_default_device("cuda")                                                  # ... then every factory -> cuda
starts = (torch.zeros(B, dtype=torch.long), torch.zeros(B, dtype=torch.long))
res = {}
for dense in (True, False):
    ts = []
    with torch.no_grad():
        for i in range(warmup + steps):
            mask = torch.nn.functional.dropout(torch.ones(B, 128, N), p=0.5)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = orc.forward_loss(sd, data, training=True, fps_start=starts, dropout_mask=mask, dense_axis=dense)
            float(out["total"])
            torch.cuda.synchronize()
            if i >= warmup:
                ts.append(time.perf_counter() - t0)
    ts.sort()
    res["dense_axis" if dense else "closed_form_axis"] = {"ms_per_step": 1e3 * ts[len(ts) // 2],
                                                          "clouds_per_s": B / ts[len(ts) // 2]}
print(json.dumps({"impl": "oracle port, eager torch-CUDA", "B": B, "N": N, "K": K, "steps": steps, "warmup": warmup,
                  "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30, **res}))
