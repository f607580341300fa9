"""GPU box: how exact is one MLP layer?  3xTF32 tcgen05 kernel vs the fp32 SIMT kernel vs torch fp32 matmul, all against
the float64 product of the same operands (max |err| / max |ref|, rms err / rms ref, the proportional
coefficient c of err ~ c * ref, and the rms of the residual after removing it).  Feeds DESIGN.md section 2."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from point2cyl_b200 import _lib, ops  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"
res = []
for M, K, N in [(16384, 64, 64), (16384, 64, 128), (16384, 128, 128), (16384, 128, 256), (4096, 256, 512),
                (4096, 512, 1024), (4096, 1280, 256)]:
    g = torch.Generator().manual_seed(K + N)
    X = torch.relu(torch.randn(M, K, generator=g)).to(DEV)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    ref = X.double() @ W.double().t()

    def err(Y):
        d = Y.double() - ref
        c = float((d * ref).sum() / (ref * ref).sum())          # proportional part: d ~ c * ref (truncating accumulate)
        r = d - c * ref
        return (float(d.abs().max() / ref.abs().max()), float(d.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()), c,
                float(r.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()))

    row = {"M": M, "K": K, "N": N}
    ws = ops.weight_operand(X, W, N, K, False, 0, _lib.PREC_3XTF32)
    row["3xtf32"] = err(ops.linear(X, W, None, K=K, precision=_lib.PREC_3XTF32, w_split=ws))
    row["simt_fp32"] = err(ops.linear(X, W, None, K=K, precision=_lib.PREC_FP32))
    row["torch_fp32"] = err(X @ W.t())
    res.append(row)
    print(json.dumps(row), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/precision_probe.json", "w"), indent=1)
