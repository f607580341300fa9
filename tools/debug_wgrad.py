import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point2cyl_b200 import ops, _lib
DEV = "cuda"
M, N, K = 4096, 128, 128
for name, dY, X in [
    ("ones", torch.ones(M, N), torch.ones(M, K)),
    ("dY=e_n(col idx), X=1", torch.arange(N).float().repeat(M, 1), torch.ones(M, K)),
    ("dY=1, X=col idx", torch.ones(M, N), torch.arange(K).float().repeat(M, 1)),
    ("rand", torch.randn(M, N), torch.randn(M, K)),
]:
    dW = torch.zeros(N, K, device=DEV); db = torch.zeros(N, device=DEV)
    ops.wgrad(dY.to(DEV), X.to(DEV), K, dW, db, precision=_lib.PREC_3XTF32)
    torch.cuda.synchronize()
    ref = dY.double().T @ X.double()
    print(name, "| dW[0,:4]", dW[0, :4].tolist(), "dW[5,:4]", dW[5, :4].tolist(), "dW[:4,7]", dW[:4, 7].tolist(),
          "| ref[0,:4]", ref[0, :4].tolist(), "ref[5,:4]", ref[5, :4].tolist(), "| db[:4]", db[:4].tolist(),
          "| nonzero frac", float((dW != 0).float().mean()), "maxerr", float((dW.cpu().double() - ref).abs().max()))
