#!/usr/bin/env python
"""Three representative backbone layers, each launched 3x (for `ncu --set full --import-source on -k regex:gemm_tc`):
stem 3x3 64->64 @300x500, res4 3x3 256->256 dil 2 @74x124, 1x1 1024->256 @9176 rows."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from drn_wsod_pytorch_b200 import ops

def conv(H, W, Cin, Cout, k, dil):
    x = torch.randn(1, H, W, Cin, device="cuda").bfloat16()
    w = (torch.randn(Cout, k * k * Cin, device="cuda") / (k * k * Cin) ** 0.5).bfloat16()
    packed = {"w": w, "scale": None, "bias": torch.zeros(Cout, device="cuda"), "cout": Cout}
    for _ in range(3):
        ops.conv_bf16_tc(x, packed, k, dil, True)
    torch.cuda.synchronize()

conv(300, 500, 64, 64, 3, 1)
conv(74, 124, 256, 256, 3, 2)
conv(1, 9176, 1024, 256, 1, 1)
