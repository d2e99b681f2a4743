#!/usr/bin/env python
"""Time one 3x3 conv through drn_conv_igemm_bf16_tc (graph-replayed) -- used with DRN_TC_DEBUG ablations.
    python tools/conv_probe.py H W Cin Cout dil"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from drn_wsod_pytorch_b200 import ops
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gemm_sweep import timed

H, W, Cin, Cout, dil = [int(x) for x in sys.argv[1:6]]
x = torch.randn(1, H, W, Cin, device="cuda").bfloat16()
w = (torch.randn(Cout, 9 * Cin, device="cuda") / (9 * Cin) ** 0.5).bfloat16()
packed = {"w": w, "scale": None, "bias": torch.zeros(Cout, device="cuda"), "cout": Cout}
us = timed(lambda: ops.conv_bf16_tc(x, packed, 3, dil, True))
print(f"conv3x3 {H}x{W} {Cin}->{Cout} dil{dil} DEBUG={os.environ.get('DRN_TC_DEBUG', '0')}: {us:.1f} us")
