#!/usr/bin/env python
"""Hardware probe (GPU box): does a UMMA smem descriptor whose start address is offset by r*128 B inside a
128B-swizzled tile read rows r.. with the right swizzle phase?  DRN_TC_DEBUG = 8 | shift<<8 [| 16 = set base_offset].
Rows [0, 128-shift) of every 128-row tile must then match the plain GEMM; the last `shift` rows read past the tile."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, torch
sys.path.insert(0, %r)
from drn_wsod_pytorch_b200 import ops
M, K, N = 256, 128, 64
g = torch.Generator().manual_seed(0)
a = torch.randn(M, K, generator=g).bfloat16().cuda()
w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().cuda()
packed = {"w": w, "scale": None, "bias": torch.zeros(N, device="cuda"), "cout": N}
out = ops.conv_bf16_tc(a.view(1, M, 1, K), packed, 1, 1, False, out_dtype=torch.float32).view(M, N)
ref = a.float() @ w.float().t()
err = (out - ref).abs().amax(1)
print("rows_ok", [int((err[t*128:(t+1)*128] < 1e-2).sum()) for t in range(2)], "first_bad", [int((err[t*128:(t+1)*128] >= 1e-2).nonzero()[0]) if (err[t*128:(t+1)*128] >= 1e-2).any() else -1 for t in range(2)])
''' % ROOT
for shift in (1, 2, 3, 8):
    for bo in (0, 16):
        env = dict(os.environ, DRN_TC_DEBUG=str(8 | bo | (shift << 8)), DRN_TC_CTA_GROUP="1")
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
        print(f"shift={shift} base_offset={'set' if bo else '0'}:", r.stdout.strip() or r.stderr.strip()[-300:])
