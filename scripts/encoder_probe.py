#!/usr/bin/env python
"""GPU probe: per-batch time of the AlexNet hash-head encoder (batch 128 = 1280 crops) and the CLI end to end."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hashgan_b200.encoder import AlexNetHashEncoder, AlexNetWeights
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
conv = sys.argv[2] if len(sys.argv) > 2 else "fp32"
tf32 = conv != "fp32"
enc = AlexNetHashEncoder(AlexNetWeights.synthetic(64, 0), lrn=True, conv=conv)
img = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (B, 3072), dtype=np.uint8)).cuda()
for _ in range(2): out = enc(img)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 5
for _ in range(n): out = enc(img)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
flop = 2 * 10 * B * 720.5e6
print(f"encoder batch {B} conv={conv}: {ms:.2f} ms  -> {B / ms * 1e3:.0f} images/s, {flop / ms / 1e9:.1f} TFLOP/s effective; 54k db ~ {54000 / B * ms / 1e3:.1f} s")
