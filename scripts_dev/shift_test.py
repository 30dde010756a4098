"""Experiment: does a UMMA smem descriptor whose start address is shifted by one 128-byte row
(inside a 1024-byte swizzle atom) read the rows TMA wrote there?  UB_DBG_SHIFT=1|2."""
import os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, '.')
torch.backends.cudnn.allow_tf32 = False
from unet_segmentation_b200 import ops
n, h, w, ci, co = 2, 37, 29, 64, 128
g = torch.Generator(device='cuda').manual_seed(0)
x = torch.randn(n, ci, h, w, device='cuda', generator=g).bfloat16().float()
wt = (torch.randn(co, ci, 3, 3, device='cuda', generator=g) * 0.05).bfloat16().float()
wf, _ = ops.pack_conv3x3(wt)
y, _, _ = ops.conv3x3_forward(ops.nhwc(x), None, wf, None, epilogue=1)
ref = F.conv2d(x, wt).permute(0, 2, 3, 1).reshape(-1, co)
out = y.float().reshape(-1, co)
rows = torch.arange(out.shape[0], device='cuda')
keep = (rows % 128) != 127          # the last row of every tile reads past the A stage
err = ((out - ref)[keep].norm() / ref[keep].norm()).item()
err_last = ((out - ref)[~keep].norm() / ref[~keep].norm()).item()
print(f"UB_DBG_SHIFT={os.environ.get('UB_DBG_SHIFT')} rel-L2 rows!=127: {err:.3e}   rows==127: {err_last:.3e}")
