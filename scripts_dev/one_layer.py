"""Run one conv layer's fprop / dgrad / wgrad a few times (for ncu captures).
usage: one_layer.py H Ci Co [N] [reps]"""
import sys
import torch
sys.path.insert(0, '.')
from unet_segmentation_b200 import ops

h, ci, co = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
N = int(sys.argv[4]) if len(sys.argv) > 4 else 16
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
x = torch.randn(N, h, h, ci, device='cuda').to(torch.bfloat16)
w = torch.randn(co, ci, 3, 3, device='cuda') * 0.05
wf, wd = ops.pack_conv3x3(w)
dy = torch.randn(N, h - 2, h - 2, co, device='cuda').to(torch.bfloat16)
for _ in range(reps):
    ops.conv3x3_forward(x, None, wf, None, epilogue=0)
    ops.conv3x3_dgrad(dy, wd)
    ops.conv3x3_wgrad(x, None, dy)
torch.cuda.synchronize()
print("ok")
