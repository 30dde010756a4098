"""Kernel-class breakdown of one eval forward over a batch of overlap tiles (dev aid)."""
import ctypes as C, sys, time
import torch
sys.path.insert(0, '.')
from oracle import unet_ref
from unet_segmentation_b200 import _lib, tiling
from unet_segmentation_b200.unet import UNet

lib = _lib.load()
model = UNet(1, 2)
model.load_state_dict(unet_ref.make_state_dict(1, 2, seed=0))
model = model.cuda().eval()
bt, S = int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 1212
x = (0.4 + 0.2 * torch.rand(bt, 1, S, S)).cuda()
for _ in range(2):
    model.predict_mask(x)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    model.predict_mask(x)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / 3 * 1e3
out = S - 184
print(f"forward batch {bt} x {S}^2: {ms:.2f} ms  -> {bt*out*out/ms/1e3:.0f} Mpix/s (tile outputs only)")
plan = next(iter(model._plans.values()))
lib.ub_plan_profile_enable(plan.handle, 1)
model.predict_mask(x)
torch.cuda.synchronize()
n = lib.ub_plan_profile_classes()
msa = (C.c_double * n)(); fl = (C.c_double * n)(); by = (C.c_double * n)(); ln = (C.c_int * n)()
lib.ub_plan_profile_collect(plan.handle, msa, fl, by, ln)
for c in range(n):
    if ln[c]:
        print(f"  {lib.ub_plan_profile_class_name(c).decode():22s} {msa[c]:8.3f} ms  {ln[c]:3d} groups"
              f"  {fl[c]/msa[c]/1e9 if fl[c] else 0:7.0f} TF/s  {by[c]/msa[c]/1e6:7.0f} GB/s")
img = (0.4 + 0.2 * torch.rand(512, 512)).repeat(16, 16).cuda()
for name, fn in [("extract", lambda: tiling.extract_tiles(img, [(0, 0)] * bt, S, 92))]:
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    print(f"  {name}: {(time.perf_counter()-t0)*1e3:.2f} ms")
