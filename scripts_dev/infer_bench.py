"""Overlap-tile inference timing for a few tile sizes (dev aid)."""
import sys, time
import torch
sys.path.insert(0, '.')
from oracle import unet_ref
from unet_segmentation_b200 import tiling
from unet_segmentation_b200.unet import UNet

model = UNet(1, 2)
model.load_state_dict(unet_ref.make_state_dict(1, 2, seed=0))
model = model.cuda().eval()
size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
img = (0.4 + 0.2 * torch.rand(512, 512)).repeat(size // 512, size // 512).cuda()
for tile_in, bt in [(572, 8), (700, 8), (1212, 2), (1212, 4), (1212, 8), (1372, 4)]:
    n = len(tiling.plan_tiles(size, size, tile_in)[2])
    try:
        tiling.overlap_tile_predict(model, img, tile_in=tile_in, batch_tiles=bt)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tiling.overlap_tile_predict(model, img, tile_in=tile_in, batch_tiles=bt)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"tile_in {tile_in} batch {bt} tiles {n}: {dt*1e3:.1f} ms  {size*size/dt/1e6:.0f} Mpix/s arena {model.arena_bytes()/2**30:.1f} GiB")
    except Exception as e:
        print(tile_in, bt, "failed:", str(e)[:200])
    model._plans.clear()
