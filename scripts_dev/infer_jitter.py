"""Per-call device + host time of overlap_tile_predict on a 1024^2 image (1 GPU): looks for outliers
(graph re-capture, allocator, host stalls) behind the run-to-run spread of bench.py's 1024^2 figure."""
import sys
import time

import torch

sys.path.insert(0, ".")
from oracle import unet_ref
from unet_segmentation_b200 import tiling
from unet_segmentation_b200.unet import UNet

dev = torch.device("cuda", 0)
model = UNet(1, 2)
model.load_state_dict(unet_ref.make_state_dict(1, 2, seed=0))
model = model.to(dev).eval()
size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
img = (0.4 + 0.2 * torch.rand(512, 512)).repeat(size // 512, size // 512).to(dev)
tile_in, ranks_used, bt = tiling.choose_plan(size, size, 1)
for _ in range(3):
    tiling.overlap_tile_predict(model, img, tile_in=tile_in, batch_tiles=bt)
torch.cuda.synchronize()
dev_ms, host_ms = [], []
for _ in range(30):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    tiling.overlap_tile_predict(model, img, tile_in=tile_in, batch_tiles=bt)
    e1.record()
    host_ms.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
    dev_ms.append(e0.elapsed_time(e1))
plans = [p for p in model._plans.values()]
print("tile_in", tile_in, "bt", bt, "graph replays", [p.graph_replays() for p in plans])
print("device ms:", " ".join(f"{x:.2f}" for x in dev_ms))
print("host   ms:", " ".join(f"{x:.2f}" for x in host_ms))
# back-to-back (no sync between calls), as bench.py times it
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(20):
    tiling.overlap_tile_predict(model, img, tile_in=tile_in, batch_tiles=bt)
e1.record()
host = (time.perf_counter() - t0) * 1e3 / 20
torch.cuda.synchronize()
print(f"back-to-back: device {e0.elapsed_time(e1) / 20:.3f} ms/call, host enqueue {host:.3f} ms/call")
