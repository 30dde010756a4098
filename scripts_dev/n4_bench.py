"""Timing of the row-N4 kernels (weight maps, elastic deformation) on one GPU, CUDA events.
    python scripts_dev/n4_bench.py            # batch 16 x 512^2, alpha 2000, sigma 20"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import elastic_ref, weight_map_ref  # noqa: E402  (CPU baseline beside the kernels)
from unet_segmentation_b200 import input_pipeline as ip  # noqa: E402


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    n, h, w = 16, 512, 512
    g = torch.Generator(device="cuda").manual_seed(0)
    img = torch.randint(0, 256, (n, h, w), device="cuda", dtype=torch.uint8, generator=g)
    lab = (torch.randint(0, 40, (n, h, w), device="cuda", generator=g) *
           (torch.rand(n, h, w, device="cuda", generator=g) > 0.5)).to(torch.uint8)
    noise = torch.rand(2, n, h, w, dtype=torch.float64, device="cuda", generator=g)
    out = {}
    ms = timed(lambda: ip.weight_maps_from_labels(lab, dtype=torch.float64))
    out["weight_map_f64"] = {"ms": ms, "GB/s": n * h * w * (2 * 1 + 8) / ms / 1e6, "img/s": n / ms * 1e3}
    ms = timed(lambda: ip.weight_maps_from_labels(lab, dtype=torch.float32))
    out["weight_map_f32"] = {"ms": ms, "GB/s": n * h * w * (2 * 1 + 4) / ms / 1e6, "img/s": n / ms * 1e3}
    ms = timed(lambda: ip.elastic_deform(img, lab, 2000, 20, noise=noise))
    # algorithmic bytes: noise 16 B read, two fields written + read twice (8 B x 2 x 3), pixels 4 B
    out["elastic_deform"] = {"ms": ms, "img/s": n / ms * 1e3,
                             "GB/s_algorithmic": n * h * w * (16 + 48 + 4) / ms / 1e6}
    li, ll = img[0].cpu().numpy(), lab[0].cpu().numpy()
    t = time.time()
    for k in range(3):
        elastic_ref.elastic_deform_scipy(li, ll, 2000, 20, k)
    out["cpu_elastic_scipy_img/s"] = 3 / (time.time() - t)
    t = time.time()
    for k in range(3):
        weight_map_ref.weight_map_full(ll.astype(np.uint16))
    out["cpu_weight_map_full_img/s"] = 3 / (time.time() - t)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
