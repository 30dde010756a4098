"""Per-layer timing of the tcgen05 conv kernels at BASELINE config 2 shapes (N=16, 512^2)."""
import sys
import torch
sys.path.insert(0, '.')
from unet_segmentation_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
# (name, Hin, Cin (total), Cout)   conv output is Hin-2
LAYERS = [("inc.b", 510, 64, 64), ("d1.a", 254, 64, 128), ("d1.b", 252, 128, 128),
          ("d2.a", 125, 128, 256), ("d2.b", 123, 256, 256), ("d3.a", 60, 256, 512),
          ("d3.b", 58, 512, 512), ("d4.a", 28, 512, 1024), ("d4.b", 26, 1024, 1024),
          ("up1.a", 48, 1024, 512), ("up1.b", 46, 512, 512), ("up2.a", 88, 512, 256),
          ("up2.b", 86, 256, 256), ("up3.a", 168, 256, 128), ("up3.b", 166, 128, 128),
          ("up4.a", 328, 128, 64), ("up4.b", 326, 64, 64)]


def timeit(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


tot = {"fprop": 0.0, "dgrad": 0.0, "wgrad": 0.0}
print(f"{'layer':8s} {'M':>9s} {'Ci':>5s} {'Co':>5s} {'GFLOP':>8s} | fprop ms TF/s | dgrad ms TF/s | wgrad ms TF/s")
for name, h, ci, co in LAYERS:
    x = torch.randn(N, h, h, ci, device='cuda').to(torch.bfloat16)
    w = torch.randn(co, ci, 3, 3, device='cuda') * 0.05
    wf, wd = ops.pack_conv3x3(w)
    dy = torch.randn(N, h - 2, h - 2, co, device='cuda').to(torch.bfloat16)
    gf = 2.0 * N * (h - 2) ** 2 * co * 9 * ci / 1e9
    t_f = timeit(lambda: ops.conv3x3_forward(x, None, wf, None, epilogue=0))
    t_d = timeit(lambda: ops.conv3x3_dgrad(dy, wd))
    t_w = timeit(lambda: ops.conv3x3_wgrad(x, None, dy))
    tot["fprop"] += t_f; tot["dgrad"] += t_d; tot["wgrad"] += t_w
    print(f"{name:8s} {N*(h-2)**2:9d} {ci:5d} {co:5d} {gf:8.1f} | {t_f:7.3f} {gf/t_f:6.0f} | {t_d:7.3f} {gf/t_d:6.0f} | {t_w:7.3f} {gf/t_w:6.0f}")
    del x, w, wf, wd, dy
print("totals ms:", tot)
