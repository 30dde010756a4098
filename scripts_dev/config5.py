"""BASELINE config 5: wide UNet (base 128), 1024^2 crops, batch 8 — memory / correctness / timing probe."""
import sys, time, torch
sys.path.insert(0, '.')
from oracle import unet_ref
from unet_segmentation_b200.loss import WeightedCrossEntropyLoss
from unet_segmentation_b200.unet import UNet
N, S = int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 1024
torch.manual_seed(0)
m = UNet(1, 2, base_channels=128).cuda().train()
crit = WeightedCrossEntropyLoss()
opt = torch.optim.SGD(m.parameters(), lr=1e-4, momentum=0.99)
img, t, w = unet_ref.synthetic_batch(N, S, device='cuda')
print("params", sum(p.numel() for p in m.parameters()), "out", unet_ref.out_size(S))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(4):
    if it == 2:
        torch.cuda.synchronize(); e0.record()
    opt.zero_grad(set_to_none=True)
    loss = crit(m(img), t, w); loss.backward(); opt.step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 2
flop = 14234.6e9 * N * (S / 1024) ** 2
print(f"loss {float(loss.detach()):.4f} arena {m.arena_bytes()/1e9:.1f} GB  {ms:.1f} ms/step  {N/ms*1e3:.1f} img/s  ~{flop/ms/1e9:.0f} TFLOP/s  max mem {torch.cuda.max_memory_allocated()/1e9:.1f} GB (torch)")
assert torch.isfinite(loss)
