import time, torch, sys
sys.path.insert(0, '.')
from unet_segmentation_b200.unet import UNet
from unet_segmentation_b200.loss import WeightedCrossEntropyLoss
from oracle import unet_ref
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
m = UNet(1, 2).cuda().train()
crit = WeightedCrossEntropyLoss()
img, t, w = unet_ref.synthetic_batch(N, 512, device='cuda')
for it in range(3):
    loss = crit(m(img), t, w); loss.backward()
torch.cuda.synchronize()
print("arena GB", m.arena_bytes() / 1e9, "loss", float(loss))
e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
K = 5
fw = bw = 0
for it in range(K):
    e0.record(); out = m(img); loss = crit(out, t, w); e1.record(); loss.backward(); e2.record()
    torch.cuda.synchronize(); fw += e0.elapsed_time(e1); bw += e1.elapsed_time(e2)
print(f"N={N} fwd {fw/K:.2f} ms bwd {bw/K:.2f} ms  -> {N/((fw+bw)/K)*1e3:.1f} img/s")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for it in range(2):
        loss = crit(m(img), t, w); loss.backward()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
