"""Eager vs CUDA-graph replay of the bench.py training step (N=16 x 512^2, FusedSGD): what do the ~185
launches per step cost when they are issued one by one?   python scripts_dev/graph_step_bench.py"""
import sys
import torch
sys.path.insert(0, '.')
from oracle import unet_ref
from unet_segmentation_b200.loss import WeightedCrossEntropyLoss
from unet_segmentation_b200.optim import FusedSGD
from unet_segmentation_b200.unet import UNet

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
model = UNet(1, 2)
model.load_state_dict(unet_ref.make_state_dict(1, 2, seed=0))
model = model.cuda().train()
crit = WeightedCrossEntropyLoss()
opt = FusedSGD(model, lr=1e-4, momentum=0.99)
img, t, w = unet_ref.synthetic_batch(N, 512, device='cuda')


def step():
    opt.zero_grad(set_to_none=True)
    loss = crit(model(img), t, w)
    loss.backward()
    opt.step()
    return loss


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for _ in range(3):
    step()
eager = timeit(step)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    loss_s = step()
graphed = timeit(g.replay)
eager2 = timeit(step)
print(f"N={N}: eager {eager:.3f} ms/step, graph replay {graphed:.3f} ms/step, eager again {eager2:.3f}; loss {float(loss_s):.4f}")
