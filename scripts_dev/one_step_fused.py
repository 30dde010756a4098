"""Training steps at BASELINE config 2 with the fused optimizer, exactly the bench.py step (for ncu launch lists)."""
import sys
import torch
sys.path.insert(0, '.')
from oracle import unet_ref
from unet_segmentation_b200.loss import WeightedCrossEntropyLoss
from unet_segmentation_b200.unet import UNet

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
model = UNet(1, 2)
model.load_state_dict(unet_ref.make_state_dict(1, 2, seed=0))
model = model.cuda().train()
crit = WeightedCrossEntropyLoss()
from unet_segmentation_b200.optim import FusedSGD
opt = FusedSGD(model, lr=1e-4, momentum=0.99)
img, t, w = unet_ref.synthetic_batch(N, 512, device='cuda')
for _ in range(steps):
    opt.zero_grad(set_to_none=True)
    loss = crit(model(img), t, w)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("loss", float(loss.detach()))
