"""Diagnostic: bilinear=True UNet vs the fp32 oracle in eval mode and at several training sizes."""
import sys
import torch
sys.path.insert(0, '.')
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from oracle import unet_ref
from unet_segmentation_b200.unet import UNet


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


for bilinear in (False, True):
    sd = unet_ref.make_state_dict(1, 2, seed=2, bilinear=bilinear)
    m = UNet(1, 2, bilinear)
    m.load_state_dict(sd)
    m = m.cuda()
    sdc = {k: v.cuda() for k, v in sd.items()}
    for n, size in [(2, 252), (2, 380), (2, 508)]:
        img, t, w = unet_ref.synthetic_batch(n, size=size, seed=31, device="cuda")
        m.train()
        with torch.no_grad():
            ref = unet_ref.unet_forward(dict(sdc), img, training=True, buffers_out={})
        out = m(img)
        print(f"bilinear={bilinear} train {n}x{size}: logits rel-L2 {rel(out, ref):.3e}")
    g = torch.Generator().manual_seed(7)
    for k in [k for k in sdc if k.endswith("running_mean")]:
        nf = sdc[k].numel()
        sdc[k] = (torch.randn(nf, generator=g) * 0.1).cuda()
        sdc[k.replace("running_mean", "running_var")] = (0.5 + torch.rand(nf, generator=g)).cuda()
    m.load_state_dict(sdc)
    m.eval()
    img, _, _ = unet_ref.synthetic_batch(1, size=316, seed=5, device="cuda")
    with torch.no_grad():
        ref = unet_ref.unet_forward(sdc, img, training=False)
        out = m(img)
    print(f"bilinear={bilinear} eval 1x316: logits rel-L2 {rel(out, ref):.3e}")
