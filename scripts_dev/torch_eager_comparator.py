"""Informal GPU comparator: the reference architecture in STOCK PyTorch on the same
B200 — cuDNN convolutions, torch.autocast(bf16), channels_last, cudnn.benchmark — one training step
(forward + weighted CE + backward + SGD) at BASELINE configs[1] (batch 16 x 512^2), CUDA events.
Not part of the product and not used by bench.py; it answers "what does the reference cost on this
GPU when PyTorch's own kernels run it" next to libunetb200's number.
    python scripts_dev/torch_eager_comparator.py [batch] [steps]"""
import json
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, ".")
from oracle import unet_ref  # noqa: E402  (synthetic batch + reference-keyed init only)


def double_conv(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, 3), nn.BatchNorm2d(cout), nn.ReLU(True),
                         nn.Conv2d(cout, cout, 3), nn.BatchNorm2d(cout), nn.ReLU(True))


class StockUNet(nn.Module):
    def __init__(self):
        super().__init__()
        c = [64, 128, 256, 512, 1024]
        self.enc = nn.ModuleList([double_conv(1, c[0])] + [double_conv(c[i - 1], c[i]) for i in range(1, 5)])
        self.up = nn.ModuleList([nn.ConvTranspose2d(c[4 - j], c[3 - j], 2, 2) for j in range(4)])
        self.dec = nn.ModuleList([double_conv(c[4 - j], c[3 - j]) for j in range(4)])
        self.head = nn.Conv2d(64, 2, 1)

    def forward(self, x):
        feats = []
        for i, e in enumerate(self.enc):
            x = e(x if i == 0 else F.max_pool2d(x, 2))
            feats.append(x)
        for j in range(4):
            u = self.up[j](x)
            s = feats[3 - j]
            dh, dw = (s.shape[2] - u.shape[2]) // 2, (s.shape[3] - u.shape[3]) // 2
            x = self.dec[j](torch.cat([s[:, :, dh:dh + u.shape[2], dw:dw + u.shape[3]], u], 1))
        return self.head(x)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    model = StockUNet().cuda().to(memory_format=torch.channels_last).train()
    opt = torch.optim.SGD(model.parameters(), lr=1e-4, momentum=0.99)
    img, t, w = unet_ref.synthetic_batch(n, 512, device="cuda")
    img = img.contiguous(memory_format=torch.channels_last)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = model(img)
        loss = (F.cross_entropy(logits.float(), t, reduction="none") * w).mean()
        loss.backward()
        opt.step()
        return loss

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"comparator": "stock PyTorch eager: cuDNN + autocast(bf16) + channels_last + cudnn.benchmark",
                      "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
                      "batch": n, "ms_per_step": ms, "img_per_s": n / ms * 1e3, "loss": float(loss),
                      "max_mem_GB": torch.cuda.max_memory_allocated() / 1e9}))


if __name__ == "__main__":
    main()
