"""Decode what a row-shifted UMMA descriptor reads: identity weights on tap (r=0,s=1)."""
import os, sys, torch
sys.path.insert(0, '.')
from unet_segmentation_b200 import ops
n, h, w, ci, co = 1, 10, 40, 64, 64
wt = torch.zeros(co, ci, 3, 3, device='cuda')
wt[torch.arange(64), torch.arange(64), 0, 1] = 1.0
wf, _ = ops.pack_conv3x3(wt)
Wo = w - 2
for name, x in (("chan", torch.arange(64, device='cuda').float().view(1, 1, 1, 64).expand(n, h, w, 64)),
                ("wcol", torch.arange(w, device='cuda').float().view(1, 1, w, 1).expand(n, h, w, 64)),
                ("hrow", torch.arange(h, device='cuda').float().view(1, h, 1, 1).expand(n, h, w, 64))):
    y, _, _ = ops.conv3x3_forward(x.contiguous().bfloat16(), None, wf, None, epilogue=1)
    out = y.float().reshape(-1, co)
    print(f"--- {name} (UB_DBG_SHIFT={os.environ.get('UB_DBG_SHIFT')})")
    for m in list(range(0, 18)) + [37, 38, 39, 126, 127, 128, 129]:
        if m >= out.shape[0]: continue
        q, p = m % Wo, m // Wo
        row = out[m]
        if name == "chan":
            chunks = [int(row[8 * j].item()) // 8 for j in range(8)]
            ok = all(int(row[c].item()) == c for c in range(64))
            print(f"m={m:3d} (p={p},q={q:2d}) chunk order {chunks} {'OK' if ok else ''}")
        else:
            vals = sorted(set(int(v) for v in row.tolist()))
            exp = q + 1 if name == "wcol" else p
            print(f"m={m:3d} (p={p},q={q:2d}) read {name}={vals} expected {exp}")
