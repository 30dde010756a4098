"""Generate tests/golden/*.npz from the REAL reference (run in the authoring container only).

    python oracle/make_golden.py

Imports /root/reference (read-only, bytecode writing disabled) and records, for seeded weights and
inputs, what the reference's own UNet + WeightedCrossEntropyLoss compute on CPU in fp32, plus a
subset of the reference's shipped mask -> instance-label pairs for get_instance_masks and of its
instance mask -> stored weight map pairs for calculate_weight_map.
The GPU box has no /root/reference: tests there compare against these fixtures.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_modules():
    unet_mod = _load("ref_unet_model", os.path.join(REF, "models", "unet_model.py"))
    loss_mod = _load("ref_losses", os.path.join(REF, "utils", "losses.py"))
    sys.path.insert(0, REF)
    try:
        train_mod = _load("ref_train", os.path.join(REF, "scripts", "train.py"))
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k.split(".")[0] in ("models", "utils")]:
            del sys.modules[k]
    return unet_mod.UNet, loss_mod.WeightedCrossEntropyLoss, train_mod


def sample_idx(numel, k=64):
    return np.unique(np.linspace(0, numel - 1, num=min(k, numel)).astype(np.int64))


def unet_case(UNet, Loss, train_mod, n, size, seed_w, seed_x, training, bilinear=False):
    sys.path.insert(0, ROOT)
    from oracle.unet_ref import synthetic_batch

    torch.manual_seed(seed_w)
    model = UNet(n_channels=1, n_classes=2, bilinear=bilinear)
    model.apply(train_mod.init_weights)
    img, t, w = synthetic_batch(n, size=size, seed=seed_x)
    out = {}
    if training:
        model.train()
        logits = model(img)
        loss = Loss()(logits, t, w)
        loss.backward()
        out["loss"] = np.float64(loss.item())
        for name, p in model.named_parameters():
            g = p.grad.detach().flatten()
            out[f"gnorm/{name}"] = np.float64(g.double().norm().item())
            idx = sample_idx(g.numel())
            out[f"gidx/{name}"] = idx
            out[f"gval/{name}"] = g[idx].numpy()
        for name, b in model.named_buffers():
            if name.startswith(("inc.", "up4.")):
                out[f"buf/{name}"] = b.detach().numpy()
    else:
        # non-trivial running statistics, as after some training
        g = torch.Generator().manual_seed(99)
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=g))
        model.eval()
        with torch.no_grad():
            logits = model(img)
    out["logits"] = logits.detach().numpy()
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    UNet, Loss, train_mod = reference_modules()
    blob = {}
    cases = {"train_n2_s188": (2, 188, 0, 11, True), "train_n1_s220": (1, 220, 3, 12, True),
             "eval_n1_s252": (1, 252, 0, 13, False)}
    for name, (n, size, sw, sx, tr) in cases.items():
        for k, v in unet_case(UNet, Loss, train_mod, n, size, sw, sx, tr).items():
            blob[f"{name}/{k}"] = v
        blob[f"{name}/meta"] = np.array([n, size, sw, sx, int(tr)])
    np.savez_compressed(os.path.join(OUT, "unet_golden.npz"), **blob)
    bil = {}     # UNet(1, 2, bilinear=True): nn.Upsample branch of the reference's Up block
    for k, v in unet_case(UNet, Loss, train_mod, 1, 220, 3, 12, True, bilinear=True).items():
        bil[f"train_n1_s220_bilinear/{k}"] = v
    bil["train_n1_s220_bilinear/meta"] = np.array([1, 220, 3, 12, 1])
    np.savez_compressed(os.path.join(OUT, "unet_bilinear_golden.npz"), **bil)

    # the real operating point (BASELINE configs[0]: one 1x1x512x512 image -> 324x324 logits): the
    # reference-held vector that is checked at the north-star tolerance itself (2e-2 / 99.9 %)
    big = {}
    for name, (n, size, sw, sx, tr) in {"train_n1_s512": (1, 512, 0, 1234, True),
                                        "eval_n1_s512": (1, 512, 0, 77, False)}.items():
        for k, v in unet_case(UNet, Loss, train_mod, n, size, sw, sx, tr).items():
            big[f"{name}/{k}"] = v
        big[f"{name}/meta"] = np.array([n, size, sw, sx, int(tr)])
    np.savez_compressed(os.path.join(OUT, "unet_golden_512.npz"), **big)

    from PIL import Image

    base = os.path.join(REF, "data/raw/processed/predictions/DIC-C2DH-HeLa")
    ccl = {}
    for i in (0, 17, 41, 83):
        m = np.array(Image.open(os.path.join(base, "01_RES", f"mask{i:03d}.tif")))
        inst = np.array(Image.open(os.path.join(base, "01_RES_INST", f"m{i:03d}.tif")))
        ccl[f"mask{i:03d}"] = m.astype(np.uint8)
        ccl[f"inst{i:03d}"] = inst.astype(np.uint16)
    np.savez_compressed(os.path.join(OUT, "ccl_golden.npz"), **ccl)

    # calculate_weight_map (scripts/preprocess_data.py): instance masks + the maps the reference stores
    train = os.path.join(REF, "data/raw/train/DIC-C2DH-HeLa")
    wm = {}
    for seq, num in (("01", "002"), ("01", "041"), ("01", "067")):
        wm[f"labels{seq}_{num}"] = np.array(
            Image.open(os.path.join(train, f"{seq}_ST", "SEG", f"man_seg{num}.tif"))).astype(np.uint16)
        wm[f"wmap{seq}_{num}"] = np.load(
            os.path.join(train, f"{seq}_ST", "WEIGHT_MAPS", f"weight_map_{num}.npy"))
    np.savez_compressed(os.path.join(OUT, "weight_map_golden.npz"), **wm)
    for f in os.listdir(OUT):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
