"""CPU tests: the oracle restatements against the reference's golden vectors (and against the live
reference when /root/reference is mounted), and the C-ABI surface. No GPU needed."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"

from oracle import ccl_ref, elastic_ref, unet_ref, weight_map_ref  # noqa: E402


def _case(blob, name):
    return {k[len(name) + 1:]: blob[k] for k in blob.files if k.startswith(name + "/")}


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLD, "unet_golden.npz"))


@pytest.mark.parametrize("name", ["train_n2_s188", "train_n1_s220", "eval_n1_s252",
                                  "train_n1_s220_bilinear"])
def test_unet_oracle_matches_reference_golden(golden, name):
    bilinear = name.endswith("_bilinear")
    if bilinear:
        golden = np.load(os.path.join(GOLD, "unet_bilinear_golden.npz"))
    c = _case(golden, name)
    n, size, sw, sx, training = [int(v) for v in c["meta"]]
    sd = unet_ref.make_state_dict(1, 2, seed=sw, bilinear=bilinear)
    img, t, w = unet_ref.synthetic_batch(n, size=size, seed=sx)
    if training:
        params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()
                  and "running" not in k}
        full = dict(sd)
        full.update(params)
        bufs = {}
        logits = unet_ref.unet_forward(full, img, training=True, buffers_out=bufs)
        loss = unet_ref.weighted_cross_entropy(logits, t, w)
        loss.backward()
        assert abs(float(loss) - float(c["loss"])) <= 1e-6 * abs(float(c["loss"]))
        for k, p in params.items():
            g = p.grad.flatten()
            ref_norm = float(c[f"gnorm/{k}"])
            assert abs(float(g.double().norm()) - ref_norm) <= 1e-4 * ref_norm + 1e-9, k
            np.testing.assert_allclose(g[torch.from_numpy(c[f"gidx/{k}"])].numpy(), c[f"gval/{k}"],
                                       rtol=1e-3, atol=1e-5 * ref_norm + 1e-9)
        for k in [k for k in c if k.startswith("buf/")]:
            np.testing.assert_allclose(bufs[k[4:]].numpy(), c[k], rtol=1e-5, atol=1e-6)
    else:
        g = torch.Generator().manual_seed(99)
        for k in [k for k in sd if k.endswith("running_mean")]:
            nf = sd[k].numel()
            sd[k] = torch.randn(nf, generator=g) * 0.1
            sd[k.replace("running_mean", "running_var")] = 0.5 + torch.rand(nf, generator=g)
        with torch.no_grad():
            logits = unet_ref.unet_forward(sd, img, training=False)
    np.testing.assert_allclose(logits.detach().numpy(), c["logits"], rtol=1e-4, atol=1e-5)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_unet_oracle_matches_live_reference():
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location("ref_unet_model_live",
                                                  os.path.join(REF, "models", "unet_model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(5)
    model = mod.UNet(1, 2).train()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    x = torch.rand(1, 1, 204, 204)
    bufs = {}
    mine = unet_ref.unet_forward(sd, x, training=True, buffers_out=bufs)
    ref = model(x)
    assert torch.allclose(mine, ref, rtol=1e-5, atol=1e-6)
    new = model.state_dict()
    for k, v in bufs.items():
        assert torch.allclose(v.float(), new[k].float(), rtol=1e-5, atol=1e-6), k
    model.eval()
    with torch.no_grad():
        assert torch.allclose(unet_ref.unet_forward(new, x, training=False), model(x), rtol=1e-5,
                              atol=1e-6)
    assert unet_ref.out_size(512) == 324 and unet_ref.out_size(572) == 388


def test_drop_in_module_tree_matches_reference_state_dict():
    """Same keys, shapes and dtypes as the reference module tree and same seeded init."""
    from unet_segmentation_b200.unet import UNet

    torch.manual_seed(0)
    m = UNet(1, 2)
    ours = m.state_dict()
    ref = unet_ref.make_state_dict(1, 2, seed=0, init=False)
    assert list(ours.keys()) == list(ref.keys())
    assert len(list(m.parameters())) == 82 and len(list(m.buffers())) == 54
    for k in ref:
        assert ours[k].shape == ref[k].shape and ours[k].dtype == ref[k].dtype, k
        assert torch.equal(ours[k], ref[k]), k
    assert sum(p.numel() for p in m.parameters()) == 31_042_434
    # canonical order handed to the C library == named_parameters order
    assert [id(p) for p in m._ordered_params()] == [id(p) for p in m.parameters()]
    bn_names = [n for n, mod in m.named_modules() if isinstance(mod, torch.nn.BatchNorm2d)]
    assert [id(b) for b in m._ordered_bns()] == [id(dict(m.named_modules())[n]) for n in bn_names]


def test_drop_in_rejects_cpu():
    from unet_segmentation_b200.loss import WeightedCrossEntropyLoss
    from unet_segmentation_b200.unet import UNet

    for m in (UNet(1, 2), UNet(1, 2, bilinear=True)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m(torch.zeros(1, 1, 188, 188))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        WeightedCrossEntropyLoss()(torch.zeros(1, 2, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long),
                                   torch.ones(1, 4, 4))


def test_drop_in_bilinear_module_tree_matches_reference_state_dict():
    """UNet(..., bilinear=True) (models/unet_model.py:40-43,78-81): nn.Upsample up-sampling without
    parameters, first decoder convolutions over prev + skip channels (1536 -> 512 ...)."""
    from unet_segmentation_b200.unet import UNet

    torch.manual_seed(0)
    m = UNet(1, 2, bilinear=True)
    ours = m.state_dict()
    ref = unet_ref.make_state_dict(1, 2, seed=0, init=False, bilinear=True)
    assert list(ours.keys()) == list(ref.keys())
    assert len(list(m.parameters())) == 74
    for k in ref:
        assert ours[k].shape == ref[k].shape and torch.equal(ours[k], ref[k]), k
    assert tuple(ours["up1.conv.double_conv.0.weight"].shape) == (512, 1536, 3, 3)
    assert isinstance(m.up1.up, torch.nn.Upsample) and m.up1.up.align_corners is True
    assert [id(p) for p in m._ordered_params()] == [id(p) for p in m.parameters()]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_bilinear_oracle_matches_live_reference():
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location("ref_unet_bilinear",
                                                  os.path.join(REF, "models", "unet_model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(5)
    ref = mod.UNet(1, 2, bilinear=True)
    sd = unet_ref.make_state_dict(1, 2, seed=5, init=False, bilinear=True)
    rsd = ref.state_dict()
    assert list(sd) == list(rsd) and all(torch.equal(sd[k], rsd[k]) for k in sd)
    img, t, w = unet_ref.synthetic_batch(1, size=220, seed=3)
    ref.train()
    logits = ref(img)
    loss = (torch.nn.functional.cross_entropy(logits, t, reduction="none") * w).mean()
    loss.backward()
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and "running" not in k}
    full = dict(sd)
    full.update(params)
    o_logits = unet_ref.unet_forward(full, img, training=True, buffers_out={})
    o_loss = unet_ref.weighted_cross_entropy(o_logits, t, w)
    o_loss.backward()
    assert torch.allclose(logits, o_logits, atol=1e-5, rtol=1e-5)
    assert abs(float(loss) - float(o_loss)) <= 1e-6 * abs(float(loss))
    for name, p in ref.named_parameters():
        g, og = p.grad, params[name].grad
        assert float((g - og).norm()) <= 1e-4 * float(g.norm()) + 1e-7, name


def test_ccl_oracle_matches_reference_golden_pairs():
    blob = np.load(os.path.join(GOLD, "ccl_golden.npz"))
    ids = sorted(k[4:] for k in blob.files if k.startswith("mask"))
    assert len(ids) >= 4
    for i in ids:
        out = ccl_ref.get_instance_masks(blob[f"mask{i}"], min_size=15)
        assert out.dtype == np.uint16
        assert np.array_equal(out, blob[f"inst{i}"]), i


def test_ccl_pure_python_agrees_on_small_cases():
    rng = np.random.default_rng(3)
    for shape, p in [((9, 13), 0.5), ((16, 16), 0.7), ((1, 30), 0.6), ((12, 12), 0.0)]:
        m = (rng.random(shape) < p).astype(np.uint8) * 255
        for ms in (1, 4, 15):
            assert np.array_equal(ccl_ref.get_instance_masks(m, ms),
                                  ccl_ref.get_instance_masks_pure(m, ms))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_ccl_oracle_matches_all_shipped_pairs():
    from PIL import Image

    base = os.path.join(REF, "data/raw/processed/predictions/DIC-C2DH-HeLa")
    n = 0
    for i in range(0, 84, 7):
        m = np.array(Image.open(os.path.join(base, "01_RES", f"mask{i:03d}.tif")))
        inst = np.array(Image.open(os.path.join(base, "01_RES_INST", f"m{i:03d}.tif")))
        assert np.array_equal(ccl_ref.get_instance_masks(m, 15), inst.astype(np.uint16))
        n += 1
    assert n == 12


def test_weight_map_oracle_matches_reference_stored_maps():
    """calculate_weight_map (scripts/preprocess_data.py:17-77): the step-by-step restatement (real
    distance transforms) and the closed form the CUDA kernel implements both reproduce the maps the
    reference stores, bit for bit."""
    blob = np.load(os.path.join(GOLD, "weight_map_golden.npz"))
    ids = sorted(k[len("labels"):] for k in blob.files if k.startswith("labels"))
    assert len(ids) >= 3
    for i in ids:
        lab, stored = blob[f"labels{i}"], blob[f"wmap{i}"]
        assert stored.dtype == np.float64 and len(np.unique(stored)) == 2
        closed = weight_map_ref.weight_map_closed_form(lab)
        assert closed.dtype == np.float64 and np.array_equal(closed, stored), i
    full = weight_map_ref.weight_map_full(blob[f"labels{ids[0]}"])
    assert full.dtype == np.float64 and np.array_equal(full, blob[f"wmap{ids[0]}"])


def test_weight_map_closed_form_equals_full_restatement_on_edge_cases():
    rng = np.random.default_rng(5)
    cases = [np.zeros((12, 20), np.uint16), np.full((9, 7), 3, np.uint8)]
    one = np.zeros((24, 31), np.uint16); one[5:14, 8:20] = 7; cases.append(one)
    three = one.copy(); three[16:22, 2:30] = 2; three[0:3, 0:4] = 300; cases.append(three)
    cases.append((rng.integers(0, 6, (33, 17)) * (rng.random((33, 17)) < 0.4)).astype(np.uint16))
    for lab in cases:
        for w0, sigma in ((10, 5), (3.5, 0.0)):
            full = weight_map_ref.weight_map_full(lab, w0, sigma)
            closed = weight_map_ref.weight_map_closed_form(lab, w0, sigma)
            assert np.array_equal(full.astype(np.float64), closed), (lab.shape, w0, sigma)
    wbg, wfg = weight_map_ref.class_balance_weights(0, 100)
    assert (wbg, wfg) == (np.float32(1.0), np.float32(0.0))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_weight_map_oracle_matches_live_reference_and_all_stored_maps():
    import glob

    from PIL import Image

    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location(
        "ref_preprocess", os.path.join(REF, "scripts", "preprocess_data.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    lab = np.zeros((40, 56), np.uint16)
    for k, (y, x, h, w) in enumerate([(3, 4, 10, 12), (20, 30, 15, 20), (25, 2, 8, 9)]):
        lab[y:y + h, x:x + w] = 5 * k + 1
        got = ref.calculate_weight_map(lab, 10, 5)
        assert np.array_equal(got, weight_map_ref.weight_map_full(lab))
        assert np.array_equal(got, weight_map_ref.weight_map_closed_form(lab))
    empty = ref.calculate_weight_map(np.zeros((8, 8), np.uint16), 10, 5)
    assert empty.dtype == np.float32 == weight_map_ref.weight_map_full(np.zeros((8, 8), np.uint16)).dtype
    base = os.path.join(REF, "data/raw/train/DIC-C2DH-HeLa/01_ST")
    files = sorted(glob.glob(os.path.join(base, "WEIGHT_MAPS", "weight_map_*.npy")))
    assert len(files) == 84
    for f in files:
        num = os.path.basename(f)[len("weight_map_"):-4]
        m = np.array(Image.open(os.path.join(base, "SEG", f"man_seg{num}.tif")))
        assert np.array_equal(weight_map_ref.weight_map_closed_form(m), np.load(f)), f


_ELASTIC_CASES = [((96, 80), 2000, 20, 123), ((64, 48), 300, 3, 7), ((33, 17), 2000, 2, 9),
                  ((5, 7), 50, 0.8, 1), ((1, 9), 20, 1, 2)]


def _elastic_inputs(shape, seed):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, shape).astype(np.uint8)
    lab = (rng.integers(0, 60000, shape) * (rng.random(shape) < 0.5)).astype(np.uint16)
    return img, lab


@pytest.mark.parametrize("shape,alpha,sigma,seed", _ELASTIC_CASES)
def test_elastic_step_by_step_oracle_equals_scipy_restatement(shape, alpha, sigma, seed):
    """The plain-numpy arithmetic the CUDA kernels mirror is bit-identical to the reference's scipy
    calls (utils/augmentations.py:27-37), incl. displacements of several image sizes (repeated
    reflection) and one-pixel-high images."""
    img, lab = _elastic_inputs(shape, seed)
    si, sm = elastic_ref.elastic_deform_scipy(img, lab, alpha, sigma, seed)
    u, v = elastic_ref.reference_noise(seed, shape)
    ti, tm = elastic_ref.elastic_deform_steps(img, lab, u, v, alpha, sigma)
    assert si.dtype == ti.dtype == np.uint8 and sm.dtype == tm.dtype == np.uint16
    assert np.array_equal(si, ti) and np.array_equal(sm, tm)
    assert not np.array_equal(si, img)                      # the deformation does something


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_elastic_oracle_matches_live_reference():
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location(
        "ref_augmentations", os.path.join(REF, "utils", "augmentations.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    for shape, alpha, sigma, seed in _ELASTIC_CASES + [((512, 512), 2000, 20, 2024)]:
        img, lab = _elastic_inputs(shape, seed)
        ri, rm = ref.elastic_deform_image_and_mask(img, lab, alpha, sigma, random_state=seed)
        si, sm = elastic_ref.elastic_deform_scipy(img, lab, alpha, sigma, seed)
        u, v = elastic_ref.reference_noise(seed, shape)
        ti, tm = elastic_ref.elastic_deform_steps(img, lab, u, v, alpha, sigma)
        assert np.array_equal(ri, si) and np.array_equal(rm, sm), shape
        assert np.array_equal(ri, ti) and np.array_equal(rm, tm), shape


def test_c_abi_exports_every_declared_symbol():
    from unet_segmentation_b200 import _lib

    lib = _lib.load()
    declared = _lib.declared_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), f"libunetb200.so does not export {name}"
    assert set(declared) == set(_lib._SIGNATURES)
    assert lib.ub_version() >= 100


def test_host_side_helpers_of_the_elastic_pipeline():
    """The two host-side pieces of input_pipeline.elastic_deform: the Gaussian taps are scipy's own
    kernel (impulse response of gaussian_filter1d), and reference_noise replays the reference's
    RandomState draws (dx first, then dy; utils/augmentations.py:27-28)."""
    from scipy.ndimage import gaussian_filter1d

    from unet_segmentation_b200 import input_pipeline as ip

    for sigma in (20, 5, 2.5, 0.8):
        taps = ip.gaussian_taps(sigma)
        r = (len(taps) - 1) // 2
        assert r == int(4.0 * sigma + 0.5)
        impulse = np.zeros(2 * r + 1)
        impulse[r] = 1.0
        assert np.array_equal(gaussian_filter1d(impulse, sigma, mode="constant"), taps)
        assert np.array_equal(taps, elastic_ref.gaussian_taps(sigma))
    with pytest.raises(ValueError):
        ip.gaussian_taps(0.0)
    noise = ip.reference_noise([11, 12], (6, 5))
    assert noise.shape == (2, 2, 6, 5) and noise.dtype == torch.float64
    for k, seed in enumerate((11, 12)):
        u, v = elastic_ref.reference_noise(seed, (6, 5))
        assert np.array_equal(noise[0, k].numpy(), u) and np.array_equal(noise[1, k].numpy(), v)
