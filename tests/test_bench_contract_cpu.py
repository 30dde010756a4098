"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) runs the
unmodified reference (baseline/_ref or /root/reference; the oracle port only when neither exists) on
the host cores and prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "unet_512x512_train_images_per_sec"
    assert d["unit"] == "img/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["value"] - 1e3 / d["ms_per_step"]) < 1e-6 * d["value"] + 1e-9
    from oracle import ref_loader

    expect = "reference" if ref_loader.reference_root() else "port"
    assert d["cpu_baseline"]["kind"] == expect and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "img/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--gpus", "2", "--steps", "1", "--warmup", "1"], capture_output=True,
                       text=True, env=env, timeout=120, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_device():
    import torch

    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


def test_reference_loader_runs_the_unmodified_reference_files():
    """oracle/ref_loader.py executes the reference's own files (never the drop-in shims): the classes
    come from modules whose __file__ lies under baseline/_ref or /root/reference, byte-identical to the
    reference tree when that is mounted, and their seeded init equals the oracle's state dict."""
    import hashlib

    import torch

    from oracle import ref_loader, unet_ref

    ref = ref_loader.load_reference()
    if ref is None:
        return                                   # GPU box without baseline/_ref: nothing to check
    mod = sys.modules[ref.UNet.__module__] if ref.UNet.__module__ in sys.modules else None
    path = ref.UNet.__init__.__code__.co_filename
    assert path.startswith(ref.root) and "unet_segmentation_b200" not in path
    if os.path.isdir("/root/reference") and ref.root != "/root/reference":
        for f in ("models/unet_model.py", "utils/losses.py", "scripts/train.py"):
            a = hashlib.sha256(open(os.path.join(ref.root, f), "rb").read()).hexdigest()
            b = hashlib.sha256(open(os.path.join("/root/reference", f), "rb").read()).hexdigest()
            assert a == b, f
    torch.manual_seed(0)
    m = ref.UNet(1, 2)
    m.apply(ref.init_weights)
    sd = unet_ref.make_state_dict(1, 2, seed=0)
    assert all(torch.equal(sd[k], v) for k, v in m.state_dict().items())
    del mod


def test_forward_flop_model_matches_the_survey_figures():
    """bench.forward_flops (layer-by-layer FLOPs of one eval forward, the `executed_tflops` of the
    inference block) against SURVEY §8d: 223.86 GFLOP per 512^2 image, 2.04 MFLOP per output pixel for
    572 -> 388 tiles at stride 384, 1 359.1 GFLOP per 1084-tile; wide net: 4 745.7 GFLOP at 1024^2."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert abs(bench.forward_flops(512) / 1e9 - 223.860) < 0.01
    assert abs(bench.forward_flops(572) / 384 ** 2 / 1e6 - 2.04) < 0.005
    assert abs(bench.forward_flops(1084) / 1e9 - 1359.1) < 0.1
    assert abs(bench.forward_flops(1024, base=128) / 1e9 - 4745.7) < 0.5
    # training step = 3 x forward minus the first layer's data gradient (2 * 510^2 * 64 * 9 FLOP), SURVEY §8d
    assert abs(3 * bench.forward_flops(512) - 2.0 * 510 ** 2 * 64 * 9 - bench.FLOP_PER_IMG_STEP) < 1e-4 * bench.FLOP_PER_IMG_STEP
