"""CPU test of the evidence tooling: scripts_dev/traffic_by_class.py turns an `ncu --nvtx
--print-nvtx-rename kernel --csv` log into the per-class DRAM traffic JSON that bench.py reads for
`roofline.traffic`; the committed JSON must carry every tensor-core class bench.py can name."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HEADER = ('"ID","Process ID","Process Name","Host Name","thread Domain:Push/Pop_Range:PL_Type:PL_Value:'
          'CLR_Type:Color:Msg_Type:Msg","Id:Domain:Start/Stop_Range:PL_Type:PL_Value:CLR_Type:Color:'
          'Msg_Type:Msg","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC",'
          '"Section Name","Metric Name","Metric Unit","Metric Value"')


def _rows(kid, name, read, write, ns):
    base = f'"{kid}","1","python","127.0.0.1","","","{name}","1","7","(256, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics"'
    return [f'{base},"dram__bytes_read.sum","byte","{read}"',
            f'{base},"dram__bytes_write.sum","Kbyte","{write / 1e3}"',
            f'{base},"gpu__time_duration.sum","ns","{ns:,}"']


def test_traffic_by_class_aggregates_the_last_step(tmp_path):
    lines = ["==PROF== Connected to process 1", HEADER]
    kid = 0
    for step in range(2):
        seq = [("first_conv_fp32/void ub::fc1_cov_kernel<4>(...)", 100, 0, 1000),
               ("first_conv_fp32/void ub::fc1_apply_kernel<4>(...)", 400, 2000, 2000),
               ("conv3x3_fprop/void ub::igemm_kmajor_kernel<256, 0, 2>(...)", 1000 * (step + 1), 3000, 5000),
               ("bn_apply_relu_pool/void ub::bn_apply_relu_kernel<0>(...)", 700, 700, 1500),
               ("ub::wce_fwd_bwd_kernel(...)", 10, 10, 100),
               ("first_conv_fp32/void ub::fc1_bwd_kernel<2>(...)", 50, 50, 500),
               ("conv3x3_wgrad/void ub::igemm_wgrad_kernel<64, 1>(...)", 9000, 0, 7000),
               ("conv3x3_wgrad/void ub::wgrad_reduce_kernel<9>(...)", 100, 200, 300)]
        for name, rd, wr, ns in seq:
            lines += _rows(kid, name, rd, wr, ns)
            kid += 1
    log = tmp_path / "traffic.csv"
    log.write_text("\n".join(lines) + "\n")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts_dev", "traffic_by_class.py"), str(log)],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout)
    assert out["kernels_in_step"] == 8                       # only the second step is counted
    assert out["conv3x3_fprop"]["dram_read_bytes_per_step"] == 2000.0
    assert out["conv3x3_fprop"]["dram_bytes_per_step"] == 5000.0
    assert out["conv3x3_wgrad"]["kernels_per_step"] == 2
    assert out["conv3x3_wgrad"]["dram_bytes_per_step"] == 9300.0
    assert out["first_conv_fp32"]["kernels_per_step"] == 3
    assert abs(out["first_conv_fp32"]["ncu_time_us_per_step"] - 3.5) < 1e-9
    assert out["other"]["kernels_per_step"] == 1


def test_committed_traffic_json_covers_the_conv_classes():
    with open(os.path.join(ROOT, "profiles", "r01_traffic_by_class.json")) as fh:
        tj = json.load(fh)
    for cls in ("conv3x3_fprop", "conv3x3_dgrad", "conv3x3_wgrad", "bn_apply_relu_pool", "bn_relu_backward"):
        assert tj[cls]["dram_bytes_per_step"] > 1e9, cls
    assert tj["kernels_in_step"] == 184
