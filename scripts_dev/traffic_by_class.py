"""Per-kernel-class DRAM traffic and time of ONE training step from an ncu CSV taken with

    UB_NVTX=1 ncu --nvtx --print-nvtx-rename kernel --clock-control none --csv \
        --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --log-file gpurun_out/traffic.csv python scripts_dev/one_step_fused.py 16 2

(the library names each kernel class with an NVTX range under UB_NVTX=1, so ncu reports the class
instead of the function name). Writes the JSON bench.py reads for roofline.traffic.
    python scripts_dev/traffic_by_class.py gpurun_out/traffic.csv > profiles/r01_traffic_by_class.json
"""
import collections
import csv
import json
import sys

CLASSES = ["conv3x3_fprop", "conv3x3_dgrad", "conv3x3_wgrad", "convT_fprop", "convT_dgrad",
           "convT_wgrad", "bn_apply_relu_pool", "bn_relu_backward", "first_conv_fp32", "head_1x1"]


def main(path):
    lines = open(path).read().splitlines()
    start = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
    per_id = collections.OrderedDict()
    for r in csv.DictReader(lines[start:]):
        rec = per_id.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "")
        if r["Metric Name"].startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        elif r["Metric Name"].startswith("gpu__time"):
            v *= {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(unit, 1e-3)
        rec[r["Metric Name"]] = v
    recs = list(per_id.values())
    names = [r["name"] for r in recs]
    # one step = [forward first_conv group, next forward first_conv group): groups alternate fwd, bwd
    groups = []
    for k, n in enumerate(names):
        if "first_conv_fp32" in n and (k == 0 or "first_conv_fp32" not in names[k - 1]):
            groups.append(k)
    if len(groups) < 2:
        raise SystemExit("no NVTX-renamed kernels found: was the run made with UB_NVTX=1 and --nvtx?")
    begin = groups[-2]
    out = {"source": "ncu --nvtx --print-nvtx-rename kernel, dram__bytes_{read,write}.sum + "
                     "gpu__time_duration.sum, last of 2 training steps of scripts_dev/one_step_fused.py 16 2 "
                     "(N=16, 512^2; cold-cache serialised replays)",
           "kernels_in_step": len(recs) - begin}
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for r in recs[begin:]:
        cls = next((c for c in CLASSES if c in r["name"]), "other")
        a = agg[cls]
        a[0] += 1
        a[1] += r.get("dram__bytes_read.sum", 0.0)
        a[2] += r.get("dram__bytes_write.sum", 0.0)
        a[3] += r.get("gpu__time_duration.sum", 0.0)
    for cls, (n, rd, wr, us) in agg.items():
        out[cls] = {"kernels_per_step": n, "dram_read_bytes_per_step": rd,
                    "dram_write_bytes_per_step": wr, "dram_bytes_per_step": rd + wr,
                    "ncu_time_us_per_step": us}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
