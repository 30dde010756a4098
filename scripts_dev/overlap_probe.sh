#!/bin/bash
# One-box A/B of the weight-gradient / backward-chain overlap knobs (round 2). Usage: bash scripts_dev/overlap_probe.sh
F="--steps 20 --warmup 5 --no-infer --no-wide --no-cpu-baseline"
run() {   # name, bench flags, env...
  local name=$1; shift; local flags=$1; shift
  env "$@" python bench.py $F $flags > gpurun_out/ovl_$name.log 2>gpurun_out/ovl_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/ovl_{name}.log") if l.startswith("{")][-1])
    print(f"{name:34s} {d['value']:8.1f} img/s {d['ms_per_step']:7.3f} ms  loss {d['final_loss']:.6f}")
except Exception as e:
    print(name, "FAILED", e)
PY
}
run base            ""                       UB_X=0
run prio            "--stream-priority -1"   UB_X=0
run defer           ""                       UB_WGRAD_DEFER=1
run defer_prio      "--stream-priority -1"   UB_WGRAD_DEFER=1
run maxkb128        ""                       UB_WGRAD_MAXKB=128
run maxkb128_prio   "--stream-priority -1"   UB_WGRAD_MAXKB=128
run defer_kb128_prio "--stream-priority -1"  UB_WGRAD_DEFER=1 UB_WGRAD_MAXKB=128
run defer_kb64_prio "--stream-priority -1"   UB_WGRAD_DEFER=1 UB_WGRAD_MAXKB=64
run defer_kb128_prio_a2 "--stream-priority -1" UB_WGRAD_DEFER=1 UB_WGRAD_MAXKB=128 UB_BNBWD_APPLY_CTAS=2
run defer_kb128_a2  ""                       UB_WGRAD_DEFER=1 UB_WGRAD_MAXKB=128 UB_BNBWD_APPLY_CTAS=2
run base2           ""                       UB_X=0
