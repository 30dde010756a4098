#!/bin/bash
# Data-parallel scaling probe on one 8 x B200 box: the bench.py training step at N = 1 and N = 8 with
# different NCCL settings for the per-stage gradient all-reduce (the ring kernel
# competes with the HBM-bound BN-backward kernels for SMs and bandwidth).
#   gpurun --gpus 8 -- bash scripts_dev/scale_probe.sh
set -u
OUT=gpurun_out
FLAGS="--steps 20 --warmup 5 --no-infer --no-wide --no-cpu-baseline"
run() {  # name, env..., then N
  local name=$1; shift
  local n=$1; shift
  if [ "$n" = 1 ]; then
    env "$@" python bench.py --gpus 1 $FLAGS > $OUT/scale_${name}.json 2> $OUT/scale_${name}.err
  else
    env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port 29517 bench.py --gpus $n $FLAGS > $OUT/scale_${name}.json 2> $OUT/scale_${name}.err
  fi
  python - "$name" $OUT/scale_${name}.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
    print(f"{sys.argv[1]:28s} n={d['n_gpus']} {d['value']:9.1f} img/s  {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['value']:9.1f}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run n1 1 UB_X=0
run n8_default 8 UB_X=0
run n8_maxctas4 8 NCCL_MAX_CTAS=4
run n8_maxctas8 8 NCCL_MAX_CTAS=8
run n8_nvls 8 NCCL_ALGO=NVLS
run n8_nvls_ctas4 8 NCCL_ALGO=NVLS NCCL_MAX_CTAS=4
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
  --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 2 --warmup 3 --no-infer --no-wide \
  --no-cpu-baseline 2>&1 | grep -E "NVLS|Algo|algo|Channel|nChannels|Connected" | sort | uniq -c | sort -rn | head -20 > $OUT/scale_nccl_info.txt
