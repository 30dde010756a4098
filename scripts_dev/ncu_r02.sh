#!/bin/bash
# ncu evidence of round 2 (one B200, via gpurun): per-launch device times of one bench.py-identical
# training step, DRAM traffic per kernel class (NVTX-named), and --set full captures of the tensor-core
# and BN kernels of the SECOND step. Every ncu run follows a plain run of the same command that exited 0
# (B200_PROFILING.md). Only CSV summaries are kept (gpurun brings back at most 64 MiB).
set -u
OUT=gpurun_out
CMD="python scripts_dev/one_step_fused.py 16 2"
$CMD > $OUT/ncu_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r02_launches.csv $CMD > $OUT/ncu_a.log 2>&1
echo "launch list rc=$?"
UB_NVTX=1 $CMD > $OUT/ncu_plain_nvtx.log 2>&1 && \
UB_NVTX=1 timeout 600 ncu --nvtx --print-nvtx-rename kernel --clock-control none --csv -c 600 \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --log-file $OUT/r02_traffic.csv $CMD > $OUT/ncu_b.log 2>&1
echo "traffic rc=$?"
# --set full over the tensor-core kernels of the second step (launches 215..): ~60 kernels
$CMD > $OUT/ncu_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none -k regex:'igemm_' -s 72 -c 72 -f -o /tmp/r02_full_gemm $CMD > $OUT/ncu_c.log 2>&1
echo "full gemm rc=$?"
python scripts_dev/ncu_summary.py /tmp/r02_full_gemm.ncu-rep > $OUT/r02_ncu_gemm.csv 2>$OUT/ncu_sum_c.err
$CMD > $OUT/ncu_plain3.log 2>&1 && \
timeout 600 ncu --set full --clock-control none -k regex:'bn_|fc1_|wgrad_reduce|sgd_|head_|wce_' -s 80 -c 80 -f -o /tmp/r02_full_ew $CMD > $OUT/ncu_d.log 2>&1
echo "full elementwise rc=$?"
python scripts_dev/ncu_summary.py /tmp/r02_full_ew.ncu-rep > $OUT/r02_ncu_elementwise.csv 2>$OUT/ncu_sum_d.err
ls -la /tmp/*.ncu-rep; du -sh $OUT
