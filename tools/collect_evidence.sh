#!/bin/bash
# Runs ON the GPU box (gpurun): the default bench, the reference arm, and the ncu captures that profiles/ summarises.
# Every ncu pass runs only after the same command has exited 0 without ncu; numbers printed under ncu are never bench values.
set -u
O=gpurun_out
mkdir -p $O
python bench.py --layers-out $O/r2_layers.json > $O/r2_bench.json 2> $O/r2_bench.err || { echo "bench failed"; tail -5 $O/r2_bench.err; exit 1; }
python bench.py --impl reference --steps 5 --warmup 1 > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err
FWD="python bench.py --no-graph --steps 2 --warmup 1 --no-cpu --no-train --no-pose --no-eager --no-fp32"
$FWD > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv $FWD > $O/ncu_fwd.log 2>&1
FWD1="python bench.py --no-graph --steps 1 --warmup 1 --no-cpu --no-train --no-pose --no-eager --no-fp32"
ncu --set full --clock-control none -k regex:"dwconv_smem_kernel|dwconv_half_kernel|dw_col_kernel|conv_gemm_kernel|conv3x3_halo_kernel" -c 46 -f -o /tmp/r2_fwd $FWD1 > $O/ncu_fwd_full.log 2>&1 \
  && ncu -i /tmp/r2_fwd.ncu-rep --page raw --csv > $O/r2_prof_fwd.csv
ncu --set full --clock-control none -k regex:"stem_kernel|se_fused_kernel|gap_kernel|head_mix_kernel|upsample_out_kernel" -c 14 -f -o /tmp/r2_misc $FWD1 > $O/ncu_misc_full.log 2>&1 \
  && ncu -i /tmp/r2_misc.ncu-rep --page raw --csv > $O/r2_prof_misc.csv
TRAIN_B=32 TRAIN_STEPS=3 python tools/train_probe.py > $O/r2_train_probe.log 2>&1 \
  && TRAIN_B=32 TRAIN_STEPS=3 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_train32_launches.csv python tools/train_probe.py > /dev/null 2>&1
TRAIN_B=256 TRAIN_STEPS=3 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_train256_launches.csv python tools/train_probe.py > /dev/null 2>&1
# (a --set full capture of a whole B=32 training step was tried here: ncu failed to shut the target down after ~360 kernels x 40 passes
# and the call ran into its time limit; per-kernel full captures of the training kernels are taken with tools/train_probe.py and -k instead)
ls -la $O | grep r2_ | tail -20
