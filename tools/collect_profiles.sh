#!/bin/bash
# Round-end profile collection on one B200 (run under gpurun): the plain run first, then the launch list,
# then one --set full capture of every kernel of one step.  Outputs under gpurun_out/.
set -x
ARGS="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-other"
timeout 300 python bench.py $ARGS > gpurun_out/r02d_plain.json 2> gpurun_out/r02d_plain.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02d_launches.csv \
  python bench.py $ARGS > gpurun_out/r02d_launches.log 2>&1
timeout 900 ncu --set full --clock-control none \
  -k regex:'hbm_wave_kernel|hbm_static_kernel|hbm_init|obmc_kernel_v4|upsample_kernel_words|downsample_kernel_words|wavelet_inv_fast|wavelet_level_kernel|edgeextend' \
  -s 34 -c 34 -o gpurun_out/r02d_full python bench.py $ARGS > gpurun_out/r02d_full.log 2>&1
# the report itself is too large to travel (64 MiB limit on gpurun_out): keep its raw page as CSV
ncu -i gpurun_out/r02d_full.ncu-rep --page raw --csv > gpurun_out/r02d_full_raw.csv 2>/dev/null
rm -f gpurun_out/r02d_full.ncu-rep
ls -la gpurun_out/r02d_*
