#!/bin/bash
# Round-end profile collection on one B200 (run under gpurun): the plain run first, then the launch list,
# then one --set full capture of every kernel of one step, then the side captures (fused vs per-level
# wavelet DRAM bytes, the rough full search).  Outputs under gpurun_out/; tools/summarise_profiles.py
# turns them into the files committed under profiles/.
set -x
R=${1:-r02k}
ARGS="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-other"
timeout 300 python bench.py $ARGS > gpurun_out/${R}_plain.json 2> gpurun_out/${R}_plain.err || exit 1
# (the set-up's torch fill / copy kernels -- several hundred since every picture of a batch differs -- are filtered out)
K='hbm_wave_kernel|hbm_level_kernel|hbm_static_kernel|hbm_init|obmc_kernel|upsample_kernel|downsample_kernel|wavelet_inv|wavelet_level_kernel|edgeextend'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
  python bench.py $ARGS > gpurun_out/${R}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none \
  -k regex:"$K" -s 34 -c 34 -o gpurun_out/${R}_full python bench.py $ARGS > gpurun_out/${R}_full.log 2>&1
# the report itself is too large to travel (64 MiB limit on gpurun_out): keep its raw page as CSV
ncu -i gpurun_out/${R}_full.ncu-rep --page raw --csv > gpurun_out/${R}_full_raw.csv 2>/dev/null
rm -f gpurun_out/${R}_full.ncu-rep
# NS-1: DRAM bytes and time of the fused level-1+0 launch next to the per-level pair (32 pictures, 2160p s32, depth 2)
timeout 300 python tools/time_wavelet_fused.py > gpurun_out/${R}_wavelet_fused_times.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:'wavelet_inv' -c 60 --csv --log-file gpurun_out/${R}_wavelet_fused_dram.csv python tools/time_wavelet_fused.py only-d2 \
  > gpurun_out/${R}_wavelet_fused_dram.log 2>&1
# the rough full search (+-12, 32 1080p pictures)
timeout 300 ncu --set full --clock-control none -k regex:rough_full_kernel -s 2 -c 1 -o gpurun_out/${R}_rough python tools/profile_rough.py \
  > gpurun_out/${R}_rough.log 2>&1
ncu -i gpurun_out/${R}_rough.ncu-rep --page raw --csv > gpurun_out/${R}_rough_raw.csv 2>/dev/null
rm -f gpurun_out/${R}_rough.ncu-rep
ls -la gpurun_out/${R}_*
