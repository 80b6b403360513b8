#!/bin/bash
# ncu evidence of the current build, taken on the GPU box (one GPU):  tools/ncu_capture.sh <tag> [c2] [c3] [c5]
#   <tag>_launches_c2.csv   every launch of two fits with its device time (--metrics gpu__time_duration.sum)
#   <tag>_full_<wl>.ncu-rep + _raw.csv   `--set full` of the streaming kernels of ONE fit (all product variants)
# The raw CSVs are turned into profiles/<tag>_ncu_full_*_summary.csv and profiles/<tag>_traffic.json by tools/ncu_traffic.py.
tag=$1; shift
mkdir -p gpurun_out
Q="--steps 1 --warmup 1 --no-e2e --no-cpu --no-strong-c4 --no-models"
for wl in "$@"; do
  case $wl in
    launches) wl=c2
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:xb:: -c 1200 --csv --log-file gpurun_out/${tag}_launches_c2.csv \
          python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-strong-c4 --no-models > gpurun_out/${tag}_launches_c2.log 2>&1 ;;
    c2)
      # (the library's kernels only — namespace xb: the synthetic field alone is built by ~3000 torch launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:xb:: -c 1200 --csv --log-file gpurun_out/${tag}_launches_c2.csv \
          python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-strong-c4 --no-models > gpurun_out/${tag}_launches_c2.log 2>&1
      # a fit = 14 project_tc launches (10 passes + 4 k-column applies); capture the 14 of the timed fit
      timeout 1500 ncu --set full --clock-control none --import-source on -k regex:project_tc_kernel -s 14 -c 14 -f \
          -o gpurun_out/${tag}_full_c2 python bench.py $Q > gpurun_out/${tag}_full_c2.log 2>&1
      ncu -i gpurun_out/${tag}_full_c2.ncu-rep --page raw --csv > gpurun_out/${tag}_full_c2_raw.csv 2>/dev/null
      rm -f gpurun_out/${tag}_full_c2.ncu-rep ;;   # (gpurun brings back at most 64 MiB: the CSV pages are what is kept)
    c3)
      timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"project_tc_kernel|gram_bf16_kernel" -s 70 -c 12 -f \
          -o gpurun_out/${tag}_full_c3 python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu > gpurun_out/${tag}_full_c3.log 2>&1
      ncu -i gpurun_out/${tag}_full_c3.ncu-rep --page raw --csv > gpurun_out/${tag}_full_c3_raw.csv 2>/dev/null
      rm -f gpurun_out/${tag}_full_c3.ncu-rep ;;
    c5)
      timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"varimax_tc2_kernel|varimax_exact_mma_kernel" -s 40 -c 3 -f \
          -o gpurun_out/${tag}_full_c5 python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_full_c5.log 2>&1
      ncu -i gpurun_out/${tag}_full_c5.ncu-rep --page raw --csv > gpurun_out/${tag}_full_c5_raw.csv 2>/dev/null
      rm -f gpurun_out/${tag}_full_c5.ncu-rep ;;
  esac
done
ls -la gpurun_out/${tag}_* | head -20
