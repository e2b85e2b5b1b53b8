#!/bin/bash
# Captures one launch of each hot kernel with `ncu --set full` on the GPU box, at the launch size bench.py uses.
#   gpurun --timeout 900 -- bash tools/ncu_capture.sh c3 r02
# Writes gpurun_out/<tag>_{scan,sigma}_<cfg>.ncu-rep and gpurun_out/<tag>_<cfg>_sha.txt (hash of the kernel sources the
# capture belongs to).  Back on the CPU box:  python tools/ncu_summary.py counters <cfg> <tag> <permutations>
# turns them into profiles/kernel_counters_<cfg>.json (what bench.py's roofline reads) and the text summaries.
CFG=${1:-c3}
TAG=${2:-r02}
PERMS=${3:-100000}
python -c "import bench; print(bench.kernel_source_sha())" > gpurun_out/${TAG}_${CFG}_sha.txt
for K in scan sigma; do
  PAT=scan_kernel; [ $K = sigma ] && PAT=sigma_sort_kernel
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$PAT -s 3 -c 1 -f -o gpurun_out/${TAG}_${K}_${CFG} \
    python bench.py --config $CFG --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --perms $PERMS > gpurun_out/${TAG}_${K}_${CFG}.log 2>&1
done
# launch list of a short bench run (cold-cache, serialised: compare shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_${CFG}.csv \
  python bench.py --config $CFG --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline --perms $PERMS > gpurun_out/${TAG}_launches_${CFG}.log 2>&1
ls -la gpurun_out/${TAG}_*
