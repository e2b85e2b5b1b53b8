#!/bin/bash
# Developer harness: benches every library build found under build/variants/ (git-ignored; built by hand with different
# compile-time knobs) by swapping it into the package on the GPU box's scratch copy.  Usage: gpurun -- bash tools/variants_bench.sh
L=dual_threshold_optimization_b200/lib/libdto_b200.so
cp $L /tmp/orig.so
for V in build/variants/*.so; do
  T=$(basename $V .so)
  cp $V $L
  python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/var_$T.json 2> gpurun_out/var_$T.err
  python -c "
import json;d=json.load(open('gpurun_out/var_$T.json'));print('$T', d['value'], d['kernel_ms'])"
done
cp /tmp/orig.so $L
