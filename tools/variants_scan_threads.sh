L=dual_threshold_optimization_b200/lib/libdto_b200.so
cp $L /tmp/orig.so
for T in 256 320 352 384; do
  cp build/variants/libdto_b200_$T.so $L
  python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/var_$T.json 2> gpurun_out/var_$T.err
  python -c "
import json;d=json.load(open('gpurun_out/var_$T.json'));print($T, d['value'], d['kernel_ms'])"
done
cp /tmp/orig.so $L
