K='test_data or ten_gene or background_330 or degenerate or pairing_is_uniform or table_cache or generator_is_philox'
for tool in memcheck racecheck; do
  ( echo "# compute-sanitizer --tool $tool python -c 'import __graft_entry__ as g; g.smoke()'"; timeout 600 compute-sanitizer --tool $tool python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | grep -v "^$" | tail -6 ) > gpurun_out/san_${tool}.txt
  ( echo "# compute-sanitizer --tool $tool python -m pytest tests -m gpu -x -q -k \"$K\""; timeout 900 compute-sanitizer --tool $tool python -m pytest tests -m gpu -x -q -k "$K" 2>&1 | grep -v "^$" | tail -8 ) > gpurun_out/san_${tool}_tests.txt
done
for f in gpurun_out/san_*.txt; do tail -n 3 $f; done
