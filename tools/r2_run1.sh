out=gpurun_out/r2a; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60 > $out/tests.log
timeout 900 python bench.py --steps 5 --warmup 3 --workload c3_terrain_10M_ao --with-hiprt-reference > $out/bench_c3_ao.json 2> $out/bench_c3_ao.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"traceTrianglesKernel" -s 4 -c 2 -o $out/trace_c3_ao python tools/profile_ao.py 16 1 > $out/ncu_ao.log 2>&1
for f in $out/*.log $out/*.json $out/*.err; do echo "== $f"; tail -c 2500 $f; done
