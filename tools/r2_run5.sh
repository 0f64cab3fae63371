out=gpurun_out/r2e; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_qnodes.py -m gpu -q -x 2>&1 | tail -40 > $out/tests_q.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $out/tests.log
timeout 600 python bench.py --no-cpu-baseline > $out/bench.json 2> $out/bench.err
HIPRT_B200_TRACE_VARIANT=3 timeout 600 python bench.py --no-cpu-baseline --no-hiprt-reference > $out/bench_v3.json 2> $out/bench_v3.err
for f in $out/*.log $out/*.json $out/*.err; do echo "== $f"; tail -c 3000 $f; done
