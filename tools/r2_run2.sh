out=gpurun_out/r2b; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $out/tests.log
hiprt_b200/csrc/build/cache_peaks > $out/cache_peaks.json 2> $out/cache_peaks.err
mkdir -p profiles; cp $out/cache_peaks.json profiles/r02_cache_peaks.json
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $out/bench_ref.json 2> $out/bench_ref.err
for f in $out/*.log $out/*.json $out/*.err; do echo "== $f"; tail -c 3000 $f; done
