# One full verification visit to a 1-GPU box: every GPU test, the headline bench (both arms), the high-quality build bench, a sanitizer pass.
out=gpurun_out/${1:-r2final}; mkdir -p $out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $out/tests.log
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $out/bench_ref.json 2> $out/bench_ref.err
timeout 600 python tools/bench_sbvh.py $out/sbvh.json > $out/sbvh.log 2>&1
timeout 300 python tools/profile_sbvh.py terrain10M sphere1M needles200k > $out/sbvh_timeline.txt 2>&1
timeout 300 python tools/bench_scene.py > $out/scene_c4.json 2> $out/scene_c4.err
timeout 300 python -m pytest tests/test_gpu_reference.py -m gpu -q -s -k scene_hits 2>&1 | grep "scene hits vs" > $out/scene_vs_reference.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-hiprt-reference --no-graph > $out/ncu_list.log 2>&1
python tools/summarize_launches.py $out/launches.csv > $out/launches_summary.txt 2>&1
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_qnodes.py tests/test_gpu_multilevel.py tests/test_gpu_frame.py::test_single_gpu_frame_matches_the_oracle tests/test_gpu_scene.py tests/test_gpu_update.py tests/test_gpu_sbvh.py tests/test_gpu_pathtrace.py -m gpu -q -x -k "not soup and not two_ranks and not large and not needles20k" > $out/sanitizer_memcheck.log 2>&1; echo "compute-sanitizer exit code $?" >> $out/sanitizer_memcheck.log
for f in $out/tests.log $out/bench.err $out/bench_ref.err $out/scene_c4.json $out/scene_vs_reference.txt; do echo "== $f"; tail -c 1500 $f; done; tail -4 $out/sbvh.log | cut -c1-400; tail -6 $out/sanitizer_memcheck.log
