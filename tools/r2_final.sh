# One full verification visit to a 1-GPU box: every GPU test, the headline bench (both arms), probes, a sanitizer pass.
out=gpurun_out/${1:-r2final}; mkdir -p $out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $out/tests.log
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $out/bench_ref.json 2> $out/bench_ref.err
timeout 300 python tools/bench_scene.py > $out/scene_c4.json 2> $out/scene_c4.err
timeout 300 python tools/probe_tail.py > $out/probe_tail.txt 2>&1
timeout 300 python tools/probe_size.py > $out/probe_size.txt 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_qnodes.py tests/test_gpu_multilevel.py tests/test_gpu_frame.py::test_single_gpu_frame_matches_the_oracle tests/test_gpu_scene.py tests/test_gpu_update.py -m gpu -q -x -k "not soup and not two_ranks" > $out/sanitizer_memcheck.log 2>&1; echo "compute-sanitizer exit code $?" >> $out/sanitizer_memcheck.log
for f in $out/tests.log $out/bench.err $out/bench_ref.err $out/scene_c4.json; do echo "== $f"; tail -c 1800 $f; done; tail -5 $out/sanitizer_memcheck.log
