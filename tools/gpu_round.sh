#!/bin/bash
# One GPU-box visit: GPU tests, bench lines, ncu launch list + full capture of the top kernels.
# usage: tools/gpu_round.sh <tag> [what...]   (what: tests bench bench_ref c3 ref_arm scene ncu_list ncu_full ncu_build)
tag=$1; shift
what=${@:-tests bench}
out=gpurun_out/$tag; mkdir -p $out
for w in $what; do
  case $w in
    tests)    timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > $out/tests.log ;;
    bench)    timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err ;;
    bench_ref) timeout 900 python bench.py --steps 10 --warmup 3 --with-hiprt-reference > $out/bench_with_ref.json 2> $out/bench_with_ref.err ;;
    c3)       timeout 900 python bench.py --steps 10 --warmup 3 --workload c3_terrain_10M --with-hiprt-reference > $out/bench_c3.json 2> $out/bench_c3.err ;;
    ref_arm)  timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $out/bench_refarm.json 2> $out/bench_refarm.err ;;
    ncu_list) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_list.log 2>&1 ;;
    ncu_full) timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"traceTrianglesKernel" -s 2 -c 2 -o $out/trace_c2 python tools/profile_trace.py c2_sphere_1M 2 > $out/ncu_full.log 2>&1 ;;
    ncu_build) timeout 600 ncu --set full --clock-control none --import-source on -k regex:"Kernel|onesweepPass" -s 11 -c 11 -o $out/build_c3 python tools/profile_build.py c3_terrain_10M 0 2 > $out/ncu_build.log 2>&1
               timeout 600 ncu --set full --clock-control none --import-source on -k regex:"hplocKernel" -s 1 -c 1 -o $out/hploc_c3 python tools/profile_build.py c3_terrain_10M 2 2 > $out/ncu_hploc.log 2>&1 ;;
    c3ao)     timeout 900 python bench.py --steps 5 --warmup 3 --workload c3_terrain_10M_ao --with-hiprt-reference > $out/bench_c3_ao.json 2> $out/bench_c3_ao.err ;;
    scene)    timeout 300 python tools/bench_scene.py > $out/scene_c4.json 2> $out/scene_c4.err ;;
  esac
done
for f in $out/*.log $out/*.json $out/*.err; do echo "== $f"; tail -c 3000 $f; done
