out=gpurun_out/r2k; mkdir -p $out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-hiprt-reference --no-graph > $out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"traceTrianglesKernel" -s 5 -c 2 -o $out/trace_c3_frame python tools/profile_frame.py 4 > $out/ncu_frame.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"hplocKernel|collapsePackKernel|compressBoxNodesKernel|onesweepPass|pairTrianglesKernel" -s 8 -c 8 -o $out/build_c3 python tools/profile_frame.py 0 > $out/ncu_build.log 2>&1
tail -3 $out/ncu_frame.log $out/ncu_build.log; wc -l $out/launches.csv
