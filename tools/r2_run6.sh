out=gpurun_out/r2f; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"traceTrianglesKernel" -s 4 -c 2 -o $out/trace_c3_ao_q python tools/profile_ao.py 16 2 > $out/ncu_ao.log 2>&1
cat $out/ncu_ao.log | tail -5
