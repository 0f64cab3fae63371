out=gpurun_out/${1:-r2sbvh}; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_sbvh.py -m gpu -q -x -s 2>&1 | tail -40 > $out/sbvh_tests.log
cat $out/sbvh_tests.log
