python -m pytest tests/test_gpu_qnodes.py -m gpu -q 2>&1 | tail -3
python tools/bench_build.py c3_terrain_10M_ao 1 2>&1 | tail -1
