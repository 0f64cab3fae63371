out=gpurun_out/r2h; mkdir -p $out
timeout 600 python bench.py --no-cpu-baseline --no-hiprt-reference > $out/bench1.json 2> $out/bench1.err
timeout 300 python -m pytest tests/test_gpu_frame.py -m gpu -q 2>&1 | tail -5 > $out/tests_frame.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 > $out/bench2.json 2> $out/bench2.err
timeout 600 $TR bench.py --gpus 2 --no-graph > $out/bench2_nograph.json 2> $out/bench2_nograph.err
for f in $out/*.log $out/*.json $out/*.err; do echo "== $f"; tail -c 1500 $f; done
