out=gpurun_out/r2c; mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_frame.py tests/test_gpu_multilevel.py tests/test_gpu_reference.py tests/test_gpu_configs.py -m gpu -q 2>&1 | tail -30 > $out/tests.log
timeout 600 python bench.py --no-cpu-baseline > $out/bench1.json 2> $out/bench1.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $out/bench2_p2p.json 2> $out/bench2_p2p.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --exchange nccl > $out/bench2_nccl.json 2> $out/bench2_nccl.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --replicate rebuild --tile 64 > $out/bench2_t64.json 2> $out/bench2_t64.err
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > $out/bench2_ref.json 2> $out/bench2_ref.err
for f in $out/*.log $out/*.json $out/*.err; do echo "== $f"; tail -c 2500 $f; done
