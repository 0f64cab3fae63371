out=gpurun_out/r2i; mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
for n in 8 4; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $n --steps 20 > $out/bench$n.json 2> $out/bench$n.err
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 8 --steps 20 --tile 64 > $out/bench8_t64.json 2> $out/bench8_t64.err
for f in $out/*.json $out/*.err; do echo "== $f"; tail -c 1200 $f; done
