# One visit to an N-GPU box: the C5 render and batch-build halves on N GPUs.  usage: r2_multi.sh <N> <outdir>
n=${1:-8}; out=gpurun_out/${2:-r2multi}; mkdir -p $out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) "$@"; }
run tools/bench_pathtrace.py --size 4096 --spp 64 2>$out/pt_n$n.err | tail -1 > $out/pt_n$n.json
run tools/bench_batch.py 100000 5 2>$out/batch_n$n.err | tail -1 > $out/batch_n$n.json
for f in $out/pt_n$n.json $out/batch_n$n.json; do echo "== $f"; cut -c1-330 $f; done; tail -3 $out/pt_n$n.err
