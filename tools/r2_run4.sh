out=gpurun_out/r2d; mkdir -p $out
timeout 900 python tools/probe_shard.py > $out/probe_shard.txt 2>&1
cat $out/probe_shard.txt
