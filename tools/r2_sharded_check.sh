# First multi-GPU check of the kr-row sharded field solve; run on a
# 2- or 8-GPU box:  gpurun --gpus N -- 'bash tools/r2_sharded_check.sh N'
# Prints / stores the replicated and the sharded bench lines back to back.
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
  --master-port 29516 tools/sharded_parity.py 2>&1 | tail -3 | tee gpurun_out/sharded_parity_${N}gpu.txt
for flag in "--replicated-solve" ""; do
  tag=replicated; [ -z "$flag" ] && tag=sharded
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline $flag \
    > gpurun_out/bench_${N}gpu_${tag}.json 2> gpurun_out/bench_${N}gpu_${tag}.err
  tail -c 600 gpurun_out/bench_${N}gpu_${tag}.json
done
# 64-row contraction tiles for the owned-row outputs (opt-in until validated): parity on one
# GPU with the virtual-shard tests, then the sharded bench again
CHB_DHT_TILE64=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "sharded or dht_contraction or hermitian" \
  2>&1 | tail -3 | tee gpurun_out/tile64_pytest.txt
CHB_DHT_TILE64=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
  --master-port 29518 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline \
  > gpurun_out/bench_${N}gpu_sharded_tile64.json 2> gpurun_out/bench_${N}gpu_sharded_tile64.err
tail -c 600 gpurun_out/bench_${N}gpu_sharded_tile64.json
