# Multi-GPU check list for the kr-row sharded field solve; run on an
# N-GPU box:   gpurun --gpus N --timeout 900 -- 'bash tools/r2_sharded_check.sh N'
# (round 1 covered N = 2 and 4; N = 8 and cfg5 are open).  Everything lands in gpurun_out/.
# SECTIONS="parity bench variants peer cfg5" (default: all) picks what runs; TAG prefixes the outputs.
N=${1:-8}
SECTIONS=${SECTIONS:-parity replicated bench weak variants tile128 peer cfg5}
has() { case " $SECTIONS " in *" $1 "*) return 0;; esac; return 1; }
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"

# 1. parity: three PIC steps, sharded against replicated, on every rank
has parity && timeout 300 $TR --master-port 29516 tools/sharded_parity.py 2>&1 | tail -2 | tee gpurun_out/sharded_parity_${N}gpu.txt

# 2. bench: replicated, sharded, sharded with the collectives on a high-priority stream
#    (the contraction kernel holds whole SMs, NCCL only gets them between waves)
bench() {  # tag, env, extra flags
  env $2 timeout 300 $TR --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline $BENCH_FLAGS $3 \
    > gpurun_out/bench_${N}gpu_$1.json 2> gpurun_out/bench_${N}gpu_$1.err
  python - "$1" <<PY
import json, sys
d = json.load(open("gpurun_out/bench_${N}gpu_%s.json" % sys.argv[1]))
print("%-22s %.3f ms/step  %.3g particle-steps/s  e2e %.1f ms/step" %
      (sys.argv[1], d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]))
print("   phases", {k: round(v, 3) for k, v in d["phases_ms"].items() if v > 0.05})
print("   parity", d.get("parity"))
PY
}
has replicated && bench replicated "A=0" "--replicated-solve"
has bench && bench sharded "A=0" ""
has weak && bench sharded_weak "A=0" "--scaling weak"
has variants && bench sharded_hiprio "TORCH_NCCL_HIGH_PRIORITY=1" ""
has tile128 && bench sharded_tile128 "CHB_DHT_TILE64=0" ""

# 2b. E/B partial sums through own kernels over peer memory instead of NCCL (compiled, never
#     run before): parity first, then the bench, P2P and NVSwitch-multicast flavours
has peer && for mode in ${PEER_MODES:-1 multimem}; do
  [ -z "$SKIP_PEER_PARITY" ] && CHB_PEER_EXCHANGE=$mode timeout 300 $TR --master-port 29519 tools/sharded_parity.py 2>&1 | tail -2 \
    | tee gpurun_out/sharded_parity_${N}gpu_peer_$mode.txt
  bench sharded_peer_$mode "CHB_PEER_EXCHANGE=$mode" ""
done

# 3. cfg5 (Nx=16384, Nr=1024, M=2, 32 ppc): replicated against sharded solve
has cfg5 && for flag in ${CFG5_FLAGS:---replicated-solve --sharded-solve}; do
  timeout 600 $TR --master-port 29518 examples/lpa_script_large.py --cfg5 --steps 10 $flag 2>&1 \
    | grep "ms/step" | tee -a gpurun_out/cfg5_${N}gpu.txt
done
