# usage (on the GPU box): bash tools/prof_round.sh TAG
# launch list of two PIC steps + ncu --set full captures of the main kernels (cfg3), taken at
# the same point of the run bench.py times (a few steps after the lattice start)
set -x
TAG=$1
PS=${PRESTEPS:-5}
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv python tools/profile_step.py --steps 2 --presteps $PS > /dev/null 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"dht_gemm|depose_kernel|gather_push_kernel|fft_pow2_kernel|fft_damp_kernel|sort_scatter_kernel|sort_fixup_kernel|psatd_kernel" \
    -f -o gpurun_out/${TAG}_step python tools/profile_step.py --steps 1 --presteps $PS > /dev/null 2>&1
ls -la gpurun_out | grep ${TAG}
