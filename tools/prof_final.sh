set -x
TAG=$1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv python tools/profile_step.py --steps 2 > /dev/null 2>&1
for k in dht_gemm_kernel depose_kernel gather_push_kernel fft_pow2_kernel index_kernel sort_scatter_kernel sort_fixup_kernel psatd_kernel; do
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 2 -s 2 -f -o gpurun_out/${TAG}_$k python tools/profile_step.py --steps 2 > /dev/null 2>&1
done
ls gpurun_out | grep ${TAG}
