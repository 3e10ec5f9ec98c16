set -x
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_a.csv python tools/profile_step.py --steps 1 > /dev/null 2>&1
for k in fft_pow2_kernel dht_gemm_kernel gather_push_kernel sort_scatter_kernel sort_fixup_kernel depose_kernel index_kernel; do
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 2 -f -o gpurun_out/r1a_$k python tools/profile_step.py --steps 1 > /dev/null 2>&1
done
ls -la gpurun_out
