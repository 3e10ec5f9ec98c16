TAG=$1; shift
for k in "$@"; do
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 2 -f -o gpurun_out/${TAG}_$k python tools/profile_step.py --steps 2 > /dev/null 2>&1
done
ls gpurun_out | grep ${TAG}
