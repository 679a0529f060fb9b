# round 2, call 17: -i 2 mapping time against genome size (trend of the kernel times)
mkdir -p gpurun_out
for cfg in "100e6 256" "200e6 256" "300e6 256"; do
  echo "== $cfg" >> gpurun_out/r2_hprobe2.log
  LNR_TRACE=1 timeout 60 python tools/hindex_probe.py $cfg >> gpurun_out/r2_hprobe2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_hprobe2.log
done
tail -c 8000 gpurun_out/r2_hprobe2.log
