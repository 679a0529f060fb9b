# round 2, call 16: where does -i 2 stall at scale? (every step under its own short timeout)
mkdir -p gpurun_out
for cfg in "50e6 1024" "500e6 1024" "3.1e9 1024"; do
  echo "== $cfg" >> gpurun_out/r2_hprobe.log
  LNR_TRACE=1 timeout 150 python tools/hindex_probe.py $cfg >> gpurun_out/r2_hprobe.log 2>&1; echo "rc=$?" >> gpurun_out/r2_hprobe.log
done
tail -c 6000 gpurun_out/r2_hprobe.log
