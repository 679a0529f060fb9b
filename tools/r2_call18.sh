# round 2, call 18: which launch of the -i 2 mapping does not return at 200 Mbase?
mkdir -p gpurun_out
LNR_TRACE=1 LNR_TRACE_SYNC=1 timeout 40 python tools/hindex_probe.py 200e6 64 > gpurun_out/r2_hprobe3.log 2>&1; echo "rc=$?" >> gpurun_out/r2_hprobe3.log
tail -n 25 gpurun_out/r2_hprobe3.log
