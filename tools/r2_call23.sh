# round 2, call 23: memcheck over every entry point on a small case (new kernels included)
mkdir -p gpurun_out
timeout 200 python tools/sanitize_small.py > gpurun_out/r2_sanitize_plain.log 2>&1 || { tail -5 gpurun_out/r2_sanitize_plain.log; exit 1; }
timeout 800 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/r2_sanitize.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|done|cords" gpurun_out/r2_sanitize.log | head -30
