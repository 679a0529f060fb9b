# round 2, call 27: software prefetch in the window extension: distance / level sweep (side runs)
mkdir -p gpurun_out
export LNR_BENCH_NO_SMALL=1
for d in 0 3 6 12 -3 -6; do
  LNR_PF_DIST=$d timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench27_pf$d.json 2> gpurun_out/r2_bench27_pf$d.err; echo "pf $d rc=$?"
done
