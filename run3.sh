set -x
mkdir -p gpurun_out
# launch list of the bench command (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300
# full capture of the apply kernels (3 launches after the warm-up solve)
ncu --set full --clock-control none --import-source on -k regex:apply_kernel -s 8 -c 3 -o gpurun_out/prof_apply_r01 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/prof_apply.log 2>&1
tail -2 gpurun_out/prof_apply.log | cut -c1-300
# full capture of the final update kernel
ncu --set full --clock-control none --import-source on -k regex:BiEndBody -s 2 -c 1 -o gpurun_out/prof_final_r01 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/prof_final.log 2>&1
ls -la gpurun_out
