mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
P='import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"],1), round(d["roofline"]["achieved"]), round(d["roofline"]["frac"],3), {k:round(v,4) for k,v in d["kernel_ms_per_iteration"].items()}, round(d["iteration_roofline"]["frac_of_nominal_8TBs"],3), round(d["e2e"]["value"],1))'
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/b.err | tee gpurun_out/bench_v3.json | python -c "$P" V3

timeout 300 python bench.py --solver cg --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/b.err | tee gpurun_out/bench_cg_v3.json | python -c "$P" CG_V3
timeout 300 python bench.py --cell hex --n 215 --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/b.err | tee gpurun_out/bench_hex_v3.json | python -c "$P" HEX_V3
