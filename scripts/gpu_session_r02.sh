#!/bin/bash
# GPU sessions of round 2, one block per gpurun call; outputs go to gpurun_out/r02_<what>/ (summaries are copied to
# profiles/ by hand afterwards).
#
#   gpurun --timeout 1500 -- 'bash scripts/gpu_session_r02.sh a'
set -u
what=${1:-a}
out=gpurun_out/r02_$what
mkdir -p "$out"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
case "$what" in
  a)
    # 1. can two ranks share ONE GPU (time-sliced contexts, CUDA-IPC peer mapping of the same device)? If so the
    #    world-size-2 bit-exactness tests run on the driver's 1-GPU test box.
    ( time timeout 600 $TR --nproc-per-node 2 --master-port 29531 tests/_dist_worker.py p2p ) > "$out/two_ranks_one_gpu.log" 2>&1
    echo "two ranks on one GPU: rc=$?"; tail -4 "$out/two_ranks_one_gpu.log"
    # 2. statement grouping at full size (VERDICT task 3)
    python scripts/solver_sweep.py --solvers cg,bicgstab,cgs,bicgstabl,tfqmr,tfqmr1,idrs,gmres,fgmres,richardson,grouped_idrs,grouped_bicgstabl \
        --out "$out/solver_sweep_as_written.json" > "$out/as_written.log" 2>&1
    python scripts/solver_sweep.py --grouping 1 --solvers cg,bicgstab,cgs,bicgstabl,tfqmr,tfqmr1,idrs,gmres,fgmres,richardson \
        --out "$out/solver_sweep_grouping1.json" > "$out/grouping1.log" 2>&1
    python scripts/solver_sweep.py --grouping 2 --solvers cg,bicgstab,cgs,bicgstabl,tfqmr,tfqmr1,idrs,gmres,fgmres,richardson \
        --out "$out/solver_sweep_grouping2.json" > "$out/grouping2.log" 2>&1
    tail -2 "$out/as_written.log" "$out/grouping1.log" "$out/grouping2.log"
    # 3. what one rank of the 8-GPU run costs without any communication: 1.23 M tets on one GPU, per kernel slot
    python bench.py --axis 59 --steps 200 --warmup 20 --no-cpu-baseline > "$out/bench_n1_axis59_bicgstab.json" 2> "$out/bench_axis59.err"
    python bench.py --axis 59 --steps 200 --warmup 20 --no-cpu-baseline --solver cg > "$out/bench_n1_axis59_cg.json" 2>> "$out/bench_axis59.err"
    tail -c 700 "$out/bench_n1_axis59_bicgstab.json"
    # 4. config 5: the 1e8 / 2e8-cell points
    python scripts/apply_sweep.py --cells hexlat --sizes 1e8,2e8 --out "$out/apply_sweep_hexlat.json" > "$out/config5.log" 2>&1
    tail -3 "$out/config5.log"
    ;;
  b)
    # the persistent whole-solve kernel: parity first (bounded spins so that a protocol bug ends in an error, not a hang)
    export SB_SPIN_TIMEOUT_S=20
    timeout 900 python -m pytest tests/test_gpu_mega.py -x -q 2>&1 | tail -25 > "$out/pytest_mega.log"; cat "$out/pytest_mega.log"
    timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 > "$out/pytest_gpu.log"; tail -12 "$out/pytest_gpu.log"
    ( time timeout 600 $TR --nproc-per-node 2 --master-port 29531 tests/_dist_worker.py p2p ) > "$out/two_ranks_one_gpu.log" 2>&1
    echo "two ranks on one GPU: rc=$?"; tail -6 "$out/two_ranks_one_gpu.log"
    unset SB_SPIN_TIMEOUT_S
    python bench.py --axis 59 --steps 200 --warmup 20 --no-cpu-baseline > "$out/bench_n1_axis59.json" 2> "$out/bench_axis59.err"
    tail -c 1500 "$out/bench_n1_axis59.json"
    python bench.py --steps 200 --warmup 20 > "$out/bench_n1.json" 2> "$out/bench_n1.err"
    tail -c 2500 "$out/bench_n1.json"; tail -3 "$out/bench_n1.err"
    python scripts/solver_sweep.py --solvers cg,bicgstab,cgs,bicgstabl,tfqmr,tfqmr1,idrs,gmres,fgmres,richardson,grouped_idrs,grouped_bicgstabl \
        --out "$out/solver_sweep_as_written.json" > "$out/as_written.log" 2>&1
    python scripts/solver_sweep.py --grouping 1 --solvers cg,bicgstab,cgs,bicgstabl,tfqmr,tfqmr1,idrs,gmres,fgmres,richardson \
        --out "$out/solver_sweep_grouping1.json" > "$out/grouping1.log" 2>&1
    python scripts/solver_sweep.py --grouping 2 --solvers cg,bicgstab,cgs,bicgstabl,tfqmr,tfqmr1,idrs,gmres,fgmres,richardson \
        --out "$out/solver_sweep_grouping2.json" > "$out/grouping2.log" 2>&1
    tail -2 "$out/as_written.log" "$out/grouping1.log" "$out/grouping2.log"
    python scripts/apply_sweep.py --cells hexlat --sizes 1e8,2e8 --out "$out/apply_sweep_hexlat.json" > "$out/config5.log" 2>&1
    tail -3 "$out/config5.log"
    ;;
  c)
    # persistent kernel v2 (dynamic tiles, bulk-copy staging of the element-wise steps): parity, then A/B against stepwise
    export SB_SPIN_TIMEOUT_S=20
    timeout 900 python -m pytest tests/test_gpu_mega.py tests/test_gpu_scale.py -x -q 2>&1 | tail -25 > "$out/pytest_mega.log"; tail -5 "$out/pytest_mega.log"
    ( time timeout 600 $TR --nproc-per-node 2 --master-port 29531 tests/_dist_worker.py p2p ) > "$out/two_ranks_one_gpu.log" 2>&1
    echo "two ranks on one GPU: rc=$?"; tail -4 "$out/two_ranks_one_gpu.log"
    unset SB_SPIN_TIMEOUT_S
    for ax in 59 119; do for sch in persistent stepwise; do
      python bench.py --axis $ax --steps 200 --warmup 20 --no-cpu-baseline --schedule $sch > "$out/bench_n1_axis${ax}_$sch.json" 2>> "$out/bench.err"
      python - "$out/bench_n1_axis${ax}_$sch.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
o = d["other_solver"]
print(sys.argv[1], "bicgstab", round(d["value"]), "it/s", round(1e3 * d["ms_per_step"], 1), "us;  cg", round(o["value"]), "it/s", round(1e3 * o["ms_per_step"], 1), "us")
print("   phases", d["phases"] and {k: round(v, 1) for k, v in d["phases"]["us_per_step"].items()}, d["phases"] and {k: round(v, 1) for k, v in d["phases"]["us_barrier_wait_for_last_cta"].items()})
PY
    done; done
    ;;
  d)
    # persistent kernel v3 (balanced grid, last-arriver reduction, cross-barrier prefetch) + the CGS launch list
    export SB_SPIN_TIMEOUT_S=20
    timeout 900 python -m pytest tests/test_gpu_mega.py tests/test_gpu_scale.py -x -q 2>&1 | tail -25 > "$out/pytest_mega.log"; tail -5 "$out/pytest_mega.log"
    ( time timeout 600 $TR --nproc-per-node 2 --master-port 29531 tests/_dist_worker.py p2p ) > "$out/two_ranks_one_gpu.log" 2>&1
    echo "two ranks on one GPU: rc=$?"; tail -4 "$out/two_ranks_one_gpu.log"
    unset SB_SPIN_TIMEOUT_S
    for ax in 59 84 119; do for sch in persistent stepwise; do
      python bench.py --axis $ax --steps 200 --warmup 20 --no-cpu-baseline --schedule $sch > "$out/bench_n1_axis${ax}_$sch.json" 2>> "$out/bench.err"
      python - "$out/bench_n1_axis${ax}_$sch.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
o = d["other_solver"]
print(sys.argv[1], "bicgstab", round(d["value"]), "it/s", round(1e3 * d["ms_per_step"], 1), "us;  cg", round(o["value"]), "it/s", round(1e3 * o["ms_per_step"], 1), "us")
print("   phases", d["phases"] and {k: round(v, 1) for k, v in d["phases"]["us_per_step"].items()}, d["phases"] and {k: round(v, 1) for k, v in d["phases"]["us_in_barrier_cta0"].items()})
PY
    done; done
    ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file "$out/launches_cgs_10M.csv" \
        python scripts/solver_sweep.py --solvers cgs --steps 4 --repeats 1 > "$out/ncu_cgs.log" 2>&1
    python scripts/summarize_ncu.py launches "$out/launches_cgs_10M.csv" "$out/launches_cgs_10M" "CGS, reference template on DeviceVector, 10.1 M tets" > /dev/null 2>&1; tail -15 "$out/launches_cgs_10M_summary.txt"
    ;;
  e)
    # stepwise schedule with folded reductions (the default now): the whole GPU suite, then A/B against the persistent kernel
    export SB_SPIN_TIMEOUT_S=30
    timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > "$out/pytest_gpu.log"; tail -8 "$out/pytest_gpu.log"
    unset SB_SPIN_TIMEOUT_S
    python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; tail -4 "$out/smoke.log"
    for ax in 59 84 119; do for sch in auto persistent; do
      python bench.py --axis $ax --steps 200 --warmup 20 --no-cpu-baseline --schedule $sch > "$out/bench_n1_axis${ax}_$sch.json" 2>> "$out/bench.err"
      python - "$out/bench_n1_axis${ax}_$sch.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
o = d["other_solver"]
print(sys.argv[1], "bicgstab", round(d["value"]), "it/s", round(1e3 * d["ms_per_step"], 1), "us;  cg", round(o["value"]), "it/s", round(1e3 * o["ms_per_step"], 1), "us; launches", d["gpu_launches"])
print("   stepwise slots us", {k: round(1e3 * v, 1) for k, v in d["stepwise"]["kernel_ms_per_iteration"].items()})
PY
    done; done
    ;;
  f)
    # folded reductions v2 (CTA 0 reduces and raises a flag; the other CTAs only acquire it)
    export SB_SPIN_TIMEOUT_S=30
    timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > "$out/pytest_gpu.log"; tail -8 "$out/pytest_gpu.log"
    unset SB_SPIN_TIMEOUT_S
    for ax in 59 84 119; do
      python bench.py --axis $ax --steps 200 --warmup 20 --no-cpu-baseline > "$out/bench_n1_axis${ax}_auto.json" 2>> "$out/bench.err"
      python - "$out/bench_n1_axis${ax}_auto.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
o = d["other_solver"]
print(sys.argv[1], "bicgstab", round(d["value"]), "it/s", round(1e3 * d["ms_per_step"], 1), "us;  cg", round(o["value"]), "it/s", round(1e3 * o["ms_per_step"], 1), "us; launches", d["gpu_launches"])
print("   stepwise slots us", {k: round(1e3 * v, 1) for k, v in d["stepwise"]["kernel_ms_per_iteration"].items()})
PY
    done
    ;;
esac
