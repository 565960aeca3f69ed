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
  i)
    # tuning bits of the stepwise schedule on one GPU: the whole GPU suite (incl. two ranks sharing the GPU: push-on-produce,
    # in-kernel reducer, PDL), then the A/B at the per-rank sizes of 8 / 4 / 1 GPUs
    export SB_SPIN_TIMEOUT_S=30
    timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > "$out/pytest_gpu.log"; tail -8 "$out/pytest_gpu.log"
    unset SB_SPIN_TIMEOUT_S
    for ax in 59 75 119; do
      timeout 600 python scripts/scale_ab.py --axis $ax --out "$out/ab_n1_axis$ax.json" > "$out/ab_n1_axis$ax.jsonl" 2> "$out/ab_n1_axis$ax.log"
      grep "^\[ab\]" "$out/ab_n1_axis$ax.log"
    done
    ;;
  j)
    # one GPU: lazy halo push (two ranks sharing the GPU: protocol + bits), batched in-kernel reducer, rolled run-time
    # expression interpreter; A/B at the per-rank size of 8 GPUs and at full size; CGS through the reference template
    export SB_SPIN_TIMEOUT_S=30
    timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > "$out/pytest_gpu.log"; tail -8 "$out/pytest_gpu.log"
    unset SB_SPIN_TIMEOUT_S
    for ax in 59 119; do
      timeout 600 python scripts/scale_ab.py --axis $ax --variants off,stream,red,red+stream,stream+pdlfa --out "$out/ab_n1_axis$ax.json" > "$out/ab_n1_axis$ax.jsonl" 2> "$out/ab_n1_axis$ax.log"
      grep "^\[ab\]" "$out/ab_n1_axis$ax.log"
    done
    python scripts/solver_sweep.py --solvers cgs,bicgstab,tfqmr --out "$out/solver_sweep_cgs.json" > "$out/solver_sweep_cgs.log" 2>&1
    tail -4 "$out/solver_sweep_cgs.log"
    ;;
  k)
    # one GPU, the round's record: the whole GPU suite, smoke, the bench line, the ncu launch list of the same command,
    # one `ncu --set full` capture of the apply kernels, time to solution against the reference CPU solve, every solver
    export SB_SPIN_TIMEOUT_S=30
    timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > "$out/pytest_gpu.log"; tail -5 "$out/pytest_gpu.log"
    unset SB_SPIN_TIMEOUT_S
    python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; tail -4 "$out/smoke.log"
    python bench.py --steps 200 --warmup 20 > "$out/bench_n1.json" 2> "$out/bench_n1.err"; tail -c 600 "$out/bench_n1.json"; tail -2 "$out/bench_n1.err"
    python bench.py --impl reference --steps 20 --warmup 3 > "$out/bench_reference_arm.json" 2> "$out/bench_reference_arm.err"; tail -c 400 "$out/bench_reference_arm.json"
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_bench.csv" \
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline > "$out/ncu_launches.log" 2>&1
    python scripts/summarize_ncu.py launches "$out/launches_bench.csv" "$out/launches_bench" "bench.py --steps 2 --warmup 1 (BiCGStab + CG legs, 10.1 M tets, stepwise schedule)" > /dev/null 2>&1
    head -16 "$out/launches_bench_summary.txt"
    ncu --set full --clock-control none --import-source on -k regex:apply_kernel_tma -c 8 -o "$out/apply_full" \
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline --single-solver > "$out/ncu_full.log" 2>&1
    ls -la "$out"/apply_full* 2>&1 | tail -2
    timeout 900 python scripts/time_to_solution.py --cpu --out "$out/time_to_solution_10M.json" > "$out/tts.jsonl" 2> "$out/tts.err"; tail -3 "$out/tts.err"; cut -c1-400 "$out/tts.jsonl"
    python scripts/solver_sweep.py --solvers cg,bicgstab,cgs,bicgstabl,tfqmr,tfqmr1,idrs,gmres,fgmres,richardson,grouped_idrs,grouped_bicgstabl \
        --out "$out/solver_sweep_default.json" > "$out/solver_sweep.log" 2>&1
    tail -3 "$out/solver_sweep.log" | cut -c1-300
    # the strong-scaling reference point of a 30 M-cell problem (its 8-GPU leg: block h5)
    python bench.py --axis 171 --steps 100 --warmup 10 --no-cpu-baseline > "$out/bench_n1_30M.json" 2> "$out/bench_n1_30M.err"; tail -c 300 "$out/bench_n1_30M.json"
    ;;
  k2)
    # one GPU: the files of block k once more (that call's directory exceeded the transfer limit because of a 73 MB ncu
    # report and nothing came back): the bench line, a SMALL `ncu --set full` capture of the two BiCGStab apply kernels
    # (raw page exported on the box, the report itself dropped when it is large), the single-GPU point of the 20 M-cell
    # strong-scaling pair
    python bench.py --steps 200 --warmup 20 > "$out/bench_n1.json" 2> "$out/bench_n1.err"; tail -c 300 "$out/bench_n1.json"; tail -2 "$out/bench_n1.err"
    ncu --set full --clock-control none -k regex:apply_kernel_tma --launch-skip 2 -c 4 -o "$out/apply_full" \
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline --single-solver > "$out/ncu_full.log" 2>&1
    ncu -i "$out/apply_full.ncu-rep" --page raw --csv > "$out/apply_full_raw.csv" 2> /dev/null
    ls -la "$out"/apply_full* | tail -3
    if [ "$(stat -c %s "$out/apply_full.ncu-rep" 2>/dev/null || echo 0)" -gt 30000000 ]; then rm -f "$out/apply_full.ncu-rep"; fi
    timeout 300 python scripts/scale_ab.py --axis 150 --variants default,off --no-profile --out "$out/ab_n1_axis150.json" > "$out/ab_n1_axis150.jsonl" 2> "$out/ab_n1_axis150.log"
    grep "^\[ab\]" "$out/ab_n1_axis150.log"
    du -sh "$out"
    ;;
  h5)
    # N GPUs: a larger strong-scaling point with the library's defaults (axis 161 = 25.0 M cells, 3.13 M per rank at N = 8);
    # `h5 1 161` is its single-GPU reference
    N=${2:-8}
    AX=${3:-161}
    timeout 200 $TR --nproc-per-node $N --master-port 29561 scripts/scale_ab.py --axis $AX --variants default --no-profile \
        --out "$out/ab_n${N}_axis$AX.json" > "$out/ab_n${N}_axis$AX.jsonl" 2> "$out/ab_n${N}_axis$AX.log"
    grep "^\[ab\]" "$out/ab_n${N}_axis$AX.log"
    ;;
  ab)
    # (record of a finished experiment: the -D hooks and the ab/ libraries were removed once it was decided;
    #  results: profiles/r02_ab_apply_registers.txt, r02_ab_apply_without_reducer_call.txt)
    # one GPU: the TMA apply kernel compiled for 72 registers (__launch_bounds__(256, 3), the tree's build) against 64
    # registers (__launch_bounds__(256, 4): what ptxas chose in round 1), same source otherwise (ab/libstormb200_regs64.so,
    # built with -DSB_TMA_REG_CTAS3=4); two problem sizes, slot times included
    for ax in 119 59; do
      timeout 300 python scripts/scale_ab.py --axis $ax --variants default,off --out "$out/ab_regs72_axis$ax.json" > "$out/ab_regs72_axis$ax.jsonl" 2> "$out/ab_regs72_axis$ax.log"
      echo "72 registers, axis $ax"; grep "^\[ab\]" "$out/ab_regs72_axis$ax.log"
    done
    cp stormruler_b200/libstormb200.so /tmp/libstormb200_tree.so
    cp ab/libstormb200_regs64.so stormruler_b200/libstormb200.so
    for ax in 119 59; do
      timeout 300 python scripts/scale_ab.py --axis $ax --variants default,off --out "$out/ab_regs64_axis$ax.json" > "$out/ab_regs64_axis$ax.jsonl" 2> "$out/ab_regs64_axis$ax.log"
      echo "64 registers, axis $ax"; grep "^\[ab\]" "$out/ab_regs64_axis$ax.log"
    done
    cp /tmp/libstormb200_tree.so stormruler_b200/libstormb200.so
    ;;
  ab2)
    # one GPU: the tree's library against A/B builds of the same source (ab/libstormb200_<X>.so; C = the TMA apply without
    # the out-of-line reducer call, D = C without the policy branch in the refill path), 10.1 M cells, slot times included
    cp stormruler_b200/libstormb200.so /tmp/libstormb200_tree.so
    for v in tree C D tree; do
      if [ $v = tree ]; then cp /tmp/libstormb200_tree.so stormruler_b200/libstormb200.so; else cp ab/libstormb200_$v.so stormruler_b200/libstormb200.so; fi
      timeout 300 python scripts/scale_ab.py --axis 119 --variants default,off --out "$out/ab_lib_${v}_axis119.json" > "$out/ab_lib_${v}.jsonl" 2> "$out/ab_lib_$v.log"
      echo "library $v"; grep "^\[ab\]" "$out/ab_lib_$v.log"
    done
    cp /tmp/libstormb200_tree.so stormruler_b200/libstormb200.so
    ;;
  k3)
    # one GPU, after the reducer role left the apply kernels and the default element-wise kernels: the whole GPU suite
    # (the in-kernel-reducer variants run in the mega and the two-ranks-on-one-GPU tests), the bench line, a quick A/B
    export SB_SPIN_TIMEOUT_S=30
    timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > "$out/pytest_gpu.log"; tail -5 "$out/pytest_gpu.log"
    unset SB_SPIN_TIMEOUT_S
    python bench.py --steps 200 --warmup 20 > "$out/bench_n1.json" 2> "$out/bench_n1.err"; tail -c 300 "$out/bench_n1.json"; tail -2 "$out/bench_n1.err"
    timeout 300 python scripts/scale_ab.py --axis 59 --variants default,off,red+stream --out "$out/ab_n1_axis59.json" > "$out/ab_n1_axis59.jsonl" 2> "$out/ab_n1_axis59.log"
    grep "^\[ab\]" "$out/ab_n1_axis59.log"
    ;;
  k4)
    # one GPU: the whole GPU suite on the round's last code state
    export SB_SPIN_TIMEOUT_S=30
    timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > "$out/pytest_gpu.log"; tail -5 "$out/pytest_gpu.log"
    unset SB_SPIN_TIMEOUT_S
    python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; tail -4 "$out/smoke.log"
    ;;
  g)
    # two GPUs: the distributed tests with a GPU per rank, the tuning A/B on the strong-scaling problem, the bench line,
    # config 3 and config 5 at N = 2
    export SB_SPIN_TIMEOUT_S=60
    timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q 2>&1 | tail -15 > "$out/pytest_multigpu_n2.log"; cat "$out/pytest_multigpu_n2.log"
    unset SB_SPIN_TIMEOUT_S
    timeout 900 $TR --nproc-per-node 2 --master-port 29540 scripts/scale_ab.py --axis 119 --out "$out/ab_n2_axis119.json" \
        --variants off,stream,noack,push,push+stream,red,red+stream,red+noack,red+push,red+push+stream,pdlfa,push+stream+pdlfa,folded,persistent \
        > "$out/ab_n2_axis119.jsonl" 2> "$out/ab_n2_axis119.log"
    grep "^\[ab\]" "$out/ab_n2_axis119.log"
    timeout 600 $TR --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 --steps 200 --warmup 20 > "$out/bench_n2_bicgstab.json" 2> "$out/bench_n2.err"
    for f in bench_n2_bicgstab; do python - "$out/$f.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d["value"]), "it/s", round(1e3 * d["ms_per_step"], 1), "us", d["config"]["schedule"], "e2e", round(d["e2e"]["value"]))
print("   slots us", {k: round(1e3 * v, 1) for k, v in d["stepwise"]["kernel_ms_per_iteration"].items()})
print("   waits us", {k: round(1e3 * v, 1) for k, v in d["stepwise"]["in_kernel_wait_ms_per_iteration"].items()})
PY
    done
    timeout 600 $TR --nproc-per-node 2 --master-port 29544 scripts/config3_gmres.py --axis 119 --steps 100 > "$out/config3_gmres_fused_10M_n2.json" 2> "$out/config3.err"
    timeout 600 $TR --nproc-per-node 2 --master-port 29545 scripts/config3_gmres.py --axis 119 --steps 100 --solver fgmres --path generic --precond jacobi > "$out/config3_fgmres_jacobi_generic_10M_n2.json" 2>> "$out/config3.err"
    tail -c 400 "$out/config3_gmres_fused_10M_n2.json"; tail -c 400 "$out/config3_fgmres_jacobi_generic_10M_n2.json"; tail -3 "$out/config3.err"
    timeout 900 $TR --nproc-per-node 2 --master-port 29546 scripts/apply_sweep.py --cells tet,hex --sizes 1e6,1e7,3e7 --out "$out/apply_sweep_n2.json" > "$out/config5_n2.log" 2>&1
    tail -6 "$out/config5_n2.log"
    ;;
  h)
    # eight GPUs (charged 8x: keep it short). Order = value: the tuning A/B on the strong-scaling problem (slot times, halo
    # and all-reduce waits, maxima over the ranks), the bit-exactness logs, config 4 at full size, N = 4, configs 3 and 5
    N=${2:-8}
    AB="off,stream,stream+noack,push+stream,red+stream,stream+pdlfa,folded,folded+stream,persistent"
    timeout 300 $TR --nproc-per-node $N --master-port 29550 scripts/scale_ab.py --axis 119 --out "$out/ab_n${N}_axis119.json" --variants $AB \
        > "$out/ab_n${N}_axis119.jsonl" 2> "$out/ab_n${N}_axis119.log"
    grep "^\[ab\]" "$out/ab_n${N}_axis119.log"
    export SB_SPIN_TIMEOUT_S=60
    timeout 400 python -m pytest tests/test_multigpu.py -m gpu -q -rA 2>&1 | tail -15 > "$out/pytest_multigpu_n$N.log"; cat "$out/pytest_multigpu_n$N.log"
    ( time timeout 200 $TR --nproc-per-node 4 --master-port 29555 tests/_dist_worker.py p2p ) > "$out/dist_worker_p2p_n4.log" 2>&1; tail -4 "$out/dist_worker_p2p_n4.log"
    unset SB_SPIN_TIMEOUT_S
    timeout 240 $TR --nproc-per-node 4 --master-port 29551 scripts/scale_ab.py --axis 119 --out "$out/ab_n4_axis119.json" --variants off,stream,persistent \
        > "$out/ab_n4_axis119.jsonl" 2> "$out/ab_n4_axis119.log"
    grep "^\[ab\]" "$out/ab_n4_axis119.log"
    timeout 240 $TR --nproc-per-node $N --master-port 29552 scripts/scale_ab.py --axis 119 --partition slab --out "$out/ab_n${N}_axis119_slab.json" --variants off,stream,persistent \
        > "$out/ab_n${N}_axis119_slab.jsonl" 2> "$out/ab_n${N}_axis119_slab.log"
    grep "^\[ab\]" "$out/ab_n${N}_axis119_slab.log"
    ;;
  h3)
    # eight GPUs, second session: lazy halo push against the best of the first session, METIS and slab partitions; the
    # bit-exactness worker at 8 ranks (it runs the lazy variants too); then config 4 at full size, configs 3 and 5
    N=${2:-8}
    timeout 300 $TR --nproc-per-node $N --master-port 29550 scripts/scale_ab.py --axis 119 --out "$out/ab_n${N}_axis119.json" \
        --variants stream+noack,lazy+stream,lazy+stream+pdlfa,lazy+stream+red,lazy,persistent > "$out/ab_n${N}_axis119.jsonl" 2> "$out/ab_n${N}_axis119.log"
    grep "^\[ab\]" "$out/ab_n${N}_axis119.log"
    timeout 240 $TR --nproc-per-node $N --master-port 29552 scripts/scale_ab.py --axis 119 --partition slab --out "$out/ab_n${N}_axis119_slab.json" \
        --variants stream,lazy+stream > "$out/ab_n${N}_axis119_slab.jsonl" 2> "$out/ab_n${N}_axis119_slab.log"
    grep "^\[ab\]" "$out/ab_n${N}_axis119_slab.log"
    export SB_SPIN_TIMEOUT_S=60
    ( time timeout 200 $TR --nproc-per-node $N --master-port 29555 tests/_dist_worker.py p2p ) > "$out/dist_worker_p2p_n$N.log" 2>&1; tail -4 "$out/dist_worker_p2p_n$N.log"
    unset SB_SPIN_TIMEOUT_S
    timeout 300 $TR --nproc-per-node $N --master-port 29556 scripts/config4_projection.py --axis 368 --lattice --steps 10 > "$out/config4_368_n${N}_lattice.json" 2> "$out/config4.err"
    tail -c 1500 "$out/config4_368_n${N}_lattice.json"; tail -3 "$out/config4.err"
    timeout 300 $TR --nproc-per-node $N --master-port 29558 scripts/config3_gmres.py --axis 119 --steps 100 > "$out/config3_gmres_fused_10M_n$N.json" 2> "$out/config3.err"
    timeout 300 $TR --nproc-per-node 4 --master-port 29557 scripts/config3_gmres.py --axis 119 --steps 100 > "$out/config3_gmres_fused_10M_n4.json" 2>> "$out/config3.err"
    tail -c 300 "$out/config3_gmres_fused_10M_n4.json"; tail -c 300 "$out/config3_gmres_fused_10M_n$N.json"; tail -3 "$out/config3.err"
    timeout 400 $TR --nproc-per-node $N --master-port 29559 scripts/apply_sweep.py --cells hexlat --sizes 1e7,1e8,2e8 --out "$out/apply_sweep_hexlat_n$N.json" > "$out/config5_n$N.log" 2>&1
    tail -4 "$out/config5_n$N.log"
    ;;
  h4)
    # eight GPUs, third session: lazy push with the batched forwarding loop; the communication-free floor of the same
    # partition (SB_DEBUG=6: no halo exchange, rank-local sums -- numbers only, the iterates are meaningless); N = 4
    N=${2:-8}
    timeout 300 $TR --nproc-per-node $N --master-port 29550 scripts/scale_ab.py --axis 119 --out "$out/ab_n${N}_axis119.json" \
        --variants stream+noack,lazy+stream > "$out/ab_n${N}_axis119.jsonl" 2> "$out/ab_n${N}_axis119.log"
    grep "^\[ab\]" "$out/ab_n${N}_axis119.log"
    SB_DEBUG=6 timeout 300 $TR --nproc-per-node $N --master-port 29551 scripts/scale_ab.py --axis 119 --out "$out/ab_n${N}_axis119_nocomm.json" \
        --variants off,stream > "$out/ab_n${N}_axis119_nocomm.jsonl" 2> "$out/ab_n${N}_axis119_nocomm.log"
    grep "^\[ab\]" "$out/ab_n${N}_axis119_nocomm.log"
    timeout 240 $TR --nproc-per-node $N --master-port 29552 scripts/scale_ab.py --axis 119 --partition slab --out "$out/ab_n${N}_axis119_slab.json" \
        --variants lazy+stream > "$out/ab_n${N}_axis119_slab.jsonl" 2> "$out/ab_n${N}_axis119_slab.log"
    grep "^\[ab\]" "$out/ab_n${N}_axis119_slab.log"
    timeout 240 $TR --nproc-per-node 4 --master-port 29553 scripts/scale_ab.py --axis 119 --out "$out/ab_n4_axis119.json" \
        --variants stream,lazy+stream > "$out/ab_n4_axis119.jsonl" 2> "$out/ab_n4_axis119.log"
    grep "^\[ab\]" "$out/ab_n4_axis119.log"
    ;;
  h2)
    # eight GPUs, second call: config 4 at full size, configs 3 and 5
    N=${2:-8}
    timeout 300 $TR --nproc-per-node $N --master-port 29556 scripts/config4_projection.py --axis 368 --lattice --steps 10 > "$out/config4_368_n${N}_lattice.json" 2> "$out/config4.err"
    tail -c 1200 "$out/config4_368_n${N}_lattice.json"; tail -3 "$out/config4.err"
    timeout 300 $TR --nproc-per-node $N --master-port 29558 scripts/config3_gmres.py --axis 119 --steps 100 > "$out/config3_gmres_fused_10M_n$N.json" 2>> "$out/config3.err"
    timeout 300 $TR --nproc-per-node 4 --master-port 29557 scripts/config3_gmres.py --axis 119 --steps 100 > "$out/config3_gmres_fused_10M_n4.json" 2> "$out/config3.err"
    tail -c 300 "$out/config3_gmres_fused_10M_n4.json"; tail -c 300 "$out/config3_gmres_fused_10M_n$N.json"
    timeout 400 $TR --nproc-per-node $N --master-port 29559 scripts/apply_sweep.py --cells hexlat --sizes 1e7,1e8,2e8 --out "$out/apply_sweep_hexlat_n$N.json" > "$out/config5_n$N.log" 2>&1
    tail -4 "$out/config5_n$N.log"
    ;;
esac
