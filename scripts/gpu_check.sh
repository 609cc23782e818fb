#!/bin/bash
# Run on the GPU box via gpurun: every stage has its own timeout so a hung kernel cannot eat the lease.
# usage: scripts/gpu_check.sh [stage ...]   (default: all test stages)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/smi.txt 2>&1
run() { # name timeout cmd...
  local name=$1; local to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc" | tee -a gpurun_out/summary.txt
  tail -n ${TAILN:-12} "gpurun_out/$name.log" | cut -c1-600 | tee -a gpurun_out/summary.txt
}
PT="python -m pytest -q --tb=short -p no:cacheprovider -m gpu"
STAGES=${@:-"tests smoke"}
for st in $STAGES; do
  case $st in
    tests) run tests 1500 $PT tests ;;
    smoke) run smoke 300 python -c "import __graft_entry__ as g; g.smoke()" ;;
    benchquick) run benchquick 600 python bench.py --sde-steps 100 --steps 1 --warmup 1 --no-cpu-baseline --cd-clouds 256 ;;
    bench) run bench 900 python bench.py ;;
    benchref) run benchref 600 python bench.py --impl reference --steps 1 --warmup 0 ;;
    launches) run launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --sde-steps 2 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-eager --cd-clouds 16 --no-secondary ;;
    ncufull)
      for k in ${NCU_KERNELS:-gemm_tc2_kernel qkv_attention_kernel layernorm_mod_kernel sde_step_kernel attention_tc_kernel}; do
        run ncu_$k 400 ncu --set full --clock-control none --import-source on -k regex:$k -s ${NCU_SKIP:-6} -c 2 -f -o gpurun_out/prof_$k \
          python bench.py --sde-steps 2 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-eager --cd-clouds 16 --no-secondary
      done ;;
    ncuemd) run ncu_emd 400 ncu --set full --clock-control none --import-source on -k regex:approx_match_kernel -c 1 -f -o gpurun_out/prof_approx_match_kernel \
          python bench.py --sde-steps 2 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-eager --cd-clouds 16 --emd-clouds 13 --completion-batch 8 ;;
    ncucd) run ncu_pairwise_cd 400 ncu --set full --clock-control none --import-source on -k regex:pairwise_cd_kernel -s 2 -c 1 -f -o gpurun_out/prof_pairwise_cd_kernel \
          python bench.py --sde-steps 2 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-eager --cd-clouds 128 --no-secondary ;;
    *) echo "unknown stage $st" ;;
  esac
done
