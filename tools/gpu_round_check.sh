#!/bin/bash
# GPU-box check of the denoise pass: parity tests, a launch list and one full ncu capture of nlm_kernel.
# usage (from the repo root, through gpurun): bash tools/gpu_round_check.sh nlm | full
set -x
mkdir -p gpurun_out
if [ "$1" = "nlm" ]; then
  (time timeout 400 python -m pytest tests/test_gpu_nlmeans.py -m gpu -x -q) > gpurun_out/r01n_nlm_tests.log 2>&1
  echo "nlm tests rc=$?" >> gpurun_out/r01n_nlm_tests.log
  tail -15 gpurun_out/r01n_nlm_tests.log
  timeout 120 python tools/nlm_pass.py 16 3 5 > gpurun_out/r01n_nlm_pass.log 2>&1
  B2SR_NLM_PACKED=0 timeout 120 python tools/nlm_pass.py 16 3 5 > gpurun_out/r01n_nlm_pass_generic.log 2>&1
  timeout 120 python tools/nlm_pass.py 16 10 5 > gpurun_out/r01n_nlm_pass_l10.log 2>&1
  for th in 12 8; do B2SR_NLM_TH=$th timeout 120 python tools/nlm_pass.py 16 3 5 > gpurun_out/r01n_nlm_pass_th$th.log 2>&1; done
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:nlm_kernel -c 1 -o gpurun_out/r01n_nlm_kernel python tools/nlm_pass.py 4 3 2 > gpurun_out/r01n_ncu_full.log 2>&1
  timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 50 --csv --log-file gpurun_out/r01n_nlm_launches.csv python tools/nlm_pass.py 16 3 3 > /dev/null 2>&1
  tail -n 3 gpurun_out/r01n_nlm_pass*.log gpurun_out/r01n_ncu_full.log
else
  (time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r01o_gpu_tests.log 2>&1
  echo "gpu tests rc=$?" >> gpurun_out/r01o_gpu_tests.log
  tail -8 gpurun_out/r01o_gpu_tests.log
  timeout 600 python bench.py > gpurun_out/r01o_bench.json 2> gpurun_out/r01o_bench.err
  cat gpurun_out/r01o_bench.json
  timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01o_smoke.log 2>&1; tail -2 gpurun_out/r01o_smoke.log
  B2SR_E2E_TAPER=0 timeout 300 python bench.py --no-cpu-baseline --no-extra --steps 6 > gpurun_out/r01o_bench_equal_chunks.json 2>/dev/null
  tail -c 600 gpurun_out/r01o_bench_equal_chunks.json
fi
