set -x
python tests/stress_gpu.py 100 7 2>&1 | tail -4
compute-sanitizer --tool memcheck python tests/bringup_gpu.py pipe 2x_Compact_Pretrain 40 301 2>&1 | tail -6
compute-sanitizer --tool memcheck python tools/seg_check.py --n 2 --h 40 --w 300 --reps 1 2>&1 | tail -8
for th in 16 12 8; do B2SR_NLM_TH=$th python tools/nlm_pass.py 16 3 5 2>&1 | head -1; done
for th in 16 12 8; do B2SR_NLM_TH=$th python tools/nlm_pass.py 16 10 5 2>&1 | head -1; done
