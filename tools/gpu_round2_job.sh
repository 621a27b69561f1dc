set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02f_gpu_tests.log 2>&1; tail -4 gpurun_out/r02f_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; tail -3 gpurun_out/r02f_smoke.log
python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; head -c 300 gpurun_out/r02f_bench.json; echo
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:tc_pipe_kernel -s 2 -c 1 --csv --log-file gpurun_out/r02_pipe_traffic.csv python tools/profile_pass.py --frames 16 --passes 4 > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s 354 -c 354 --csv --log-file gpurun_out/r02_valar_traffic.csv python tools/valar_batch.py --passes 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_launches_bench.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_pipe_kernel -s 2 -c 1 -o gpurun_out/r02_pipe_full python tools/profile_pass.py --frames 16 --passes 4 > /dev/null 2>&1
ls -la gpurun_out/ | tail -8
