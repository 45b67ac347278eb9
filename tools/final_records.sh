set -x
timeout 500 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu.log 2>&1; tail -2 gpurun_out/r02_pytest_gpu.log
SCSPLIT_B200_LIB=$PWD/scsplit_b200/libscsplit_b200_bounds.so timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu_bounds_build.log 2>&1; tail -2 gpurun_out/r02_pytest_gpu_bounds_build.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 600 gpurun_out/r02_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 12 --warmup 4 --quick > gpurun_out/r02_ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_estep_lk|k_bayes' -s 12 -c 2 -o gpurun_out/r02_lk_final python bench.py --steps 12 --warmup 4 --quick > gpurun_out/r02_ncu_full.log 2>&1
ls -la gpurun_out/r02_lk_final.ncu-rep gpurun_out/r02_launches.csv
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
