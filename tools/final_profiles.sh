mkdir -p gpurun_out
# 1. bounds-checked build: small cases + the parity suite (incl. the truncation tests)
( PGM_LIB=$PWD/photogrammetry_b200/libpgmatch_checked.so python tools/sanitize_cases.py; echo "exit code $?"; PGM_LIB=$PWD/photogrammetry_b200/libpgmatch_checked.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2; echo "exit code $?" ) > gpurun_out/r02_checked_build.log 2>&1
tail -4 gpurun_out/r02_checked_build.log
# 2. launch lists (ncu, cold caches, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_pair_8k.csv python tools/devprof.py 8192 U > gpurun_out/ncu_pair.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_l2_fp16.csv python tools/devbench_l2.py > gpurun_out/ncu_l2.log 2>&1
tail -3 gpurun_out/ncu_l2.log
# 3. ncu --set full of the tail kernel (single 8k pair) and of the float pair kernel
ncu --set full --clock-control none --import-source on -k regex:tail_kernel -s 6 -c 1 -o gpurun_out/r02_tail_full -f python tools/devprof.py 8192 U > gpurun_out/ncu_tail_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:l2_topk_pair_kernel -s 13 -c 1 -o gpurun_out/r02_l2_full -f python tools/devbench_l2.py > gpurun_out/ncu_l2_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
