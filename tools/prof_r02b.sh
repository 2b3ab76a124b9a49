cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_gputests.log 2>&1; tail -3 gpurun_out/r02b_gputests.log
timeout 500 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -c 600 gpurun_out/r02b_bench.json
B="python bench.py --steps 2 --warmup 3 --no-extras --sets 256"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02b_ncu_launches.csv $B > gpurun_out/r02b_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_kmp_predict_tab|k_potrf_wpp|k_trtri_wpp|k_kernel_build_comp" -s 4 -c 4 -o gpurun_out/r02b_hot python tools/pipeline_once.py > gpurun_out/r02b_ncu_pipe.log 2>&1
ls -la gpurun_out/r02b*
