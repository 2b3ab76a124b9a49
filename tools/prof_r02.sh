# Round-2 profiling recipe: launch list of the bench step + ncu --set full of the hot kernels (run under gpurun, 1 GPU).
set -x
cd $GRAFT_REPO_ROOT
B="python bench.py --steps 2 --warmup 3 --no-extras --sets 256"
$B > gpurun_out/prof_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_ncu_launches.csv $B > gpurun_out/prof_ncu_bench.log 2>&1
python tools/pipeline_once.py 79 > gpurun_out/prof_plain_pipe.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_kmp_predict_fast|k_potrf_tma|k_trtri_tma|k_kernel_build_comp|k_solve_alpha" -s 5 -c 5 -o gpurun_out/r02_hot python tools/pipeline_once.py 79 > gpurun_out/prof_ncu_pipe.log 2>&1
python tools/dgemm_once.py > gpurun_out/prof_plain_dgemm.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_dgemm_nt_tma -s 2 -c 2 -o gpurun_out/r02_dgemm python tools/dgemm_once.py > gpurun_out/prof_ncu_dgemm.log 2>&1
python tools/em_gmr_once.py 2000000 > gpurun_out/prof_plain_em.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_gmm_em -s 2 -c 2 -o gpurun_out/r02_em python tools/em_gmr_once.py 2000000 > gpurun_out/prof_ncu_em.log 2>&1
python tools/gmr_once.py > gpurun_out/prof_plain_gmr.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_gmr -s 2 -c 2 -o gpurun_out/r02_gmr2 python tools/gmr_once.py > gpurun_out/prof_ncu_gmr.log 2>&1
tail -2 gpurun_out/prof_plain_pipe.log gpurun_out/prof_plain_dgemm.log gpurun_out/prof_ncu_em.log
ls -la gpurun_out/*.ncu-rep
