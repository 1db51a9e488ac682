set -x
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_cfg2.csv python scripts/ncu_target.py cfg2 3 > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_cost_fused2|k_cost_sum2|k_cost_quad2|k_sort_pass|k_segment|k_jtj_dmma" -s 9 -c 9 -o gpurun_out/prof_r02_cost python scripts/ncu_target.py cfg2 2 > gpurun_out/prof_r02_cost.log 2>&1
timeout 300 python scripts/fma_deviation.py cfg2 2>&1 | tail -2
timeout 500 compute-sanitizer --tool memcheck python scripts/sanitize_target.py tiny > gpurun_out/r02_sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/r02_sanitizer_memcheck.log
timeout 500 compute-sanitizer --tool racecheck python scripts/sanitize_target.py tiny > gpurun_out/r02_sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/r02_sanitizer_racecheck.log
