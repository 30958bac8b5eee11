set -x
timeout 200 python bench.py > gpurun_out/r01b_bench_c2.log 2>&1
timeout 120 python bench.py --workload C3 --dense-eta 0.25 --steps 60 --warmup 10 --no-cpu > gpurun_out/r01b_bench_c3.log 2>&1
timeout 120 python bench.py --workload C4 --dense-eta 0.25 --steps 20 --warmup 5 --no-cpu > gpurun_out/r01b_bench_c4.log 2>&1
timeout 120 python bench.py --workload C5 --dense-eta 0.25 --steps 20 --warmup 5 --no-cpu > gpurun_out/r01b_bench_c5.log 2>&1
timeout 150 ncu --set full --import-source on --clock-control none -k regex:walk_team -s 1 -c 1 -o gpurun_out/prof_team28_c5_final -f python tools/perf_probe.py C5_256 0 > gpurun_out/ncu_team28_final.log 2>&1
timeout 150 ncu --set full --import-source on --clock-control none -k regex:walk_team -s 1 -c 1 -o gpurun_out/prof_team24_c4_final -f python tools/perf_probe.py C4_512 0 > gpurun_out/ncu_team24_final.log 2>&1
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c5_team.csv python bench.py --workload C5 --steps 4 --warmup 3 --no-cpu > gpurun_out/ncu_launches_c5.log 2>&1
tail -c 600 gpurun_out/r01b_bench_c5.log
