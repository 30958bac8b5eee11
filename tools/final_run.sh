# end-of-round bundle on one B200 (run through gpurun from the repo root); outputs under gpurun_out/
set -x
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/r02g_gputests.log 2>&1; tail -3 gpurun_out/r02g_gputests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02g_bench_c3.json 2> gpurun_out/r02g_bench_c3.err; tail -c 400 gpurun_out/r02g_bench_c3.json
timeout 200 python bench.py --workload C4 --dense-eta 0.23 --steps 5 --warmup 3 --no-cpu --no-dropin --others NONE 2>/dev/null | tail -1 > gpurun_out/r02g_bench_c4_dense.json
timeout 200 python bench.py --workload C3 --dense-eta 0.25 --steps 20 --warmup 5 --no-cpu --no-dropin --others NONE 2>/dev/null | tail -1 > gpurun_out/r02g_bench_c3_dense.json
timeout 200 python tools/perf_probe.py C3_1024,C3_2048,C3_4096,C2,C2_2048,C4_512,C5_256 0 > gpurun_out/r02g_probe.jsonl 2>&1; cat gpurun_out/r02g_probe.jsonl | cut -c1-100
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02g_launches_c3.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-dropin --others NONE > gpurun_out/r02g_ncu_launches.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:walk_team -s 1 -c 1 -o gpurun_out/r02g_c4_team4 -f python tools/perf_probe.py C4_512 0 10 > gpurun_out/r02g_ncu_c4.log 2>&1
ls -la gpurun_out/*.ncu-rep
