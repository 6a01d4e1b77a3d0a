# round-2 measurement suite (one GPU): tests, bench, reference arm, launch list, ncu captures
set -x
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r02k_pytest.log
timeout 900 python bench.py > gpurun_out/r02k_bench_1gpu.json 2> gpurun_out/r02k_bench_1gpu.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02k_bench_reference.json 2> gpurun_out/r02k_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02k_launches_profile5.csv python bench.py --workload profile5 --steps 2 --warmup 1 --no_e2e --no_cpu_baseline --no_api_e2e > gpurun_out/r02k_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_tc_pass_a|k_tc_pass_b|k_loo_gram_mma' --launch-skip 16 -c 6 -o gpurun_out/r02k_full -f python tools/passb_probe.py profile5 1 > gpurun_out/r02k_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_tc_pass_a|k_tc_pass_b2' --launch-skip 6 -c 2 -o gpurun_out/r02k_full_tiled -f python tools/pa_probe.py > gpurun_out/r02k_ncu_full_tiled.log 2>&1
tail -2 gpurun_out/r02k_pytest.log; ls gpurun_out | grep r02k | head -20
