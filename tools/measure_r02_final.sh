# final round-2 one-GPU run: GPU tests, the contract bench line, the reference arm (kernels unchanged since the r02k ncu captures)
set -x
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r02n_pytest.log
timeout 600 python bench.py > gpurun_out/r02n_bench_1gpu.json 2> gpurun_out/r02n_bench_1gpu.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02n_bench_reference.json 2> gpurun_out/r02n_bench_reference.err
tail -2 gpurun_out/r02n_pytest.log; tail -c 400 gpurun_out/r02n_bench_1gpu.json; tail -c 300 gpurun_out/r02n_bench_reference.json
