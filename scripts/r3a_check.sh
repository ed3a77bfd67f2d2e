set -x
python -m pytest tests/test_gpu_scale.py -x -q -k "operator_is_rebuilt or bench_tolerance" 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r3a_bench_n1.json 2> gpurun_out/r3a_bench_n1.err
python bench.py --steps 20 --warmup 5 --no-cpu --no-checks --rebuild-operator > gpurun_out/r3a_bench_n1_rebuild.json 2>> gpurun_out/r3a_bench_n1.err
PE_TIMING=1 python bench.py --steps 10 --warmup 0 --no-cpu --no-checks --solver-options 65 > /dev/null 2> gpurun_out/r3a_phase_timing_n1.txt
python bench.py --config 4 --steps 10 --no-cpu > gpurun_out/r3a_cfg4_fgmres.json 2>> gpurun_out/r3a_bench_n1.err
tail -3 gpurun_out/r3a_bench_n1.err
