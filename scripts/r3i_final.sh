set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r3i_bench_n1.json 2> gpurun_out/r3i_bench_n1.err
tail -2 gpurun_out/r3i_bench_n1.err
python -m pytest tests/test_gpu_scale.py -x -q 2>&1 | tail -6
