set -x
python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/r3d_bench_n1.json 2> gpurun_out/r3d_bench_n1.err
tail -2 gpurun_out/r3d_bench_n1.err
python bench.py --config 5 --n-refine 13 --steps 2 > gpurun_out/r3d_cfg5.json 2> gpurun_out/r3d_cfg5.err
tail -2 gpurun_out/r3d_cfg5.err
