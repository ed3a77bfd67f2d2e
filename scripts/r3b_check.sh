set -x
python -m pytest tests/test_gpu_esm.py -x -q 2>&1 | tail -8
python scripts/bench_cfg5.py 12 2 0.2 1048576 0 > gpurun_out/r3b_cfg5_q2rt2_dmma.json 2> gpurun_out/r3b.err
python scripts/bench_cfg5.py 12 2 0.2 1048576 1 > gpurun_out/r3b_cfg5_q2rt2_dfma.json 2>> gpurun_out/r3b.err
cat gpurun_out/r3b_cfg5_q2rt2_dmma.json gpurun_out/r3b_cfg5_q2rt2_dfma.json
tail -3 gpurun_out/r3b.err
