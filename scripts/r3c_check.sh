set -x
python -m pytest tests/test_gpu_esm.py -x -q 2>&1 | tail -4
python scripts/bench_cfg5.py 12 2 0.2 1048576 0 > gpurun_out/r3c_cfg5_q2rt2_dmma.json 2> gpurun_out/r3c.err
cat gpurun_out/r3c_cfg5_q2rt2_dmma.json
ncu --set full --clock-control none --import-source on -k regex:k_esm_cell_mma -s 1 -c 1 -o gpurun_out/r3c_k_esm_cell_mma -f python scripts/bench_cfg5.py 10 2 0.2 1048576 0 > gpurun_out/r3c_ncu.log 2>&1
tail -3 gpurun_out/r3c.err gpurun_out/r3c_ncu.log
