set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; echo rc=$?
python bench.py --config 4 --n-refine 10 --steps 10 > gpurun_out/r2z_cfg4_fgmres.json 2>/dev/null
python bench.py --config 4 --n-refine 10 --steps 10 --krylov 0 > gpurun_out/r2z_cfg4_minres.json 2>/dev/null
PE_TIMING=1 python bench.py --steps 4 --warmup 1 --no-checks --no-cpu --solver-options 65 > /dev/null 2> gpurun_out/r2z_phase_timing_n1.txt
python scripts/sweep_asm_modes.py 10 > gpurun_out/r2z_asm_modes.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-checks > gpurun_out/r2z_ncu_list.log 2>&1
for k in k_spmv_stream k_sub_stream k_mg_jacobi k_assemble_interior3 k_multi_dot k_gs_update; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/r2z_$k -f python bench.py --steps 1 --warmup 1 --no-cpu --no-checks > gpurun_out/r2z_ncu_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:k_assemble_gatherB2 -s 3 -c 1 -o gpurun_out/r2z_k_assemble_gatherB2 -f python scripts/profile_gather.py 1 > gpurun_out/r2z_ncu_gatherB2.log 2>&1
ls -la gpurun_out/r2z_*
