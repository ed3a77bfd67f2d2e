set -x
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r3h_bench_n1.json 2> gpurun_out/r3h_bench_n1.err
tail -2 gpurun_out/r3h_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r3h_bench_n1.json').read())
print(d['value'], d['ms_per_step'], d['solver']['iterations'][:10], d['parity_check']['ok'], d['solution_check']['ok'], d['solution_check']['u_rel_l2'], d['solution_check']['p_rel_l2'])"
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4
