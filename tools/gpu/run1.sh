timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s12_pytest.log 2>&1; tail -3 gpurun_out/s12_pytest.log
for s in 1 0; do XM_SORT_HARD=$s timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s12_bench_$s.json 2> gpurun_out/s12_bench.err; tail -3 gpurun_out/s12_bench.err; python -c "
import json;d=json.load(open('gpurun_out/s12_bench_$s.json'));print($s, {k:d[k] for k in ['value','ms_per_step','e2e','gpu_launches','first_pass','full_pass']})"; done
