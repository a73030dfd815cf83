timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s19_pytest.log 2>&1; tail -3 gpurun_out/s19_pytest.log | cut -c1-200
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s19_bench.json 2> gpurun_out/s19_bench.err; tail -2 gpurun_out/s19_bench.err | cut -c1-200; python -c "
import json;d=json.load(open('gpurun_out/s19_bench.json'));print({k:d.get(k) for k in ['value','ms_per_step','e2e','gpu_launches','first_pass','full_pass']})"
timeout 600 python bench.py --no-cpu-baseline --paired --reads 500000 --steps 2 > gpurun_out/s19_bench_paired.json 2> gpurun_out/s19_bench.err; python -c "
import json;d=json.load(open('gpurun_out/s19_bench_paired.json'));print('paired', {k:d.get(k) for k in ['value','ms_per_step','e2e','first_pass','full_pass']})"
