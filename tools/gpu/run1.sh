timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s17_pytest.log 2>&1; tail -2 gpurun_out/s17_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s17_bench.json 2> gpurun_out/s17_bench.err; tail -2 gpurun_out/s17_bench.err; python -c "
import json;d=json.load(open('gpurun_out/s17_bench.json'));print({k:d.get(k) for k in ['value','ms_per_step','e2e','gpu_launches','first_pass','full_pass','index_build_s']})"
