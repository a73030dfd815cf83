timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s13_pytest.log 2>&1; tail -3 gpurun_out/s13_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s13_bench.json 2> gpurun_out/s13_bench.err; tail -3 gpurun_out/s13_bench.err; python -c "
import json;d=json.load(open('gpurun_out/s13_bench.json'));print({k:d[k] for k in ['value','ms_per_step','e2e','gpu_launches','first_pass','full_pass']})"
XM_QCYCLES=1 timeout 300 python tools/probe_qcycles.py --reads 1000000 2>&1 | sed -n 4,5p
