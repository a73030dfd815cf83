timeout 900 python bench.py > gpurun_out/r1l_bench_n1.json 2> gpurun_out/r1l_bench_n1.err; tail -3 gpurun_out/r1l_bench_n1.err; python -c "
import json;d=json.load(open('gpurun_out/r1l_bench_n1.json'));print({k:d.get(k) for k in ['value','ms_per_step','e2e','gpu_launches','parity','cpu_baseline','clocks']}); print(d['roofline'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1l_bench_reference_arm.json 2>/dev/null; cut -c1-400 gpurun_out/r1l_bench_reference_arm.json
