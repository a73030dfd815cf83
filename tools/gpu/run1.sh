timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s18_pytest.log 2>&1; tail -25 gpurun_out/s18_pytest.log | cut -c1-250
