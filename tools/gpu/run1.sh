XM_HOST_TIMES=1 timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 2 2>&1 >/dev/null | grep "\[xm\]" | tail -8
