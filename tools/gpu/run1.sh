echo "== unrolled"; XM_LIB_PATH=$PWD/mapper_b200/libxm_unroll.so timeout 300 python tools/probe_qcycles.py --reads 1000000 2>&1 | sed -n 1,4p | cut -c1-220
echo "== default"; timeout 300 python tools/probe_qcycles.py --reads 1000000 2>&1 | sed -n 1,4p | cut -c1-220
