for b in 512 1024; do echo "== easy block $b"; XM_LIB_PATH=$PWD/mapper_b200/libxm_e$b.so timeout 300 python tools/probe_qcycles.py --reads 1000000 2>&1 | sed -n 1,2p | cut -c1-160; done
echo "== default"; timeout 300 python tools/probe_qcycles.py --reads 1000000 2>&1 | sed -n 1,2p | cut -c1-160
