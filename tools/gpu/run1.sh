for b in 512 768; do echo "== block $b"; XM_FULL_WARPS=$((b/32)) XM_LIB_PATH=$PWD/mapper_b200/libxm_$b.so timeout 300 python tools/probe_qcycles.py --reads 1000000 2>&1 | sed -n 2,2p; done
