"""Reference-time setup at a given size: device index build, duplication table (device scan vs host detector)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from mapper_b200 import capi, synth

total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 250_000_000
n_contigs = int(sys.argv[2]) if len(sys.argv) > 2 else 50
ref = synth.random_reference(total, seed=5, n_contigs=n_contigs, repeat_fraction=0.05, repeat_len=(300, 3000))
g = capi.XMapper(synth.DEFAULT_PARAMS, device=0)
ref = sorted(ref, key=lambda x: len(x[1]))   # the reference sorts its contigs by length
g.set_reference([synth.pack_contig(c) for _, c in ref], [len(c) for _, c in ref])
t0 = time.time(); g.build_index(150); t1 = time.time()
print("index build (device): %.2f s" % (t1 - t0))
for rep in range(2):
    t0 = time.time(); g.build_duplications(-1, -1, 2, 1000); t1 = time.time()
    print("duplications, device scan + host merge: %.3f s" % (t1 - t0))
dev = [g.get_duplications(c).copy() for c in range(n_contigs)]
t0 = time.time(); g.build_duplications(-1, -1, 2, 1000, host=True); t1 = time.time()
print("duplications, host detector: %.3f s" % (t1 - t0))
same = all(np.array_equal(dev[c], g.get_duplications(c)) for c in range(n_contigs))
print("tables equal:", same, " starts:", sum(len(d) for d in dev))
