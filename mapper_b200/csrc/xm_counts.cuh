// xmapper_b200 — count accumulation for --out-vcf / --out-mutations on the device.
// Included by xm_capi.cu after BatchD / OutArena are defined.
//
// What MatchDatabase -> Alignments -> AlignmentsSection -> RegionAlignments -> DirectionalAlignments keep per reference position
// (QV/MatchDatabase.java:16-59, QV/Alignments.java:89-156, QV/DirectionalAlignments.java:20-96) is split in two here:
//   * dense planes  int32 [contig][region: 0 middle, 1 end][dir: 0 forward, 1 reverse][pos]   referenceCounts (:26-28)
//   * a sparse table of variants: one entry per (position, region, direction, insertion column or "at the position", allele) with the
//     summed scaled weight (:29-38, :41-55) and the example read the reference would have kept (betterExample :63-96, a total order:
//     longer read, closer to the read's middle, earlier position, later name, smaller id - so the winner is independent of arrival order)
// xm_counts_kernel walks every sequence alignment of every choice once and feeds both.  The table is kept as raw records that are
// radix-sorted by key and reduced by key (sum, best example) whenever they are read, merged with other GPUs or outgrow their buffer.
#pragma once
#include <cub/device/device_reduce.cuh>

struct CountsD {
  int32_t* planes;            // [contig_off[c]*4 + ((region*2+dir)*len + pos)]
  const int64_t* contig_off;  // prefix of contig lengths
  double end_fraction;        // MatchDatabase.queryEndFraction
};
// key = (contig_off[contig] + pos) << 21 | region << 20 | dir << 19 | (insertion column + 1) << 3 | allele   (allele: index into "ACGTN-")
struct VarRec {
  unsigned long long key;
  int32_t count;       // sum of (int)(weight * 100)
  int32_t ex_len;      // example read: length,
  int32_t ex_delta;    //   | |index| - length / 2 |,
  int32_t ex_index;    //   index in the read (negative: deletion, DirectionalAlignments.java:33-35),
  long long ex_order;  //   host-supplied name order key (larger = later name), 0 when the host gave none,
  long long ex_gid;    //   global sequence id << 1 | reversed ("-rev" view)
};
struct VarOut {
  VarRec* recs; unsigned long long* n; unsigned long long cap;   // raw records are appended at recs[*n]; *n keeps counting past cap
  const int64_t* order;   // per sequence of this batch, or nullptr
  long long first_gid;    // global id of sequence 0 of this batch
};
__device__ __forceinline__ bool var_better(const VarRec& a, const VarRec& b) {  // true: b replaces a (betterExample :63-96)
  if (a.ex_len != b.ex_len) return b.ex_len > a.ex_len;
  if (a.ex_index != b.ex_index) {
    if (a.ex_delta != b.ex_delta) return b.ex_delta < a.ex_delta;
    return b.ex_index < a.ex_index;
  }
  if (a.ex_order != b.ex_order) return a.ex_order < b.ex_order;
  return (b.ex_gid >> 1) < (a.ex_gid >> 1);
}
struct VarMerge {
  __device__ __forceinline__ VarRec operator()(const VarRec& a, const VarRec& b) const {
    VarRec r = var_better(a, b) ? b : a;
    r.key = a.key; r.count = a.count + b.count;
    return r;
  }
};
__device__ __forceinline__ void var_emit(const VarOut& V, unsigned long long key, int32_t count, int qlen, int index, long long sid, int reversed) {
  const unsigned long long at = atomicAdd(V.n, 1ull);
  if (at >= V.cap) return;   // counted, not stored: the host grows the buffer and re-runs the batch's accumulation
  VarRec r;
  r.key = key; r.count = count; r.ex_len = qlen;
  const int ai = index < 0 ? -index : index;
  const int d = ai - qlen / 2;
  r.ex_delta = d < 0 ? -d : d; r.ex_index = index;
  r.ex_order = V.order ? V.order[sid] : 0;
  r.ex_gid = ((V.first_gid + sid) << 1) | (long long)(reversed ? 1 : 0);
  V.recs[at] = r;
}
// One thread per query.  All weights are Java floats: weight = 1f / numChoices (MatchDatabase.groupByReference :40) times
// 1f / numAlignmentsCoveringIndexB (WeightedAlignment.getWeight :19-28, QueryAlignment.java:97-120,203-214).
template <bool PLANES>
__global__ void xm_counts_kernel(RefD ref, BatchD batch, OutArena out, CountsD C, VarOut V, int n_queries) {
  int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= n_queries) return;
  const OutQuery& oq = out.q[qi];
  if (oq.status != 0) return;
  long long s0 = batch.first_seq[qi];
  for (int comp = 0; comp < oq.n_comp; comp++) {
    int nch = oq.n_choice[comp];
    if (nch < 1) continue;
    const float weight = 1.0f / (float)nch;
    for (int k = 0; k < nch; k++) {
      const OutChoice& ch = out.choices[oq.choice_first[comp] + k];
      // QueryAlignment.computeOverlap :203-214 (alignments found by the aligner are reference-contiguous)
      int min_overlap = -1, max_overlap = -1;
      for (int s = 0; s < ch.n_sa; s++) {
        const OutSA& o = out.sas[ch.sa_first + s];
        const int32_t* ob = out.blocks + 4 * o.block_first;
        int mn = ob[1], mx = ob[4 * (o.n_blocks - 1) + 1] + ob[4 * (o.n_blocks - 1) + 3];
        if (min_overlap < 0 || mn >= min_overlap) min_overlap = mn;
        if (max_overlap < 0 || mx <= max_overlap) max_overlap = mx;
      }
      for (int s = 0; s < ch.n_sa; s++) {
        const OutSA& sa = out.sas[ch.sa_first + s];
        const int mate = (oq.n_comp == 2) ? comp : s;
        const long long sid = s0 + mate;
        SeqView qv; qv.w = batch.packed + batch.seq_word_off[sid]; qv.len = batch.seq_len[sid]; qv.rc = sa.reversed; qv.bytes = nullptr;
        SeqView rv = ref.contig(sa.contig, 0);
        const int32_t* bl0 = out.blocks + 4 * sa.block_first;
        const int first_start_a = bl0[0];
        const int last_end_a = bl0[4 * (sa.n_blocks - 1)] + bl0[4 * (sa.n_blocks - 1) + 2];
        const double end_limit = (double)qv.len * C.end_fraction;  // Alignments.isNearQueryEnd :153-156
        const long long base = C.contig_off[sa.contig] * 4;
        const unsigned long long gbase = (unsigned long long)C.contig_off[sa.contig];
        const int dir = sa.reversed ? 1 : 0;
        auto scaled_at = [&](int rb) -> int32_t {
          int num = ch.n_sa;
          if (ch.n_sa >= 2 && (rb < min_overlap || rb >= max_overlap)) num = 1;  // getNumAlignmentsCoveringIndexB :97-111
          const float pos_w = (num != 0) ? 1.0f / (float)num : 0.0f;
          const float wgt = weight * pos_w;
          return (int32_t)(wgt * 100.0f);
        };
        auto region_at = [&](int qa) -> int { const int dist = min(qa - first_start_a, last_end_a - qa - 1); return ((double)dist < end_limit) ? 1 : 0; };
        auto key_of = [&](int rb, int region, int ins_plus_1, int allele) -> unsigned long long {
          return ((gbase + (unsigned long long)rb) << 21) | ((unsigned long long)region << 20) | ((unsigned long long)dir << 19) | ((unsigned long long)ins_plus_1 << 3) | (unsigned long long)allele;
        };
        for (int b = 0; b < sa.n_blocks; b++) {
          const int32_t* bl = bl0 + 4 * b;
          const int a0 = bl[0], b0 = bl[1], al = bl[2], blen = bl[3];
          if (al == blen) {                              // Alignments.addMatchOnSequence :104-119
            for (int i = 0; i < al; i++) {
              const int qa = a0 + i, rb = b0 + i;
              const uint8_t code = qv.at(qa);
              if (bp_is_ambiguous(code)) continue;       // DirectionalAlignments.add :21-25
              const int region = region_at(qa);
              if (code == rv.at(rb)) { if (PLANES) atomicAdd(&C.planes[base + (long long)(region * 2 + dir) * rv.len + rb], scaled_at(rb)); }
              else if (V.recs) var_emit(V, key_of(rb, region, 0, code == 1 ? 0 : code == 2 ? 1 : code == 4 ? 2 : code == 8 ? 3 : 5), scaled_at(rb), qv.len, code == 0 ? -qa : qa, sid, dir);
            }
          } else if (al > blen) {                        // insertion :120-132: column i of the insertion after reference position b0 - 1
            if (!V.recs || b0 < 1) continue;
            const int region = region_at(a0);
            const int32_t sc = scaled_at(b0 - 1);
            for (int i = 0; i < al && i < 65534; i++) {
              const uint8_t code = qv.at(a0 + i);
              const int allele = code == 1 ? 0 : code == 2 ? 1 : code == 4 ? 2 : code == 8 ? 3 : code == 0 ? 5 : 4;   // ambiguous -> 'N' :46-47
              var_emit(V, key_of(b0 - 1, region, i + 1, allele), sc, qv.len, a0 + i, sid, dir);
            }
          } else if (V.recs) {                           // deletion :133-149: a '-' at every deleted reference position
            for (int i = 0; i < blen; i++) {
              const int rb = b0 + i;
              if (rv.at(rb) == 0) { if (PLANES) atomicAdd(&C.planes[base + (long long)(region_at(a0 + i) * 2 + dir) * rv.len + rb], scaled_at(rb)); continue; }  // a '-' in the reference equals the dash (:26-28)
              var_emit(V, key_of(rb, region_at(a0 + i), 0, 5), scaled_at(rb), qv.len, -a0, sid, dir);
            }
          }
        }
      }
    }
  }
}

// sorts recs[0, n) by key and reduces equal keys (VarMerge); the result lands in recs[0, *n_out).  tmp buffers are grown as needed.
struct VarScratch { DevBuf keys_a, keys_b, idx_a, idx_b, gathered, out_keys, n_out, tmp; };
__global__ void xm_var_keys_kernel(const VarRec* recs, unsigned long long n, unsigned long long* keys, uint32_t* idx) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { keys[i] = recs[i].key; idx[i] = (uint32_t)i; }
}
__global__ void xm_var_gather_kernel(const VarRec* recs, const uint32_t* idx, unsigned long long n, VarRec* out) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = recs[idx[i]];
}
static bool var_sort_reduce(VarRec* recs, unsigned long long n, VarScratch& S, cudaStream_t st, unsigned long long* n_out_host, std::string& err) {
  *n_out_host = 0;
  if (n == 0) return true;
  if (n >= (1ull << 31)) { err = "variant table: more than 2^31 raw records"; return false; }
  const int ni = (int)n;
  size_t t1 = 0, t2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t1, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, ni, 0, 64, st);
  cub::DeviceReduce::ReduceByKey(nullptr, t2, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const VarRec*)nullptr, (VarRec*)nullptr, (int*)nullptr, VarMerge(), ni, st);
  const size_t tb = (t1 > t2 ? t1 : t2) + 256;
  if (!S.keys_a.ensure(n * 8) || !S.keys_b.ensure(n * 8) || !S.idx_a.ensure(n * 4) || !S.idx_b.ensure(n * 4) || !S.gathered.ensure(n * sizeof(VarRec)) || !S.out_keys.ensure(n * 8) ||
      !S.n_out.ensure(16) || !S.tmp.ensure(tb)) { err = "out of device memory (variant table reduce)"; return false; }
  const unsigned blocks = (unsigned)((n + 255) / 256);
  xm_var_keys_kernel<<<blocks, 256, 0, st>>>(recs, n, (unsigned long long*)S.keys_a.p, (uint32_t*)S.idx_a.p);
  size_t q = tb;
  if (cub::DeviceRadixSort::SortPairs(S.tmp.p, q, (const unsigned long long*)S.keys_a.p, (unsigned long long*)S.keys_b.p, (const uint32_t*)S.idx_a.p, (uint32_t*)S.idx_b.p, ni, 0, 64, st) != cudaSuccess) { err = "variant table sort failed"; return false; }
  xm_var_gather_kernel<<<blocks, 256, 0, st>>>(recs, (const uint32_t*)S.idx_b.p, n, (VarRec*)S.gathered.p);
  q = tb;
  if (cub::DeviceReduce::ReduceByKey(S.tmp.p, q, (const unsigned long long*)S.keys_b.p, (unsigned long long*)S.out_keys.p, (const VarRec*)S.gathered.p, recs, (int*)S.n_out.p, VarMerge(), ni, st) != cudaSuccess) { err = "variant table reduce failed"; return false; }
  int n_out = 0;
  if (cudaMemcpyAsync(&n_out, S.n_out.p, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { err = std::string("variant table reduce: ") + cudaGetErrorString(cudaGetLastError()); return false; }
  *n_out_host = (unsigned long long)n_out;
  return true;
}
