// xmapper_b200 — host-side result container: turns the device result arena (per-query records + bump-allocated
// choices / sequence alignments / blocks) into the CSR arrays documented in include/xmapper_b200.h.
#pragma once
#include "xm_types.h"
#include <vector>

namespace xm {

struct ResultsHost {
  std::vector<int64_t> q_comp_off, comp_choice_off, choice_sa_off, sa_block_off;
  std::vector<double> choice_f64, sa_f64;
  std::vector<int32_t> choice_inner, sa_contig, blocks, q_status;
  std::vector<uint8_t> sa_reversed;
  std::vector<int64_t> stats;
  std::vector<int64_t> q_cycles;  // debug probe (XM_QCYCLES=1)
  // slab mode (the CUDA library): the CSR arrays were assembled on the device and copied in ONE transfer into a
  // pinned host slab; the eleven arrays are views into it.  The vectors above stay empty.
  const char* slab = nullptr;
  int64_t slab_off[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, slab_n[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

  // q: n_queries records; choices/sas/blocks: host copies of the arena
  void assemble(int n_queries, const OutQuery* q, const OutChoice* choices, const OutSA* sas, const int32_t* blk) {
    q_comp_off.assign(1, 0); comp_choice_off.assign(1, 0); choice_sa_off.assign(1, 0); sa_block_off.assign(1, 0);
    q_status.resize((size_t)n_queries);
    for (int i = 0; i < n_queries; i++) {
      const OutQuery& oq = q[i];
      q_status[(size_t)i] = oq.status;
      int ncomp = oq.status == 0 ? oq.n_comp : 1;
      for (int c = 0; c < ncomp; c++) {
        int nch = oq.status == 0 ? oq.n_choice[c] : 0;
        for (int k = 0; k < nch; k++) {
          const OutChoice& ch = choices[oq.choice_first[c] + k];
          choice_f64.push_back(ch.spacing); choice_f64.push_back(ch.multiplier); choice_f64.push_back(ch.bonus); choice_f64.push_back(ch.total);
          choice_inner.push_back(ch.inner);
          for (int s = 0; s < ch.n_sa; s++) {
            const OutSA& sa = sas[ch.sa_first + s];
            sa_contig.push_back(sa.contig); sa_reversed.push_back((uint8_t)sa.reversed);
            sa_f64.push_back(sa.penalty); sa_f64.push_back(sa.aligned);
            const int32_t* b = blk + 4 * sa.block_first;
            blocks.insert(blocks.end(), b, b + 4 * (size_t)sa.n_blocks);
            sa_block_off.push_back((int64_t)blocks.size() / 4);
          }
          choice_sa_off.push_back((int64_t)sa_contig.size());
        }
        comp_choice_off.push_back((int64_t)choice_inner.size());
      }
      q_comp_off.push_back((int64_t)comp_choice_off.size() - 1);
    }
  }
  int64_t array(int which, const void** ptr) const {
    if (slab != nullptr && which >= 0 && which < 11) { *ptr = slab + slab_off[which]; return slab_n[which]; }
    switch (which) {
      case 0: *ptr = q_comp_off.data(); return (int64_t)q_comp_off.size();
      case 1: *ptr = comp_choice_off.data(); return (int64_t)comp_choice_off.size();
      case 2: *ptr = choice_sa_off.data(); return (int64_t)choice_sa_off.size();
      case 3: *ptr = sa_block_off.data(); return (int64_t)sa_block_off.size();
      case 4: *ptr = choice_f64.data(); return (int64_t)choice_f64.size();
      case 5: *ptr = sa_f64.data(); return (int64_t)sa_f64.size();
      case 6: *ptr = choice_inner.data(); return (int64_t)choice_inner.size();
      case 7: *ptr = sa_contig.data(); return (int64_t)sa_contig.size();
      case 8: *ptr = blocks.data(); return (int64_t)blocks.size();
      case 9: *ptr = q_status.data(); return (int64_t)q_status.size();
      case 10: *ptr = sa_reversed.data(); return (int64_t)sa_reversed.size();
      case 11: *ptr = stats.data(); return (int64_t)stats.size();
      case 12: *ptr = q_cycles.data(); return (int64_t)q_cycles.size();
    }
    return -1;
  }
};

}  // namespace xm
