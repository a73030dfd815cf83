// ORACLE — TEST INFRASTRUCTURE ONLY (see xo_core.h header).
// extern "C" surface for ctypes (tests/, bench.py cpu_baseline / --impl reference, smoke()).
#include "xo_worker.h"
#include "xo_ancestry.h"
#include <sstream>
#include <iomanip>
#include <cstdlib>
#include <cstring>

using namespace xo;

namespace {

struct Ctx {
  SeqDb db;
  std::vector<std::pair<std::string, std::string>> pending;
  std::unique_ptr<Index> index;
  std::unique_ptr<DupDetector> dup;
  std::vector<std::vector<uint8_t>> ancestors;  // xo_infer_ancestors: forward "-anc" sequences, database order
  std::string lastError;
};

struct CParams {  // mirrors struct xm_params in include/xmapper_b200.h (kept separate on purpose)
  double mutation, ins_start, ins_ext, del_start, del_ext, max_error_rate, unaligned, ambiguity, max_penalty_span;
  int32_t max_num_matches;
  int32_t reserved;
};
Params toParams(const CParams* c) {
  Params p;
  p.MutationPenalty = c->mutation; p.InsertionStart_Penalty = c->ins_start; p.InsertionExtension_Penalty = c->ins_ext;
  p.DeletionStart_Penalty = c->del_start; p.DeletionExtension_Penalty = c->del_ext; p.MaxErrorRate = c->max_error_rate;
  p.UnalignedPenalty = c->unaligned; p.AmbiguityPenalty = c->ambiguity; p.Max_PenaltySpan = c->max_penalty_span;
  p.MaxNumMatches = c->max_num_matches;
  return p;
}

std::string fmtd(double v) { std::ostringstream o; o << std::setprecision(17) << v; return o.str(); }

void jsonSeqAln(std::ostringstream& o, const SeqAln& a, const SeqDb& db) {
  const Seq* b = a.seqB();
  o << "{\"contig\":\"" << b->name << "\",\"contig_index\":" << (b->id / 2) << ",\"reversed\":" << (a.referenceReversed ? "true" : "false")
    << ",\"penalty\":" << fmtd(a.penalty) << ",\"aligned_penalty\":" << fmtd(a.alignedPenalty) << ",\"start_b\":" << a.startB()
    << ",\"aligned_a\":\"" << a.alignedTextA() << "\",\"aligned_b\":\"" << a.alignedTextB() << "\",\"query_text\":\"" << a.seqA()->text()
    << "\",\"blocks\":[";
  for (size_t i = 0; i < a.sections.size(); i++) {
    auto& s = a.sections[i];
    if (i) o << ",";
    o << "[" << s.aStart << "," << s.bStart << "," << s.aLen << "," << s.bLen << "]";
  }
  o << "]}";
  (void)db;
}
std::string jsonAlns(const QueryAlns& r, const SeqDb& db) {
  std::ostringstream o;
  o << "{\"components\":[";
  for (size_t c = 0; c < r.comps.size(); c++) {
    if (c) o << ",";
    o << "[";
    for (size_t k = 0; k < r.comps[c].size(); k++) {
      const QueryAln& qa = *r.comps[c][k];
      if (k) o << ",";
      o << "{\"penalty\":" << fmtd(qa.totalPenalty) << ",\"spacing_penalty\":" << fmtd(qa.spacingPenalty) << ",\"overlap_multiplier\":" << fmtd(qa.overlapMultiplier)
        << ",\"duplication_bonus\":" << fmtd(qa.duplicationBonus) << ",\"inner_distance\":" << qa.innerDistance << ",\"seqs\":[";
      for (size_t s = 0; s < qa.comps.size(); s++) { if (s) o << ","; jsonSeqAln(o, *qa.comps[s], db); }
      o << "]}";
    }
    o << "]";
  }
  o << "]}";
  return o.str();
}
char* dupstr(const std::string& s) { char* r = (char*)malloc(s.size() + 1); memcpy(r, s.c_str(), s.size() + 1); return r; }

struct Results {
  std::vector<int64_t> q_comp_off, comp_choice_off, choice_sa_off, sa_block_off;
  std::vector<double> choice_f64, sa_f64;
  std::vector<int32_t> choice_inner, sa_contig, blocks, q_status;
  std::vector<uint8_t> sa_reversed;
  std::vector<int64_t> stats;  // probes, seeds, hits, pathCalls, pathSteps, pathCells, straightCalls, immediateAccepts
};

void appendResults(Results& R, const QueryAlns& qa) {
  for (auto& comp : qa.comps) {
    for (auto& ch : comp) {
      R.choice_f64.push_back(ch->spacingPenalty); R.choice_f64.push_back(ch->overlapMultiplier);
      R.choice_f64.push_back(ch->duplicationBonus); R.choice_f64.push_back(ch->totalPenalty);
      R.choice_inner.push_back(ch->innerDistance);
      for (auto& sa : ch->comps) {
        R.sa_contig.push_back((int32_t)(sa->seqB()->id / 2));
        R.sa_reversed.push_back(sa->referenceReversed ? 1 : 0);
        R.sa_f64.push_back(sa->penalty); R.sa_f64.push_back(sa->alignedPenalty);
        for (auto& b : sa->sections) { R.blocks.push_back(b.aStart); R.blocks.push_back(b.bStart); R.blocks.push_back(b.aLen); R.blocks.push_back(b.bLen); }
        R.sa_block_off.push_back((int64_t)R.blocks.size() / 4);
      }
      R.choice_sa_off.push_back((int64_t)R.sa_contig.size());
    }
    R.comp_choice_off.push_back((int64_t)R.choice_inner.size());
  }
  R.q_comp_off.push_back((int64_t)R.comp_choice_off.size());
}

}  // namespace

extern "C" {

void* xo_create() { return new Ctx(); }
void xo_destroy(void* c) { delete (Ctx*)c; }
const char* xo_last_error(void* c) { return ((Ctx*)c)->lastError.c_str(); }
void xo_free(void* p) { free(p); }

int xo_add_contig(void* c, const char* name, const char* text) {
  try { ((Ctx*)c)->pending.push_back({name, text}); return 0; } catch (std::exception& e) { ((Ctx*)c)->lastError = e.what(); return -1; }
}
// sort_by_length = 1: Mapper.sortAndComplementReference (M/Mapper.java:1151-1172): descending length, ties in input order.
// sort_by_length = 0: input order (Api.newDatabase / unit tests).
int xo_finalize_reference(void* cv, int sort_by_length) {
  Ctx* c = (Ctx*)cv;
  try {
    std::vector<size_t> order(c->pending.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    if (sort_by_length) std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return c->pending[a].second.size() > c->pending[b].second.size(); });
    for (size_t i : order) c->db.addWithRC(makeSeq(c->pending[i].first, c->pending[i].second));
    c->db.computeMetrics();
    c->pending.clear();
    return 0;
  } catch (std::exception& e) { c->lastError = e.what(); return -1; }
}
int xo_num_contigs(void* cv) { return (int)((Ctx*)cv)->db.seqs.size() / 2; }
int xo_contig_length(void* cv, int i) { return ((Ctx*)cv)->db.seqs[(size_t)i * 2]->length(); }
const char* xo_contig_name(void* cv, int i) { return ((Ctx*)cv)->db.seqs[(size_t)i * 2]->name.c_str(); }
// copies unpacked 4-bit codes of forward contig i
void xo_contig_codes(void* cv, int i, uint8_t* out) { auto s = ((Ctx*)cv)->db.seqs[(size_t)i * 2]; memcpy(out, s->codes.data(), s->codes.size()); }

int xo_create_index(void* cv, int min_interesting, int max_short_matches, int enable_gapmers, int threads) {
  Ctx* c = (Ctx*)cv;
  try {
    c->index = std::make_unique<Index>(&c->db, min_interesting, max_short_matches, enable_gapmers != 0);
    c->index->numThreads = threads;
    return 0;
  } catch (std::exception& e) { c->lastError = e.what(); return -1; }
}
int xo_build_index_through(void* cv, int n) {
  Ctx* c = (Ctx*)cv;
  try {
    if (c->index->maxBuilt < 1) c->index->tableFor(1);
    if (n > c->index->maxBuilt) c->index->buildThrough(n);
    return c->index->maxBuilt;
  } catch (std::exception& e) { c->lastError = e.what(); return -1; }
}
int xo_index_min_interesting(void* cv) { return ((Ctx*)cv)->index->minInterestingSize; }
int xo_index_max_built(void* cv) { return ((Ctx*)cv)->index->maxBuilt; }
// table export: returns capacity; counts via out params
int xo_index_table_info(void* cv, int n, int* capacity, int* max_count, int64_t* n_positions) {
  Ctx* c = (Ctx*)cv;
  if (n < 0 || n > c->index->maxBuilt) return -1;
  const LenTable& T = c->index->tables[(size_t)n];
  *capacity = T.capacity; *max_count = T.maxCount; *n_positions = (int64_t)T.positions.size();
  return 0;
}
// offsets: capacity+1 int64, -1 marks an overfull bucket start is kept (overfull flags returned separately)
int xo_index_table_copy(void* cv, int n, int64_t* offsets, uint32_t* positions, uint8_t* overfull) {
  Ctx* c = (Ctx*)cv;
  const LenTable& T = c->index->tables[(size_t)n];
  if (T.offsets.empty()) { offsets[0] = 0; offsets[1] = 0; overfull[0] = 0; return 0; }
  memcpy(offsets, T.offsets.data(), T.offsets.size() * sizeof(int64_t));
  if (!T.positions.empty()) memcpy(positions, T.positions.data(), T.positions.size() * sizeof(uint32_t));
  memcpy(overfull, T.overfull.data(), T.overfull.size());
  return 0;
}

// min_len/max_len < 0: DuplicationDetector.chooseMin/MaxDuplicationLength
int xo_create_dup_detector(void* cv, int min_len, int max_len, int min_copies, int window) {
  Ctx* c = (Ctx*)cv;
  try {
    if (min_len < 0) min_len = Index::chooseMinDuplicationLength(&c->db);
    if (max_len < 0) max_len = Index::chooseMaxDuplicationLength(&c->db);
    c->dup = std::make_unique<DupDetector>(c->index.get(), min_len, max_len, min_copies, window);
    return 0;
  } catch (std::exception& e) { c->lastError = e.what(); return -1; }
}
int xo_dup_detect(void* cv) {
  Ctx* c = (Ctx*)cv;
  try { c->dup->detect(); return 0; } catch (std::exception& e) { c->lastError = e.what(); return -1; }
}
// --infer-ancestors (M/Mapper.java:675-681): the context's DupDetector must have been created with min_copies = 3, window = 1
int xo_infer_ancestors(void* cv, double dissimilarity_threshold, int verify_no_duplicate_analyses) {
  Ctx* c = (Ctx*)cv;
  try {
    AncestryDetector a(c->dup.get(), dissimilarity_threshold);
    a.verifyNoDuplicateAnalyses = verify_no_duplicate_analyses != 0;
    c->ancestors = a.run();
    return (int)c->ancestors.size();
  } catch (std::exception& e) { c->lastError = e.what(); return -1; }
}
void xo_ancestor_codes(void* cv, int contig, uint8_t* out) {
  Ctx* c = (Ctx*)cv;
  const auto& v = c->ancestors[(size_t)contig];
  memcpy(out, v.data(), v.size());
}
double xo_dup_granularity(void* cv) { return ((Ctx*)cv)->dup->detectionGranularity(); }
// duplication start keys on forward contig i
int64_t xo_dup_count(void* cv, int contig) {
  Ctx* c = (Ctx*)cv;
  auto it = c->dup->bySeq.find(c->db.seqs[(size_t)contig * 2]);
  return it == c->dup->bySeq.end() ? 0 : (int64_t)it->second.size();
}
void xo_dup_copy(void* cv, int contig, int32_t* out) {
  Ctx* c = (Ctx*)cv;
  auto it = c->dup->bySeq.find(c->db.seqs[(size_t)contig * 2]);
  if (it == c->dup->bySeq.end()) return;
  size_t k = 0;
  for (auto& e : it->second) out[k++] = e.first;
}

// Single query, text in, JSON out (Api.align). seq2 may be NULL.
char* xo_align_json(void* cv, const CParams* cp, const char* seq1, const char* seq2, double expected_inner, double per_penalty) {
  Ctx* c = (Ctx*)cv;
  try {
    Params p = toParams(cp);
    std::vector<std::unique_ptr<Seq>> own;
    Query q;
    std::vector<Seq*> rcs;
    own.push_back(makeSeq("query1", seq1)); q.seqs.push_back(own.back().get());
    if (seq2) { own.push_back(makeSeq("query2", seq2)); q.seqs.push_back(own.back().get()); }
    size_t n = q.seqs.size();
    for (size_t i = 0; i < n; i++) { own.push_back(makeRC(q.seqs[i])); rcs.push_back(own.back().get()); }
    q.expectedInnerDistance = expected_inner;
    q.spacingDeviationPerUnitPenalty = seq2 ? per_penalty : 1;
    if (!seq2) q.expectedInnerDistance = 0;
    c->index->prepare();
    Worker w(c->index.get(), c->dup.get(), p);
    QueryAlns r = w.align(q, rcs);
    return dupstr(jsonAlns(r, c->db));
  } catch (std::exception& e) { c->lastError = e.what(); return nullptr; }
}

// Batch: QV-packed 4-bit reads (base i of a sequence at bits 4*(i&3) of its 16-bit word i>>2).
void* xo_align_batch(void* cv, const CParams* cp, int n_queries, const uint16_t* packed, const int64_t* seq_word_off,
                     const int32_t* seq_len, const uint8_t* n_seqs_per_query, const double* expected_inner,
                     const double* per_penalty, int n_threads) {
  Ctx* c = (Ctx*)cv;
  try {
    Params p = toParams(cp);
    c->index->prepare();
    std::vector<int64_t> firstSeq((size_t)n_queries + 1, 0);
    for (int i = 0; i < n_queries; i++) firstSeq[(size_t)i + 1] = firstSeq[(size_t)i] + n_seqs_per_query[i];
    int nt = std::max(1, n_threads);
    std::vector<Results> parts((size_t)nt);
    std::vector<OracleStats> st((size_t)nt);
    std::vector<int> imm((size_t)nt, 0);
    std::vector<std::string> errs((size_t)nt);
    auto work = [&](int t) {
      int lo = (int)((long long)n_queries * t / nt), hi = (int)((long long)n_queries * (t + 1) / nt);
      Worker w(c->index.get(), c->dup.get(), p);
      Results& R = parts[(size_t)t];
      for (int qi = lo; qi < hi; qi++) {
        std::vector<std::unique_ptr<Seq>> own;
        Query q; std::vector<Seq*> rcs;
        for (int64_t s = firstSeq[(size_t)qi]; s < firstSeq[(size_t)qi + 1]; s++) {
          auto sq = std::make_unique<Seq>();
          sq->name = "q";
          int L = seq_len[s];
          sq->codes.resize((size_t)L);
          const uint16_t* w16 = packed + seq_word_off[s];
          for (int i = 0; i < L; i++) sq->codes[(size_t)i] = (uint8_t)((w16[i >> 2] >> ((i & 3) << 2)) & 15);
          q.seqs.push_back(sq.get());
          own.push_back(std::move(sq));
        }
        size_t n = q.seqs.size();
        for (size_t i = 0; i < n; i++) { own.push_back(makeRC(q.seqs[i])); rcs.push_back(own.back().get()); }
        q.expectedInnerDistance = n > 1 ? expected_inner[qi] : 0;
        q.spacingDeviationPerUnitPenalty = n > 1 ? per_penalty[qi] : 1;
        bool empty_seq = false;
        for (size_t i = 0; i < n; i++) if (q.seqs[i]->codes.empty()) empty_seq = true;
        if (empty_seq) {  // a zero-length sequence: no reference test covers it; both sides refuse it with the same status (the product's Q_INTERNAL)
          QueryAlns r; r.comps.emplace_back();
          appendResults(R, r);
          R.q_status.push_back(-6);
          continue;
        }
        try {
          QueryAlns r = w.align(q, rcs);
          appendResults(R, r);
          R.q_status.push_back(0);
        } catch (std::exception& e) {
          errs[(size_t)t] = e.what();
          QueryAlns r; r.comps.emplace_back();
          appendResults(R, r);
          R.q_status.push_back(-1);
        }
      }
      st[(size_t)t] = w.stats; imm[(size_t)t] = w.numImmediateAccepts;
    };
    if (nt == 1) work(0);
    else { std::vector<std::thread> th; for (int t = 0; t < nt; t++) th.emplace_back(work, t); for (auto& x : th) x.join(); }
    // merge parts in query order, rebasing offsets
    Results* R = new Results();
    R->q_comp_off.push_back(0); R->comp_choice_off.push_back(0); R->choice_sa_off.push_back(0); R->sa_block_off.push_back(0);
    OracleStats tot; int totImm = 0;
    for (int t = 0; t < nt; t++) {
      Results& P = parts[(size_t)t];
      int64_t compBase = (int64_t)R->comp_choice_off.size() - 1, choiceBase = (int64_t)R->choice_inner.size();
      int64_t saBase = (int64_t)R->sa_contig.size(), blockBase = (int64_t)R->blocks.size() / 4;
      // P's offset arrays were built without leading zeros
      for (auto v : P.q_comp_off) R->q_comp_off.push_back(compBase + v);
      for (auto v : P.comp_choice_off) R->comp_choice_off.push_back(choiceBase + v);
      for (auto v : P.choice_sa_off) R->choice_sa_off.push_back(saBase + v);
      for (auto v : P.sa_block_off) R->sa_block_off.push_back(blockBase + v);
      R->choice_f64.insert(R->choice_f64.end(), P.choice_f64.begin(), P.choice_f64.end());
      R->sa_f64.insert(R->sa_f64.end(), P.sa_f64.begin(), P.sa_f64.end());
      R->choice_inner.insert(R->choice_inner.end(), P.choice_inner.begin(), P.choice_inner.end());
      R->sa_contig.insert(R->sa_contig.end(), P.sa_contig.begin(), P.sa_contig.end());
      R->blocks.insert(R->blocks.end(), P.blocks.begin(), P.blocks.end());
      R->q_status.insert(R->q_status.end(), P.q_status.begin(), P.q_status.end());
      R->sa_reversed.insert(R->sa_reversed.end(), P.sa_reversed.begin(), P.sa_reversed.end());
      tot.probes += st[(size_t)t].probes; tot.seeds += st[(size_t)t].seeds; tot.hits += st[(size_t)t].hits;
      tot.pathAlignerCalls += st[(size_t)t].pathAlignerCalls; tot.pathAlignerSteps += st[(size_t)t].pathAlignerSteps;
      tot.pathAlignerCells += st[(size_t)t].pathAlignerCells; tot.straightCalls += st[(size_t)t].straightCalls;
      totImm += imm[(size_t)t];
      if (!errs[(size_t)t].empty()) c->lastError = errs[(size_t)t];
    }
    R->stats = {tot.probes, tot.seeds, tot.hits, tot.pathAlignerCalls, tot.pathAlignerSteps, tot.pathAlignerCells, tot.straightCalls, totImm};
    return R;
  } catch (std::exception& e) { c->lastError = e.what(); return nullptr; }
}
void xo_results_free(void* r) { delete (Results*)r; }
// which: 0 q_comp_off 1 comp_choice_off 2 choice_sa_off 3 sa_block_off (int64) | 4 choice_f64 5 sa_f64 (f64)
//        6 choice_inner 7 sa_contig 8 blocks 9 q_status (int32) | 10 sa_reversed (u8) | 11 stats (int64)
int64_t xo_results_array(void* rv, int which, const void** ptr) {
  Results* R = (Results*)rv;
  switch (which) {
    case 0: *ptr = R->q_comp_off.data(); return (int64_t)R->q_comp_off.size();
    case 1: *ptr = R->comp_choice_off.data(); return (int64_t)R->comp_choice_off.size();
    case 2: *ptr = R->choice_sa_off.data(); return (int64_t)R->choice_sa_off.size();
    case 3: *ptr = R->sa_block_off.data(); return (int64_t)R->sa_block_off.size();
    case 4: *ptr = R->choice_f64.data(); return (int64_t)R->choice_f64.size();
    case 5: *ptr = R->sa_f64.data(); return (int64_t)R->sa_f64.size();
    case 6: *ptr = R->choice_inner.data(); return (int64_t)R->choice_inner.size();
    case 7: *ptr = R->sa_contig.data(); return (int64_t)R->sa_contig.size();
    case 8: *ptr = R->blocks.data(); return (int64_t)R->blocks.size();
    case 9: *ptr = R->q_status.data(); return (int64_t)R->q_status.size();
    case 10: *ptr = R->sa_reversed.data(); return (int64_t)R->sa_reversed.size();
    case 11: *ptr = R->stats.data(); return (int64_t)R->stats.size();
  }
  return -1;
}

// ---- unit-level hooks for the reference's own unit tests ----

// T/PathAligner_Test.java:50-62 — PathAligner alone on (a, b), offset 0
char* xo_test_path_aligner(const CParams* cp, const char* a, const char* b, double max_ins, double max_del) {
  try {
    Params p = toParams(cp);
    auto sa = makeSeq("a", a); auto sb = makeSeq("b", b);
    PathAligner pa;
    Analysis an; an.maxIns = max_ins; an.maxDel = max_del;
    SeqAlnP r = pa.align(SeqSection{sa.get(), 0, sa->length()}, SeqSection{sb.get(), 0, sb->length()}, p, an);
    if (!r) return dupstr("null");
    std::ostringstream o; SeqDb d; jsonSeqAln(o, *r, d);
    return dupstr(o.str());
  } catch (std::exception& e) { return dupstr(std::string("error:") + e.what()); }
}
// T/HashBlockAligner_Test.java:59-72 — HashBlock_Aligner(StraightAligner(PathAligner_Runner))
char* xo_test_hashblock_aligner(const CParams* cp, const char* a, const char* b, double max_ins, double max_del) {
  try {
    Params p = toParams(cp);
    auto sa = makeSeq("a", a); auto sb = makeSeq("b", b);
    std::unique_ptr<LocalAligner> al = std::make_unique<PathAligner>();
    al = std::make_unique<StraightAligner>(std::move(al));
    al = std::make_unique<HashBlockAligner>(std::move(al));
    Analysis an; an.maxIns = max_ins; an.maxDel = max_del;
    SeqAlnP r = al->align(SeqSection{sa.get(), 0, sa->length()}, SeqSection{sb.get(), 0, sb->length()}, p, an);
    if (!r) return dupstr("null");
    std::ostringstream o; SeqDb d; jsonSeqAln(o, *r, d);
    return dupstr(o.str());
  } catch (std::exception& e) { return dupstr(std::string("error:") + e.what()); }
}
// T/Counting_HashBlockPath_Test.java — offsets of findGoodPositionsHavingPriorityUpTo(priority) for one query
char* xo_test_counting_path(void* cv, const CParams* cp, const char* query, int priority) {
  Ctx* c = (Ctx*)cv;
  try {
    Params p = toParams(cp);
    auto q = makeSeq("query", query); auto rq = makeRC(q.get());
    Pyramid pyr(q.get());
    CountingPath path(&pyr, c->index.get(), &c->db, q.get(), rq.get(), p);
    CounterList l = path.findGoodPositionsHavingPriorityUpTo(priority);
    std::ostringstream o; o << "[";
    for (size_t i = 0; i < l->size(); i++) { if (i) o << ","; o << "[" << ((*l)[i]->match.reversed() ? 1 : 0) << "," << (*l)[i]->match.offset << "]"; }
    o << "]";
    return dupstr(o.str());
  } catch (std::exception& e) { c->lastError = e.what(); return nullptr; }
}
// T/HashBlockPaths_Counter_Test.java:77-100 — findGoodPositionsHavingPriority(0) for (seq1, rc(seq2text))
char* xo_test_paths_counter(void* cv, const CParams* cp, const char* seq1, const char* seq2_as_read, int expected_inner, int max_inner) {
  Ctx* c = (Ctx*)cv;
  try {
    Params p = toParams(cp);
    auto q1 = makeSeq("seq1", seq1); auto r1 = makeRC(q1.get());
    auto q2 = makeSeq("seq2", seq2_as_read); auto r2 = makeRC(q2.get());
    // the test passes both sequences as-is (no extra reverse complement inside makePath)
    Pyramid p1(q1.get()), p2(q2.get());
    CountingPath c1(&p1, c->index.get(), &c->db, q1.get(), r1.get(), p), c2(&p2, c->index.get(), &c->db, q2.get(), r2.get(), p);
    PathsCounter pc({&c1, &c2}, expected_inner, max_inner);
    QMList m = pc.findGoodPositionsHavingPriority(0);
    std::ostringstream o; o << "[";
    for (size_t i = 0; i < m->size(); i++) {
      if (i) o << ",";
      o << "{\"inner\":" << (*m)[i].totalDistanceBetweenComponents() << ",\"across\":" << (*m)[i].totalDistanceAcross() << "}";
    }
    o << "]";
    return dupstr(o.str());
  } catch (std::exception& e) { c->lastError = e.what(); return nullptr; }
}
// T/HashBlock_Test.java — forward/reverse-complement symmetry of every single block of every row. returns #blocks checked, <0 on failure
int xo_test_hash_symmetry(const char* text) {
  auto s = makeSeq("q", text); auto r = makeRC(s.get());
  Pyramid ps(s.get());
  int checked = 0;
  auto hashSequence = [&](const Seq* seq, int startIndex, int endIndex, HB& out) -> bool {
    Pyramid pr(seq);
    for (int level = 0;; level++) {
      const MB* b = pr.get(level)->get(startIndex);
      if (b == nullptr) return false;
      if (b->single) { if (b->hb.end() == endIndex) { out = b->hb; return true; } }
      else for (auto& p : b->poss) if (p.has && p.hb.end() == endIndex) { out = p.hb; return true; }
    }
  };
  for (int level = 0;; level++) {
    Row* row = ps.get(level);
    if (row->getAfter(-1) == nullptr) break;
    int i = -1;
    while (true) {
      const MB* b = row->getAfter(i);
      if (b == nullptr) break;
      i = b->startIndex();
      if (!b->single) continue;
      const HB& blk = b->hb;
      HB rb;
      if (!hashSequence(r.get(), s->length() - blk.end(), s->length() - blk.start, rb)) return -1;
      if (rb.fwd != blk.rev || rb.rev != blk.fwd) return -2;
      if (blk.rml != rb.rmr || blk.rmr != rb.rml) return -3;
      if (blk.nrml != rb.nrmr || blk.nrmr != rb.nrml) return -4;
      if (!blk.primary() && !blk.secondary()) return -5;
      HB e, re;
      bool he = withGapAndExtension(blk, s.get(), e), hre = withGapAndExtension(rb, r.get(), re);
      if (he != hre) return -6;
      if (he) { if (re.fwd != e.rev || re.rev != e.fwd) return -7; }
      checked++;
    }
  }
  return checked;
}

// T/MultiHashBlock_Test.java:90-117,129-165 — hashString(): for every row, the block at index 0 and its possibilities that end at the
// end of the sequence.  Returns 1 if hashing `ambiguous` yields a possibility equal (start, end, forward hash) to the single block
// that spans all of `text`, 0 if not, -1 if `text` itself is not spanned by exactly one block.
static void xo_hash_string(const Seq* seq, std::vector<HB>& out) {
  Pyramid p(seq);
  for (int level = 0;; level++) {
    Row* row = p.get(level);
    if (row == nullptr) break;
    const MB* b = row->get(0);
    if (b == nullptr) break;
    if (b->single) { if (b->hb.end() == seq->length()) out.push_back(b->hb); }
    else for (auto& c : b->poss) if (c.has && c.hb.end() == seq->length()) out.push_back(c.hb);
  }
}
int xo_test_multi_expand(const char* text, const char* ambiguous) {
  try {
    auto s = makeSeq("q", text); auto a = makeSeq("q", ambiguous);
    std::vector<HB> plain, amb;
    xo_hash_string(s.get(), plain);
    if (plain.size() != 1) return -1;
    xo_hash_string(a.get(), amb);
    for (auto& h : amb) if (h.start == plain[0].start && h.end() == plain[0].end() && h.fwd == plain[0].fwd) return 1;
    return 0;
  } catch (std::exception&) { return -2; }
}

// T/BasepairsTest.java:9-47 — AlignmentParameters.getPenalty(encoded, encoded) (M/AlignmentParameters.java:156-180)
double xo_test_base_penalty(const CParams* cp, const char* q, const char* r) {
  Params p = toParams(cp);
  return p.basePenalty(bp_encode(q[0]), bp_encode(r[0]));
}
// T/SequenceDatabase_Test.java:17-41 / T/PackedMap_Test.java:14-48 — encodePosition / decodePosition over sequences whose total
// length exceeds 2^31 (lengths only: the bases are never touched).  Returns the number of (sequence, offset) pairs that round-trip.
int64_t xo_test_position_roundtrip(int n_sequences, int64_t length) {
  std::vector<long long> starts((size_t)n_sequences);
  long long t = 0;
  for (int i = 0; i < n_sequences; i++) { starts[(size_t)i] = t; t += length; }   // SequenceDatabase.computeMetrics :240-248
  int64_t ok = 0;
  const long long offs[4] = {0, 100, length - 100, length - 1};
  for (int i = 0; i < n_sequences; i++) for (long long off : offs) {
    long long enc = starts[(size_t)i] + off;                                       // encodePosition :88-105
    size_t idx = (size_t)(std::upper_bound(starts.begin(), starts.end(), enc) - starts.begin()) - 1;   // decodePosition :170-209
    if ((int)idx == i && enc - starts[idx] == off) ok++;
  }
  return ok;
}

}  // extern "C"
