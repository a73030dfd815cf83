// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's ancestor inference (--infer-ancestors):
//   M/AncestryDetector.java:89-147 (unionRecentAncestors), :153-337 (analyze), :339-421 (bounds), :423-439 (scores),
//   QV/SimilarityAnalysis.java, M/OverriddenSequence.java:19-27 (a position can be overridden once).
// Input: a DupDetector built with minNumInterestingCopies = 3 and windowSize = 1 over the original reference
// (M/Mapper.java:675-681); output: the forward "-anc" sequences (IUPAC unions written where a recent common ancestor was inferred).
// Set iteration orders (HashSet<SimilarityAnalysis>, Set<Duplication>) do not influence the result: every per-step quantity is a
// count or a per-analysis update, and a position is written at most once (OverriddenSequence throws otherwise).
// Pinned by the seven exact strings of T/AncestryDetector_Test.java:10-91 (tests/test_oracle_junit.py).
#pragma once
#include "xo_index.h"
#include <map>
#include <set>
#include <stdexcept>

namespace xo {

struct SimilarityAnalysis {  // QV/SimilarityAnalysis.java
  Seq* sequence; int startIndex, boundIndex, currentIndex, bestIndex; double cumulativeScore, bestScore;
  SimilarityAnalysis(Seq* s, int start, int bound, double initial)
      : sequence(s), startIndex(start), boundIndex(bound), currentIndex(start), bestIndex(start), cumulativeScore(initial), bestScore(initial) {}
  void addScore(double v) { cumulativeScore += v; if (cumulativeScore > bestScore) { bestScore = cumulativeScore; bestIndex = currentIndex; } }
  bool reachedEndOfSequence() const { return currentIndex < 0 || currentIndex >= sequence->length(); }
};

struct AncestryDetector {
  DupDetector* dup; double dissimilarityThreshold; bool verifyNoDuplicateAnalyses = false;
  std::map<const Seq*, std::map<int, uint8_t>> overrides;

  AncestryDetector(DupDetector* d, double threshold) : dup(d), dissimilarityThreshold(threshold) {}
  double matchScore(int length) const { return dissimilarityThreshold * length; }          // :423-425
  double mismatchScore(int length) const { return -length + matchScore(length); }          // :427-431
  static int centerOf(int start, int length) { return start + length / 2; }
  static int middleBetween(int l, int r) { return (l + r) / 2; }

  void write(const Seq* seq, int index, uint8_t allele) {                                   // :339-351 + OverriddenSequence.putEncoded
    auto& m = overrides[seq];
    if (m.count(index)) throw std::runtime_error("Cannot override " + seq->name + "[" + std::to_string(index) + "]: already overridden");
    m[index] = allele;
  }
  typedef std::map<int, DupDetector::Dup> DupMap;
  static DupMap::const_iterator interestingBefore(int index, const DupMap& m) {             // :353-366
    while (true) {
      auto it = m.lower_bound(index);
      if (it == m.begin()) return m.end();
      --it;
      if (it->second.numInstances >= 3) return it;
      index = it->first;
    }
  }
  static DupMap::const_iterator interestingAfter(int index, const DupMap& m) {              // :368-382
    while (true) {
      auto it = m.upper_bound(index);
      if (it == m.end()) return m.end();
      if (it->second.numInstances >= 3) return it;
      index = it->first;
    }
  }
  // :384-421; false = null
  bool analysisBounds(const DupDetector::DupGroup& d, const SeqPos& sp, int polarity, SimilarityAnalysis& out) {
    Seq* sequence = sp.seq;
    int startIndex = sp.start;
    const DupMap& here = dup->bySeq[sequence];
    int middle = centerOf(startIndex, d.length);
    int initial = polarity > 0 ? middle + 1 : middle;
    int bound;
    if (polarity > 0) {
      bound = sequence->length();
      auto nx = interestingAfter(startIndex, here);
      if (nx != here.end()) bound = middleBetween(middle, centerOf(nx->first, nx->second.length)) + 1;
    } else {
      bound = -1;
      auto pv = interestingBefore(startIndex, here);
      if (pv != here.end()) bound = middleBetween(centerOf(pv->first, pv->second.length), middle);
    }
    out = SimilarityAnalysis(sequence, initial, bound, matchScore(d.length));
    if ((out.boundIndex - out.startIndex) * polarity < 0) return false;
    return true;
  }
  void analyze(const std::shared_ptr<DupDetector::DupGroup>& g, int polarity) {             // :157-337
    if ((int)g->starts.size() < 3) return;
    std::vector<std::unique_ptr<SimilarityAnalysis>> pool;
    std::set<SimilarityAnalysis*> available, interested;
    for (const SeqPos& sp : g->starts) {
      SimilarityAnalysis a(nullptr, 0, 0, 0);
      if (!analysisBounds(*g, sp, polarity, a)) continue;
      pool.push_back(std::make_unique<SimilarityAnalysis>(a));
      available.insert(pool.back().get());
      const DupMap& here = dup->bySeq[sp.seq];
      auto it = here.find(sp.start);
      if (it != here.end() && it->second.group.get() == g.get()) interested.insert(pool.back().get());
    }
    std::vector<uint8_t> popular;
    const uint8_t noAncestor = 0;  // Basepairs.encode('-')
    while (interested.size() >= 1 && available.size() >= 3) {
      std::set<SimilarityAnalysis*> noLongerInterested, noLongerAvailable;
      for (auto* s : interested) if (s->currentIndex == s->boundIndex) noLongerInterested.insert(s);
      int counts[16] = {0};
      for (auto* s : available) {
        int cp = s->currentIndex;
        if (cp < 0 || cp >= s->sequence->length()) { noLongerAvailable.insert(s); if (interested.count(s)) noLongerInterested.insert(s); }
        else counts[s->sequence->at(cp)]++;
      }
      int bestCount = 0; uint8_t best = 0; bool tie = false;
      for (int item = 0; item < 16; item++) {     // HashMap<Byte,Integer> iterates in ascending byte value here (hash = value, 16 buckets)
        int c = counts[item];
        if (c == 0) continue;
        if (c > bestCount) { bestCount = c; best = (uint8_t)item; tie = false; }
        else if (c == bestCount) tie = true;
      }
      if (tie) best = noAncestor;
      popular.push_back(best);
      for (auto* s : noLongerInterested) {
        bool hasNeighbor = !s->reachedEndOfSequence();
        bool hasScore = s->cumulativeScore >= 0;
        if (hasNeighbor && hasScore) s->addScore(mismatchScore(3) * -1);
        interested.erase(s);
      }
      for (auto* s : noLongerAvailable) available.erase(s);
      for (auto* s : available) {
        uint8_t here = s->sequence->at(s->currentIndex);
        s->addScore(here == best ? matchScore(1) : mismatchScore(1));
        if (s->cumulativeScore < 0) { noLongerAvailable.insert(s); if (interested.count(s)) noLongerInterested.insert(s); }
      }
      for (auto* s : noLongerAvailable) available.erase(s);
      for (auto* s : noLongerInterested) interested.erase(s);
      for (auto* s : available) s->currentIndex += polarity;
      for (auto* s : noLongerInterested) {
        for (int offset = 0; offset < (int)popular.size(); offset++) {
          int index = s->startIndex + offset * polarity;
          if (index == s->boundIndex) break;
          uint8_t anc = popular[(size_t)offset];
          uint8_t item = s->sequence->at(index);
          if ((anc != item && anc != noAncestor) || verifyNoDuplicateAnalyses) write(s->sequence, index, (uint8_t)(anc | item));
          if (index == s->bestIndex) break;
        }
      }
    }
  }
  // unionRecentAncestors :89-147: returns the forward sequences with their overrides applied, in database order
  std::vector<std::vector<uint8_t>> run() {
    dup->detect();
    std::set<DupDetector::DupGroup*> seen;
    std::vector<std::shared_ptr<DupDetector::DupGroup>> all;       // DuplicationDetector.getAll :79-90
    for (auto& e : dup->bySeq) for (auto& p : e.second) if (seen.insert(p.second.group.get()).second) all.push_back(p.second.group);
    for (auto& g : all) { analyze(g, -1); analyze(g, 1); }
    std::vector<std::vector<uint8_t>> out;
    for (Seq* s : dup->index->db->seqs) {
      if (s->complementedFrom) continue;                           // reverse sequences are regenerated from the forward overrides
      std::vector<uint8_t> codes = s->codes;
      auto it = overrides.find(s);
      if (it != overrides.end()) for (auto& o : it->second) codes[(size_t)o.first] = o.second;
      out.push_back(std::move(codes));
    }
    return out;
  }
};

}  // namespace xo
