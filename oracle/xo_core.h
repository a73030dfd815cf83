// ORACLE — TEST INFRASTRUCTURE ONLY.
// CPU restatement of mathjeff/Mapper @ ae7f346a (X-Mapper) query-alignment hot path.
// Nothing under oracle/ is linked into, imported by or executed from the product
// (mapper_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use it, as the checker / CPU baseline.
//
// Parity pin: the reference is Java and cannot be compiled in this container (no JVM);
// this restatement is pinned by the reference's own JUnit known-answer tests, transcribed
// to tests/golden/ (see tests/golden/make_vectors.py).  Seeding-level hash VALUES are not
// pinned by any reference fixture — "behaviour-pinned, not value-pinned".
//
// Citations: M/ = src/main/java/mapper/, QV/ = deps/QuickVariants/QuickVariants/src/main/java/mapper/
#pragma once
#include <cstdint>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <memory>
#include <algorithm>
#include <limits>
#include <stdexcept>

namespace xo {

// ---------- Java numeric semantics ----------
// (int)double in Java saturates and maps NaN to 0 (JLS 5.1.3); C++ is UB out of range.
inline int j2i(double v) {
  if (v != v) return 0;
  if (v >= 2147483647.0) return 2147483647;
  if (v <= -2147483648.0) return (-2147483647 - 1);
  return (int)v;
}
inline long long j2l(double v) {
  if (v != v) return 0;
  if (v >= 9223372036854775807.0) return std::numeric_limits<long long>::max();
  if (v <= -9223372036854775808.0) return std::numeric_limits<long long>::min();
  return (long long)v;
}
inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
inline int32_t wmul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
inline double nextUp(double v) { return std::nextafter(v, std::numeric_limits<double>::infinity()); }
// Math.abs(int) keeps Integer.MIN_VALUE negative
inline int32_t jabs(int32_t v) { return v < 0 ? (int32_t)(0u - (uint32_t)v) : v; }
static const int JMAX = 2147483647;

// ---------- QV/Basepairs.java:32-130 ----------
inline uint8_t bp_encode(char c) {
  switch (c) {
    case '-': return 0; case 'A': return 1; case 'C': return 2; case 'M': return 3;
    case 'G': return 4; case 'R': return 5; case 'S': return 6; case 'V': return 7;
    case 'T': return 8; case 'W': return 9; case 'Y': return 10; case 'H': return 11;
    case 'K': return 12; case 'D': return 13; case 'B': return 14; case 'N': return 15;
  }
  throw std::invalid_argument(std::string("Cannot encode ") + c + " as a basepair");
}
inline char bp_decode(uint8_t e) { return "-ACMGRSVTWYHKDBN"[e & 15]; }
inline bool bp_canMatch(uint8_t a, uint8_t b) { return (a & b) != 0; }
inline int bp_numChoices(uint8_t e) { return (e & 8) / 8 + (e & 4) / 4 + (e & 2) / 2 + (e & 1); }
inline double bp_falseNegRate(uint8_t e) { return (bp_numChoices(e) - 1.0) / 3.0; }
inline uint8_t bp_complement(uint8_t e) {
  uint8_t r = 0;
  if (e & 8) r += 1;
  if (e & 4) r += 2;
  if (e & 2) r += 4;
  if (e & 1) r += 8;
  return r;
}
inline bool bp_isAmbiguous(uint8_t e) { return e != 0 && e != 1 && e != 2 && e != 4 && e != 8; }
inline bool bp_isFullyAmbiguous(uint8_t e) { return bp_numChoices(e) > 3; }

// ---------- QV/Sequence.java, ReverseComplementSequence.java ----------
// Identity semantics of the Java objects are kept by using Seq* everywhere.
struct Seq {
  std::string name;
  std::vector<uint8_t> codes;  // decompressed 4-bit codes (QV/Sequence.java:76-93)
  Seq* complementedFrom = nullptr;  // non-null for "-rev" sequences
  long long id = 0;
  int length() const { return (int)codes.size(); }
  uint8_t at(int i) const { return codes[i]; }
  std::string text() const { return range(0, length()); }
  std::string range(int start, int count) const {
    std::string s;
    for (int i = start; i < start + count; i++) s.push_back(bp_decode(codes.at(i)));
    return s;
  }
  bool reversed() const { return complementedFrom != nullptr; }
};

inline std::unique_ptr<Seq> makeSeq(const std::string& name, const std::string& text) {
  // QV/SequenceBuilder.java:57 upper-cases; Basepairs.encode throws on anything else
  auto s = std::make_unique<Seq>();
  s->name = name;
  s->codes.reserve(text.size());
  for (char c : text) {
    if (c >= 'a' && c <= 'z') c = (char)(c - 'a' + 'A');
    s->codes.push_back(bp_encode(c));
  }
  return s;
}
inline std::unique_ptr<Seq> makeRC(Seq* fwd) {
  auto s = std::make_unique<Seq>();
  s->name = fwd->name + "-rev";
  int n = fwd->length();
  s->codes.resize(n);
  for (int i = 0; i < n; i++) s->codes[i] = bp_complement(fwd->codes[n - 1 - i]);
  s->complementedFrom = fwd;
  s->id = fwd->id;
  return s;
}

// ---------- M/AlignmentParameters.java ----------
struct Params {
  double MutationPenalty = 0, InsertionStart_Penalty = 0, InsertionExtension_Penalty = 0;
  double DeletionStart_Penalty = 0, DeletionExtension_Penalty = 0;
  double MaxErrorRate = 0, UnalignedPenalty = 0, AmbiguityPenalty = 0;
  int MaxNumMatches = JMAX;
  double Max_PenaltySpan = 0;
  bool StartingInsertionStartFree = false;
  double startingInsertionStartPenalty() const { return StartingInsertionStartFree ? 0 : InsertionStart_Penalty; }
  double minPossibleNonzeroPenalty() const {  // :39-44
    double r = MutationPenalty;
    r = std::min(r, startingInsertionStartPenalty() + InsertionStart_Penalty);
    r = std::min(r, DeletionStart_Penalty + DeletionExtension_Penalty);
    return r;
  }
  double basePenalty(uint8_t q, uint8_t r) const {  // :156-180
    if (!bp_canMatch(r, q)) return MutationPenalty;
    return AmbiguityPenalty * bp_falseNegRate((uint8_t)(q | r));
  }
};

// ---------- QV/AlignedBlock.java ----------
struct ABlock {
  const Seq* a; const Seq* b;
  int aStart, bStart, aLen, bLen;
  int aEnd() const { return aStart + aLen; }
  int bEnd() const { return bStart + bLen; }
  int offset() const { return bStart - aStart; }
  bool sameIndelType(const ABlock& o) const { return ((aLen > 0) == (o.aLen > 0)) && ((bLen > 0) == (o.bLen > 0)); }
  bool equalsBlock(const ABlock& o) const {
    return bStart == o.bStart && aStart == o.aStart && bLen == o.bLen && aLen == o.aLen && b == o.b && a == o.a;
  }
  bool hasAmbiguous() const {
    for (int i = aStart; i < aStart + aLen; i++) if (bp_isAmbiguous(a->at(i))) return true;
    for (int i = bStart; i < bStart + bLen; i++) if (bp_isAmbiguous(b->at(i))) return true;
    return false;
  }
};

inline double blockPenalty(const Params& p, const ABlock& blk) {  // M/AlignmentParameters.java:106-126
  double pen = 0;
  if (blk.aLen == blk.bLen) {
    for (int i = 0; i < blk.aLen; i++) pen += p.basePenalty(blk.a->at(blk.aStart + i), blk.b->at(blk.bStart + i));
  } else if (blk.aLen > 0) {
    pen += p.InsertionStart_Penalty;
    pen += p.InsertionExtension_Penalty * blk.aLen;
  } else {
    pen += p.DeletionStart_Penalty;
    pen += p.DeletionExtension_Penalty * blk.bLen;
  }
  return pen;
}
inline double blockPenaltyRange(const Params& p, const ABlock& blk, int startB, int endB) {  // :128-154
  double pen = 0;
  if (blk.aLen == blk.bLen) {
    for (int i = 0; i < blk.aLen; i++) {
      int bi = blk.bStart + i;
      if (bi >= startB && bi < endB) pen += p.basePenalty(blk.a->at(blk.aStart + i), blk.b->at(bi));
    }
  } else if (blk.bStart < endB && blk.bEnd() > startB) {
    if (blk.aLen > 0) { pen += p.InsertionStart_Penalty; pen += p.InsertionExtension_Penalty * blk.aLen; }
    else { pen += p.DeletionStart_Penalty; pen += p.DeletionExtension_Penalty * blk.bLen; }
  }
  return pen;
}

// ---------- QV/SequenceAlignment.java ----------
struct SeqAln {
  std::vector<ABlock> sections;
  bool referenceReversed = false;
  double penalty = 0, alignedPenalty = 0;
  const Seq* seqA() const { return sections[0].a; }
  const Seq* seqB() const { return sections[0].b; }
  int startA() const { return sections.front().aStart; }
  int endA() const { return sections.back().aEnd(); }
  int startB() const { return sections.front().bStart; }
  int endB() const { return sections.back().bEnd(); }
  int lengthA() const { return endA() - startA(); }
  int startOffset() const { return sections[0].offset(); }
  bool hasIndel() const { return sections.size() > 1; }
  bool hasAmbiguous() const { for (auto& s : sections) if (s.hasAmbiguous()) return true; return false; }
  int lengthABefore(int indexB) const {  // :98-117
    int total = 0;
    for (auto& b : sections) {
      if (indexB <= b.bStart) break;
      if (b.aLen < 1) continue;
      if (b.aLen > b.bLen) total += b.aLen;
      else if (indexB < b.bEnd()) total += indexB - b.bStart;
      else total += b.aLen;
    }
    return total;
  }
  int lengthAAfter(int indexB) const {  // :119-139
    int total = 0;
    for (auto& b : sections) {
      if (indexB >= b.bEnd()) continue;
      if (b.aLen < 1) continue;
      if (b.aLen > b.bLen) total += b.aLen;
      else if (indexB > b.bStart) total += b.bEnd() - indexB;
      else total += b.aLen;
    }
    return total;
  }
  int insertAOrBLength() const {
    int t = 0;
    for (auto& b : sections) if (b.aLen != b.bLen) t += b.aLen + b.bLen;
    return t;
  }
  int countNumIndels() const { int c = 0; for (auto& b : sections) if (b.aLen != b.bLen) c++; return c; }
  std::string alignedTextA() const {
    std::string r;
    for (auto& b : sections) { if (b.aLen > 0) r += b.a->range(b.aStart, b.aLen); else r += std::string(b.bLen, '-'); }
    return r;
  }
  std::string alignedTextB() const {
    std::string r;
    for (auto& b : sections) { if (b.bLen > 0) r += b.b->range(b.bStart, b.bLen); else r += std::string(b.aLen, '-'); }
    return r;
  }
  bool equalsAln(const SeqAln& o) const {  // :363-378
    if (sections.size() != o.sections.size()) return false;
    if (referenceReversed != o.referenceReversed) return false;
    for (size_t i = 0; i < sections.size(); i++) if (!sections[i].equalsBlock(o.sections[i])) return false;
    return true;
  }
  int hashCode() const { return sections[0].offset(); }
};
typedef std::shared_ptr<SeqAln> SeqAlnP;

// M/AlignmentParameters.java:73-95
inline SeqAlnP newSeqAln(const Params& p, const std::vector<ABlock>& sections, bool referenceReversed) {
  int alignedQueryLength = 0;
  double total = 0;
  for (auto& b : sections) { total += blockPenalty(p, b); alignedQueryLength += b.aLen; }
  if (!sections.empty()) {
    if (p.StartingInsertionStartFree && sections[0].bLen == 0) total -= p.InsertionStart_Penalty;
  }
  double aligned = total;
  if (!sections.empty()) {
    int unalignedLen = sections[0].a->length() - alignedQueryLength;
    double unalignedPenalty = (double)unalignedLen * p.UnalignedPenalty;
    total += unalignedPenalty;
  }
  auto r = std::make_shared<SeqAln>();
  r->sections = sections;
  r->referenceReversed = referenceReversed;
  r->penalty = total;
  r->alignedPenalty = aligned;
  return r;
}
inline double alnPenaltyRange(const Params& p, const SeqAln& a, int startB, int endB) {  // :97-103
  double t = 0;
  for (auto& b : a.sections) t += blockPenaltyRange(p, b, startB, endB);
  return t;
}

// ---------- QV/QueryAlignment.java ----------
struct QueryAln {
  std::vector<SeqAlnP> comps;
  double spacingPenalty = 0, overlapMultiplier = 0, duplicationBonus = 0, totalPenalty = 0;
  int innerDistance = 0;
  bool hasIndel() const { for (auto& c : comps) if (c->hasIndel()) return true; return false; }
  bool hasAmbiguous() const { for (auto& c : comps) if (c->hasAmbiguous()) return true; return false; }
  int hashCode() const {  // :226-233
    int32_t h = 0;
    for (auto& c : comps) h = wadd(wmul(h, 1001), c->hashCode());
    return h;
  }
  bool equalsQA(const QueryAln& o) const {  // :235-256
    if (spacingPenalty != o.spacingPenalty) return false;
    if (overlapMultiplier != o.overlapMultiplier) return false;
    if (duplicationBonus != o.duplicationBonus) return false;
    if (totalPenalty != o.totalPenalty) return false;
    if (innerDistance != o.innerDistance) return false;
    if (comps.size() != o.comps.size()) return false;
    for (size_t i = 0; i < comps.size(); i++) if (!comps[i]->equalsAln(*o.comps[i])) return false;
    return true;
  }
};
typedef std::shared_ptr<QueryAln> QueryAlnP;

// QV/Query.java
struct Query {
  std::vector<Seq*> seqs;  // as read (mate 2 NOT yet reverse-complemented)
  double expectedInnerDistance = 0;
  double spacingDeviationPerUnitPenalty = 1;
  int length() const { int t = 0; for (auto s : seqs) t += s->length(); return t; }
};

// QV/QueryAlignments.java: one list of choices per "sub-query"
struct QueryAlns {
  std::vector<std::vector<QueryAlnP>> comps;
};

struct SeqSection {  // M/SequenceSection.java
  const Seq* seq; int start, end;
  int length() const { return end - start; }
};

}  // namespace xo
