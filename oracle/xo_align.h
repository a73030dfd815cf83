// ORACLE — TEST INFRASTRUCTURE ONLY (see xo_core.h header).
// The scoring + traceback cascade.
// Follows M/StraightAligner.java, M/SkipHighAmbiguity_Aligner.java, M/HashBlock_Aligner.java,
// M/HashBlock_Matcher.java, M/CountMap.java, M/BlockAligner.java, M/PathAligner.java,
// M/AlignmentNode.java, M/AlignmentAnalysis.java, M/PenaltyAnalysis.java.
#pragma once
#include "xo_path.h"
#include <unordered_map>

namespace xo {

struct Matcher {  // M/HashBlock_Matcher.java
  static const int NO_MATCHES = -1, MULTIPLE_MATCHES = -2, UNKNOWN = -3;
  const Seq* query; const Seq* reference;
  int referenceStart, referenceLength, blockLength, sectionLength, maxSectionIndex, numPossibilities, maxPossibility;
  std::vector<std::unique_ptr<std::vector<int>>> locations;  // null entries are meaningful (§9-7)
  Matcher(const Seq* q, const SeqSection& refSection, int sectionLen) {
    if (sectionLen < 1) sectionLen = 1;
    blockLength = j2i(std::log((double)(sectionLen * 5)) / std::log(4.0) + 1);
    if (blockLength < 3) blockLength = 3;
    reference = refSection.seq; referenceStart = refSection.start; referenceLength = refSection.length();
    sectionLength = sectionLen; query = q;
    maxSectionIndex = getSectionIndex(reference->length() - 1);
    numPossibilities = j2i(std::pow(4.0, (double)blockLength));
    maxPossibility = numPossibilities - 1;
  }
  static int encodedCharToInt(uint8_t b) {
    switch (b) { case 1: return 0; case 2: return 1; case 4: return 2; case 8: return 3; }
    throw std::invalid_argument("invalid encoded char");
  }
  int getSectionIndex(int referenceIndex) const { return (referenceIndex - referenceStart) / sectionLength; }
  int encodeBlock(const Seq* s, int index) const {  // :79-91
    if (index + blockLength > s->length()) return UNKNOWN;
    int sum = 0;
    for (int i = 0; i < blockLength; i++) {
      uint8_t here = s->at(index + i);
      if (bp_isAmbiguous(here)) return UNKNOWN;
      sum = sum * 4 + encodedCharToInt(here);
    }
    return sum;
  }
  void indexSection(int sectionIndex, std::vector<int>& section) {  // :40-77 (stale previousEncoded after an ambiguity kept)
    for (int i = 0; i < numPossibilities; i++) section[(size_t)i] = NO_MATCHES;
    int previousEncoded = UNKNOWN;
    int startIndex = referenceStart + sectionIndex * sectionLength;
    int endIndex = std::min(startIndex + sectionLength, referenceStart + referenceLength - blockLength);
    for (int i = startIndex; i < endIndex; i++) {
      int encoded;
      if (previousEncoded == UNKNOWN) encoded = encodeBlock(reference, i);
      else {
        uint8_t nextChar = reference->at(i + blockLength - 1);
        if (bp_isAmbiguous(nextChar)) encoded = UNKNOWN;
        else encoded = ((previousEncoded * 4) & maxPossibility) + encodedCharToInt(nextChar);
      }
      if (encoded == UNKNOWN) continue;
      int existing = section[(size_t)encoded];
      if (existing == NO_MATCHES) section[(size_t)encoded] = i; else section[(size_t)encoded] = MULTIPLE_MATCHES;
      previousEncoded = encoded;
    }
  }
  std::vector<int>* getSection(int index) {  // :203-215
    if ((int)locations.size() > index) return locations[(size_t)index].get();
    while ((int)locations.size() <= index) locations.push_back(nullptr);
    auto sec = std::make_unique<std::vector<int>>((size_t)numPossibilities);
    indexSection(index, *sec);
    locations[(size_t)index] = std::move(sec);
    return locations[(size_t)index].get();
  }
  bool canPositionsMatch(int queryIndex, int referenceIndex) const {  // :159-171
    if (referenceIndex + blockLength > referenceStart + referenceLength) return false;
    for (int i = 0; i < blockLength; i++) {
      if (!bp_canMatch(query->at(queryIndex), reference->at(referenceIndex))) return false;
      queryIndex++; referenceIndex++;
    }
    return true;
  }
  int scanSection(int queryIndex, int sectionIndex) const {  // :143-157
    int result = NO_MATCHES;
    int startIndex = referenceStart + sectionIndex * sectionLength;
    int endIndex = startIndex + sectionLength;
    for (int i = startIndex; i < endIndex; i++) {
      if (canPositionsMatch(queryIndex, i)) { if (result == NO_MATCHES) result = i; else return MULTIPLE_MATCHES; }
    }
    return result;
  }
  int lookup(int queryIndex, int minReferenceIndex, int maxReferenceIndex) {  // :98-141
    if (minReferenceIndex < 0) return UNKNOWN;
    if (maxReferenceIndex > reference->length()) return UNKNOWN;
    int encoded = encodeBlock(query, queryIndex);
    if (encoded < 0) return UNKNOWN;
    int matched = NO_MATCHES;
    int minSection = std::max(0, getSectionIndex(minReferenceIndex));
    int maxSection = std::min(maxSectionIndex, getSectionIndex(maxReferenceIndex));
    for (int si = minSection; si <= maxSection; si++) {
      std::vector<int>* section = getSection(si);
      int lookedUp;
      if (sectionLength < 3) lookedUp = scanSection(queryIndex, si);
      else { if (section != nullptr) lookedUp = (*section)[(size_t)encoded]; else return UNKNOWN; }
      if (lookedUp == UNKNOWN) return UNKNOWN;
      if (lookedUp == MULTIPLE_MATCHES) return MULTIPLE_MATCHES;
      if (lookedUp == NO_MATCHES) continue;
      if (lookedUp < minReferenceIndex || lookedUp > maxReferenceIndex) continue;
      if (matched != NO_MATCHES) return MULTIPLE_MATCHES;
      matched = lookedUp;
    }
    return matched;
  }
};

struct Analysis {  // M/AlignmentAnalysis.java
  std::shared_ptr<Matcher> matcher;
  int predictedBestOffset = 0, lastCheckedOffset = 0;
  bool confident = false;
  double maxIns = 1000000, maxDel = 1000000;
  Analysis child() const { return *this; }
};

struct LocalAligner {  // M/LocalAligner.java
  virtual ~LocalAligner() {}
  virtual SeqAlnP align(const SeqSection& q, const SeqSection& r, const Params& p, Analysis& a) = 0;
  OracleStats* stats = nullptr;
  virtual void setStats(OracleStats* s) { stats = s; }
};

struct StraightAligner : LocalAligner {  // M/StraightAligner.java
  std::unique_ptr<LocalAligner> next;
  explicit StraightAligner(std::unique_ptr<LocalAligner> n) : next(std::move(n)) {}
  void setStats(OracleStats* s) override { stats = s; next->setStats(s); }
  static SeqAlnP straightAlignment(const SeqSection& q, const SeqSection& r, const Params& p, const Analysis& a) {  // :73-94
    int qs = q.start, qe = q.end, rs = r.start, re = r.end, off = a.predictedBestOffset;
    if (qs + off > rs) rs = qs + off; else qs = rs - off;
    if (qe + off < re) re = qe + off; else qe = re - off;
    std::vector<ABlock> blocks{ABlock{q.seq, r.seq, qs, rs, qe - qs, re - rs}};
    return newSeqAln(p, blocks, q.seq->complementedFrom != nullptr);
  }
  SeqAlnP align(const SeqSection& q, const SeqSection& r, const Params& p, Analysis& a) override {  // :13-71
    a.lastCheckedOffset = a.predictedBestOffset;
    if (stats) stats->straightCalls++;
    SeqAlnP simple = straightAlignment(q, r, p, a);
    double simplePenalty = simple->alignedPenalty;
    double maxInteresting = q.length() * p.MaxErrorRate;
    double indelPenalty = std::min(p.startingInsertionStartPenalty() + p.InsertionExtension_Penalty, p.DeletionStart_Penalty + p.DeletionExtension_Penalty);
    if (simplePenalty <= 0) return simple;
    if (a.confident) {
      if (simplePenalty <= indelPenalty || (a.maxIns <= 0 && a.maxDel <= 0)) {
        if (simplePenalty <= maxInteresting) return simple;
        return nullptr;
      }
      if (indelPenalty > maxInteresting) return nullptr;
    }
    double rate = simple->alignedPenalty / q.length();
    Params sub = p;
    sub.MaxErrorRate = std::min(rate, p.MaxErrorRate);
    SeqAlnP aln = next->align(q, r, sub, a);
    if (aln == nullptr || aln->alignedPenalty >= simplePenalty) {
      if (simplePenalty <= maxInteresting) return simple;
    }
    return aln;
  }
};

struct SkipHighAmbiguityAligner : LocalAligner {  // M/SkipHighAmbiguity_Aligner.java
  std::unique_ptr<LocalAligner> next;
  explicit SkipHighAmbiguityAligner(std::unique_ptr<LocalAligner> n) : next(std::move(n)) {}
  void setStats(OracleStats* s) override { stats = s; next->setStats(s); }
  SeqAlnP align(const SeqSection& q, const SeqSection& r, const Params& p, Analysis& a) override {
    int numAmb = 0;
    for (int i = r.start; i < r.end; i++) {
      char c = bp_decode(r.seq->at(i));
      if (c != 'A' && c != 'C' && c != 'G' && c != 'T' && c != '-') numAmb++;
    }
    if (numAmb >= r.length() / 4) return nullptr;
    return next->align(q, r, p, a);
  }
};

struct CountMap {  // M/CountMap.java
  int mostPopularKey = 0, mostPopularCount = 0;
  bool haveCounts = false;
  std::unordered_map<int, int> counts;
  void add(int key, int value) {
    if (key == mostPopularKey || mostPopularCount == 0) {
      mostPopularCount += value; mostPopularKey = key;
      if (haveCounts) counts[mostPopularKey] = mostPopularCount;
    } else {
      if (!haveCounts) { haveCounts = true; counts[mostPopularKey] = mostPopularCount; }
      int count;
      auto it = counts.find(key);
      if (it == counts.end()) count = value; else count = it->second + value;
      counts[key] = count;
      if (count > mostPopularCount) { mostPopularKey = key; mostPopularCount = count; }
    }
  }
};

struct PenaltyAnalysis {  // M/PenaltyAnalysis.java
  double minPossiblePenalty = 0, maxIns = 0, maxDel = 0;
  int offsetWithMost = 0, numWithBest = 0;
};

struct HashBlockAligner : LocalAligner {  // M/HashBlock_Aligner.java
  std::unique_ptr<LocalAligner> next;
  explicit HashBlockAligner(std::unique_ptr<LocalAligner> n) : next(std::move(n)) {}
  void setStats(OracleStats* s) override { stats = s; next->setStats(s); }

  SeqAlnP align(const SeqSection& q, const SeqSection& r, const Params& p, Analysis& a) override {  // :21-81
    double maxInteresting = p.MaxErrorRate * q.length();
    if (q.length() > r.length()) return next->align(q, r, p, a);
    PenaltyAnalysis pa = analyzePenalty(q, r, p, a);
    if (pa.minPossiblePenalty > maxInteresting) return nullptr;
    int offsetWithMost = pa.offsetWithMost;
    int numWithBest = pa.numWithBest;
    Analysis sub = a.child();
    sub.maxIns = pa.maxIns; sub.maxDel = pa.maxDel;
    double extra = numWithBest * p.MutationPenalty + pa.minPossiblePenalty;
    if (extra > maxInteresting) { sub.predictedBestOffset = offsetWithMost; sub.confident = true; }
    else { if (!a.confident) sub.predictedBestOffset = offsetWithMost; }
    if (a.confident && sub.predictedBestOffset == a.predictedBestOffset) sub.confident = true;
    SeqSection rsub = r;
    if (sub.confident) {
      int maxDelLen = j2i((double)pa.maxDel / (double)p.DeletionExtension_Penalty);
      int maxInsLen = j2i((double)pa.maxIns / (double)p.InsertionExtension_Penalty);
      int maxIndel = std::max(maxDelLen, maxInsLen);
      int rs = std::max(r.start, q.start + sub.predictedBestOffset - maxIndel);
      int re = std::min(r.end, q.end + sub.predictedBestOffset + maxIndel);
      rsub = SeqSection{r.seq, rs, re};
    }
    if (rsub.length() < r.length()) return this->align(q, rsub, p, sub);
    return next->align(q, rsub, p, sub);
  }

  static double minIndelPenaltyForBlockMismatches(int numMismatches, const Params& p) {  // :286-310
    numMismatches = std::max(1, numMismatches);
    double perInitialIndel = std::min(p.startingInsertionStartPenalty() + p.InsertionExtension_Penalty, p.DeletionStart_Penalty + p.DeletionExtension_Penalty);
    double perExtension = std::min(p.InsertionExtension_Penalty, p.DeletionExtension_Penalty);
    double perSubsequentIndel = std::min(p.InsertionStart_Penalty + p.InsertionExtension_Penalty, p.DeletionStart_Penalty + p.DeletionExtension_Penalty);
    double perSubsequentChange = std::min(p.MutationPenalty, perSubsequentIndel);
    if (numMismatches <= 1) return perInitialIndel;
    if (numMismatches <= 2) return perInitialIndel + perExtension;
    return perInitialIndel + perExtension + (numMismatches - 2) * perSubsequentChange;
  }
  static bool isTooManyMismatches(int n, const Params& p, double maxInteresting) { return minIndelPenaltyForBlockMismatches(n, p) > maxInteresting; }

  static double maxExtLongInsertion(int numMismatches, double totalPenalty, const Params& p, int blockLength) {  // :322-354
    double available = totalPenalty - p.startingInsertionStartPenalty();
    double onlySNPs = numMismatches * p.MutationPenalty;
    double perBlockExt = blockLength * p.InsertionExtension_Penalty;
    double extraPerBlockExt = perBlockExt - p.MutationPenalty;
    if (extraPerBlockExt <= 0) return available;
    if (numMismatches < 2) return available;
    double shortExt = 2 * p.InsertionExtension_Penalty;
    if (shortExt > available) return available;
    double shortSNPs = 2 * p.MutationPenalty;
    double maxIncreasePastAllSNPs = available - onlySNPs;
    double maxForBlockExtensions = maxIncreasePastAllSNPs + shortSNPs - shortExt;
    double maxNumBlockExtensions = maxForBlockExtensions / extraPerBlockExt;
    double r = (maxNumBlockExtensions * blockLength + 2) * p.InsertionExtension_Penalty;
    r = std::min(r, available);
    if (r < shortExt) r = 0;
    return r;
  }
  static double maxExtManyInsertions(int numMismatches, double totalPenalty, const Params& p) {  // :356-376
    double available = totalPenalty + (p.InsertionStart_Penalty - p.startingInsertionStartPenalty());
    double onlySNPs = numMismatches * p.MutationPenalty;
    double perShortIndel = p.InsertionStart_Penalty + 2 * p.InsertionExtension_Penalty;
    double extraPerShortIndel = perShortIndel - 2 * p.MutationPenalty;
    if (extraPerShortIndel <= 0) return available;
    double maxNum = (available - onlySNPs) / extraPerShortIndel;
    if (maxNum < 1) maxNum = 0;
    double r = maxNum * 2 * p.InsertionExtension_Penalty;
    return std::min(r, available);
  }
  static double maxExtManyDeletions(int numMismatches, double totalPenalty, const Params& p) {  // :378-400
    double available = totalPenalty;
    double onlySNPs = numMismatches * p.MutationPenalty;
    double perShortIndel = p.DeletionStart_Penalty + 2 * p.DeletionExtension_Penalty;
    double extraPerShortIndel = perShortIndel - 2 * p.MutationPenalty;
    if (extraPerShortIndel <= 0) return available;
    double maxNum = (available - onlySNPs) / extraPerShortIndel;
    if (maxNum < 1) maxNum = 0;
    double r = maxNum * 2 * p.DeletionExtension_Penalty;
    r = std::min(r, available);
    if (r < 0) r = 0;
    return r;
  }

  PenaltyAnalysis analyzePenalty(const SeqSection& q, const SeqSection& r, const Params& p, Analysis& a) {  // :94-283
    const Seq* query = q.seq; const Seq* reference = r.seq;
    std::shared_ptr<Matcher> matcher = a.matcher;
    double maxInteresting = p.MaxErrorRate * q.length();
    int numMismatches = 0;
    int maxNonmatchingBlockEnd = q.start;
    CountMap counts;
    int numLateIns = 0, numLateDel = 0;
    int minPossibleOffset = r.start - q.start;
    int maxPossibleOffset = r.end - q.end;
    int lookupUncertainty = maxPossibleOffset - minPossibleOffset;
    if (matcher == nullptr || std::abs(matcher->sectionLength - lookupUncertainty) > lookupUncertainty / 2) {
      matcher = std::make_shared<Matcher>(query, r, lookupUncertainty);
      if (a.matcher == nullptr) a.matcher = matcher;
    }
    int blockLength = matcher->blockLength;
    int maxBlockStart = q.end - blockLength;
    for (int bs = q.start; bs <= maxBlockStart; bs++) {
      if (bs >= maxNonmatchingBlockEnd) {
        int position = matcher->lookup(bs, bs + minPossibleOffset, bs + maxPossibleOffset + 1);
        int offset = position - bs;
        if (position == Matcher::UNKNOWN || position == Matcher::MULTIPLE_MATCHES) continue;
        if (position == Matcher::NO_MATCHES) {
          numMismatches++;
          maxNonmatchingBlockEnd = bs + blockLength;
          if (isTooManyMismatches(numMismatches, p, maxInteresting)) break;
          continue;
        }
        int otherStart = position;
        int reverseCount = std::min(bs - maxNonmatchingBlockEnd, otherStart);
        bool foundMismatch = false;
        for (int i = 1; i <= reverseCount; i++) {
          int ia = bs - i, ib = otherStart - i;
          if (!bp_canMatch(query->at(ia), reference->at(ib))) {
            numMismatches++; foundMismatch = true; maxNonmatchingBlockEnd = bs + blockLength; break;
          }
        }
        if (!foundMismatch) {
          int forwardShift = q.end - bs;
          for (int i = blockLength; i < forwardShift; i++) {
            int ia = bs + i, ib = otherStart + i;
            uint8_t ca = query->at(ia);
            uint8_t cb = (ib < r.end) ? reference->at(ib) : (uint8_t)0;
            if (!bp_canMatch(ca, cb)) { numMismatches++; foundMismatch = true; maxNonmatchingBlockEnd = ia + 1; break; }
          }
          if (!foundMismatch) maxNonmatchingBlockEnd = q.end;
          int numOther = 0;
          int forwardShift2 = maxNonmatchingBlockEnd - bs - blockLength;
          for (int i = blockLength; i < forwardShift2; i++) {
            int ia = bs + i;
            int res = matcher->lookup(ia, ia + minPossibleOffset, ia + maxPossibleOffset + 1);
            int offset2 = res - ia;
            if (res >= 0 && offset2 == offset) { numOther++; i = i - 1 + blockLength; }
          }
          if (offset != counts.mostPopularKey && counts.mostPopularCount > 0) {
            if (offset > counts.mostPopularKey) numLateDel += numOther; else numLateIns += numOther;
          }
          counts.add(offset, numOther);
        }
        if (foundMismatch) { if (isTooManyMismatches(numMismatches, p, maxInteresting)) break; }
        else counts.add(offset, 1);
      }
    }
    int mostPopularOffset = counts.mostPopularKey;
    int mostPopularCount = counts.mostPopularCount;
    PenaltyAnalysis res;
    res.minPossiblePenalty = minIndelPenaltyForBlockMismatches(numMismatches, p);
    bool couldDiffer = mostPopularCount < 1 || a.lastCheckedOffset != mostPopularOffset;
    if (couldDiffer) {
      double mismatchPenalty = numMismatches * p.MutationPenalty;
      if (res.minPossiblePenalty > mismatchPenalty) res.minPossiblePenalty = mismatchPenalty;
    }
    // setMaxExtensionPenalty :313-319
    double longIns = maxExtLongInsertion(numMismatches + numLateDel, maxInteresting, p, blockLength);
    double manyIns = maxExtManyInsertions(numMismatches + numLateIns, maxInteresting, p);
    res.maxIns = std::max(longIns, manyIns);
    res.maxDel = maxExtManyDeletions(numMismatches + numLateIns, maxInteresting, p);
    if (res.maxIns > a.maxIns) res.maxIns = a.maxIns;
    if (res.maxDel > a.maxDel) res.maxDel = a.maxDel;
    if (mostPopularCount < 1) mostPopularOffset = a.predictedBestOffset;
    res.offsetWithMost = mostPopularOffset;
    res.numWithBest = mostPopularCount;
    return res;
  }
};

struct BlockAligner : LocalAligner {  // M/BlockAligner.java
  std::unique_ptr<LocalAligner> next;
  explicit BlockAligner(std::unique_ptr<LocalAligner> n) : next(std::move(n)) {}
  void setStats(OracleStats* s) override { stats = s; next->setStats(s); }

  SeqAlnP align(const SeqSection& q, const SeqSection& r, const Params& p, Analysis& a) override {  // :17-36
    double maxInteresting = p.MaxErrorRate * q.length();
    std::vector<SeqAlnP> alns;
    if (!initialAlignments(q, r, p, a, alns) || alns.empty()) return nullptr;
    bool even = false;
    while (alns.size() > 1) {
      std::vector<SeqAlnP> joined;
      if (!joinAlignments(alns, r, p, maxInteresting, a, even, joined)) return nullptr;
      alns = joined;
      even = !even;
    }
    return alns[0];
  }
  bool initialAlignments(const SeqSection& q, const SeqSection& r, const Params& p, Analysis& a, std::vector<SeqAlnP>& result) {  // :39-96
    const Seq* query = q.seq;
    double maxInteresting = p.MaxErrorRate * query->length();
    int numBasesToEncode = j2i(std::log((double)r.length() / std::log(4.0))) + 1;  // sic :48
    int numHashblocks = q.length() / numBasesToEncode + 1;
    int targetPerBlock = j2i(std::sqrt((double)numHashblocks)) + 1;
    int targetBlockSize = targetPerBlock * numBasesToEncode;
    int numBlocks = q.length() / targetBlockSize;
    result.assign((size_t)std::max(0, numBlocks), nullptr);
    double usedPenalty = 0;
    int numRemaining = numBlocks;
    while (true) {
      bool failed = false, failedThenFound = false;
      int startPosition = q.start;
      for (int i = 0; i < numBlocks; i++) {
        int endPosition = q.start + (int)((long long)q.length() * (i + 1) / numBlocks);
        if (result[(size_t)i] == nullptr) {
          SeqSection qsub{query, startPosition, endPosition};
          double averagePenalty = (maxInteresting - usedPenalty) / numRemaining;
          SeqAlnP sub = alignPiece(qsub, r, averagePenalty, p, i == 0, a);
          if (sub != nullptr) {
            if (failed) failedThenFound = true;
            numRemaining--;
            result[(size_t)i] = sub;
            usedPenalty += sub->alignedPenalty;
          } else failed = true;
        }
        startPosition = endPosition;
      }
      if (numRemaining < 1) return true;
      if (!failedThenFound) return false;
    }
  }
  bool joinAlignments(const std::vector<SeqAlnP>& alns, const SeqSection& r, const Params& p, double maxInteresting, Analysis& a, bool allowSimpleMerges, std::vector<SeqAlnP>& result) {  // :99-144
    double usedPenalty = 0;
    for (auto& x : alns) usedPenalty += x->alignedPenalty;
    for (int i = 0; i < (int)alns.size(); i += 2) {
      SeqAlnP merge;
      SeqAlnP left = alns[(size_t)i];
      if (i + 1 < (int)alns.size()) {
        SeqAlnP right = alns[(size_t)i + 1];
        merge = doTryMerge(left, right, p);
        if (merge == nullptr) {
          usedPenalty -= left->alignedPenalty;
          usedPenalty -= right->alignedPenalty;
          SeqSection qsub{left->seqA(), left->startA(), right->endA()};
          merge = alignPiece(qsub, r, maxInteresting - usedPenalty, p, i == 0, a);
          if (merge == nullptr) return false;
          usedPenalty += merge->alignedPenalty;
        } else {
          if (!allowSimpleMerges) { result.push_back(left); i--; continue; }
        }
      } else merge = left;
      result.push_back(merge);
    }
    return true;
  }
  static SeqAlnP doTryMerge(const SeqAlnP& left, const SeqAlnP& right, const Params& p) {  // :158-212
    if (left->endB() != right->startB()) return nullptr;
    const ABlock& l = left->sections.back();
    const ABlock& rr = right->sections.front();
    if (!l.sameIndelType(rr)) return nullptr;
    if (l.aEnd() != rr.aStart) return nullptr;
    if (l.bEnd() != rr.bStart) return nullptr;
    ABlock mid{l.a, l.b, l.aStart, l.bStart, l.aLen + rr.aLen, l.bLen + rr.bLen};
    std::vector<ABlock> sections;
    for (size_t i = 0; i + 1 < left->sections.size(); i++) sections.push_back(left->sections[i]);
    sections.push_back(mid);
    for (size_t i = 1; i < right->sections.size(); i++) sections.push_back(right->sections[i]);
    return newSeqAln(p, sections, left->referenceReversed);
  }
  SeqAlnP alignPiece(const SeqSection& q, const SeqSection& r, double maxPenalty, const Params& p, bool firstPiece, Analysis& parent) {  // :215-249
    if (maxPenalty < 0) return nullptr;
    SeqSection rsub = r;
    if (parent.confident) {
      int maxInsLen = j2i((double)parent.maxIns / (double)p.InsertionExtension_Penalty);
      int maxDelLen = j2i((double)parent.maxDel / (double)p.DeletionExtension_Penalty);
      int maxIndel = std::max(maxInsLen, maxDelLen);
      // Java int arithmetic wraps; operands stay far from the limits for real inputs
      int rs = std::max(r.start, wadd(wadd(q.start, parent.predictedBestOffset), -maxIndel));
      int re = std::min(r.end, wadd(wadd(q.end, parent.predictedBestOffset), maxIndel));
      if (re > rs) rsub = SeqSection{r.seq, rs, re};
    }
    Params sub = p;
    if (!firstPiece) sub.StartingInsertionStartFree = true;
    sub.MaxErrorRate = maxPenalty / q.length();
    Analysis child = parent.child();
    child.confident = false;
    return next->align(q, rsub, sub, child);
  }
};

struct PathAligner : LocalAligner {  // M/PathAligner.java (+ PathAligner_Runner: fresh state per call)
  static constexpr double disallowed = 1000000.0;
  struct Node { int x, y; double penalty, insX, insY; bool main, other; };
  struct Slot { bool present = false; Node n; };
  // per-call state
  Params prm; double maxInteresting;
  const Seq* query; const Seq* reference;
  int startIndexA, endIndexA, startIndexB, endIndexB, textALength, textBLength;
  std::vector<uint8_t> qc, rc;
  Analysis* analysis;
  int diagonal, stepDelta; bool searchReverse, mayQueryExtendPastEndOfReference;
  int startX, startY, goalX, goalY;
  std::map<double, std::vector<Node>> prioritized;
  std::vector<std::vector<Slot>> located;
  double activePenalty;

  bool chooseSearchReverse() {  // :17-53
    int sumMis = 0, numMis = 0, sumMatch = 0, numMatch = 0;
    int offset = analysis->predictedBestOffset;
    int si = std::max(startIndexA, startIndexB - offset);
    int ei = std::min(endIndexA, endIndexB - offset);
    int length = ei - si;
    for (int i = 0; i < length; i++) {
      int j = i - diagonal;
      if (j >= 0 && j < (int)rc.size()) {
        if (!bp_canMatch(qc[(size_t)i], rc[(size_t)j])) { sumMis += i; numMis++; } else { sumMatch += i; numMatch++; }
      }
    }
    if (numMis > 1 && numMatch > 1) return (sumMis / numMis) > (sumMatch / numMatch);
    return true;
  }
  int signedDist(int x, int y) const { return x - y - diagonal; }
  double estimateOverallPenalty(const Node& n) const {  // :475-521
    if (!analysis->confident) return n.penalty;
    int sd = signedDist(n.x, n.y);
    if (n.main) {
      if (sd * stepDelta > 0) {
        double ie = std::fabs(sd * prm.InsertionExtension_Penalty);
        if (ie > analysis->maxIns) return disallowed;
      } else {
        double de = std::fabs(sd * prm.DeletionExtension_Penalty);
        if (de > analysis->maxDel) return disallowed;
      }
      if (n.other) return n.penalty;
      double indel = std::min(prm.InsertionStart_Penalty + prm.InsertionExtension_Penalty, prm.DeletionStart_Penalty + prm.DeletionExtension_Penalty);
      return n.penalty + indel;
    }
    if (sd * stepDelta < 0) {
      double ie = std::fabs(sd * prm.InsertionExtension_Penalty);
      if (ie > analysis->maxIns) return disallowed;
      double is = std::min(prm.InsertionStart_Penalty, n.insX - n.penalty);
      return n.penalty + is + ie;
    } else {
      double de = std::fabs(sd * prm.DeletionExtension_Penalty);
      if (de > analysis->maxDel) return disallowed;
      double ds = std::min(prm.DeletionStart_Penalty, n.insY - n.penalty);
      return n.penalty + ds + de;
    }
  }
  static int encodeDiag(int x, int y) { int e = (y - x) * 2; if (e < 0) e = -e - 1; return e; }
  void saveNode(const Node& n) {  // :523-539
    if (n.x < 0 || n.y < 0) return;
    while ((int)located.size() <= n.x) located.emplace_back();
    auto& d = located[(size_t)n.x];
    int e = encodeDiag(n.x, n.y);
    while ((int)d.size() <= e) d.emplace_back();
    d[(size_t)e].present = true; d[(size_t)e].n = n;
  }
  const Node* getNode(int x, int y) const {  // :541-553
    if (x < 0 || (int)located.size() <= x) return nullptr;  // Java: negative x throws; never happens (x >= 0 by construction)
    auto& d = located[(size_t)x];
    int e = encodeDiag(x, y);
    if (e >= (int)d.size()) return nullptr;
    return d[(size_t)e].present ? &d[(size_t)e].n : nullptr;
  }
  void putNode(const Node& n) {  // :446-473
    double est = estimateOverallPenalty(n);
    if (est < activePenalty) est = activePenalty;
    prioritized[est].push_back(n);
    saveNode(n);
  }
  void update(int x, int y) {  // :555-571
    if (x <= 0 || x > textALength) return;
    if (y <= 0 || y > textBLength) return;
    Node nn;
    if (computeUpdated(x, y, nn)) putNode(nn);
  }
  bool computeUpdated(int x, int y, Node& out) {  // :573-719
    const Node* existing = getNode(x, y);
    const Node* left = getNode(x - stepDelta, y);
    const Node* up = getNode(x, y - stepDelta);
    const Node* diag = getNode(x - stepDelta, y - stepDelta);
    double insX = disallowed, insY = disallowed, overlay = disallowed;
    if (diag != nullptr) overlay = diag->penalty + prm.basePenalty(qc[(size_t)x - 1], rc[(size_t)y - 1]);
    if (left != nullptr) {
      if (y == goalY && mayQueryExtendPastEndOfReference) insX = left->penalty + prm.UnalignedPenalty;
      else {
        bool allowed = true;
        int pa = x - 1 - stepDelta, pb = y - 1;
        if (pa >= 0 && pa < textALength && pb >= 0 && pb < textBLength) {
          if (!bp_canMatch(qc[(size_t)pa], rc[(size_t)pb])) allowed = false;
        }
        if (allowed) {
          int na = x - 1, nb = y - 1 + stepDelta;
          if (na >= 0 && na < textALength && nb >= 0 && nb < textBLength) {
            uint8_t a = qc[(size_t)na], b = rc[(size_t)nb];
            if (prm.basePenalty(a, b) == 0) allowed = false;
            else if (bp_isFullyAmbiguous(a) || bp_isFullyAmbiguous(b)) allowed = false;
          }
        }
        double newIns = allowed ? left->penalty + prm.InsertionStart_Penalty + prm.InsertionExtension_Penalty : disallowed;
        double extIns = left->insX + prm.InsertionExtension_Penalty;
        insX = std::min(extIns, newIns);
      }
    }
    if (up != nullptr) {
      bool allowed = true;
      int pa = x - 1, pb = y - 1 - stepDelta;
      if (pa >= 0 && pa < textALength && pb >= 0 && pb < textBLength) {
        if (!bp_canMatch(qc[(size_t)pa], rc[(size_t)pb])) allowed = false;
      }
      if (allowed) {
        int na = x - 1 + stepDelta, nb = y - 1;
        if (na >= 0 && na < textALength && nb >= 0 && nb < textBLength) {
          uint8_t a = qc[(size_t)na], b = rc[(size_t)nb];
          if (prm.basePenalty(a, b) == 0) allowed = false;
          else if (bp_isFullyAmbiguous(a) || bp_isFullyAmbiguous(b)) allowed = false;
        }
      }
      double newDel = allowed ? up->penalty + prm.DeletionStart_Penalty + prm.DeletionExtension_Penalty : disallowed;
      double extDel = up->insY + prm.DeletionExtension_Penalty;
      insY = std::min(extDel, newDel);
    }
    double best = std::min(std::min(overlay, insX), insY);
    if (existing == nullptr || best < existing->penalty || insX < existing->insX || insY < existing->insY) {
      bool m = false, o = false;
      if (best != disallowed) {
        if (best == overlay) { m = diag->main; o = diag->other; }
        else if (best == insX) { m = left->main; o = left->other; }
        else { m = up->main; o = up->other; }
        if (std::abs(signedDist(x, y)) == 0) m = true; else o = true;
      }
      out = Node{x, y, best, insX, insY, m, o};
      return true;
    }
    return false;
  }
  void explore(int x, int y) { update(x + stepDelta, y); update(x, y + stepDelta); update(x + stepDelta, y + stepDelta); }

  static bool canRemoveSection(const ABlock& b) {  // :358-366
    if (b.aLen <= 0 && b.bLen <= 0) return true;
    if ((b.aStart <= 0 && b.aLen <= 0) || (b.bStart <= 0 && b.bLen <= 0)) return true;
    return false;
  }
  SeqAlnP justify(std::vector<ABlock>& s) {  // :307-352
    for (int i = 1; i < (int)s.size() - 1; i++) {
      while (true) {
        ABlock left = s[(size_t)i - 1], middle = s[(size_t)i], right = s[(size_t)i + 1];
        if ((middle.aLen > 0) == (middle.bLen > 0)) break;
        if (left.aLen == 0 || left.bLen == 0) break;
        if (right.aLen == 0 || right.bLen == 0) break;
        if (middle.aLen > 0) { if (left.a->at(left.aEnd() - 1) != middle.a->at(middle.aEnd() - 1)) break; }
        else { if (left.b->at(left.bEnd() - 1) != middle.b->at(middle.bEnd() - 1)) break; }
        left.aLen -= 1; left.bLen -= 1;
        middle.aStart -= 1; middle.bStart -= 1;
        right.aStart -= 1; right.bStart -= 1; right.aLen += 1; right.bLen += 1;
        s[(size_t)i - 1] = left; s[(size_t)i] = middle; s[(size_t)i + 1] = right;
      }
    }
    while (true) {
      if (s.empty()) throw std::runtime_error("PathAligner.justify: sections exhausted (IndexOutOfBounds in the reference)");
      if (!canRemoveSection(s[0])) break;
      s.erase(s.begin());
    }
    return newSeqAln(prm, s, query->complementedFrom != nullptr);
  }

  SeqAlnP align(const SeqSection& q, const SeqSection& r, const Params& p, Analysis& a) override {  // :55-293
    prm = p;
    maxInteresting = q.length() * p.MaxErrorRate;
    prioritized.clear(); located.clear();
    query = q.seq; startIndexA = q.start; endIndexA = q.end;
    qc.assign(query->codes.begin() + q.start, query->codes.begin() + q.end);
    reference = r.seq; startIndexB = r.start; endIndexB = r.end;
    rc.assign(reference->codes.begin() + r.start, reference->codes.begin() + r.end);
    textALength = q.length(); textBLength = r.length();
    analysis = &a;
    diagonal = startIndexB - (startIndexA + a.predictedBestOffset);
    searchReverse = chooseSearchReverse();
    activePenalty = 0;
    if (searchReverse) { stepDelta = -1; mayQueryExtendPastEndOfReference = startIndexB == 0; }
    else { stepDelta = 1; mayQueryExtendPastEndOfReference = endIndexB == reference->length(); }
    int width = textALength + 2, height = endIndexB - startIndexB + 2;
    if (searchReverse) { startX = width - 1; startY = height - 1; goalX = 1; goalY = 1; }
    else { startX = 0; startY = 0; goalX = width - 2; goalY = height - 2; }
    if (stats) { stats->pathAlignerCalls++; stats->pathAlignerCells += (long long)textALength * (long long)textBLength; }

    if (textBLength >= textALength) {
      double sisp = p.startingInsertionStartPenalty();
      if (!mayQueryExtendPastEndOfReference) sisp = disallowed;
      int cnt = std::max(0, textBLength - textALength) + 1;
      for (int i = 0; i < cnt; i++) putNode(Node{startX, startY + i * stepDelta, 0, sisp, disallowed, false, false});
    } else {
      int cnt = std::max(0, textALength - textBLength) + 1;
      for (int i = 0; i < cnt; i++) putNode(Node{startX + i * stepDelta, startY, 0, disallowed, disallowed, false, false});
    }
    if (mayQueryExtendPastEndOfReference) {
      int cnt = j2i(a.maxIns / p.DeletionExtension_Penalty);
      for (int i = 1; i < cnt; i++) putNode(Node{startX + i * stepDelta, startY, i * p.UnalignedPenalty, disallowed, disallowed, false, false});
    }

    bool haveLast = false; Node lastNode{};
    while (!haveLast) {
      if (prioritized.empty()) throw std::runtime_error("PathAligner: priority queue empty (NullPointerException in the reference)");
      auto it = prioritized.begin();
      activePenalty = it->first;
      for (size_t i = 0; i < it->second.size(); i++) {
        if (stats) stats->pathAlignerSteps++;
        Node node = it->second[i];
        if (activePenalty > maxInteresting + 0.000001) return nullptr;
        if (node.x == goalX) { lastNode = node; haveLast = true; break; }
        explore(node.x, node.y);
      }
      prioritized.erase(it);
    }
    int i = lastNode.x, j = lastNode.y;
    std::vector<ABlock> blocks;
    auto need = [&](int x, int y) -> const Node& {
      const Node* n = getNode(x, y);
      if (!n) throw std::runtime_error("PathAligner traceback: missing node (NullPointerException in the reference)");
      return *n;
    };
    while (i != startX && j != startY) {
      const Node& node = need(i, j);
      double best = node.penalty, ix = node.insX, iy = node.insY;
      if (best == ix) {
        int oldI = i;
        i -= stepDelta;
        while (i != startX) {
          const Node& other = need(i, j);
          double newIns = other.penalty + p.InsertionStart_Penalty + p.InsertionExtension_Penalty;
          double extIns = other.insX + p.InsertionExtension_Penalty;
          if (newIns < extIns) break;
          i -= stepDelta;
        }
        if (searchReverse) blocks.push_back(ABlock{query, reference, startIndexA + oldI - 1, startIndexB + j - 1, i - oldI, 0});
        else blocks.push_back(ABlock{query, reference, startIndexA + i, startIndexB + j, oldI - i, 0});
      } else if (best == iy) {
        int oldJ = j;
        j -= stepDelta;
        while (j != startY) {
          const Node& other = need(i, j);
          double newDel = other.penalty + p.DeletionStart_Penalty + p.DeletionExtension_Penalty;
          double extDel = other.insY + p.DeletionExtension_Penalty;
          if (newDel < extDel) break;
          j -= stepDelta;
        }
        if (searchReverse) blocks.push_back(ABlock{query, reference, startIndexA + i - 1, startIndexB + oldJ - 1, 0, j - oldJ});
        else blocks.push_back(ABlock{query, reference, startIndexA + i, startIndexB + j, 0, oldJ - j});
      } else {
        int oldI = i, oldJ = j;
        i -= stepDelta; j -= stepDelta;
        while (i != startX && j != startY) {
          const Node& other = need(i, j);
          if (other.penalty == other.insX || other.penalty == other.insY) break;
          i -= stepDelta; j -= stepDelta;
        }
        if (searchReverse) blocks.push_back(ABlock{query, reference, startIndexA + oldI - 1, startIndexB + oldJ - 1, i - oldI, j - oldJ});
        else blocks.push_back(ABlock{query, reference, startIndexA + i, startIndexB + j, oldI - i, oldJ - j});
      }
    }
    if (!searchReverse) std::reverse(blocks.begin(), blocks.end());
    if (blocks.empty()) return nullptr;
    SeqAlnP result = justify(blocks);
    if (result->alignedPenalty > maxInteresting) return nullptr;
    return result;
  }
};

// M/QueryMatch_Aligner.java:18-29
inline std::unique_ptr<LocalAligner> buildAlignerChain() {
  std::unique_ptr<LocalAligner> a = std::make_unique<PathAligner>();
  a = std::make_unique<StraightAligner>(std::move(a));
  a = std::make_unique<HashBlockAligner>(std::move(a));
  a = std::make_unique<StraightAligner>(std::move(a));
  a = std::make_unique<BlockAligner>(std::move(a));
  a = std::make_unique<HashBlockAligner>(std::move(a));
  a = std::make_unique<SkipHighAmbiguityAligner>(std::move(a));
  a = std::make_unique<StraightAligner>(std::move(a));
  return a;
}

}  // namespace xo
