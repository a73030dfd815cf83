// ORACLE — TEST INFRASTRUCTURE ONLY (see xo_core.h header).
// Per-query driver and pair model.
// Follows M/QueryMatch_Aligner.java and M/AlignerWorker.java:256-644.
// AlignmentCache (M/AlignmentCache.java) is a result memo keyed by query text and is not restated.
//
// Stated deviation: QueryMatch_Aligner.withoutDuplicates orders results by
// java.util.HashSet iteration; emulated here for the non-treeified case (bucket = spread(hash) & (n-1),
// insertion order within a bucket).
#pragma once
#include "xo_align.h"

namespace xo {

struct QueryMatchAligner {  // M/QueryMatch_Aligner.java
  Params parameters;
  const Query* query;
  std::unique_ptr<LocalAligner> aligner;
  std::vector<QueryAlnP> goodAlignments;
  double bestPenalty = (double)JMAX;
  std::vector<std::unique_ptr<Seq>> joinedStore;  // keeps "joined" sequences alive
  OracleStats* stats = nullptr;

  QueryMatchAligner(const Query* q, const Params& p, OracleStats* st) : parameters(p), query(q), stats(st) {
    aligner = buildAlignerChain();
    aligner->setStats(st);
  }
  static double divideRoundUp(double a, double b) {  // :56-61
    double r = a / b;
    if (r * b < a) r = nextUp(r);
    return r;
  }
  QueryAlnP align(const QueryMatch& match, double extraSpacing) {  // :39-54
    QueryAlnP aln = doAlign(match, extraSpacing);
    if (aln != nullptr) {
      if (aln->totalPenalty < bestPenalty) {
        bestPenalty = aln->totalPenalty;
        double newTargetPenalty = aln->totalPenalty + parameters.Max_PenaltySpan;
        double newTargetErrorRate = divideRoundUp(newTargetPenalty, (double)query->length());
        if (newTargetErrorRate < parameters.MaxErrorRate) parameters.MaxErrorRate = newTargetErrorRate;
      }
      goodAlignments.push_back(aln);
    }
    return aln;
  }
  std::vector<QueryAlnP> getBestAlignments() {  // :71-83
    double maxAnywhere = query->length() * parameters.MaxErrorRate;
    double cutoff = bestPenalty + parameters.Max_PenaltySpan;
    if (cutoff > maxAnywhere) cutoff = maxAnywhere;
    std::vector<QueryAlnP> best;
    for (auto& a : goodAlignments) if (a->totalPenalty <= cutoff) best.push_back(a);
    return withoutDuplicates(best);
  }
  static std::vector<QueryAlnP> withoutDuplicates(const std::vector<QueryAlnP>& alns) {  // :86-92
    if (alns.size() <= 1) return alns;
    // new HashSet<>(collection): HashMap(max((int)(n/.75f)+1, 16)) -> table = tableSizeFor(cap); no resize during addAll
    int cap = std::max((int)((float)alns.size() / .75f) + 1, 16);
    int n = 1; while (n < cap) n <<= 1;
    std::vector<std::vector<QueryAlnP>> table((size_t)n);
    for (auto& a : alns) {
      uint32_t h = (uint32_t)a->hashCode();
      h ^= (h >> 16);
      auto& bucket = table[h & (uint32_t)(n - 1)];
      bool dup = false;
      for (auto& e : bucket) if (e->hashCode() == a->hashCode() && e->equalsQA(*a)) { dup = true; break; }
      if (!dup) bucket.push_back(a);
    }
    std::vector<QueryAlnP> r;
    for (auto& b : table) for (auto& e : b) r.push_back(e);
    return r;
  }

  int getSpacing(const QueryMatch& m) { return m.comps.size() < 2 ? 0 : m.totalDistanceBetweenComponents(); }
  double computeSpacingPenalty(double innerDistance) {  // :530-546
    double expected = query->expectedInnerDistance;
    int totalLength = query->length();
    if (innerDistance < 0 && innerDistance > -1 * totalLength) return 0;
    double per = query->spacingDeviationPerUnitPenalty;
    int pen = j2i(std::fabs(innerDistance - expected) / per);
    return (double)pen;
  }

  QueryAlnP doAlign(const QueryMatch& match, double extraSpacing) {  // :94-272
    double innerDistance = getSpacing(match) + extraSpacing;
    double spacingPenalty = computeSpacingPenalty(innerDistance);
    double overlapMultiplier = 1, duplicationBonus = 0;
    double maxAllowed = match.queryTotalLength() * parameters.MaxErrorRate;
    maxAllowed = nextUp(maxAllowed);
    if (innerDistance > 0) {
      double minPossible = spacingPenalty + match.priority * parameters.MutationPenalty;
      if (minPossible > maxAllowed) return nullptr;
    }
    std::vector<SeqAlnP> result;
    bool haveResult = false;
    double componentsPenalty = 0;
    int numSeqs = (int)match.comps.size();
    if (numSeqs > 1 && innerDistance < 0) {
      Seq* joined = tryJoinQuerySequences(match);
      if (joined != nullptr) {
        SeqAlnP joinedAln = computeJoinedAlignment(joined, match);
        if (!splitAlignment(joinedAln, match, result)) return nullptr;
        haveResult = true;
        for (auto& c : result) componentsPenalty += c->penalty;
      }
    }
    if (!haveResult) {
      result.assign((size_t)numSeqs, nullptr);
      std::vector<bool> remainingPresent((size_t)numSeqs, true);
      int numRemaining = numSeqs;
      bool forward = match.hintForward;
      int first, stepi, last;
      if (forward) { first = 0; stepi = 1; last = numSeqs; } else { first = numSeqs - 1; stepi = -1; last = -1; }
      double maxTotalComponentPenalty;
      if (innerDistance < 0 && numSeqs > 1) {
        double queryTotalLength = match.queryTotalLength();
        double estimatedOverlap = std::min(-1 * innerDistance, (double)std::min(match.comps[0].a->length(), match.comps[1].a->length()));
        double estimatedUniqueLength = queryTotalLength - estimatedOverlap;
        maxTotalComponentPenalty = divideRoundUp(maxAllowed - spacingPenalty, queryTotalLength) * estimatedUniqueLength * 2;
      } else maxTotalComponentPenalty = maxAllowed - spacingPenalty;
      while (true) {
        int numBases = 0;
        for (int i = 0; i < numSeqs; i++) if (remainingPresent[(size_t)i]) numBases += match.comps[(size_t)i].a->length();
        if (numBases < 1) break;
        double avg = divideRoundUp(maxTotalComponentPenalty - componentsPenalty, (double)numBases);
        Params prem = parameters;
        prem.MaxErrorRate = avg;
        bool found = false;
        for (int i = first; i != last; i += stepi) {
          if (remainingPresent[(size_t)i]) {
            SeqAlnP sa = alignMatch(match.comps[(size_t)i], prem);
            if (sa != nullptr) {
              result[(size_t)i] = sa; found = true; remainingPresent[(size_t)i] = false;
              componentsPenalty += sa->penalty; numRemaining--;
              break;
            }
          }
        }
        if (numRemaining < 1) break;
        if (!found) return nullptr;
      }
    }
    double totalUsed = componentsPenalty;
    if (innerDistance < 0) {
      duplicationBonus = computeDuplicationBonus(result);
      totalUsed -= duplicationBonus;
      double multiplied = multiplyPenaltyForOverlap(result, totalUsed);
      if (totalUsed != 0) overlapMultiplier = multiplied / totalUsed; else overlapMultiplier = 1;
      totalUsed = multiplied;
    }
    totalUsed += spacingPenalty;
    if (totalUsed > maxAllowed) return nullptr;
    int actualInner = result.size() > 1 ? result[1]->startB() - result[0]->endB() : 0;
    auto qa = std::make_shared<QueryAln>();
    qa->comps = result; qa->spacingPenalty = spacingPenalty; qa->overlapMultiplier = overlapMultiplier;
    qa->duplicationBonus = duplicationBonus; qa->totalPenalty = totalUsed; qa->innerDistance = actualInner;
    return qa;
  }

  Seq* tryJoinQuerySequences(const QueryMatch& match) {  // :274-284
    const SeqMatch& m1 = match.comps[0]; const SeqMatch& m2 = match.comps[1];
    int offset = m2.offset - m1.offset;
    if (offset >= 0) return tryJoin(m1.a, m2.a, offset);
    return tryJoin(m2.a, m1.a, -offset);
  }
  Seq* tryJoin(const Seq* s1, const Seq* s2, int offset) {  // :287-319
    int suffixStart = s1->length() - offset;
    if (suffixStart < 0) return nullptr;
    int end2 = std::min(s2->length(), s1->length() - offset);
    for (int i2 = 0; i2 < end2; i2++) if (s1->at(i2 + offset) != s2->at(i2)) return nullptr;
    auto j = std::make_unique<Seq>();
    j->name = "joined";
    j->codes = s1->codes;
    // sequence2.getRange(suffixStartIndex, endIndex - suffixStartIndex): throws in Java if suffixStart > length
    if (suffixStart > s2->length()) throw std::runtime_error("tryJoinQuerySequences: suffix start beyond sequence2 (exception in the reference)");
    for (int i = suffixStart; i < s2->length(); i++) j->codes.push_back(s2->at(i));
    Seq* r = j.get();
    joinedStore.push_back(std::move(j));
    return r;
  }
  SeqAlnP computeJoinedAlignment(Seq* joined, const QueryMatch& orig) {  // :321-331
    int joinedOffset = std::min(orig.comps[0].offset, orig.comps[1].offset);
    SeqMatch jm{joined, orig.comps[0].b, joinedOffset, true};
    Params sub = parameters;
    sub.MaxErrorRate = nextUp(sub.MaxErrorRate);
    return alignMatch(jm, sub);
  }
  bool splitAlignment(const SeqAlnP& joinedAln, const QueryMatch& qm, std::vector<SeqAlnP>& out) {  // :332-363
    if (joinedAln == nullptr) return false;
    const SeqMatch& m1 = qm.comps[0]; const SeqMatch& m2 = qm.comps[1];
    const Seq* s1 = m1.a; const Seq* s2 = m2.a;
    int offset = m2.offset - m1.offset;
    SeqAlnP a1, a2;
    if (offset >= 0) {
      a1 = extract(joinedAln, 0, s1->length(), s1, m1.reversed());
      a2 = extract(joinedAln, offset, s2->length() + offset, s2, m2.reversed());
    } else {
      a2 = extract(joinedAln, 0, s2->length(), s2, m2.reversed());
      a1 = extract(joinedAln, -offset, s1->length() - offset, s1, m1.reversed());
    }
    if (a1 == nullptr || a2 == nullptr) return false;
    out.clear(); out.push_back(a1); out.push_back(a2);
    return true;
  }
  SeqAlnP extract(const SeqAlnP& joinedAln, int queryStart, int queryEnd, const Seq* q, bool reverse) {  // :365-405
    bool referenceReversed = joinedAln->referenceReversed != reverse;
    const Seq* reference = joinedAln->seqB();
    std::vector<ABlock> blocks;
    for (auto& block : joinedAln->sections) {
      if (block.aStart >= queryEnd) break;
      if (block.aEnd() <= queryStart) continue;
      int selStart = std::max(block.aStart, queryStart);
      int selEnd = std::min(block.aEnd(), queryEnd);
      int qLen = selEnd - selStart;
      int rLen, rStart;
      if (block.aLen == block.bLen) { rLen = qLen; rStart = selStart + block.offset(); }
      else if (block.aLen > block.bLen) { rLen = 0; rStart = block.bStart; }
      else { rLen = block.bLen; rStart = selStart + block.offset(); }
      blocks.push_back(ABlock{q, reference, selStart - queryStart, rStart, qLen, rLen});
    }
    if (blocks.empty()) return nullptr;
    return newSeqAln(parameters, blocks, referenceReversed);
  }

  SeqAlnP alignMatch(const SeqMatch& sm, const Params& p) {  // :412-462
    SeqSection qsec{sm.a, sm.startA(), sm.endA()};
    double maxInteresting = qsec.length() * p.MaxErrorRate;
    int maxIndelLength = j2i(std::max((double)0, (double)(maxInteresting - p.DeletionStart_Penalty) / p.DeletionExtension_Penalty));
    int maxShift;
    int bestOffset = sm.offset;
    if (sm.fromHashblockMatch) maxShift = maxIndelLength;
    else {
      maxShift = j2i((double)maxInteresting * (double)query->spacingDeviationPerUnitPenalty);
      if (maxShift < 0) return nullptr;
      if (bestOffset + sm.a->length() > sm.b->length()) bestOffset = sm.b->length() - sm.a->length();
      if (bestOffset < 0) bestOffset = 0;
      qsec = SeqSection{sm.a, 0, sm.a->length()};
    }
    SeqSection rsec{sm.b, std::max(0, sm.startB() - maxShift), std::min(sm.endB() + maxShift, sm.b->length())};
    Analysis an;
    an.maxIns = maxInteresting - p.InsertionStart_Penalty;
    an.maxDel = maxInteresting - p.DeletionStart_Penalty;
    an.predictedBestOffset = bestOffset;
    an.confident = sm.fromHashblockMatch;
    return aligner->align(qsec, rsec, p, an);
  }

  double multiplyPenaltyForOverlap(const std::vector<SeqAlnP>& comps, double totalPenalty) {  // :464-504
    if (comps.size() < 2) return totalPenalty;
    const SeqAln& first = *comps[0]; const SeqAln& second = *comps[1];
    double overlappingLengthB = std::min(first.endB(), second.endB()) - std::max(first.startB(), second.startB());
    if (overlappingLengthB <= 0) return totalPenalty;
    int uniqueLengthA;
    if (first.startB() <= second.startB()) uniqueLengthA = first.lengthABefore(second.startB()) + second.lengthA() + first.lengthAAfter(second.endB());
    else uniqueLengthA = second.lengthABefore(first.startB()) + first.lengthA() + second.lengthAAfter(first.endB());
    double deletion = std::min(first.insertAOrBLength(), second.insertAOrBLength());
    uniqueLengthA = j2i((double)uniqueLengthA - deletion);  // uniqueLengthA -= deletion (compound assignment narrows)
    if (uniqueLengthA <= 0) return totalPenalty;
    int totalLengthA = first.lengthA() + second.lengthA();
    return divideRoundUp(totalPenalty, (double)uniqueLengthA) * totalLengthA;
  }
  double computeDuplicationBonus(const std::vector<SeqAlnP>& comps) {  // :506-520
    if (comps.size() < 2) return 0;
    const SeqAln& a = *comps[0]; const SeqAln& b = *comps[1];
    double overlappingLength = std::min(a.endB(), b.endB()) - std::max(a.startB(), b.startB());
    if (overlappingLength < 0) return 0;
    return (alnPenaltyRange(parameters, a, b.startB(), b.endB()) + alnPenaltyRange(parameters, b, a.startB(), a.endB())) / 2;
  }
};

struct Worker {  // M/AlignerWorker.java
  Index* index; SeqDb* seqdb; DupDetector* dup; Params parameters;
  int shortestHashblockLength;
  OracleStats stats;
  int numImmediateAccepts = 0;
  Worker(Index* ix, DupDetector* d, const Params& p) : index(ix), seqdb(ix->db), dup(d), parameters(p) {
    shortestHashblockLength = ix->minInterestingSize;
  }
  double getPenaltyLowerBound(int k) const {  // :487-491
    double mutationPenalty = k * parameters.MutationPenalty;
    double indelPenalty = shortestHashblockLength * k * parameters.DeletionExtension_Penalty;
    return std::min(mutationPenalty, indelPenalty);
  }

  bool quicklyConfident(const QueryAlnP& best, const QueryMatch& bestMatch) {  // :494-587
    if (best == nullptr) return false;
    if (best->hasIndel()) return false;
    const Seq* originalReference = bestMatch.comps[0].b;
    int matchStart = bestMatch.startIndexB();
    int matchEnd = bestMatch.endIndexB();
    bool hasNearbyDuplication = false;
    double granularity = dup->detectionGranularity();
    double penalty = best->totalPenalty;
    double numberOfMutations = (penalty + parameters.Max_PenaltySpan) / parameters.MutationPenalty;
    double existingMutationRate = numberOfMutations / bestMatch.queryTotalLength();
    if (penalty <= 0 && parameters.Max_PenaltySpan < parameters.minPossibleNonzeroPenalty()) return true;
    double probabilityMutationInSection = 1 - std::pow(1 - existingMutationRate, granularity);
    double acceptableProbability = 1.0 / (double)seqdb->totalFR;
    double numberOfUnmatchedBlocks = std::log(acceptableProbability) / std::log(probabilityMutationInSection);
    double totalLengthForHighConfidence = numberOfUnmatchedBlocks * granularity;
    double matchMiddle = (double)((matchStart + matchEnd) / 2);
    double interestingWindow = std::max(totalLengthForHighConfidence, (double)((matchEnd - matchStart + 1) / 2));
    int windowStart = j2i(matchMiddle - interestingWindow);
    int windowEnd = j2i(matchMiddle + interestingWindow);
    int key;
    if (dup->mayContainDuplicationInRange(originalReference, windowStart, windowEnd, key)) hasNearbyDuplication = true;
    else if (matchStart <= interestingWindow) hasNearbyDuplication = true;
    else if (matchEnd >= originalReference->length() - interestingWindow) hasNearbyDuplication = true;
    if (hasNearbyDuplication) return false;
    if (best->hasAmbiguous()) return false;
    return true;
  }

  QueryAlns unaligned() { QueryAlns r; r.comps.emplace_back(); return r; }
  QueryAlns single(const std::vector<QueryAlnP>& choices) { QueryAlns r; r.comps.push_back(choices); return r; }

  // alignToAncestralReference :306-484. queryRC[i] = reverse complement of query.seqs[i] (owned by caller)
  QueryAlns align(const Query& query, const std::vector<Seq*>& queryRC) {
    double maxInterestingPenalty = query.length() * parameters.MaxErrorRate;
    int maxInnerDistance = j2i(maxInterestingPenalty * query.spacingDeviationPerUnitPenalty + query.expectedInnerDistance);
    std::vector<std::unique_ptr<Pyramid>> pyramids;
    std::vector<std::unique_ptr<CountingPath>> paths;
    std::vector<CountingPath*> pathPtrs;
    for (size_t i = 0; i < query.seqs.size(); i++) {
      const Seq* qs = query.seqs[i]; const Seq* rcq = queryRC[i];
      if (i > 0) std::swap(qs, rcq);  // mate 2 is reverse-complemented before seeding (:317-318)
      pyramids.push_back(std::make_unique<Pyramid>(qs));
      paths.push_back(std::make_unique<CountingPath>(pyramids.back().get(), index, seqdb, qs, rcq, parameters));
      paths.back()->stats = &stats; paths.back()->path.stats = &stats;
      pathPtrs.push_back(paths.back().get());
    }
    PathsCounter path(pathPtrs, j2i(query.expectedInnerDistance), maxInnerDistance);
    QueryAlnP optimisticBestAlignment;
    bool haveOptimisticMatch = false; QueryMatch optimisticBestMatch;
    int numMismatches = 0;
    QMList bestMatches = path.optimisticGetBestMatches();
    QueryMatchAligner aligner(&query, parameters, &stats);
    if (bestMatches->size() == 1) {
      optimisticBestMatch = (*bestMatches)[0]; haveOptimisticMatch = true;
      optimisticBestAlignment = aligner.align(optimisticBestMatch, 0);
      if (quicklyConfident(optimisticBestAlignment, optimisticBestMatch)) { numImmediateAccepts++; return single({optimisticBestAlignment}); }
    }
    if (optimisticBestAlignment != nullptr) {
      while (true) {
        double possiblePenalty = getPenaltyLowerBound(numMismatches);
        if (possiblePenalty > optimisticBestAlignment->totalPenalty + parameters.Max_PenaltySpan) { numImmediateAccepts++; return single({optimisticBestAlignment}); }
        QMList matches = path.findGoodPositionsHavingPriority(numMismatches);
        numMismatches++;
        bool done = false;
        for (auto& m : *matches) if (!optimisticBestMatch.samePosition(m)) { done = true; break; }
        if (done) break;
      }
    }
    double bestPenalty = (double)JMAX;
    int candidateNumMismatches = 0;
    while (true) {
      double estimatedPenalty = getPenaltyLowerBound(candidateNumMismatches);
      if (estimatedPenalty > bestPenalty + parameters.Max_PenaltySpan) break;
      if (candidateNumMismatches > path.getNumBlocks()) break;
      QMList candidates = path.findGoodPositionsHavingPriority(candidateNumMismatches);
      for (auto& m : *candidates) {
        QueryAlnP aln;
        if (haveOptimisticMatch && m.samePosition(optimisticBestMatch)) aln = optimisticBestAlignment;
        else aln = aligner.align(m, 0);
        if (aln != nullptr) { if (bestPenalty > aln->totalPenalty) bestPenalty = aln->totalPenalty; }
      }
      if (estimatedPenalty >= maxInterestingPenalty) break;
      candidateNumMismatches++;
    }
    if (aligner.getBestAlignments().size() < 1 && query.seqs.size() > 1) {
      QMList partial = path.findPartiallyGoodPositions();
      for (auto& m : *partial) {
        QueryAlnP aln = aligner.align(m, 0);
        if (aln != nullptr) { if (bestPenalty > aln->totalPenalty) bestPenalty = aln->totalPenalty; }
      }
    }
    std::vector<QueryAlnP> bestAlignments = aligner.getBestAlignments();
    QueryAlns result = single(bestAlignments);
    if (bestAlignments.size() < 1 && query.seqs.size() > 1) result = getUnpairedAlignments(query, path);
    if ((long long)bestAlignments.size() > (long long)parameters.MaxNumMatches) return unaligned();
    return result;
  }

  QueryAlns getUnpairedAlignments(const Query& query, PathsCounter& path) {  // :602-644
    QueryAlns out;
    out.comps.resize(2);
    double expectedInner = query.expectedInnerDistance;
    for (int si = 0; si < (int)query.seqs.size(); si++) {
      const Seq* sequence = query.seqs[(size_t)si];
      double maxSub = sequence->length() * parameters.MaxErrorRate;
      int maxNumMutations = j2i(maxSub / parameters.MutationPenalty);
      std::vector<SeqMatch> candidates = path.findGoodComponentMatches(si, maxNumMutations);
      Query subQuery;
      subQuery.seqs.push_back(query.seqs[(size_t)si]);
      subQuery.expectedInnerDistance = query.expectedInnerDistance;
      subQuery.spacingDeviationPerUnitPenalty = query.spacingDeviationPerUnitPenalty;
      QueryMatchAligner subAligner(&subQuery, parameters, &stats);
      for (auto& sm : candidates) {
        int minInner;
        if (si % 2 == 1) minInner = sm.startB(); else minInner = sm.b->length() - sm.endB();
        double inner = minInner;
        if (inner < expectedInner) inner = expectedInner;
        double spacingPenalty = inner / query.spacingDeviationPerUnitPenalty;
        if (spacingPenalty > maxSub) continue;
        QueryMatch qm; qm.comps.push_back(sm); qm.priority = -1; qm.hintForward = false;
        subAligner.align(qm, inner);
      }
      out.comps[(size_t)si] = subAligner.getBestAlignments();
      // keep joined sequences alive (none are created for single-sequence sub-queries)
    }
    return out;
  }
};

}  // namespace xo
