// ORACLE — TEST INFRASTRUCTURE ONLY (see xo_core.h header).
// Seed walk, candidate binning, mate pairing.
// Follows M/HashBlockPath.java, M/Counting_HashBlockPath.java, M/HashBlockMatch_Counter.java,
// M/HashBlockPaths_Counter.java, M/QueryMatch.java, M/SequenceMatch.java.
//
// Stated deviation: Counting_HashBlockPath iterates HashMap<Sequence,...>.values()
// (identity hash => order undefined in the reference itself) in tryEnsureGoodMatchCounter and
// getAllPositions; here contigs are visited in database order.
#pragma once
#include "xo_index.h"
#include <deque>

#ifdef XO_TRACE
#include <cstdio>
#define XO_T(...) fprintf(stderr, __VA_ARGS__)
#else
#define XO_T(...)
#endif
namespace xo {

struct SeqMatch {  // M/SequenceMatch.java
  const Seq* a = nullptr; const Seq* b = nullptr; int offset = 0;
  bool fromHashblockMatch = true;
  int startB() const { return std::max(0, offset); }
  int endB() const { return std::min(offset + a->length(), b->length()); }
  int startA() const { return startB() - offset; }
  int endA() const { return endB() - offset; }
  int length() const { return endB() - startB(); }
  bool reversed() const { return a->complementedFrom != nullptr; }
  bool equalsSM(const SeqMatch& o) const { return offset == o.offset && a == o.a && b == o.b; }
};

struct QueryMatch {  // M/QueryMatch.java
  std::vector<SeqMatch> comps;
  int priority = 0;
  bool hintForward = false;
  int queryTotalLength() const { int t = 0; for (auto& c : comps) t += c.a->length(); return t; }
  bool reversed() const { return comps[0].reversed(); }
  int startIndexB() const { return std::min(comps.front().startB(), comps.back().startB()); }
  int endIndexB() const { return std::max(comps.front().startB(), comps.back().startB()); }  // sic :57-61
  int distance(const SeqMatch& a, const SeqMatch& b) const {  // :124-133
    if (a.b != b.b) return JMAX;
    if (reversed()) return a.startB() - b.endB();
    return b.startB() - a.endB();
  }
  int totalDistanceBetweenComponents() const {
    int t = 0;
    for (size_t i = 1; i < comps.size(); i++) t = wadd(t, distance(comps[i - 1], comps[i]));
    return t;
  }
  int totalDistanceAcross() const {
    if (reversed()) return comps.front().endB() - comps.back().startB();
    return comps.back().endB() - comps.front().startB();
  }
  bool samePosition(const QueryMatch& o) const {  // :83-95 (this.reversed is never set => always false)
    if (comps.size() != o.comps.size()) return false;
    for (size_t i = 0; i < comps.size(); i++) if (!comps[i].equalsSM(o.comps[i])) return false;
    return true;
  }
};
typedef std::shared_ptr<std::vector<QueryMatch>> QMList;

struct Counter {  // M/HashBlockMatch_Counter.java
  SeqMatch match;
  const std::vector<HB>* history;
  int numMatches = 0, numDistinctMismatches = 0, lastMismatchedPosition = 0;
  bool hasLastMatched = false; long long lastMatchedIdent = 0;
  size_t historyProcessed = 0;
  bool good = false; int priority = 0;
  Counter* next = nullptr; Counter* prev = nullptr;
  void update() {
    while (historyProcessed < history->size()) { updateOne((*history)[historyProcessed]); historyProcessed++; }
  }
  void updateOne(const HB& block) {  // :83-97
    if (!(hasLastMatched && block.ident == lastMatchedIdent)) {
      if (block.start >= lastMismatchedPosition) {
        if (match.offset + block.end() <= match.b->length()) { numDistinctMismatches++; lastMismatchedPosition = block.end(); }
      }
    }
  }
  int getNumDistinctMismatches() { update(); return numDistinctMismatches; }
  void addMatch(const HB& block) { numMatches++; hasLastMatched = true; lastMatchedIdent = block.ident; }
  void setGood() { good = true; priority = getNumDistinctMismatches(); }
};
typedef std::shared_ptr<std::vector<Counter*>> CounterList;

struct OracleStats {  // per-query counters used by the roofline arithmetic
  long long probes = 0, seeds = 0, hits = 0, pathAlignerCalls = 0, pathAlignerSteps = 0, pathAlignerCells = 0;
  long long straightCalls = 0;
};

struct HashBlockPath {  // M/HashBlockPath.java
  Pyramid* pyramid; Index* database; const Seq* query;
  int batchIndex = -1;
  MB dummy; const MB* currentBlock;
  bool haveGapmer = false; HB currentGapmer;
  bool havePrev = false, havePrevPrev = false; int32_t prevFwd = 0, prevPrevFwd = 0;
  long long gapmerSerial = 0;
  OracleStats* stats = nullptr;
  HashBlockPath(Pyramid* p, Index* d, const Seq* q) : pyramid(p), database(d), query(q) {
    dummy.single = true; dummy.hb = HB(); dummy.hb.ident = -1;  // new HashBlock(0, 0) :20
    currentBlock = &dummy;
  }
  bool getNextInterestingBlock(HB& out) {  // :27-50 (previousBlock is never assigned: §9-6)
    if (currentBlock == nullptr) return false;
    while (true) {
      if (!getNextBlockWithGoodNumberOfMatches(out)) return false;
      if (recentlySeen(out)) continue;
      break;
    }
    return true;
  }
  bool recentlySeen(const HB& block) {  // :52-65
    bool r = false;
    if (havePrev && block.fwd == prevFwd) r = true;
    else if (havePrevPrev && block.fwd == prevPrevFwd) r = true;
    havePrevPrev = havePrev; prevPrevFwd = prevFwd;
    havePrev = true; prevFwd = block.fwd;
    return r;
  }
  bool getNextBlockWithGoodNumberOfMatches(HB& out) {  // :68-96
    while (true) {
      if (!advanceToNextPosition()) return false;
      HB ext;
      if (!withGap(ext)) continue;
      if (!hasFewEnoughMatches(ext)) continue;
      out = ext;
      return true;
    }
  }
  void moveDown() {  // :99-108
    batchIndex--;
    currentBlock = pyramid->get(batchIndex)->getAfter(currentBlock->startIndex());
    haveGapmer = false;
  }
  void moveUpOrRight() {  // :111-122
    const HB& left = currentBlock->hb;
    const MB* up = pyramid->get(batchIndex + 1)->get(left.start);
    if (up != nullptr && up->startIndex() <= left.start) { batchIndex++; currentBlock = up; haveGapmer = false; }
    else moveRight();
  }
  void moveRight() {  // :125-128
    currentBlock = pyramid->get(batchIndex)->getAfter(currentBlock->startIndex());
    haveGapmer = false;
  }
  void skipMultiblocks() {  // :130-140
    while (true) {
      if (currentBlock == nullptr || currentBlock->single) return;
      if (batchIndex > 0) moveDown(); else moveRight();
    }
  }
  bool advanceToNextPosition() {  // :143-195
    const HB& single = currentBlock->hb;
    if (maxGapmerNumBasepairsUsed(single.len) < database->minInterestingSize && database->enableGapmers) {
      moveUpOrRight();
    } else {
      HB ext;
      if (withGap(ext)) {
        int numMatches = database->numMatchesLowerBound(ext);
        if (stats) stats->probes++;
        if (numMatches < 6) { if (batchIndex > 0) moveDown(); else moveRight(); }
        else if (numMatches > getMaxNumMatchesAllowed(ext)) moveUpOrRight();
        else moveRight();
      } else {
        int typical = single.len * 3 / 2;
        if (typical <= database->minInterestingSize && database->enableGapmers) moveUpOrRight();
        else { if (batchIndex > 0) moveDown(); else moveRight(); }
      }
    }
    skipMultiblocks();
    return currentBlock != nullptr;
  }
  bool withGap(HB& out) {  // :197-203
    if (!database->enableGapmers) { out = currentBlock->hb; return true; }
    if (!haveGapmer) {
      HB g;
      if (!withGapAndExtension(currentBlock->hb, query, g)) return false;
      if (currentBlock->hb.gapDir != 0) g.ident = (1LL << 60) + (gapmerSerial++);  // fresh Gapped_HashBlock object
      currentGapmer = g; haveGapmer = true;
    }
    out = currentGapmer;
    return true;
  }
  int getMaxNumMatchesAllowed(const HB& block) {  // :205-219
    if (block.len >= query->length() / 6) return database->maxNumMatchesAllowed(block);
    if (block.rmr) return 5;
    return block.used + 1;
  }
  bool hasFewEnoughMatches(const HB& block) {
    if (stats) stats->probes++;
    return database->numMatchesLowerBound(block) <= getMaxNumMatchesAllowed(block);
  }
};

struct CountingPath {  // M/Counting_HashBlockPath.java
  static const int usualNumberOfMatchesRequiredBeforeInvestigating = 1;
  HashBlockPath path;
  Pyramid* pyramid; Index* database; SeqDb* seqdb;
  const Seq* query; const Seq* rcQuery;
  // [0] = "forwardMatchCounters" (holds REVERSED matches, §9-5), [1] = "reverseMatchCounters"
  std::map<long long, std::map<int, Counter*>> counters[2];  // key: contig id
  std::deque<Counter> counterStore;
  std::vector<Counter*> goodMatchCounters;
  bool foundGoodMatchCounter = false;
  std::vector<HB> history;
  int numBlocksMatchingAnywhere = 0, numMatchCounters = 0;
  int maxNonoverlappingBlockVisited = 0, numNonoverlappingBlocksVisited = 0;
  int minNumDistinctMismatches = -1;
  bool done = false;
  int maxIndelLengthToConsider = 0;
  std::deque<HB> pendingBlocks;
  CounterList previousHighPriority; CounterList previousAllPositions;
  OracleStats* stats = nullptr;

  CountingPath(Pyramid* p, Index* d, SeqDb* sdb, const Seq* q, const Seq* rcq, const Params& prm)
      : path(p, d, q), pyramid(p), database(d), seqdb(sdb), query(q), rcQuery(rcq) {
    int maxPossibleIndel = j2i((q->length() * prm.MaxErrorRate - prm.DeletionStart_Penalty) / prm.DeletionExtension_Penalty);  // :33
    maxIndelLengthToConsider = maxPossibleIndel / 2;
  }

  bool step() {  // :40-179
    if (done) return false;
    HB queryBlock; std::vector<SeqPos> matches;
    if (!getNextInterestingMatch(queryBlock, matches)) {
      done = true;
      if (numBlocksMatchingAnywhere < usualNumberOfMatchesRequiredBeforeInvestigating) tryEnsureGoodMatchCounter();
      return false;
    }
    history.push_back(queryBlock);
    XO_T("seed start=%d len=%d used=%d fwd=%d rev=%d count=%d invert=%d\n", queryBlock.start, queryBlock.len, queryBlock.used, queryBlock.fwd, queryBlock.rev, (int)matches.size(), (int)!queryBlock.primary());
    if (stats) { stats->seeds++; stats->hits += (long long)matches.size(); }
    int queryBlockNumMatches = (int)matches.size();
    for (auto& ref : matches) {
      const Seq* cms = ref.seq;
      int numMismatched = 0, numMatched = 0;
      for (int distance = 1; distance < 20; distance++) {
        int checkOffset = -distance;
        int qi = queryBlock.start + checkOffset;
        if (qi >= 0 && qi < query->length()) {
          int ri = ref.start + checkOffset;
          if (ri >= 0 && ri < cms->length()) {
            if (!bp_canMatch(query->at(qi), cms->at(ri))) numMismatched++; else numMatched++;
          }
        }
        checkOffset = queryBlock.len - 1 + distance;
        qi = queryBlock.start + checkOffset;
        if (qi >= 0 && qi < query->length()) {
          int ri = ref.start + checkOffset;
          if (ri >= 0 && ri < cms->length()) {
            if (!bp_canMatch(query->at(qi), cms->at(ri))) numMismatched++; else numMatched++;
          }
        }
        if (numMatched < numMismatched) break;
        if (numMatched >= numMismatched + queryBlock.used) break;
      }
      XO_T("  hit seq=%d rstart=%d mism=%d mat=%d\n", (int)ref.seq->id, ref.start, numMismatched, numMatched);
      if (numMismatched > numMatched) continue;
      SeqMatch full;
      if (cms->complementedFrom != nullptr) {
        const Seq* forwardRef = cms->complementedFrom;
        int reverseQueryBlockStart = query->length() - queryBlock.end();
        int reverseReferenceBlockStart = cms->length() - (ref.start + queryBlock.len);
        full = SeqMatch{rcQuery, forwardRef, reverseReferenceBlockStart - reverseQueryBlockStart, true};
      } else {
        full = SeqMatch{query, cms, ref.start - queryBlock.start, true};
      }
      updateMatches(full, queryBlock, queryBlockNumMatches);
    }
    if (queryBlock.start >= maxNonoverlappingBlockVisited) {
      maxNonoverlappingBlockVisited = queryBlock.end();
      numNonoverlappingBlocksVisited++;
    }
    numBlocksMatchingAnywhere++;
    minNumDistinctMismatches = -1;
    return true;
  }

  void updateMatches(const SeqMatch& sm, const HB& queryBlock, int queryBlockNumMatches) {  // :193-252
    int offset = sm.offset;
    auto& all = sm.reversed() ? counters[0] : counters[1];
    auto& onSeq = all[sm.b->id];
    Counter* cur = nullptr;
    auto it = onSeq.find(offset);
    if (it != onSeq.end()) cur = it->second;
    if (cur == nullptr) {
      counterStore.emplace_back();
      cur = &counterStore.back();
      cur->match = sm; cur->history = &history;
      cur->numDistinctMismatches = numNonoverlappingBlocksVisited;
      cur->lastMismatchedPosition = queryBlock.start;
      cur->historyProcessed = history.size() - 1;
      onSeq[offset] = cur;
      numMatchCounters++;
      auto lo = onSeq.lower_bound(offset);  // == the new entry
      if (lo != onSeq.begin()) {
        auto pv = std::prev(lo);
        if (std::abs(pv->first - offset) <= maxIndelLengthToConsider) { cur->prev = pv->second; pv->second->next = cur; }
      }
      auto nx = std::next(lo);
      if (nx != onSeq.end()) {
        if (std::abs(nx->first - offset) <= maxIndelLengthToConsider) { cur->next = nx->second; nx->second->prev = cur; }
      }
    }
    Counter* previousCounter = cur->prev;
    if (previousCounter) addMatch(sm, queryBlock, previousCounter, queryBlockNumMatches);
    Counter* nextCounter = cur->next;
    if (nextCounter) addMatch(sm, queryBlock, nextCounter, queryBlockNumMatches);
    bool updateThisOne = true;
    if ((previousCounter && previousCounter->good) || (nextCounter && nextCounter->good)) {
      if (!cur->good) updateThisOne = false;
    }
    if (updateThisOne) addMatch(sm, queryBlock, cur, queryBlockNumMatches);
  }

  void addMatch(const SeqMatch& full, const HB& queryBlock, Counter* counter, int queryBlockNumMatches) {  // :254-279
    counter->addMatch(queryBlock);
    counter->update();
    if (counter->numMatches <= usualNumberOfMatchesRequiredBeforeInvestigating) {
      if (counter->numMatches == usualNumberOfMatchesRequiredBeforeInvestigating) {
        foundGoodMatchCounter = true;
        declareGood(counter);
      } else {
        if (queryBlockNumMatches <= queryBlock.len) {
          int fromStart = full.offset;
          int fromEnd = full.b->length() - (full.offset + full.a->length());
          if (std::min(fromStart, fromEnd) < 0) declareGood(counter);
        }
      }
    }
  }
  void declareGood(Counter* c) { if (!c->good) { goodMatchCounters.push_back(c); c->setGood(); } }

  void tryEnsureGoodMatchCounter() {  // :292-308
    if (!foundGoodMatchCounter && numMatchCounters <= query->length()) {
      for (int s = 0; s < 2; s++) for (auto& onSeq : counters[s]) for (auto& e : onSeq.second) declareGood(e.second);
      foundGoodMatchCounter = true;
    }
  }

  bool getNextInterestingBlock(HB& out) {  // :344-368
    previousAllPositions.reset();
    while (true) {
      HB block;
      if (!path.getNextInterestingBlock(block)) {
        if (pendingBlocks.empty()) return false;
        out = pendingBlocks.front(); pendingBlocks.pop_front();
        return true;
      }
      if (block.start < maxNonoverlappingBlockVisited) { pendingBlocks.push_back(block); continue; }
      out = block;
      return true;
    }
  }
  bool getNextInterestingMatch(HB& block, std::vector<SeqPos>& matches) {  // :371-388
    while (true) {
      if (!getNextInterestingBlock(block)) return false;
      if (!database->matchBlock(block, matches)) continue;
      return true;
    }
  }

  CounterList findGoodPositionsHavingPriorityUpTo(int priority) {  // :406-433
    while (true) {
      // Java int arithmetic: priority + 1 overflows for Integer.MAX_VALUE
      if (numNonoverlappingBlocksVisited >= wadd(priority, usualNumberOfMatchesRequiredBeforeInvestigating)) break;
      if (!step()) break;
    }
    if (previousHighPriority && previousHighPriority->size() == goodMatchCounters.size()) return previousHighPriority;
    auto r = std::make_shared<std::vector<Counter*>>();
    for (auto c : goodMatchCounters) if (c->priority <= priority) r->push_back(c);
    previousHighPriority = r;
    return r;
  }
  CounterList getAllPositions() {  // :435-452
    if (!previousAllPositions) {
      auto r = std::make_shared<std::vector<Counter*>>();
      for (int s = 0; s < 2; s++) for (auto& onSeq : counters[s]) for (auto& e : onSeq.second) r->push_back(e.second);
      previousAllPositions = r;
    }
    return previousAllPositions;
  }
  int getNumGoodDistinctMismatches() {  // :458-470
    if (minNumDistinctMismatches < 0) {
      int mn = numNonoverlappingBlocksVisited - 1;
      for (auto c : goodMatchCounters) { int cnt = c->getNumDistinctMismatches(); if (mn >= cnt) mn = cnt; }
      minNumDistinctMismatches = mn;
    }
    return minNumDistinctMismatches;
  }
  CounterList getBestMatches() {  // :471-493
    auto best = std::make_shared<std::vector<Counter*>>();
    if (numBlocksMatchingAnywhere < usualNumberOfMatchesRequiredBeforeInvestigating) return best;
    int mn = getNumGoodDistinctMismatches();
    for (auto c : goodMatchCounters) if (c->getNumDistinctMismatches() <= mn) best->push_back(c);
    return best;
  }
};

struct PathsCounter {  // M/HashBlockPaths_Counter.java
  std::vector<CountingPath*> components;
  int maxOffsetBetweenComponents;
  QMList previousAssembled;
  std::vector<CounterList> previousMatchComponents; bool havePrevious = false;
  bool foundNonemptyResult = false;

  PathsCounter(std::vector<CountingPath*> comps, int expectedInner, int maxInner) : components(comps) {
    maxOffsetBetweenComponents = maxInner + comps[0]->query->length();
    (void)expectedInner;
  }
  QMList findGoodPositionsHavingPriority(int k) {  // :21-24
    return filterHavingPriority(findGoodPositionsWithPriorityUpTo(k), k);
  }
  QMList findPartiallyGoodPositions() {  // :26-50
    if (components.size() != 2) return std::make_shared<std::vector<QueryMatch>>();
    if (!foundNonemptyResult) return std::make_shared<std::vector<QueryMatch>>();
    std::vector<CounterList> pieces;
    bool foundGood = false, foundBad = false;
    for (auto c : components) {
      CounterList here = c->findGoodPositionsHavingPriorityUpTo(JMAX);
      if (here->size() == 0) { foundBad = true; here = c->getAllPositions(); } else foundGood = true;
      pieces.push_back(here);
    }
    if (foundGood && foundBad) return match(pieces);
    return std::make_shared<std::vector<QueryMatch>>();
  }
  QMList findGoodPositionsWithPriorityUpTo(int k) {  // :52-82
    std::vector<CounterList> pieces;
    for (auto c : components) {
      CounterList here = c->findGoodPositionsHavingPriorityUpTo(k);
      if (here->size() >= 1) foundNonemptyResult = true;
      pieces.push_back(here);
    }
    return match(pieces);
  }
  QMList optimisticGetBestMatches() {  // :84-98
    std::vector<CounterList> pieces;
    for (auto c : components) {
      while (true) {
        CounterList best = c->getBestMatches();
        if (best->size() == 1 || !c->step()) { pieces.push_back(best); break; }
      }
    }
    return filterHavingMinPriority(match(pieces));
  }
  std::vector<SeqMatch> findGoodComponentMatches(int sequenceIndex, int maxPriority) {  // :102-106
    CounterList l = components[(size_t)sequenceIndex]->findGoodPositionsHavingPriorityUpTo(maxPriority);
    std::vector<SeqMatch> r;
    for (auto c : *l) r.push_back(c->match);
    return r;
  }
  int getNumBlocks() { int t = 0; for (auto c : components) t += c->numBlocksMatchingAnywhere; return t; }

  QMList match(const std::vector<CounterList>& comps) {  // :116-133 (list identity cache, §9-18)
    bool same = havePrevious;
    if (same) for (size_t i = 0; i < comps.size(); i++) if (previousMatchComponents[i].get() != comps[i].get()) { same = false; break; }
    if (!same) {
      previousAssembled = matchWithoutCache(comps);
      previousMatchComponents = comps; havePrevious = true;
    }
    return previousAssembled;
  }
  QMList matchWithoutCache(const std::vector<CounterList>& comps) {  // :136-246
    auto results = std::make_shared<std::vector<QueryMatch>>();
    if (comps.size() == 1) {
      for (auto c : *comps[0]) { QueryMatch q; q.comps.push_back(c->match); q.priority = c->priority; q.hintForward = false; results->push_back(q); }
      return results;
    }
    // [0] forward, [1] reverse: contig -> offset -> counter
    std::map<const Seq*, std::map<int, Counter*>> matching[2];
    std::vector<std::pair<Counter*, Counter*>> matched;
    bool lastComponentIsLargest = comps.size() <= 1 || comps[0]->size() <= comps[1]->size();
    for (int i = 0; i < (int)comps.size(); i++) {
      int componentIndex = lastComponentIsLargest ? i : 1 - i;
      for (Counter* counter : *comps[(size_t)componentIndex]) {
        const SeqMatch& m = counter->match;
        int maxReverseOffset = m.a->length() / 2;
        bool sequenceMatchReversed = m.reversed();
        bool queryMatchReversed = (sequenceMatchReversed == (componentIndex % 2 == 0));
        auto& onSeq = matching[queryMatchReversed ? 1 : 0][m.b];
        int offset = m.offset;
        if (i == 0) {
          onSeq[offset] = counter;
        } else {
          int searchStart, searchEnd;
          bool otherSequenceExpectEarlier = (queryMatchReversed == lastComponentIsLargest);
          if (otherSequenceExpectEarlier) { searchStart = offset - maxReverseOffset; searchEnd = offset + maxOffsetBetweenComponents; }
          else { searchStart = offset - maxOffsetBetweenComponents; searchEnd = offset + maxReverseOffset; }
          std::vector<Counter*> nearby;
          if (searchStart <= searchEnd) {  // TreeMap.subMap throws if fromKey > toKey; cannot happen for sane inputs
            for (auto it = onSeq.lower_bound(searchStart); it != onSeq.end() && it->first <= searchEnd; ++it) nearby.push_back(it->second);
          }
          if (queryMatchReversed && nearby.size() > 1) std::reverse(nearby.begin(), nearby.end());
          for (Counter* nb : nearby) {
            if (lastComponentIsLargest) matched.push_back({nb, counter}); else matched.push_back({counter, nb});
          }
        }
      }
    }
    for (auto& g : matched) {  // assembleQueryMatches :248-265
      QueryMatch q;
      q.comps.push_back(g.first->match); q.comps.push_back(g.second->match);
      q.hintForward = g.first->getNumDistinctMismatches() < g.second->getNumDistinctMismatches();
      q.priority = countPriority(g.first, g.second);
      results->push_back(q);
    }
    return results;
  }
  static int countPriority(Counter* c1, Counter* c2) {  // :314-334
    const SeqMatch& m1 = c1->match; const SeqMatch& m2 = c2->match;
    if (m1.startB() < m2.endB() && m1.endB() > m2.startB()) return std::max(std::max(0, c1->priority), c2->priority);
    return c1->priority + c2->priority;
  }
  QMList filterHavingPriority(QMList matches, int k) {  // :267-294
    auto r = std::make_shared<std::vector<QueryMatch>>();
    for (auto& m : *matches) if (m.priority == k) r->push_back(m);
    return r;
  }
  QMList filterHavingMinPriority(QMList matches) {  // :296-304 (selects the MAX priority, §9-5)
    int mn = -1;
    for (auto& m : *matches) if (mn < 0 || mn < m.priority) mn = m.priority;
    return filterHavingPriority(matches, mn);
  }
};

}  // namespace xo
