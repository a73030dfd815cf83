// ORACLE — TEST INFRASTRUCTURE ONLY (see xo_core.h header).
// Reference side: sequence database, hash-block index (one bucket table per numBasepairsUsed),
// duplication detector.
// Follows QV/SequenceDatabase.java, M/HashBlock_Database.java, M/HashBlock_Buffer.java,
// M/PackedMap.java, QV/ByteKeyStore.java (semantics only: sorted bucket, overfull bucket = unknown;
// the delta codec is storage), M/Readable_HashBlock_Database.java, M/DuplicationDetector.java,
// M/Readable_DuplicationDetector.java.
//
// Stated deviation: ByteKeyStore marks a bucket overfull when its pending+encoded
// BYTE count exceeds maxCount*bytesPerPosition; for delta-compressed stores (maxCount*bytes >= 800)
// the encoded size can differ from the raw size by a few bytes and the outcome depends on add order
// in the reference itself.  Here a bucket is overfull iff its final entry count > maxCount.
#pragma once
#include <memory>
#include "xo_hash.h"
#include <set>
#include <thread>
#include <mutex>
#include <atomic>

namespace xo {

struct SeqPos { Seq* seq; int start; };

struct SeqDb {  // QV/SequenceDatabase.java
  std::vector<std::unique_ptr<Seq>> owned;
  std::vector<Seq*> seqs;  // [fwd0, rc0, fwd1, rc1, ...] when built by addWithRC
  std::vector<long long> starts;
  long long totalForward = 0, totalFR = 0;
  int numBitsPerPosition = 8;

  Seq* addWithRC(std::unique_ptr<Seq> f) {
    Seq* fp = f.get();
    add(std::move(f));
    add(makeRC(fp));
    return fp;
  }
  void add(std::unique_ptr<Seq> s) {  // :44-51
    s->id = (long long)seqs.size();
    if (s->complementedFrom == nullptr) totalForward += s->length();
    totalFR += s->length();
    seqs.push_back(s.get());
    owned.push_back(std::move(s));
  }
  static int log2RoundUp(long long value) {  // :76-86
    int numBits = 1; long long e = 2;
    while (true) { if (e >= value) return numBits; numBits++; e *= 2; }
  }
  void computeMetrics() {  // :65-74, :240-248
    numBitsPerPosition = std::max(log2RoundUp(totalFR), 8);
    starts.resize(seqs.size());
    long long t = 0;
    for (auto s : seqs) { starts[(size_t)s->id] = t; t += s->length(); }
  }
  long long encodePosition(const Seq* s, int start) const { return starts[(size_t)s->id] + start; }
  SeqPos decodePosition(long long enc) const {  // :170-209 (result: last sequence whose start <= enc)
    size_t idx = (size_t)(std::upper_bound(starts.begin(), starts.end(), enc) - starts.begin()) - 1;
    return SeqPos{seqs[idx], (int)(enc - starts[idx])};
  }
  Seq* reverseComplementOf(const Seq* s) const {  // :277-288
    if (s->complementedFrom) return s->complementedFrom;
    int id = (int)s->id;
    return seqs[(size_t)(id + 1 - (id % 2) * 2)];
  }
  std::vector<Seq*> forwardOnly() const {
    std::vector<Seq*> r;
    for (auto s : seqs) if (!s->complementedFrom) r.push_back(s);
    return r;
  }
};

// One PackedMap (M/PackedMap.java) in CSR form.
struct LenTable {
  int used = 0;
  int capacity = 1;
  int maxCount = 1;
  std::vector<long long> offsets;       // capacity+1
  std::vector<uint32_t> positions;      // global positions, ascending per bucket
  std::vector<uint8_t> overfull;        // per bucket
  bool empty() const { return positions.empty() && offsets.empty(); }
  int bucketOf(int32_t key) const {     // getPackedKey :197-202
    int r = key % capacity;
    if (r < 0) r += capacity;
    return r;
  }
  int numMatchesLowerBound(int32_t key) const {  // :210-218
    if (offsets.empty()) return 0;
    int b = bucketOf(key);
    if (overfull[b]) return JMAX;
    return (int)(offsets[b + 1] - offsets[b]);
  }
};

struct Index {  // M/HashBlock_Database.java (+ Readable view)
  SeqDb* db;
  bool enableGapmers = true;
  int minInterestingSize = 1;
  int maxNumShortMatches = 5;
  int maxBuilt = 0;           // maxFullySetUpSize
  int numThreads = 1;
  std::vector<LenTable> tables;  // index = numBasepairsUsed
  std::mutex growMutex;

  Index(SeqDb* d, int minInteresting = -1, int maxShort = -1, bool gapmers = true) : db(d), enableGapmers(gapmers) {
    // :51-55  (int)Math.max(log(N+1)/log(4) - 2, 1): max first, then cast
    if (minInteresting <= 0) minInterestingSize = j2i(std::max((std::log((double)(db->totalForward + 1)) / std::log(4.0)) - 2, 1.0));
    else minInterestingSize = minInteresting;
    maxNumShortMatches = maxShort < 0 ? 5 : maxShort;  // :68-89
  }
  static int chooseMinDuplicationLength(const SeqDb* d) { return SeqDb::log2RoundUp(d->totalForward); }  // M/DuplicationDetector.java:17-31
  static int chooseMaxDuplicationLength(const SeqDb* d) { return chooseMinDuplicationLength(d) * 2; }

  int estimateRequiredCapacity(int n) const {  // :620-665
    int anchorBlockSize = enableGapmers ? n * 2 / 3 : n;
    double sizeProbability = std::min(1.0, 2.0 / anchorBlockSize);
    double offsetProbability = std::min(1.0, 2.0 / anchorBlockSize);
    double blockPossibilityProbability = sizeProbability * offsetProbability;
    long long maxNumSequencesOfThisLength = (n <= 16) ? (1LL << (n * 2)) : (1LL << 32);
    long long maxNumStored = maxNumSequencesOfThisLength / 2;
    long long maxNumExistentHashcodes = j2l((double)maxNumStored * blockPossibilityProbability);
    long long effectiveSize = db->totalForward;
    long long numBlocksOfThisSize = j2l((double)effectiveSize * blockPossibilityProbability);
    double existenceFraction = 1 - std::pow((double)((double)maxNumExistentHashcodes - 1.0) / (double)maxNumExistentHashcodes, (double)numBlocksOfThisSize);
    int uniqueCount = j2i((double)maxNumExistentHashcodes * existenceFraction);
    int result = uniqueCount;
    if (result % 2 == 0) result++;
    return result;
  }
  int maxCountFor(int n) const {  // :569-576
    int m = n * n;
    if (m < maxNumShortMatches) m = maxNumShortMatches;
    if (m > 32766) m = 32766;
    if (m < 1) m = 1;
    return m;
  }

  struct Entry { uint32_t bucket; uint32_t pos; };

  // Builds every table for numBasepairsUsed in (maxBuilt, through].
  void buildThrough(int through) {
    std::lock_guard<std::mutex> lock(growMutex);
    if (through <= maxBuilt) return;
    int lo = maxBuilt + 1, hi = through;
    if ((int)tables.size() <= hi) tables.resize((size_t)hi + 1);
    std::vector<int> cap((size_t)hi + 1, 1);
    for (int n = lo; n <= hi; n++) { int c = estimateRequiredCapacity(n); if (c < 1) c = 1; cap[n] = c; }
    // jobs: 50 kbp slices (:218-235)
    struct Job { Seq* seq; int s, e; };
    std::vector<Job> jobs;
    for (Seq* s : db->forwardOnly()) {
      int target = 50000;
      int nj = (s->length() + target - 1) / target;
      int prevStart = 0;
      for (int i = 1; i <= nj; i++) {
        int st = (int)((long long)s->length() * (long long)i / (long long)nj);
        jobs.push_back({s, prevStart, st});
        prevStart = st;
      }
    }
    int nt = std::max(1, numThreads);
    std::vector<std::vector<std::vector<Entry>>> single((size_t)nt), multi((size_t)nt);
    for (int t = 0; t < nt; t++) { single[t].resize((size_t)hi + 1); multi[t].resize((size_t)hi + 1); }
    std::atomic<size_t> next(0);
    auto work = [&](int t) {
      while (true) {
        size_t j = next.fetch_add(1);
        if (j >= jobs.size()) break;
        hashJob(jobs[j].seq, jobs[j].s, jobs[j].e, lo, hi, cap, single[t], multi[t]);
      }
    };
    if (nt == 1) work(0);
    else { std::vector<std::thread> th; for (int t = 0; t < nt; t++) th.emplace_back(work, t); for (auto& x : th) x.join(); }
    // finalize each length
    auto fin = [&](int n) {
      LenTable& T = tables[(size_t)n];
      T.used = n;
      size_t total = 0, totalMulti = 0;
      for (int t = 0; t < nt; t++) { total += single[t][n].size(); totalMulti += multi[t][n].size(); }
      if (total + totalMulti == 0) { T.capacity = 1; T.maxCount = 1; return; }  // :383-389 PackedMap(1,1,..)
      T.capacity = cap[n]; T.maxCount = maxCountFor(n);
      std::vector<Entry> all; all.reserve(total + totalMulti);
      for (int t = 0; t < nt; t++) all.insert(all.end(), single[t][n].begin(), single[t][n].end());
      if (totalMulti) {
        // PackedMap.add(preventDuplicates): a multi-block possibility already present in its bucket is skipped (:117-131)
        std::set<std::pair<uint32_t, uint32_t>> seen;
        for (auto& e : all) seen.insert({e.bucket, e.pos});
        for (int t = 0; t < nt; t++) for (auto& e : multi[t][n]) if (seen.insert({e.bucket, e.pos}).second) all.push_back(e);
      }
      std::sort(all.begin(), all.end(), [](const Entry& a, const Entry& b) { return a.bucket != b.bucket ? a.bucket < b.bucket : a.pos < b.pos; });
      T.offsets.assign((size_t)T.capacity + 1, 0);
      T.overfull.assign((size_t)T.capacity, 0);
      std::vector<long long> cnt((size_t)T.capacity, 0);
      for (auto& e : all) cnt[e.bucket]++;
      size_t kept = 0;
      for (int b = 0; b < T.capacity; b++) { if (cnt[b] > T.maxCount) T.overfull[b] = 1; else kept += (size_t)cnt[b]; }
      T.positions.reserve(kept);
      long long off = 0;
      size_t i = 0;
      for (int b = 0; b < T.capacity; b++) {
        T.offsets[b] = off;
        if (!T.overfull[b]) { for (long long k = 0; k < cnt[b]; k++) T.positions.push_back(all[i + (size_t)k].pos); off += cnt[b]; }
        i += (size_t)cnt[b];
      }
      T.offsets[(size_t)T.capacity] = off;
    };
    if (nt == 1) { for (int n = lo; n <= hi; n++) fin(n); }
    else {
      std::atomic<int> nn(lo);
      std::vector<std::thread> th;
      for (int t = 0; t < nt; t++) th.emplace_back([&]() { while (true) { int n = nn.fetch_add(1); if (n > hi) break; fin(n); } });
      for (auto& x : th) x.join();
    }
    maxBuilt = hi;
  }

  // hashSequenceThroughSize (:490-528) + addHashblocks/addToBlocksBySize (:530-618) + PackedMap.process (:99-121).
  // Enumerates every block of every level whose start lies in [s, e); the Java traversal creates exactly
  // those (plus longer ones that addToBlocksBySize discards).
  void hashJob(Seq* seq, int s, int e, int lo, int hi, const std::vector<int>& cap,
               std::vector<std::vector<Entry>>& single, std::vector<std::vector<Entry>>& multi) {
    Pyramid pyr(seq, s - 1);
    Seq* rc = db->reverseComplementOf(seq);
    auto addOne = [&](const HB& blk, bool isMulti) {
      HB g;
      if (enableGapmers) { if (!withGapAndExtension(blk, seq, g)) return; } else g = blk;
      int n = g.used;
      if (n < minInterestingSize) return;
      if (n < lo || n > hi) return;
      auto& dst = isMulti ? multi[n] : single[n];
      int c = cap[n];
      if (g.primary()) {
        int r = g.fwd % c; if (r < 0) r += c;
        dst.push_back({(uint32_t)r, (uint32_t)db->encodePosition(seq, g.start)});
      }
      if (g.secondary()) {
        int r = g.rev % c; if (r < 0) r += c;
        dst.push_back({(uint32_t)r, (uint32_t)db->encodePosition(rc, rc->length() - g.end())});
      }
    };
    for (int level = 0;; level++) {
      Row* row = pyr.get(level);
      bool any = false, anyShort = false;
      auto visit = [&](const MB* b) {
        any = true;
        if (b->minLength() <= hi) anyShort = true;
        if (b->single) addOne(b->hb, false);
        else for (auto& p : b->poss) if (p.has) addOne(p.hb, true);
      };
      if (level == 0) {
        for (int i = s; i < e; i++) visit(row->get(i));
      } else {
        // force the row through e, then walk its block list in order (getAfter() scans backwards from the end)
        ParentRow* pr = static_cast<ParentRow*>(row);
        pr->getAfter(e - 1);
        for (const MB& b : pr->blockList) { int st = b.startIndex(); if (st >= e) break; if (st >= s) visit(&b); }
      }
      if (!any || !anyShort) break;
    }
  }

  // ---- Readable_HashBlock_Database ----
  const LenTable& tableFor(int n) {  // getContainingMap :97-106 + lazy growth :108-113, M/HashBlock_Database.java:183-193
    if (n > maxBuilt) {
      int target = (maxBuilt < 1) ? std::max(chooseMaxDuplicationLength(db), n) : n * 2;
      buildThrough(target);
    }
    return tables[(size_t)n];
  }
  void prepare() { tableFor(1); }
  int numMatchesLowerBound(const HB& b) {  // :72-80
    if (b.used < minInterestingSize) return JMAX;
    return tableFor(b.used).numMatchesLowerBound(b.lookupKey());
  }
  int maxNumMatchesAllowed(const HB& b) {  // :82-90
    if (b.used < minInterestingSize) return -1;
    return tableFor(b.used).maxCount;
  }
  // matchBlock :22-38. returns false for "null" (unknown); hits appended otherwise
  bool matchBlock(const HB& b, std::vector<SeqPos>& out) {
    out.clear();
    if (b.used < minInterestingSize) return false;
    const LenTable& T = tableFor(b.used);
    int32_t key = b.lookupKey();
    bool invert = !b.primary();
    int cnt = T.numMatchesLowerBound(key);
    if (cnt > T.maxCount) return false;  // PackedMap.get :162-165 (overfull -> MAX_VALUE)
    if (T.offsets.empty()) return true;
    int bk = T.bucketOf(key);
    for (long long i = T.offsets[bk]; i < T.offsets[bk + 1]; i++) {
      SeqPos p = db->decodePosition((long long)T.positions[(size_t)i]);
      if (invert) {  // :52-57
        Seq* r = db->reverseComplementOf(p.seq);
        p = SeqPos{r, r->length() - p.start - b.len};
      }
      out.push_back(p);
    }
    return true;
  }
};

// M/DuplicationDetector.java + M/Readable_DuplicationDetector.java: only the keys (duplication starts) are
// observable by the aligner (mayContainDuplicationInRange), values are kept for compareDuplications.
struct DupDetector {
  Index* index;
  int minSize, maxSize, minCopies, windowSize;
  bool enableGapmers;
  struct DupGroup { int length; std::vector<SeqPos> starts; };  // M/Duplication.java: one object shared by all of its positions
  struct Dup { int length; int numInstances; std::shared_ptr<DupGroup> group; };
  std::map<const Seq*, std::map<int, Dup>> bySeq;
  bool detected = false;
  std::mutex mu;

  DupDetector(Index* ix, int minLen, int maxLen, int minNumCopies, int window)
      : index(ix), minSize(minLen), maxSize(maxLen), minCopies(minNumCopies), windowSize(window), enableGapmers(ix->enableGapmers) {}
  double detectionGranularity() const { return enableGapmers ? (double)(minSize * 5 / 8) : (double)minSize; }  // :67-77
  int windowNumber(int i) const { return i / windowSize; }

  int compareDuplications(int start1, const Dup& d1, int start2, const Dup& d2) const {  // :406-436
    if (windowSize > 1) { if (windowNumber(start1) != windowNumber(start2)) return 0; }
    int end1 = start1 + d1.length, end2 = start2 + d2.length;
    if (start1 <= start2 && end1 >= end2) return 1;
    if (start1 >= start2 && end1 <= end2) return -1;
    if (windowSize > 1) {
      int cd = d1.numInstances - d2.numInstances;
      if (cd != 0) return cd;
      if (start1 != start2) return start1 - start2;
    }
    return 0;
  }
  void saveDuplications(std::map<const Seq*, std::map<int, Dup>>& blocks) {  // :332-400
    for (auto& entry : blocks) {
      auto& all = bySeq[entry.first];
      for (auto& pos : entry.second) {
        int dupStart = pos.first;
        const Dup& nd = pos.second;
        bool insert = true;
        while (true) {
          auto it = all.upper_bound(dupStart);  // floorEntry
          if (it != all.begin()) {
            --it;
            int c = compareDuplications(dupStart, nd, it->first, it->second);
            if (c > 0) { insert = false; break; }
            if (c < 0) { all.erase(it); continue; }
          }
          break;
        }
        while (true) {
          auto it = all.lower_bound(dupStart);  // ceilingEntry
          if (it != all.end()) {
            int c = compareDuplications(dupStart, nd, it->first, it->second);
            if (c > 0) { insert = false; break; }
            if (c < 0) { all.erase(it); continue; }
          }
          break;
        }
        if (insert) all[dupStart] = nd;
      }
    }
  }
  void detect() {  // :97-104, process :129-214
    std::lock_guard<std::mutex> lock(mu);
    if (detected) return;
    SeqDb* db = index->db;
    index->tableFor(minSize + 1);  // ensureHashed :121
    for (int blockLength = minSize; blockLength <= maxSize; blockLength++) {
      const LenTable& T = index->tableFor(blockLength);
      int numBlocks = T.capacity;  // getNumHashKeys
      std::map<const Seq*, std::map<int, Dup>> blocks;
      for (int hashcode = 0; hashcode < numBlocks; hashcode++) {
        // lookupByForwardHash :41-52
        bool known = !(T.numMatchesLowerBound(hashcode) > T.maxCount);
        if (known) {
          std::vector<SeqPos> matches;
          if (!T.offsets.empty()) {
            int bk = T.bucketOf(hashcode);
            for (long long i = T.offsets[bk]; i < T.offsets[bk + 1]; i++) matches.push_back(db->decodePosition((long long)T.positions[(size_t)i]));
          }
          size_t nf = matches.size();
          for (size_t i = 0; i < nf; i++) {
            Seq* r = db->reverseComplementOf(matches[i].seq);
            matches.push_back(SeqPos{r, r->length() - matches[i].start - blockLength});
          }
          int numForwardMatches = (int)matches.size() / 2;
          if (numForwardMatches >= minCopies) {
            std::map<std::string, std::vector<SeqPos>> byText;
            for (auto& p : matches) {
              int prefixLength = (blockLength + 3) / 4;
              std::string text = p.seq->range(p.start, prefixLength) + p.seq->range(p.start + blockLength - prefixLength, prefixLength);
              bool amb = false;
              for (char c : text) if (c != 'A' && c != 'C' && c != 'G' && c != 'T' && c != '-') amb = true;
              if (!amb) byText[text].push_back(p);
            }
            for (auto& g : byText) {
              // removeDuplicatePositions: unique (sequence,start)
              std::set<std::pair<const Seq*, int>> uniq;
              for (auto& p : g.second) uniq.insert({p.seq, p.start});
              Dup d{blockLength, (int)uniq.size(), std::make_shared<DupGroup>()};
              d.group->length = blockLength;
              for (auto& u : uniq) d.group->starts.push_back(SeqPos{const_cast<Seq*>(u.first), u.second});
              if (d.numInstances >= minCopies) for (auto& u : uniq) blocks[u.first][u.second] = d;  // groupDuplicationsBySequence :252-269
            }
          }
        }
        if (hashcode % 10000 == 9999 || hashcode == numBlocks - 1) { saveDuplications(blocks); blocks.clear(); }
      }
    }
    detected = true;
  }
  // Readable_DuplicationDetector.mayContainDuplicationInRange :28-47. returns false for null
  bool mayContainDuplicationInRange(const Seq* sequence, int startIndex, int endIndex, int& key) {
    if (!detected) detect();
    int windowStart = windowNumber(startIndex), windowEnd = windowNumber(endIndex);
    auto sit = bySeq.find(sequence);
    if (sit == bySeq.end()) return false;
    auto& m = sit->second;
    auto it = m.upper_bound(endIndex);
    if (it != m.begin()) {
      --it;
      int w = windowNumber(it->first);
      if (w >= windowStart && w <= windowEnd) { key = it->first; return true; }
    }
    auto it2 = m.lower_bound(startIndex);
    if (it2 != m.end()) {
      int w = windowNumber(it2->first);
      if (w >= windowStart && w <= windowEnd) { key = it2->first; return true; }
    }
    return false;
  }
};

}  // namespace xo
