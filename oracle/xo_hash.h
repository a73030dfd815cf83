// ORACLE — TEST INFRASTRUCTURE ONLY (see xo_core.h header).
// Seed generation: content-defined hash blocks, gapmers, lazy rows.
// Follows M/HashBlock.java, M/Gapped_HashBlock.java, M/HashBlock_BaseRow.java,
// M/HashBlock_ParentRow.java, M/HashBlock_Stream.java, M/HashBlock_Pyramid.java,
// M/MultiHashBlock.java, M/ConditionalHashBlock.java, M/SequenceCondition.java.
// HashBlock_Compiler* is a process-wide memo of ParentRow results and is not restated
// (SURVEY.md §9-15).
#pragma once
#include "xo_core.h"
#include <deque>
#include <unordered_map>

namespace xo {

struct HB {  // M/HashBlock.java fields :385-397
  int start = 0, len = 0, used = 0;
  int32_t fwd = 0, rev = 0;
  int gapDir = 0, extra = 0;
  bool rml = false, rmr = false, nrml = false, nrmr = false;
  long long ident = 0;  // stands in for Java object identity (HashBlockMatch_Counter.update uses !=)
  int end() const { return start + len; }
  bool primary() const { return (rml != rmr) ? rml : (fwd >= rev); }   // :332-337
  bool secondary() const { return (rml != rmr) ? rmr : (fwd <= rev); } // :339-343
  int32_t lookupKey() const { return primary() ? fwd : rev; }          // :322-326
};

inline int maxGapmerNumBasepairsUsed(int startingLength) { return startingLength + startingLength * 9 / 8 + 1; }  // :12-14

inline int32_t mergeHash(int lLen, int32_t lH, int rLen, int32_t rH) {  // :261-269
  long long rotatedLeft = (long long)((unsigned long long)((long long)lH + 1) * (unsigned long long)(54323LL + 323LL * (long long)rLen));
  long long rotatedRight = (long long)((unsigned long long)(long long)wadd(rH, 1) * (unsigned long long)(long long)lLen);
  long long top = (long long)((unsigned long long)rotatedLeft + (unsigned long long)rotatedRight);
  return wadd((int32_t)(uint32_t)(unsigned long long)top, (int32_t)(uint32_t)(unsigned long long)(top >> 32));
}

inline HB baseBlock(uint8_t code, int index) {  // HashBlock(char,int) :60-65 + hashChar :171-188
  HB b;
  b.start = index; b.len = 1; b.used = 1;
  if (code == 1) b.fwd = 0; else if (code == 2) b.fwd = 1; else if (code == 4) b.fwd = 2; else b.fwd = 3;
  if (b.fwd / 2 == 0) b.rml = true;
  b.rmr = !b.rml;
  if (b.fwd % 2 == 0) b.nrml = true;
  b.nrmr = !b.nrml;
  b.rev = 3 - b.fwd;
  return b;
}

inline HB mergeBlocks(const HB& L, const HB& R) {  // HashBlock(seq,start,len,left,right) :20-44 + mergeHashes :192-259
  HB b;
  b.start = L.start; b.len = R.end() - L.start; b.used = b.len;
  b.fwd = mergeHash(L.len, L.fwd, R.len, R.fwd);
  b.rev = mergeHash(R.len, R.rev, L.len, L.rev);
  b.rml = b.rmr = true; b.nrml = b.nrmr = true;
  const HB* anchor = nullptr; const HB* other = nullptr;
  if (L.fwd != R.rev) {
    if (L.fwd > R.rev) { anchor = &R; other = &L; } else { anchor = &L; other = &R; }
  }
  if (anchor != nullptr) {
    if (b.fwd != b.rev) {
      bool isReverse = b.fwd < b.rev;
      bool invert = isReverse == (anchor == &R);
      bool aL = anchor->nrml, aR = anchor->nrmr;
      if (aL && aR) { if (anchor == &R) aR = false; else aL = false; }
      bool oL = other->nrml, oR = other->nrmr;
      if (oL && oR) { if (other == &R) oL = false; else oR = false; }
      b.rml = aL != invert; b.rmr = aR != invert;
      b.nrml = oL != invert; b.nrmr = oR != invert;
    }
  }
  if (L.len != R.len) {
    b.rml = (L.len > R.len); b.rmr = !b.rml;
    b.nrml = !b.rml; b.nrmr = !b.nrml;
  }
  if (b.fwd != b.rev) {
    if (b.rml && b.rmr) { b.rml = (b.fwd > b.rev); b.rmr = !b.rml; }
    if (b.nrml && b.nrmr) { b.nrml = b.rml; b.nrmr = !b.nrml; }
  }
  // gap direction :27-40
  if (b.rml != b.rmr) b.gapDir = b.rml ? 1 : -1;
  else if (L.fwd != R.rev) b.gapDir = (L.fwd > R.rev) ? 1 : -1;
  b.extra = (L.len + R.len - b.len) / 4;
  return b;
}

inline int extCharToInt(uint8_t code) {  // charToInt :152-169 applied to decoded char
  switch (code) { case 1: return 1; case 2: return 2; case 4: return 3; case 8: return 4; }
  return 0;
}

// M/HashBlock.java:67-150. Returns 0 = null, 1 = result written to out (out may equal *this when gapDir == 0)
inline bool withGapAndExtension(const HB& b, const Seq* seq, HB& out) {
  int32_t extensionHash = 0;
  int target = b.len;
  target += jabs(std::max(b.fwd, b.rev)) % 3;
  target += b.extra;
  int gapLength = b.len / 2;
  int extensionLength = target - gapLength;
  if (b.gapDir == 0) { out = b; return true; }
  HB r;
  if (b.gapDir < 0) {
    int extensionEnd = b.start - gapLength;
    int extensionStart = extensionEnd - extensionLength;
    if (extensionStart < 0) return false;
    for (int i = extensionEnd - 1; i >= extensionStart; i--) {
      extensionHash = wmul(extensionHash, 7654337);
      extensionHash = wadd(extensionHash, extCharToInt(seq->at(i)));
    }
    r.start = extensionStart; r.len = extensionLength + gapLength + b.len;
  } else {
    int extensionStart = b.end() + gapLength;
    int extensionEnd = extensionStart + extensionLength;
    if (extensionEnd > seq->length()) return false;
    for (int i = extensionStart; i < extensionEnd; i++) {
      extensionHash = wmul(extensionHash, 7654337);
      extensionHash = wadd(extensionHash, extCharToInt(bp_complement(seq->at(i))));
    }
    r.start = b.start; r.len = b.len + gapLength + extensionLength;
  }
  r.fwd = wadd(b.fwd, extensionHash);
  r.rev = wadd(b.rev, extensionHash);
  r.used = b.len + extensionLength;
  if (r.used > maxGapmerNumBasepairsUsed(b.len)) throw std::runtime_error("gapmer numBasepairsUsed larger than expected");
  out = r;
  return true;
}

// ---- ambiguity support: SequenceCondition / ConditionalHashBlock / MultiHashBlock ----
struct Cond {  // M/SequenceCondition.java
  std::vector<std::pair<int, char>> kv;  // sorted by position
  // returns false on conflict
  static bool intersect(const Cond& a, const Cond& b, Cond& out) {  // :27-106
    if (b.kv.empty()) { out = a; return true; }
    if (a.kv.empty()) { out = b; return true; }
    size_t i = 0, j = 0;
    out.kv.clear();
    while (i < a.kv.size() && j < b.kv.size()) {
      if (a.kv[i].first < b.kv[j].first) out.kv.push_back(a.kv[i++]);
      else if (b.kv[j].first < a.kv[i].first) out.kv.push_back(b.kv[j++]);
      else {
        if (a.kv[i].second != b.kv[j].second) return false;
        out.kv.push_back(a.kv[i]); i++; j++;
      }
    }
    while (i < a.kv.size()) out.kv.push_back(a.kv[i++]);
    while (j < b.kv.size()) out.kv.push_back(b.kv[j++]);
    return true;
  }
};
struct CHB { bool has = false; HB hb; Cond cond; };  // M/ConditionalHashBlock.java (has == hashBlock != null)
struct MB {  // IMultiHashBlock: either a HashBlock (single) or a MultiHashBlock
  bool single = true;
  HB hb;
  std::vector<CHB> poss;
  int startIndex() const {  // M/MultiHashBlock.java:22-33
    if (single) return hb.start;
    int mn = -1;
    for (auto& p : poss) if (p.has) { int v = p.hb.start; if (mn < 0 || mn > v) mn = v; }
    return mn;
  }
  int endIndex() const {
    if (single) return hb.end();
    int mx = -1;
    for (auto& p : poss) if (p.has) { int v = p.hb.end(); if (mx < v) mx = v; }
    return mx;
  }
  int minLength() const {
    if (single) return hb.len;
    int mn = -1;
    for (auto& p : poss) if (p.has) { int v = p.hb.len; if (mn < 0 || mn > v) mn = v; }
    return mn;
  }
};

struct Row {  // M/HashBlock_Row.java
  virtual ~Row() {}
  virtual const MB* get(int index) = 0;
  virtual const MB* getAfter(int index) = 0;
  virtual int level() const = 0;
  const Seq* seq = nullptr;
};

struct BaseRow : Row {  // M/HashBlock_BaseRow.java
  std::unordered_map<int, std::unique_ptr<MB>> blocks;
  explicit BaseRow(const Seq* s) { seq = s; }
  const MB* get(int index) override {  // :27-59
    if (index >= seq->length()) return nullptr;
    auto it = blocks.find(index);
    if (it != blocks.end()) return it->second.get();
    auto mb = std::make_unique<MB>();
    uint8_t code = seq->at(index);
    if (bp_isAmbiguous(code)) {
      mb->single = false;
      static const uint8_t opts[4] = {1, 2, 4, 8};
      for (uint8_t o : opts) {
        if (bp_canMatch(code, o)) {
          CHB c; c.has = true; c.hb = baseBlock(o, index);
          c.hb.ident = ((long long)0 << 40) | (long long)index;
          c.cond.kv.push_back({index, bp_decode(o)});
          mb->poss.push_back(c);
        }
      }
    } else {
      mb->hb = baseBlock(code, index);
      mb->hb.ident = (long long)index;  // level 0
    }
    const MB* r = mb.get();
    blocks[index] = std::move(mb);
    return r;
  }
  const MB* getAfter(int index) override { return get(index + 1); }
  int level() const override { return 0; }
};

struct ParentRow : Row {  // M/HashBlock_ParentRow.java (assumeOnlyUsedOnce = false behaviour)
  static const int maxNumCombinationsToExpand = 64;
  Row* prev;
  int maxPositionChecked = -1;
  int lvl;
  std::deque<MB> blockList;
  // startAfter: the enumeration for index building begins at startAfter (Java: skipTo(startIndex-1), :62-67)
  ParentRow(Row* previous, int startAfter = -1) : prev(previous), maxPositionChecked(startAfter) {
    seq = previous->seq; lvl = previous->level() + 1;
  }
  int level() const override { return lvl; }
  const MB* get(int index) override {  // :21-26
    const MB* next = getAfter(index - 1);
    if (next != nullptr && next->startIndex() == index) return next;
    return nullptr;
  }
  const MB* getAfter(int position) override {  // :28-60
    if (position < maxPositionChecked) {
      // Java scans backwards from the end; block starts are ascending, so a binary search finds the same block
      size_t lo = 0, hi = blockList.size();
      while (lo < hi) { size_t mid = (lo + hi) / 2; if (blockList[mid].startIndex() > position) hi = mid; else lo = mid + 1; }
      if (lo < blockList.size()) return &blockList[lo];
    }
    while (true) {
      if (maxPositionChecked >= seq->length()) break;
      if (!blockList.empty()) {
        const MB& last = blockList.back();
        if (last.startIndex() > position) return &last;
      }
      maybeMakeBlock();
    }
    return nullptr;
  }
  static bool shouldMerge(const HB& l, const HB& r) {  // :200-208
    if (l.end() < r.start) return false;
    if (l.rmr) return true;
    if (r.rml) return true;
    return false;
  }
  HB doMerge(const HB& l, const HB& r) {
    HB m = mergeBlocks(l, r);
    m.ident = ((long long)lvl << 40) | (long long)m.start;
    return m;
  }
  void maybeMakeBlock() {  // :69-127
    int afterIndex = maxPositionChecked;
    const MB* leftBlock = prev->getAfter(afterIndex);
    if (leftBlock == nullptr) { maxPositionChecked = seq->length(); return; }
    int index = leftBlock->startIndex();
    maxPositionChecked = index;
    const MB* rightBlock = prev->getAfter(index);
    if (rightBlock != nullptr) {
      if (leftBlock->single && rightBlock->single) {
        if (shouldMerge(leftBlock->hb, rightBlock->hb)) {
          MB m; m.single = true; m.hb = doMerge(leftBlock->hb, rightBlock->hb);
          blockList.push_back(m);
        }
      } else {
        std::vector<CHB> mergeOptions;
        std::vector<CHB> leftPoss;
        if (leftBlock->single) { CHB c; c.has = true; c.hb = leftBlock->hb; leftPoss.push_back(c); }
        else leftPoss = leftBlock->poss;
        for (auto& leftOption : leftPoss) {
          if (leftOption.has) expand(leftOption.hb, leftOption.cond, index, mergeOptions);
          else { CHB c; c.has = false; c.cond = leftOption.cond; mergeOptions.push_back(c); }
        }
        if (!mergeOptions.empty() && (int)mergeOptions.size() <= maxNumCombinationsToExpand) {
          bool hasNonEmpty = false;
          for (auto& c : mergeOptions) if (c.has) hasNonEmpty = true;
          if (hasNonEmpty) { MB m; m.single = false; m.poss = mergeOptions; blockList.push_back(m); }
        }
      }
    }
  }
  void expand(const HB& leftBlock, const Cond& startingCondition, int startIndex, std::vector<CHB>& results) {  // :137-191
    const MB* next = prev->getAfter(startIndex);
    if (next == nullptr) return;
    bool foundAnIntersection = false;
    std::vector<CHB> nextPoss;
    if (next->single) { CHB c; c.has = true; c.hb = next->hb; nextPoss.push_back(c); }
    else nextPoss = next->poss;
    int nextStart = next->startIndex();
    for (auto& rightOption : nextPoss) {
      Cond inter;
      if (!Cond::intersect(startingCondition, rightOption.cond, inter)) {
        if (foundAnIntersection) break;
        continue;
      }
      foundAnIntersection = true;
      if ((int)results.size() > maxNumCombinationsToExpand) return;
      if (!rightOption.has) { expand(leftBlock, inter, nextStart, results); continue; }
      CHB c; c.cond = inter;
      if (shouldMerge(leftBlock, rightOption.hb)) { c.has = true; c.hb = doMerge(leftBlock, rightOption.hb); }
      else c.has = false;
      results.push_back(c);
    }
  }
};

// M/HashBlock_Pyramid.java + M/HashBlock_Stream.java: lazy list of rows
struct Pyramid {
  const Seq* seq;
  std::vector<std::unique_ptr<Row>> rows;
  int startAfter;
  explicit Pyramid(const Seq* s, int startAfter_ = -1) : seq(s), startAfter(startAfter_) {}
  Row* get(int index) {
    while ((int)rows.size() <= index) {
      if (rows.empty()) rows.push_back(std::make_unique<BaseRow>(seq));
      else rows.push_back(std::make_unique<ParentRow>(rows.back().get(), startAfter));
    }
    return rows[index].get();
  }
};

}  // namespace xo
