// ORACLE — TEST INFRASTRUCTURE ONLY (see kz_core.hpp header).  parity unpinned for bitstreams;
// BWT pinned by the 'mississippi' KAT (BWT.java:45-50).
//
// Transform stage: Null, LZ (LZXCodec, extra=false/true), LZP (LZPCodec), ROLZ (ROLZCodec1), BWTBlockCodec + BWT,
// SBRT (RANK/MTFT), SRT, ZRLT, Sequence, TransformFactory.
#pragma once
#include "kz_core.hpp"
#include "kz_entropy.hpp"
#include <memory>

namespace kzo {

struct Transform {
  virtual ~Transform() {}
  virtual bool forward(Slice& src, Slice& dst) = 0;
  virtual bool inverse(Slice& src, Slice& dst) = 0;
  virtual int getMaxEncodedLength(int srcLen) = 0;
};

// common precondition block shared by most transforms (e.g. ZRLT.java:57-63)
static inline bool basicCheck(const Slice& in, const Slice& out) {
  return !((in.index < 0) || (out.index < 0) || (in.length < 0) ||
           ((i64)in.index + in.length > in.cap()) || (out.index > out.cap()));
}

// ---- NullTransform (transform/NullTransform.java:36-66) ----------------------------------------
struct NullTransform : Transform {
  static bool doCopy(Slice& input, Slice& output) {
    if (input.length == 0) return true;
    if (!basicCheck(input, output)) return false;
    const int count = input.length;
    if (output.length - output.index < count) return false;
    if ((input.arr != output.arr) || (input.index != output.index))
      memmove(output.p() + output.index, input.p() + input.index, count);
    input.index += count; output.index += count;
    return true;
  }
  bool forward(Slice& s, Slice& d) override { return doCopy(s, d); }
  bool inverse(Slice& s, Slice& d) override { return doCopy(s, d); }
  int getMaxEncodedLength(int n) override { return n; }
};

// ---- ZRLT (transform/ZRLT.java) --------------------------------------------------------------------
struct ZRLT : Transform {
  // forward, ZRLT.java:54-136
  bool forward(Slice& input, Slice& output) override {
    if (input.length == 0) return true;
    if (!basicCheck(input, output)) return false;
    if (input.arr == output.arr) return false;
    const int count = input.length;
    if (output.length - output.index < getMaxEncodedLength(count)) return false;
    const u8* src = input.p(); u8* dst = output.p();
    int srcIdx = input.index, dstIdx = output.index;
    const int srcEnd = srcIdx + count, dstEnd = dstIdx + count;
    bool res = true;
    if (dstIdx < dstEnd) {
      while (srcIdx < srcEnd) {
        if (src[srcIdx] == 0) {
          int runLength = 1;
          while ((srcIdx + runLength < srcEnd) && (src[srcIdx + runLength] == src[srcIdx])) runLength++;
          srcIdx += runLength;
          runLength++;
          int log2 = log2i((u32)runLength);
          if (dstIdx >= dstEnd - log2) { res = false; break; }
          while (log2 > 0) { log2--; dst[dstIdx++] = (u8)((runLength >> log2) & 1); }
          continue;
        }
        const int val = src[srcIdx];
        if (val >= 0xFE) {
          if (dstIdx >= dstEnd - 1) { res = false; break; }
          dst[dstIdx] = 0xFF; dst[dstIdx + 1] = (u8)(val - 0xFE); dstIdx += 2;
        } else {
          if (dstIdx >= dstEnd) { res = false; break; }
          dst[dstIdx] = (u8)(val + 1); dstIdx++;
        }
        srcIdx++;
      }
    }
    input.index = srcIdx; output.index = dstIdx;
    return res && (srcIdx == srcEnd);
  }
  // inverse, ZRLT.java:146-233
  bool inverse(Slice& input, Slice& output) override {
    if (input.length == 0) return true;
    if (!basicCheck(input, output)) return false;
    if (input.arr == output.arr) return false;
    const int count = input.length;
    int srcIdx = input.index, dstIdx = output.index;
    const u8* src = input.p(); u8* dst = output.p();
    const int srcEnd = srcIdx + count, dstEnd = output.length;
    int runLength = 0;
    while (true) {
      int val = src[srcIdx];
      if (val <= 1) {
        runLength = 1;
        bool brk = false;
        do {
          runLength += (runLength + val);
          srcIdx++;
          if (srcIdx >= srcEnd) { brk = true; break; }
          val = src[srcIdx];
        } while (val <= 1);
        if (brk) break;
        runLength--;
        if (runLength > 0) {
          if (dstIdx + runLength >= dstEnd) break;
          while (runLength > 0) { runLength--; dst[dstIdx++] = 0; }
        }
      }
      if (val == 0xFF) {
        srcIdx++;
        if (srcIdx >= srcEnd) break;
        dst[dstIdx] = (u8)(0xFE + src[srcIdx]);
      } else {
        dst[dstIdx] = (u8)(val - 1);
      }
      srcIdx++; dstIdx++;
      if ((srcIdx >= srcEnd) || (dstIdx >= dstEnd)) break;
    }
    if (runLength > 0) {
      runLength--;
      if (dstIdx + runLength > dstEnd) return false;
      while (runLength > 0) { runLength--; dst[dstIdx++] = 0; }
    }
    input.index = srcIdx; output.index = dstIdx;
    return srcIdx == srcEnd;
  }
  int getMaxEncodedLength(int n) override { return n; }
};

// ---- SBRT (transform/SBRT.java), modes MTF=1, RANK=2, TIMESTAMP=3 ----------------------------------
struct SBRT : Transform {
  int mode;
  explicit SBRT(int m) : mode(m) {}
  static bool check(const Slice& input, const Slice& output) {   // SBRT.java:90-106
    if ((input.index < 0) || (output.index < 0) || (input.length < 0) || (output.length < 0) ||
        (input.index > input.length) || (output.index > output.length) ||
        ((i64)input.index + input.length > input.cap()) || ((i64)output.index + output.length > output.cap()))
      return false;
    if (input.arr == output.arr) return false;
    const int count = input.length;
    if ((count > input.length - input.index) || (count > output.length - output.index)) return false;
    if (output.index + count > output.cap()) return false;
    return true;
  }
  // forward, SBRT.java:87-151
  bool forward(Slice& input, Slice& output) override {
    if (input.length == 0) return true;
    if (!check(input, output)) return false;
    const int count = input.length;
    const u8* src = input.p() + input.index; u8* dst = output.p() + output.index;
    int p[256], q[256], s2r[256], r2s[256];
    const int m1 = (mode == 3) ? 0 : -1, m2 = (mode == 1) ? 0 : -1, s = (mode == 2) ? 1 : 0;
    for (int i = 0; i < 256; i++) { p[i] = 0; q[i] = 0; s2r[i] = i; r2s[i] = i; }
    for (int i = 0; i < count; i++) {
      const int c = src[i];
      int r = s2r[c];
      dst[i] = (u8)r;
      const int qc = ((i & m1) + (p[c] & m2)) >> s;
      p[c] = i; q[c] = qc;
      while ((r > 0) && (q[r2s[r - 1]] <= qc)) { r2s[r] = r2s[r - 1]; s2r[r2s[r]] = r; r--; }
      r2s[r] = c; s2r[c] = r;
    }
    input.index += count; output.index += count;
    return true;
  }
  // inverse, SBRT.java:154-214
  bool inverse(Slice& input, Slice& output) override {
    if (input.length == 0) return true;
    if (!check(input, output)) return false;
    const int count = input.length;
    const u8* src = input.p() + input.index; u8* dst = output.p() + output.index;
    int p[256], q[256], r2s[256];
    const int m1 = (mode == 3) ? 0 : -1, m2 = (mode == 1) ? 0 : -1, s = (mode == 2) ? 1 : 0;
    for (int i = 0; i < 256; i++) { p[i] = 0; q[i] = 0; r2s[i] = i; }
    for (int i = 0; i < count; i++) {
      int r = src[i];
      const int c = r2s[r];
      dst[i] = (u8)c;
      const int qc = ((i & m1) + (p[c] & m2)) >> s;
      p[c] = i; q[c] = qc;
      while ((r > 0) && (q[r2s[r - 1]] <= qc)) { r2s[r] = r2s[r - 1]; r--; }
      r2s[r] = c;
    }
    input.index += count; output.index += count;
    return true;
  }
  int getMaxEncodedLength(int n) override { return n; }
};

// ---- SRT (transform/SRT.java) -----------------------------------------------------------------------
struct SRT : Transform {
  enum { MAX_HEADER_SIZE = 4 * 256 };
  // preprocess, SRT.java:266-302 (shell sort of present symbols by freq desc, then symbol asc)
  static int preprocess(const int* freqs, u8* symbols) {
    int nbSymbols = 0;
    for (int i = 0; i < 256; i++) if (freqs[i] > 0) symbols[nbSymbols++] = (u8)i;
    int h = 4;
    while (h < nbSymbols) h = h * 3 + 1;
    while (true) {
      h /= 3;
      for (int i = h; i < nbSymbols; i++) {
        const int t = symbols[i];
        int b = i - h;
        while ((b >= 0) && ((freqs[symbols[b]] < freqs[t]) || ((freqs[t] == freqs[symbols[b]]) && (t < symbols[b])))) {
          symbols[b + h] = symbols[b];
          b -= h;
        }
        symbols[b + h] = (u8)t;
      }
      if (h == 1) break;
    }
    return nbSymbols;
  }
  static int encodeHeader(const int* freqs, u8* dst, int dstIdx) {   // SRT.java:312-325
    for (int i = 0; i < 256; i++) {
      u32 f = (u32)freqs[i];
      while (f >= 128) { dst[dstIdx++] = (u8)(0x80 | f); f >>= 7; }
      dst[dstIdx++] = (u8)f;
    }
    return dstIdx;
  }
  static int decodeHeader(const u8* src, int srcIdx, int srcCap, int* freqs) {   // SRT.java:335-353
    for (int i = 0; i < 256; i++) {
      if (srcIdx >= srcCap) throw JavaException("AIOOBE in SRT.decodeHeader");
      int val = src[srcIdx++];
      int res = val & 0x7F;
      int shift = 7;
      while (val >= 128) {
        if (srcIdx >= srcCap) throw JavaException("AIOOBE in SRT.decodeHeader");
        val = src[srcIdx++];
        res |= ((val & 0x7F) << shift);
        if (shift > 21) break;
        shift += 7;
      }
      freqs[i] = res;
    }
    return srcIdx;
  }
  // forward, SRT.java:73-168 (NB: indexes src from 0 and uses encodeHeader's absolute return as a
  // length — SURVEY E-4 — restated as written)
  bool forward(Slice& input, Slice& output) override {
    if (input.length == 0) return true;
    if (!basicCheck(input, output)) return false;
    if (input.arr == output.arr) return false;
    const int count = input.length;
    if (output.length - output.index < getMaxEncodedLength(count)) return false;
    const u8* src = input.p();
    const int srcEnd = input.index + count;
    int freqs[256], r2s[256], s2r[256], buckets[256];
    u8 symbols[256];
    for (int i = 0; i < 256; i++) { freqs[i] = 0; r2s[i] = 0; s2r[i] = 0; buckets[i] = 0; }
    for (int i = input.index, b = 0; i < srcEnd;) {
      const u8 val = src[i];
      const int c = val;
      if (freqs[c] == 0) { r2s[b] = c; s2r[c] = (int)(int8_t)b; b++; }
      int j = i + 1;
      while ((j < count) && (src[j] == val)) j++;
      freqs[c] += (j - i);
      i = j;
    }
    int nbSymbols = preprocess(freqs, symbols);
    for (int i = 0, bucketPos = 0; i < nbSymbols; i++) {
      const int c = symbols[i];
      buckets[c] = bucketPos;
      bucketPos += freqs[c];
    }
    const int headerSize = encodeHeader(freqs, output.p(), output.index);
    output.index += headerSize;
    const int dstIdx = output.index;
    u8* dst = output.p();
    for (int i = 0; i < count;) {
      const int c = src[i];
      int r = s2r[c] & 0xFF;
      int p = buckets[c];
      dst[dstIdx + p] = (u8)r;
      p++;
      if (r != 0) {
        do { r2s[r] = r2s[r - 1]; s2r[r2s[r]] = r; r--; } while (r != 0);
        r2s[0] = c; s2r[c] = 0;
      }
      i++;
      while ((i < count) && (src[i] == c)) { dst[dstIdx + p] = 0; p++; i++; }
      buckets[c] = p;
    }
    input.index += count; output.index += count;
    return true;
  }
  // inverse, SRT.java:178-257
  bool inverse(Slice& input, Slice& output) override {
    if (input.length == 0) return true;
    if (!basicCheck(input, output)) return false;
    if (input.arr == output.arr) return false;
    int freqs[256];
    const int headerSize = decodeHeader(input.p(), input.index, input.cap(), freqs);
    input.index += headerSize;
    const int count = input.length - headerSize;
    if (count > output.length - output.index) return false;
    const u8* src = input.p();
    const int srcIdx = input.index;
    u8 symbols[256];
    int nbSymbols = preprocess(freqs, symbols);
    int buckets[256], bucketEnds[256], r2s[256];
    for (int i = 0; i < 256; i++) { buckets[i] = 0; bucketEnds[i] = 0; r2s[i] = 0; }
    for (int i = 0, bucketPos = 0; i < nbSymbols; i++) {
      const int c = symbols[i];
      if ((srcIdx + bucketPos < 0) || (srcIdx + bucketPos >= input.length)) return false;
      r2s[src[srcIdx + bucketPos]] = c;
      buckets[c] = bucketPos + 1;
      bucketPos += freqs[c];
      bucketEnds[c] = bucketPos;
    }
    int c = r2s[0];
    u8* dst = output.p();
    const int dstIdx = output.index;
    for (int i = 0; i < count; i++) {
      dst[dstIdx + i] = (u8)c;
      if (buckets[c] < bucketEnds[c]) {
        if (srcIdx + buckets[c] >= input.cap()) throw JavaException("AIOOBE in SRT.inverse");
        const int r = src[srcIdx + buckets[c]];
        buckets[c]++;
        if (r == 0) continue;
        for (int s = 0; s < r; s++) r2s[s] = r2s[s + 1];
        r2s[r] = c;
        c = r2s[0];
      } else {
        if (nbSymbols == 1) continue;
        nbSymbols--;
        for (int s = 0; s < nbSymbols; s++) r2s[s] = r2s[s + 1];
        c = r2s[0];
      }
    }
    input.index += count; output.index += count;
    return true;
  }
  int getMaxEncodedLength(int n) override { return n + MAX_HEADER_SIZE; }
};

// ---- suffix array (oracle's own SA-IS; BWT.forward only needs *a* correct suffix sort, SURVEY B-5) --
namespace sais {
static inline void getBuckets(const int* C, int* B, int k, bool end) {
  int sum = 0;
  for (int i = 0; i < k; i++) { sum += C[i]; B[i] = end ? sum : sum - C[i]; }
}
template <typename T>
static void induce(const T* s, int* SA, const std::vector<bool>& t, int n, int k, const int* C, int* B) {
  getBuckets(C, B, k, false);
  for (int i = 0; i < n; i++) {
    const int j = SA[i] - 1;
    if (SA[i] > 0 && !t[j]) SA[B[s[j]]++] = j;
  }
  getBuckets(C, B, k, true);
  for (int i = n - 1; i >= 0; i--) {
    const int j = SA[i] - 1;
    if (SA[i] > 0 && t[j]) SA[--B[s[j]]] = j;
  }
}
// s[n-1] must be a unique smallest sentinel
template <typename T>
static void run(const T* s, int* SA, int n, int k) {
  std::vector<bool> t(n);
  t[n - 1] = true;
  for (int i = n - 2; i >= 0; i--) t[i] = (s[i] < s[i + 1]) || (s[i] == s[i + 1] && t[i + 1]);
  auto isLMS = [&](int i) { return i > 0 && t[i] && !t[i - 1]; };
  std::vector<int> C(k, 0), B(k);
  for (int i = 0; i < n; i++) C[s[i]]++;
  getBuckets(C.data(), B.data(), k, true);
  for (int i = 0; i < n; i++) SA[i] = -1;
  for (int i = 1; i < n; i++) if (isLMS(i)) SA[--B[s[i]]] = i;
  induce(s, SA, t, n, k, C.data(), B.data());
  int n1 = 0;
  for (int i = 0; i < n; i++) if (isLMS(SA[i])) SA[n1++] = SA[i];
  for (int i = n1; i < n; i++) SA[i] = -1;
  int name = 0, prev = -1;
  for (int i = 0; i < n1; i++) {
    const int pos = SA[i];
    bool diff = false;
    for (int d = 0; d < n; d++) {
      if (prev == -1 || s[pos + d] != s[prev + d] || t[pos + d] != t[prev + d]) { diff = true; break; }
      if (d > 0 && (isLMS(pos + d) || isLMS(prev + d))) break;
    }
    if (diff) { name++; prev = pos; }
    SA[n1 + (pos >> 1)] = name - 1;
  }
  for (int i = n - 1, j = n - 1; i >= n1; i--) if (SA[i] >= 0) SA[j--] = SA[i];
  int* SA1 = SA; int* s1 = SA + n - n1;
  if (name < n1) run<int>(s1, SA1, n1, name);
  else for (int i = 0; i < n1; i++) SA1[s1[i]] = i;
  getBuckets(C.data(), B.data(), k, true);
  for (int i = 1, j = 0; i < n; i++) if (isLMS(i)) s1[j++] = i;
  for (int i = 0; i < n1; i++) SA1[i] = s1[SA1[i]];
  for (int i = n1; i < n; i++) SA[i] = -1;
  for (int i = n1 - 1; i >= 0; i--) { const int j = SA[i]; SA[i] = -1; SA[--B[s[j]]] = j; }
  induce(s, SA, t, n, k, C.data(), B.data());
}
// suffix array of the n suffixes of T (shorter-is-smaller): append a sentinel, drop its slot
static inline void suffixArray(const u8* T, int n, std::vector<int>& SA) {
  std::vector<int> s(n + 1);
  for (int i = 0; i < n; i++) s[i] = T[i] + 1;
  s[n] = 0;
  std::vector<int> sa(n + 1);
  run<int>(s.data(), sa.data(), n + 1, 257);
  SA.assign(sa.begin() + 1, sa.end());
}
}  // namespace sais

// ---- BWT (transform/BWT.java) ------------------------------------------------------------------------
struct BWT {
  enum { MAX_BLOCK_SIZE = 1 << 30, NB_FASTBITS = 17, MASK_FASTBITS = (1 << 17) - 1, BLOCK_SIZE_THRESHOLD1 = 256,
         BLOCK_SIZE_THRESHOLD2 = 8 * 1024 * 1024 };
  int primaryIndexes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int jobs = 1;
  bool asref = true;   // keep the BWT.java:152-156 `dst.index + dst.length > dst.array.length` clause (SURVEY E-1)
  static int getBWTChunks(int size) { return (size < BLOCK_SIZE_THRESHOLD1) ? 1 : 8; }
  int getPrimaryIndex(int n) const { return primaryIndexes[n]; }
  bool setPrimaryIndex(int n, int p) { if (p < 0 || n < 0 || n >= 8) return false; primaryIndexes[n] = p; return true; }

  // BWT.java:151-170 (forward) / 198-216 (inverse).  With asref == false ("fixed", SURVEY E-1) the three
  // clauses that make every call coming from BWTBlockCodec fail are dropped: BWTBlockCodec advances
  // output.index past its header without shrinking output.length (forward: dst.index + dst.length >
  // dst.array.length), and advances input.index while setting input.length = n - header (inverse:
  // count > src.length - src.index).  As written, BWTBlockCodec.forward AND .inverse always return false.
  bool check(const Slice& src, const Slice& dst) const {
    if ((src.index < 0) || (dst.index < 0) || (src.length < 0) || (dst.length <= 0) || (dst.index > dst.length) ||
        (src.index + src.length > src.cap()))
      return false;
    if (asref && ((src.index > src.length) || (dst.index + dst.length > dst.cap()))) return false;
    if (src.arr == dst.arr) return false;
    const int count = src.length;
    if (asref && (count > src.length - src.index)) return false;
    if (count > dst.length - dst.index) return false;
    if (count > MAX_BLOCK_SIZE) return false;
    if (dst.index + count > dst.cap()) return false;
    return true;
  }

  // forward, BWT.java:148-191 + DivSufSort.computeBWT (DivSufSort.java:204-227) restated through the
  // mathematical definition (SURVEY B-5): any correct suffix sort gives these bytes and indexes.
  bool forward(Slice& src, Slice& dst) {
    if (src.length == 0) return true;
    if (!check(src, dst)) return false;
    const int count = src.length;
    if (count == 1) { dst.p()[dst.index++] = src.p()[src.index++]; return true; }
    const u8* T = src.p() + src.index; u8* out = dst.p() + dst.index;
    std::vector<int> SA;
    sais::suffixArray(T, count, SA);
    const int chunks = getBWTChunks(count);
    const int st = count / chunks;
    const int step = (st * chunks != count) ? st + 1 : st;
    int pIdx = -1;
    out[0] = T[count - 1];
    for (int i = 0; i < count; i++) {
      const int s = SA[i];
      if (s == 0) { pIdx = i; continue; }
      if (s % step == 0) primaryIndexes[s / step] = i + 1;
      out[(pIdx < 0) ? i + 1 : i] = T[s - 1];
    }
    primaryIndexes[0] = pIdx + 1;
    src.index += count; dst.index += count;
    return true;
  }

  // inverse, BWT.java:195-235
  bool inverse(Slice& src, Slice& dst) {
    if (src.length == 0) return true;
    if (!check(src, dst)) return false;
    const int count = src.length;
    if (count == 1) { dst.p()[dst.index++] = src.p()[src.index++]; return true; }
    if (count <= BLOCK_SIZE_THRESHOLD2) return inverseMergeTPSI(src, dst, count);
    return inverseBiPSIv2(src, dst, count);
  }

  // inverseMergeTPSI, BWT.java:245-374
  bool inverseMergeTPSI(Slice& src, Slice& dst, int count) {
    std::vector<u32> data(std::max(count, 64));
    const u8* input = src.p() + src.index; u8* output = dst.p() + dst.index;
    int b[257];
    const int pIdx = getPrimaryIndex(0);
    if ((pIdx <= 0) || (pIdx > count)) return false;
    histogramOrder0(input, 0, count, b, false);
    for (int i = 0, sum = 0; i < 256; i++) { const int tmp = b[i]; b[i] = sum; sum += tmp; }
    const int val0 = input[0];
    data[b[val0]] = 0xFF00 | val0; b[val0]++;
    for (int i = 1; i < pIdx; i++) { const int val = input[i]; data[b[val]] = ((u32)(i - 1) << 8) | val; b[val]++; }
    for (int i = pIdx; i < count; i++) { const int val = input[i]; data[b[val]] = ((u32)i << 8) | val; b[val]++; }
    if (getBWTChunks(count) != 8) {
      for (int i = 0, t = pIdx - 1; i < count; i++) {
        if (t < 0 || t >= (int)data.size()) throw JavaException("AIOOBE in inverseMergeTPSI");
        const u32 ptr = data[t]; output[i] = (u8)ptr; t = (int)(ptr >> 8);
      }
    } else {
      const int ckSize = ((count & 7) == 0) ? count >> 3 : (count >> 3) + 1;
      int t[8];
      for (int k = 0; k < 8; k++) t[k] = getPrimaryIndex(k) - 1;
      for (int k = 0; k < 8; k++) if (t[k] < 0) return false;
      for (int k = 0; k < 8; k++) if (t[k] >= count) return false;
      const int end = count - ckSize * 7;
      int n = 0;
      auto stepk = [&](int k) {
        if (t[k] < 0 || t[k] >= (int)data.size()) throw JavaException("AIOOBE in inverseMergeTPSI");
        const u32 ptr = data[t[k]];
        if (n + ckSize * k >= dst.cap() - dst.index) throw JavaException("AIOOBE in inverseMergeTPSI");
        output[n + ckSize * k] = (u8)ptr; t[k] = (int)(ptr >> 8);
      };
      while (n < end) { for (int k = 0; k < 8; k++) stepk(k); n++; }
      while (n < ckSize) { for (int k = 0; k < 7; k++) stepk(k); n++; }
    }
    src.index += count; dst.index += count;
    return true;
  }

  // inverseBiPSIv2 + InverseBiPSIv2Task, BWT.java:384-544, 568-674 (single job: one task runs all chunks)
  bool inverseBiPSIv2(Slice& src, Slice& dst, int count) {
    std::vector<int> data(std::max(count + 1, 64), 0);
    std::vector<int> b(65536, 0);
    std::vector<uint16_t> fastBits(MASK_FASTBITS + 1, 0);
    int freqs_[257];
    const u8* inArr = src.p(); u8* outArr = dst.p();
    const int srcIdx = src.index, dstIdx = dst.index, srcIdx2 = src.index - 1;
    const int pIdx = getPrimaryIndex(0);
    if ((pIdx <= 0) || (pIdx > count)) return false;
    for (int i = 1; i < 8; i++) { const int p = getPrimaryIndex(i); if ((p <= 0) || (p > count)) return false; }
    histogramOrder0(inArr, srcIdx, srcIdx + count, freqs_, false);
    for (int sum = 1, c = 0; c < 256; c++) {
      const int f = sum;
      sum += freqs_[c];
      freqs_[c] = f;
      if (f != sum) {
        const int c256 = c << 8;
        const int hi = (sum < pIdx) ? sum : pIdx;
        for (int i = f; i < hi; i++) b[c256 | inArr[srcIdx + i]]++;
        const int lo = (f - 1 > pIdx) ? f - 1 : pIdx;
        for (int i = lo; i < sum - 1; i++) b[c256 | inArr[srcIdx + i]]++;
      }
    }
    const int lastc = inArr[srcIdx];
    int shift = 0;
    while ((count >> shift) > MASK_FASTBITS) shift++;
    for (int v = 0, sum = 1, c = 0; c < 256; c++) {
      if (c == lastc) sum++;
      for (int d = 0; d < 256; d++) {
        const int s = sum;
        sum += b[(d << 8) | c];
        b[(d << 8) | c] = s;
        if (s != sum) for (; v <= ((sum - 1) >> shift); v++) fastBits[v] = (uint16_t)((c << 8) | d);
      }
    }
    for (int i = 0; i < pIdx; ++i) {
      const int c = inArr[srcIdx + i];
      const int p = freqs_[c]; freqs_[c]++;
      if (p < pIdx) { const int idx = (c << 8) | inArr[srcIdx + p]; data[b[idx]] = i; b[idx]++; }
      else if (p > pIdx) { const int idx = (c << 8) | inArr[srcIdx2 + p]; data[b[idx]] = i; b[idx]++; }
    }
    for (int i = pIdx; i < count; i++) {
      const int c = inArr[srcIdx + i];
      const int p = freqs_[c]; freqs_[c]++;
      if (p < pIdx) { const int idx = (c << 8) | inArr[srcIdx + p]; data[b[idx]] = i + 1; b[idx]++; }
      else if (p > pIdx) { const int idx = (c << 8) | inArr[srcIdx2 + p]; data[b[idx]] = i + 1; b[idx]++; }
    }
    for (int c = 0; c < 256; c++) {
      const int c256 = c << 8;
      for (int d = 0; d < c; d++) { const int tmp = b[(d << 8) | c]; b[(d << 8) | c] = b[c256 | d]; b[c256 | d] = tmp; }
    }
    const int chunks = getBWTChunks(count);
    const int st = count / chunks;
    const int ckSize = (chunks * st == count) ? st : st + 1;
    // one task: firstChunk 0, lastChunk = chunks, start = dstIdx (InverseBiPSIv2Task.call, :597-673)
    {
      const int total = count;
      int start = dstIdx;
      int c = 0;
      const int lastChunk = chunks;
      const int outCap = dst.cap();
      auto put = [&](int i, int v) { if (i < 0 || i >= outCap) throw JavaException("AIOOBE in biPSIv2"); outArr[i] = (u8)v; };
      auto look = [&](int p) -> int {
        if (p < 0 || (p >> shift) > MASK_FASTBITS) throw JavaException("AIOOBE in biPSIv2");
        int s = fastBits[p >> shift];
        while (true) { if (s >= 65536) throw JavaException("AIOOBE in biPSIv2"); if (b[s] > p) break; s++; }
        return s;
      };
      auto nxt = [&](int p) -> int { if (p < 0 || p >= (int)data.size()) throw JavaException("AIOOBE in biPSIv2"); return data[p]; };
      if (start + 4 * ckSize < total) {
        for (; c + 3 < lastChunk; c += 4) {
          const int end = start + ckSize;
          int p0 = getPrimaryIndex(c), p1 = getPrimaryIndex(c + 1), p2 = getPrimaryIndex(c + 2), p3 = getPrimaryIndex(c + 3);
          for (int i = start + 1; i <= end; i += 2) {
            const int s0 = look(p0), s1 = look(p1), s2 = look(p2), s3 = look(p3);
            put(i - 1, s0 >> 8); put(i, s0);
            put(1 * ckSize + i - 1, s1 >> 8); put(1 * ckSize + i, s1);
            put(2 * ckSize + i - 1, s2 >> 8); put(2 * ckSize + i, s2);
            put(3 * ckSize + i - 1, s3 >> 8); put(3 * ckSize + i, s3);
            p0 = nxt(p0); p1 = nxt(p1); p2 = nxt(p2); p3 = nxt(p3);
          }
          start = end + 3 * ckSize;
        }
      }
      for (; c < lastChunk; c++) {
        const int end = std::min(start + ckSize, total - 1);
        int p = getPrimaryIndex(c);
        for (int i = start + 1; i <= end; i += 2) {
          const int s = look(p);
          put(i - 1, s >> 8); put(i, s);
          p = nxt(p);
        }
        start = end;
      }
    }
    outArr[dstIdx + count - 1] = (u8)lastc;
    src.index += count; dst.index += count;
    return true;
  }
};

// ---- BWTBlockCodec (transform/BWTBlockCodec.java) --------------------------------------------------
struct BWTBlockCodec : Transform {
  BWT bwt;
  int bsVersion;
  explicit BWTBlockCodec(const Ctx& ctx) : bsVersion(ctx.bsVersion) { bwt.asref = (ctx.bwtBounds != 0); bwt.jobs = ctx.jobs; }
  // forward, BWTBlockCodec.java:71-128
  bool forward(Slice& input, Slice& output) override {
    if (input.length == 0) return true;
    if ((input.index < 0) || (output.index < 0) || (input.length < 0) || (input.index + input.length > input.cap()) ||
        (output.index > output.cap()))
      return false;
    if (input.arr == output.arr) return false;
    const int blockSize = input.length;
    const int maxEncodedLength = getMaxEncodedLength(blockSize);
    if ((output.length < output.index) || (output.length > output.cap()) || (maxEncodedLength < blockSize) ||
        (output.length - output.index < maxEncodedLength))
      return false;
    int logBlockSize = log2i((u32)blockSize);
    if ((blockSize & (blockSize - 1)) != 0) logBlockSize++;
    const int pIndexSize = (logBlockSize + 7) >> 3;
    if ((pIndexSize <= 0) || (pIndexSize >= 5)) return false;
    const int chunks = BWT::getBWTChunks(blockSize);
    const int logNbChunks = log2i((u32)chunks);
    if (logNbChunks > 7) return false;
    const int idx0 = output.index;
    output.index += (1 + chunks * pIndexSize);
    if (!bwt.forward(input, output)) return false;
    const u8 mode = (u8)((logNbChunks << 2) | (pIndexSize - 1));
    u8* oa = output.p();
    for (int i = 0, idx = idx0 + 1; i < chunks; i++) {
      const int primaryIndex = bwt.getPrimaryIndex(i) - 1;
      int shift = (pIndexSize - 1) << 3;
      while (shift >= 0) { oa[idx++] = (u8)(primaryIndex >> shift); shift -= 8; }
    }
    oa[idx0] = mode;
    return true;
  }
  // inverse, BWTBlockCodec.java:138-213 (bsVersion > 5 branch)
  bool inverse(Slice& input, Slice& output) override {
    if (input.length == 0) return true;
    if (!basicCheck(input, output)) return false;
    if (input.arr == output.arr) return false;
    const int blockSize = input.length;
    const u8* ia = input.p();
    const u8 mode = ia[input.index++];
    const int logNbChunks = ((int8_t)mode >> 2) & 0x07;
    const int pIndexSize = (mode & 0x03) + 1;
    const int chunks = 1 << logNbChunks;
    const int headerSize = 1 + chunks * pIndexSize;
    if (blockSize < headerSize) return false;
    if (chunks != BWT::getBWTChunks(blockSize - headerSize)) return false;
    for (int i = 0; i < chunks; i++) {
      int shift = (pIndexSize - 1) << 3;
      i64 primaryIndex = 0;
      while (shift >= 0) { primaryIndex = (primaryIndex << 8) | ia[input.index++]; shift -= 8; }
      if (primaryIndex >= 0x7FFFFFFFLL) return false;
      if (!bwt.setPrimaryIndex(i, (int)primaryIndex + 1)) return false;
    }
    input.length = blockSize - headerSize;
    return bwt.inverse(input, output);
  }
  int getMaxEncodedLength(int n) override { return n + 33; }
};

// ---- LZ: LZCodec -> LZXCodec (transform/LZCodec.java) ------------------------------------------------
struct LZX : Transform {
  enum { HASH_SEED = 0x1E35A7BD, HASH_LOG1 = 16, HASH_LOG2 = 19, MAX_DISTANCE1 = (1 << 16) - 2, MAX_DISTANCE2 = (1 << 24) - 2,
         MIN_MATCH4 = 4, MIN_MATCH6 = 6, MAX_MATCH = 65535 + 254 + 4, MIN_BLOCK_LENGTH = 24 };
  bool extra; Ctx* ctx;
  LZX(Ctx* c, bool extra_) : extra(extra_), ctx(c) {}

  static bool differentInts(const u8* a, int s, int d) { return le32(a + s) != le32(a + d); }   // LZCodec.java:82-85
  // emitLength, LZCodec.java:211-231 (bounds-checked: Java would throw AIOOBE)
  static int emitLength(std::vector<u8>& block, int idx, int length) {
    auto need = [&](int k) { if (idx + k > (int)block.size()) throw JavaException("AIOOBE in LZ emitLength"); };
    if (length < 254) { need(1); block[idx] = (u8)length; return idx + 1; }
    if (length < 65536 + 254) {
      length -= 254; need(3);
      block[idx] = 254; block[idx + 1] = (u8)(length >> 8); block[idx + 2] = (u8)length;
      return idx + 3;
    }
    length -= 255; need(4);
    block[idx] = 255; block[idx + 1] = (u8)(length >> 16); block[idx + 2] = (u8)(length >> 8); block[idx + 3] = (u8)length;
    return idx + 4;
  }
  static int readLength(const u8* a, int& index) {    // LZCodec.java:241-258
    int res = a[index++];
    if (res < 254) return res;
    if (res == 254) { res += (a[index++] << 8); res += a[index++]; return res; }
    res += (a[index] << 16); res += (a[index + 1] << 8); res += a[index + 2];
    index += 3;
    return res;
  }
  static int findMatch(const u8* src, int srcIdx, int ref, int maxMatch) {    // LZCodec.java:271-287
    int bestLen = 0;
    while (bestLen + 8 <= maxMatch) {
      const u64 diff = le64(src + srcIdx + bestLen) ^ le64(src + ref + bestLen);
      if (diff != 0) { bestLen += (__builtin_ctzll(diff) >> 3); break; }
      bestLen += 8;
    }
    return bestLen;
  }
  int hash(const u8* block, int idx) const {     // LZCodec.java:904-911
    if (extra) return (int)(((le64(block + idx) << 24) * (u64)(i64)HASH_SEED) >> (64 - HASH_LOG2));
    return (int)(((le64(block + idx) << 24) * (u64)(i64)HASH_SEED) >> (64 - HASH_LOG1));
  }
  int getMaxEncodedLength(int srcLen) override { return ((srcLen <= 1024) ? srcLen + 16 : srcLen + (srcLen / 64)) + 2; }

  // LZCodec.forward wrapper (LZCodec.java:52-63) + LZXCodec.forward (LZCodec.java:299-597)
  bool forward(Slice& input, Slice& output) override {
    if (input.length == 0) return true;
    if (!basicCheck(input, output)) return false;
    if (input.arr == output.arr) return false;
    const int count = input.length;
    if (output.length - output.index < getMaxEncodedLength(count)) return false;
    if (count < MIN_BLOCK_LENGTH) return false;
    std::vector<i32> hashes(extra ? (1 << HASH_LOG2) : (1 << HASH_LOG1), 0);
    const int minBufSize = std::max(count / 5, 256);
    std::vector<u8> mBuf(minBufSize), mLenBuf(minBufSize), tkBuf(minBufSize);
    const int srcIdx0 = input.index, dstIdx0 = output.index;
    const u8* src = input.p();
    // dst writes go through a bounds-checked lambda-free path: dst capacity >= maxEncodedLength holds
    std::vector<u8>& dstv = *output.arr;
    u8* dst = output.p();
    const int srcEnd = srcIdx0 + count - 16 - 2;
    const int maxDist = (srcEnd < 4 * MAX_DISTANCE1) ? MAX_DISTANCE1 : MAX_DISTANCE2;
    dst[dstIdx0 + 12] = (maxDist == MAX_DISTANCE1) ? 0 : 1;
    int mm = MIN_MATCH4;
    if (ctx != nullptr) {
      if (ctx->dataType == DT_DNA) mm = MIN_MATCH6;
      else if (ctx->dataType == DT_SMALL_ALPHABET) return false;
    }
    dst[dstIdx0 + 12] |= (u8)(((mm - 2) & 0x07) << 1);
    const int minMatch = mm;
    int srcIdx = srcIdx0, anchor = srcIdx0, dstIdx = dstIdx0 + 13;
    int mIdx = 0, mLenIdx = 0, tkIdx = 0;
    int repd[2] = {count, count};
    int repIdx = 0, srcInc = 0;
    auto tkPut = [&](int v) { if (tkIdx >= (int)tkBuf.size()) throw JavaException("AIOOBE tkBuf"); tkBuf[tkIdx++] = (u8)v; };

    while (srcIdx < srcEnd) {
      int bestLen = 0;
      const int h0 = hash(src, srcIdx);
      const int ref0 = hashes[h0];
      hashes[h0] = srcIdx;
      const int srcIdx1 = srcIdx + 1;
      int ref = srcIdx1 - repd[repIdx];
      const int minRef = std::max(srcIdx - maxDist, srcIdx0);
      if ((ref > minRef) && !differentInts(src, ref, srcIdx1)) {
        bestLen = findMatch(src, srcIdx1, ref, std::min(srcEnd - srcIdx1, (int)MAX_MATCH));
      } else {
        ref = srcIdx1 - repd[repIdx ^ 1];
        if ((ref > minRef) && !differentInts(src, ref, srcIdx1))
          bestLen = findMatch(src, srcIdx1, ref, std::min(srcEnd - srcIdx1, (int)MAX_MATCH));
      }
      if (bestLen < minMatch) {
        ref = ref0;
        if ((ref > minRef) && !differentInts(src, ref, srcIdx))
          bestLen = findMatch(src, srcIdx, ref, std::min(srcEnd - srcIdx, (int)MAX_MATCH));
        if (bestLen < minMatch) {
          srcIdx = srcIdx1 + (srcInc >> 6);
          srcInc++;
          repIdx = 0;
          continue;
        }
        if ((ref != srcIdx - repd[0]) && (ref != srcIdx - repd[1])) {
          const int h1 = hash(src, srcIdx1);
          const int ref1 = hashes[h1];
          hashes[h1] = srcIdx1;
          if ((ref1 > minRef + 1) && !differentInts(src, ref1 + bestLen - 3, srcIdx1 + bestLen - 3)) {
            const int maxMatch = std::min(srcEnd - srcIdx1, (int)MAX_MATCH);
            const int bestLen1 = findMatch(src, srcIdx1, ref1, maxMatch);
            if (bestLen1 >= bestLen) { ref = ref1; bestLen = bestLen1; srcIdx = srcIdx1; }
          }
          if (extra) {
            const int srcIdx2 = srcIdx1 + 1;
            const int h2 = hash(src, srcIdx2);
            const int ref2 = hashes[h2];
            hashes[h2] = srcIdx2;
            if ((ref2 > minRef + 2) && !differentInts(src, ref2 + bestLen - 3, srcIdx2 + bestLen - 3)) {
              const int maxMatch = std::min(srcEnd - srcIdx2, (int)MAX_MATCH);
              const int bestLen2 = findMatch(src, srcIdx2, ref2, maxMatch);
              if (bestLen2 >= bestLen) { ref = ref2; bestLen = bestLen2; srcIdx = srcIdx2; }
            }
          }
        }
        while ((srcIdx > anchor) && (ref > minRef) && (src[srcIdx - 1] == src[ref - 1])) { bestLen++; ref--; srcIdx--; }
        if (bestLen > MAX_MATCH) { ref += (bestLen - MAX_MATCH); srcIdx += (bestLen - MAX_MATCH); bestLen = MAX_MATCH; }
      } else {
        if ((bestLen >= MAX_MATCH) || (src[srcIdx] != src[ref - 1])) {
          srcIdx++;
          const int h1 = hash(src, srcIdx);
          hashes[h1] = srcIdx;
        } else {
          bestLen++; ref--;
        }
      }
      srcInc = 0;
      const int dist = srcIdx - ref;
      int token, mLenTh;
      if (dist == repd[0]) { token = 0x00; mLenTh = 3; }
      else if (dist == repd[1]) { token = 0x04; mLenTh = 3; }
      else {
        if (mIdx + 2 >= (int)mBuf.size()) throw JavaException("AIOOBE mBuf");
        mBuf[mIdx] = (u8)(dist >> 16);
        const int inc1 = dist >= 65536 ? 1 : 0; mIdx += inc1;
        mBuf[mIdx] = (u8)(dist >> 8);
        const int inc2 = dist >= 256 ? 1 : 0; mIdx += inc2;
        mBuf[mIdx++] = (u8)dist;
        token = (inc1 + inc2 + 1) << 3;
        mLenTh = 7;
      }
      const int mLen = bestLen - minMatch;
      if (mLen >= mLenTh) { token += mLenTh; mLenIdx = emitLength(mLenBuf, mLenIdx, mLen - mLenTh); }
      else token += mLen;
      repd[1] = repd[0]; repd[0] = dist; repIdx = 1;
      const int litLen = srcIdx - anchor;
      if (litLen == 0) {
        tkPut(token);
      } else {
        if (litLen >= 7) {
          if (litLen >= (1 << 24)) return false;
          tkPut((7 << 5) | token);
          dstIdx = emitLength(dstv, dstIdx, litLen - 7);
        } else {
          tkPut((litLen << 5) | token);
        }
        // emitLiterals copies in 8-byte chunks (LZCodec.java:945-950); the overshoot lands in bytes that
        // are overwritten later or lie beyond the final length; only [dstIdx, dstIdx+litLen) matters.
        if (dstIdx + ((litLen + 7) & ~7) > (int)dstv.size()) throw JavaException("AIOOBE LZ emitLiterals");
        memcpy(dst + dstIdx, src + anchor, litLen);
        dstIdx += litLen;
      }
      if (mIdx >= (int)mBuf.size() - 8) {
        mBuf.resize((mBuf.size() * 3) / 2);
        if (mLenIdx >= (int)mLenBuf.size() - 4) mLenBuf.resize((mLenBuf.size() * 3) / 2);
      }
      anchor = srcIdx + bestLen;
      while (srcIdx + 4 < anchor) {
        srcIdx += 4;
        hashes[hash(src, srcIdx - 3)] = srcIdx - 3;
        hashes[hash(src, srcIdx - 2)] = srcIdx - 2;
        hashes[hash(src, srcIdx - 1)] = srcIdx - 1;
        hashes[hash(src, srcIdx)] = srcIdx;
      }
      while (++srcIdx < anchor) hashes[hash(src, srcIdx)] = srcIdx;
    }
    const int litLen = count - anchor;   // NB: as written (assumes srcIdx0 == 0)
    if (dstIdx + litLen + tkIdx + mIdx + mLenIdx >= output.index + count) return false;
    if (litLen >= 7) { tkPut(7 << 5); dstIdx = emitLength(dstv, dstIdx, litLen - 7); }
    else tkPut(litLen << 5);
    memcpy(dst + dstIdx, src + anchor, litLen);
    dstIdx += litLen;
    put_le32(dst + dstIdx0, (u32)dstIdx);
    put_le32(dst + dstIdx0 + 4, (u32)tkIdx);
    put_le32(dst + dstIdx0 + 8, (u32)mIdx);
    memcpy(dst + dstIdx, tkBuf.data(), tkIdx); dstIdx += tkIdx;
    memcpy(dst + dstIdx, mBuf.data(), mIdx); dstIdx += mIdx;
    memcpy(dst + dstIdx, mLenBuf.data(), mLenIdx); dstIdx += mLenIdx;
    input.index += count;
    output.index = dstIdx;
    return (dstIdx - dstIdx0) <= count - (count / 100);
  }

  // LZXCodec.inverseV6, LZCodec.java:626-756 (bsVersion >= 6)
  bool inverse(Slice& input, Slice& output) override {
    if (input.length == 0) return true;
    if (!basicCheck(input, output)) return false;
    if (input.arr == output.arr) return false;
    if (input.length < 13) return false;
    const int count = input.length;
    const int srcIdx0 = input.index, dstIdx0 = output.index;
    const u8* src = input.p(); u8* dst = output.p();
    const int srcCap = input.cap();
    const int dstEnd = output.cap();
    const i32 tkLen = (i32)le32(src + srcIdx0), mIdxLen = (i32)le32(src + srcIdx0 + 4), mLenLen = (i32)le32(src + srcIdx0 + 8);
    if ((tkLen < 0) || (mIdxLen < 0) || (mLenLen < 0)) return false;
    if ((tkLen < 13) || (tkLen > count) || (mIdxLen > count - tkLen) || (mLenLen > count - tkLen - mIdxLen)) return false;
    int tkIdx = srcIdx0 + tkLen;
    int mIdx = tkIdx + mIdxLen;
    int mLenIdx = mIdx + mLenLen;
    const int srcEnd = tkIdx - 13;
    const int litEnd = tkIdx;
    const int maxDist = ((src[srcIdx0 + 12] & 1) == 0) ? MAX_DISTANCE1 : MAX_DISTANCE2;
    const int minMatch = (((int8_t)src[srcIdx0 + 12] >> 1) & 0x07) + 2;
    int srcIdx = srcIdx0 + 13, dstIdx = dstIdx0;
    int repd0 = count, repd1 = count;
    auto rd = [&](int i) -> int { if (i < 0 || i >= srcCap) throw JavaException("AIOOBE LZ inverse src"); return src[i]; };
    auto readLen = [&](int& index) -> int {
      if (index + 4 > srcCap) { int r = rd(index); if (r >= 254) { rd(index + 1); rd(index + 2); if (r == 255) rd(index + 3); } }
      return readLength(src, index);
    };
    while (true) {
      const int token = rd(tkIdx++);
      if (token >= 32) {
        const int litLen = (token >= 0xE0) ? 7 + readLen(srcIdx) : token >> 5;
        if ((litLen > dstEnd - dstIdx) || (litLen > litEnd - srcIdx)) { input.index = srcIdx; output.index = dstIdx; return false; }
        // (arraycopy near the end, 8-byte chunk copy otherwise; chunk copy may touch <= 7 bytes past litLen)
        if (!(srcIdx + litLen >= srcEnd)) {
          const int padded = (litLen + 7) & ~7;
          if (srcIdx + padded > srcCap || dstIdx + padded > dstEnd) throw JavaException("AIOOBE LZ inverse emitLiterals");
          memcpy(dst + dstIdx, src + srcIdx, padded);
        } else {
          memcpy(dst + dstIdx, src + srcIdx, litLen);
        }
        srcIdx += litLen; dstIdx += litLen;
        if (srcIdx >= srcEnd) break;
      }
      int mLen, dist;
      const int f = token & 0x18;
      if (f == 0) {
        mLen = token & 0x03;
        mLen += (mLen == 3) ? minMatch + readLen(mLenIdx) : minMatch;
        dist = ((token & 0x04) == 0) ? repd0 : repd1;
      } else {
        mLen = token & 0x07;
        mLen += (mLen == 7 ? minMatch + readLen(mLenIdx) : minMatch);
        dist = rd(mIdx++);
        if (f == 0x18) { dist = (dist << 8) | rd(mIdx++); dist = (dist << 8) | rd(mIdx++); }
        else if (f == 0x10) { dist = (dist << 8) | rd(mIdx++); }
      }
      repd1 = repd0; repd0 = dist;
      const int mEnd = dstIdx + mLen;
      int ref = dstIdx - dist;
      if ((ref < dstIdx0) || (dist > maxDist) || (mEnd > dstEnd)) { input.index = srcIdx; output.index = dstIdx; return false; }
      if (dist >= 16) {
        do {
          if (dstIdx + 16 > dstEnd) throw JavaException("AIOOBE LZ inverse copy");   // System.arraycopy bound
          memmove(dst + dstIdx, dst + ref, 16);
          ref += 16; dstIdx += 16;
        } while (dstIdx < mEnd);
      } else {
        for (int i = 0; i < mLen; i++) dst[dstIdx + i] = dst[ref + i];
      }
      dstIdx = mEnd;
    }
    output.index = dstIdx;
    input.index = srcIdx0 + count;
    return srcIdx == srcEnd + 13;
  }
};

// ---- ROLZ: ROLZCodec -> ROLZCodec1 (transform/ROLZCodec.java) ----------------------------------------
static const int DNA_SYMBOLS_[12] = {'a', 'c', 'g', 'n', 't', 'u', 'A', 'C', 'G', 'N', 'T', 'U'};
static const char NUMERIC_SYMBOLS_[] = "0123456789+-*/=,.:; ";
static const char BASE64_SYMBOLS_[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
// Global.detectSimpleType, Global.java:556-608
static inline int detectSimpleType(int count, const int* freqs0) {
  if (count == 0) return DT_UNDEFINED;
  int sum = 0;
  for (int i = 0; i < 12; i++) sum += freqs0[DNA_SYMBOLS_[i]];
  if (sum > count - count / 12) return DT_DNA;
  sum = 0;
  for (int i = 0; i < 20; i++) sum += freqs0[(u8)NUMERIC_SYMBOLS_[i]];
  if (sum == count) return DT_NUMERIC;
  sum = (freqs0[0x3D] == 1) ? 1 : 0;
  for (int i = 0; i < 64; i++) sum += freqs0[(u8)BASE64_SYMBOLS_[i]];
  if (sum == count) return DT_BASE64;
  sum = 0;
  for (int i = 0; i < 256; i++) sum += (freqs0[i] > 0) ? 1 : 0;
  if (sum == 256) return DT_BIN;
  if (sum <= 4) return DT_SMALL_ALPHABET;
  return DT_UNDEFINED;
}

struct ROLZ1 : Transform {
  enum { HASH_SIZE = 65536, CHUNK_SIZE = 16 * 1024 * 1024, HASH = 200002979, HASH_MASK = ~(CHUNK_SIZE - 1),
         MAX_BLOCK_SIZE = 1 << 30, MIN_BLOCK_SIZE = 64, MIN_MATCH3 = 3, MIN_MATCH4 = 4, MIN_MATCH7 = 7,
         MAX_MATCH = 3 + 65535, LOG_POS_CHECKS = 4 };
  int logPosChecks = LOG_POS_CHECKS, maskChecks = 15, posChecks = 16, minMatch = 3;
  std::vector<i32> counters, matches;
  Ctx* ctx;
  explicit ROLZ1(Ctx* c) : counters(1 << 16, 0), ctx(c) {}

  static int getKey1(const u8* buf, int idx) { return (int)le16(buf + idx); }                        // ROLZCodec.java:123-125
  static int getKey2(const u8* buf, int idx) { return (int)((i64)(le64(buf + idx) * (u64)(i64)HASH) >> 40) & 0xFFFF; }   // :135-137
  static i32 hash(const u8* buf, int idx) { return (i32)(((le32(buf + idx) << 8) * (u32)HASH) & (u32)HASH_MASK); }     // :147-149
  int getMaxEncodedLength(int n) override { return (n <= 512) ? n + 64 : n; }

  // findMatch, ROLZCodec.java:365-406.  sba = (buf, length = endChunk, index = startChunk)
  int findMatch(const u8* buf, int sbaLength, int sbaIndex, int pos, i32 hash32, int counter, int base) {
    int bestLen = 0, bestIdx = -1;
    const int maxMatch = std::min((int)MAX_MATCH, sbaLength - pos) - 8;
    for (int i = counter; i > counter - posChecks; i--) {
      i32 ref = matches[base + (i & maskChecks)];
      if ((ref & HASH_MASK) != hash32) continue;
      ref = (ref & ~HASH_MASK) + sbaIndex;
      if (buf[ref + bestLen] != buf[pos + bestLen]) continue;
      int n = 0;
      while (n < maxMatch) {
        const u64 diff = le64(buf + ref + n) ^ le64(buf + pos + n);
        if (diff != 0) { n += (__builtin_ctzll(diff) >> 3); break; }
        n += 8;
      }
      if (n > bestLen) { bestIdx = counter - i; bestLen = n; }
    }
    return (bestLen < minMatch) ? -1 : (bestIdx << 16) | (bestLen - minMatch);
  }
  struct Buf { std::vector<u8> a; int index = 0; void put(int v) { if (index >= (int)a.size()) throw JavaException("AIOOBE ROLZ buf"); a[index++] = (u8)v; } };
  static void emitLength(Buf& lenBuf, int length) {     // ROLZCodec.java:670-683
    if (length >= 1 << 7) {
      if (length >= 1 << 14) {
        if (length >= 1 << 21) lenBuf.put(0x80 | (length >> 21));
        lenBuf.put(0x80 | (length >> 14));
      }
      lenBuf.put(0x80 | (length >> 7));
    }
    lenBuf.put(length & 0x7F);
  }
  static int readLength(Buf& lenBuf) {                  // ROLZCodec.java:969-989
    auto g = [&]() -> int { if (lenBuf.index >= (int)lenBuf.a.size()) throw JavaException("AIOOBE ROLZ readLength"); return (int8_t)lenBuf.a[lenBuf.index++]; };
    int next = g();
    int length = next & 0x7F;
    if ((next & 0x80) != 0) {
      next = g(); length = (length << 7) | (next & 0x7F);
      if ((next & 0x80) != 0) {
        next = g(); length = (length << 7) | (next & 0x7F);
        if ((next & 0x80) != 0) { next = g(); length = (length << 7) | (next & 0x7F); }
      }
    }
    return length;
  }

  // ROLZCodec.forward wrapper (:200-214) + ROLZCodec1.forward (:419-661)
  bool forward(Slice& input, Slice& output) override {
    if (input.length == 0) return true;
    if (!basicCheck(input, output)) return false;
    if (input.length < MIN_BLOCK_SIZE) return false;
    if (input.arr == output.arr) return false;
    if (input.length > MAX_BLOCK_SIZE) return false;
    const int count = input.length;
    if (output.length - output.index < getMaxEncodedLength(count)) return false;
    const u8* src = input.p(); u8* dst = output.p();
    const int dstLen = output.cap();
    const int srcEnd = input.index + count - 4;
    put_be32(dst + output.index, (u32)count);
    int sizeChunk = std::min(count, (int)CHUNK_SIZE);
    int startChunk = input.index;
    Buf litBuf, lenBuf, mIdxBuf, tkBuf;
    litBuf.a.resize(getMaxEncodedLength(sizeChunk)); lenBuf.a.resize(sizeChunk / 5);
    mIdxBuf.a.resize(sizeChunk / 4); tkBuf.a.resize(sizeChunk / 4);
    std::fill(counters.begin(), counters.end(), 0);
    const int litOrder = (count < (1 << 17)) ? 0 : 1;
    int flags = litOrder;
    minMatch = MIN_MATCH3;
    int delta = 2;
    if (ctx != nullptr) {
      int dtp = ctx->dataType;
      if (dtp == DT_UNDEFINED) {
        int freqs0[257];
        histogramOrder0(src, 0, count, freqs0, false);
        dtp = detectSimpleType(count, freqs0);
        if (dtp != DT_UNDEFINED) ctx->dataType = dtp;
      }
      switch (dtp) {
        case DT_EXE: delta = 3; flags |= 8; break;
        case DT_MULTIMEDIA: delta = 8; minMatch = MIN_MATCH4; flags |= 2; break;
        case DT_DNA: delta = 8; minMatch = MIN_MATCH7; flags |= 4; break;
        default: break;
      }
    }
    const int mm = minMatch, dt = delta;
    flags |= (logPosChecks << 4);
    dst[output.index + 4] = (u8)flags;
    int dstIdx = output.index + 5;
    if (matches.empty()) matches.assign((size_t)HASH_SIZE << logPosChecks, 0);
    while (startChunk < srcEnd) {
      litBuf.index = 0; lenBuf.index = 0; mIdxBuf.index = 0; tkBuf.index = 0;
      std::fill(matches.begin(), matches.end(), 0);
      const int endChunk = std::min(startChunk + sizeChunk, srcEnd);
      sizeChunk = endChunk - startChunk;
      int srcIdx = startChunk;
      const int sbaLength = endChunk, sbaIndex = startChunk;
      const int n = std::min(srcEnd - startChunk, 8);
      for (int j = 0; j < n; j++) litBuf.put(src[srcIdx++]);
      int firstLitIdx = srcIdx;
      int srcInc = 0;
      while (srcIdx < endChunk) {
        int key = (mm == MIN_MATCH3) ? getKey1(src, srcIdx - dt) : getKey2(src, srcIdx - dt);
        int base = key << logPosChecks;
        i32 hash32 = hash(src, srcIdx);
        int counter = counters[key];
        int match = findMatch(src, sbaLength, sbaIndex, srcIdx, hash32, counter, base);
        counters[key] = (counters[key] + 1) & maskChecks;
        matches[base + counters[key]] = hash32 | (srcIdx - sbaIndex);
        if (match == -1) { srcIdx++; srcIdx += (srcInc >> 6); srcInc++; continue; }
        {
          key = (mm == MIN_MATCH3) ? getKey1(src, srcIdx + 1 - dt) : getKey2(src, srcIdx + 1 - dt);
          base = key << logPosChecks;
          hash32 = hash(src, srcIdx + 1);
          counter = counters[key];
          const int match2 = findMatch(src, sbaLength, sbaIndex, srcIdx + 1, hash32, counter, base);
          if ((match2 >= 0) && ((match2 & 0xFFFF) > (match & 0xFFFF))) {
            match = match2;
            srcIdx++;
            counters[key] = (counters[key] + 1) & maskChecks;
            matches[base + counters[key]] = hash32 | (srcIdx - sbaIndex);
          }
        }
        const int litLen = srcIdx - firstLitIdx;
        const int token = (litLen < 31) ? (litLen << 3) : 0xF8;
        const int mLen = match & 0xFFFF;
        if (mLen >= 7) { tkBuf.put(token | 0x07); emitLength(lenBuf, mLen - 7); }
        else tkBuf.put(token | mLen);
        if (litLen >= 31) emitLength(lenBuf, litLen - 31);
        if (litBuf.index + litLen > (int)litBuf.a.size()) throw JavaException("AIOOBE ROLZ litBuf");
        memcpy(&litBuf.a[litBuf.index], src + firstLitIdx, litLen);
        litBuf.index += litLen;
        mIdxBuf.put((u32)match >> 16);
        srcIdx += (mLen + mm);
        firstLitIdx = srcIdx;
        srcInc = 0;
      }
      srcIdx = sizeChunk;
      const int litLen = srcIdx - (firstLitIdx - startChunk);
      if (tkBuf.index != 0) { const int token = (litLen >= 31) ? 0xF8 : (litLen << 3); tkBuf.put(token); }
      if (litLen >= 31) emitLength(lenBuf, litLen - 31);
      if (litLen < 0 || litBuf.index + litLen > (int)litBuf.a.size()) throw JavaException("AIOOBE ROLZ litBuf");
      memcpy(litBuf.a.data() + litBuf.index, src + firstLitIdx, litLen);
      litBuf.index += litLen;
      BitWriter obs;
      obs.writeBits((u32)litBuf.index, 32); obs.writeBits((u32)tkBuf.index, 32);
      obs.writeBits((u32)lenBuf.index, 32); obs.writeBits((u32)mIdxBuf.index, 32);
      { ans::Encoder litEnc(obs, litOrder); litEnc.encode(litBuf.a.data(), 0, litBuf.index); }
      { ans::Encoder mEnc(obs, 0, 32768);
        mEnc.encode(tkBuf.a.data(), 0, tkBuf.index);
        mEnc.encode(lenBuf.a.data(), 0, lenBuf.index);
        mEnc.encode(mIdxBuf.a.data(), 0, mIdxBuf.index); }
      obs.close();
      const int bufLen = (int)obs.buf.size();
      if (dstIdx + bufLen > dstLen) { output.index = dstIdx; input.index = srcIdx; return false; }
      memcpy(dst + dstIdx, obs.buf.data(), bufLen);
      dstIdx += bufLen;
      startChunk = endChunk;
    }
    if (dstIdx + 4 > dstLen) { output.index = dstIdx; input.index = startChunk; return false; }
    if (dstIdx + 4 > output.length) {
      input.index = srcEnd;
    } else {
      dst[dstIdx++] = src[srcEnd]; dst[dstIdx++] = src[srcEnd + 1]; dst[dstIdx++] = src[srcEnd + 2]; dst[dstIdx++] = src[srcEnd + 3];
      input.index = srcEnd + 4;
    }
    output.index = dstIdx;
    return (input.index == srcEnd + 4) && ((dstIdx - output.index) < count);
  }

  // ROLZCodec.inverse wrapper (:217-229) + ROLZCodec1.inverse (:696-960), bsVersion >= 4
  bool inverse(Slice& input, Slice& output) override {
    if (input.length == 0) return true;
    if (!basicCheck(input, output)) return false;
    if (input.arr == output.arr) return false;
    if (input.length > MAX_BLOCK_SIZE) return false;
    const int count = input.length;
    const u8* src = input.p(); u8* dst = output.p();
    const int srcEnd = input.index + count;
    if (input.index + 5 > input.cap()) throw JavaException("AIOOBE ROLZ inverse header");
    const int szBlock = (i32)be32(src + input.index) - 4;
    if ((szBlock <= 0) || (szBlock > output.length)) return false;
    const int dstEnd = output.index + szBlock;
    int sizeChunk = std::min(szBlock, (int)CHUNK_SIZE);
    int startChunk = output.index;
    Buf litBuf, lenBuf, mIdxBuf, tkBuf;
    litBuf.a.assign(sizeChunk, 0); lenBuf.a.assign((sizeChunk / 5) + 4, 0);
    mIdxBuf.a.assign(sizeChunk / 4, 0); tkBuf.a.assign(sizeChunk / 4, 0);
    std::fill(counters.begin(), counters.end(), 0);
    const int flags = src[input.index + 4];
    const int litOrder = flags & 0x01;
    minMatch = MIN_MATCH3;
    int delta = 2;
    logPosChecks = flags >> 4;
    if ((logPosChecks < 2) || (logPosChecks > 8)) return false;
    if (matches.size() < ((size_t)HASH_SIZE << logPosChecks)) matches.assign((size_t)HASH_SIZE << logPosChecks, 0);
    posChecks = 1 << logPosChecks; maskChecks = posChecks - 1;
    switch (flags & 0x0E) {
      case 2: minMatch = MIN_MATCH4; delta = 8; break;
      case 4: minMatch = MIN_MATCH7; delta = 8; break;
      case 8: delta = 3; break;
      default: break;
    }
    const int mm = minMatch, dt = delta;
    int srcIdx = input.index + 5;
    const int dstCap = output.cap();
    while (startChunk < dstEnd) {
      litBuf.index = 0; lenBuf.index = 0; mIdxBuf.index = 0; tkBuf.index = 0;
      std::fill(matches.begin(), matches.end(), 0);
      const int endChunk = std::min(startChunk + sizeChunk, dstEnd);
      sizeChunk = endChunk - startChunk;
      int dstIdx = output.index;
      bool onlyLiterals = false;
      int litLenDecoded = 0, tkLen = 0, mLenLen = 0, mIdxLen = 0;
      {
        if (srcEnd - srcIdx < 0) throw JavaException("ROLZ inverse: negative stream length");
        BitReader ibs(src + srcIdx, (u64)(srcEnd - srcIdx) * 8);
        int litLen = (i32)ibs.readBits(32);
        tkLen = (i32)ibs.readBits(32); mLenLen = (i32)ibs.readBits(32); mIdxLen = (i32)ibs.readBits(32);
        const int firstLitLen = std::min(sizeChunk, 8);
        auto fail = [&]() { input.index = srcIdx; output.index = dstIdx; return false; };
        if ((litLen < 0) || (tkLen < 0) || (mLenLen < 0) || (mIdxLen < 0)) return fail();
        // NB: litBuf.length etc. are SliceByteArray.length == 0 in Java (constructed with (array, 0) -> length = array.length?)
        if ((litLen > (int)litBuf.a.size()) || (tkLen > (int)tkBuf.a.size()) || (mLenLen > (int)lenBuf.a.size() - 4) ||
            (mIdxLen > (int)mIdxBuf.a.size()))
          return fail();
        if ((litLen < firstLitLen) || (litLen > sizeChunk) || ((tkLen == 0) && (mIdxLen != 0)) ||
            ((tkLen > 0) && (mIdxLen + 1 != tkLen)))
          return fail();
        litLenDecoded = litLen;
        { ans::Decoder litDec(ibs, litOrder); litDec.decode(litBuf.a.data(), 0, litLen); }
        { ans::Decoder mDec(ibs, 0, 32768);
          mDec.decode(tkBuf.a.data(), 0, tkLen);
          mDec.decode(lenBuf.a.data(), 0, mLenLen);
          mDec.decode(mIdxBuf.a.data(), 0, mIdxLen); }
        onlyLiterals = tkLen == 0;
        srcIdx += (int)((ibs.read() + 7) >> 3);
      }
      if (onlyLiterals) {
        if (litLenDecoded != sizeChunk) { input.index = srcIdx; output.index = dstIdx; return false; }
        memcpy(dst + output.index, litBuf.a.data(), sizeChunk);
        startChunk = endChunk;
        output.index += sizeChunk;
        continue;
      }
      const int n = std::min(dstEnd - dstIdx, 8);
      for (int j = 0; j < n; j++) dst[dstIdx++] = litBuf.a[litBuf.index++];
      while (dstIdx < endChunk) {
        if (tkBuf.index >= (int)tkBuf.a.size()) throw JavaException("AIOOBE ROLZ tkBuf");
        const int token = tkBuf.a[tkBuf.index++];
        int matchLen = token & 0x07;
        if (matchLen == 7) {
          if (lenBuf.index >= mLenLen) { output.index = dstIdx; input.index = srcIdx; return false; }
          matchLen = readLength(lenBuf) + 7;
        }
        int litLen;
        if (token < 0xF8) litLen = token >> 3;
        else {
          if (lenBuf.index >= mLenLen) { output.index = dstIdx; input.index = srcIdx; return false; }
          litLen = readLength(lenBuf) + 31;
        }
        if (litLen > 0) {
          int srcInc = 0;
          const int n0 = dstIdx - output.index;
          if (litBuf.index + litLen > (int)litBuf.a.size() || dstIdx + litLen > dstCap) throw JavaException("AIOOBE ROLZ literal copy");
          memcpy(dst + dstIdx, &litBuf.a[litBuf.index], litLen);
          for (int j = 0; j < litLen; j++) {
            const int key = (mm == MIN_MATCH3) ? getKey1(dst, dstIdx + j - dt) : getKey2(dst, dstIdx + j - dt);
            counters[key] = (counters[key] + 1) & maskChecks;
            matches[(key << logPosChecks) + counters[key]] = n0 + j;
            j += (srcInc >> 6);
            srcInc++;
          }
          litBuf.index += litLen;
          dstIdx += litLen;
          if (dstIdx >= endChunk) {
            if (dstIdx == endChunk) break;
            output.index = dstIdx; input.index = srcIdx; return false;
          }
        }
        if (dstIdx + matchLen + mm > dstEnd) { output.index = dstIdx; input.index = srcIdx; return false; }
        const int key = (mm == MIN_MATCH3) ? getKey1(dst, dstIdx - dt) : getKey2(dst, dstIdx - dt);
        const int base = key << logPosChecks;
        if (mIdxBuf.index >= (int)mIdxBuf.a.size()) throw JavaException("AIOOBE ROLZ mIdxBuf");
        const int matchIdx = mIdxBuf.a[mIdxBuf.index++];
        int ref = output.index + matches[base + ((counters[key] - matchIdx) & maskChecks)];
        const int savedIdx = dstIdx;
        { // emitCopy, ROLZCodec.java:162-179 (byte-wise forward copy semantics)
          int ml = matchLen + minMatch;
          if (ref < 0 || dstIdx + ml > dstCap) throw JavaException("AIOOBE ROLZ emitCopy");
          while (ml-- > 0) dst[dstIdx++] = dst[ref++];
        }
        counters[key] = (counters[key] + 1) & maskChecks;
        matches[base + counters[key]] = savedIdx - output.index;
      }
      if ((tkBuf.index != tkLen) || (mIdxBuf.index != mIdxLen) || (litBuf.index != litLenDecoded) || (lenBuf.index != mLenLen)) {
        output.index = dstIdx; input.index = srcIdx; return false;
      }
      startChunk = endChunk;
      output.index = dstIdx;
    }
    if ((output.index + 4 > output.length) || (srcEnd - srcIdx != 4)) { input.index = srcIdx; return false; }
    dst[output.index++] = src[srcIdx++]; dst[output.index++] = src[srcIdx++];
    dst[output.index++] = src[srcIdx++]; dst[output.index++] = src[srcIdx++];
    input.index = srcIdx;
    return input.index == srcEnd;
  }
};

// ---- ROLZX = ROLZCodec2 (transform/ROLZCodec.java:1016-1428) with its binary arithmetic coder ROLZEncoder / ROLZDecoder
// (:1431-1597, :1599-1770).  Selected here by the ROLZX id (the reference looks for "ROLZX" in ctx["transform"], :106-112).
struct ROLZ2 : Transform {
  enum { HASH_SIZE = 65536, CHUNK_SIZE = 16 * 1024 * 1024, MAX_BLOCK_SIZE = 1 << 30, MIN_BLOCK_SIZE = 64, MIN_MATCH3 = 3, MIN_MATCH7 = 7,
         MAX_MATCH = 3 + 255, LOG_POS_CHECKS = 5, MATCH_FLAG = 0, LITERAL_FLAG = 1, LITERAL_CTX = 0, MATCH_CTX = 1 };
  static constexpr u32 HASH_MASK = ~(u32)(CHUNK_SIZE - 1);
  static constexpr u64 TOP = 0x00FFFFFFFFFFFFFFULL, MASK_0_56 = 0x00FFFFFFFFFFFFFFULL, MASK_0_32 = 0x00000000FFFFFFFFULL;
  int logPosChecks = LOG_POS_CHECKS, maskChecks = 31, posChecks = 32, minMatch = 3;
  std::vector<i32> counters, matches;
  Ctx* ctx;
  explicit ROLZ2(Ctx* c) : counters(1 << 16, 0), matches((size_t)HASH_SIZE << LOG_POS_CHECKS, 0), ctx(c) {}
  int getMaxEncodedLength(int n) override { return (n <= 16384) ? n + 1024 : n + (n / 32); }     // :1417-1421

  struct Coder {                       // the state ROLZEncoder and ROLZDecoder share
    u64 low = 0, high = TOP, current = 0;
    std::vector<i32> probs[2];
    int logSizes[2];
    int c1 = 1, cx = 0, pIdx = LITERAL_FLAG;
    u8* arr; int cap; int* index;
    Coder(int litLogSize, int mLogSize, u8* a, int cap_, int* idx) : arr(a), cap(cap_), index(idx) {
      probs[MATCH_CTX].assign((size_t)256 << mLogSize, 0xFFFF >> 1);
      probs[LITERAL_CTX].assign((size_t)256 << litLogSize, 0xFFFF >> 1);
      logSizes[MATCH_CTX] = mLogSize; logSizes[LITERAL_CTX] = litLogSize;
    }
    void setContext(int n, u8 c) { pIdx = n; cx = (int)c << logSizes[pIdx]; }
    void need(int k) const { if (*index < 0 || *index + k > cap) throw JavaException("AIOOBE in ROLZ coder"); }
    // ROLZEncoder.encodeBit, :1553-1580
    void encodeBit(int bit) {
      i32& p = probs[pIdx][cx + c1];
      const u64 split = (((high - low) >> 4) * (u64)((u32)p >> 4)) >> 8;
      if (bit == 0) { low += (split + 1); p -= (p >> 5); c1 += c1; }
      else { high = low + split; p -= (((p - 0xFFFF) >> 5) + 1); c1 += (c1 + 1); }
      while (((low ^ high) >> 24) == 0) {
        need(4);
        put_be32(arr + *index, (u32)(high >> 32));
        *index += 4;
        low <<= 32;
        high = (high << 32) | MASK_0_32;
      }
    }
    void encodeBits(int val, int n) { c1 = 1; do { n--; encodeBit(val & (1 << n)); } while (n != 0); }        // :1526-1535
    void encode9Bits(int val) { c1 = 1; for (int m = 0x100; m != 0; m >>= 1) encodeBit(val & m); }          // :1538-1550
    void disposeEncoder() {                                                                                  // :1583-1590
      need(8);
      for (int i = 0; i < 8; i++) { arr[*index + i] = (u8)((i64)low >> 56); low <<= 8; }
      *index += 8;
    }
    // ROLZDecoder: constructor :1625-1648, decodeBit :1733-1764
    void initDecoder() {
      need(8);
      current = 0;
      for (int i = 0; i < 8; i++) current = (current << 8) | (u64)arr[*index + i];
      *index += 8;
      pIdx = LITERAL_CTX;
    }
    int decodeBit() {
      i32& p = probs[pIdx][cx + c1];
      const u64 mid = low + ((((high - low) >> 4) * (u64)((u32)p >> 4)) >> 8);
      int bit;
      if ((i64)mid >= (i64)current) { bit = 1; high = mid; p -= (((p - 0xFFFF) >> 5) + 1); c1 += (c1 + 1); }
      else { bit = 0; low = mid + 1; p -= (p >> 5); c1 += c1; }
      while (((low ^ high) >> 24) == 0) {
        low = (low << 32) & MASK_0_56;
        high = ((high << 32) | MASK_0_32) & MASK_0_56;
        need(4);
        const u64 val = (u64)be32(arr + *index);
        current = ((current << 32) | val) & MASK_0_56;
        *index += 4;
      }
      return bit;
    }
    int decodeBits(int n) { c1 = 1; const int mask = (1 << n) - 1; do { decodeBit(); n--; } while (n != 0); return c1 & mask; }   // :1702-1715
    int decode9Bits() { c1 = 1; for (int i = 0; i < 9; i++) decodeBit(); return c1 & 0x1FF; }                                   // :1718-1730
  };

  // findMatch, :1114-1173.  sba = (buf, length = endChunk, index = startChunk)
  int findMatch(const u8* buf, int bufCap, int sbaLength, int sbaIndex, int pos, int key) {
    const int base = key << logPosChecks;
    const i32 hash32 = ROLZ1::hash(buf, pos);
    const int counter = counters[key];
    int bestLen = 0, bestIdx = -1;
    const int maxMatch = std::min((int)MAX_MATCH, sbaLength - pos) - 8;
    for (int i = counter; i > counter - posChecks; i--) {
      i32 ref = matches[base + (i & maskChecks)];
      if ((u32)(ref & (i32)HASH_MASK) != (u32)hash32) continue;
      ref = (ref & ~(i32)HASH_MASK) + sbaIndex;
      if (ref + bestLen >= bufCap || pos + bestLen >= bufCap) throw JavaException("AIOOBE in ROLZ2.findMatch");
      if (buf[ref + bestLen] != buf[pos + bestLen]) continue;
      int n = 0;
      while (n < maxMatch) {
        const u64 diff = le64(buf + ref + n) ^ le64(buf + pos + n);
        if (diff != 0) { n += (__builtin_ctzll(diff) >> 3); break; }
        n += 8;
      }
      if (n > bestLen) {
        bestIdx = counter - i; bestLen = n;
        if (bestLen == maxMatch) break;
      }
    }
    counters[key] = (counters[key] + 1) & maskChecks;
    matches[base + counters[key]] = hash32 | (pos - sbaIndex);
    return (bestLen < minMatch) ? -1 : (bestIdx << 16) | (bestLen - minMatch);
  }

  // forward: outer guards ROLZCodec.java:207-237, then :1176-1294
  bool forward(Slice& input, Slice& output) override {
    if (input.length == 0) return true;
    if (!basicCheck(input, output)) return false;
    if (input.length < MIN_BLOCK_SIZE) return false;
    if (input.arr == output.arr) return false;
    if (input.length > MAX_BLOCK_SIZE) return false;
    const int count = input.length;
    if (output.length - output.index < getMaxEncodedLength(count)) return false;
    const u8* src = input.p(); u8* dst = output.p();
    const int srcEnd = input.index + count - 4;
    put_be32(dst + output.index, (u32)count);
    int sizeChunk = std::min(count, (int)CHUNK_SIZE);
    int startChunk = input.index;
    minMatch = MIN_MATCH3;
    int delta = 2, flags = 0;
    if (ctx != nullptr) {
      int dtp = ctx->dataType;
      if (dtp == DT_UNDEFINED) {
        int freqs0[257];
        histogramOrder0(src, 0, count, freqs0, false);
        dtp = detectSimpleType(count, freqs0);
        if (dtp != DT_UNDEFINED) ctx->dataType = dtp;
      }
      if (dtp == DT_EXE) { delta = 3; flags |= 8; }
      else if (dtp == DT_DNA) { delta = 8; minMatch = MIN_MATCH7; flags |= 4; }
    }
    const int mm = minMatch, dt = delta;
    dst[output.index + 4] = (u8)flags;
    int sbaIndex = output.index + 5;
    Coder re(9, logPosChecks, dst, output.cap(), &sbaIndex);
    int srcIdx = input.index;
    std::fill(counters.begin(), counters.end(), 0);
    while (startChunk < srcEnd) {
      std::fill(matches.begin(), matches.end(), 0);
      const int endChunk = std::min(startChunk + sizeChunk, srcEnd);
      srcIdx = startChunk;
      const int n = std::min(srcEnd - startChunk, 8);
      re.setContext(LITERAL_CTX, 0);
      for (int j = 0; j < n; j++) { re.encode9Bits((LITERAL_FLAG << 8) | src[srcIdx]); srcIdx++; }
      while (srcIdx < endChunk) {
        re.setContext(LITERAL_CTX, src[srcIdx - 1]);
        const int key = (mm == MIN_MATCH3) ? ROLZ1::getKey1(src, srcIdx - dt) : ROLZ1::getKey2(src, srcIdx - dt);
        const int match = findMatch(src, input.cap(), endChunk, startChunk, srcIdx, key);
        if (match < 0) { re.encode9Bits((LITERAL_FLAG << 8) | src[srcIdx]); srcIdx++; continue; }
        const int matchLen = match & 0xFFFF;
        re.encode9Bits((MATCH_FLAG << 8) | matchLen);
        re.setContext(MATCH_CTX, src[srcIdx - 1]);
        const int matchIdx = (int)((u32)match >> 16);
        re.encodeBits(matchIdx, logPosChecks);
        srcIdx += (matchLen + minMatch);
      }
      startChunk = endChunk;
    }
    for (int i = 0; i < 4; i++, srcIdx++) {
      re.setContext(LITERAL_CTX, src[srcIdx - 1]);
      re.encode9Bits((LITERAL_FLAG << 8) | src[srcIdx]);
    }
    re.disposeEncoder();
    input.index = srcIdx;
    output.index = sbaIndex;
    return (input.index == srcEnd + 4);            // (the second clause, (output.index - sba1.index) < count, is 0 < count: always true)
  }

  // inverse: outer guards ROLZCodec.java:239-256, then :1296-1414
  bool inverse(Slice& input, Slice& output) override {
    if (input.length == 0) return true;
    if (!basicCheck(input, output)) return false;
    if (input.arr == output.arr) return false;
    if (input.length > MAX_BLOCK_SIZE) return false;
    const int count = input.length;
    u8* src = input.p(); u8* dst = output.p();
    const int srcCap = input.cap(), dstCap = output.cap();
    const int srcEnd = input.index + count;
    if (input.index + 5 > srcCap) throw JavaException("AIOOBE in ROLZ2.inverse");
    const int szBlock = (int)be32(src + input.index);
    if ((szBlock <= 0) || (szBlock > output.length)) return false;
    const int dstEnd = output.index + szBlock;
    int sizeChunk = std::min(szBlock, (int)CHUNK_SIZE);
    int startChunk = output.index;
    minMatch = MIN_MATCH3;
    int delta = 2;
    int srcIdx = input.index + 4;
    const int flags = src[srcIdx++];
    const int bsVersion = (ctx == nullptr) ? 6 : ctx->bsVersion;
    if (bsVersion >= 4) {
      if ((flags & 0x0E) == 8) delta = 3;
      else if ((flags & 0x0E) == 4) { delta = 8; minMatch = MIN_MATCH7; }
    } else if ((bsVersion >= 3) && (flags == 1)) minMatch = MIN_MATCH7;
    const int mm = minMatch, dt = delta;
    int sbaIndex = srcIdx;
    Coder rd(9, logPosChecks, src, srcCap, &sbaIndex);
    rd.initDecoder();
    std::fill(counters.begin(), counters.end(), 0);
    auto put = [&](int i, int v) { if (i < 0 || i >= dstCap) throw JavaException("AIOOBE in ROLZ2.inverse"); dst[i] = (u8)v; };
    auto at = [&](int i) -> int { if (i < 0 || i >= dstCap) throw JavaException("AIOOBE in ROLZ2.inverse"); return dst[i]; };
    while (startChunk < dstEnd) {
      std::fill(matches.begin(), matches.end(), 0);
      const int endChunk = (startChunk + sizeChunk < dstEnd) ? startChunk + sizeChunk : dstEnd;
      int dstIdx = output.index;
      const int n = (bsVersion < 3) ? 2 : std::min(dstEnd - startChunk, 8);
      rd.setContext(LITERAL_CTX, 0);
      for (int j = 0; j < n; j++) {
        const int val1 = rd.decode9Bits();
        if ((val1 >> 8) == MATCH_FLAG) { output.index = dstIdx; return false; }
        put(dstIdx++, val1);
      }
      while (dstIdx < endChunk) {
        const int savedIdx = dstIdx;
        if (dstIdx - dt < 0 || dstIdx - dt + (mm == MIN_MATCH3 ? 2 : 8) > dstCap) throw JavaException("AIOOBE in ROLZ2.inverse");
        const int key = (mm == MIN_MATCH3) ? ROLZ1::getKey1(dst, dstIdx - dt) : ROLZ1::getKey2(dst, dstIdx - dt);
        const int base = key << logPosChecks;
        rd.setContext(LITERAL_CTX, (u8)at(dstIdx - 1));
        const int val = rd.decode9Bits();
        if ((val >> 8) == LITERAL_FLAG) {
          put(dstIdx++, val);
        } else {
          const int matchLen = val & 0xFF;
          if (dstIdx + matchLen + 3 > dstEnd) { output.index = dstIdx; return false; }
          rd.setContext(MATCH_CTX, (u8)at(dstIdx - 1));
          const int matchIdx = rd.decodeBits(logPosChecks);
          const int ref = output.index + matches[base + ((counters[key] - matchIdx) & maskChecks)];
          const int len = matchLen + mm;
          if (ref < 0 || dstIdx + len > dstCap) throw JavaException("AIOOBE in ROLZ2 emitCopy");
          for (int k = 0; k < len; k++) dst[dstIdx + k] = dst[ref + k];
          dstIdx += len;
        }
        counters[key] = (counters[key] + 1) & maskChecks;
        matches[base + counters[key]] = savedIdx - output.index;
      }
      startChunk = endChunk;
      output.index = dstIdx;
    }
    input.index = sbaIndex;
    return input.index == srcEnd;
  }
};

// ---- BWTS (transform/BWTS.java): the bijective Burrows-Wheeler transform (Scott), no primary index.  Not on the CUDA path
// (SURVEY §8f rank 3); here so that the checker is ready.  Suffix array from the oracle's own SA-IS (any correct suffix sort will do).
struct BWTS : Transform {
  enum { MAX_BLOCK_SIZE = 1024 * 1024 * 1024 };
  static bool guards(const Slice& src, const Slice& dst) {                  // BWTS.java:63-83
    if ((src.index < 0) || (dst.index < 0) || (src.length < 0) || (dst.length <= 0) || (src.index > src.length) || (dst.index > dst.length) ||
        ((i64)src.index + src.length > src.cap()) || ((i64)dst.index + dst.length > dst.cap())) return false;
    if (src.arr == dst.arr) return false;
    const int count = src.length;
    if ((count > src.length - src.index) || (count > dst.length - dst.index)) return false;
    if (count > MAX_BLOCK_SIZE) return false;
    if (dst.index + count > dst.cap()) return false;
    return true;
  }
  // moveLyndonWordHead, :163-197
  static int moveLyndonWordHead(std::vector<int>& sa, std::vector<int>& isa, const u8* data, int count, int start, int size, int rank) {
    const int end = start + size;
    while (rank + 1 < count) {
      const int nextStart0 = sa[rank + 1];
      if (nextStart0 <= end) break;
      int nextStart = nextStart0, k = 0;
      while ((k < size) && (nextStart < count) && (data[start + k] == data[nextStart])) { k++; nextStart++; }
      if ((k == size) && (nextStart >= count)) throw JavaException("AIOOBE in BWTS.moveLyndonWordHead");
      if ((k == size) && (rank < isa[nextStart])) break;
      if ((k < size) && (nextStart < count) && (data[start + k] < data[nextStart])) break;
      sa[rank] = nextStart0;
      isa[nextStart0] = rank;
      rank++;
    }
    sa[rank] = start;
    isa[start] = rank;
    return rank;
  }
  // forward, :60-160
  bool forward(Slice& src, Slice& dst) override {
    if (src.length == 0) return true;
    if (!guards(src, dst)) return false;
    const int count = src.length;
    const u8* input = src.p() + src.index; u8* output = dst.p() + dst.index;
    if (count < 2) { output[0] = input[0]; src.index++; dst.index++; return true; }
    std::vector<int> sa, isa((size_t)count, 0);                  // (Java: int[count] each; an index of `count` throws there and here)
    sais::suffixArray(input, count, sa);
    for (int i = 0; i < count; i++) isa[sa[i]] = i;
    int min = isa[0], idxMin = 0;
    for (int i = 1; (i < count) && (min > 0); i++) {
      if (isa[i] >= min) continue;
      int refRank = moveLyndonWordHead(sa, isa, input, count, idxMin, i - idxMin, min);
      for (int j = i - 1; j > idxMin; j--) {
        int testRank = isa[j];
        const int startRank = testRank;
        while (testRank < count - 1) {
          const int nextRankStart = sa[testRank + 1];
          if ((j > nextRankStart) || (input[j] != input[nextRankStart])) break;
          if (nextRankStart + 1 >= count) throw JavaException("AIOOBE in BWTS.forward");
          if (refRank < isa[nextRankStart + 1]) break;
          sa[testRank] = nextRankStart;
          isa[nextRankStart] = testRank;
          testRank++;
        }
        sa[testRank] = j;
        isa[j] = testRank;
        refRank = testRank;
        if (startRank == testRank) break;
      }
      min = isa[i];
      idxMin = i;
    }
    min = count;
    auto before = [&](int i) -> u8 { if (i < 1) throw JavaException("AIOOBE in BWTS.forward"); return input[i - 1]; };   // input[srcIdx - 1 + i]
    for (int i = 0; i < count; i++) {
      if (isa[i] >= min) { output[isa[i]] = before(i); continue; }
      if (min < count) output[min] = before(i);
      min = isa[i];
    }
    output[0] = input[count - 1];
    src.index += count; dst.index += count;
    return true;
  }
  // inverse, :200-262
  bool inverse(Slice& src, Slice& dst) override {
    if (src.length == 0) return true;
    if (!guards(src, dst)) return false;
    const int count = src.length;
    const u8* input = src.p() + src.index; u8* output = dst.p() + dst.index;
    if (count < 2) { output[0] = input[0]; src.index++; dst.index++; return true; }
    int buckets[256] = {0};
    std::vector<int> lf((size_t)count, 0);
    for (int i = 0; i < count; i++) buckets[input[i]]++;
    for (int i = 0, sum = 0; i < 256; i++) { sum += buckets[i]; buckets[i] = sum - buckets[i]; }
    for (int i = 0; i < count; i++) lf[i] = buckets[input[i]]++;
    for (int i = 0, j = count - 1; j >= 0; i++) {
      if (i >= count) throw JavaException("AIOOBE in BWTS.inverse");
      if (lf[i] < 0) continue;
      int p = i;
      do {
        output[j] = input[p];
        j--;
        const int t = lf[p];
        lf[p] = -1;
        p = t;
      } while (lf[p] >= 0);
    }
    src.index += count; dst.index += count;
    return true;
  }
  int getMaxEncodedLength(int n) override { return n; }
};

// ---- RLT (transform/RLT.java) ---------------------------------------------------------------------------------------------------
struct RLT : Transform {
  enum { RUN_LEN_ENCODE1 = 224, RUN_LEN_ENCODE2 = (255 - RUN_LEN_ENCODE1) << 8, RUN_THRESHOLD = 3,
         MAX_RUN = 0xFFFF + RUN_LEN_ENCODE2 + RUN_THRESHOLD - 1, MAX_RUN4 = MAX_RUN - 4, DEFAULT_ESCAPE = 0xFB };
  Ctx* ctx;
  explicit RLT(Ctx* c) : ctx(c) {}
  static int emitRunLength(u8* dst, int dstCap, int dstIdx, int run) {       // RLT.java:233-249
    auto put = [&](int i, int v) { if (i < 0 || i >= dstCap) throw JavaException("AIOOBE in RLT.emitRunLength"); dst[i] = (u8)v; };
    run -= RUN_THRESHOLD;
    if (run >= RUN_LEN_ENCODE1) {
      if (run < RUN_LEN_ENCODE2) { run -= RUN_LEN_ENCODE1; put(dstIdx++, RUN_LEN_ENCODE1 + (run >> 8)); }
      else { run -= RUN_LEN_ENCODE2; put(dstIdx++, 0xFF); put(dstIdx++, run >> 8); }
    }
    put(dstIdx, run);
    return dstIdx + 1;
  }
  // forward, RLT.java:62-231
  bool forward(Slice& input, Slice& output) override {
    if (input.length == 0) return true;
    if (!basicCheck(input, output)) return false;
    if (input.length < 16) return false;
    if (input.arr == output.arr) return false;
    const int count = input.length;
    if (output.length - output.index < getMaxEncodedLength(count)) return false;
    const u8* src = input.p(); u8* dst = output.p();
    const int srcCap = input.cap(), dstCap = output.cap();
    auto get = [&](int i) -> int { if (i < 0 || i >= srcCap) throw JavaException("AIOOBE in RLT.forward"); return src[i]; };
    auto put = [&](int i, int v) { if (i < 0 || i >= dstCap) throw JavaException("AIOOBE in RLT.forward"); dst[i] = (u8)v; };
    int dt = DT_UNDEFINED;
    bool findBestEscape = true;
    if (ctx != nullptr) {
      dt = ctx->dataType;
      if ((dt == DT_DNA) || (dt == DT_BASE64) || (dt == DT_UTF8)) return false;
      const int e = ctx->entropyType;
      if (e == E_NONE || e == E_ANS0 || e == E_HUFFMAN || e == E_RANGE) findBestEscape = false;
    }
    int escape = DEFAULT_ESCAPE;
    int srcIdx = input.index, dstIdx = output.index;
    const int srcEnd = srcIdx + count, srcEnd4 = srcEnd - 4;
    const int dstEnd = dstCap;
    if (findBestEscape) {
      int freqs[257];
      histogramOrder0(src, srcIdx, srcEnd, freqs, false);
      if (dt == DT_UNDEFINED) {
        dt = detectSimpleType(count, freqs);
        if ((ctx != nullptr) && (dt != DT_UNDEFINED)) ctx->dataType = dt;
        if ((dt == DT_DNA) || (dt == DT_BASE64) || (dt == DT_UTF8)) return false;
      }
      int minIdx = 0;
      if (freqs[minIdx] > 0) {
        for (int i = 1; i < 256; i++) {
          if (freqs[i] < freqs[minIdx]) { minIdx = i; if (freqs[i] == 0) break; }
        }
      }
      escape = minIdx;
    }
    bool res = true;
    int run = 0;
    int prev = get(srcIdx++);
    put(dstIdx++, escape);
    put(dstIdx++, prev);
    if (prev == escape) put(dstIdx++, 0);
    while (true) {
      if (prev == get(srcIdx)) {
        srcIdx++; run++;
        if (prev == get(srcIdx)) {
          srcIdx++; run++;
          if (prev == get(srcIdx)) {
            srcIdx++; run++;
            if (prev == get(srcIdx)) {
              srcIdx++; run++;
              if ((run < MAX_RUN4) && (srcIdx < srcEnd4)) continue;
            }
          }
        }
      }
      if (run > RUN_THRESHOLD) {
        if (dstIdx + 6 >= dstEnd) { res = false; break; }
        put(dstIdx++, prev);
        if (prev == escape) put(dstIdx++, 0);
        put(dstIdx++, escape);
        dstIdx = emitRunLength(dst, dstCap, dstIdx, run);
      } else if (prev != escape) {
        if (dstIdx + run >= dstEnd) { res = false; break; }
        while (run-- > 0) put(dstIdx++, prev);
      } else {
        if (dstIdx + 2 * run >= dstEnd) { res = false; break; }
        while (run-- > 0) { put(dstIdx++, escape); put(dstIdx++, 0); }
      }
      prev = get(srcIdx);
      srcIdx++;
      run = 1;
      if (srcIdx >= srcEnd4) break;
    }
    if (res) {
      if (prev != escape) {
        if (dstIdx + run < dstEnd) { while (run-- > 0) put(dstIdx++, prev); }
      } else {
        if (dstIdx + 2 * run < dstEnd) { while (run-- > 0) { put(dstIdx++, escape); put(dstIdx++, 0); } }
      }
      while ((srcIdx < srcEnd) && (dstIdx < dstEnd)) {
        if (get(srcIdx) == escape) {
          if (dstIdx + 2 >= dstEnd) { res = false; break; }
          put(dstIdx++, escape); put(dstIdx++, 0);
          srcIdx++;
          continue;
        }
        put(dstIdx++, get(srcIdx++));
      }
      res &= (srcIdx == srcEnd);
    }
    res &= ((dstIdx - output.index) < (srcIdx - input.index));
    input.index = srcIdx; output.index = dstIdx;
    return res;
  }
  // inverse, RLT.java:252-352
  bool inverse(Slice& input, Slice& output) override {
    if (input.length == 0) return true;
    if (!basicCheck(input, output)) return false;
    if (input.arr == output.arr) return false;
    const int count = input.length;
    int srcIdx = input.index, dstIdx = output.index;
    const u8* src = input.p(); u8* dst = output.p();
    const int srcCap = input.cap(), dstCap = output.cap();
    auto get = [&](int i) -> int { if (i < 0 || i >= srcCap) throw JavaException("AIOOBE in RLT.inverse"); return src[i]; };
    auto put = [&](int i, int v) { if (i < 0 || i >= dstCap) throw JavaException("AIOOBE in RLT.inverse"); dst[i] = (u8)v; };
    const int srcEnd = srcIdx + count;
    const int dstEnd = dstCap;
    bool res = true;
    const int escape = get(srcIdx++);
    if (get(srcIdx) == escape) {
      srcIdx++;
      if ((srcIdx < srcEnd) && (get(srcIdx) != 0)) return false;
      put(dstIdx++, escape);
      srcIdx++;
    }
    while (srcIdx < srcEnd) {
      if (get(srcIdx) != escape) {
        if (dstIdx >= dstEnd) break;
        put(dstIdx++, get(srcIdx++));
        continue;
      }
      srcIdx++;
      if (srcIdx >= srcEnd) { res = false; break; }
      if (dstIdx - 1 < 0) throw JavaException("AIOOBE in RLT.inverse");
      const int val = dst[dstIdx - 1];
      int run = get(srcIdx++);
      if (run == 0) {
        if (dstIdx >= dstEnd) break;
        put(dstIdx++, escape);
        continue;
      }
      if (run == 0xFF) {
        if (srcIdx >= srcEnd - 1) { res = false; break; }
        run = (get(srcIdx) << 8) | get(srcIdx + 1);
        srcIdx += 2;
        run += RUN_LEN_ENCODE2;
      } else if (run >= RUN_LEN_ENCODE1) {
        if (srcIdx >= srcEnd) { res = false; break; }
        run = ((run - RUN_LEN_ENCODE1) << 8) | get(srcIdx++);
        run += RUN_LEN_ENCODE1;
      }
      run += (RUN_THRESHOLD - 1);
      if ((dstIdx + run > dstEnd) || (run > MAX_RUN)) { res = false; break; }
      while (run-- > 0) put(dstIdx++, val);
    }
    res &= (srcIdx == srcEnd);
    input.index = srcIdx; output.index = dstIdx;
    return res;
  }
  int getMaxEncodedLength(int srcLen) override { return (srcLen <= 512) ? srcLen + 32 : srcLen; }        // RLT.java:355-357
};

// ---- LZPCodec (transform/LZCodec.java:973-1287; selected by ctx["lz"] == LZP_TYPE, LZCodec.java:57-58) ------------------
// bsVersion >= 4 everywhere on this path, so minMatch is 64 both ways (LZCodec.java:988-997, 1137).
struct LZP : Transform {
  enum { HASH_SEED = 0x7FEB352D, HASH_LOG = 16, HASH_SHIFT = 32 - HASH_LOG, MIN_MATCH64 = 64, MIN_BLOCK_LENGTH = 128, MATCH_FLAG = 0xFC };
  std::vector<i32> hashes;
  static int findMatch(const u8* src, int srcIdx, int ref, int maxMatch) {       // :1267-1280
    int bestLen = 0;
    while (bestLen + 8 <= maxMatch) {
      const u64 diff = le64(src + srcIdx + bestLen) ^ le64(src + ref + bestLen);
      if (diff != 0) { bestLen += (__builtin_ctzll(diff) >> 3); break; }
      bestLen += 8;
    }
    return bestLen;
  }
  // forward, :1003-1118
  bool forward(Slice& input, Slice& output) override {
    if (input.length == 0) return true;
    if (!basicCheck(input, output)) return false;
    const int count = input.length;
    if (output.length - output.index < getMaxEncodedLength(count)) return false;
    if (count < MIN_BLOCK_LENGTH) return false;
    hashes.assign(1 << HASH_LOG, 0);
    const int srcIdx0 = input.index, dstIdx0 = output.index;
    const u8* src = input.p(); u8* dst = output.p();
    const int dstCap = output.cap();
    auto put = [&](int i, int v) { if (i < 0 || i >= dstCap) throw JavaException("AIOOBE in LZP.forward"); dst[i] = (u8)v; };
    const int srcEnd = srcIdx0 + count;
    const int dstEnd = dstIdx0 + count - (count >> 6);
    int srcIdx = srcIdx0, dstIdx = dstIdx0;
    for (int k = 0; k < 4; k++) put(dstIdx + k, src[srcIdx + k]);
    u32 ctx = le32(src + srcIdx);
    srcIdx += 4; dstIdx += 4;
    const int minMatch = MIN_MATCH64;
    while ((srcIdx < srcEnd - minMatch) && (dstIdx < dstEnd)) {
      const u32 h = ((u32)HASH_SEED * ctx) >> HASH_SHIFT;
      const int ref = hashes[h];
      hashes[h] = srcIdx;
      int bestLen = 0;
      if ((ref != 0) && (le32(src + ref + minMatch - 4) == le32(src + srcIdx + minMatch - 4)))
        bestLen = findMatch(src, srcIdx, ref, srcEnd - srcIdx);
      if (bestLen < minMatch) {
        const int val = src[srcIdx];
        ctx = (ctx << 8) | (u32)val;
        put(dstIdx++, src[srcIdx++]);
        if ((ref != 0) && (val == MATCH_FLAG)) {
          if (dstIdx >= dstEnd) return false;
          put(dstIdx++, 0xFF);
        }
        continue;
      }
      srcIdx += bestLen;
      ctx = le32(src + srcIdx - 4);
      put(dstIdx++, MATCH_FLAG);
      bestLen -= minMatch;
      while (bestLen >= 254) {
        bestLen -= 254;
        put(dstIdx++, 0xFE);
        if (dstIdx >= dstEnd) break;
      }
      if (dstIdx >= dstEnd) return false;
      put(dstIdx++, bestLen);
    }
    while ((srcIdx < srcEnd) && (dstIdx < dstEnd)) {
      const u32 h = ((u32)HASH_SEED * ctx) >> HASH_SHIFT;
      const int ref = hashes[h];
      hashes[h] = srcIdx;
      const int val = src[srcIdx];
      ctx = (ctx << 8) | (u32)val;
      put(dstIdx++, src[srcIdx++]);
      if ((ref != 0) && (val == MATCH_FLAG)) {
        if (dstIdx >= dstEnd) return false;
        put(dstIdx++, 0xFF);
      }
    }
    input.index = srcIdx; output.index = dstIdx;
    return (srcIdx == srcIdx0 + count) && (dstIdx < dstEnd);
  }
  // inverse, :1121-1263
  bool inverse(Slice& input, Slice& output) override {
    if (input.length == 0) return true;
    if (!basicCheck(input, output)) return false;
    const int count = input.length;
    const u8* src = input.p(); u8* dst = output.p();
    const int srcCap = input.cap(), dstCap = output.cap();
    auto get = [&](int i) -> int { if (i < 0 || i >= srcCap) throw JavaException("AIOOBE in LZP.inverse"); return src[i]; };
    auto put = [&](int i, int v) { if (i < 0 || i >= dstCap) throw JavaException("AIOOBE in LZP.inverse"); dst[i] = (u8)v; };
    const int srcEnd = input.index + count;
    const int dstEnd = output.length;
    int srcIdx = input.index, dstIdx = output.index;
    const int minMatch = MIN_MATCH64;
    if (output.length - output.index < count) return false;
    hashes.assign(1 << HASH_LOG, 0);
    for (int k = 0; k < 4; k++) put(dstIdx + k, get(srcIdx + k));
    u32 ctx = le32(dst + dstIdx);
    srcIdx += 4; dstIdx += 4;
    while (srcIdx < srcEnd) {
      const u32 h = ((u32)HASH_SEED * ctx) >> HASH_SHIFT;
      const int ref = hashes[h];
      hashes[h] = dstIdx;
      if ((ref == 0) || (src[srcIdx] != MATCH_FLAG)) {
        if (dstIdx >= dstEnd) return false;
        put(dstIdx, src[srcIdx]);
        ctx = (ctx << 8) | (u32)dst[dstIdx];
        srcIdx++; dstIdx++;
        continue;
      }
      srcIdx++;
      if (srcIdx >= srcEnd) return false;
      if (src[srcIdx] == 0xFF) {
        if (dstIdx >= dstEnd) return false;
        put(dstIdx, MATCH_FLAG);
        ctx = (ctx << 8) | (u32)MATCH_FLAG;
        srcIdx++; dstIdx++;
        continue;
      }
      int mLen = minMatch;
      if (src[srcIdx] == 0xFE) {
        while ((srcIdx < srcEnd) && (src[srcIdx] == 0xFE)) { srcIdx++; mLen += 254; }
        if (srcIdx >= srcEnd) return false;
      }
      mLen += src[srcIdx++];
      if (dstIdx + mLen > dstEnd) return false;
      if (dstIdx + mLen > dstCap) throw JavaException("AIOOBE in LZP.inverse");
      for (int i = 0; i < mLen; i++) dst[dstIdx + i] = dst[ref + i];
      dstIdx += mLen;
      ctx = le32(dst + dstIdx - 4);
    }
    input.index = srcIdx; output.index = dstIdx;
    return srcIdx == srcEnd;
  }
  int getMaxEncodedLength(int srcLen) override { return (srcLen <= 1024) ? srcLen + 16 : srcLen + (srcLen / 64); }   // :1283-1285
};

// ---- TransformFactory.newFunctionToken (TransformFactory.java:273-351) -------------------------------
static inline std::unique_ptr<Transform> newTransform(Ctx& ctx, int type) {
  switch (type) {
    case T_NONE: return std::unique_ptr<Transform>(new NullTransform());
    case T_LZ: ctx.lzType = T_LZ; return std::unique_ptr<Transform>(new LZX(&ctx, false));
    case T_LZX: ctx.lzType = T_LZX; return std::unique_ptr<Transform>(new LZX(&ctx, true));
    case T_LZP: ctx.lzType = T_LZP; return std::unique_ptr<Transform>(new LZP());
    case T_RLT: return std::unique_ptr<Transform>(new RLT(&ctx));
    case T_BWTS: return std::unique_ptr<Transform>(new BWTS());
    case T_ROLZ: return std::unique_ptr<Transform>(new ROLZ1(&ctx));
    case T_ROLZX: ctx.rolzExtra = 1; return std::unique_ptr<Transform>(new ROLZ2(&ctx));
    case T_BWT: return std::unique_ptr<Transform>(new BWTBlockCodec(ctx));
    case T_RANK: ctx.sbrtMode = 2; return std::unique_ptr<Transform>(new SBRT(2));
    case T_MTFT: ctx.sbrtMode = 1; return std::unique_ptr<Transform>(new SBRT(1));
    case T_SRT: return std::unique_ptr<Transform>(new SRT());
    case T_ZRLT: return std::unique_ptr<Transform>(new ZRLT());
    default: throw JavaException("transform type outside the oracle's scope");
  }
}

// ---- Sequence (transform/Sequence.java) + TransformFactory.newFunction (:240-270) ---------------------
struct Sequence {
  enum { SKIP_MASK = 0xFF };
  std::vector<std::unique_ptr<Transform>> transforms;
  u8 skipFlags = 0;
  Sequence(Ctx& ctx, u64 functionType) {
    int nbtr = 0;
    for (int i = 0; i < 8; i++) if (((functionType >> (42 - 6 * i)) & 63) != T_NONE) nbtr++;
    if (nbtr == 0) nbtr = 1;
    const int len = nbtr;
    for (int i = 0; i < len; i++) {
      const int t = (int)((functionType >> (42 - 6 * i)) & 63);
      if ((t != T_NONE) || (i == 0)) transforms.push_back(newTransform(ctx, t));
    }
  }
  int getNbFunctions() const { return (int)transforms.size(); }
  int getMaxEncodedLength(int srcLength) {
    int requiredSize = srcLength;
    for (auto& t : transforms) requiredSize = std::max(requiredSize, t->getMaxEncodedLength(requiredSize));
    return requiredSize;
  }
  // forward, Sequence.java:56-127
  bool forward(Slice& src, Slice& dst) {
    int count = src.length;
    if ((count < 0) || (count > src.cap() - src.index)) return false;
    skipFlags = SKIP_MASK;
    if (src.length == 0) return true;
    if (!basicCheck(src, dst)) return false;
    const int blockSize = count;
    const int requiredSize = getMaxEncodedLength(count);
    Slice* sa[2] = {&src, &dst};
    Slice* sa1 = sa[0]; Slice* sa2 = sa[1];
    int saIdx = 0;
    for (size_t i = 0; i < transforms.size(); i++) {
      if (sa2->length < requiredSize) {
        sa2->length = requiredSize;
        if (sa2->cap() < sa2->length) sa2->arr->assign(sa2->length, 0);
      }
      const int savedIIdx = sa1->index, savedOIdx = sa2->index, savedLength = sa1->length;
      sa1->length = count;
      if (!transforms[i]->forward(*sa1, *sa2)) {
        if (sa1->arr != sa2->arr) memcpy(sa2->p() + savedOIdx, sa1->p() + savedIIdx, count);
        sa1->index = savedIIdx; sa2->index = savedOIdx; sa1->length = savedLength;
        continue;
      }
      skipFlags &= ~(1 << (7 - i));
      count = sa2->index - savedOIdx;
      sa1->index = savedIIdx; sa2->index = savedOIdx; sa1->length = savedLength;
      saIdx ^= 1;
      sa1 = sa[saIdx]; sa2 = sa[saIdx ^ 1];
    }
    if (saIdx != 1) {
      if (count > sa[1]->cap() - sa[1]->index) skipFlags = SKIP_MASK;
      else memmove(sa[1]->p() + sa[1]->index, sa[0]->p() + sa[0]->index, count);
    }
    src.index += blockSize;
    dst.index += count;
    return skipFlags != SKIP_MASK;
  }
  // inverse, Sequence.java:137-207
  bool inverse(Slice& src, Slice& dst) {
    if (src.length == 0) return true;
    if (!basicCheck(src, dst)) return false;
    int count = src.length;
    if (skipFlags == SKIP_MASK) {
      if (src.arr != dst.arr) memcpy(dst.p() + dst.index, src.p() + src.index, count);
      src.index += count; dst.index += count;
      return true;
    }
    const int blockSize = count;
    bool res = true;
    Slice* sa[2] = {&src, &dst};
    int saIdx = 0;
    for (int i = (int)transforms.size() - 1; i >= 0; i--) {
      if ((skipFlags & (1 << (7 - i))) != 0) continue;
      Slice* sa1 = sa[saIdx];
      saIdx ^= 1;
      Slice* sa2 = sa[saIdx];
      const int savedIIdx = sa1->index, savedOIdx = sa2->index, savedILen = sa1->length, savedOLen = sa2->length;
      sa1->length = count;
      sa2->length = dst.cap();
      if (sa2->cap() < sa2->length) sa2->arr->assign(sa2->length, 0);
      res = transforms[i]->inverse(*sa1, *sa2);
      count = sa2->index - savedOIdx;
      sa1->index = savedIIdx; sa2->index = savedOIdx; sa1->length = savedILen; sa2->length = savedOLen;
      if (!res) break;
    }
    if (res && (saIdx != 1)) {
      if (count > sa[1]->cap() - sa[1]->index) res = false;
      else memmove(sa[1]->p() + sa[1]->index, sa[0]->p() + sa[0]->index, count);
    }
    if (count > dst.length) return false;
    src.index += blockSize;
    dst.index += count;
    return res;
  }
};

}  // namespace kzo
