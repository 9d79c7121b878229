// ORACLE — TEST INFRASTRUCTURE ONLY (see kz_core.hpp header).  parity unpinned (no JVM, no goldens
// in the reference); every function cites the Java it restates.
//
// Entropy stage: EntropyUtils, ExpGolomb, Null, Huffman, ANS (order 0/1), FPAQ.
#pragma once
#include "kz_core.hpp"

namespace kzo {

// =================================================================================================
// EntropyUtils (entropy/EntropyUtils.java)
// =================================================================================================

// encodeAlphabet, EntropyUtils.java:38-75.  alphabet sorted ascending, alphabet.length == 256 here.
static inline int encodeAlphabet(BitWriter& obs, const int* alphabet, int alphabetLen, int count) {
  if ((alphabetLen & (alphabetLen - 1)) != 0) return -1;
  if (alphabetLen > 256 || count > alphabetLen) return -1;
  if (count == 0) { obs.writeBit(0); obs.writeBit(1); }
  else if (count == 256) { obs.writeBit(0); obs.writeBit(0); }
  else {
    obs.writeBit(1);
    u8 masks[32]; memset(masks, 0, 32);
    for (int i = 0; i < count; i++) masks[alphabet[i] >> 3] |= (u8)(1 << (alphabet[i] & 7));
    const int lastMask = alphabet[count - 1] >> 3;
    obs.writeBits((u64)lastMask, 5);
    for (int i = 0; i <= lastMask; i++) obs.writeBits(masks[i], 8);
  }
  return count;
}

// decodeAlphabet, EntropyUtils.java:86-122
static inline int decodeAlphabet(BitReader& ibs, int* alphabet) {
  if (ibs.readBit() == 0) {
    if (ibs.readBit() == 1) return 0;
    for (int i = 0; i < 256; i++) alphabet[i] = i;
    return 256;
  }
  const int lastMask = (int)ibs.readBits(5);
  int count = 0;
  for (int i = 0; i <= lastMask; i++) {
    const int mask = (int)ibs.readBits(8);
    for (int j = 0; j < 8; j++)
      if (mask & (1 << j)) alphabet[count++] = (i << 3) + j;
  }
  return count;
}

// normalizeFrequencies, EntropyUtils.java:141-250.  `alphabetLen` = alphabet.length in Java.
static inline int normalizeFrequencies(int* freqs, int* alphabet, int alphabetLen, int totalFreq, int scale) {
  if (alphabetLen > 256) throw JavaException("Invalid alphabet size parameter");
  if (scale < (1 << 8) || scale > (1 << 16)) throw JavaException("Invalid scale parameter");
  if (alphabetLen == 0 || totalFreq == 0) return 0;
  int alphabetSize = 0;
  if (totalFreq == scale) {                       // :156-163 (note: scans 256 entries regardless)
    for (int i = 0; i < 256; i++)
      if (freqs[i] != 0) alphabet[alphabetSize++] = i;
    return alphabetSize;
  }
  int sumScaledFreq = 0, sumFreq = 0, idxMax = 0;
  for (int i = 0; i < alphabetLen; i++) {         // :170-190
    alphabet[i] = 0;
    const int f = freqs[i];
    if (f == 0) continue;
    i64 sf = (i64)freqs[i] * scale;
    const int scaledFreq = (sf <= totalFreq) ? 1 : (int)((sf + ((i64)totalFreq >> 1)) / (i64)totalFreq);
    alphabet[alphabetSize++] = i;
    sumScaledFreq += scaledFreq;
    freqs[i] = scaledFreq;
    sumFreq += f;
    if (scaledFreq > freqs[idxMax]) idxMax = i;
    if (sumFreq >= totalFreq) break;
  }
  if (alphabetSize == 0) return 0;
  if (alphabetSize == 1) { freqs[alphabet[0]] = scale; return 1; }
  if (sumScaledFreq == scale) return alphabetSize;
  int delta = sumScaledFreq - scale;
  const int errThr = freqs[idxMax] >> 4;
  if (std::abs(delta) <= errThr) { freqs[idxMax] -= delta; return alphabetSize; }
  if (delta < 0) { delta += errThr; freqs[idxMax] += errThr; }
  else { delta -= errThr; freqs[idxMax] -= errThr; }
  const int inc = (delta > 0) ? -1 : 1;
  delta = std::abs(delta);
  int round = 0;
  while ((++round < 6) && (delta > 0)) {
    int adjustments = 0;
    for (int i = 0; i < alphabetSize; i++) {
      const int idx = alphabet[i];
      if (freqs[idx] <= 2) continue;
      freqs[idx] += inc;
      adjustments++;
      delta--;
      if (delta == 0) break;
    }
    if (adjustments == 0) break;
  }
  freqs[idxMax] = std::max(freqs[idxMax] - delta, 1);
  return alphabetSize;
}

// writeVarInt / readVarInt, EntropyUtils.java:259-300
static inline int writeVarInt(BitWriter& bs, i32 value) {
  int res = 0;
  u32 v = (u32)value;
  if (value >= 128 || value < 0) {
    bs.writeBits(0x80 | (v & 0x7F), 8); v >>= 7; res++;
    while (v >= 128) { bs.writeBits(0x80 | (v & 0x7F), 8); v >>= 7; res++; }
  }
  bs.writeBits(v, 8);
  return res;
}
static inline i32 readVarInt(BitReader& bs) {
  int value = (int)bs.readBits(8);
  u32 res = value & 0x7F;
  int shift = 7;
  while (value >= 128) {
    value = (int)bs.readBits(8);
    res |= ((u32)(value & 0x7F) << shift);
    if (shift == 28) break;
    shift += 7;
  }
  return (i32)res;
}

// =================================================================================================
// Global.computeHistogramOrder0/1 (Global.java:274-427) — semantic restatement (the Java unrolls 4x)
// =================================================================================================
static inline void histogramOrder0(const u8* block, int start, int end, int* freqs, bool withTotal) {
  for (int i = 0; i < 256; i++) freqs[i] = 0;
  if (withTotal) freqs[256] = end - start;
  for (int i = start; i < end; i++) freqs[block[i]]++;
}
// freqs is [256][257]; Global.java:341-427.  Quarter-interleaved walk when end-start >= 32.
static inline void histogramOrder1(const u8* block, int start, int end, int (*freqs)[257], bool withTotal) {
  const int quarter = (end - start) >> 2;
  int n0 = start, n1 = start + quarter, n2 = start + 2 * quarter, n3 = start + 3 * quarter;
  if (end - start < 32) {
    int prv = 0;
    for (int i = start; i < end; i++) {
      freqs[prv][block[i]]++;
      if (withTotal) freqs[prv][256]++;
      prv = block[i];
    }
    return;
  }
  int prv0 = 0, prv1 = block[n1 - 1], prv2 = block[n2 - 1], prv3 = block[n3 - 1];
  for (; n0 < start + quarter; n0++, n1++, n2++, n3++) {
    const int c0 = block[n0], c1 = block[n1], c2 = block[n2], c3 = block[n3];
    freqs[prv0][c0]++; freqs[prv1][c1]++; freqs[prv2][c2]++; freqs[prv3][c3]++;
    if (withTotal) { freqs[prv0][256]++; freqs[prv1][256]++; freqs[prv2][256]++; freqs[prv3][256]++; }
    prv0 = c0; prv1 = c1; prv2 = c2; prv3 = c3;
  }
  for (; n3 < end; n3++) {
    freqs[prv3][block[n3]]++;
    if (withTotal) freqs[prv3][256]++;
    prv3 = block[n3];
  }
}

// =================================================================================================
// ExpGolomb (signed flavour used by Huffman).  Encoder: ExpGolombEncoder.java:123-132 is a 256-entry
// table of (nbits<<9)|bits; this restates the code the table holds, checked
// against the decoder (ExpGolombDecoder.java:41-60) and the KATs +1 -> 0100, -1 -> 0101.
// =================================================================================================
static inline void expGolombEncodeSigned(BitWriter& bs, int8_t val) {
  if (val == 0) { bs.writeBit(1); return; }
  const int a = (val < 0) ? -(int)val : (int)val;       // 1..128
  const int sgn = (val < 0) ? 1 : 0;
  const int lg = log2i((u32)(a + 1));                   // >= 1
  // '0', (lg-1) zeros, '1', then lg+1 bits of ((a+1-2^lg)<<1 | sign)
  const u32 tail = (((u32)(a + 1 - (1 << lg))) << 1) | (u32)sgn;
  const int nbits = 2 * lg + 2;
  const u32 bits = (1u << (lg + 1)) | tail;
  bs.writeBits(bits, nbits);
}
static inline int8_t expGolombDecodeSigned(BitReader& bs) {
  if (bs.readBit() == 1) return 0;
  int lg = 1;
  while (bs.readBit() == 0) lg++;
  i64 res = (i64)bs.readBits(lg + 1);
  const i64 sgn = res & 1;
  res = (i64)((u64)res >> 1) + (1 << lg) - 1;
  return (int8_t)((res - sgn) ^ -sgn);
}

// =================================================================================================
// Null entropy codec (NullEntropyEncoder.java / NullEntropyDecoder.java)
// =================================================================================================
static inline int nullEncode(BitWriter& bs, const u8* block, int blkptr, int count) {
  bs.writeBytesBits(block + blkptr, (i64)count * 8);
  return count;
}
static inline int nullDecode(BitReader& bs, u8* block, int blkptr, int count) {
  bs.readBytesBits(block + blkptr, (i64)count * 8);
  return count;
}

// =================================================================================================
// Huffman (HuffmanCommon.java, HuffmanEncoder.java, HuffmanDecoder.java)
// =================================================================================================
namespace huff {
enum { LOG_MAX_CHUNK_SIZE = 14, MAX_CHUNK_SIZE = 1 << 14, MAX_SYMBOL_SIZE_V4 = 12,
       BUFFER_SIZE = (14 << 8) + 256 };

// HuffmanCommon.generateCanonicalCodes, HuffmanCommon.java:71-111
static inline int generateCanonicalCodes(const short* sizes, int* codes, int* symbols, int count, int maxSymbolSize) {
  if (count > 1) {
    std::vector<u8> buf(BUFFER_SIZE, 0);
    for (int i = 0; i < count; i++) {
      const int s = symbols[i];
      if (((s & 0xFF) != s) || (sizes[s] > maxSymbolSize)) return -1;
      buf[((sizes[s] - 1) << 8) | s] = 1;
    }
    int n = 0;
    for (int i = 0; i < BUFFER_SIZE; i++) {
      if (buf[i] == 0) continue;
      symbols[n++] = i & 0xFF;
      if (n == count) break;
    }
  }
  int code = 0;
  int curLen = sizes[symbols[0]];
  for (int i = 0; i < count; i++) {
    const int s = symbols[i];
    code <<= (sizes[s] - curLen);
    curLen = sizes[s];
    codes[s] = code;
    code++;
  }
  return count;
}

// HuffmanEncoder.computeInPlaceSizesPhase1/2, HuffmanEncoder.java:317-376
static inline void phase1(int* data, int n) {
  for (int s = 0, r = 0, t = 0; t < n - 1; t++) {
    int sum = 0;
    for (int i = 0; i < 2; i++) {
      if ((s >= n) || ((r < t) && (data[r] < data[s]))) { sum += data[r]; data[r] = t; r++; continue; }
      sum += data[s];
      if (s > t) data[s] = 0;
      s++;
    }
    data[t] = sum;
  }
}
static inline int phase2(int* data, int n) {
  if (n < 2) return 0;
  int levelTop = n - 2, depth = 1, i = n, totalNodesAtLevel = 2;
  while (i > 0) {
    int k = levelTop;
    while ((k > 0) && (data[k - 1] >= levelTop)) k--;
    const int internalNodesAtLevel = levelTop - k;
    const int leavesAtLevel = totalNodesAtLevel - internalNodesAtLevel;
    for (int j = 0; j < leavesAtLevel; j++) data[--i] = depth;
    totalNodesAtLevel = internalNodesAtLevel << 1;
    levelTop = k;
    depth++;
  }
  return depth - 1;
}

// HuffmanEncoder.computeCodeLengths, HuffmanEncoder.java:285-308
static inline int computeCodeLengths(short* sizes, int* ranks, int count) {
  std::sort(ranks, ranks + count);
  int freqs[256]; memset(freqs, 0, sizeof(freqs));
  for (int i = 0; i < count; i++) {
    freqs[i] = (int)((u32)ranks[i] >> 8);
    ranks[i] &= 0xFF;
    if (freqs[i] == 0) return 0;
  }
  phase1(freqs, count);
  const int maxCodeLen = phase2(freqs, count);
  for (int i = 0; i < count; i++) sizes[ranks[i]] = (short)freqs[i];
  return maxCodeLen;
}

// HuffmanEncoder.limitCodeLengths, HuffmanEncoder.java:191-273
static inline int limitCodeLengths(const int* alphabet, int* freqs, short* sizes, int* ranks, int count) {
  int n = 0, debt = 0;
  // NB: the Java loop has no n < count guard; sizes[ranks[n]] for n >= count reads ranks[n] == 0 slots
  // of a 256-int array, harmless there; keep the read in bounds here.
  while (n < 256 && sizes[ranks[n]] >= MAX_SYMBOL_SIZE_V4) {
    debt += (sizes[ranks[n]] - MAX_SYMBOL_SIZE_V4);
    sizes[ranks[n]] = MAX_SYMBOL_SIZE_V4;
    n++;
  }
  std::vector<std::vector<int>> ll(6);   // FIFO lists; head index tracked separately
  size_t head[6] = {0, 0, 0, 0, 0, 0};
  while (n < count) {
    const int idx = MAX_SYMBOL_SIZE_V4 - 1 - sizes[ranks[n]];
    if ((idx >= 6) || (debt < (1 << idx))) break;
    ll[idx].push_back(ranks[n]);
    n++;
  }
  int idx = 5;
  while ((debt > 0) && (idx >= 0)) {
    if ((head[idx] >= ll[idx].size()) || (debt < (1 << idx))) { idx--; continue; }
    const int r = ll[idx][head[idx]++];
    sizes[r]++;
    debt -= (1 << idx);
  }
  idx = 0;
  while ((debt > 0) && (idx < 6)) {
    if (head[idx] >= ll[idx].size()) { idx++; continue; }
    const int r = ll[idx][head[idx]++];
    sizes[r]++;
    debt -= (1 << idx);
  }
  if (debt > 0) {
    std::vector<int> f(count), symbols(count);
    int totalFreq = 0;
    for (int i = 0; i < count; i++) { f[i] = freqs[alphabet[i]]; totalFreq += f[i]; }
    // alphabet.length == count in this call (may be < 256): loop bound of the scaling pass
    // NB: the totalFreq == scale shortcut scans 256 entries of f in Java (AIOOBE if count < 256 and
    // totalFreq == 2048); restated with a bounds-checked scan.
    if (totalFreq == (MAX_CHUNK_SIZE >> 3) && count < 256) throw JavaException("AIOOBE in normalizeFrequencies shortcut");
    normalizeFrequencies(f.data(), symbols.data(), count, totalFreq, MAX_CHUNK_SIZE >> 3);
    for (int i = 0; i < count; i++) {
      freqs[alphabet[i]] = f[i];
      ranks[i] = (f[i] << 8) | alphabet[i];
    }
    return computeCodeLengths(sizes, ranks, count);
  }
  return MAX_SYMBOL_SIZE_V4;
}

struct Encoder {
  BitWriter& bs;
  int alphabet[256];
  int codes[256];
  int chunkSize = MAX_CHUNK_SIZE;
  explicit Encoder(BitWriter& b) : bs(b) { for (int i = 0; i < 256; i++) codes[i] = i; }

  // HuffmanEncoder.updateFrequencies, HuffmanEncoder.java:103-178
  int updateFrequencies(int* freqs) {
    int count = 0;
    short sizes[256]; memset(sizes, 0, sizeof(sizes));
    for (int i = 0; i < 256; i++) {
      codes[i] = 0;
      if (freqs[i] > 0) alphabet[count++] = i;
    }
    encodeAlphabet(bs, alphabet, 256, count);
    if (count == 0) return 0;
    if (count == 1) {
      codes[alphabet[0]] = 1 << 24;
      sizes[alphabet[0]] = 1;
    } else {
      int ranks[256]; memset(ranks, 0, sizeof(ranks));
      for (int i = 0; i < count; i++) ranks[i] = (freqs[alphabet[i]] << 8) | alphabet[i];
      int maxCodeLen = computeCodeLengths(sizes, ranks, count);
      if (maxCodeLen == 0) throw JavaException("Could not generate Huffman codes: invalid code length 0");
      if (maxCodeLen > MAX_SYMBOL_SIZE_V4) {
        maxCodeLen = limitCodeLengths(alphabet, freqs, sizes, ranks, count);
        if (maxCodeLen == 0) throw JavaException("Could not generate Huffman codes: invalid code length 0");
      }
      if (maxCodeLen > MAX_SYMBOL_SIZE_V4) {
        int n = 0;
        for (int i = 0; i < count; i++) { codes[alphabet[i]] = n; sizes[alphabet[i]] = 8; n++; }
      } else {
        generateCanonicalCodes(sizes, codes, ranks, count, MAX_SYMBOL_SIZE_V4);
      }
    }
    short prevSize = 2;
    for (int i = 0; i < count; i++) {
      const int s = alphabet[i];
      const short currSize = sizes[s];
      codes[s] |= (currSize << 24);
      expGolombEncodeSigned(bs, (int8_t)(currSize - prevSize));
      prevSize = currSize;
    }
    return count;
  }

  // HuffmanEncoder.encodeChunk, HuffmanEncoder.java:419-493.  The Java packs each fragment through a
  // 64-bit shift register into a byte buffer; the emitted bits are the plain concatenation of the
  // codes, which is what is written here.
  void encodeChunk(const u8* block, int blkptr, int count) {
    const int szFrag = count / 4;
    BitWriter frag[4];
    int nbBits[4];
    for (int j = 0; j < 4; j++) {
      const int start = blkptr + j * szFrag;
      for (int i = start; i < start + szFrag; i++) {
        const int code = codes[block[i]];
        frag[j].writeBits((u64)(code & 0xFFFFFF), (int)((u32)code >> 24));
      }
      nbBits[j] = (int)frag[j].written();
      frag[j].close();
    }
    for (int j = 0; j < 4; j++) writeVarInt(bs, nbBits[j]);
    for (int j = 0; j < 4; j++) bs.writeBytesBits(frag[j].buf.data(), nbBits[j]);
    for (int i = 4 * szFrag; i < count; i++) bs.writeBits(block[blkptr + i], 8);
  }

  // HuffmanEncoder.encode, HuffmanEncoder.java:380-416
  int encode(const u8* block, int blkptr, int count) {
    if (count == 0) return 0;
    const int end = blkptr + count;
    int startChunk = blkptr;
    int freqs[257];
    while (startChunk < end) {
      const int sizeChunk = std::min(chunkSize, end - startChunk);
      if (sizeChunk < 32) {
        bs.writeBytesBits(block + startChunk, 8 * (i64)sizeChunk);
      } else {
        histogramOrder0(block, startChunk, startChunk + sizeChunk, freqs, false);
        if (updateFrequencies(freqs) > 1) encodeChunk(block, startChunk, sizeChunk);
      }
      startChunk += sizeChunk;
    }
    return count;
  }
};

struct Decoder {
  BitReader& bs;
  int codes[256]; int alphabet[256]; short sizes[256];
  short table[1 << MAX_SYMBOL_SIZE_V4];
  std::vector<u8> buffer;
  int chunkSize = MAX_CHUNK_SIZE;
  explicit Decoder(BitReader& b) : bs(b) { for (int i = 0; i < 256; i++) { sizes[i] = 8; codes[i] = i; } }

  // HuffmanDecoder.readLengths, HuffmanDecoder.java:115-154
  int readLengths() {
    const int count = decodeAlphabet(bs, alphabet);
    if (count == 0) return 0;
    int curSize = 2;
    for (int i = 0; i < count; i++) {
      const int s = alphabet[i];
      if ((s & 0xFF) != s) throw BitStreamError("incorrect Huffman symbol");
      codes[s] = 0;
      curSize += expGolombDecodeSigned(bs);
      if ((curSize <= 0) || (curSize > MAX_SYMBOL_SIZE_V4)) throw BitStreamError("incorrect size for Huffman symbol");
      sizes[s] = (short)curSize;
    }
    if (generateCanonicalCodes(sizes, codes, alphabet, count, MAX_SYMBOL_SIZE_V4) < 0)
      throw BitStreamError("max code length exceeded");
    return count;
  }

  // HuffmanDecoder.buildDecodingTables, HuffmanDecoder.java:162-191
  void buildDecodingTables(int count) {
    for (int i = 0; i < (1 << MAX_SYMBOL_SIZE_V4); i++) table[i] = 7;
    int length = 0;
    const int shift = MAX_SYMBOL_SIZE_V4;
    for (int i = 0; i < count; i++) {
      const int s = alphabet[i];
      if (sizes[s] > length) length = sizes[s];
      const short val = (short)((sizes[s] << 8) | s);
      const int code = codes[s];
      int idx = code << (shift - length);
      const int end = idx + (1 << (shift - length));
      while (idx < end) table[idx++] = val;
    }
  }

  static inline u64 be64(const u8* p) { return __builtin_bswap64(le64(p)); }

  // HuffmanDecoder.decodeChunk, HuffmanDecoder.java:404-587 (4 streams kept in arrays, same state
  // machine: 56-bit refill, 4 symbols per refill, final consumed-bits check)
  bool decodeChunk(u8* block, int blkptr, int count) {
    int szBits[4];
    for (int j = 0; j < 4; j++) szBits[j] = readVarInt(bs);
    for (int j = 0; j < 4; j++) if (szBits[j] < 0) return false;
    std::fill(buffer.begin(), buffer.end(), 0);
    const int stride = (int)buffer.size() / 4;
    int base[4], idx[4];
    for (int j = 0; j < 4; j++) { base[j] = j * stride; idx[j] = base[j]; }
    for (int j = 0; j < 4; j++) {
      if ((szBits[j] >> 3) > (int)buffer.size() - idx[j]) throw JavaException("Invalid bit count");  // readBits(byte[]) guard
      bs.readBytesBits(buffer.data() + idx[j], szBits[j]);
    }
    u64 state[4] = {0, 0, 0, 0}; int bits[4] = {0, 0, 0, 0};
    const int szFrag = count / 4;
    int blockIdx[4]; for (int j = 0; j < 4; j++) blockIdx[j] = blkptr + j * szFrag;
    const int MASK = (1 << MAX_SYMBOL_SIZE_V4) - 1;
    int n = 0; int bsh[4];
    auto refill = [&](int j) {
      const int shift = (56 - bits[j]) & -8;
      // Java: (state << shift) | (readLong64 >>> (63 - shift) >>> 1); shift in {0,8,..,56}
      if (idx[j] + 8 > (int)buffer.size()) throw JavaException("AIOOBE in Huffman refill");
      const u64 w = be64(buffer.data() + idx[j]);
      state[j] = ((shift == 0) ? state[j] : (state[j] << shift)) | ((w >> (63 - shift)) >> 1);
      bsh[j] = bits[j] + shift - MAX_SYMBOL_SIZE_V4;
      idx[j] += (shift >> 3);
    };
    auto sym = [&](int j) -> int {
      const int val = table[(int)(u32)((i64)state[j] >> (bsh[j] & 63)) & MASK];
      bsh[j] -= (int)((u32)(int)val >> 8);   // val is a non-negative short
      return val;
    };
    while (n < szFrag - 4) {
      for (int j = 0; j < 4; j++) refill(j);
      int v[4][4];
      for (int k = 0; k < 4; k++) for (int j = 0; j < 4; j++) v[j][k] = sym(j);
      for (int j = 0; j < 4; j++) {
        bits[j] = bsh[j] + MAX_SYMBOL_SIZE_V4;
        for (int k = 0; k < 4; k++) block[blockIdx[j] + k] = (u8)v[j][k];
        blockIdx[j] += 4;
      }
      n += 4;
    }
    for (int j = 0; j < 4; j++) refill(j);
    while (n < szFrag) {
      for (int j = 0; j < 4; j++) block[blockIdx[j]++] = (u8)sym(j);
      n++;
    }
    for (int i = 4 * szFrag; i < count; i++) block[blkptr + i] = (u8)bs.readBits(8);
    for (int j = 0; j < 4; j++)
      if ((((idx[j] - base[j]) << 3) - (bsh[j] + MAX_SYMBOL_SIZE_V4)) != szBits[j]) return false;
    return true;
  }

  // HuffmanDecoder.decodeV6, HuffmanDecoder.java:353-390 (bsVersion >= 6)
  int decode(u8* block, int blkptr, int count) {
    if (count == 0) return 0;
    if ((int)buffer.size() < 2 * chunkSize) buffer.assign(2 * chunkSize, 0);
    int startChunk = blkptr;
    const int end = blkptr + count;
    while (startChunk < end) {
      const int sizeChunk = std::min(chunkSize, end - startChunk);
      const int endChunk = startChunk + sizeChunk;
      if (sizeChunk < 32) {
        bs.readBytesBits(block + startChunk, 8 * (i64)sizeChunk);
      } else {
        const int alphabetSize = readLengths();
        if (alphabetSize <= 0) return startChunk - blkptr;
        if (alphabetSize == 1) {
          for (int i = startChunk; i < endChunk; i++) block[i] = (u8)alphabet[0];
        } else {
          buildDecodingTables(alphabetSize);
          if (!decodeChunk(block, startChunk, endChunk - startChunk)) return startChunk - blkptr;
        }
      }
      startChunk = endChunk;
    }
    return count;
  }
};
}  // namespace huff

// =================================================================================================
// ANS range codec (ANSRangeEncoder.java, ANSRangeDecoder.java)
// =================================================================================================
namespace ans {
enum { ANS_TOP = 1 << 15, DEFAULT_ANS0_CHUNK_SIZE = 16384, DEFAULT_LOG_RANGE = 12, MAX_CHUNK_SIZE = 1 << 27 };

// ANSRangeEncoder.Symbol, ANSRangeEncoder.java:466-497
struct EncSymbol {
  i32 xMax = 0, bias = 0, cmplFreq = 0, invShift = 0; u64 invFreq = 0;
  void reset(int cumFreq, int freq, int logRange) {
    if (freq >= (1 << logRange)) freq = (1 << logRange) - 1;
    xMax = jmul((i32)(((u32)ANS_TOP >> logRange) << 16), freq);
    cmplFreq = (1 << logRange) - freq;
    if (freq < 2) {
      invFreq = 0xFFFFFFFFULL; invShift = 32; bias = cumFreq + (1 << logRange) - 1;
    } else {
      int shift = 0;
      while (freq > (1 << shift)) shift++;
      invFreq = (((1ULL << (shift + 31)) + (u64)freq - 1) / (u64)freq) & 0xFFFFFFFFULL;
      invShift = 32 + shift - 1;
      bias = cumFreq;
    }
  }
};

struct Encoder {
  BitWriter& bs;
  int order, logRange, chunkSize;
  std::vector<int> freqsStore;       // [dim][257]
  std::vector<EncSymbol> symbols;    // [dim][256]
  std::vector<u8> buffer;
  int dim;
  int (*freqs)[257];

  // ANSRangeEncoder(bs, order, chunkSize, logRange), ANSRangeEncoder.java:77-115;
  // the (bs, ctx, order) ctor (:117-140) = chunkSize 16384, logRange 12.
  Encoder(BitWriter& b, int order_, int chunkSize_ = DEFAULT_ANS0_CHUNK_SIZE, int logRange_ = DEFAULT_LOG_RANGE)
      : bs(b), order(order_) {
    dim = 255 * order + 1;
    freqsStore.assign((size_t)dim * 257, 0);
    freqs = reinterpret_cast<int (*)[257]>(freqsStore.data());
    symbols.resize((size_t)dim * 256);
    logRange = (order == 0) ? logRange_ : std::max(logRange_ - 1, 8);
    chunkSize = (int)std::min(((i64)chunkSize_) << (8 * order), (i64)MAX_CHUNK_SIZE);
  }

  // encodeHeader, ANSRangeEncoder.java:211-252
  bool encodeHeader(int alphabetSize, const int* alphabet, const int* frequencies, int lr) {
    const int encoded = encodeAlphabet(bs, alphabet, 256, alphabetSize);
    if (encoded < 0) return false;
    if (encoded <= 1) return true;
    const int chkSize = (alphabetSize >= 64) ? 8 : 6;
    int llr = 3;
    while ((1 << llr) <= lr) llr++;
    for (int i = 1; i < alphabetSize; i += chkSize) {
      int max = frequencies[alphabet[i]] - 1;
      int logMax = 0;
      const int endj = (i + chkSize < alphabetSize) ? i + chkSize : alphabetSize;
      for (int j = i + 1; j < endj; j++)
        if (frequencies[alphabet[j]] - 1 > max) max = frequencies[alphabet[j]] - 1;
      while ((1 << logMax) <= max) logMax++;
      bs.writeBits((u64)logMax, llr);
      if (logMax == 0) continue;
      for (int j = i; j < endj; j++) bs.writeBits((u64)(frequencies[alphabet[j]] - 1), logMax);
    }
    return true;
  }

  // updateFrequencies, ANSRangeEncoder.java:171-200
  int updateFrequencies(int lr) {
    int res = 0;
    bs.writeBits((u64)(lr - 8), 3);
    int alphabet[256];
    for (int k = 0; k < dim; k++) {
      int* f = freqs[k];
      EncSymbol* symb = &symbols[(size_t)k * 256];
      const int alphabetSize = normalizeFrequencies(f, alphabet, 256, f[256], 1 << lr);
      if (alphabetSize > 0) {
        int sum = 0;
        for (int i = 0, count = 0; (i < 256) && (count < alphabetSize); i++) {
          if (f[i] == 0) continue;
          symb[i].reset(sum, f[i], lr);
          sum += f[i];
          count++;
        }
      }
      encodeHeader(alphabetSize, alphabet, f, lr);
      res += alphabetSize;
    }
    return res;
  }

  // rebuildStatistics, ANSRangeEncoder.java:419-449
  int rebuildStatistics(const u8* block, int start, int end, int lr) {
    std::fill(freqsStore.begin(), freqsStore.end(), 0);
    if (order == 0) {
      histogramOrder0(block, start, end, freqs[0], true);
    } else {
      const int quarter = (end - start) >> 2;
      if (quarter == 0) {
        histogramOrder1(block, start, end, freqs, true);
      } else {
        for (int q = 0; q < 4; q++) histogramOrder1(block, start + q * quarter, start + (q + 1) * quarter, freqs, true);
      }
    }
    return updateFrequencies(lr);
  }

  // encodeSymbol, ANSRangeEncoder.java:315-328
  inline i32 encodeSymbol(int& idx, i32 st, const EncSymbol& sym) {
    const int x = (st >= sym.xMax) ? 1 : 0;
    if (idx - x < 0) throw JavaException("AIOOBE in ANS encodeSymbol");
    buffer[idx] = (u8)st; idx -= x;
    buffer[idx] = (u8)(st >> 8); idx -= x;
    st >>= (-x & 16);
    const i32 q = (i32)(((i64)st * (i64)sym.invFreq) >> sym.invShift);
    return jadd(jadd(st, sym.bias), jmul(q, sym.cmplFreq));
  }

  // encodeChunk, ANSRangeEncoder.java:337-407
  void encodeChunk(const u8* block, int start, int end) {
    i32 st0 = ANS_TOP, st1 = ANS_TOP, st2 = ANS_TOP, st3 = ANS_TOP;
    int n = (int)buffer.size() - 1;
    const int end4 = start + ((end - start) & -4);
    for (int i = end - 1; i >= end4; i--) buffer[n--] = block[i];
    int idx = n;
    if (order == 0) {
      const EncSymbol* symb = &symbols[0];
      for (int i = end4 - 1; i > start; i -= 4) {
        st0 = encodeSymbol(idx, st0, symb[block[i]]);
        st1 = encodeSymbol(idx, st1, symb[block[i - 1]]);
        st2 = encodeSymbol(idx, st2, symb[block[i - 2]]);
        st3 = encodeSymbol(idx, st3, symb[block[i - 3]]);
      }
    } else {
      const int quarter = (end4 - start) >> 2;
      int i0 = start + 1 * quarter - 2, i1 = start + 2 * quarter - 2, i2 = start + 3 * quarter - 2, i3 = end4 - 2;
      if (i0 + 1 < 0) throw JavaException("AIOOBE in ANS1 encodeChunk");
      int prv0 = block[i0 + 1], prv1 = block[i1 + 1], prv2 = block[i2 + 1], prv3 = block[i3 + 1];
      for (; i0 >= start; i0--, i1--, i2--, i3--) {
        const int cur0 = block[i0]; st0 = encodeSymbol(idx, st0, symbols[(size_t)cur0 * 256 + prv0]);
        const int cur1 = block[i1]; st1 = encodeSymbol(idx, st1, symbols[(size_t)cur1 * 256 + prv1]);
        const int cur2 = block[i2]; st2 = encodeSymbol(idx, st2, symbols[(size_t)cur2 * 256 + prv2]);
        const int cur3 = block[i3]; st3 = encodeSymbol(idx, st3, symbols[(size_t)cur3 * 256 + prv3]);
        prv0 = cur0; prv1 = cur1; prv2 = cur2; prv3 = cur3;
      }
      st0 = encodeSymbol(idx, st0, symbols[prv0]);
      st1 = encodeSymbol(idx, st1, symbols[prv1]);
      st2 = encodeSymbol(idx, st2, symbols[prv2]);
      st3 = encodeSymbol(idx, st3, symbols[prv3]);
    }
    n = idx; n++;
    writeVarInt(bs, (i32)buffer.size() - n);
    bs.writeBits((u64)(u32)st0, 32); bs.writeBits((u64)(u32)st1, 32);
    bs.writeBits((u64)(u32)st2, 32); bs.writeBits((u64)(u32)st3, 32);
    if ((int)buffer.size() != n) bs.writeBytesBits(buffer.data() + n, 8 * (i64)((int)buffer.size() - n));
  }

  // encode, ANSRangeEncoder.java:263-305
  int encode(const u8* block, int blkptr, int count) {
    if (count <= 32) { bs.writeBytesBits(block + blkptr, 8 * (i64)count); return count; }
    const int end = blkptr + count;
    int sizeChunk = chunkSize;
    int startChunk = blkptr;
    for (auto& s : symbols) s = EncSymbol();
    const int size = std::max(std::min(sizeChunk + (sizeChunk >> 3), 2 * count), 65536);
    if ((int)buffer.size() < size) buffer.assign(size, 0);
    while (startChunk < end) {
      const int endChunk = std::min(startChunk + sizeChunk, end);
      const int alphabetSize = rebuildStatistics(block, startChunk, endChunk, logRange);
      if ((alphabetSize <= 1) && (order == 0)) { startChunk = endChunk; continue; }
      encodeChunk(block, startChunk, endChunk);
      startChunk = endChunk;
    }
    return count;
  }
};

struct DecSymbol { i32 cumFreq = 0, freq = 0; };

struct Decoder {
  BitReader& bs;
  int order, chunkSize, logRange = DEFAULT_LOG_RANGE, dim;
  std::vector<int> freqsStore;                 // [dim][256]
  std::vector<std::vector<u8>> f2s;            // [dim][scale]
  std::vector<DecSymbol> symbols;              // [dim][256]
  std::vector<u8> buffer;

  // ANSRangeDecoder(bs, ctx, order[, chunkSize]), ANSRangeDecoder.java:94-146 (bsVersion >= 4)
  Decoder(BitReader& b, int order_, int chunkSize_ = DEFAULT_ANS0_CHUNK_SIZE) : bs(b), order(order_) {
    dim = 255 * order + 1;
    chunkSize = (int)std::min(((i64)chunkSize_) << (8 * order), (i64)MAX_CHUNK_SIZE);
    freqsStore.assign((size_t)dim * 256, 0);
    f2s.resize(dim);
    symbols.resize((size_t)dim * 256);
  }

  // decodeHeader, ANSRangeDecoder.java:452-544
  int decodeHeader(int* alphabet) {
    logRange = (int)(8 + bs.readBits(3));
    if (logRange < 8 || logRange > 15) throw BitStreamError("Invalid bitstream: range");
    int res = 0;
    const int scale = 1 << logRange;
    for (int k = 0; k < dim; k++) {
      int alphabetSize = decodeAlphabet(bs, alphabet);
      if (alphabetSize == 0) continue;
      int llr = 3;
      while ((1 << llr) <= logRange) llr++;
      int* f = &freqsStore[(size_t)k * 256];
      if (alphabetSize != 256) for (int i = 255; i >= 0; i--) f[i] = 0;
      if ((int)f2s[k].size() < scale) f2s[k].assign(scale, 0);
      const int chkSize = (alphabetSize >= 64) ? 8 : 6;
      int sum = 0;
      for (int i = 1; i < alphabetSize; i += chkSize) {
        const int logMax = (int)bs.readBits(llr);
        if ((1 << logMax) > scale) throw BitStreamError("incorrect frequency size");
        const int endj = (i + chkSize < alphabetSize) ? i + chkSize : alphabetSize;
        for (int j = i; j < endj; j++) {
          const int freq = (logMax == 0) ? 1 : (int)(1 + bs.readBits(logMax));
          if (freq <= 0 || freq >= scale) throw BitStreamError("incorrect frequency");
          f[alphabet[j]] = freq;
          sum += freq;
        }
      }
      if (scale <= sum) throw BitStreamError("incorrect frequency (first)");
      f[alphabet[0]] = scale - sum;
      sum = 0;
      DecSymbol* symb = &symbols[(size_t)k * 256];
      u8* freq2sym = f2s[k].data();
      for (int i = 0; i < 256; i++) {
        if (f[i] == 0) continue;
        for (int j = f[i] - 1; j >= 0; j--) freq2sym[sum + j] = (u8)i;
        symb[i].cumFreq = sum;
        symb[i].freq = (f[i] >= (1 << logRange)) ? (1 << logRange) - 1 : f[i];
        sum += f[i];
      }
      res += alphabetSize;
    }
    return res;
  }

  // decodeSymbol, ANSRangeDecoder.java:333-347
  inline i32 decodeSymbol(int& idx, i32 st, const DecSymbol& sym, int mask) {
    st = jadd(jmul(sym.freq, jushr(st, logRange)), (st & mask)) - sym.cumFreq;
    if (st < ANS_TOP) {
      if (idx + 1 >= (int)buffer.size()) throw JavaException("AIOOBE in ANS decodeSymbol");
      st = (i32)(((u32)st << 8) | buffer[idx]);
      st = (i32)(((u32)st << 8) | buffer[idx + 1]);
      idx += 2;
    }
    return st;
  }

  // decodeChunkV2, ANSRangeDecoder.java:357-440
  bool decodeChunk(u8* block, int start, int end) {
    const i32 sz = readVarInt(bs);
    if (sz >= MAX_CHUNK_SIZE) return false;     // NB: negative sz passes this test in Java too
    i32 st0 = (i32)bs.readBits(32), st1 = (i32)bs.readBits(32), st2 = (i32)bs.readBits(32), st3 = (i32)bs.readBits(32);
    if (start == end) return true;
    const int minBufSize = std::max(2 * (end - start), 256);
    if ((int)buffer.size() < minBufSize) buffer.assign(minBufSize, 0);
    std::fill(buffer.begin(), buffer.end(), 0);
    if (sz < 0 || (sz > (int)buffer.size())) throw JavaException("Invalid bit count");   // readBits(byte[]) guard
    bs.readBytesBits(buffer.data(), 8 * (i64)sz);
    const int mask = (1 << logRange) - 1;
    const int end4 = start + ((end - start) & -4);
    int idx = 0;
    if (order == 0) {
      const u8* freq2sym = f2s[0].data();
      const DecSymbol* symb = &symbols[0];
      for (int i = start; i < end4; i += 4) {
        const int cur3 = freq2sym[st3 & mask]; block[i] = (u8)cur3; st3 = decodeSymbol(idx, st3, symb[cur3], mask);
        const int cur2 = freq2sym[st2 & mask]; block[i + 1] = (u8)cur2; st2 = decodeSymbol(idx, st2, symb[cur2], mask);
        const int cur1 = freq2sym[st1 & mask]; block[i + 2] = (u8)cur1; st1 = decodeSymbol(idx, st1, symb[cur1], mask);
        const int cur0 = freq2sym[st0 & mask]; block[i + 3] = (u8)cur0; st0 = decodeSymbol(idx, st0, symb[cur0], mask);
      }
    } else {
      const int quarter = (end4 - start) >> 2;
      int i0 = start, i1 = start + quarter, i2 = start + 2 * quarter, i3 = start + 3 * quarter;
      int prv0 = 0, prv1 = 0, prv2 = 0, prv3 = 0;
      auto F = [&](int prv, i32 st) -> int {
        if (f2s[prv].empty()) throw JavaException("AIOOBE f2s");   // context never declared
        return f2s[prv][st & mask];
      };
      for (; i0 < start + quarter; i0++, i1++, i2++, i3++) {
        const int cur3 = F(prv3, st3); block[i3] = (u8)cur3; st3 = decodeSymbol(idx, st3, symbols[(size_t)prv3 * 256 + cur3], mask);
        const int cur2 = F(prv2, st2); block[i2] = (u8)cur2; st2 = decodeSymbol(idx, st2, symbols[(size_t)prv2 * 256 + cur2], mask);
        const int cur1 = F(prv1, st1); block[i1] = (u8)cur1; st1 = decodeSymbol(idx, st1, symbols[(size_t)prv1 * 256 + cur1], mask);
        const int cur0 = F(prv0, st0); block[i0] = (u8)cur0; st0 = decodeSymbol(idx, st0, symbols[(size_t)prv0 * 256 + cur0], mask);
        prv3 = cur3; prv2 = cur2; prv1 = cur1; prv0 = cur0;
      }
    }
    int n = idx;
    for (int i = end4; i < end; i++) {
      if (n >= (int)buffer.size()) throw JavaException("AIOOBE in ANS tail");
      block[i] = buffer[n++];
    }
    return n == sz;
  }

  // decode, ANSRangeDecoder.java:189-236
  int decode(u8* block, int blkptr, int count) {
    if (count <= 32) { bs.readBytesBits(block + blkptr, 8 * (i64)count); return count; }
    const int end = blkptr + count;
    int startChunk = blkptr;
    for (auto& s : symbols) s = DecSymbol();
    int alphabet[256];
    while (startChunk < end) {
      const int endChunk = std::min(startChunk + chunkSize, end);
      const int alphabetSize = decodeHeader(alphabet);
      if (alphabetSize == 0) return startChunk - blkptr;
      if ((order == 0) && (alphabetSize == 1)) {
        for (int i = startChunk; i < endChunk; i++) block[i] = (u8)alphabet[0];
      } else {
        if (!decodeChunk(block, startChunk, endChunk)) break;
      }
      startChunk = endChunk;
    }
    return count;
  }
};
}  // namespace ans

// =================================================================================================
// FPAQ (FPAQEncoder.java, FPAQDecoder.java)
// =================================================================================================
namespace fpaq {
static const u64 TOP = 0x00FFFFFFFFFFFFFFULL, MASK_24_56 = 0x00FFFFFFFF000000ULL, MASK_0_24 = 0x0000000000FFFFFFULL,
                 MASK_0_32 = 0x00000000FFFFFFFFULL, MASK_0_56 = 0x00FFFFFFFFFFFFFFULL;
enum { DEFAULT_CHUNK_SIZE = 4 * 1024 * 1024, MAX_BLOCK_SIZE = 1 << 30, PSCALE = 65536 };

struct Encoder {
  BitWriter& bs;
  u64 low = 0, high = TOP;
  bool disposed = false;
  std::vector<u8> sba; int sbaIndex = 0;
  int probs[4][256]; int* p;
  explicit Encoder(BitWriter& b) : bs(b) {
    for (int i = 0; i < 4; i++) for (int j = 0; j < 256; j++) probs[i][j] = PSCALE >> 1;
    p = probs[0];
  }
  // flush, FPAQEncoder.java:208-213
  inline void flush() {
    if (sbaIndex + 4 > (int)sba.size()) throw JavaException("AIOOBE in FPAQ flush");
    put_be32(&sba[sbaIndex], (u32)(high >> 24));
    sbaIndex += 4;
    low <<= 32;
    high = (high << 32) | MASK_0_32;
  }
  // encodeBit, FPAQEncoder.java:182-199
  inline void encodeBit(int bit, int pIdx) {
    const u64 split = (((high - low) >> 8) * (u64)(i64)p[pIdx]) >> 8;
    if (bit == 0) { low += (split + 1); p[pIdx] -= (p[pIdx] >> 6); }
    else { high = low + split; p[pIdx] -= ((p[pIdx] - PSCALE + 64) >> 6); }
    while (((low ^ high) & MASK_24_56) == 0) flush();
  }
  // encode, FPAQEncoder.java:128-173
  int encode(const u8* block, int blkptr, int count) {
    if (count > MAX_BLOCK_SIZE) return -1;
    if (count == 0) return 0;
    int startChunk = blkptr;
    const int end = blkptr + count;
    while (startChunk < end) {
      const int chunkSize = std::min((int)DEFAULT_CHUNK_SIZE, end - startChunk);
      if ((int)sba.size() < (chunkSize + (chunkSize >> 3))) sba.assign(chunkSize + (chunkSize >> 3), 0);
      sbaIndex = 0;
      const int endChunk = startChunk + chunkSize;
      p = probs[0];
      for (int i = startChunk; i < endChunk; i++) {
        const int val = block[i];
        const int bits = val + 256;
        encodeBit(val & 0x80, 1);
        encodeBit(val & 0x40, bits >> 7);
        encodeBit(val & 0x20, bits >> 6);
        encodeBit(val & 0x10, bits >> 5);
        encodeBit(val & 0x08, bits >> 4);
        encodeBit(val & 0x04, bits >> 3);
        encodeBit(val & 0x02, bits >> 2);
        encodeBit(val & 0x01, bits >> 1);
        p = probs[val >> 6];
      }
      writeVarInt(bs, sbaIndex);
      bs.writeBytesBits(sba.data(), 8 * (i64)sbaIndex);
      startChunk += chunkSize;
      if (startChunk < end) bs.writeBits(low | MASK_0_24, 56);
    }
    return count;
  }
  // dispose, FPAQEncoder.java:232-238
  void dispose() {
    if (disposed) return;
    disposed = true;
    bs.writeBits(low | MASK_0_24, 56);
  }
};

struct Decoder {
  BitReader& bs;
  u64 low = 0, high = TOP, current = 0;
  std::vector<u8> sba; int sbaIndex = 0, bufLimit = 0;
  int probs[4][256]; int* p; int ctx = 1;
  explicit Decoder(BitReader& b) : bs(b) {
    for (int i = 0; i < 4; i++) for (int j = 0; j < 256; j++) probs[i][j] = PSCALE >> 1;
    p = probs[0];
  }
  // read, FPAQDecoder.java:322-335
  inline void read() {
    low = (low << 32) & MASK_0_56;
    high = ((high << 32) | MASK_0_32) & MASK_0_56;
    if (sbaIndex + 4 > bufLimit) {
      current = (current << 32) & MASK_0_56;
      sbaIndex = bufLimit + 1;
      return;
    }
    const u64 val = be32(&sba[sbaIndex]);
    current = ((current << 32) | val) & MASK_0_56;
    sbaIndex += 4;
  }
  // decodeBitV2, FPAQDecoder.java:290-314 (comparisons are on Java signed longs; all values < 2^56)
  inline int decodeBit(int pred) {
    const u64 split = ((((high - low) >> 8) * (u64)(i64)pred) >> 8) + low;
    int bit;
    if ((i64)split >= (i64)current) {
      bit = 1; high = split;
      p[ctx] -= ((p[ctx] - PSCALE + 64) >> 6);
      ctx = (ctx << 1) + 1;
    } else {
      bit = 0; low = split + 1;   // -~split
      p[ctx] -= (p[ctx] >> 6);
      ctx = ctx << 1;
    }
    while (((low ^ high) & MASK_24_56) == 0) read();
    return bit;
  }
  // decode, FPAQDecoder.java:161-242 (bsVersion >= 4 branch)
  int decode(u8* block, int blkptr, int count) {
    if (count > MAX_BLOCK_SIZE) return -1;
    if (count == 0) return 0;
    int startChunk = blkptr;
    const int end = blkptr + count;
    while (startChunk < end) {
      const i32 szBytes = readVarInt(bs);
      if (szBytes >= 2 * count) return 0;
      if (szBytes < 0) throw JavaException("negative size");
      const int bufSize = std::max(szBytes + (szBytes >> 2), 1024);
      if ((int)sba.size() < bufSize) sba.assign(bufSize, 0);
      current = bs.readBits(56);
      if (bufSize > szBytes) std::fill(sba.begin() + szBytes, sba.begin() + bufSize, 0);
      bs.readBytesBits(sba.data(), 8 * (i64)szBytes);
      bufLimit = szBytes;
      sbaIndex = 0;
      const int chunkSize = std::min((int)DEFAULT_CHUNK_SIZE, end - startChunk);
      const int endChunk = startChunk + chunkSize;
      p = probs[0];
      for (int i = startChunk; i < endChunk; i++) {
        ctx = 1;
        for (int k = 0; k < 8; k++) decodeBit(p[ctx]);
        block[i] = (u8)ctx;
        if (sbaIndex > szBytes) return 0;
        p = probs[(ctx & 0xFF) >> 6];
      }
      if (sbaIndex > szBytes) return 0;
      startChunk = endChunk;
    }
    return count;
  }
};
}  // namespace fpaq

// =================================================================================================
// RangeEncoder / RangeDecoder (entropy/RangeEncoder.java, RangeDecoder.java): order-0 range coder, 32 KiB chunks, each with its
// own statistics; 28 bits leave the coder at a time.  Not on the CUDA path yet: here so that the checker is
// ready and ctx["entropy"] = "RANGE" means something to RLT.
// =================================================================================================
namespace range {
static constexpr u64 TOP_RANGE = 0x0FFFFFFFFFFFFFFFULL, BOTTOM_RANGE = 0x000000000000FFFFULL, RANGE_MASK = 0x0FFFFFFF00000000ULL;
enum { DEFAULT_CHUNK_SIZE = 1 << 15, DEFAULT_LOG_RANGE = 12 };

struct Encoder {
  BitWriter& bs; int chunkSize, logRange;
  u64 low = 0, rng = TOP_RANGE; int shift = 0;
  int alphabet[256], freqs[257]; u64 cumFreqs[257];
  explicit Encoder(BitWriter& b, int chunkSize_ = DEFAULT_CHUNK_SIZE, int logRange_ = DEFAULT_LOG_RANGE) : bs(b), chunkSize(chunkSize_), logRange(logRange_) {}
  // encodeHeader, RangeEncoder.java:182-219
  bool encodeHeader(int alphabetSize, int lr) {
    const int encoded = encodeAlphabet(bs, alphabet, 256, alphabetSize);
    if (encoded < 0) return false;
    if (encoded == 0) return true;
    bs.writeBits((u64)(lr - 8), 3);
    const int chkSize = (alphabetSize >= 64) ? 8 : 6;
    int llr = 3;
    while ((1 << llr) <= lr) llr++;
    for (int i = 1; i < alphabetSize; i += chkSize) {
      int max = freqs[alphabet[i]] - 1, logMax = 0;
      const int endj = (i + chkSize < alphabetSize) ? i + chkSize : alphabetSize;
      for (int j = i + 1; j < endj; j++) if (freqs[alphabet[j]] - 1 > max) max = freqs[alphabet[j]] - 1;
      while ((1 << logMax) <= max) logMax++;
      bs.writeBits((u64)logMax, llr);
      if (logMax == 0) continue;
      for (int j = i; j < endj; j++) bs.writeBits((u64)(freqs[alphabet[j]] - 1), logMax);
    }
    return true;
  }
  // updateFrequencies + rebuildStatistics, :160-176, :298-301
  int rebuildStatistics(const u8* block, int start, int end, int lr) {
    histogramOrder0(block, start, end, freqs, false);
    const int alphabetSize = normalizeFrequencies(freqs, alphabet, 256, end - start, 1 << lr);
    if (alphabetSize > 0) { cumFreqs[0] = 0; for (int i = 0; i < 256; i++) cumFreqs[i + 1] = cumFreqs[i] + (u64)freqs[i]; }
    encodeHeader(alphabetSize, lr);
    return alphabetSize;
  }
  // encodeByte, :270-295 (Java longs: 64-bit wrap-around, `range > BOTTOM_RANGE` compares signed)
  void encodeByte(u8 b) {
    const u64 cumFreq = cumFreqs[b], freq = cumFreqs[b + 1] - cumFreq;
    rng >>= shift;
    low += cumFreq * rng;
    rng *= freq;
    while (true) {
      if (((low ^ (low + rng)) & RANGE_MASK) != 0) {
        if ((i64)rng > (i64)BOTTOM_RANGE) break;
        rng = (0 - low) & BOTTOM_RANGE;
      }
      bs.writeBits(low >> 32, 28);
      rng <<= 28;
      low <<= 28;
    }
  }
  // encode, :223-266
  int encode(const u8* block, int blkptr, int count) {
    if (count == 0) return 0;
    const int end = blkptr + count;
    int startChunk = blkptr;
    while (startChunk < end) {
      const int endChunk = (startChunk + chunkSize < end) ? startChunk + chunkSize : end;
      rng = TOP_RANGE; low = 0;
      int lr = logRange;
      while ((lr > 8) && ((1 << lr) > endChunk - startChunk)) lr--;
      if (rebuildStatistics(block, startChunk, endChunk, lr) <= 1) { startChunk = endChunk; continue; }
      shift = lr;
      for (int i = startChunk; i < endChunk; i++) encodeByte(block[i]);
      bs.writeBits(low, 60);
      startChunk = endChunk;
    }
    return count;
  }
};

struct Decoder {
  BitReader& bs; int chunkSize;
  u64 code = 0, low = 0, rng = TOP_RANGE; int shift = 0;
  int alphabet[256], freqs[256]; u64 cumFreqs[257];
  std::vector<short> f2s;
  explicit Decoder(BitReader& b, int chunkSize_ = DEFAULT_CHUNK_SIZE) : bs(b), chunkSize(chunkSize_) {}
  // decodeHeader, RangeDecoder.java:131-207
  int decodeHeader() {
    const int alphabetSize = decodeAlphabet(bs, alphabet);
    if (alphabetSize == 0) return 0;
    if (alphabetSize != 256) for (int i = 0; i < 256; i++) freqs[i] = 0;
    const int logRange = (int)(8 + bs.readBits(3));
    if ((logRange < 8) || (logRange > 15)) throw BitStreamError("Invalid bitstream: range");
    const int scale = 1 << logRange;
    shift = logRange;
    int sum = 0;
    const int chkSize = (alphabetSize >= 64) ? 8 : 6;
    int llr = 3;
    while ((1 << llr) <= logRange) llr++;
    for (int i = 1; i < alphabetSize; i += chkSize) {
      const int logMax = (int)bs.readBits(llr);
      if ((1 << logMax) > scale) throw BitStreamError("Invalid bitstream: incorrect frequency size");
      const int endj = (i + chkSize < alphabetSize) ? i + chkSize : alphabetSize;
      for (int j = i; j < endj; j++) {
        const int freq = (logMax == 0) ? 1 : (int)(1 + bs.readBits(logMax));
        if ((freq <= 0) || (freq >= scale)) throw BitStreamError("Invalid bitstream: incorrect frequency");
        freqs[alphabet[j]] = freq;
        sum += freq;
      }
    }
    if (scale <= sum) throw BitStreamError("Invalid bitstream: incorrect frequency");
    freqs[alphabet[0]] = scale - sum;
    cumFreqs[0] = 0;
    if ((int)f2s.size() < scale) f2s.assign(scale, 0);
    for (int i = 0; i < 256; i++) {
      cumFreqs[i + 1] = cumFreqs[i] + (u64)freqs[i];
      const int base = (int)cumFreqs[i];
      for (int j = freqs[i] - 1; j >= 0; j--) {
        if (base + j >= (int)f2s.size()) throw JavaException("AIOOBE in RangeDecoder.decodeHeader");
        f2s[base + j] = (short)i;
      }
    }
    return alphabetSize;
  }
  // decodeByte, :252-278 (the division is Java's signed long division)
  u8 decodeByte() {
    rng >>= shift;
    if (rng == 0) throw JavaException("ArithmeticException in RangeDecoder.decodeByte");
    const int count = (int)((i64)(code - low) / (i64)rng);
    if (count < 0 || count >= (int)f2s.size()) throw JavaException("AIOOBE in RangeDecoder.decodeByte");
    const int symbol = f2s[count];
    const u64 cumFreq = cumFreqs[symbol], freq = cumFreqs[symbol + 1] - cumFreq;
    low += cumFreq * rng;
    rng *= freq;
    while (true) {
      if (((low ^ (low + rng)) & RANGE_MASK) != 0) {
        if ((i64)rng > (i64)BOTTOM_RANGE) break;
        rng = (0 - low) & BOTTOM_RANGE;
      }
      code = (code << 28) | bs.readBits(28);
      rng <<= 28;
      low <<= 28;
    }
    return (u8)symbol;
  }
  // decode, :210-249
  int decode(u8* block, int blkptr, int count) {
    if (count == 0) return 0;
    const int end = blkptr + count;
    int startChunk = blkptr;
    while (startChunk < end) {
      const int endChunk = (startChunk + chunkSize < end) ? startChunk + chunkSize : end;
      const int alphabetSize = decodeHeader();
      if (alphabetSize == 0) return startChunk - blkptr;
      if (alphabetSize == 1) {
        for (int i = startChunk; i < endChunk; i++) block[i] = (u8)alphabet[0];
        startChunk = endChunk;
        continue;
      }
      rng = TOP_RANGE; low = 0;
      code = bs.readBits(60);
      for (int i = startChunk; i < endChunk; i++) block[i] = decodeByte();
      startChunk = endChunk;
    }
    return count;
  }
};
}  // namespace range

// =================================================================================================
// EntropyCodecFactory.newEncoder/newDecoder (EntropyCodecFactory.java:113-203) + the host's
// encode-then-dispose sequence (CompressedOutputStream.java:907-916)
// =================================================================================================
static inline int entropyEncode(int type, BitWriter& bs, const u8* block, int count) {
  switch (type) {
    case E_NONE: return nullEncode(bs, block, 0, count);
    case E_HUFFMAN: { huff::Encoder e(bs); return e.encode(block, 0, count); }
    case E_ANS0: { ans::Encoder e(bs, 0); return e.encode(block, 0, count); }
    case E_ANS1: { ans::Encoder e(bs, 1); return e.encode(block, 0, count); }
    case E_FPAQ: { fpaq::Encoder e(bs); int r = e.encode(block, 0, count); e.dispose(); return r; }
    case E_RANGE: { range::Encoder e(bs); return e.encode(block, 0, count); }
    default: throw JavaException("Unknown entropy codec type");
  }
}
static inline int entropyDecode(int type, BitReader& bs, u8* block, int count) {
  switch (type) {
    case E_NONE: return nullDecode(bs, block, 0, count);
    case E_HUFFMAN: { huff::Decoder d(bs); return d.decode(block, 0, count); }
    case E_ANS0: { ans::Decoder d(bs, 0); return d.decode(block, 0, count); }
    case E_ANS1: { ans::Decoder d(bs, 1); return d.decode(block, 0, count); }
    case E_FPAQ: { fpaq::Decoder d(bs); return d.decode(block, 0, count); }
    case E_RANGE: { range::Decoder d(bs); return d.decode(block, 0, count); }
    default: throw JavaException("Unsupported entropy codec type");
  }
}

}  // namespace kzo
