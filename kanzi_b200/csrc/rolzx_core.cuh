// rolzx_core.cuh — ROLZX = ROLZCodec2 (K/transform/ROLZCodec.java:1016-1428) and its adaptive binary arithmetic coder
// (ROLZEncoder :1431-1597, ROLZDecoder :1599-1770), one thread per block.
//
// Everything here is one dependent chain: a literal is nine coded bits, each bit's probability cell is chosen by the bits before
// it, the interval update needs the probability, the match finder's ring depends on every earlier position, and the decoder
// cannot know a byte before it has decoded the previous one.  What is done about the chain's cost on a GPU:
//   * the 32 ring slots of a key are read as eight 16-byte loads and filtered by their 8-bit hash in registers; only the slots
//     that pass (1 in 256 by chance, plus the real candidates) are walked in the reference's order, with its early exits;
//   * the 1 KiB probability row of the current context byte is prefetched into L1 before the nine dependent look-ups start;
//   * probabilities are 16-bit cells (they never leave [0, 65535]), so a block's tables take 272 KiB of scratch.
// Compiled for the device by rolzx.cu and, unchanged, for the host by tests/native/sibling_hostcheck.cpp (test infrastructure:
// held against the CPU oracle; the product never runs it on the host).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RZX_HD __host__ __device__ __forceinline__
#else
#define RZX_HD static inline
#endif

#define RZX_HASH_SIZE 65536
#define RZX_LOG_POS 5
#define RZX_POS_MASK 31
#define RZX_CHUNK (16 << 20)
#define RZX_HASH 200002979u
#define RZX_HASH_MASK 0xFF000000u
#define RZX_MAX_MATCH (3 + 255)
#define RZX_LIT_CELLS (256 << 9)
#define RZX_MATCH_CELLS (256 << RZX_LOG_POS)
#define RZX_MATCH_INTS ((size_t)RZX_HASH_SIZE << RZX_LOG_POS)
#define RZX_TOP 0x00FFFFFFFFFFFFFFull
#define RZX_MASK32 0x00000000FFFFFFFFull

struct alignas(16) RzxQuad { int32_t v[4]; };

struct RzxCoder {
  uint64_t low, high, current;
  uint16_t* lit; uint16_t* mat;        // probability cells: [256 << 9] and [256 << 5], all 0x7FFF at the start
  uint16_t* p;                         // the table of the current context kind, offset by the context byte's row
  int c1;
  uint8_t* buf; int index; int limit;  // coded bytes: encoder writes / decoder reads at buf[index], never at or beyond limit
  int overrun;                         // the coder wanted bytes beyond limit (Java: ArrayIndexOutOfBounds or stale bytes; here: block fails)
};

RZX_HD uint64_t rzx_ld64(const uint8_t* p) {
#if defined(__CUDA_ARCH__)
  const uintptr_t a = (uintptr_t)p;
  const uint64_t* q = (const uint64_t*)(a & ~(uintptr_t)7);
  const int sh = (int)(a & 7) * 8;
  const uint64_t w0 = q[0];
  if (sh == 0) return w0;
  return (w0 >> sh) | (q[1] << (64 - sh));
#else
  uint64_t v; memcpy(&v, p, 8); return v;
#endif
}
RZX_HD int rzx_ctz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)v) - 1;
#else
  return __builtin_ctzll(v);
#endif
}
RZX_HD int rzx_ctz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __ffs((int)v) - 1;
#else
  return __builtin_ctz(v);
#endif
}
RZX_HD void rzx_prefetch_row(const uint16_t* row) {                // 512 cells = 1 KiB = eight 128-byte lines
#if defined(__CUDA_ARCH__)
  const char* c = (const char*)row;
#pragma unroll
  for (int i = 0; i < 8; i++) asm volatile("prefetch.global.L1 [%0];" ::"l"(c + 128 * i));
#else
  (void)row;
#endif
}
RZX_HD uint32_t rzx_ld32le(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
RZX_HD int rzx_key1(const uint8_t* buf, int idx) { return (int)buf[idx] | ((int)buf[idx + 1] << 8); }                                   // :123-125
RZX_HD int rzx_key2(const uint8_t* buf, int idx) {                                                                                  // :135-137
  uint64_t v = 0;
  for (int k = 7; k >= 0; k--) v = (v << 8) | buf[idx + k];
  return (int)((int64_t)(v * (uint64_t)RZX_HASH) >> 40) & 0xFFFF;
}
RZX_HD uint32_t rzx_hash(const uint8_t* buf, int idx) { return ((rzx_ld32le(buf + idx) << 8) * RZX_HASH) & RZX_HASH_MASK; }           // :147-149

RZX_HD void rzx_coder_init(RzxCoder& C, uint16_t* lit, uint16_t* mat, uint8_t* buf, int index, int limit) {
  C.low = 0; C.high = RZX_TOP; C.current = 0; C.lit = lit; C.mat = mat; C.p = lit; C.c1 = 1; C.buf = buf; C.index = index; C.limit = limit; C.overrun = 0;
}
RZX_HD void rzx_ctx_literal(RzxCoder& C, int prevByte) { C.p = C.lit + ((size_t)prevByte << 9); rzx_prefetch_row(C.p); }           // setContext(LITERAL_CTX, b)
RZX_HD void rzx_ctx_match(RzxCoder& C, int prevByte) { C.p = C.mat + ((size_t)prevByte << RZX_LOG_POS); }                            // setContext(MATCH_CTX, b)

// ROLZEncoder.encodeBit (:1553-1580): Java long arithmetic is 64-bit wrap-around; >>> is the logical shift
RZX_HD void rzx_encode_bit(RzxCoder& C, int bit) {
  uint16_t* cell = C.p + C.c1;
  const int pr = *cell;
  const uint64_t split = (((C.high - C.low) >> 4) * (uint64_t)(pr >> 4)) >> 8;
  if (bit == 0) { C.low += split + 1; *cell = (uint16_t)(pr - (pr >> 5)); C.c1 += C.c1; }
  else { C.high = C.low + split; *cell = (uint16_t)(pr - (((pr - 0xFFFF) >> 5) + 1)); C.c1 += C.c1 + 1; }
  while (((C.low ^ C.high) >> 24) == 0) {
    if (C.index + 4 <= C.limit) {
      const uint32_t w = (uint32_t)(C.high >> 32);
      C.buf[C.index] = (uint8_t)(w >> 24); C.buf[C.index + 1] = (uint8_t)(w >> 16); C.buf[C.index + 2] = (uint8_t)(w >> 8); C.buf[C.index + 3] = (uint8_t)w;
    } else C.overrun = 1;
    C.index += 4;
    C.low <<= 32;
    C.high = (C.high << 32) | RZX_MASK32;
  }
}
RZX_HD void rzx_encode9(RzxCoder& C, int val) {                    // encode9Bits (:1538-1550)
  C.c1 = 1;
  for (int m = 0x100; m != 0; m >>= 1) rzx_encode_bit(C, val & m);
}
RZX_HD void rzx_encode_bits(RzxCoder& C, int val, int n) {         // encodeBits (:1526-1535), n >= 1
  C.c1 = 1;
  do { n--; rzx_encode_bit(C, val & (1 << n)); } while (n != 0);
}
RZX_HD void rzx_encoder_dispose(RzxCoder& C) {                     // :1583-1590
  for (int i = 0; i < 8; i++) {
    if (C.index + i < C.limit) C.buf[C.index + i] = (uint8_t)(C.low >> 56); else C.overrun = 1;
    C.low <<= 8;
  }
  C.index += 8;
}
RZX_HD uint32_t rzx_next_be32(RzxCoder& C) {                       // bytes the stream does not hold read as zero and flag the block
  uint32_t v = 0;
  for (int k = 0; k < 4; k++) {
    v <<= 8;
    if (C.index + k < C.limit) v |= C.buf[C.index + k]; else C.overrun = 1;
  }
  C.index += 4;
  return v;
}
RZX_HD void rzx_decoder_start(RzxCoder& C) {                       // ROLZDecoder constructor (:1625-1648)
  const uint64_t hi = rzx_next_be32(C);
  C.current = (hi << 32) | (uint64_t)rzx_next_be32(C);
}
// ROLZDecoder.decodeBit (:1733-1764); `mid >= current` compares Java longs, i.e. signed
RZX_HD void rzx_decode_bit(RzxCoder& C) {
  uint16_t* cell = C.p + C.c1;
  const int pr = *cell;
  const uint64_t mid = C.low + ((((C.high - C.low) >> 4) * (uint64_t)(pr >> 4)) >> 8);
  if ((int64_t)mid >= (int64_t)C.current) { C.high = mid; *cell = (uint16_t)(pr - (((pr - 0xFFFF) >> 5) + 1)); C.c1 += C.c1 + 1; }
  else { C.low = mid + 1; *cell = (uint16_t)(pr - (pr >> 5)); C.c1 += C.c1; }
  while (((C.low ^ C.high) >> 24) == 0) {
    C.low = (C.low << 32) & RZX_TOP;
    C.high = ((C.high << 32) | RZX_MASK32) & RZX_TOP;
    C.current = ((C.current << 32) | (uint64_t)rzx_next_be32(C)) & RZX_TOP;
  }
}
RZX_HD int rzx_decode9(RzxCoder& C) {                              // decode9Bits (:1718-1730)
  C.c1 = 1;
  for (int i = 0; i < 9; i++) rzx_decode_bit(C);
  return C.c1 & 0x1FF;
}
RZX_HD int rzx_decode_bits(RzxCoder& C, int n) {                   // decodeBits (:1702-1715)
  C.c1 = 1;
  const int mask = (1 << n) - 1;
  do { rzx_decode_bit(C); n--; } while (n != 0);
  return C.c1 & mask;
}

// ROLZCodec2.findMatch (:1114-1173) for the chunk [sbaIndex, sbaLength) of buf.  Returns -1 or (bestIdx << 16) | (bestLen - minMatch).
RZX_HD int rzx_find_match(const uint8_t* buf, int sbaLength, int sbaIndex, int pos, int key, int32_t* matches, int32_t* counters, int minMatch) {
  const int base = key << RZX_LOG_POS;
  const uint32_t hash32 = rzx_hash(buf, pos);
  const int counter = counters[key];
  // slots whose stored hash equals this position's, as a mask over their age (0 = the slot written last): the order the reference walks them in
  uint32_t cand = 0;
  const RzxQuad* row = (const RzxQuad*)(matches + base);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int q = 0; q < 8; q++) {
    const RzxQuad e = row[q];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; k++)
      if (((uint32_t)e.v[k] & RZX_HASH_MASK) == hash32) cand |= 1u << ((counter - (4 * q + k)) & RZX_POS_MASK);
  }
  int bestLen = 0, bestIdx = -1;
  const int maxMatch = ((RZX_MAX_MATCH < sbaLength - pos) ? RZX_MAX_MATCH : (sbaLength - pos)) - 8;
  while (cand != 0) {
    const int age = rzx_ctz32(cand);
    cand &= cand - 1;
    const int ref = (int)((uint32_t)matches[base + ((counter - age) & RZX_POS_MASK)] & ~RZX_HASH_MASK) + sbaIndex;
    if (buf[ref + bestLen] != buf[pos + bestLen]) continue;
    int n = 0;
    while (n < maxMatch) {
      const uint64_t diff = rzx_ld64(buf + ref + n) ^ rzx_ld64(buf + pos + n);
      if (diff != 0) { n += rzx_ctz64(diff) >> 3; break; }
      n += 8;
    }
    if (n > bestLen) {
      bestIdx = age; bestLen = n;
      if (bestLen == maxMatch) break;
    }
  }
  const int c = (counter + 1) & RZX_POS_MASK;
  counters[key] = c;
  matches[base + c] = (int32_t)(hash32 | (uint32_t)(pos - sbaIndex));
  return (bestLen < minMatch) ? -1 : (bestIdx << 16) | (bestLen - minMatch);
}

// One chunk of ROLZCodec2.forward's main loop (:1240-1281): `matches` all zero, `counters` carried from the chunk before.
RZX_HD void rzx_forward_chunk(const uint8_t* src, int startChunk, int endChunk, int srcEnd, RzxCoder& C, int32_t* matches, int32_t* counters, int mm, int dt) {
  int srcIdx = startChunk;
  const int n = (srcEnd - startChunk < 8) ? (srcEnd - startChunk) : 8;
  rzx_ctx_literal(C, 0);
  for (int j = 0; j < n; j++) { rzx_encode9(C, 0x100 | src[srcIdx]); srcIdx++; }
  while (srcIdx < endChunk) {
    rzx_ctx_literal(C, src[srcIdx - 1]);
    const int key = (mm == 3) ? rzx_key1(src, srcIdx - dt) : rzx_key2(src, srcIdx - dt);
    const int match = rzx_find_match(src, endChunk, startChunk, srcIdx, key, matches, counters, mm);
    if (match < 0) { rzx_encode9(C, 0x100 | src[srcIdx]); srcIdx++; continue; }
    const int matchLen = match & 0xFFFF;
    rzx_encode9(C, matchLen);                                      // MATCH_FLAG = 0 in bit 8
    rzx_ctx_match(C, src[srcIdx - 1]);
    rzx_encode_bits(C, (int)((uint32_t)match >> 16), RZX_LOG_POS);
    srcIdx += matchLen + mm;
  }
}
// the four trailing literals and the coder's last eight bytes (:1283-1289); srcIdx = srcEnd
RZX_HD void rzx_forward_tail(const uint8_t* src, int srcIdx, RzxCoder& C) {
  for (int i = 0; i < 4; i++, srcIdx++) {
    rzx_ctx_literal(C, src[srcIdx - 1]);
    rzx_encode9(C, 0x100 | src[srcIdx]);
  }
  rzx_encoder_dispose(C);
}

// One chunk of ROLZCodec2.inverse's loop (:1351-1405).  *outIndex = output.index on entry (the chunk's base for ring positions) and on
// exit; dstEnd = szBlock, dstCap = bytes dst really holds.  Returns false where the reference returns false or would throw.
RZX_HD bool rzx_inverse_chunk(uint8_t* dst, int startChunk, int endChunk, int dstEnd, int dstCap, int* outIndex, RzxCoder& C,
                              int32_t* matches, int32_t* counters, int mm, int dt) {
  const int base0 = *outIndex;
  int dstIdx = base0;
  const int n = (dstEnd - startChunk < 8) ? (dstEnd - startChunk) : 8;
  rzx_ctx_literal(C, 0);
  for (int j = 0; j < n; j++) {
    const int val1 = rzx_decode9(C);
    if ((val1 >> 8) == 0) { *outIndex = dstIdx; return false; }
    if (dstIdx >= dstCap) return false;
    dst[dstIdx++] = (uint8_t)val1;
  }
  while (dstIdx < endChunk) {
    const int savedIdx = dstIdx;
    if (dstIdx - dt < 0 || dstIdx >= dstCap) return false;
    const int key = (mm == 3) ? rzx_key1(dst, dstIdx - dt) : rzx_key2(dst, dstIdx - dt);
    const int base = key << RZX_LOG_POS;
    rzx_ctx_literal(C, dst[dstIdx - 1]);
    const int val = rzx_decode9(C);
    if ((val >> 8) == 1) {
      dst[dstIdx++] = (uint8_t)val;
    } else {
      const int matchLen = val & 0xFF;
      if (dstIdx + matchLen + 3 > dstEnd) { *outIndex = dstIdx; return false; }
      rzx_ctx_match(C, dst[dstIdx - 1]);
      const int matchIdx = rzx_decode_bits(C, RZX_LOG_POS);
      const int ref = base0 + matches[base + ((counters[key] - matchIdx) & RZX_POS_MASK)];
      const int len = matchLen + mm;
      if (dstIdx + len > dstCap) return false;
      for (int k = 0; k < len; k++) dst[dstIdx + k] = dst[ref + k];
      dstIdx += len;
    }
    const int c = (counters[key] + 1) & RZX_POS_MASK;
    counters[key] = c;
    matches[base + c] = savedIdx - base0;
  }
  *outIndex = dstIdx;
  return true;
}
