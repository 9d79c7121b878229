// lzp_core.cuh — the LZP parse and replay loops (K/transform/LZCodec.java:973-1287, LZPCodec), one thread per block.
//
// LZP predicts: the hash of the last four bytes names ONE earlier position; a match of 64+ bytes against it is coded as a
// flag byte plus a length, everything else is a literal.  Whether a position enters the table depends on every decision before
// it (positions inside a match never do), so the parse is a serial chain per block; the batch of blocks supplies the
// parallelism.  What the loops below do about the chain's cost on a GPU (a lone warp issues one instruction every few cycles
// and waits hundreds for an L2 round trip, DESIGN.md "What a lone warp costs"):
//   * forward: the table lookups of the next four positions are issued together, on the guess that all four are literals
//     (their hashes need only the source bytes); a later position that hashes like an earlier one of the same group takes
//     that position instead of the value loaded, exactly what the serial order would have stored.  A match drops the rest
//     of the group.  The group's literal bytes and its four 4-byte probes at +60 come from two 8-byte loads;
//   * inverse: the table is read only when the coded byte is the flag (a literal never looks at what it replaces).
//
// Compiled for the device by lzp.cu and, unchanged, for the host by tests/native/lzp_hostcheck.cu, which holds these same
// loops against the CPU oracle (test infrastructure: the product never runs them on the host).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define LZP_HD __host__ __device__ __forceinline__
#else
#define LZP_HD static inline
#endif

#define LZP_HASH_SEED 0x7FEB352Du
#define LZP_HASH_SHIFT 16
#define LZP_MIN_MATCH 64
#define LZP_MIN_BLOCK 128
#define LZP_MATCH_FLAG 0xFC
#define LZP_TABLE_INTS 65536

LZP_HD uint64_t lzp_ld64(const uint8_t* p) {
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
LZP_HD uint32_t lzp_ld32(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
LZP_HD int lzp_ctz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)v) - 1;
#else
  return __builtin_ctzll(v);
#endif
}
// LZPCodec.findMatch (:1267-1280)
LZP_HD int lzp_find_match(const uint8_t* src, int srcIdx, int ref, int maxMatch) {
  int bestLen = 0;
  while (bestLen + 8 <= maxMatch) {
    const uint64_t diff = lzp_ld64(src + srcIdx + bestLen) ^ lzp_ld64(src + ref + bestLen);
    if (diff != 0) { bestLen += lzp_ctz64(diff) >> 3; break; }
    bestLen += 8;
  }
  return bestLen;
}

// One position of the main loop once its table entry `ref` is known (:1046-1086).  Returns 0 = go on, 1 = a match was
// coded (the caller's group is void), -1 = `return false`.
LZP_HD int lzp_fwd_position(const uint8_t* src, int srcEnd, uint8_t* dst, int dstEnd, int ref, uint32_t probe, int val,
                            int& srcIdx, int& dstIdx, uint32_t& ctx) {
  int bestLen = 0;
  if (ref != 0 && lzp_ld32(src + ref + LZP_MIN_MATCH - 4) == probe) bestLen = lzp_find_match(src, srcIdx, ref, srcEnd - srcIdx);
  if (bestLen < LZP_MIN_MATCH) {
    ctx = (ctx << 8) | (uint32_t)val;
    dst[dstIdx++] = (uint8_t)val;
    srcIdx++;
    if (ref != 0 && val == LZP_MATCH_FLAG) {
      if (dstIdx >= dstEnd) return -1;
      dst[dstIdx++] = 0xFF;
    }
    return 0;
  }
  srcIdx += bestLen;
  ctx = lzp_ld32(src + srcIdx - 4);
  dst[dstIdx++] = LZP_MATCH_FLAG;
  bestLen -= LZP_MIN_MATCH;
  while (bestLen >= 254) {
    bestLen -= 254;
    dst[dstIdx++] = 0xFE;
    if (dstIdx >= dstEnd) break;
  }
  if (dstIdx >= dstEnd) return -1;
  dst[dstIdx++] = (uint8_t)bestLen;
  return 1;
}

// LZPCodec.forward (:1003-1118) for a slice with index 0 whose guards (:1004-1018) passed: count >= 128, `hashes` all zero,
// dst holds count - (count >> 6) bytes or more.  Returns the Java boolean; *outLen = output.index.
LZP_HD bool lzp_forward_core(const uint8_t* src, int count, uint8_t* dst, int32_t* hashes, int* outLen) {
  const int srcEnd = count;
  const int dstEnd = count - (count >> 6);
  dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
  uint32_t ctx = lzp_ld32(src);
  int srcIdx = 4, dstIdx = 4;
  const int mainEnd = srcEnd - LZP_MIN_MATCH;
  *outLen = 0;
  while (srcIdx < mainEnd && dstIdx < dstEnd) {
    if (srcIdx + 72 <= srcEnd) {
      // group of four positions p .. p+3 (all < mainEnd: p + 72 <= srcEnd): bytes [p, p+8) and [p+60, p+68) are inside the block
      const int p = srcIdx;
      const uint64_t lit = lzp_ld64(src + p);
      const uint64_t probes = lzp_ld64(src + p + LZP_MIN_MATCH - 4);
      const uint32_t c0 = ctx, c1 = (c0 << 8) | (uint32_t)(lit & 0xFF), c2 = (c1 << 8) | (uint32_t)((lit >> 8) & 0xFF),
                     c3 = (c2 << 8) | (uint32_t)((lit >> 16) & 0xFF);
      const uint32_t h0 = (LZP_HASH_SEED * c0) >> LZP_HASH_SHIFT, h1 = (LZP_HASH_SEED * c1) >> LZP_HASH_SHIFT,
                     h2 = (LZP_HASH_SEED * c2) >> LZP_HASH_SHIFT, h3 = (LZP_HASH_SEED * c3) >> LZP_HASH_SHIFT;
      const int r0 = hashes[h0];
      int r1 = hashes[h1], r2 = hashes[h2], r3 = hashes[h3];
      if (h1 == h0) r1 = p;
      if (h2 == h0) r2 = p;
      if (h2 == h1) r2 = p + 1;
      if (h3 == h0) r3 = p;
      if (h3 == h1) r3 = p + 1;
      if (h3 == h2) r3 = p + 2;
      int st;
      hashes[h0] = p;
      st = lzp_fwd_position(src, srcEnd, dst, dstEnd, r0, (uint32_t)probes, (int)(lit & 0xFF), srcIdx, dstIdx, ctx);
      if (st < 0) return false;
      if (st > 0 || dstIdx >= dstEnd) continue;
      hashes[h1] = p + 1;
      st = lzp_fwd_position(src, srcEnd, dst, dstEnd, r1, (uint32_t)(probes >> 8), (int)((lit >> 8) & 0xFF), srcIdx, dstIdx, ctx);
      if (st < 0) return false;
      if (st > 0 || dstIdx >= dstEnd) continue;
      hashes[h2] = p + 2;
      st = lzp_fwd_position(src, srcEnd, dst, dstEnd, r2, (uint32_t)(probes >> 16), (int)((lit >> 16) & 0xFF), srcIdx, dstIdx, ctx);
      if (st < 0) return false;
      if (st > 0 || dstIdx >= dstEnd) continue;
      hashes[h3] = p + 3;
      st = lzp_fwd_position(src, srcEnd, dst, dstEnd, r3, (uint32_t)(probes >> 24), (int)((lit >> 24) & 0xFF), srcIdx, dstIdx, ctx);
      if (st < 0) return false;
    } else {
      const uint32_t h = (LZP_HASH_SEED * ctx) >> LZP_HASH_SHIFT;
      const int ref = hashes[h];
      hashes[h] = srcIdx;
      if (lzp_fwd_position(src, srcEnd, dst, dstEnd, ref, lzp_ld32(src + srcIdx + LZP_MIN_MATCH - 4), src[srcIdx], srcIdx, dstIdx, ctx) < 0) return false;
    }
  }
  while (srcIdx < srcEnd && dstIdx < dstEnd) {                     // the last 64 bytes: literals only (:1090-1111)
    const uint32_t h = (LZP_HASH_SEED * ctx) >> LZP_HASH_SHIFT;
    const int ref = hashes[h];
    hashes[h] = srcIdx;
    const int val = src[srcIdx];
    ctx = (ctx << 8) | (uint32_t)val;
    dst[dstIdx++] = src[srcIdx++];
    if (ref != 0 && val == LZP_MATCH_FLAG) {
      if (dstIdx >= dstEnd) return false;
      dst[dstIdx++] = 0xFF;
    }
  }
  *outLen = dstIdx;
  return (srcIdx == count) && (dstIdx < dstEnd);
}

// LZPCodec.inverse (:1121-1263) for slices with index 0: count = src.length >= 4, dstEnd = min(dst.length, bytes dst really
// holds), `hashes` all zero.  Returns the Java boolean; *outLen = output.index (meaningful when true).
LZP_HD bool lzp_inverse_core(const uint8_t* src, int count, uint8_t* dst, int dstEnd, int32_t* hashes, int* outLen) {
  const int srcEnd = count;
  *outLen = 0;
  if (dstEnd < count || dstEnd < 4) return false;
  dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
  uint32_t ctx = lzp_ld32(dst);
  int srcIdx = 4, dstIdx = 4;
  while (srcIdx < srcEnd) {
    const uint32_t h = (LZP_HASH_SEED * ctx) >> LZP_HASH_SHIFT;
    const int b = src[srcIdx];
    int ref = 0;
    if (b == LZP_MATCH_FLAG) ref = hashes[h];
    hashes[h] = dstIdx;
    if (ref == 0) {                                                // also: any byte but the flag, whatever the table holds
      if (dstIdx >= dstEnd) return false;
      dst[dstIdx++] = (uint8_t)b;
      ctx = (ctx << 8) | (uint32_t)b;
      srcIdx++;
      continue;
    }
    srcIdx++;
    if (srcIdx >= srcEnd) return false;
    if (src[srcIdx] == 0xFF) {
      if (dstIdx >= dstEnd) return false;
      dst[dstIdx++] = LZP_MATCH_FLAG;
      ctx = (ctx << 8) | (uint32_t)LZP_MATCH_FLAG;
      srcIdx++;
      continue;
    }
    int mLen = LZP_MIN_MATCH;
    if (src[srcIdx] == 0xFE) {
      while (srcIdx < srcEnd && src[srcIdx] == 0xFE) { srcIdx++; mLen += 254; }
      if (srcIdx >= srcEnd) return false;
    }
    mLen += src[srcIdx++];
    if (dstIdx + mLen > dstEnd) return false;
    int i = 0;
    if (dstIdx - ref >= 8) {                                       // the eight source bytes are all written already
      for (; i + 8 <= mLen; i += 8) {
        const uint64_t w = lzp_ld64(dst + ref + i);
        uint8_t* o = dst + dstIdx + i;
        o[0] = (uint8_t)w; o[1] = (uint8_t)(w >> 8); o[2] = (uint8_t)(w >> 16); o[3] = (uint8_t)(w >> 24);
        o[4] = (uint8_t)(w >> 32); o[5] = (uint8_t)(w >> 40); o[6] = (uint8_t)(w >> 48); o[7] = (uint8_t)(w >> 56);
      }
    }
    for (; i < mLen; i++) dst[dstIdx + i] = dst[ref + i];
    dstIdx += mLen;
    ctx = lzp_ld32(dst + dstIdx - 4);
  }
  *outLen = dstIdx;
  return srcIdx == srcEnd;
}
