// rlt_core.cuh — the RLT scan loops (K/transform/RLT.java:62-352), one thread per block.
//
// Run-length coding with an escape byte: runs of 4+ become `byte, escape, length (1-3 bytes)`, shorter runs stay literal, a literal
// escape is followed by 0.  The reference's forward loop has a shape of its own (runs counted four bytes at a time, cut at
// MAX_RUN4, the last four bytes of the block handled by a separate literal loop, destination checks that depend on the run in
// hand); it is kept step for step because the output depends on it.  Compiled for the device by rlt.cu and, unchanged, for the
// host by tests/native/rlt_hostcheck.cpp (test infrastructure: held against the CPU oracle; the product never runs it on the host).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RLT_HD __host__ __device__ __forceinline__
#else
#define RLT_HD static inline
#endif

#define RLT_RUN_LEN_ENCODE1 224
#define RLT_RUN_LEN_ENCODE2 ((255 - RLT_RUN_LEN_ENCODE1) << 8)
#define RLT_RUN_THRESHOLD 3
#define RLT_MAX_RUN (0xFFFF + RLT_RUN_LEN_ENCODE2 + RLT_RUN_THRESHOLD - 1)
#define RLT_MAX_RUN4 (RLT_MAX_RUN - 4)
#define RLT_DEFAULT_ESCAPE 0xFB
// DataType ordinals (K/Global.java:40-90), as in kzg.h
#define RLT_DT_UNDEFINED 0
#define RLT_DT_NUMERIC 4
#define RLT_DT_BASE64 5
#define RLT_DT_DNA 6
#define RLT_DT_BIN 7
#define RLT_DT_UTF8 8
#define RLT_DT_SMALL_ALPHABET 9

// Global.detectSimpleType (K/Global.java:556-608) on a 256-bin histogram
RLT_HD int rlt_detect_type(int count, const uint32_t* f) {
  if (count == 0) return RLT_DT_UNDEFINED;
  int sum = (int)(f['a'] + f['c'] + f['g'] + f['n'] + f['t'] + f['u'] + f['A'] + f['C'] + f['G'] + f['N'] + f['T'] + f['U']);
  if (sum > count - count / 12) return RLT_DT_DNA;
  sum = (int)(f['+'] + f['-'] + f['*'] + f['/'] + f['='] + f[','] + f['.'] + f[':'] + f[';'] + f[' ']);
  for (int c = '0'; c <= '9'; c++) sum += (int)f[c];
  if (sum == count) return RLT_DT_NUMERIC;
  sum = (f[0x3D] == 1) ? 1 : 0;
  for (int c = 'A'; c <= 'Z'; c++) sum += (int)f[c];
  for (int c = 'a'; c <= 'z'; c++) sum += (int)f[c];
  for (int c = '0'; c <= '9'; c++) sum += (int)f[c];
  sum += (int)(f['+'] + f['/']);
  if (sum == count) return RLT_DT_BASE64;
  sum = 0;
  for (int i = 0; i < 256; i++) sum += (f[i] > 0) ? 1 : 0;
  if (sum == 256) return RLT_DT_BIN;
  if (sum <= 4) return RLT_DT_SMALL_ALPHABET;
  return RLT_DT_UNDEFINED;
}
// the rarest byte, the first absent one if there is any (RLT.java:129-141)
RLT_HD int rlt_best_escape(const uint32_t* f) {
  int minIdx = 0;
  if (f[0] > 0) {
    for (int i = 1; i < 256; i++) {
      if (f[i] < f[minIdx]) { minIdx = i; if (f[i] == 0) break; }
    }
  }
  return minIdx;
}
RLT_HD int rlt_emit_run_length(uint8_t* dst, int dstIdx, int run) {                 // :233-249
  run -= RLT_RUN_THRESHOLD;
  if (run >= RLT_RUN_LEN_ENCODE1) {
    if (run < RLT_RUN_LEN_ENCODE2) { run -= RLT_RUN_LEN_ENCODE1; dst[dstIdx++] = (uint8_t)(RLT_RUN_LEN_ENCODE1 + (run >> 8)); }
    else { run -= RLT_RUN_LEN_ENCODE2; dst[dstIdx++] = 0xFF; dst[dstIdx++] = (uint8_t)(run >> 8); }
  }
  dst[dstIdx] = (uint8_t)run;
  return dstIdx + 1;
}

// RLT.forward from :143 on, for slices with index 0: count >= 16, dstEnd = dst.array.length (at least 3).  Returns the Java
// boolean; *outLen = output.index.
RLT_HD bool rlt_forward_core(const uint8_t* src, int count, uint8_t* dst, int dstEnd, int escape, int* outLen) {
  int srcIdx = 0, dstIdx = 0;
  const int srcEnd = count, srcEnd4 = srcEnd - 4;
  bool res = true;
  int run = 0;
  int prev = src[srcIdx++];
  dst[dstIdx++] = (uint8_t)escape;
  dst[dstIdx++] = (uint8_t)prev;
  if (prev == escape) dst[dstIdx++] = 0;
  while (true) {
    if (prev == src[srcIdx]) {
      srcIdx++; run++;
      if (prev == src[srcIdx]) {
        srcIdx++; run++;
        if (prev == src[srcIdx]) {
          srcIdx++; run++;
          if (prev == src[srcIdx]) {
            srcIdx++; run++;
            if (run < RLT_MAX_RUN4 && srcIdx < srcEnd4) continue;
          }
        }
      }
    }
    if (run > RLT_RUN_THRESHOLD) {
      if (dstIdx + 6 >= dstEnd) { res = false; break; }
      dst[dstIdx++] = (uint8_t)prev;
      if (prev == escape) dst[dstIdx++] = 0;
      dst[dstIdx++] = (uint8_t)escape;
      dstIdx = rlt_emit_run_length(dst, dstIdx, run);
    } else if (prev != escape) {
      if (dstIdx + run >= dstEnd) { res = false; break; }
      while (run-- > 0) dst[dstIdx++] = (uint8_t)prev;
    } else {
      if (dstIdx + 2 * run >= dstEnd) { res = false; break; }
      while (run-- > 0) { dst[dstIdx++] = (uint8_t)escape; dst[dstIdx++] = 0; }
    }
    prev = src[srcIdx];
    srcIdx++;
    run = 1;
    if (srcIdx >= srcEnd4) break;
  }
  if (res) {
    if (prev != escape) {
      if (dstIdx + run < dstEnd) { while (run-- > 0) dst[dstIdx++] = (uint8_t)prev; }
    } else {
      if (dstIdx + 2 * run < dstEnd) { while (run-- > 0) { dst[dstIdx++] = (uint8_t)escape; dst[dstIdx++] = 0; } }
    }
    while (srcIdx < srcEnd && dstIdx < dstEnd) {
      if (src[srcIdx] == escape) {
        if (dstIdx + 2 >= dstEnd) { res = false; break; }
        dst[dstIdx++] = (uint8_t)escape; dst[dstIdx++] = 0;
        srcIdx++;
        continue;
      }
      dst[dstIdx++] = src[srcIdx++];
    }
    res = res && (srcIdx == srcEnd);
  }
  res = res && (dstIdx < srcIdx);
  *outLen = dstIdx;
  return res;
}

// RLT.inverse (:252-352) for slices with index 0; dstEnd = dst.array.length.  Where the Java code would throw (a one-byte
// input, a run before any byte was written) the block fails here too.  Returns the Java boolean; *outLen = output.index.
RLT_HD bool rlt_inverse_core(const uint8_t* src, int count, uint8_t* dst, int dstEnd, int* outLen) {
  *outLen = 0;
  if (count < 2 || dstEnd < 1) return false;
  int srcIdx = 0, dstIdx = 0;
  const int srcEnd = count;
  bool res = true;
  const int escape = src[srcIdx++];
  if (src[srcIdx] == escape) {
    srcIdx++;
    if (srcIdx < srcEnd && src[srcIdx] != 0) return false;
    dst[dstIdx++] = (uint8_t)escape;
    srcIdx++;
  }
  while (srcIdx < srcEnd) {
    const int b = src[srcIdx];
    if (b != escape) {
      if (dstIdx >= dstEnd) break;
      dst[dstIdx++] = (uint8_t)b;
      srcIdx++;
      continue;
    }
    srcIdx++;
    if (srcIdx >= srcEnd) { res = false; break; }
    if (dstIdx < 1) return false;
    const uint8_t val = dst[dstIdx - 1];
    int run = src[srcIdx++];
    if (run == 0) {
      if (dstIdx >= dstEnd) break;
      dst[dstIdx++] = (uint8_t)escape;
      continue;
    }
    if (run == 0xFF) {
      if (srcIdx >= srcEnd - 1) { res = false; break; }
      run = ((int)src[srcIdx] << 8) | (int)src[srcIdx + 1];
      srcIdx += 2;
      run += RLT_RUN_LEN_ENCODE2;
    } else if (run >= RLT_RUN_LEN_ENCODE1) {
      if (srcIdx >= srcEnd) { res = false; break; }
      run = ((run - RLT_RUN_LEN_ENCODE1) << 8) | (int)src[srcIdx++];
      run += RLT_RUN_LEN_ENCODE1;
    }
    run += (RLT_RUN_THRESHOLD - 1);
    if (dstIdx + run > dstEnd || run > RLT_MAX_RUN) { res = false; break; }
    while (run-- > 0) dst[dstIdx++] = val;
  }
  res = res && (srcIdx == srcEnd);
  *outLen = dstIdx;
  return res;
}
