// TEST INFRASTRUCTURE: kanzi_b200/csrc/lzp_core.cuh, rlt_core.cuh and rolzx_core.cuh (the loops lane 0 of lzp_kernel / rlt_kernel / rolzx_kernel runs) compiled for the
// host, so that the CPU test suite can hold the very same source against the oracle (tests/test_sibling_hostcheck.py).  Not part of the product: the
// library exports nothing of this and never runs a codec on the host.
#include "../../kanzi_b200/csrc/lzp_core.cuh"
#include "../../kanzi_b200/csrc/rlt_core.cuh"
#include "../../kanzi_b200/csrc/rolzx_core.cuh"
#include <vector>

extern "C" int lzp_host_forward(const uint8_t* src, int count, uint8_t* dst, int* outLen) {
  std::vector<int32_t> hashes(LZP_TABLE_INTS, 0);
  return lzp_forward_core(src, count, dst, hashes.data(), outLen) ? 1 : 0;
}
extern "C" int lzp_host_inverse(const uint8_t* src, int count, uint8_t* dst, int dstEnd, int* outLen) {
  std::vector<int32_t> hashes(LZP_TABLE_INTS, 0);
  return lzp_inverse_core(src, count, dst, dstEnd, hashes.data(), outLen) ? 1 : 0;
}

// RLT.forward as rlt_kernel<true> runs it: best != 0 = ctx["entropy"] asks for the rarest byte as escape; *dataType in/out.
extern "C" int rlt_host_forward(const uint8_t* src, int count, uint8_t* dst, int dstEnd, int best, int* dataType, int* outLen) {
  *outLen = 0;
  int dt = *dataType;
  if (count < 16 || dstEnd < ((count <= 512) ? count + 32 : count)) return 0;
  if (dt == RLT_DT_DNA || dt == RLT_DT_BASE64 || dt == RLT_DT_UTF8) return 0;
  int escape = RLT_DEFAULT_ESCAPE;
  if (best) {
    std::vector<uint32_t> f(256, 0);
    for (int i = 0; i < count; i++) f[src[i]]++;
    if (dt == RLT_DT_UNDEFINED) {
      dt = rlt_detect_type(count, f.data());
      if (dt != RLT_DT_UNDEFINED) *dataType = dt;
      if (dt == RLT_DT_DNA || dt == RLT_DT_BASE64 || dt == RLT_DT_UTF8) return 0;
    }
    escape = rlt_best_escape(f.data());
  }
  return rlt_forward_core(src, count, dst, dstEnd, escape, outLen) ? 1 : 0;
}
extern "C" int rlt_host_inverse(const uint8_t* src, int count, uint8_t* dst, int dstEnd, int* outLen) {
  return rlt_inverse_core(src, count, dst, dstEnd, outLen) ? 1 : 0;
}

// ROLZX as rolzx_kernel drives it: chunk loop around rzx_forward_chunk / rzx_inverse_chunk.  dataType: DataType ordinal the block
// was classified as (the kernel's histogram step; the test passes the oracle's answer).  Returns 1 / 0, -1 = coder overran dst.
extern "C" int rolzx_host_forward(const uint8_t* src, int count, uint8_t* dst, int limit, int dataType, int* outLen) {
  *outLen = 0;
  if (count < 64) return 0;
  int mm = 3, dt = 2, flags = 0;
  if (dataType == 3) { dt = 3; flags |= 8; } else if (dataType == 6) { dt = 8; mm = 7; flags |= 4; }
  std::vector<uint16_t> probs(RZX_LIT_CELLS + RZX_MATCH_CELLS, 0x7FFF);
  std::vector<int32_t> tab(RZX_MATCH_INTS + RZX_HASH_SIZE + 4, 0);
  int32_t* matches = (int32_t*)(((uintptr_t)tab.data() + 15) & ~(uintptr_t)15);
  int32_t* counters = matches + RZX_MATCH_INTS;
  dst[0] = (uint8_t)(count >> 24); dst[1] = (uint8_t)(count >> 16); dst[2] = (uint8_t)(count >> 8); dst[3] = (uint8_t)count; dst[4] = (uint8_t)flags;
  RzxCoder C;
  rzx_coder_init(C, probs.data(), probs.data() + RZX_LIT_CELLS, dst, 5, limit);
  const int total = count - 4, sizeChunk = count < RZX_CHUNK ? count : RZX_CHUNK;
  for (int start = 0; start < total;) {
    memset(matches, 0, RZX_MATCH_INTS * 4);
    const int end = (start + sizeChunk < total) ? start + sizeChunk : total;
    rzx_forward_chunk(src, start, end, total, C, matches, counters, mm, dt);
    start = end;
  }
  rzx_forward_tail(src, total, C);
  if (C.overrun) return -1;
  *outLen = C.index;
  return 1;
}
extern "C" int rolzx_host_inverse(const uint8_t* src, int count, uint8_t* dst, int dstLen, int* outLen) {
  *outLen = 0;
  if (count < 13) return 0;
  const int sz = (int)(((uint32_t)src[0] << 24) | ((uint32_t)src[1] << 16) | ((uint32_t)src[2] << 8) | (uint32_t)src[3]);
  if (sz <= 0 || sz > dstLen) return 0;
  int mm = 3, dt = 2;
  const int flags = src[4];
  if ((flags & 0x0E) == 8) dt = 3; else if ((flags & 0x0E) == 4) { dt = 8; mm = 7; }
  std::vector<uint16_t> probs(RZX_LIT_CELLS + RZX_MATCH_CELLS, 0x7FFF);
  std::vector<int32_t> tab(RZX_MATCH_INTS + RZX_HASH_SIZE + 4, 0);
  int32_t* matches = (int32_t*)(((uintptr_t)tab.data() + 15) & ~(uintptr_t)15);
  int32_t* counters = matches + RZX_MATCH_INTS;
  RzxCoder C;
  rzx_coder_init(C, probs.data(), probs.data() + RZX_LIT_CELLS, const_cast<uint8_t*>(src), 5, count);
  rzx_decoder_start(C);
  const int sizeChunk = sz < RZX_CHUNK ? sz : RZX_CHUNK;
  int outIndex = 0;
  for (int start = 0; start < sz;) {
    memset(matches, 0, RZX_MATCH_INTS * 4);
    const int end = (start + sizeChunk < sz) ? start + sizeChunk : sz;
    if (!rzx_inverse_chunk(dst, start, end, sz, dstLen, &outIndex, C, matches, counters, mm, dt)) return 0;
    start = end;
  }
  if (C.overrun || C.index != count) return 0;
  *outLen = outIndex;
  return 1;
}
