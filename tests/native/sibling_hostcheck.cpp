// TEST INFRASTRUCTURE: kanzi_b200/csrc/lzp_core.cuh and rlt_core.cuh (the loops lane 0 of lzp_kernel / rlt_kernel runs) compiled for the
// host, so that the CPU test suite can hold the very same source against the oracle (tests/test_sibling_hostcheck.py).  Not part of the product: the
// library exports nothing of this and never runs a codec on the host.
#include "../../kanzi_b200/csrc/lzp_core.cuh"
#include "../../kanzi_b200/csrc/rlt_core.cuh"
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
