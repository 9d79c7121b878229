// TEST INFRASTRUCTURE: kanzi_b200/csrc/lzp_core.cuh (the loops lane 0 of lzp_kernel runs) compiled for the host, so that the CPU
// test suite can hold the very same source against the oracle (tests/test_lzp_hostcheck.py).  Not part of the product: the
// library exports nothing of this and never runs a codec on the host.
#include "../../kanzi_b200/csrc/lzp_core.cuh"
#include <vector>

extern "C" int lzp_host_forward(const uint8_t* src, int count, uint8_t* dst, int* outLen) {
  std::vector<int32_t> hashes(LZP_TABLE_INTS, 0);
  return lzp_forward_core(src, count, dst, hashes.data(), outLen) ? 1 : 0;
}
extern "C" int lzp_host_inverse(const uint8_t* src, int count, uint8_t* dst, int dstEnd, int* outLen) {
  std::vector<int32_t> hashes(LZP_TABLE_INTS, 0);
  return lzp_inverse_core(src, count, dst, dstEnd, hashes.data(), outLen) ? 1 : 0;
}
