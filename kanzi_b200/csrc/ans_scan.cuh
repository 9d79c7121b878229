// ans_scan.cuh — chunk-header walk of an ANS payload (shared by ans.cu and rolz.cu)
#pragma once
#include "kzg_common.cuh"
#include "kzg_entropy.cuh"

__device__ __forceinline__ int ans_skip_alphabet(BitReaderD& br) {   // EntropyUtils.decodeAlphabet sizes only
  if (br.read(1) == 0) return (br.read(1) == 1) ? 0 : 256;
  const int lastMask = (int)br.read(5);
  int count = 0;
  for (int i = 0; i <= lastMask; i++) count += __popc(br.read(8));
  return count;
}

// Walks the chunks of one ANSRangeDecoder.decode call (ANSRangeDecoder.java:189-236) of `len` bytes starting at
// br.pos; fills ci[0..nChunks).  Returns 0, or a negative status when the decoder would bail out.
__device__ __forceinline__ int ans_scan_stream(BitReaderD& br, int len, int chunkSize, int order, KzgChunkInfo* ci) {
  if (len <= 32) {   // raw (ANSRangeDecoder.decode :193-196)
    br.pos += (u64)len * 8;
    return br.overrun() ? -KZG_ERR_PROCESS_BLOCK : 0;
  }
  const int nChunks = (len + chunkSize - 1) / chunkSize;
  const int dim = 255 * order + 1;
  for (int c = 0; c < nChunks; c++) {
    KzgChunkInfo info;
    info.hdrBit = (i64)br.pos;
    const int lr = 8 + (int)br.read(3);
    int llr = 3;
    while ((1 << llr) <= lr) llr++;
    int res = 0;
    for (int k = 0; k < dim; k++) {
      const int alphabetSize = ans_skip_alphabet(br);
      if (alphabetSize == 0) continue;
      const int chkSize = (alphabetSize >= 64) ? 8 : 6;
      for (int i = 1; i < alphabetSize; i += chkSize) {
        const int logMax = (int)br.read(llr);
        const int endj = (i + chkSize < alphabetSize) ? i + chkSize : alphabetSize;
        br.pos += (u64)(logMax * (endj - i));
      }
      res += alphabetSize;
      if (br.overrun()) break;
    }
    info.alphabetSize = res;
    info.sz = 0; info.payBit = 0;
    info.st[0] = info.st[1] = info.st[2] = info.st[3] = 0;
    if (res == 0 || br.overrun()) return -KZG_ERR_PROCESS_BLOCK;   // decode returns early (:218-219)
    if (!(order == 0 && res == 1)) {
      const i32 sz = read_varint(br);
      if (sz < 0 || sz >= (1 << 27)) return -KZG_ERR_PROCESS_BLOCK;
      info.st[0] = br.read(32); info.st[1] = br.read(32); info.st[2] = br.read(32); info.st[3] = br.read(32);
      info.sz = sz;
      info.payBit = (i64)br.pos;
      br.pos += (u64)sz * 8;
    }
    if (br.overrun()) return -KZG_ERR_PROCESS_BLOCK;
    ci[c] = info;
  }
  return 0;
}
