// lzp.cu — LZPCodec (K/transform/LZCodec.java:973-1287; TransformFactory.LZP_TYPE = 14) as a transform stage.
// One warp per block: the lanes clear the block's 65536-entry table, lane 0 runs the chain (lzp_core.cuh says why it is a
// chain and what is done about its latency).  The table lives in the stage's hash area (256 KiB per block, L2-resident).
#include "kzg_transforms.cuh"
#include "kzg_xf_kernels.cuh"
#include "lzp_core.cuh"

template <bool FORWARD>
__global__ void __launch_bounds__(32) lzp_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P) {
  const int lane = threadIdx.x, b = blockIdx.x;
  KzgBlock& B = blocks[b];
  int* res = P.result + 2 * b;
  if (lane == 0) { res[0] = 0; res[1] = 0; }
  if (B.status != 0 || !P.enabled[b]) return;
  const int count = B.curLen;
  i32* hashes = P.hashBuf + (size_t)b * LZP_TABLE_INTS;
  if (FORWARD) {                                                                         // :1013-1018
    const int maxEnc = (count <= 1024) ? count + 16 : count + count / 64;
    if (P.dstLimit[b] < maxEnc || count < LZP_MIN_BLOCK || B.cap < count) return;
  }
  else if (count < 1) return;
  for (int i = lane; i < LZP_TABLE_INTS; i += 32) hashes[i] = 0;
  __syncwarp();
  if (lane != 0) return;
  int outLen = 0;
  bool ok;
  if (FORWARD) ok = lzp_forward_core(B.cur, count, B.alt, hashes, &outLen);
  else ok = lzp_inverse_core(B.cur, count, B.alt, min(kzg_dst_limit(B, P.dstLimit[b]), B.cap), hashes, &outLen);
  res[0] = ok ? 1 : 0;
  res[1] = ok ? outLen : 0;
}

void kzg_lzp_scratch(i32 maxLen, bool forward, size_t* perBlockBytes, size_t* hashInts) {
  (void)maxLen; (void)forward; (void)perBlockBytes;
  *hashInts = std::max(*hashInts, (size_t)LZP_TABLE_INTS);
}

int kzg_lzp_launch(cudaStream_t s, bool forward, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, i32 maxLen) {
  (void)maxLen;
  if (forward) KZG_PROF("lzp_forward_kernel", s, (lzp_kernel<true><<<nBlocks, 32, 0, s>>>(d_blocks, P)));
  else KZG_PROF("lzp_inverse_kernel", s, (lzp_kernel<false><<<nBlocks, 32, 0, s>>>(d_blocks, P)));
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(1);
  return 0;
}
