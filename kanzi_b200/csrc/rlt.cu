// rlt.cu — RLT (K/transform/RLT.java; TransformFactory.RLT_TYPE = 5) as a transform stage.
// One warp per block.  When ctx["entropy"] asks for the best escape byte (anything but NONE / ANS0 / HUFFMAN / RANGE,
// RLT.java:101-107) the 32 lanes histogram the block in shared memory and lane 0 picks the rarest byte and, if the block's data
// type is still undefined, classifies it (Global.detectSimpleType) — a DNA or BASE64 block is then left alone, as in the
// reference.  The scan itself is the reference's loop on lane 0 (rlt_core.cuh): its output depends on how that loop cuts runs.
#include "kzg_transforms.cuh"
#include "kzg_xf_kernels.cuh"
#include "rlt_core.cuh"

template <bool FORWARD>
__global__ void __launch_bounds__(32) rlt_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P) {
  __shared__ u32 freqs[256];
  __shared__ int sEscape, sStop;
  const int lane = threadIdx.x, b = blockIdx.x;
  KzgBlock& B = blocks[b];
  int* res = P.result + 2 * b;
  if (lane == 0) { res[0] = 0; res[1] = 0; }
  if (B.status != 0 || !P.enabled[b]) return;
  const int count = B.curLen;
  const u8* src = B.cur;
  u8* dst = B.alt;
  int outLen = 0;
  bool ok = false;
  if (FORWARD) {
    const int maxEnc = (count <= 512) ? count + 32 : count;                                // :355-357
    const int dstEnd = min(P.dstLimit[b], B.cap);
    if (count < 16 || P.dstLimit[b] < maxEnc || dstEnd < 16) return;                       // :71-80
    int dt = B.dataType;
    if (dt == KZG_DT_DNA || dt == KZG_DT_BASE64 || dt == KZG_DT_UTF8) return;              // :96-99
    const int e = ((P.flags >> 8) & 0xF) - 1;                                               // ctx["entropy"], absent = NONE
    const bool best = !(e < 0 || e == KZG_E_NONE || e == KZG_E_ANS0 || e == KZG_E_HUFFMAN || e == 4 /* RANGE */);
    int escape = RLT_DEFAULT_ESCAPE;
    if (best) {
      for (int i = lane; i < 256; i += 32) freqs[i] = 0;
      __syncwarp();
      for (int i = lane; i < count; i += 32) atomicAdd(&freqs[src[i]], 1u);
      __syncwarp();
      if (lane == 0) {
        int stop = 0;
        if (dt == KZG_DT_UNDEFINED) {
          dt = rlt_detect_type(count, freqs);
          if (dt != KZG_DT_UNDEFINED) B.dataType = dt;
          if (dt == KZG_DT_DNA || dt == KZG_DT_BASE64 || dt == KZG_DT_UTF8) stop = 1;
        }
        sStop = stop;
        sEscape = rlt_best_escape(freqs);
      }
      __syncwarp();
      if (sStop) return;
      escape = sEscape;
    }
    if (lane != 0) return;
    ok = rlt_forward_core(src, count, dst, dstEnd, escape, &outLen);
  } else {
    if (lane != 0) return;
    ok = rlt_inverse_core(src, count, dst, min(kzg_dst_limit(B, P.dstLimit[b]), B.cap), &outLen);
  }
  res[0] = ok ? 1 : 0;
  res[1] = ok ? outLen : 0;
}

int kzg_rlt_launch(cudaStream_t s, bool forward, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, i32 maxLen) {
  (void)maxLen;
  if (forward) KZG_PROF("rlt_forward_kernel", s, (rlt_kernel<true><<<nBlocks, 32, 0, s>>>(d_blocks, P)));
  else KZG_PROF("rlt_inverse_kernel", s, (rlt_kernel<false><<<nBlocks, 32, 0, s>>>(d_blocks, P)));
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(1);
  return 0;
}
