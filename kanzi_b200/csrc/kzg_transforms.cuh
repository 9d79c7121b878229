// kzg_transforms.cuh — parameter block shared by the transform-stage kernels (lz.cu, rolz.cu, bwt.cu,
// sbrt.cu, srt.cu, zrlt.cu) and the Sequence bookkeeping (container.cu).
#pragma once
#include "kzg_common.cuh"

#define LZF_SAMPLES 8       // samples per block the LZ forward orders its blocks by: the first LZF_SAMPLE bytes of every eighth
#define LZF_SAMPLE 4096
#define KZG_XF_LZI_EXCLUSIVE (1 << 24)   // KzgXfParams.flags: LZ inverse, see lzi_tok_chase_kernel's launch

struct KzgXfParams {
  int* result;            // [2 * nBlocks]: {boolean result of forward()/inverse(), bytes produced}
  const u8* enabled;      // [nBlocks]: stage runs for this block (preconditions / skip flags, host-evaluated)
  const int* dstLimit;    // [nBlocks]: dst.array.length (inverse) or dst slice length (forward) as Java sees it
  u8* scratch; i64 scratchStride;      // per-block scratch area
  int tkStride, mStride, mLenStride;   // LZ: sub-buffer capacities inside the scratch area
  i32* hashBuf;                        // LZ/ROLZ: global hash / match tables
  i32* aux32; i64 aux32Stride;         // BWT etc.: per-block 32-bit scratch (u32 units)
  int flags;
  // optional (first stage of a host-buffer encode): block b's input, bytes [b * lazyBlock, min(lazyN, (b + 1) * lazyBlock)), is still on
  // the host at lazyHost and belongs at lazyDev; only a 4 KiB sample at the start of each eighth of every block has been
  // uploaded.  The stage uploads the blocks itself, on the streams it deals them to, so the copies overlap its kernels.
  const u8* lazyHost; u8* lazyDev; i64 lazyN; i32 lazyBlock;
};

int kzg_lz_inverse_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, i32 maxLen);
void kzg_lzi_scratch(i32 maxLen, size_t* perBlockBytes, size_t* aux32);
int kzg_lz_forward2_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, bool extra, i32 maxLen);
void kzg_lzf_scratch(i32 maxLen, size_t* perBlockBytes);
void kzg_lzf_release();
void kzg_count_launch(int n);
