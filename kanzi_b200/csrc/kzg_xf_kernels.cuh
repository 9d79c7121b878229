// kzg_xf_kernels.cuh — launchers of the transform kernels other than LZ (zrlt.cu, sbrt.cu, srt.cu, bwt.cu, rolz.cu)
#pragma once
#include "kzg_common.cuh"
#include "kzg_transforms.cuh"

void kzg_bwt_scratch(i32 maxLen, bool forward, size_t* perBlockBytes, size_t* aux32);
void kzg_rolz_scratch(i32 maxLen, bool forward, size_t* perBlockBytes, size_t* hashInts, size_t* aux32);
void kzg_small_scratch(int type, i32 maxLen, bool forward, size_t* perBlockBytes, size_t* aux32);
int kzg_zrlt_launch(cudaStream_t s, bool forward, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, i32 maxLen);
int kzg_sbrt_launch(cudaStream_t s, bool forward, int mode, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, i32 maxLen);
int kzg_srt_launch(cudaStream_t s, bool forward, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, i32 maxLen);
int kzg_bwtblock_launch(cudaStream_t s, bool forward, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, i32 maxLen);
void kzg_lzp_scratch(i32 maxLen, bool forward, size_t* perBlockBytes, size_t* hashInts);
int kzg_lzp_launch(cudaStream_t s, bool forward, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, i32 maxLen);
void kzg_rolzx_scratch(i32 maxLen, bool forward, size_t* perBlockBytes, size_t* hashInts);
int kzg_rolzx_launch(cudaStream_t s, bool forward, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, i32 maxLen);
int kzg_rlt_launch(cudaStream_t s, bool forward, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, i32 maxLen);
int kzg_rolz_launch(cudaStream_t s, bool forward, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, i32 maxLen);
