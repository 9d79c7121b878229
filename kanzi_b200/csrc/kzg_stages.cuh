// kzg_stages.cuh — dispatch of the transform stages other than LZ, plus small helper launches
#pragma once
#include "kzg_common.cuh"
#include "kzg_transforms.cuh"

// adds the per-block scratch a stage needs (bytes, hash ints, aux 32-bit words)
void kzg_stage_scratch(int type, i32 maxLen, bool forward, size_t* perBlockBytes, size_t* hashInts, size_t* aux32);
// launches the stage kernel(s) for every enabled block; results in P.result
int kzg_stage_launch(cudaStream_t s, int type, bool forward, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, i32 maxLen);
// length-only part of a codec's guard block for a single call with (src: length srcLen, index 0), (dst: length dstLen, array dstCap).
// returns 1 = go on, 0 = the Java call returns false, < 0 error
int kzg_stage_precheck(int type, bool forward, const kzg_ctx* ctx, i32 srcLen, i32 dstLen, i32 dstCap);
// raw BWT entry points of include/kzg.h
int kzg_bwt_raw(cudaStream_t s, bool forward, const u8* src, i32 n, u8* dst, i32* primaryIndexes8);

int kzg_rawbits_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, const u8* d_stream);
int kzg_magic_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks);
int kzg_null_forward_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, const u8* enabled, int stage);
