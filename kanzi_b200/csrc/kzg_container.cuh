// kzg_container.cuh — launchers of container.cu
#pragma once
#include "kzg_common.cuh"
int kzg_commit_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, const int* result, const u8* enabled, int stage, int forward);
// container != 0: d_out receives the whole .knz bit stream (stream header bytes must already be at its
// start, rest zero-filled); d_totalBits[0] = stream bit length or -1.  container == 0: each block's
// entropy payload is written at bit blocks[b].srcBit of d_out (caller presets srcBit) — no headers.
int kzg_assemble_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, KzgSeg* segs, int segsPerBlock, u8* hdrBytes, int nbFunctions,
                        int container, u8* d_out, i64 headerBits, i64* d_totalBits, i64 capBytes);
#define KZG_HDR_STRIDE 16      // bytes of block header staging per block: mode + skip flags + 4 length bytes + checksum byte + XXHash (<= 8)
int kzg_xxh_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, int verify);
