// kzg_entropy.cuh — parameter blocks shared by the entropy-stage kernels (ans.cu, huffman.cu, fpaq.cu)
// and the container assembly (container.cu).
#pragma once
#include "kzg_common.cuh"

// what the chunk scan learns about one chunk of an entropy payload (decode side)
struct KzgChunkInfo {
  i64 hdrBit;         // absolute bit offset of the chunk header in the stream
  i64 payBit;         // absolute bit offset of the chunk's coded bytes / first fragment
  i32 sz;             // ANS: coded byte count
  i32 alphabetSize;   // sum over contexts
  u32 st[4];          // ANS: initial states st0..st3; Huffman: fragment bit lengths
};

struct KzgEntParams {
  int entropy;        // KZG_E_* this launch serves (blocks with another id are ignored)
  int chunkSize;      // bytes per chunk (ANS0/Huffman 16384, ANS1/FPAQ 4 MiB)
  int maxChunks;      // chunks per block the per-block arrays are strided by
  // encode side: per-chunk scratch and the segment list (segsPerChunk entries per chunk)
  u8* hdrBuf; int hdrStride;
  u8* payBuf; int payStride;
  u32* tabBuf; i64 tabStride;     // u32 units (order-1 tables)
  const int* slotBase;            // optional: first scratch slot of block b (default b * maxChunks); lets sparse users (ROLZ) pack slots
  KzgSeg* segs; int segsPerBlock;  // block b, chunk c, k-th segment: segs[b*segsPerBlock + 1 + c*segsPerChunk + k] (slot 0 = raw copy)
  // decode side
  const u8* stream;               // compressed stream (device), >= 16 bytes of slack after the end
  KzgChunkInfo* chunks;
};

void kzg_count_launch(int n);

int kzg_ans_encode_launch(cudaStream_t s, const KzgBlock* d_blocks, int nBlocks, const KzgEntParams& P, int order);
int kzg_ans_decode_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, const KzgEntParams& P, int order, bool withScan = true);
size_t kzg_ans1_enc_tab_u32();
size_t kzg_ans1_dec_tab_u32();

int kzg_huff_encode_launch(cudaStream_t s, const KzgBlock* d_blocks, int nBlocks, const KzgEntParams& P);
int kzg_huff_decode_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, const KzgEntParams& P);

int kzg_fpaq_encode_launch(cudaStream_t s, const KzgBlock* d_blocks, int nBlocks, const KzgEntParams& P);
int kzg_fpaq_decode_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, const KzgEntParams& P);
int kzg_bitcopy_launch(cudaStream_t s, const KzgSeg* segs, i64 nSegs);
