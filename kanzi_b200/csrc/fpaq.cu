// fpaq.cu — FPAQ (adaptive order-0 binary arithmetic coder) kernels (sm_100a).
//
// Replaces K/entropy/FPAQEncoder.java and FPAQDecoder.java.  The format leaves no
// parallelism inside a block: probabilities, low/high (and the decoder's `current`) carry across the
// 4 MiB chunks of a block (FPAQEncoder.java:140-170), every bit depends on the previous one.  One thread
// per block runs the coder (probability tables in shared memory); parallelism = blocks in flight.
// This stage is latency-bound by construction; it is here for bit-exact coverage of the chain
// BWT+SRT+ZRLT&FPAQ, not for a roofline.
#include "kzg_common.cuh"
#include "kzg_entropy.cuh"

#define FQ_TOP 0x00FFFFFFFFFFFFFFull
#define FQ_MASK_24_56 0x00FFFFFFFF000000ull
#define FQ_MASK_0_24 0x0000000000FFFFFFull
#define FQ_MASK_0_32 0x00000000FFFFFFFFull
#define FQ_MASK_0_56 0x00FFFFFFFFFFFFFFull
#define FQ_PSCALE 65536
#define FQ_CHUNK (4 << 20)

// ---- encode: grid nBlocks, 32 threads (lane 0 works); per chunk 2 segments: varint | bytes + 56-bit flush ----------
__global__ void __launch_bounds__(32) fpaq_encode_kernel(const KzgBlock* __restrict__ blocks, KzgEntParams P) {
  __shared__ int probs[4][256];
  const int b = blockIdx.x;
  const KzgBlock& B = blocks[b];
  for (int i = threadIdx.x; i < 1024; i += 32) (&probs[0][0])[i] = FQ_PSCALE >> 1;
  __syncwarp();
  if (threadIdx.x != 0) return;
  KzgSeg* segs = P.segs + (i64)b * P.segsPerBlock + 1;
  const int count = (B.status == 0 && B.entropy == P.entropy) ? B.curLen : 0;
  const u8* __restrict__ data = B.cur;
  u64 low = 0, high = FQ_TOP;
  int c = 0;
  for (int startChunk = 0; c < P.maxChunks; c++, startChunk += FQ_CHUNK) {
    if (startChunk >= count) { segs[2 * c] = KzgSeg{nullptr, 0, 0, 0}; segs[2 * c + 1] = KzgSeg{nullptr, 0, 0, 0}; continue; }
    const int chunkSize = min(FQ_CHUNK, count - startChunk);
    const i64 gidx = (i64)b * P.maxChunks + c;
    u8* hdr = P.hdrBuf + gidx * (i64)P.hdrStride;
    u8* sba = P.payBuf + gidx * (i64)P.payStride;
    const int sbaCap = chunkSize + (chunkSize >> 3);      // Java array size (:145-146); overflow = ArrayIndexOutOfBounds
    int idx = 0;
    bool overflow = false;
    int* p = probs[0];
    int val = data[startChunk];
    for (int i = startChunk; i < startChunk + chunkSize; i++) {
      const int nxt = (i + 1 < startChunk + chunkSize) ? (int)data[i + 1] : 0;      // (the next byte's load overlaps this byte's bits)
      // the byte's path through the binary context tree is known up front: all eight probabilities are loaded before the arithmetic
      // (no two nodes of one path coincide; the previous byte's updates precede these loads in program order)
      int pr[8];
      #pragma unroll
      for (int k = 0; k < 8; k++) pr[k] = p[(val | 0x100) >> (8 - k)];
      #pragma unroll
      for (int k = 7; k >= 0; k--) {
        const int bit = (val >> k) & 1;
        const int pv = pr[7 - k];
        // encodeBit (:182-199)
        const u64 split = (((high - low) >> 8) * (u64)pv) >> 8;
        if (bit == 0) { low += split + 1; p[(val | 0x100) >> (k + 1)] = pv - (pv >> 6); }
        else { high = low + split; p[(val | 0x100) >> (k + 1)] = pv - ((pv - FQ_PSCALE + 64) >> 6); }
        while (((low ^ high) & FQ_MASK_24_56) == 0) {       // flush (:208-213)
          if (idx + 4 > sbaCap) { overflow = true; idx = 0; }
          const u32 w = (u32)(high >> 24);
          *reinterpret_cast<u32*>(sba + idx) = __byte_perm(w, 0, 0x0123);      // big-endian, idx stays a multiple of 4
          idx += 4;
          low <<= 32;
          high = (high << 32) | FQ_MASK_0_32;
        }
      }
      p = probs[val >> 6];
      val = nxt;
    }
    if (overflow) { atomicExch((int*)&blocks[b].status, -KZG_ERR_PROCESS_BLOCK); }
    BitWriterD bw(hdr);
    write_varint(bw, idx);
    bw.flush();
    // 56 bits of (low | MASK_0_24): between chunks (:168-169) and from dispose() after the last one (:232-238)
    const u64 fl = (low | FQ_MASK_0_24) & FQ_MASK_0_56;
    for (int k = 0; k < 7; k++) sba[idx + k] = (u8)(fl >> (48 - 8 * k));
    segs[2 * c] = KzgSeg{hdr, 0, 0, (u64)bw.bits()};
    segs[2 * c + 1] = KzgSeg{sba, 0, 0, (u64)(idx + 7) * 8};
  }
}

// ---- decode: one thread per block -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) fpaq_decode_kernel(KzgBlock* __restrict__ blocks, KzgEntParams P) {
  __shared__ int probs[4][256];
  const int b = blockIdx.x;
  KzgBlock& B = blocks[b];
  for (int i = threadIdx.x; i < 1024; i += 32) (&probs[0][0])[i] = FQ_PSCALE >> 1;
  __syncwarp();
  if (threadIdx.x != 0) return;
  if (!(B.status == 0 && B.entropy == P.entropy)) return;
  const int count = B.preLen;
  u8* __restrict__ out = B.cur;
  const u8* __restrict__ stream = P.stream;
  BitReaderD br(stream, (u64)B.srcBit, (u64)(B.srcBit + B.srcBits));
  u64 low = 0, high = FQ_TOP, current = 0;
  for (int startChunk = 0; startChunk < count; startChunk += FQ_CHUNK) {
    const i32 szBytes = read_varint(br);
    if (szBytes < 0 || szBytes >= 2 * count) { B.status = -KZG_ERR_PROCESS_BLOCK; return; }       // sanity check (:176-178)
    current = ((u64)br.read(24) << 32) | (u64)br.read(32);
    const u64 payBit = br.pos;
    br.pos += (u64)szBytes * 8;
    if (br.overrun()) { B.status = -KZG_ERR_PROCESS_BLOCK; return; }
    int idx = 0;
    const int chunkSize = min(FQ_CHUNK, count - startChunk);
    int* p = probs[0];
    // the next 32 coded bits are fetched one refill ahead (a refill never waits for global memory)
    u32 ahead = (szBytes >= 4) ? get_bits(stream, payBit, 32) : 0u;
    int pv = p[1];                       // probability of the node about to be decoded
    for (int i = startChunk; i < startChunk + chunkSize; i++) {
      int ctx = 1;
      #pragma unroll
      for (int k = 0; k < 8; k++) {
        // both children's probabilities are requested before the bit is known: the shared-memory latency overlaps the arithmetic
        // (at the last level the only successor is the root of the next byte's table, fetched after this byte's updates below)
        int c0 = 0, c1 = 0;
        if (k < 7) { c0 = p[2 * ctx]; c1 = p[2 * ctx + 1]; }
        // decodeBitV2 (:290-314)
        const u64 split = ((((high - low) >> 8) * (u64)pv) >> 8) + low;
        if (split >= current) { high = split; p[ctx] = pv - ((pv - FQ_PSCALE + 64) >> 6); ctx = (ctx << 1) + 1; pv = c1; }
        else { low = split + 1; p[ctx] = pv - (pv >> 6); ctx = ctx << 1; pv = c0; }
        while (((low ^ high) & FQ_MASK_24_56) == 0) {       // read (:322-335)
          low = (low << 32) & FQ_MASK_0_56;
          high = ((high << 32) | FQ_MASK_0_32) & FQ_MASK_0_56;
          if (idx + 4 > szBytes) { current = (current << 32) & FQ_MASK_0_56; idx = szBytes + 1; }
          else {
            current = ((current << 32) | (u64)ahead) & FQ_MASK_0_56; idx += 4;
            if (idx + 4 <= szBytes) ahead = get_bits(stream, payBit + 8ull * idx, 32);
          }
        }
      }
      out[i] = (u8)ctx;
      if (idx > szBytes) { B.status = -KZG_ERR_PROCESS_BLOCK; return; }
      p = probs[(ctx & 0xFF) >> 6];
      pv = p[1];
    }
  }
  B.entBits = (i64)br.pos - B.srcBit;
}

int kzg_fpaq_encode_launch(cudaStream_t s, const KzgBlock* d_blocks, int nBlocks, const KzgEntParams& P) {
  fpaq_encode_kernel<<<nBlocks, 32, 0, s>>>(d_blocks, P);
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(1);
  return 0;
}
int kzg_fpaq_decode_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, const KzgEntParams& P) {
  fpaq_decode_kernel<<<nBlocks, 32, 0, s>>>(d_blocks, P);
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(1);
  return 0;
}
