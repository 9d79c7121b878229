// container.cu — Sequence bookkeeping, block-record assembly and bit-granular concatenation (sm_100a).
//
// What K/transform/Sequence.java (skip flags, ping-pong slices) and the tail of EncodingTask.encodeBlock
// (COS:861-1035: block header, "transformed copy" fallback, 8-bit header checksum, 5-bit length-of-length +
// length + payload at an arbitrary bit offset) do per block is done here for every block of a batch at
// once.  Chunk bit strings produced by the entropy kernels are stitched into the final .knz bit stream by
// one HBM-bound pass (kzg_bitcopy_kernel): each 32-bit destination word is built from two source words
// with a funnel shift; words shared by two segments are merged with atomicOr into a zero-filled stream.
#include "kzg_common.cuh"
#include "kzg_entropy.cuh"
#include "kzg_transforms.cuh"
#include "kzg_container.cuh"

// ---- Sequence.forward / inverse bookkeeping after one transform stage -----------------------------------
// forward (Sequence.java:95-114): success -> clear skip bit 7-i, swap roles; failure -> input passes through.
// inverse (Sequence.java:168-184): failure is fatal for the block.
__global__ void kzg_commit_kernel(KzgBlock* __restrict__ blocks, int nBlocks, const int* __restrict__ result,
                                  const u8* __restrict__ enabled, int stage, int forward) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nBlocks) return;
  KzgBlock& B = blocks[b];
  if (B.status != 0 || !enabled[b]) return;
  const int ok = result[2 * b], len = result[2 * b + 1];
  if (ok) {
    u8* t = B.cur; B.cur = B.alt;
    B.curLen = len;
    if (forward) {
      B.alt = (t == B.aux0) ? B.aux1 : t;          // never write into the caller's input block
      B.skipFlags &= ~(1 << (7 - stage));
    } else {
      B.stagesLeft--;
      B.alt = (B.stagesLeft == 1) ? B.aux0 : t;    // the last inverse stage writes the block's final destination
    }
  } else if (!forward) {
    B.status = -KZG_ERR_PROCESS_BLOCK;      // "Transform inverse failed" (CIS:1342-1344)
  }
}

// ---- block checksums: K/util/hash/XXHash32.java:94-142 (the published XXH32) and K/util/hash/XXHash64.java (Kanzi's own variant:
// the accumulators are folded with `(v << 1) | (v >>> 31)` on 64-bit values, :127-128), seed = 0x4B414E5A (COS:195-200).
// The four accumulators are four dependent chains over the block: lanes 0-3 run one each, lane 0 finishes.  verify == 0:
// hash B.aux0[0, origLen) into B.xxh (encode: the original bytes); verify != 0: hash the decoded block and compare (CIS:1348-1370).
__device__ __forceinline__ u32 xxh32_round(u32 acc, u32 val) { acc += val * 2246822519u; return ((acc << 13) | (acc >> 19)) * 2654435761u; }
__device__ __forceinline__ u64 xxh64_round(u64 acc, u64 val) { acc += val * 0xC2B2AE3D27D4EB4Full; return ((acc << 31) | (acc >> 33)) * 0x9E3779B185EBCA87ull; }
__device__ __forceinline__ u64 xxh64_merge(u64 acc, u64 val) { acc ^= xxh64_round(0, val); return acc * 0x9E3779B185EBCA87ull + 0x85EBCA77C2B2AE63ull; }
__global__ void __launch_bounds__(32) kzg_xxh_kernel(KzgBlock* __restrict__ blocks, int verify) {
  const int b = blockIdx.x, lane = threadIdx.x;
  KzgBlock& B = blocks[b];
  if (B.status != 0 || B.chkBytes == 0) return;
  const u8* __restrict__ data = verify ? B.cur : B.aux0;
  const int length = verify ? B.curLen : B.origLen;
  const u32 SEED = 0x4B414E5Au;
  u64 h = 0;
  int idx = 0;
  if (B.chkBytes == 4) {
    const u32 P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
    u32 h32;
    if (length >= 16) {
      const int stripes = length >> 4;
      u32 v = (lane == 0) ? SEED + P1 + P2 : ((lane == 1) ? SEED + P2 : ((lane == 2) ? SEED : SEED - P1));
      if (lane < 4) {
        const u8* p = data + 4 * lane;
        if ((reinterpret_cast<uintptr_t>(data) & 3) == 0) { const u32* q = reinterpret_cast<const u32*>(p); for (int s = 0; s < stripes; s++) v = xxh32_round(v, q[4 * s]); }
        else for (int s = 0; s < stripes; s++) { const u8* x = p + 16 * s; v = xxh32_round(v, (u32)x[0] | ((u32)x[1] << 8) | ((u32)x[2] << 16) | ((u32)x[3] << 24)); }
      }
      const u32 v1 = __shfl_sync(0xFFFFFFFFu, v, 0), v2 = __shfl_sync(0xFFFFFFFFu, v, 1), v3 = __shfl_sync(0xFFFFFFFFu, v, 2), v4 = __shfl_sync(0xFFFFFFFFu, v, 3);
      h32 = ((v1 << 1) | (v1 >> 31)) + ((v2 << 7) | (v2 >> 25)) + ((v3 << 12) | (v3 >> 20)) + ((v4 << 18) | (v4 >> 14));
      idx = stripes << 4;
    } else h32 = SEED + P5;
    if (lane != 0) return;
    h32 += (u32)length;
    while (idx <= length - 4) { const u8* x = data + idx; h32 += ((u32)x[0] | ((u32)x[1] << 8) | ((u32)x[2] << 16) | ((u32)x[3] << 24)) * P3; h32 = ((h32 << 17) | (h32 >> 15)) * P4; idx += 4; }
    while (idx < length) { h32 += (u32)data[idx] * P5; h32 = ((h32 << 11) | (h32 >> 21)) * P1; idx++; }
    h32 ^= h32 >> 15; h32 *= P2; h32 ^= h32 >> 13; h32 *= P3;
    h = (u64)(h32 ^ (h32 >> 16));
  } else {
    const u64 P1 = 0x9E3779B185EBCA87ull, P2 = 0xC2B2AE3D27D4EB4Full, P3 = 0x165667B19E3779F9ull, P4 = 0x85EBCA77C2B2AE63ull, P5 = 0x27D4EB2F165667C5ull;
    const u64 S64 = (u64)SEED;
    u64 h64;
    auto ld64 = [&](const u8* x) { u64 r = 0; for (int i = 7; i >= 0; i--) r = (r << 8) | x[i]; return r; };
    if (length >= 32) {
      const int stripes = length >> 5;
      u64 v = (lane == 0) ? S64 + P1 + P2 : ((lane == 1) ? S64 + P2 : ((lane == 2) ? S64 : S64 - P1));
      if (lane < 4) {
        const u8* p = data + 8 * lane;
        if ((reinterpret_cast<uintptr_t>(data) & 7) == 0) { const u64* q = reinterpret_cast<const u64*>(p); for (int s = 0; s < stripes; s++) v = xxh64_round(v, q[4 * s]); }
        else for (int s = 0; s < stripes; s++) v = xxh64_round(v, ld64(p + 32 * s));
      }
      const u64 v1 = __shfl_sync(0xFFFFFFFFu, v, 0), v2 = __shfl_sync(0xFFFFFFFFu, v, 1), v3 = __shfl_sync(0xFFFFFFFFu, v, 2), v4 = __shfl_sync(0xFFFFFFFFu, v, 3);
      h64 = ((v1 << 1) | (v1 >> 31)) + ((v2 << 7) | (v2 >> 25)) + ((v3 << 12) | (v3 >> 20)) + ((v4 << 18) | (v4 >> 14));      // as written (XXHash64.java:127-128)
      h64 = xxh64_merge(h64, v1); h64 = xxh64_merge(h64, v2); h64 = xxh64_merge(h64, v3); h64 = xxh64_merge(h64, v4);
      idx = stripes << 5;
    } else h64 = S64 + P5;
    if (lane != 0) return;
    h64 += (u64)(i64)length;
    while (idx + 8 <= length) { h64 ^= xxh64_round(0, ld64(data + idx)); h64 = ((h64 << 27) | (h64 >> 37)) * P1 + P4; idx += 8; }
    while (idx + 4 <= length) { const u8* x = data + idx; const i32 w = (i32)((u32)x[0] | ((u32)x[1] << 8) | ((u32)x[2] << 16) | ((u32)x[3] << 24)); h64 ^= (u64)(i64)w * P1; h64 = ((h64 << 23) | (h64 >> 41)) * P2 + P3; idx += 4; }
    while (idx < length) { h64 ^= (u64)data[idx] * P5; h64 = ((h64 << 11) | (h64 >> 53)) * P1; idx++; }
    h64 ^= h64 >> 33; h64 *= P2; h64 ^= h64 >> 29; h64 *= P3;
    h = h64 ^ (h64 >> 32);
  }
  if (!verify) B.xxh = h;
  else if (h != B.xxh) B.status = -KZG_ERR_CRC_CHECK;       // "Corrupted bitstream: invalid checksum" (CIS:1352-1370)
}
int kzg_xxh_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, int verify) {
  KZG_PROF("kzg_xxh_kernel", s, (kzg_xxh_kernel<<<nBlocks, 32, 0, s>>>(d_blocks, verify)));
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(1);
  return 0;
}

// ---- per-block record layout (one warp per block) --------------------------------------------------------------
// scans the block's chunk segments (relative bit offsets), decides the transformed-copy fallback
// (COS:926-973) and builds the block header bytes + checksum (COS:861-896, 977-985).
__global__ void __launch_bounds__(32) kzg_block_layout_kernel(KzgBlock* __restrict__ blocks, KzgSeg* __restrict__ segs,
                                                               int segsPerBlock, u8* __restrict__ hdrBytes, int nbFunctions, int container) {
  const int b = blockIdx.x, lane = threadIdx.x;
  KzgBlock& B = blocks[b];
  KzgSeg* S = segs + (i64)b * segsPerBlock;
  if (B.status != 0) { if (lane == 0) { B.written = 0; B.entBits = 0; } return; }
  // exclusive scan of segment bit lengths (segment 0 is reserved for a raw copy of the data)
  u64 running = 0;
  if (B.entropy == KZG_E_NONE) {
    // NullEntropyEncoder (K/entropy/NullEntropyEncoder.java:44-58): the bytes themselves
    if (lane == 0) S[0] = KzgSeg{B.cur, 0, 0, (u64)B.curLen * 8};
    for (int i = 1 + lane; i < segsPerBlock; i += 32) S[i].nBits = 0;
    running = (u64)B.curLen * 8;
  } else {
    if (lane == 0) S[0].nBits = 0;
    for (int base = 1; base < segsPerBlock; base += 32) {
      const int i = base + lane;
      const u64 n = (i < segsPerBlock) ? S[i].nBits : 0;
      u64 incl = n;
      for (int o = 1; o < 32; o <<= 1) { const u64 t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
      if (i < segsPerBlock) S[i].dstBit = running + incl - n;
      running += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
  }
  __syncwarp();
  if (lane != 0) return;
  B.entBits = (i64)running;
  if (!container) { B.written = (i64)running; B.hdrBytes = 0; return; }
  const int post = B.curLen;
  const int dataSize = (post < 256) ? 1 : (ilog2((u32)post) >> 3) + 1;
  int mode = B.mode & 0x80;                       // COPY_BLOCK_MASK preset by the host for blocks <= 15 bytes (COS:764-767)
  mode |= (((dataSize - 1) & 0x03) << 5);
  const int skipFlags = B.skipFlags & 0xFF;
  int headerSkipFlags = skipFlags;
  bool skipByte = false;
  if ((mode & 0x80) || (nbFunctions <= 4)) {
    mode |= (skipFlags >> 4);
    headerSkipFlags = (mode & 0x80) ? 0 : (((mode << 4) | 0x0F) & 0xFF);
  } else {
    mode |= 0x10; skipByte = true;
  }
  const int chk = B.chkBytes;                     // XXHash32 / 64 of the original bytes follows the header checksum (COS:892-895)
  int hb = 1 + (skipByte ? 1 : 0) + dataSize + 1 + chk;
  i64 written = (i64)hb * 8 + (i64)running;
  if (!(mode & 0x80)) {
    const i64 entropyPayloadBytes = (written + 7) >> 3;
    if ((i64)post < entropyPayloadBytes) {        // transformed copy (COS:926-973)
      mode |= 0x80 | 0x10;
      skipByte = (nbFunctions > 4);
      headerSkipFlags = skipByte ? skipFlags : (((mode << 4) | 0x0F) & 0xFF);
      hb = 1 + (skipByte ? 1 : 0) + dataSize + 1 + chk;
      S[0] = KzgSeg{B.cur, 0, 0, (u64)post * 8};
      for (int i = 1; i < segsPerBlock; i++) S[i].nBits = 0;
      written = (i64)hb * 8 + (i64)post * 8;
    }
  }
  u32 ck = 0x1E35A7BDu * 0x01030507u;
  ck = kzg_mix32(ck, 0x1E35A7BDu, (u32)(mode & 0xFF));
  ck = kzg_mix32(ck, 0x1E35A7BDu, (u32)(headerSkipFlags & 0xFF));
  ck = kzg_mix32(ck, 0x1E35A7BDu, (u32)post);
  ck = kzg_mix32(ck, 0x1E35A7BDu, (u32)((u64)written >> 32));
  ck = kzg_mix32(ck, 0x1E35A7BDu, (u32)written);
  ck = (ck >> 23) ^ (ck >> 3);
  u8* h = hdrBytes + b * KZG_HDR_STRIDE;
  int k = 0;
  h[k++] = (u8)mode;
  if (skipByte) h[k++] = (u8)skipFlags;
  for (int i = dataSize - 1; i >= 0; i--) h[k++] = (u8)(post >> (8 * i));
  h[k++] = (u8)ck;
  for (int i = chk - 1; i >= 0; i--) h[k++] = (u8)(B.xxh >> (8 * i));
  B.mode = mode & 0xFF; B.hdrBytes = hb; B.written = written;
}

__device__ __forceinline__ void put_bits_atomic(u32* __restrict__ dst32, u64 bitpos, u64 value, int n) {   // n <= 40
  for (int left = n; left > 0;) {
    const u64 w = bitpos >> 5;
    const int off = (int)(bitpos & 31);
    const int take = min(32 - off, left);
    const u32 v = (u32)((value >> (left - take)) & ((take == 32) ? 0xFFFFFFFFull : ((1ull << take) - 1)));
    atomicOr(&dst32[w], __byte_perm(v << (32 - off - take), 0, 0x0123));
    bitpos += take; left -= take;
  }
}

// ---- stream layout: where each block record starts (one warp; blocks are few) ---------------------------------
// record = 5 bits (lw-3) | lw bits `written` | `written` bits (COS:1024-1035); end marker 5+3 zero bits (COS:491-492)
__global__ void __launch_bounds__(32) kzg_stream_layout_kernel(KzgBlock* __restrict__ blocks, int nBlocks, KzgSeg* __restrict__ segs,
                                                                int segsPerBlock, const u8* __restrict__ hdrBytes, u32* __restrict__ dst32,
                                                                i64 headerBits, i64* __restrict__ totalBits, i64 capBits) {
  const int lane = threadIdx.x;
  u64 pos = (u64)headerBits;
  int failed = 0;
  for (int base = 0; base < nBlocks; base += 32) {
    const int b = base + lane;
    u64 rec = 0; int lw = 0; i64 written = 0;
    if (b < nBlocks) {
      if (blocks[b].status != 0) failed = 1;
      written = blocks[b].written;
      lw = (written < 8) ? 3 : ilog2((u32)(written >> 3)) + 4;
      rec = 5 + (u64)lw + (u64)written;
    }
    u64 incl = rec;
    for (int o = 1; o < 32; o <<= 1) { const u64 t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
    const u64 start = pos + incl - rec;
    if (b < nBlocks && !failed && (i64)(start + rec + 8) <= capBits) {
      put_bits_atomic(dst32, start, (u64)(lw - 3), 5);
      put_bits_atomic(dst32, start + 5, (u64)written, lw);
      const u64 p0 = start + 5 + lw;
      const u8* h = hdrBytes + b * KZG_HDR_STRIDE;
      const int hb = blocks[b].hdrBytes;
      for (int i = 0; i < hb; i++) put_bits_atomic(dst32, p0 + 8 * i, h[i], 8);
      blocks[b].srcBit = (i64)(p0 + 8ull * hb);      // where this block's entropy payload goes
    }
    pos += __shfl_sync(0xFFFFFFFFu, incl, 31);
  }
  failed = __any_sync(0xFFFFFFFFu, failed);
  if (lane == 0) totalBits[0] = failed ? -1 : (i64)(pos + 8);
}

// rebase a block's chunk segments to absolute stream positions
__global__ void kzg_seg_rebase_kernel(const KzgBlock* __restrict__ blocks, KzgSeg* __restrict__ segs, int segsPerBlock, i64 nSegs, i64 capBits) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nSegs) return;
  const int b = (int)(i / segsPerBlock);
  if (blocks[b].status != 0) { segs[i].nBits = 0; return; }
  const u64 base = (u64)blocks[b].srcBit;
  segs[i].dstBit += base;
  if ((i64)(segs[i].dstBit + segs[i].nBits) > capBits) segs[i].nBits = 0;    // caller reports the overflow from totalBits
}

// ---- the copy itself: one CTA per segment ----------------------------------------------------------------------
__global__ void __launch_bounds__(256) kzg_bitcopy_kernel(const KzgSeg* __restrict__ segs, u32* __restrict__ dst32) {
  const KzgSeg sg = segs[blockIdx.x];
  if (sg.nBits == 0) return;
  const uintptr_t sa = (uintptr_t)sg.src;
  const u32* __restrict__ src32 = (const u32*)(sa & ~(uintptr_t)3);
  const u64 srcBit = sg.srcBit + 8 * (u64)(sa & 3);
  const u64 d0 = sg.dstBit, d1 = sg.dstBit + sg.nBits;
  const u64 wFirst = d0 >> 5, wLast = (d1 - 1) >> 5;
  for (u64 w = wFirst + threadIdx.x; w <= wLast; w += blockDim.x) {
    const u64 lo = max(w << 5, d0), hi = min((w << 5) + 32, d1);
    const int nb = (int)(hi - lo);
    const u64 sbit = srcBit + (lo - d0);
    const u64 k = sbit >> 5;
    const int sh = (int)(sbit & 31);
    const u32 a = __byte_perm(src32[k], 0, 0x0123);
    const u32 bw = (sh + nb > 32) ? __byte_perm(src32[k + 1], 0, 0x0123) : 0u;
    u32 v = __funnelshift_l(bw, a, sh);          // nb valid bits at the top
    if (nb == 32) { dst32[w] = __byte_perm(v, 0, 0x0123); continue; }
    v = (v >> (32 - nb)) << (int)(((w << 5) + 32) - hi);
    atomicOr(&dst32[w], __byte_perm(v, 0, 0x0123));
  }
}

int kzg_bitcopy_launch(cudaStream_t s, const KzgSeg* segs, i64 nSegs) {
  if (nSegs <= 0) return 0;
  kzg_bitcopy_kernel<<<(unsigned)nSegs, 256, 0, s>>>(segs, (u32*)nullptr);
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(1);
  return 0;
}

// ---- host launchers ------------------------------------------------------------------------------------------
int kzg_commit_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, const int* result, const u8* enabled, int stage, int forward) {
  kzg_commit_kernel<<<(nBlocks + 127) / 128, 128, 0, s>>>(d_blocks, nBlocks, result, enabled, stage, forward);
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(1);
  return 0;
}

int kzg_assemble_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, KzgSeg* segs, int segsPerBlock, u8* hdrBytes, int nbFunctions,
                        int container, u8* d_out, i64 headerBits, i64* d_totalBits, i64 capBytes) {
  kzg_block_layout_kernel<<<nBlocks, 32, 0, s>>>(d_blocks, segs, segsPerBlock, hdrBytes, nbFunctions, container);
  CUDA_TRY(cudaGetLastError());
  const i64 nSegs = (i64)nBlocks * segsPerBlock;
  if (container) {
    kzg_stream_layout_kernel<<<1, 32, 0, s>>>(d_blocks, nBlocks, segs, segsPerBlock, hdrBytes, (u32*)d_out, headerBits, d_totalBits, capBytes * 8);
    CUDA_TRY(cudaGetLastError());
    kzg_count_launch(1);
  }
  kzg_seg_rebase_kernel<<<(unsigned)((nSegs + 255) / 256), 256, 0, s>>>(d_blocks, segs, segsPerBlock, nSegs, capBytes * 8);
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(1);
  kzg_bitcopy_kernel<<<(unsigned)nSegs, 256, 0, s>>>(segs, (u32*)d_out);
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(2);
  return 0;
}
