// rolz.cu — ROLZ (`-t ROLZ`, ROLZCodec1) forward / inverse (sm_100a).
//
// Replaces K/transform/ROLZCodec.java's ROLZCodec1: reduced-offset LZ with
// 64 Ki contexts x 16 ring slots, whose literals / tokens / lengths / match indexes are entropy-coded
// *inside the transform* with ANS (ROLZCodec.java:611-625).  Pipeline per batch:
//   forward: parse kernel (one warp per block: the 16 ring slots of a context are probed by 16 lanes at
//            once, hash tag check first, then 8-byte compares) -> the four side buffers become four
//            "virtual blocks" fed to the same rANS kernels as the entropy stage (ans.cu) -> their chunk
//            bit strings are stitched into the block by the bit-granular copy of container.cu.
//   inverse: header walk (one thread per block, chained over the four ANS streams) -> rANS decode kernels
//            -> replay kernel (one warp per block).
// The ring table (4 MiB per block) and the counters live in global memory and stay L2-resident.
// Blocks of more than one ROLZ chunk (> 16 MiB) are not handled yet (status -KZG_ERR_BLOCK_SIZE).
#include "kzg_common.cuh"
#include "kzg_transforms.cuh"
#include "kzg_entropy.cuh"
#include "kzg_container.cuh"
#include "kzg_xf_kernels.cuh"
#include "ans_scan.cuh"
#include <algorithm>
#include <vector>

#define RZ_HASH_SIZE 65536
#define RZ_CHUNK (16 * 1024 * 1024)
#define RZ_HASH 200002979u
#define RZ_HASH_MASK (~(u32)(RZ_CHUNK - 1))
#define RZ_MAX_MATCH (3 + 65535)
#define RZ_LOGPOS 4
// pseudo entropy ids of the virtual blocks
#define RZ_E_LIT0 100      // ANSRangeEncoder(obs, 0): chunk 16384
#define RZ_E_LIT1 101      // ANSRangeEncoder(obs, 1): chunk 4 MiB
#define RZ_E_M 102         // ANSRangeEncoder(obs, 0, 32768)

struct RzLayout {          // per-block scratch layout (byte offsets inside the block's scratch area)
  size_t lit, tk, len, midx;            // side buffers
  int litCap, tkCap, lenCap, midxCap;
  size_t hdr[3], pay[3];                // ANS scratch for ids 100/101/102 (hdr/pay areas shared by the 4 virtual blocks by slot)
  size_t total;
};

struct RzParams {
  KzgBlock* vb;            // 4 virtual blocks per block
  u8* scratch; i64 scratchStride;
  i32* matches;            // [nBlocks][65536 << 4]
  i32* counters;           // [nBlocks][65536]
  RzLayout L;
  KzgChunkInfo* chunks; int maxChunks;    // decode: [4 * nBlocks][maxChunks]
};

__device__ __forceinline__ u64 rz_ld64(const u8* p) {
  const uintptr_t a = (uintptr_t)p;
  const u64* q = (const u64*)(a & ~(uintptr_t)7);
  const int sh = (int)(a & 7) * 8;
  const u64 w0 = q[0];
  if (sh == 0) return w0;
  return (w0 >> sh) | (q[1] << (64 - sh));
}
__device__ __forceinline__ u32 rz_ld32(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24); }
__device__ __forceinline__ int rz_key1(const u8* buf, int idx) { return (int)buf[idx] | ((int)buf[idx + 1] << 8); }                  // :123-125
__device__ __forceinline__ int rz_key2(const u8* buf, int idx) { return (int)((i64)(rz_ld64(buf + idx) * (u64)RZ_HASH) >> 40) & 0xFFFF; }   // :135-137
__device__ __forceinline__ u32 rz_hash(const u8* buf, int idx) { return ((rz_ld32(buf + idx) << 8) * RZ_HASH) & RZ_HASH_MASK; }       // :147-149

// Global.detectSimpleType (K/Global.java:556-608) on a 256-bin histogram
__device__ int rz_detect_type(int count, const u32* f) {
  if (count == 0) return KZG_DT_UNDEFINED;
  int sum = f['a'] + f['c'] + f['g'] + f['n'] + f['t'] + f['u'] + f['A'] + f['C'] + f['G'] + f['N'] + f['T'] + f['U'];
  if (sum > count - count / 12) return KZG_DT_DNA;
  const char* num = "0123456789+-*/=,.:; ";
  sum = 0;
  for (int i = 0; i < 20; i++) sum += f[(u8)num[i]];
  if (sum == count) return KZG_DT_NUMERIC;
  sum = (f[0x3D] == 1) ? 1 : 0;
  for (int c = 'A'; c <= 'Z'; c++) sum += f[c];
  for (int c = 'a'; c <= 'z'; c++) sum += f[c];
  for (int c = '0'; c <= '9'; c++) sum += f[c];
  sum += f['+'] + f['/'];
  if (sum == count) return KZG_DT_BASE64;
  sum = 0;
  for (int i = 0; i < 256; i++) sum += (f[i] > 0) ? 1 : 0;
  if (sum == 256) return KZG_DT_BIN;
  if (sum <= 4) return KZG_DT_SMALL_ALPHABET;
  return KZG_DT_UNDEFINED;
}

// ROLZCodec1.findMatch (:365-406): lanes 0..15 probe the ring slots counter, counter-1, ...; the strictly longest
// match wins, ties go to the slot probed first (lowest lane).  Returns -1 or (bestIdx << 16) | (bestLen - minMatch).
__device__ __forceinline__ int rz_find_match(const u8* __restrict__ buf, const i32* matches, int sbaLength, int sbaIndex,
                                             int pos, u32 hash32, int counter, int base, int minMatch, int lane) {
  const int maxMatch = min(RZ_MAX_MATCH, sbaLength - pos) - 8;
  int n = 0;
  if (lane < 16) {
    const u32 e = (u32)matches[base + ((counter - lane) & 15)];
    if ((e & RZ_HASH_MASK) == hash32) {
      const int ref = (int)(e & ~RZ_HASH_MASK) + sbaIndex;
      while (n < maxMatch) {
        const u64 diff = rz_ld64(buf + ref + n) ^ rz_ld64(buf + pos + n);
        if (diff != 0) { n += (__ffsll((long long)diff) - 1) >> 3; break; }
        n += 8;
      }
    }
  }
  // argmax with lowest-lane tie-break
  int best = n, bestLane = lane;
  for (int o = 8; o > 0; o >>= 1) {
    const int on = __shfl_xor_sync(0xFFFFFFFFu, best, o), ol = __shfl_xor_sync(0xFFFFFFFFu, bestLane, o);
    if (on > best || (on == best && ol < bestLane)) { best = on; bestLane = ol; }
  }
  best = __shfl_sync(0xFFFFFFFFu, best, 0); bestLane = __shfl_sync(0xFFFFFFFFu, bestLane, 0);
  // Java only records a slot when n > bestLen (strict, starting from 0): a best of 0 means "none"
  if (best < minMatch || best == 0) return -1;
  return (bestLane << 16) | (best - minMatch);
}

__device__ __forceinline__ int rz_emit_length(u8* buf, int idx, int cap, int length, int lane, bool& overflow) {   // :670-683
  int n = 1 + (length >= (1 << 7)) + (length >= (1 << 14)) + (length >= (1 << 21));
  if (idx + n > cap) { overflow = true; return idx; }
  if (lane == 0) {
    int k = idx;
    if (length >= 1 << 7) {
      if (length >= 1 << 14) {
        if (length >= 1 << 21) buf[k++] = (u8)(0x80 | (length >> 21));
        buf[k++] = (u8)(0x80 | (length >> 14));
      }
      buf[k++] = (u8)(0x80 | (length >> 7));
    }
    buf[k++] = (u8)(length & 0x7F);
  }
  return idx + n;
}

// ================================================================================================================
// forward parse: ROLZCodec1.forward (:419-640) up to the entropy coding of the side buffers
// ================================================================================================================
__global__ void __launch_bounds__(32) rolz_parse_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P, RzParams R) {
  __shared__ u32 hist[256];
  const int lane = threadIdx.x, b = blockIdx.x;
  KzgBlock& B = blocks[b];
  int* res = P.result + 2 * b;
  KzgBlock* vb = R.vb + 4 * b;
  if (lane < 4) { memset(&vb[lane], 0, sizeof(KzgBlock)); vb[lane].status = 1; }     // status 1 = unused slot
  if (lane == 0) { res[0] = 0; res[1] = 0; }
  __syncwarp();
  if (B.status != 0 || !P.enabled[b]) return;
  const int count = B.curLen;
  if (count < 64 || count > (1 << 30)) return;                 // MIN_BLOCK_SIZE / MAX_BLOCK_SIZE (:207-212)
  if (((count <= 512) ? count + 64 : count) > min(B.cap, kzg_dst_limit(B, P.dstLimit[b]))) return;    // output.length - output.index < getMaxEncodedLength(count)
  if (count - 4 > RZ_CHUNK) { if (lane == 0) atomicExch(&B.status, -KZG_ERR_BLOCK_SIZE); return; }
  const u8* __restrict__ src = B.cur;
  u8* __restrict__ dst = B.alt;
  u8* sc = R.scratch + (i64)b * R.scratchStride;
  u8* litBuf = sc + R.L.lit; u8* tkBuf = sc + R.L.tk; u8* lenBuf = sc + R.L.len; u8* mIdxBuf = sc + R.L.midx;
  const int sizeChunk0 = min(count, RZ_CHUNK);
  const int litCap = (sizeChunk0 <= 512) ? sizeChunk0 + 64 : sizeChunk0, lenCap = sizeChunk0 / 5, mIdxCap = sizeChunk0 / 4, tkCap = sizeChunk0 / 4;
  i32* matches = R.matches + (i64)b * (RZ_HASH_SIZE << RZ_LOGPOS);
  i32* counters = R.counters + (i64)b * RZ_HASH_SIZE;
  const int srcEnd = count - 4;
  const int litOrder = (count < (1 << 17)) ? 0 : 1;
  int flags = litOrder;
  int minMatch = 3, delta = 2;
  // dataType sniffing (:451-485)
  int dtp = B.dataType;
  if (dtp == KZG_DT_UNDEFINED) {
    for (int i = lane; i < 256; i += 32) hist[i] = 0;
    __syncwarp();
    for (int i = lane; i < count; i += 32) atomicAdd(&hist[src[i]], 1u);
    __syncwarp();
    dtp = rz_detect_type(count, hist);
    if (dtp != KZG_DT_UNDEFINED && lane == 0) B.dataType = dtp;
  }
  if (dtp == KZG_DT_EXE) { delta = 3; flags |= 8; }
  else if (dtp == KZG_DT_MULTIMEDIA) { delta = 8; minMatch = 4; flags |= 2; }
  else if (dtp == KZG_DT_DNA) { delta = 8; minMatch = 7; flags |= 4; }
  const int mm = minMatch, dt = delta;
  flags |= (RZ_LOGPOS << 4);
  for (int i = lane; i < RZ_HASH_SIZE; i += 32) counters[i] = 0;
  for (int i = lane; i < (RZ_HASH_SIZE << RZ_LOGPOS); i += 32) matches[i] = 0;
  __syncwarp();
  __threadfence_block();

  // single chunk: [startChunk = 0, endChunk = srcEnd)
  const int endChunk = min(sizeChunk0, srcEnd);
  const int sizeChunk = endChunk;
  int litIdx = 0, tkIdx = 0, lenIdx = 0, mIdxIdx = 0;
  int srcIdx = 0;
  const int nFirst = min(srcEnd, 8);
  if (lane < nFirst) litBuf[lane] = src[lane];
  litIdx = nFirst; srcIdx = nFirst;
  int firstLitIdx = srcIdx, srcInc = 0;
  bool overflow = false;
  while (srcIdx < endChunk) {
    int key = (mm == 3) ? rz_key1(src, srcIdx - dt) : rz_key2(src, srcIdx - dt);
    int base = key << RZ_LOGPOS;
    u32 hash32 = rz_hash(src, srcIdx);
    int counter = counters[key];
    int match = rz_find_match(src, matches, endChunk, 0, srcIdx, hash32, counter, base, mm, lane);
    __syncwarp();
    if (lane == 0) { counters[key] = (counter + 1) & 15; matches[base + ((counter + 1) & 15)] = (i32)(hash32 | (u32)srcIdx); }
    __syncwarp();
    if (match == -1) { srcIdx++; srcIdx += (srcInc >> 6); srcInc++; continue; }
    {
      key = (mm == 3) ? rz_key1(src, srcIdx + 1 - dt) : rz_key2(src, srcIdx + 1 - dt);
      base = key << RZ_LOGPOS;
      hash32 = rz_hash(src, srcIdx + 1);
      counter = counters[key];
      const int match2 = rz_find_match(src, matches, endChunk, 0, srcIdx + 1, hash32, counter, base, mm, lane);
      if ((match2 >= 0) && ((match2 & 0xFFFF) > (match & 0xFFFF))) {
        match = match2;
        srcIdx++;
        __syncwarp();
        if (lane == 0) { counters[key] = (counter + 1) & 15; matches[base + ((counter + 1) & 15)] = (i32)(hash32 | (u32)srcIdx); }
        __syncwarp();
      }
    }
    const int litLen = srcIdx - firstLitIdx;
    const int token = (litLen < 31) ? (litLen << 3) : 0xF8;
    const int mLen = match & 0xFFFF;
    if (tkIdx >= tkCap) { overflow = true; break; }
    if (mLen >= 7) { if (lane == 0) tkBuf[tkIdx] = (u8)(token | 0x07); tkIdx++; lenIdx = rz_emit_length(lenBuf, lenIdx, lenCap, mLen - 7, lane, overflow); }
    else { if (lane == 0) tkBuf[tkIdx] = (u8)(token | mLen); tkIdx++; }
    if (litLen >= 31) lenIdx = rz_emit_length(lenBuf, lenIdx, lenCap, litLen - 31, lane, overflow);
    if (overflow || litIdx + litLen > litCap || mIdxIdx >= mIdxCap) { overflow = true; break; }
    for (int i = lane; i < litLen; i += 32) litBuf[litIdx + i] = src[firstLitIdx + i];
    litIdx += litLen;
    if (lane == 0) mIdxBuf[mIdxIdx] = (u8)((u32)match >> 16);
    mIdxIdx++;
    srcIdx += (mLen + mm);
    firstLitIdx = srcIdx;
    srcInc = 0;
  }
  if (!overflow) {
    // last chunk literals (:588-606)
    const int litLen = sizeChunk - firstLitIdx;
    if (tkIdx != 0) {
      if (tkIdx >= tkCap) overflow = true;
      else { if (lane == 0) tkBuf[tkIdx] = (u8)((litLen >= 31) ? 0xF8 : (litLen << 3)); tkIdx++; }
    }
    if (!overflow && litLen >= 31) lenIdx = rz_emit_length(lenBuf, lenIdx, lenCap, litLen - 31, lane, overflow);
    if (!overflow && (litLen < 0 || litIdx + litLen > litCap)) overflow = true;
    if (!overflow) { for (int i = lane; i < litLen; i += 32) litBuf[litIdx + i] = src[firstLitIdx + i]; litIdx += litLen; }
  }
  if (overflow) { if (lane == 0) atomicExch(&B.status, -KZG_ERR_PROCESS_BLOCK); return; }     // Java: ArrayIndexOutOfBounds
  __syncwarp();
  if (lane == 0) {
    // block header (:433, 489-490) and the chunk's four sizes; the ANS streams follow at byte 5 + 16
    dst[0] = (u8)(count >> 24); dst[1] = (u8)(count >> 16); dst[2] = (u8)(count >> 8); dst[3] = (u8)count;
    dst[4] = (u8)flags;
    const int sz[4] = {litIdx, tkIdx, lenIdx, mIdxIdx};
    for (int k = 0; k < 4; k++) { dst[5 + 4 * k] = (u8)(sz[k] >> 24); dst[6 + 4 * k] = (u8)(sz[k] >> 16); dst[7 + 4 * k] = (u8)(sz[k] >> 8); dst[8 + 4 * k] = (u8)sz[k]; }
    u8* bufs[4] = {litBuf, tkBuf, lenBuf, mIdxBuf};
    for (int k = 0; k < 4; k++) {
      KzgBlock& V = vb[k];
      V.cur = bufs[k]; V.curLen = sz[k]; V.status = 0;
      V.entropy = (k == 0) ? (litOrder ? RZ_E_LIT1 : RZ_E_LIT0) : RZ_E_M;
    }
    res[1] = flags;        // stash; the layout kernel finishes the result
    res[0] = 2;            // "parsed"
  }
}

// after the ANS kernels: place the four streams' segments, zero the destination, append the tail bytes
__global__ void __launch_bounds__(256) rolz_layout_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P, RzParams R, KzgSeg* __restrict__ segs, int segsPerVb) {
  __shared__ u64 totalBits;
  const int b = blockIdx.x;
  KzgBlock& B = blocks[b];
  int* res = P.result + 2 * b;
  KzgSeg* S = segs + (i64)(4 * b) * segsPerVb;
  if (res[0] != 2) { for (int i = threadIdx.x; i < 4 * segsPerVb; i += blockDim.x) S[i].nBits = 0; return; }
  KzgBlock* vb = R.vb + 4 * b;
  u8* __restrict__ dst = B.alt;
  if (threadIdx.x == 0) {
    u64 pos = (u64)(5 + 16) * 8;
    bool bad = false;
    for (int k = 0; k < 4; k++) {
      if (vb[k].status != 0) bad = true;
      for (int i = 0; i < segsPerVb; i++) { KzgSeg& s = S[k * segsPerVb + i]; if (i == 0) s.nBits = 0; s.dstBit = pos; pos += s.nBits; }
    }
    totalBits = bad ? ~0ull : pos;
  }
  __syncthreads();
  const u64 tb = totalBits;
  const int count = B.curLen;
  const i64 bytes = (tb == ~0ull) ? -1 : (i64)((tb + 7) >> 3);
  // dstIdx + buf.length > dst.length (:629-633) / dstIdx + 4 > dst.length (:642-646) -> false
  const bool fits = (bytes >= 0) && (bytes + 4 <= (i64)min(B.cap, kzg_dst_limit(B, P.dstLimit[b])));      // dst slice length as the Java call sees it
  if (!fits) {
    for (int i = threadIdx.x; i < 4 * segsPerVb; i += blockDim.x) S[i].nBits = 0;
    if (threadIdx.x == 0) { res[0] = 0; res[1] = 0; if (tb == ~0ull) atomicExch(&B.status, -KZG_ERR_PROCESS_BLOCK); }
    return;
  }
  // zero the stream area (the bit copy ORs into it), keep the 21 header bytes
  for (i64 i = 21 + threadIdx.x; i < bytes + 4; i += blockDim.x) dst[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    const u8* src = B.cur;
    const int srcEnd = count - 4;
    for (int k = 0; k < 4; k++) dst[bytes + k] = src[srcEnd + k];       // last literals (:652-655)
    res[0] = 1; res[1] = (int)bytes + 4;
  }
  // segments are addressed relative to this block's dst: publish the base through srcBit of a pseudo block
}

// rebase virtual-block segments onto each block's destination buffer (dst address as a bit offset from address 0)
__global__ void rolz_seg_base_kernel(const KzgBlock* __restrict__ blocks, KzgSeg* __restrict__ segs, int segsPerVb, i64 nSegs) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nSegs) return;
  const int b = (int)(i / (4 * segsPerVb));
  segs[i].dstBit += 8ull * (u64)(uintptr_t)blocks[b].alt;
}

// ================================================================================================================
// inverse
// ================================================================================================================
struct RzDecInfo { i32 ok, szBlock, flags, litLen, tkLen, mLenLen, mIdxLen, endByte; };

// header walk: ROLZCodec1.inverse :696-838 up to (not including) the ANS decoding itself
__global__ void rolz_scan_kernel(KzgBlock* __restrict__ blocks, int nBlocks, KzgXfParams P, RzParams R, RzDecInfo* __restrict__ info) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nBlocks) return;
  KzgBlock& B = blocks[b];
  KzgBlock* vb = R.vb + 4 * b;
  for (int k = 0; k < 4; k++) { memset(&vb[k], 0, sizeof(KzgBlock)); vb[k].status = 1; }
  RzDecInfo& I = info[b];
  I.ok = 0;
  P.result[2 * b] = 0; P.result[2 * b + 1] = 0;
  if (B.status != 0 || !P.enabled[b]) return;
  const int count = B.curLen;
  const u8* __restrict__ src = B.cur;
  if (count < 5 + 16 + 4) return;
  const int szBlock = (int)(((u32)src[0] << 24) | ((u32)src[1] << 16) | ((u32)src[2] << 8) | (u32)src[3]) - 4;
  const int outLimit = min(kzg_dst_limit(B, P.dstLimit[b]), B.cap);
  if (szBlock <= 0 || szBlock > outLimit - 4) return;
  if (szBlock > RZ_CHUNK) { B.status = -KZG_ERR_BLOCK_SIZE; return; }
  const int flags = src[4];
  if ((flags >> 4) != RZ_LOGPOS) { if ((flags >> 4) < 2 || (flags >> 4) > 8) return; B.status = -KZG_ERR_INVALID_CODEC; return; }   // only logPosChecks 4 is emitted
  const int sizeChunk = szBlock;
  auto be = [&](int o) { return (i32)(((u32)src[o] << 24) | ((u32)src[o + 1] << 16) | ((u32)src[o + 2] << 8) | (u32)src[o + 3]); };
  const int litLen = be(5), tkLen = be(9), mLenLen = be(13), mIdxLen = be(17);
  const int firstLitLen = min(sizeChunk, 8);
  if ((litLen < 0) || (tkLen < 0) || (mLenLen < 0) || (mIdxLen < 0)) return;
  if ((litLen > sizeChunk) || (tkLen > sizeChunk / 4) || (mLenLen > (sizeChunk / 5) + 4 - 4) || (mIdxLen > sizeChunk / 4)) return;   // :802-808
  if ((litLen < firstLitLen) || ((tkLen == 0) && (mIdxLen != 0)) || ((tkLen > 0) && (mIdxLen + 1 != tkLen))) return;
  u8* sc = R.scratch + (i64)b * R.scratchStride;
  u8* bufs[4] = {sc + R.L.lit, sc + R.L.tk, sc + R.L.len, sc + R.L.midx};
  const int lens[4] = {litLen, tkLen, mLenLen, mIdxLen};
  const int litOrder = flags & 1;
  // positions are scanned relative to `src`; the decode kernels address the stream from address 0 (P.stream == nullptr),
  // so the block's base address (in bits) is added afterwards
  const u64 absBase = 8ull * (u64)(uintptr_t)src;
  BitReaderD br(src, 8ull * 21, 8ull * (u64)count);
  for (int k = 0; k < 4; k++) {
    KzgBlock& V = vb[k];
    const u64 pos0 = br.pos;
    const int chunkSize = (k == 0) ? (litOrder ? (4 << 20) : 16384) : 32768;
    KzgChunkInfo* ci = R.chunks + (i64)(4 * b + k) * R.maxChunks;
    const int r = ans_scan_stream(br, lens[k], chunkSize, (k == 0) ? litOrder : 0, ci);
    if (r < 0) { B.status = r; return; }
    if (lens[k] > 32) {
      const int nChunks = (lens[k] + chunkSize - 1) / chunkSize;
      for (int c = 0; c < nChunks; c++) { ci[c].hdrBit += (i64)absBase; ci[c].payBit += (i64)absBase; }
    }
    V.cur = bufs[k]; V.curLen = lens[k]; V.preLen = lens[k];
    V.entropy = (k == 0) ? (litOrder ? RZ_E_LIT1 : RZ_E_LIT0) : RZ_E_M;
    V.srcBit = (i64)(absBase + pos0); V.srcBits = (i64)(br.end - pos0);
    V.entBits = (i64)(br.pos - pos0);
  }
  for (int k = 0; k < 4; k++) vb[k].status = 0;      // published only once every stream scanned cleanly
  const u64 used = br.pos;
  I.endByte = (i32)((used + 7) >> 3);
  I.szBlock = szBlock; I.flags = flags; I.litLen = litLen; I.tkLen = tkLen; I.mLenLen = mLenLen; I.mIdxLen = mIdxLen;
  I.ok = 1;
}

__device__ __forceinline__ int rz_read_length(const u8* buf, int& idx) {     // :969-989
  int next = (int8_t)buf[idx++];
  int length = next & 0x7F;
  if (next & 0x80) {
    next = (int8_t)buf[idx++]; length = (length << 7) | (next & 0x7F);
    if (next & 0x80) {
      next = (int8_t)buf[idx++]; length = (length << 7) | (next & 0x7F);
      if (next & 0x80) { next = (int8_t)buf[idx++]; length = (length << 7) | (next & 0x7F); }
    }
  }
  return length;
}

// replay: ROLZCodec1.inverse :839-958
__global__ void __launch_bounds__(32) rolz_replay_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P, RzParams R, const RzDecInfo* __restrict__ info) {
  const int lane = threadIdx.x, b = blockIdx.x;
  KzgBlock& B = blocks[b];
  const RzDecInfo I = info[b];
  int* res = P.result + 2 * b;
  if (B.status != 0 || !P.enabled[b] || !I.ok) return;
  const KzgBlock* vb = R.vb + 4 * b;
  for (int k = 0; k < 4; k++) if (vb[k].status != 0) { if (lane == 0) atomicExch(&B.status, -KZG_ERR_PROCESS_BLOCK); return; }
  const u8* __restrict__ src = B.cur;
  u8* __restrict__ dst = B.alt;
  const int count = B.curLen;
  u8* sc = R.scratch + (i64)b * R.scratchStride;
  const u8* litBuf = sc + R.L.lit; const u8* tkBuf = sc + R.L.tk; const u8* lenBuf = sc + R.L.len; const u8* mIdxBuf = sc + R.L.midx;
  i32* matches = R.matches + (i64)b * (RZ_HASH_SIZE << RZ_LOGPOS);
  i32* counters = R.counters + (i64)b * RZ_HASH_SIZE;
  int minMatch = 3, delta = 2;
  switch (I.flags & 0x0E) { case 2: minMatch = 4; delta = 8; break; case 4: minMatch = 7; delta = 8; break; case 8: delta = 3; break; default: break; }
  const int mm = minMatch, dt = delta;
  const int dstEnd = I.szBlock, endChunk = I.szBlock, sizeChunk = I.szBlock;
  int srcIdx = I.endByte;                  // bytes consumed so far, counted from byte 0 of the block (21 header bytes + the four streams)
  int dstIdx = 0;
  if (I.tkLen == 0) {                      // only literals (:840-853)
    if (I.litLen != sizeChunk) return;
    for (int i = lane; i < sizeChunk; i += 32) dst[i] = litBuf[i];
    dstIdx = sizeChunk;
  } else {
    for (int i = lane; i < RZ_HASH_SIZE; i += 32) counters[i] = 0;
    for (int i = lane; i < (RZ_HASH_SIZE << RZ_LOGPOS); i += 32) matches[i] = 0;
    __syncwarp();
    __threadfence_block();
    int litIdx = 0, tkIdx = 0, lenIdx = 0, mIdxIdx = 0;
    const int n = min(dstEnd - dstIdx, 8);
    if (lane < n) dst[lane] = litBuf[lane];
    dstIdx = n; litIdx = n;
    __syncwarp();
    bool fail = false;
#ifdef KZG_RZ_TIMING
    long long tLit = 0, tMatch = 0, nLitTok = 0, nLit = 0, c0, c1;
#endif
    while (dstIdx < endChunk) {
#ifdef KZG_RZ_TIMING
      c0 = clock64();
#endif
      if (tkIdx >= I.tkLen) { fail = true; break; }
      const int token = tkBuf[tkIdx++];
      int matchLen = token & 0x07;
      if (matchLen == 7) { if (lenIdx >= I.mLenLen) { fail = true; break; } matchLen = rz_read_length(lenBuf, lenIdx) + 7; }
      int litLen;
      if (token < 0xF8) litLen = token >> 3;
      else { if (lenIdx >= I.mLenLen) { fail = true; break; } litLen = rz_read_length(lenBuf, lenIdx) + 31; }
      if (litLen > 0) {
        if (litIdx + litLen > I.litLen || dstIdx + litLen > dstEnd + 4) { fail = true; break; }
        for (int i = lane; i < litLen; i += 32) dst[dstIdx + i] = litBuf[litIdx + i];
        __syncwarp();
        // register the literal positions with the encoder's skip pattern (:889-900): the t-th registered literal is
        // j_t = t + sum_{u<t} (u >> 6).  Positions with different keys touch different ring rows, so 32 of them go at once:
        // lanes with the same key (match_any) take consecutive ring slots in lane order, the last 16 of a group survive.
        // number of registered literals: the largest n with j_{n-1} < litLen (j_t = t for t < 64, the common case)
        int nReg = litLen;
        if (litLen > 64) { nReg = 64; while (true) { const int k6 = nReg >> 6; if (nReg + 32 * k6 * (k6 - 1) + k6 * (nReg & 63) >= litLen) break; nReg++; } }
        for (int t0 = 0; t0 < nReg; t0 += 32) {
          const int t = t0 + lane, k6 = t >> 6;
          const int j = (k6 == 0) ? t : (t + 32 * k6 * (k6 - 1) + k6 * (t & 63));
          const bool on = t < nReg;
          const int key = on ? ((mm == 3) ? rz_key1(dst, dstIdx + j - dt) : rz_key2(dst, dstIdx + j - dt)) : (0x10000 + lane);
          const u32 peers = __match_any_sync(0xFFFFFFFFu, key);
          if (on) {
            const int rank = __popc(peers & ((1u << lane) - 1)), size = __popc(peers);
            const int c0 = counters[key];
            if (rank >= size - 16) matches[(key << RZ_LOGPOS) + ((c0 + rank + 1) & 15)] = dstIdx + j;
            if (rank == size - 1) counters[key] = (c0 + size) & 15;
          }
          __syncwarp();
        }
        __syncwarp();
        litIdx += litLen;
        dstIdx += litLen;
        if (dstIdx >= endChunk) { if (dstIdx == endChunk) break; fail = true; break; }
      }
#ifdef KZG_RZ_TIMING
      c1 = clock64(); tLit += c1 - c0; if (litLen > 0) { nLitTok++; nLit += litLen; }
#endif
      if (dstIdx + matchLen + mm > dstEnd) { fail = true; break; }
      const int key = (mm == 3) ? rz_key1(dst, dstIdx - dt) : rz_key2(dst, dstIdx - dt);
      const int base = key << RZ_LOGPOS;
      if (mIdxIdx >= I.mIdxLen) { fail = true; break; }
      const int matchIdx = mIdxBuf[mIdxIdx++];
      const int cnt = counters[key];
      const int ref = matches[base + ((cnt - matchIdx) & 15)];
      const int ml = matchLen + mm;
      if (ref < 0 || ref >= dstIdx) {        // ring slot never written (0) is position 0: valid only as an actual reference
        if (ref != 0 || dstIdx == 0) { fail = true; break; }
      }
      const int dist = dstIdx - ref;
      if (ml <= dist) { for (int i = lane; i < ml; i += 32) dst[dstIdx + i] = dst[ref + i]; }      // emitCopy (:162-179): forward byte copy
      else { for (int i = lane; i < ml; i += 32) dst[dstIdx + i] = dst[ref + (i % dist)]; }            // (overlapping: the pattern of `dist` bytes repeats)
      __syncwarp();
      if (lane == 0) { const int c = (cnt + 1) & 15; counters[key] = c; matches[base + c] = dstIdx; }
      __syncwarp();
      dstIdx += ml;
#ifdef KZG_RZ_TIMING
      tMatch += clock64() - c1;
#endif
    }
#ifdef KZG_RZ_TIMING
    if (lane == 0) printf("rolz replay block %d: %d tokens, %lld with literals (%lld literal bytes), literal part %lld cycles, match part %lld cycles\n", b, tkIdx, nLitTok, nLit, tLit, tMatch);
#endif
    if (fail) return;
    if ((tkIdx != I.tkLen) || (mIdxIdx != I.mIdxLen) || (litIdx != I.litLen) || (lenIdx != I.mLenLen)) return;
  }
  // a valid ROLZ block leaves exactly 4 raw tail bytes (:943-957)
  if ((dstIdx + 4 > min(kzg_dst_limit(B, P.dstLimit[b]), B.cap)) || (count - srcIdx != 4)) return;
  if (lane < 4) dst[dstIdx + lane] = src[srcIdx + lane];
  if (lane == 0) { res[0] = 1; res[1] = dstIdx + 4; }
}

// ================================================================================================================
// host
// ================================================================================================================
// chunk slots per ROLZ block for the three ANS flavours (lit order 0: 16 KiB chunks, lit order 1: 4 MiB, m-streams: 32 KiB)
struct RzCounts { int c100, c101, c102, maxChunks, segsPerVb; };
static RzCounts rz_counts(i32 maxLen) {
  RzCounts c;
  const size_t n = (size_t)maxLen + 64;
  c.c100 = (int)(n / 16384 + 2); c.c101 = (int)(n / (4 << 20) + 2); c.c102 = (int)((n / 4) / 32768 + 2);
  c.maxChunks = std::max(c.c100, std::max(c.c101, c.c102));
  c.segsPerVb = 1 + 2 * c.maxChunks;
  return c;
}
static const size_t RZ_HDR[3] = {512, 256 * 480 + 64, 512};
static const size_t RZ_PAY[3] = {2 * 16384 + 32, (4 << 20) + (4 << 17) + 64, 2 * 32768 + 32};

static RzLayout rz_layout(i32 maxLen) {      // per-block side buffers
  RzLayout L;
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t n = (size_t)maxLen + 64;
  size_t o = 0;
  L.litCap = (int)n; L.tkCap = (int)(n / 4 + 16); L.lenCap = (int)(n / 5 + 16); L.midxCap = (int)(n / 4 + 16);
  L.lit = o; o += al(n + 64);
  L.tk = o; o += al(n / 4 + 64);
  L.len = o; o += al(n / 5 + 64);
  L.midx = o; o += al(n / 4 + 64);
  for (int i = 0; i < 3; i++) { L.hdr[i] = 0; L.pay[i] = 0; }
  L.total = o;
  return L;
}

// per-block byte budget of everything ROLZ carves from the flat scratch pool (P.scratch, nBlocks * scratchStride bytes)
static size_t rz_pool_per_block(i32 maxLen) {
  const RzLayout L = rz_layout(maxLen);
  const RzCounts c = rz_counts(maxLen);
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t tab = std::max(kzg_ans1_enc_tab_u32(), kzg_ans1_dec_tab_u32()) * 4;
  return L.total + al(c.c100 * (RZ_HDR[0] + RZ_PAY[0])) + al(c.c101 * (RZ_HDR[1] + RZ_PAY[1] + tab)) + al(3 * c.c102 * (RZ_HDR[2] + RZ_PAY[2])) +
         al(4 * sizeof(KzgBlock)) + al(4 * (size_t)c.segsPerVb * sizeof(KzgSeg)) + al(4 * (size_t)c.maxChunks * sizeof(KzgChunkInfo)) +
         al(sizeof(RzDecInfo)) + al(3 * 4 * sizeof(int)) + 8192;
}

void kzg_rolz_scratch(i32 maxLen, bool forward, size_t* perBlockBytes, size_t* hashInts, size_t* aux32) {
  *perBlockBytes = std::max(*perBlockBytes, rz_pool_per_block(maxLen));
  *hashInts = std::max(*hashInts, (size_t)(RZ_HASH_SIZE << RZ_LOGPOS) + RZ_HASH_SIZE);
}

int kzg_rolz_launch(cudaStream_t s, bool forward, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, i32 maxLen) {
  const RzCounts C = rz_counts(maxLen);
  RzParams R;
  R.L = rz_layout(maxLen);
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t nb = (size_t)nBlocks;
  const size_t poolBytes = nb * (size_t)P.scratchStride;
  // flat carve: [side buffers: nb * L.total][ANS scratch pools per flavour][dense arrays]
  u8* pool = P.scratch;
  size_t o = 0;
  R.scratch = pool; R.scratchStride = (i64)R.L.total; o += nb * R.L.total;
  const size_t tabWords = std::max(kzg_ans1_enc_tab_u32(), kzg_ans1_dec_tab_u32());
  u8* hdrPool[3]; u8* payPool[3];
  const size_t slots[3] = {nb * C.c100, nb * C.c101, nb * 3 * C.c102};
  for (int i = 0; i < 3; i++) { hdrPool[i] = pool + o; o += al(slots[i] * RZ_HDR[i]); payPool[i] = pool + o; o += al(slots[i] * RZ_PAY[i]); }
  u32* tabPool = (u32*)(pool + o); o += al(slots[1] * tabWords * 4);
  R.vb = (KzgBlock*)(pool + o); o += al(nb * 4 * sizeof(KzgBlock));
  KzgSeg* segs = (KzgSeg*)(pool + o); o += al(nb * 4 * C.segsPerVb * sizeof(KzgSeg));
  R.chunks = (KzgChunkInfo*)(pool + o); o += al(nb * 4 * C.maxChunks * sizeof(KzgChunkInfo));
  RzDecInfo* dinfo = (RzDecInfo*)(pool + o); o += al(nb * sizeof(RzDecInfo));
  int* slotBase = (int*)(pool + o); o += al(3 * nb * 4 * sizeof(int));
  if (o > poolBytes) { kzg_set_error("rolz: scratch pool too small (%zu > %zu)", o, poolBytes); return -KZG_ERR_CREATE_CODEC; }
  R.matches = P.hashBuf;
  R.counters = P.hashBuf + nb * (size_t)(RZ_HASH_SIZE << RZ_LOGPOS);
  R.maxChunks = C.maxChunks;
  const int nVb = 4 * nBlocks;
  // first scratch slot of every virtual block, per flavour (vb 4b = literals, 4b+1..3 = tokens / lengths / indexes)
  std::vector<int> hs(3 * (size_t)nVb, 0);
  for (int b = 0; b < nBlocks; b++) {
    hs[0 * nVb + 4 * b] = b * C.c100;
    hs[1 * nVb + 4 * b] = b * C.c101;
    for (int k = 1; k < 4; k++) hs[2 * nVb + 4 * b + k] = (b * 3 + (k - 1)) * C.c102;
  }
  CUDA_TRY(cudaMemcpyAsync(slotBase, hs.data(), hs.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaStreamSynchronize(s));      // hs is stack-owned
  auto params = [&](int id) {
    KzgEntParams E;
    memset(&E, 0, sizeof(E));
    E.entropy = RZ_E_LIT0 + id;
    E.chunkSize = (id == 0) ? 16384 : (id == 1 ? (4 << 20) : 32768);
    E.maxChunks = C.maxChunks;
    E.hdrBuf = hdrPool[id]; E.hdrStride = (int)RZ_HDR[id]; E.payBuf = payPool[id]; E.payStride = (int)RZ_PAY[id];
    E.tabBuf = tabPool; E.tabStride = (i64)tabWords;
    E.slotBase = slotBase + (size_t)id * nVb;
    E.segs = segs; E.segsPerBlock = C.segsPerVb;
    E.stream = nullptr;                   // decode: srcBit is an absolute address * 8
    E.chunks = R.chunks;
    return E;
  };
  if (forward) {
    CUDA_TRY(cudaMemsetAsync(segs, 0, (size_t)nVb * C.segsPerVb * sizeof(KzgSeg), s));
    rolz_parse_kernel<<<nBlocks, 32, 0, s>>>(d_blocks, P, R);
    CUDA_TRY(cudaGetLastError());
    kzg_count_launch(1);
    for (int id = 0; id < 3; id++) {
      const KzgEntParams E = params(id);
      int r = kzg_ans_encode_launch(s, R.vb, nVb, E, id == 1 ? 1 : 0);
      if (r < 0) return r;
    }
    rolz_layout_kernel<<<nBlocks, 256, 0, s>>>(d_blocks, P, R, segs, C.segsPerVb);
    const i64 nSegs = (i64)nVb * C.segsPerVb;
    rolz_seg_base_kernel<<<(unsigned)((nSegs + 255) / 256), 256, 0, s>>>(d_blocks, segs, C.segsPerVb, nSegs);
    CUDA_TRY(cudaGetLastError());
    kzg_count_launch(2);
    return kzg_bitcopy_launch(s, segs, nSegs);
  }
  rolz_scan_kernel<<<(nBlocks + 31) / 32, 32, 0, s>>>(d_blocks, nBlocks, P, R, dinfo);
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(1);
  for (int id = 0; id < 3; id++) {
    const KzgEntParams E = params(id);
    int r = kzg_ans_decode_launch(s, R.vb, nVb, E, id == 1 ? 1 : 0, false);
    if (r < 0) return r;
  }
  rolz_replay_kernel<<<nBlocks, 32, 0, s>>>(d_blocks, P, R, dinfo);
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(1);
  return 0;
}
