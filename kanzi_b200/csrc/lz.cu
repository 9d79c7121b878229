// lz.cu — LZ (`-t LZ`, `-t LZX`) forward / inverse kernels (sm_100a).
//
// Replaces K/transform/LZCodec.java's LZXCodec.  The bitstream is defined by the
// reference's *greedy parse with a parse-dependent hash table*: which positions are inserted depends on
// the skip acceleration (srcInc >> 6), the lazy step and the repeat-offset state, so a bit-exact encoder
// replays that decision sequence.  One warp per block walks it: every lane evaluates the scalar control
// flow redundantly (no intra-warp communication on the critical path) and the lanes split the wide parts —
// 8-byte match comparisons (32 x 8 bytes per step), the ordered bulk insert of match positions into the
// hash table (__match_any picks the last writer per bucket) and the literal copies.
// -t LZ: 2^16-entry table of 24-bit positions in shared memory (192 KiB: one block per SM);
// -t LZX or blocks > 16 MiB: 32-bit table in global memory (L2-resident).
//
// The inverse lives in lz_inverse.cu (token parse + pointer jumping).
#include "kzg_common.cuh"
#include "kzg_transforms.cuh"

#define LZ_HASH_SEED 0x1E35A7BDull
#define LZ_MAX_DISTANCE1 ((1 << 16) - 2)
#define LZ_MAX_DISTANCE2 ((1 << 24) - 2)
#define LZ_MAX_MATCH (65535 + 254 + 4)
#define LZ_MIN_BLOCK_LENGTH 24

// unaligned little-endian loads built from aligned words (buffers carry >= 16 bytes of slack)
__device__ __forceinline__ u64 ld64u(const u8* __restrict__ p) {
  const uintptr_t a = (uintptr_t)p;
  const u64* q = (const u64*)(a & ~(uintptr_t)7);
  const int sh = (int)(a & 7) * 8;
  const u64 w0 = q[0];
  if (sh == 0) return w0;
  const u64 w1 = q[1];
  return (w0 >> sh) | (w1 << (64 - sh));
}
__device__ __forceinline__ u32 ld32u(const u8* __restrict__ p) {
  const uintptr_t a = (uintptr_t)p;
  const u32* q = (const u32*)(a & ~(uintptr_t)3);
  const int sh = (int)(a & 3) * 8;
  const u32 w0 = q[0];
  if (sh == 0) return w0;
  const u32 w1 = q[1];
  return __funnelshift_r(w0, w1, sh);
}

template <bool EXTRA>
__device__ __forceinline__ int lz_hash(const u8* __restrict__ src, int idx) {   // LZCodec.java:904-911
  const u64 v = (ld64u(src + idx) << 24) * LZ_HASH_SEED;
  return (int)(v >> (EXTRA ? (64 - 19) : (64 - 16)));
}

// hash table access: SMEM24 = u16 low plane + u8 high plane in shared memory, else u32 in global memory
template <bool SMEM24>
struct LzTable {
  u16* lo; u8* hi; i32* g;
  __device__ __forceinline__ int get(int h) const {
    if (SMEM24) return (int)lo[h] | ((int)hi[h] << 16);
    return g[h];
  }
  __device__ __forceinline__ void set(int h, int v) const {
    if (SMEM24) { lo[h] = (u16)v; hi[h] = (u8)(v >> 16); }
    else g[h] = v;
  }
};

// LZXCodec.findMatch (LZCodec.java:271-287): compares in 8-byte steps while bestLen + 8 <= maxMatch.
// Warp-wide: lane k compares step base+k; the first differing step decides.
__device__ __forceinline__ int lz_find_match(const u8* __restrict__ src, int srcIdx, int ref, int maxMatch, int lane) {
  const int nSteps = (maxMatch > 0) ? (maxMatch >> 3) : 0;
  for (int base = 0; base < nSteps; base += 32) {
    const int k = base + lane;
    u64 diff = 0;
    bool stop = true;           // lanes past the limit stop the search at the limit
    if (k < nSteps) {
      diff = ld64u(src + srcIdx + 8 * k) ^ ld64u(src + ref + 8 * k);
      stop = (diff != 0);
    }
    const u32 m = __ballot_sync(0xFFFFFFFFu, stop);
    if (m != 0) {
      const int first = __ffs(m) - 1;
      const int kk = base + first;
      if (kk >= nSteps) return nSteps * 8;
      const u64 d = __shfl_sync(0xFFFFFFFFu, diff, first);
      return kk * 8 + (__ffsll((long long)d) - 1) / 8;
    }
  }
  return nSteps * 8;
}

// emitLength (LZCodec.java:211-231); all lanes compute, lane 0 stores
__device__ __forceinline__ int lz_emit_length(u8* block, int idx, int length, int lane) {
  if (length < 254) { if (lane == 0) block[idx] = (u8)length; return idx + 1; }
  if (length < 65536 + 254) {
    length -= 254;
    if (lane == 0) { block[idx] = 254; block[idx + 1] = (u8)(length >> 8); block[idx + 2] = (u8)length; }
    return idx + 3;
  }
  length -= 255;
  if (lane == 0) { block[idx] = 255; block[idx + 1] = (u8)(length >> 16); block[idx + 2] = (u8)(length >> 8); block[idx + 3] = (u8)length; }
  return idx + 4;
}

__device__ __forceinline__ void warp_copy(u8* __restrict__ dst, const u8* __restrict__ src, int n, int lane) {
  for (int i = lane; i < n; i += 32) dst[i] = src[i];
}

// ordered insert of positions [from, to) (last writer per bucket wins, as the sequential loop :554-565)
template <bool EXTRA, bool SMEM24>
__device__ __forceinline__ void lz_bulk_insert(const LzTable<SMEM24>& T, const u8* __restrict__ src, int from, int to, int lane) {
  for (int base = from; base < to; base += 32) {
    const int p = base + lane;
    const bool on = p < to;
    const int h = on ? lz_hash<EXTRA>(src, p) : -1 - lane;     // distinct dummies for idle lanes
    const u32 peers = __match_any_sync(0xFFFFFFFFu, h);
    const bool winner = on && ((peers >> lane) <= 1u);        // no higher lane shares my bucket
    if (winner) T.set(h, p);
    __syncwarp();
  }
}

// ================================================================================================================
// forward: LZXCodec.forward (LZCodec.java:299-597).  One warp per block; B.cur -> B.alt.
// Outcome in res[b]: 1 = true, 0 = false (transform skipped), and on success res_len[b] = bytes produced.
// ================================================================================================================
template <bool EXTRA, bool SMEM24>
__global__ void __launch_bounds__(32) lz_forward_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P) {
  extern __shared__ __align__(16) u8 smem_raw[];
  const int lane = threadIdx.x;
  const int b = blockIdx.x;
  KzgBlock& B = blocks[b];
  int* res = P.result + 2 * b;
  if (lane == 0) { res[0] = 0; res[1] = 0; }
  if (B.status != 0 || !P.enabled[b]) return;
  const int count = B.curLen;
  const u8* __restrict__ src = B.cur;
  u8* __restrict__ dst = B.alt;
  if (count < LZ_MIN_BLOCK_LENGTH) return;                 // "if too small, skip" (:312-313)
  int mm = 4;
  if (B.dataType == KZG_DT_DNA) mm = 6;
  else if (B.dataType == KZG_DT_SMALL_ALPHABET) return;    // (:348-352)
  const int minMatch = mm;

  LzTable<SMEM24> T;
  const int hashSize = EXTRA ? (1 << 19) : (1 << 16);
  if (SMEM24) {
    T.lo = (u16*)smem_raw; T.hi = smem_raw + 2 * 65536; T.g = nullptr;
    for (int i = lane; i < 65536 / 2; i += 32) ((u32*)T.lo)[i] = 0;
    for (int i = lane; i < 65536 / 4; i += 32) ((u32*)T.hi)[i] = 0;
  } else {
    T.lo = nullptr; T.hi = nullptr; T.g = P.hashBuf + (i64)b * hashSize;
    for (int i = lane; i < hashSize; i += 32) T.g[i] = 0;
  }
  __syncwarp();

  const int tkCap = max(count / 5, 256);                   // tkBuf is never grown (:324-333; SURVEY E-3)
  u8* tkBuf = P.scratch + (i64)b * P.scratchStride;
  u8* mBuf = tkBuf + P.tkStride;
  u8* mLenBuf = mBuf + P.mStride;
  const int srcEnd = count - 16 - 2;
  const int maxDist = (srcEnd < 4 * LZ_MAX_DISTANCE1) ? LZ_MAX_DISTANCE1 : LZ_MAX_DISTANCE2;
  const u8 flagByte = (u8)(((maxDist == LZ_MAX_DISTANCE1) ? 0 : 1) | (((mm - 2) & 0x07) << 1));
  int srcIdx = 0, anchor = 0, dstIdx = 13;
  int mIdx = 0, mLenIdx = 0, tkIdx = 0;
  int repd0 = count, repd1 = count;
  int repIdx = 0, srcInc = 0;
  bool overflow = false;

  while (srcIdx < srcEnd) {
    int bestLen = 0;
    const int h0 = lz_hash<EXTRA>(src, srcIdx);
    const int ref0 = T.get(h0);
    T.set(h0, srcIdx);
    const int srcIdx1 = srcIdx + 1;
    int ref = srcIdx1 - (repIdx ? repd1 : repd0);
    const int minRef = max(srcIdx - maxDist, 0);
    const u32 cur1 = ld32u(src + srcIdx1);
    if ((ref > minRef) && (ld32u(src + ref) == cur1)) {
      bestLen = lz_find_match(src, srcIdx1, ref, min(srcEnd - srcIdx1, LZ_MAX_MATCH), lane);
    } else {
      ref = srcIdx1 - (repIdx ? repd0 : repd1);
      if ((ref > minRef) && (ld32u(src + ref) == cur1))
        bestLen = lz_find_match(src, srcIdx1, ref, min(srcEnd - srcIdx1, LZ_MAX_MATCH), lane);
    }
    if (bestLen < minMatch) {
      ref = ref0;
      if ((ref > minRef) && (ld32u(src + ref) == ld32u(src + srcIdx)))
        bestLen = lz_find_match(src, srcIdx, ref, min(srcEnd - srcIdx, LZ_MAX_MATCH), lane);
      if (bestLen < minMatch) {       // no good match
        srcIdx = srcIdx1 + (srcInc >> 6);
        srcInc++;
        repIdx = 0;
        continue;
      }
      if ((ref != srcIdx - repd0) && (ref != srcIdx - repd1)) {
        // check if better match at next position (:405-422)
        const int h1 = lz_hash<EXTRA>(src, srcIdx1);
        const int ref1 = T.get(h1);
        T.set(h1, srcIdx1);
        if ((ref1 > minRef + 1) && (ld32u(src + ref1 + bestLen - 3) == ld32u(src + srcIdx1 + bestLen - 3))) {
          const int bestLen1 = lz_find_match(src, srcIdx1, ref1, min(srcEnd - srcIdx1, LZ_MAX_MATCH), lane);
          if (bestLen1 >= bestLen) { ref = ref1; bestLen = bestLen1; srcIdx = srcIdx1; }
        }
        if (EXTRA) {
          const int srcIdx2 = srcIdx1 + 1;
          const int h2 = lz_hash<EXTRA>(src, srcIdx2);
          const int ref2 = T.get(h2);
          T.set(h2, srcIdx2);
          if ((ref2 > minRef + 2) && (ld32u(src + ref2 + bestLen - 3) == ld32u(src + srcIdx2 + bestLen - 3))) {
            const int bestLen2 = lz_find_match(src, srcIdx2, ref2, min(srcEnd - srcIdx2, LZ_MAX_MATCH), lane);
            if (bestLen2 >= bestLen) { ref = ref2; bestLen = bestLen2; srcIdx = srcIdx2; }
          }
        }
      }
      // extend backwards (:446-450)
      while ((srcIdx > anchor) && (ref > minRef) && (src[srcIdx - 1] == src[ref - 1])) { bestLen++; ref--; srcIdx--; }
      if (bestLen > LZ_MAX_MATCH) { ref += (bestLen - LZ_MAX_MATCH); srcIdx += (bestLen - LZ_MAX_MATCH); bestLen = LZ_MAX_MATCH; }
    } else {
      if ((bestLen >= LZ_MAX_MATCH) || (src[srcIdx] != src[ref - 1])) {
        srcIdx++;
        T.set(lz_hash<EXTRA>(src, srcIdx), srcIdx);
      } else {
        bestLen++; ref--;
      }
    }
    // emit match (:467-538)
    srcInc = 0;
    const int dist = srcIdx - ref;
    int token, mLenTh;
    if (dist == repd0) { token = 0x00; mLenTh = 3; }
    else if (dist == repd1) { token = 0x04; mLenTh = 3; }
    else {
      const int inc1 = dist >= 65536 ? 1 : 0, inc2 = dist >= 256 ? 1 : 0;
      if (mIdx + 3 > P.mStride) { overflow = true; break; }
      if (lane == 0) {
        int k = mIdx;
        mBuf[k] = (u8)(dist >> 16); k += inc1;
        mBuf[k] = (u8)(dist >> 8); k += inc2;
        mBuf[k] = (u8)dist;
      }
      mIdx += inc1 + inc2 + 1;
      token = (inc1 + inc2 + 1) << 3;
      mLenTh = 7;
    }
    const int mLen = bestLen - minMatch;
    if (mLen >= mLenTh) {
      token += mLenTh;
      if (mLenIdx + 4 > P.mLenStride) { overflow = true; break; }
      mLenIdx = lz_emit_length(mLenBuf, mLenIdx, mLen - mLenTh, lane);
    } else token += mLen;
    repd1 = repd0; repd0 = dist; repIdx = 1;
    const int litLen = srcIdx - anchor;
    if (tkIdx >= tkCap) { overflow = true; break; }        // Java: ArrayIndexOutOfBounds -> block error
    if (litLen == 0) {
      if (lane == 0) tkBuf[tkIdx] = (u8)token;
      tkIdx++;
    } else {
      if (litLen >= 7) {
        if (litLen >= (1 << 24)) { if (lane == 0) { res[0] = 0; } return; }
        if (lane == 0) tkBuf[tkIdx] = (u8)((7 << 5) | token);
        tkIdx++;
        dstIdx = lz_emit_length(dst, dstIdx, litLen - 7, lane);
      } else {
        if (lane == 0) tkBuf[tkIdx] = (u8)((litLen << 5) | token);
        tkIdx++;
      }
      if (dstIdx + litLen > B.cap) { overflow = true; break; }
      warp_copy(dst + dstIdx, src + anchor, litLen, lane);
      dstIdx += litLen;
    }
    // fill the hash table over the match and update positions (:553-565)
    anchor = srcIdx + bestLen;
    lz_bulk_insert<EXTRA, SMEM24>(T, src, srcIdx + 1, anchor, lane);
    srcIdx = anchor;
  }
  if (overflow) { if (lane == 0) atomicExch(&B.status, -KZG_ERR_PROCESS_BLOCK); return; }

  // emit last literals (:568-596)
  const int litLen = count - anchor;
  if (dstIdx + litLen + tkIdx + mIdx + mLenIdx >= count) return;       // res[0] stays 0 (false)
  if (tkIdx >= tkCap) { if (lane == 0) atomicExch(&B.status, -KZG_ERR_PROCESS_BLOCK); return; }
  if (litLen >= 7) {
    if (lane == 0) tkBuf[tkIdx] = (u8)(7 << 5);
    tkIdx++;
    dstIdx = lz_emit_length(dst, dstIdx, litLen - 7, lane);
  } else {
    if (lane == 0) tkBuf[tkIdx] = (u8)(litLen << 5);
    tkIdx++;
  }
  __syncwarp();
  warp_copy(dst + dstIdx, src + anchor, litLen, lane);
  dstIdx += litLen;
  if (lane == 0) {
    const u32 a = (u32)dstIdx, t = (u32)tkIdx, m = (u32)mIdx;
    dst[0] = (u8)a; dst[1] = (u8)(a >> 8); dst[2] = (u8)(a >> 16); dst[3] = (u8)(a >> 24);
    dst[4] = (u8)t; dst[5] = (u8)(t >> 8); dst[6] = (u8)(t >> 16); dst[7] = (u8)(t >> 24);
    dst[8] = (u8)m; dst[9] = (u8)(m >> 8); dst[10] = (u8)(m >> 16); dst[11] = (u8)(m >> 24);
    dst[12] = flagByte;
  }
  warp_copy(dst + dstIdx, tkBuf, tkIdx, lane); dstIdx += tkIdx;
  warp_copy(dst + dstIdx, mBuf, mIdx, lane); dstIdx += mIdx;
  warp_copy(dst + dstIdx, mLenBuf, mLenIdx, lane); dstIdx += mLenIdx;
  if (lane == 0) {
    res[1] = dstIdx;
    res[0] = (dstIdx <= count - (count / 100)) ? 1 : 0;
  }
}

int kzg_lz_forward_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, bool extra, bool smemTable) {
  if (!extra && smemTable) {
    static bool attr = false;
    const int smem = 3 * 65536;
    if (!attr) { CUDA_TRY(cudaFuncSetAttribute(lz_forward_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr = true; }
    lz_forward_kernel<false, true><<<nBlocks, 32, smem, s>>>(d_blocks, P);
  } else if (!extra) {
    lz_forward_kernel<false, false><<<nBlocks, 32, 0, s>>>(d_blocks, P);
  } else {
    lz_forward_kernel<true, false><<<nBlocks, 32, 0, s>>>(d_blocks, P);
  }
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(1);
  return 0;
}

