// lz_forward2.cu — LZ forward (`-t LZ` / `-t LZX`), two-phase bit-exact encoder (sm_100a).
//
// Replaces K/transform/LZCodec.java LZXCodec.forward (:299-597).  The reference's greedy
// parse consults a single-entry hash table whose content at position p is "the most recent inserted position
// with the same hash".  Every position below p is inserted except the ones the skip acceleration jumped over
// (srcInc >> 6, :399-400), so the table is *almost* parse-independent.  That splits the work:
//   phase 1 (data-parallel, every block at once, HBM-bound):
//     hash of every position -> stable LSD radix sort of positions by hash (2 x 8-bit passes) ->
//     prev[p] = previous position with the same hash -> len0[p] = length of the match against prev[p]
//     exactly as findMatch would report it (8-byte steps, capped at 255; 0 when the 4-byte pre-check fails).
//   phase 2 (one warp per block): the decision sequence itself.  32 upcoming visit positions are evaluated at
//     once (repeat-offset checks + the precomputed len0), a ballot finds the first position where the
//     reference would emit a match, the misses before it are committed in one step, and the match is
//     emitted with the reference's exact rules (lazy step, backward extension, token layout).
//     Positions jumped over by the acceleration are recorded in a bitmap; a candidate that falls at or below
//     the highest skipped position is resolved by walking the prev chain past skipped entries.
// The hash table of the reference no longer exists on the device; phase 2 touches prev/len0 sequentially.
#include "kzg_common.cuh"
#include "kzg_transforms.cuh"
#include <algorithm>
#include <vector>

#define LZ_HASH_SEED 0x1E35A7BDull
#define LZ_MAX_DISTANCE1 ((1 << 16) - 2)
#define LZ_MAX_DISTANCE2 ((1 << 24) - 2)
#define LZ_MAX_MATCH (65535 + 254 + 4)
#define LZ_MIN_BLOCK_LENGTH 24
#define LZF_WT 4096
#define LZF_WARPS 8

struct LzfBlock {                 // per-block scratch pointers (device)
  u32* hash;                      // hash of position p (16 or 19 bits)
  u32* sa; u32* sa2;              // positions sorted by hash
  u32* prev;                      // previous position with the same hash (0 = none)
  u8* len0;                       // findMatch(p, prev[p]) capped at 255, 0 if the candidate fails the pre-checks
  u32* skipped;                   // bitmap of positions the acceleration jumped over
  u32* hist; u32* offs;           // radix pass scratch
  u8* tk; u8* m; u8* ml;          // token / distance / match-length side buffers
  i32 n;                          // positions that take part (srcEnd + 1), 0 = block does not run
  i32 count, srcEnd, maxDist, minMatch, tkCap, mCap, mlCap;
};

__device__ __forceinline__ u64 lzf_ld64(const u8* __restrict__ p) {
  const uintptr_t a = (uintptr_t)p;
  const u64* q = (const u64*)(a & ~(uintptr_t)7);
  const int sh = (int)(a & 7) * 8;
  const u64 w0 = q[0];
  if (sh == 0) return w0;
  return (w0 >> sh) | (q[1] << (64 - sh));
}
__device__ __forceinline__ u32 lzf_ld32(const u8* __restrict__ p) {
  const uintptr_t a = (uintptr_t)p;
  const u32* q = (const u32*)(a & ~(uintptr_t)3);
  const int sh = (int)(a & 3) * 8;
  const u32 w0 = q[0];
  if (sh == 0) return w0;
  return __funnelshift_r(w0, q[1], sh);
}
// LZXCodec.findMatch (LZCodec.java:271-287), one thread
__device__ __forceinline__ int lzf_find_match(const u8* __restrict__ src, int a, int b, int maxMatch, int cap) {
  int bestLen = 0;
  while (bestLen + 8 <= maxMatch && bestLen < cap) {
    const u64 diff = lzf_ld64(src + a + bestLen) ^ lzf_ld64(src + b + bestLen);
    if (diff != 0) { bestLen += (__ffsll((long long)diff) - 1) >> 3; break; }
    bestLen += 8;
  }
  return bestLen;
}
// warp-wide exact findMatch (lane k compares step base + k)
__device__ __forceinline__ int lzf_find_match_warp(const u8* __restrict__ src, int srcIdx, int ref, int maxMatch, int lane) {
  const int nSteps = (maxMatch > 0) ? (maxMatch >> 3) : 0;
  for (int base = 0; base < nSteps; base += 32) {
    const int k = base + lane;
    u64 diff = 0;
    bool stop = true;
    if (k < nSteps) { diff = lzf_ld64(src + srcIdx + 8 * k) ^ lzf_ld64(src + ref + 8 * k); stop = (diff != 0); }
    const u32 m = __ballot_sync(0xFFFFFFFFu, stop);
    if (m != 0) {
      const int first = __ffs(m) - 1;
      const int kk = base + first;
      if (kk >= nSteps) return nSteps * 8;
      const u64 d = __shfl_sync(0xFFFFFFFFu, diff, first);
      return kk * 8 + ((__ffsll((long long)d) - 1) >> 3);
    }
  }
  return nSteps * 8;
}

// ---- phase 1a: per-block setup + hashes ----------------------------------------------------------------------------------------
__global__ void lzf_setup_kernel(const KzgBlock* __restrict__ blocks, int nBlocks, KzgXfParams P, LzfBlock* __restrict__ lb) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nBlocks) return;
  const KzgBlock& B = blocks[b];
  LzfBlock& L = lb[b];
  P.result[2 * b] = 0; P.result[2 * b + 1] = 0;
  L.n = 0;
  if (B.status != 0 || !P.enabled[b]) return;
  const int count = B.curLen;
  if (count < LZ_MIN_BLOCK_LENGTH) return;                               // (:312-313)
  int mm = 4;
  if (B.dataType == KZG_DT_DNA) mm = 6;
  else if (B.dataType == KZG_DT_SMALL_ALPHABET) return;                  // (:348-352)
  L.count = count; L.srcEnd = count - 16 - 2;
  L.maxDist = (L.srcEnd < 4 * LZ_MAX_DISTANCE1) ? LZ_MAX_DISTANCE1 : LZ_MAX_DISTANCE2;
  L.minMatch = mm;
  L.tkCap = max(count / 5, 256);                                         // tkBuf is never grown (:324-333)
  L.n = L.srcEnd + 2;                                                    // positions 0..srcEnd+1 can be visited or looked up (lazy steps)
}

template <bool EXTRA>
__global__ void lzf_hash_kernel(const KzgBlock* __restrict__ blocks, LzfBlock* __restrict__ lb) {
  const LzfBlock& L = lb[blockIdx.y];
  const int n = L.n;
  const u8* __restrict__ src = blocks[blockIdx.y].cur;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    const u64 v = (lzf_ld64(src + p) << 24) * LZ_HASH_SEED;             // LZCodec.java:904-911
    L.hash[p] = (u32)(v >> (EXTRA ? (64 - 19) : (64 - 16)));
  }
  if (blockIdx.x == 0) for (int i = threadIdx.x; i < (n + 31) / 32 + 1; i += blockDim.x) L.skipped[i] = 0;
}

// ---- phase 1b: stable LSD radix sort of positions by hash (8-bit digits) -----------------------------------------------------------
__global__ void __launch_bounds__(32 * LZF_WARPS) lzf_hist_kernel(LzfBlock* __restrict__ lb, int shift, int first) {
  __shared__ u32 cnt[LZF_WARPS][256];
  const LzfBlock& L = lb[blockIdx.y];
  const int n = L.n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x * LZF_WARPS + warp;
  const int nT = (n + LZF_WT - 1) / LZF_WT;
  if (tile >= nT) return;
  for (int i = lane; i < 256; i += 32) cnt[warp][i] = 0;
  __syncwarp();
  const int beg = tile * LZF_WT, end = min(beg + LZF_WT, n);
  for (int i = beg + lane; i < end; i += 32) {
    const u32 s = first ? (u32)i : L.sa[i];
    atomicAdd(&cnt[warp][(L.hash[s] >> shift) & 255], 1u);
  }
  __syncwarp();
  for (int d = lane; d < 256; d += 32) L.hist[(size_t)d * nT + tile] = cnt[warp][d];
}
__global__ void __launch_bounds__(1024) lzf_scan_kernel(LzfBlock* __restrict__ lb) {
  __shared__ u32 wsum[32];
  __shared__ u32 carry;
  const LzfBlock& L = lb[blockIdx.x];
  if (L.n <= 0) return;
  const int total = 256 * ((L.n + LZF_WT - 1) / LZF_WT);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < total; base += 1024) {
    const int i = base + threadIdx.x;
    const u32 v = (i < total) ? L.hist[i] : 0u;
    u32 incl = v;
    for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      u32 w = wsum[lane], wi = w;
      for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, wi, o); if (lane >= o) wi += t; }
      wsum[lane] = wi - w;
    }
    __syncthreads();
    const u32 excl = carry + wsum[warp] + incl - v;
    if (i < total) L.offs[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(32 * LZF_WARPS) lzf_scatter_kernel(LzfBlock* __restrict__ lb, int shift, int first) {
  __shared__ u32 pos[LZF_WARPS][256];
  const LzfBlock& L = lb[blockIdx.y];
  const int n = L.n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x * LZF_WARPS + warp;
  const int nT = (n + LZF_WT - 1) / LZF_WT;
  if (tile >= nT) return;
  for (int d = lane; d < 256; d += 32) pos[warp][d] = L.offs[(size_t)d * nT + tile];
  __syncwarp();
  const int beg = tile * LZF_WT, end = min(beg + LZF_WT, n);
  const u32 lower = (1u << lane) - 1;
  u32* __restrict__ out = first ? L.sa : L.sa2;       // pass 1 writes sa (from identity), pass 2 writes sa2 (from sa)
  for (int base = beg; base < end; base += 32) {
    const int i = base + lane;
    const bool on = i < end;
    u32 s = 0; int d = 256 + lane;
    if (on) { s = first ? (u32)i : L.sa[i]; d = (int)((L.hash[s] >> shift) & 255); }
    const u32 peers = __match_any_sync(0xFFFFFFFFu, d);
    if (on) out[pos[warp][d] + __popc(peers & lower)] = s;
    __syncwarp();
    if (on && (peers >> lane) <= 1u) pos[warp][d] += __popc(peers);
    __syncwarp();
  }
}
// sorted[i-1] precedes sorted[i] in (hash, position) order: same hash -> it is the previous occurrence
__global__ void lzf_prev_kernel(LzfBlock* __restrict__ lb, int extraPass) {
  const LzfBlock& L = lb[blockIdx.y];
  const int n = L.n;
  const u32* __restrict__ sorted = extraPass ? L.sa : L.sa2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const u32 s = sorted[i];
    u32 pv = 0;
    if (i > 0) { const u32 q = sorted[i - 1]; if (L.hash[q] == L.hash[s]) pv = q; }
    L.prev[s] = pv;
  }
}
// ---- phase 1c: candidate match lengths ---------------------------------------------------------------------------------------------
__global__ void lzf_cand_kernel(const KzgBlock* __restrict__ blocks, LzfBlock* __restrict__ lb) {
  const LzfBlock& L = lb[blockIdx.y];
  const int n = L.n;
  const u8* __restrict__ src = blocks[blockIdx.y].cur;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    const int ref = (int)L.prev[p];
    const int minRef = max(p - L.maxDist, 0);
    int len = 0;
    if (ref > minRef && lzf_ld32(src + ref) == lzf_ld32(src + p))
      len = lzf_find_match(src, p, ref, min(L.srcEnd - p, LZ_MAX_MATCH), 256);
    L.len0[p] = (u8)min(len, 255);
  }
}

// ---- phase 2: the walk -----------------------------------------------------------------------------------------------------------------
// (bitmap words are updated with atomics, which act at L2: read them with ld.global.cg so no stale L1 line is used)
__device__ __forceinline__ bool lzf_is_skipped(const u32* sk, int q) { return (__ldcg(sk + (q >> 5)) >> (q & 31)) & 1u; }
// the reference's table content for position x: most recent inserted position with x's hash
__device__ __forceinline__ int lzf_cand(const LzfBlock& L, int x, int lastSkip) {
  int q = (int)L.prev[x];
  while (q > 0 && q <= lastSkip && lzf_is_skipped(L.skipped, q)) q = (int)L.prev[q];
  return q;
}
__device__ __forceinline__ int lzf_emit_length(u8* block, int idx, int length, int lane) {   // emitLength (:211-231)
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

#define LZF_LOOKAHEAD 1024
struct LzfShared { volatile int pos; volatile int rep0; volatile int rep1; volatile int done; };

__device__ __forceinline__ void lzf_prefetch(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// warp 1 of the CTA: runs ahead of the walker and pulls into L1 what the walker is about to touch — the sequential
// streams (src, len0, prev), the two repeat-offset streams and the candidate lines src[prev[q]]
__device__ void lzf_prefetch_warp(const LzfBlock& L, const u8* __restrict__ src, LzfShared& S, int lane) {
  int pf = 0;
  const int n = L.n;
  while (!S.done) {
    const int pos = S.pos;
    const int target = min(pos + LZF_LOOKAHEAD, n);
    if (pf < pos) pf = pos;
    if (pf >= target) { __nanosleep(200); continue; }
    const int r0 = S.rep0, r1 = S.rep1;
    for (; pf < target; pf += 32) {
      const int q = pf + lane;
      if (q >= n) break;
      const int l0 = L.len0[q];
      if ((q & 31) == 0 && q + 128 < n) lzf_prefetch(src + q + 128);
      if (l0 > 0) { const u32 pv = L.prev[q]; lzf_prefetch(src + pv); }
      else if ((q & 31) == 0 && q + 64 < n) lzf_prefetch(L.prev + q + 64);
      if (lane == 0) { if (q + 1 - r0 > 0) lzf_prefetch(src + q + 1 - r0 + 32); if (q + 1 - r1 > 0) lzf_prefetch(src + q + 1 - r1 + 32); }
    }
  }
}

template <bool EXTRA>
__global__ void __launch_bounds__(64) lzf_walk_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P, LzfBlock* __restrict__ lb) {
  __shared__ LzfShared S;
  const int lane = threadIdx.x & 31, b = blockIdx.x;
  const LzfBlock L = lb[b];
  if (L.n <= 0) return;
  KzgBlock& B = blocks[b];
  int* res = P.result + 2 * b;
  const u8* __restrict__ src = B.cur;
  u8* __restrict__ dst = B.alt;
  const int count = L.count, srcEnd = L.srcEnd, maxDist = L.maxDist, minMatch = L.minMatch;
  if (threadIdx.x == 0) { S.pos = 0; S.rep0 = count; S.rep1 = count; S.done = 0; }
  __syncthreads();
  if (threadIdx.x >= 32) { lzf_prefetch_warp(L, src, S, lane); return; }

  const u8 flagByte = (u8)(((maxDist == LZ_MAX_DISTANCE1) ? 0 : 1) | (((minMatch - 2) & 0x07) << 1));
  u8* tkBuf = L.tk; u8* mBuf = L.m; u8* mLenBuf = L.ml;
  int srcIdx = 0, anchor = 0, dstIdx = 13;
  int mIdx = 0, mLenIdx = 0, tkIdx = 0;
  int repd0 = count, repd1 = count;
  int repIdx = 0, srcInc = 0;
  int lastSkip = -1;                         // highest position ever jumped over (may since have been re-inserted)
  bool overflow = false, giveUp = false;
  const int dbg = P.flags >> 12;
  long long t0 = clock64(); int nIter = 0, nEv = 0, nSlow = 0, nFmw = 0;

  while (srcIdx < srcEnd) {
    nIter++;
    if (lane == 0) { S.pos = srcIdx; S.rep0 = repd0; S.rep1 = repd1; }
    // ---- evaluate the next 32 visit positions, assuming the ones before each are misses ----
    const u32 stepExtra = (u32)((srcInc + lane) >> 6);
    u32 exIncl = stepExtra;
    for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, exIncl, o); if (lane >= o) exIncl += t; }
    const int p = srcIdx + lane + (int)(exIncl - stepExtra);
    const bool valid = p < srcEnd;
    bool hit = false, slow = false;
    int l0 = 0, pv = 0, repSmall = 0, repRef = 0;
    if (valid) {
      const int p1 = p + 1;
      const int minRef = max(p - maxDist, 0);
      const int maxM = min(srcEnd - p1, LZ_MAX_MATCH);
      const int rA = (lane == 0 && repIdx) ? repd1 : repd0, rB = (lane == 0 && repIdx) ? repd0 : repd1;
      const int refA = p1 - rA, refB = p1 - rB;
      // independent loads first
      const u64 n8 = lzf_ld64(src + p1);
      const u64 a8 = (refA > minRef) ? lzf_ld64(src + refA) : ~n8;
      const u64 b8 = (refB > minRef) ? lzf_ld64(src + refB) : ~n8;
      l0 = L.len0[p];
      pv = (int)L.prev[p];
      // the reference tries repd[repIdx] first and only falls to the other one when the 4-byte pre-check fails (:374-387)
      u64 diff; 
      if ((u32)(a8 ^ n8) == 0) { diff = a8 ^ n8; repRef = refA; }
      else if ((u32)(b8 ^ n8) == 0) { diff = b8 ^ n8; repRef = refB; }
      else { diff = 1; repRef = 0; }
      if (repRef > 0) repSmall = (maxM < 8) ? 0 : ((diff == 0) ? 8 : ((__ffsll((long long)diff) - 1) >> 3));
      if (repSmall >= minMatch) hit = true;
      else {
        if (lastSkip >= 0 && pv > 0 && pv <= lastSkip) slow = true;
        if (((srcInc + 31) >> 6) > 0 && pv > srcIdx) slow = true;      // may be a position this very batch jumps over
        if (!slow && l0 >= minMatch) hit = true;
      }
    }
    const u32 stopMask = __ballot_sync(0xFFFFFFFFu, hit || slow);
    const u32 validMask = __ballot_sync(0xFFFFFFFFu, valid);
    const int nValid = __popc(validMask);
    const int nMiss = stopMask ? min(__ffs(stopMask) - 1, nValid) : nValid;
    // ---- commit the misses: positions jumped over by the acceleration are recorded (they are never inserted) ----
    if (nMiss > 0) {
      const int lastEx = __shfl_sync(0xFFFFFFFFu, (int)stepExtra, nMiss - 1);
      if (lastEx > 0) {
        if (lane < nMiss && stepExtra > 0) {
          for (u32 k = 1; k <= stepExtra; k++) { const int q = p + (int)k; if (q <= srcEnd) atomicOr(&L.skipped[q >> 5], 1u << (q & 31)); }
        }
        int mx = (lane < nMiss && stepExtra > 0) ? p + (int)stepExtra : -1;
        for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
        lastSkip = max(lastSkip, min(mx, srcEnd));
        __threadfence_block();
      }
      const int lastP = __shfl_sync(0xFFFFFFFFu, p, nMiss - 1);
      srcIdx = lastP + 1 + lastEx;
      srcInc += nMiss;
      repIdx = 0;
      __syncwarp();
    }
    if (!stopMask || nMiss >= nValid) continue;
    if (srcIdx >= srcEnd) break;
    // ---- one iteration of the reference loop at srcIdx (:366-566), exact; lane `f` already holds this position's loads ----
    nEv++;
    const int f = nMiss;
    const int evL0 = __shfl_sync(0xFFFFFFFFu, l0, f);
    const int evPv = __shfl_sync(0xFFFFFFFFu, pv, f);
    const int evRepSmall = __shfl_sync(0xFFFFFFFFu, repSmall, f);
    const int evRepRef = __shfl_sync(0xFFFFFFFFu, repRef, f);
    const bool evSlow = (__shfl_sync(0xFFFFFFFFu, (int)slow, f) != 0);
    // lane f+1 holds position srcIdx + 1 when lane f's step is 1
    const int nextOk = (f + 1 < 32) && (((validMask >> (f + 1)) & 1u) != 0) && (__shfl_sync(0xFFFFFFFFu, (int)stepExtra, f) == 0);
    const int nxL0 = __shfl_sync(0xFFFFFFFFu, l0, (f + 1) & 31);
    const int nxPv = __shfl_sync(0xFFFFFFFFu, pv, (f + 1) & 31);
    int bestLen = 0;
    const int srcIdx1 = srcIdx + 1;
    const int minRef = max(srcIdx - maxDist, 0);
    int ref = evRepRef;
    if (ref > 0) bestLen = (evRepSmall < 8) ? evRepSmall : lzf_find_match_warp(src, srcIdx1, ref, min(srcEnd - srcIdx1, LZ_MAX_MATCH), lane);
    if (bestLen < minMatch) {
      // check match at position in hash table (:389-395): table content = first entry of the prev chain that was inserted
      int ref0 = evPv;
      bool firstHop = true;
      if (evSlow) { nSlow++; const int q = lzf_cand(L, srcIdx, lastSkip); firstHop = (q == evPv); ref0 = q; }
      ref = ref0;
      int hl;
      if (firstHop) {
        // len0 already is ((ref > minRef) && 4-byte check) ? findMatch(srcIdx, ref) : 0, capped at 255
        hl = (evL0 < 255) ? evL0 : lzf_find_match_warp(src, srcIdx, ref, min(srcEnd - srcIdx, LZ_MAX_MATCH), lane);
      } else if ((ref > minRef) && (lzf_ld32(src + ref) == lzf_ld32(src + srcIdx))) {
        hl = lzf_find_match_warp(src, srcIdx, ref, min(srcEnd - srcIdx, LZ_MAX_MATCH), lane);
      } else hl = 0;
      // (when the hash candidate fails its pre-check the reference keeps the too-short repeat result: a miss either way)
      bestLen = (hl >= minMatch) ? hl : 0;
      if (bestLen < minMatch) {       // no good match
        const int ex = srcInc >> 6;
        if (ex > 0) {
          if (lane == 0) for (int k = 1; k <= ex; k++) { const int q = srcIdx + k; if (q <= srcEnd) atomicOr(&L.skipped[q >> 5], 1u << (q & 31)); }
          lastSkip = max(lastSkip, min(srcIdx + ex, srcEnd));
          __threadfence_block();
          __syncwarp();
        }
        srcIdx = srcIdx1 + ex;
        srcInc++;
        repIdx = 0;
        continue;
      }
      if ((ref != srcIdx - repd0) && (ref != srcIdx - repd1)) {
        // check if better match at next position (:405-422); the table lookup there happens before srcIdx1 is inserted
        int ref1; bool hop1 = false; int l01 = 255;
        if (!(dbg & 2) && nextOk && !(lastSkip >= 0 && nxPv > 0 && nxPv <= lastSkip)) { ref1 = nxPv; hop1 = true; l01 = nxL0; }
        else ref1 = lzf_cand(L, srcIdx1, lastSkip);
        if (ref1 > minRef + 1) {
          bool sw = false; int bestLen1 = 0;
          if (hop1 && l01 < 255 && l01 != bestLen) {
            // len0 is findMatch(srcIdx1, ref1) (or 0 when its 4-byte pre-check fails, which also means < 4 <= bestLen)
            if (l01 > bestLen) { sw = true; bestLen1 = l01; }
          } else if (lzf_ld32(src + ref1 + bestLen - 3) == lzf_ld32(src + srcIdx1 + bestLen - 3)) {
            bestLen1 = lzf_find_match_warp(src, srcIdx1, ref1, min(srcEnd - srcIdx1, LZ_MAX_MATCH), lane);
            sw = (bestLen1 >= bestLen);
          }
          if (sw) { ref = ref1; bestLen = bestLen1; srcIdx = srcIdx1; }
        }
        if (EXTRA) {
          const int srcIdx2 = srcIdx1 + 1;
          const int ref2 = lzf_cand(L, srcIdx2, lastSkip);
          if ((ref2 > minRef + 2) && (lzf_ld32(src + ref2 + bestLen - 3) == lzf_ld32(src + srcIdx2 + bestLen - 3))) {
            const int bestLen2 = lzf_find_match_warp(src, srcIdx2, ref2, min(srcEnd - srcIdx2, LZ_MAX_MATCH), lane);
            if (bestLen2 >= bestLen) { ref = ref2; bestLen = bestLen2; srcIdx = srcIdx2; }
          }
        }
      }
      // extend backwards (:446-450), 8 bytes per step
      const int visited = srcIdx;
      while (true) {
        const int room = min(srcIdx - anchor, ref - minRef);
        if (room <= 0) break;
        const bool wide = !(dbg & 4) && (srcIdx >= 8) && (ref >= 8);
        int e;
        if (wide) {
          const u64 d = lzf_ld64(src + srcIdx - 8) ^ lzf_ld64(src + ref - 8);
          e = (d == 0) ? 8 : (__clzll((long long)d) >> 3);      // equal bytes counted from the end (highest byte = position - 1)
        } else {
          e = (src[srcIdx - 1] == src[ref - 1]) ? 1 : 0;
        }
        const int lim = wide ? 8 : 1;
        const int take = min(e, room);
        bestLen += take; ref -= take; srcIdx -= take;
        if (take < lim) break;
      }
      if (bestLen > LZ_MAX_MATCH) { ref += (bestLen - LZ_MAX_MATCH); srcIdx += (bestLen - LZ_MAX_MATCH); bestLen = LZ_MAX_MATCH; }
      // the match interior is inserted again (:553-565): positions jumped over earlier inside it become table entries
      if (lastSkip > srcIdx && srcIdx < visited) {
        const int hi = min(visited, lastSkip);
        for (int q = srcIdx + 1 + lane; q <= hi; q += 32) atomicAnd(&L.skipped[q >> 5], ~(1u << (q & 31)));
        __threadfence_block();
        __syncwarp();
      }
    } else {
      if ((bestLen >= LZ_MAX_MATCH) || (src[srcIdx] != src[ref - 1])) srcIdx++;
      else { bestLen++; ref--; }
    }
    // emit match (:467-538)
    srcInc = 0;
    const int dist = srcIdx - ref;
    int token, mLenTh;
    if (dist == repd0) { token = 0x00; mLenTh = 3; }
    else if (dist == repd1) { token = 0x04; mLenTh = 3; }
    else {
      const int inc1 = dist >= 65536 ? 1 : 0, inc2 = dist >= 256 ? 1 : 0;
      if (mIdx + 3 > L.mCap) { overflow = true; break; }
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
      if (mLenIdx + 4 > L.mlCap) { overflow = true; break; }
      mLenIdx = lzf_emit_length(mLenBuf, mLenIdx, mLen - mLenTh, lane);
    } else token += mLen;
    repd1 = repd0; repd0 = dist; repIdx = 1;
    const int litLen = srcIdx - anchor;
    if (tkIdx >= L.tkCap) { overflow = true; break; }        // Java: ArrayIndexOutOfBounds -> block error
    if (litLen == 0) {
      if (lane == 0) tkBuf[tkIdx] = (u8)token;
      tkIdx++;
    } else {
      if (litLen >= 7) {
        if (litLen >= (1 << 24)) { giveUp = true; break; }     // forward returns false (:523-524)
        if (lane == 0) tkBuf[tkIdx] = (u8)((7 << 5) | token);
        tkIdx++;
        dstIdx = lzf_emit_length(dst, dstIdx, litLen - 7, lane);
      } else {
        if (lane == 0) tkBuf[tkIdx] = (u8)((litLen << 5) | token);
        tkIdx++;
      }
      if (dstIdx + litLen > B.cap) { overflow = true; break; }
      for (int i = lane; i < litLen; i += 32) dst[dstIdx + i] = src[anchor + i];
      dstIdx += litLen;
    }
    anchor = srcIdx + bestLen;
    srcIdx = anchor;
  }
  if (lane == 0) S.done = 1;
  if ((dbg & 1) && lane == 0) printf("lzf block %d count %d: iters %d events %d slow %d tokens %d cycles %lld\n", b, count, nIter, nEv, nSlow, tkIdx, clock64() - t0);
  if (giveUp) return;
  if (overflow) { if (lane == 0) atomicExch(&B.status, -KZG_ERR_PROCESS_BLOCK); return; }

  // emit last literals (:568-596)
  const int litLen = count - anchor;
  if (dstIdx + litLen + tkIdx + mIdx + mLenIdx >= count) return;       // forward returns false
  if (tkIdx >= L.tkCap) { if (lane == 0) atomicExch(&B.status, -KZG_ERR_PROCESS_BLOCK); return; }
  if (litLen >= 7) {
    if (lane == 0) tkBuf[tkIdx] = (u8)(7 << 5);
    tkIdx++;
    dstIdx = lzf_emit_length(dst, dstIdx, litLen - 7, lane);
  } else {
    if (lane == 0) tkBuf[tkIdx] = (u8)(litLen << 5);
    tkIdx++;
  }
  __syncwarp();
  for (int i = lane; i < litLen; i += 32) dst[dstIdx + i] = src[anchor + i];
  dstIdx += litLen;
  if (lane == 0) {
    const u32 a = (u32)dstIdx, t = (u32)tkIdx, m = (u32)mIdx;
    dst[0] = (u8)a; dst[1] = (u8)(a >> 8); dst[2] = (u8)(a >> 16); dst[3] = (u8)(a >> 24);
    dst[4] = (u8)t; dst[5] = (u8)(t >> 8); dst[6] = (u8)(t >> 16); dst[7] = (u8)(t >> 24);
    dst[8] = (u8)m; dst[9] = (u8)(m >> 8); dst[10] = (u8)(m >> 16); dst[11] = (u8)(m >> 24);
    dst[12] = flagByte;
  }
  for (int i = lane; i < tkIdx; i += 32) dst[dstIdx + i] = tkBuf[i];
  dstIdx += tkIdx;
  for (int i = lane; i < mIdx; i += 32) dst[dstIdx + i] = mBuf[i];
  dstIdx += mIdx;
  for (int i = lane; i < mLenIdx; i += 32) dst[dstIdx + i] = mLenBuf[i];
  dstIdx += mLenIdx;
  if (lane == 0) { res[1] = dstIdx; res[0] = (dstIdx <= count - (count / 100)) ? 1 : 0; }
}

// ---- host ------------------------------------------------------------------------------------------------------------------------------
static size_t lzf_al(size_t v) { return (v + 255) / 256 * 256; }
struct LzfSizes { size_t hash, sa, prev, len0, skipped, hist, tk, m, ml, total; };
static LzfSizes lzf_sizes(i32 maxLen) {
  LzfSizes z;
  const size_t n = (size_t)maxLen + 64;
  const size_t nT = (n + LZF_WT - 1) / LZF_WT;
  z.hash = lzf_al(4 * n); z.sa = lzf_al(4 * n); z.prev = lzf_al(4 * n); z.len0 = lzf_al(n); z.skipped = lzf_al(n / 8 + 64);
  z.hist = lzf_al(4 * 256 * nT);
  z.tk = lzf_al(std::max<size_t>(n / 5, 256) + 64); z.m = lzf_al(n + 64); z.ml = lzf_al(n / 2 + 64);
  z.total = z.hash + 2 * z.sa + z.prev + z.len0 + z.skipped + 2 * z.hist + z.tk + z.m + z.ml + 1024;
  return z;
}
void kzg_lzf_scratch(i32 maxLen, size_t* perBlockBytes) { *perBlockBytes = std::max(*perBlockBytes, lzf_sizes(maxLen).total + sizeof(LzfBlock) + 256); }

int kzg_lz_forward2_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, bool extra, i32 maxLen) {
  const LzfSizes z = lzf_sizes(maxLen);
  const size_t nb = (size_t)nBlocks;
  if (nb * (z.total + 256) > nb * (size_t)P.scratchStride) { kzg_set_error("lz forward: scratch pool too small"); return -KZG_ERR_CREATE_CODEC; }
  // flat pool: [LzfBlock descriptors][per-block areas]
  LzfBlock* dlb = (LzfBlock*)P.scratch;
  u8* base = P.scratch + lzf_al(nb * sizeof(LzfBlock));
  std::vector<LzfBlock> hl(nBlocks);
  for (int b = 0; b < nBlocks; b++) {
    u8* o = base + (size_t)b * z.total;
    LzfBlock& L = hl[b];
    memset(&L, 0, sizeof(L));
    L.hash = (u32*)o; o += z.hash; L.sa = (u32*)o; o += z.sa; L.sa2 = (u32*)o; o += z.sa; L.prev = (u32*)o; o += z.prev;
    L.len0 = o; o += z.len0; L.skipped = (u32*)o; o += z.skipped; L.hist = (u32*)o; o += z.hist; L.offs = (u32*)o; o += z.hist;
    L.tk = o; o += z.tk; L.m = o; o += z.m; L.ml = o; o += z.ml;
    L.mCap = (i32)z.m - 16; L.mlCap = (i32)z.ml - 16;
  }
  CUDA_TRY(cudaMemcpyAsync(dlb, hl.data(), sizeof(LzfBlock) * nb, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaStreamSynchronize(s));          // hl is stack-owned
  lzf_setup_kernel<<<(nBlocks + 63) / 64, 64, 0, s>>>(d_blocks, nBlocks, P, dlb);
  const int gx = std::max(1, std::min((maxLen + 255) / 256, 8 * KZG_SM_COUNT));
  if (extra) lzf_hash_kernel<true><<<dim3(gx, nBlocks), 256, 0, s>>>(d_blocks, dlb);
  else lzf_hash_kernel<false><<<dim3(gx, nBlocks), 256, 0, s>>>(d_blocks, dlb);
  const int nT = (maxLen + LZF_WT - 1) / LZF_WT;
  dim3 gridT((nT + LZF_WARPS - 1) / LZF_WARPS, nBlocks);
  const int bits = extra ? 19 : 16;
  int pass = 0;
  for (int shift = 0; shift < bits; shift += 8, pass++) {
    // pass 0: identity -> sa; later passes ping-pong sa -> sa2 -> (swap by kernel argument is not possible) so copy back
    lzf_hist_kernel<<<gridT, 32 * LZF_WARPS, 0, s>>>(dlb, shift, pass == 0 ? 1 : 0);
    lzf_scan_kernel<<<nBlocks, 1024, 0, s>>>(dlb);
    lzf_scatter_kernel<<<gridT, 32 * LZF_WARPS, 0, s>>>(dlb, shift, pass == 0 ? 1 : 0);
    if (pass >= 1 && shift + 8 < bits) {       // a third pass (19-bit hash) reads sa again: move sa2 back
      for (int b = 0; b < nBlocks; b++) CUDA_TRY(cudaMemcpyAsync(hl[b].sa, hl[b].sa2, z.sa, cudaMemcpyDeviceToDevice, s));
    }
  }
  lzf_prev_kernel<<<dim3(gx, nBlocks), 256, 0, s>>>(dlb, pass == 1 ? 1 : 0);
  lzf_cand_kernel<<<dim3(gx, nBlocks), 256, 0, s>>>(d_blocks, dlb);
  if (extra) lzf_walk_kernel<true><<<nBlocks, 64, 0, s>>>(d_blocks, P, dlb);
  else lzf_walk_kernel<false><<<nBlocks, 64, 0, s>>>(d_blocks, P, dlb);
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(5 + 3 * pass);
  return 0;
}
