// lz_forward2.cu — LZ forward (`-t LZ` / `-t LZX`), bit-exact encoder (sm_100a).
//
// Replaces K/transform/LZCodec.java LZXCodec.forward (:299-597).  The reference's greedy
// parse consults a single-entry hash table whose content at position p is "the most recent inserted position
// with the same hash".  Every position below p is inserted except the ones the skip acceleration jumped over
// (srcInc >> 6, :399-400), so the table is *almost* parse-independent.  That splits the work (DESIGN.md "LZ forward"):
//   phase 1 (data parallel, HBM-bound): hash of every position -> stable LSD radix sort of packed key|position
//     elements by hash -> prev[p] / hs[] / rank[] -> len0[p] = findMatch(p, prev[p]) capped at 255; a look-back
//     through the hash class (4-byte fingerprint carried in the key) flags positions that can never hit the table.
//   head: the first 2 KiB of every block parsed for real; their jumped-over bits seed the assumed bitmap A.
//   segments (lzf_spec_kernel, one warp per 8 KiB segment): speculative parses after a warm-up, 32 visit positions per
//     batch, with A for positions before the segment; matches are logged, lookups that depended on A are marked.
//   stitch (lzf_stitch_kernel, one warp per block): adopts the segments whose entry state was the true one, parses the
//     rest itself; sparse stretches (noise, PCM) are parsed in order against a real table (lzf_direct_core).
//   check (lzf_check_marks_kernel): the segments whose marked lookups resolve differently under the produced bitmap are
//     parsed again next round with A := Kn; fixed point = the reference's parse.
//   emit: tokens, distances, lengths and literals from the match list with prefix sums.
// Every block group runs this pipeline on a CUDA stream of its own (kzg_lz_forward2_launch).  lzf_walk_kernel is the
// first-generation exact one-warp walker, kept as the fallback for blocks that do not settle within 8 rounds.
#include "kzg_common.cuh"
#include "kzg_transforms.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstdio>
#include <vector>
#include <thread>
#include <chrono>

#define LZ_HASH_SEED 0x1E35A7BDull
#define LZ_MAX_DISTANCE1 ((1 << 16) - 2)
#define LZ_MAX_DISTANCE2 ((1 << 24) - 2)
#define LZ_MAX_MATCH (65535 + 254 + 4)
#define LZ_MIN_BLOCK_LENGTH 24
#define LZF_WT 4096
#define LZF_FP_LOOKBACK 1024
#define LZF_WARPS 8

struct LzfBlock {                 // per-block scratch pointers (device)
  u64* ka; u64* kb;               // radix ping-pong of (sort key << 30 | position); key = hash | fingerprint << hash bits
  u32* hs;                        // the hash-sorted order kept: hs[i] = position | LZF_RUNSTART when it opens its hash class
  u32* rank;                      // rank[p] = index of p in hs: the chain of p is hs[rank[p]-1], hs[rank[p]-2], ... (contiguous)
  u32* prev;                      // previous position with the same hash (0 = none)
  u8* len0;                       // findMatch(p, prev[p]) capped at 255, 0 if the candidate fails the pre-checks
  u32* skipped;                   // bitmap of positions the acceleration jumped over
  u32* hist; u32* offs;           // radix pass scratch
  u8* tk; u8* m; u8* ml;          // token / distance / match-length side buffers
  i32 n;                          // positions that take part (srcEnd + 1), 0 = block does not run
  i32 count, srcEnd, maxDist, minMatch, tkCap, mCap, mlCap;
  // segment-parallel parse (phase 2/3/4)
  u32* A;                         // skipped-position bitmap every segment assumes for positions before its own start
  u32* Kn;                        // skipped-position bitmap of the stitched parse (next round's assumption)
  u32* D;                         // skipped positions each speculative segment found inside its own range
  u32* C;                         // positions before a segment whose assumed state a lookup of that segment depended on
  uint4* specEv;                  // per-segment match logs {start, length, distance, -}
  uint4* patchEv;                 // matches the stitcher had to find itself
  struct LzfSeg* seg;             // per-segment end state
  struct LzfRange* rng;           // final match list as ranges of the two logs
  uint4* fin;                     // the stitched match list, contiguous
  u32* tileSum;                   // per 1024 matches: literal-area / distance / length bytes (sums, then offsets)
  i32 segLen, nSeg, evStride, patchCap, nRng, aMax, active, needSerial;
  i32 nFin, giveUpIdx, emitGo, tkBase, mBase, mlBase;
  i32 chkDiff, chkMax, chkFlag;           // per-round verdict of lzf_check_marks_kernel
  i32 direct;                             // the stitcher parsed the whole block itself against a real table (sparse block)
  i32 fpLookback;                         // how far lzf_prev_kernel looks back through a hash class for an earlier equal fingerprint (0: dense block, flag not needed)
  i32 estHits[8];                         // per eighth of the block: positions whose hash candidate is a match of 4+ bytes (how dense the parse will be)
};
struct LzfState { i32 srcIdx, anchor, srcInc, repd0, repd1, repIdx, lastSkip, overLo, overHi; };
// rerun: a lookup of this segment's adopted parse resolved differently under the produced bitmap -> parse it again next round;
// adopted: the stitched parse of this round took (part of) the segment's own log
struct LzfSeg { LzfState entry; LzfState end; LzfState trueEntry; i32 nEv; i32 fail; i32 haveTrue; int16_t rerun; int16_t adopted; i32 nSkip; i32 pad; };
struct LzfRange { const uint4* ev; i32 count; i32 start; };

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
__global__ void lzf_setup_kernel(const KzgBlock* __restrict__ blocks, int nBlocks, KzgXfParams P, LzfBlock* __restrict__ lb, int segLen, int forceSerial) {
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
  L.segLen = segLen; L.nSeg = max(1, (L.srcEnd + segLen - 1) / segLen);
  L.nRng = 0; L.aMax = -1; L.active = forceSerial ? 0 : 1; L.needSerial = forceSerial; L.direct = 0; L.fpLookback = LZF_FP_LOOKBACK; for (int k = 0; k < 8; k++) L.estHits[k] = 0;
}

template <bool EXTRA>
__global__ void lzf_hash_kernel(const KzgBlock* __restrict__ blocks, LzfBlock* __restrict__ lb, const int* __restrict__ bmap) {
  const int b = bmap[blockIdx.y];
  const LzfBlock& L = lb[b];
  const int n = L.n;
  const u8* __restrict__ src = blocks[b].cur;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    const u64 w = lzf_ld64(src + p);
    const u64 v = (w << 24) * LZ_HASH_SEED;                              // LZCodec.java:904-911
    const u64 hash = v >> (EXTRA ? (64 - 19) : (64 - 16));
    // 16-bit (13 with the 19-bit hash) fingerprint of the first 4 bytes: second sort key
    const u64 fp = ((u32)w * 0x9E3779B1u) >> (EXTRA ? 19 : 16);
    L.ka[p] = ((hash | (fp << (EXTRA ? 19 : 16))) << 30) | (u64)p;
  }
  if (blockIdx.x == 0) for (int i = threadIdx.x; i < (n + 31) / 32 + 2; i += blockDim.x) { L.skipped[i] = 0; L.A[i] = 0; }
}

// ---- phase 1b: stable LSD radix sort of positions by hash (8-bit digits) -----------------------------------------------------------
// elements carry their key, so every pass streams: pass k reads ka (k even) or kb (k odd) and writes the other one
#define LZF_NOCAND 0x80000000u
#define LZF_RUNSTART 0x80000000u
#define LZF_POSMASK ((1ull << 30) - 1)
__global__ void __launch_bounds__(32 * LZF_WARPS) lzf_hist_kernel(LzfBlock* __restrict__ lb, const int* __restrict__ bmap, int shift, int mask, int k) {
  __shared__ u32 cnt[LZF_WARPS][256];
  const LzfBlock& L = lb[bmap[blockIdx.y]];
  const int n = L.n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x * LZF_WARPS + warp;
  const int nT = (n + LZF_WT - 1) / LZF_WT;
  if (tile >= nT) return;
  for (int i = lane; i < 256; i += 32) cnt[warp][i] = 0;
  __syncwarp();
  const int beg = tile * LZF_WT, end = min(beg + LZF_WT, n);
  const u64* __restrict__ in = (k & 1) ? L.kb : L.ka;
  for (int i = beg + lane; i < end; i += 32) atomicAdd(&cnt[warp][(int)(in[i] >> shift) & mask], 1u);
  __syncwarp();
  for (int d = lane; d < 256; d += 32) L.hist[(size_t)d * nT + tile] = cnt[warp][d];
}
__global__ void __launch_bounds__(1024) lzf_scan_kernel(LzfBlock* __restrict__ lb, const int* __restrict__ bmap) {
  __shared__ u32 wsum[32];
  __shared__ u32 carry;
  const LzfBlock& L = lb[bmap[blockIdx.x]];
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
// One CTA per tile of 4096 elements.  Every warp ranks its 512 consecutive elements by digit (16 rounds of
// __match_any_sync), the per-warp digit counts are scanned across warps and digits, the elements are staged in shared
// memory in digit order and leave from there: consecutive threads write consecutive elements of a digit run, so the
// stores cover whole sectors (element-wise scattering costs ~3x the DRAM traffic in partial-sector fills and evictions).
__global__ void __launch_bounds__(32 * LZF_WARPS, 3) lzf_scatter_kernel(LzfBlock* __restrict__ lb, const int* __restrict__ bmap, int shift, int mask, int k) {
  __shared__ u64 stage[LZF_WT];
  __shared__ u32 cnt[LZF_WARPS][256];
  __shared__ u32 digitBase[256 + 1];
  __shared__ u32 gOff[256];
  __shared__ u32 wsum[LZF_WARPS];
  const LzfBlock& L = lb[bmap[blockIdx.y]];
  const int n = L.n;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int nT = (n + LZF_WT - 1) / LZF_WT;
  if (tile >= nT) return;
  for (int i = tid; i < LZF_WARPS * 256; i += 32 * LZF_WARPS) (&cnt[0][0])[i] = 0;
  gOff[tid] = L.offs[(size_t)tid * nT + tile];
  __syncthreads();
  const int beg = tile * LZF_WT, end = min(beg + LZF_WT, n);
  const u32 lower = (1u << lane) - 1;
  const u64* __restrict__ in = (k & 1) ? L.kb : L.ka;
  u64* __restrict__ out = (k & 1) ? L.ka : L.kb;
  constexpr int R = LZF_WT / (32 * LZF_WARPS);          // rounds per warp
  u64 v[R]; u32 rk[R];
  const int wbeg = beg + warp * (32 * R);
  #pragma unroll
  for (int r = 0; r < R; r++) { const int i = wbeg + r * 32 + lane; v[r] = (i < end) ? in[i] : 0ull; }
  #pragma unroll
  for (int r = 0; r < R; r++) {
    const int i = wbeg + r * 32 + lane;
    const bool on = i < end;
    const int d = on ? ((int)(v[r] >> shift) & mask) : (256 + lane);
    const u32 peers = __match_any_sync(0xFFFFFFFFu, d);
    rk[r] = on ? (cnt[warp][d] + __popc(peers & lower)) : 0u;
    __syncwarp();
    if (on && (peers >> lane) <= 1u) cnt[warp][d] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // digit tid: exclusive scan over the warps, then over the digits
  u32 tot = 0;
  #pragma unroll
  for (int w = 0; w < LZF_WARPS; w++) { const u32 c = cnt[w][tid]; cnt[w][tid] = tot; tot += c; }
  u32 incl = tot;
  for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  u32 wb = 0;
  for (int w = 0; w < warp; w++) wb += wsum[w];
  digitBase[tid] = wb + incl - tot;
  __syncthreads();
  #pragma unroll
  for (int r = 0; r < R; r++) {
    const int i = wbeg + r * 32 + lane;
    if (i < end) { const int d = (int)(v[r] >> shift) & mask; stage[digitBase[d] + cnt[warp][d] + rk[r]] = v[r]; }
  }
  __syncthreads();
  const int count = end - beg;
  for (int j = tid; j < count; j += 32 * LZF_WARPS) {
    const u64 x = stage[j];
    const int d = (int)(x >> shift) & mask;
    out[gOff[d] + (u32)j - digitBase[d]] = x;
  }
}
// after the hash passes: sorted[i-1] precedes sorted[i] in (hash, position) order: same hash -> it is the previous occurrence.
// A position with no earlier occurrence of its 4-byte fingerprint inside its hash class can never pass the 4-byte pre-check
// (:389-395, :405-422 need bestLen >= 4) whatever the table holds: flagged LZF_NOCAND here by looking back through the class
// (at most 1024 entries: an unfinished search leaves the flag off, which is always safe).
__global__ void lzf_prev_kernel(LzfBlock* __restrict__ lb, const int* __restrict__ bmap, int nPass, int hashBits) {
  const LzfBlock& L = lb[bmap[blockIdx.y]];
  const int n = L.n;
  const u64* __restrict__ sorted = (nPass & 1) ? L.kb : L.ka;
  const u64 hmask = ((1ull << hashBits) - 1) << 30;
  const int fpShift = 30 + hashBits;
  const int lookback = L.fpLookback;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const u64 v = sorted[i];
    const u32 s = (u32)(v & LZF_POSMASK);
    u32 pv = 0;
    if (i > 0) { const u64 w = sorted[i - 1]; if (((w ^ v) & hmask) == 0) pv = (u32)(w & LZF_POSMASK); }
    bool has = false;
    if (pv != 0) {
      const u64 fp = v >> fpShift;
      int j = i - 1;
      for (int steps = 0; ; steps++, j--) {
        if (steps >= lookback) { has = true; break; }
        const u64 w = sorted[j];
        if (((w ^ v) & hmask) != 0) break;
        if ((w >> fpShift) == fp) { has = true; break; }
        if (j == 0) break;
      }
    }
    L.prev[s] = pv | (has ? 0u : LZF_NOCAND);
    L.hs[i] = s | (pv == 0 ? LZF_RUNSTART : 0u);
    L.rank[s] = (u32)i;
  }
}
// ---- phase 1c: candidate match lengths ---------------------------------------------------------------------------------------------
__global__ void lzf_cand_kernel(const KzgBlock* __restrict__ blocks, LzfBlock* __restrict__ lb, const int* __restrict__ bmap) {
  const int b = bmap[blockIdx.y];
  LzfBlock& L = lb[b];
  const int n = L.n;
  const u8* __restrict__ src = blocks[b].cur;
  // consecutive CTAs take consecutive ranges, so that a CTA's hits all fall into one or two eighths of the block
  const int per = (n + (int)gridDim.x - 1) / (int)gridDim.x;
  const int beg = blockIdx.x * per, end = min(beg + per, n);
  const int eighth = max((n + 7) / 8, 1);
  int hits = 0, hits2 = 0;
  const int e0 = beg / eighth;
  for (int p = beg + threadIdx.x; p < end; p += blockDim.x) {
    const u32 pvRaw = L.prev[p];
    const int ref = (int)(pvRaw & ~LZF_NOCAND);
    const int minRef = max(p - L.maxDist, 0);
    int len = 0;
    if (!(pvRaw & LZF_NOCAND) && ref > minRef && lzf_ld32(src + ref) == lzf_ld32(src + p))
      len = lzf_find_match(src, p, ref, min(L.srcEnd - p, LZ_MAX_MATCH), 256);
    L.len0[p] = (u8)min(len, 255);
    if (len >= 4) { if (p / eighth == e0) hits++; else hits2++; }
  }
  for (int o = 16; o > 0; o >>= 1) { hits += __shfl_xor_sync(0xFFFFFFFFu, hits, o); hits2 += __shfl_xor_sync(0xFFFFFFFFu, hits2, o); }
  if ((threadIdx.x & 31) == 0) {
    if (hits > 0) atomicAdd(&L.estHits[min(e0, 7)], hits);
    if (hits2 > 0) atomicAdd(&L.estHits[min(e0 + 1, 7)], hits2);
  }
}

// ---- phase 2: the walk -----------------------------------------------------------------------------------------------------------------
// (bitmap words are updated with atomics, which act at L2: read them with ld.global.cg so no stale L1 line is used)
__device__ __forceinline__ bool lzf_is_skipped(const u32* sk, int q) { return (__ldcg(sk + (q >> 5)) >> (q & 31)) & 1u; }
// the reference's table content for position x: most recent inserted position with x's hash
__device__ __forceinline__ int lzf_cand(const LzfBlock& L, int x, int lastSkip) {
  int q = (int)(L.prev[x] & ~LZF_NOCAND);
  while (q > 0 && q <= lastSkip && lzf_is_skipped(L.skipped, q)) q = (int)(L.prev[q] & ~LZF_NOCAND);
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
      if (l0 > 0) { const u32 pv = L.prev[q] & ~LZF_NOCAND; lzf_prefetch(src + pv); }
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
  if (L.n <= 0 || !L.needSerial) return;
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
      pv = (int)(L.prev[p] & ~LZF_NOCAND);
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

// ====================================================================================================================================
// Segment-parallel parse.  The walk above is one warp per block: ~10 cycles per dependent instruction and a few hundred
// instructions per match make it latency-bound at a few MB/s per block.  The parse state of the reference is tiny
// (srcIdx, anchor, srcInc, repd[2], repIdx) and greedy LZ parses re-synchronise quickly, so:
//   spec    one warp per 32 KiB segment parses from a fresh state at the segment start and logs its matches.  Table
//           content for positions before the segment comes from the assumed bitmap A, inside the segment from the
//           segment's own bitmap D.
//   stitch  one warp per block walks the segments in order.  Segment 0 is exact.  For the next ones it parses from the
//           true state until it emits a match the segment's log also holds, with the same previous distance and the same
//           skipped positions in between; from there the reference and the speculative parse are in the same state, so
//           the rest of the log is adopted.  The stitcher keeps the exact skipped bitmap Kn of what it produced.
//   check   if Kn == A the assumption every segment made was the truth and the stitched parse is the reference's parse
//           (induction over positions: the first differing decision would need a table lookup that resolves differently,
//           i.e. a consulted position whose bit differs).  Otherwise A := Kn and the block runs another round; the
//           correct prefix grows every round.  Blocks that do not settle fall back to the serial walk above.
//   emit    tokens, distances, lengths and literals are a pure function of the match list: one CTA per block writes them
//           with prefix sums.
#define LZF_WARMUP 2048
struct LzfNoSync { __device__ __forceinline__ bool operator()(int, int, int, const LzfState&) const { return false; } };

template <bool EXTRA, class OnMatch>
__device__ __forceinline__ void lzf_core(const LzfBlock& L, const u8* __restrict__ src, LzfState& st, const int stopAt,
                                         const u32* __restrict__ A, const int aHi, const int ownStart,
                                         u32* D, u32* C, const int dWriteBegin, const int dWriteEnd, uint4* __restrict__ ev, int& nEv, const int evCap,
                                         int& fail, const int lane, OnMatch& onMatch) {
  const int srcEnd = L.srcEnd, maxDist = L.maxDist, minMatch = L.minMatch;
  const int limit = min(srcEnd, stopAt);
  int srcIdx = st.srcIdx, anchor = st.anchor, srcInc = st.srcInc, repd0 = st.repd0, repd1 = st.repd1, repIdx = st.repIdx;
  int lastSkip = st.lastSkip;               // highest own position ever jumped over
  int overLo = st.overLo, overHi = st.overHi;

  // is position q (q > 0) absent from the reference's table?
  auto skippedAt = [&](int q) -> bool {
    if (q >= ownStart) return (q <= lastSkip) && (((__ldcg(D + (q >> 5)) >> (q & 31)) & 1u) != 0);
    return (q <= aHi) && (((__ldg(A + (q >> 5)) >> (q & 31)) & 1u) != 0);
  };
  // First inserted entry of the chain of position x, given its first entry q0 = prev[x].  The chain is contiguous in hs[]
  // (descending from rank[x] - 1), so past the first entry it is read eight entries at a time and their states tested
  // together instead of chasing prev[] one dependent load at a time.  cq: first entry before the segment (its state is an
  // assumption); unsure: an entry above `limitPos` (this batch may still jump over it) stops the walk.
  auto chain = [&](int q0, int x, int limitPos, int& cq, bool& unsure) -> int {
    unsure = false;
    if (q0 <= 0) return 0;
    if (q0 > limitPos) { unsure = true; return 0; }
    if (q0 < ownStart) cq = q0;
    if (!skippedAt(q0)) return q0;
    int i = (int)L.rank[x] - 1;               // index of q0 in hs
    if (L.hs[i] & LZF_RUNSTART) return 0;
    for (;;) {
      // entries i-1 .. i-8 (older ones); stop at the one that opens the hash class
      u32 e[8]; bool sk[8];
      #pragma unroll
      for (int k = 0; k < 8; k++) e[k] = (i - 1 - k >= 0) ? L.hs[i - 1 - k] : LZF_RUNSTART;
      #pragma unroll
      for (int k = 0; k < 8; k++) { const int q = (int)(e[k] & ~LZF_RUNSTART); sk[k] = (q > 0) && skippedAt(q); }
      #pragma unroll
      for (int k = 0; k < 8; k++) {
        const int q = (int)(e[k] & ~LZF_RUNSTART);
        if (q <= 0) return 0;
        if (q < ownStart && cq == 0) cq = q;
        if (!sk[k]) return q;
        if (e[k] & LZF_RUNSTART) return 0;
      }
      i -= 8;
    }
  };
  auto cand = [&](int x) -> int {           // table content for position x: first inserted entry of the prev chain
    const u32 r = L.prev[x];
    if (r & LZF_NOCAND) return 0;             // no earlier occurrence of these 4 bytes: whatever the table holds fails the pre-check
    int cq = 0; bool unsure;
    const int q = chain((int)r, x, 0x7FFFFFFF, cq, unsure);
    if (cq > 0) atomicOr(&C[cq >> 5], 1u << (cq & 31));        // (all lanes, same address)
    return q;
  };
  auto markSkipped = [&](int lo, int hi) {  // called by one lane: positions lo..hi were jumped over
    hi = min(min(hi, srcEnd), dWriteEnd - 1);
    lo = max(lo, dWriteBegin);
    if (lo > hi) return;
    const int w0 = lo >> 5, w1 = hi >> 5;
    for (int w = w0; w <= w1; w++) {
      u32 mask = 0xFFFFFFFFu;
      if (w == w0) mask &= 0xFFFFFFFFu << (lo & 31);
      if (w == w1) mask &= 0xFFFFFFFFu >> (31 - (hi & 31));
      atomicOr(&D[w], mask);
    }
  };

  while (srcIdx < limit) {
    // ---- evaluate the next 32 visit positions, assuming the ones before each are misses ----
    const u32 stepExtra = (u32)((srcInc + lane) >> 6);
    u32 exIncl = stepExtra;
    if (srcInc + 31 >= 64) {
      for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, exIncl, o); if (lane >= o) exIncl += t; }
    }
    const int p = srcIdx + lane + (int)(exIncl - stepExtra);
    const bool valid = p < limit;
    if (srcInc >= 64) {
      // inside a run of misses the next visits are known in advance (every batch all misses): pull the lines of the next two
      // batches towards the SM while this one is evaluated.  (srcInc' + lane) >> 6 changes at most once inside a batch, so the
      // positions come without a scan.
      int nb = __shfl_sync(0xFFFFFFFFu, p + (int)stepExtra + 1, 31);
      #pragma unroll
      for (int d = 1; d <= 2; d++) {
        const int inc = srcInc + 32 * d;
        const int e0 = inc >> 6, kc = 64 - (inc & 63);
        const int q = nb + lane * (1 + e0) + max(0, lane - kc);
        if (q < limit) {
          lzf_prefetch(src + q + 1); lzf_prefetch(L.len0 + q); lzf_prefetch(L.prev + q);
          if (q + 1 - repd0 > 0) lzf_prefetch(src + q + 1 - repd0);
          if (q + 1 - repd1 > 0) lzf_prefetch(src + q + 1 - repd1);
        }
        nb += 32 * (1 + e0) + max(0, 32 - kc);
      }
    }
    bool hit = false;
    int l0 = 0, pv = 0, repSmall = 0, repRef = 0;
    int cnd = -1;                             // table content for p (first inserted entry of its prev chain); -1: resolve after the commit
    int cq = 0;                               // first chain entry before the segment: the lookup's result hangs on the assumed bitmap from there on
    if (valid) {
      const int p1 = p + 1;
      const int minRef = max(p - maxDist, 0);
      const int maxM = min(srcEnd - p1, LZ_MAX_MATCH);
      const int rA = (lane == 0 && repIdx) ? repd1 : repd0, rB = (lane == 0 && repIdx) ? repd0 : repd1;
      const int refA = p1 - rA, refB = p1 - rB;
      const u64 n8 = lzf_ld64(src + p1);
      const u64 a8 = (refA > minRef) ? lzf_ld64(src + refA) : ~n8;
      const u64 b8 = (refB > minRef) ? lzf_ld64(src + refB) : ~n8;
      l0 = L.len0[p];
      const u32 pvRaw = L.prev[p];
      pv = (int)(pvRaw & ~LZF_NOCAND);
      if (l0 >= minMatch && pv >= 8) lzf_prefetch(src + pv - 8);     // the bytes a backward extension will compare
      // the reference tries repd[repIdx] first and only falls to the other one when the 4-byte pre-check fails (:374-387)
      u64 diff;
      if ((u32)(a8 ^ n8) == 0) { diff = a8 ^ n8; repRef = refA; }
      else if ((u32)(b8 ^ n8) == 0) { diff = b8 ^ n8; repRef = refB; }
      else { diff = 1; repRef = 0; }
      if (repRef > 0) repSmall = (maxM < 8) ? 0 : ((diff == 0) ? 8 : ((__ffsll((long long)diff) - 1) >> 3));
      // every lane walks its own chain past jumped-over entries; an entry this very batch may jump over is left for later
      const bool inBatch = ((srcInc + 31) >> 6) > 0;
      bool unsure = false;                    // flagged: nothing the table can hold passes the 4-byte pre-check
      const int q = (pvRaw & LZF_NOCAND) ? 0 : chain(pv, p, inBatch ? srcIdx : 0x7FFFFFFF, cq, unsure);
      bool tableHit;
      if (unsure) tableHit = true;
      else if (pvRaw & LZF_NOCAND) { cnd = 0; tableHit = false; }
      else {
        cnd = q;
        if (q == pv) tableHit = (l0 >= minMatch);
        else tableHit = (q > minRef) && (lzf_ld32(src + q) == lzf_ld32(src + p));     // the exact length is found at the event
      }
      hit = (repSmall >= minMatch) || tableHit;
    }
    const u32 stopMask = __ballot_sync(0xFFFFFFFFu, hit);
    const u32 validMask = __ballot_sync(0xFFFFFFFFu, valid);
    const int nValid = __popc(validMask);
    const int nMiss = stopMask ? min(__ffs(stopMask) - 1, nValid) : nValid;
    if (cq > 0 && lane <= nMiss + 1) atomicOr(&C[cq >> 5], 1u << (cq & 31));   // lookups that (may) really happen: misses, the event, its lazy step
    // ---- commit the misses: positions jumped over by the acceleration are recorded (they are never inserted) ----
    if (nMiss > 0) {
      const int lastEx = __shfl_sync(0xFFFFFFFFu, (int)stepExtra, nMiss - 1);
      if (lastEx > 0) {
        const bool mine = lane < nMiss && stepExtra > 0;
        if (mine) markSkipped(p + 1, p + (int)stepExtra);
        int mx = mine ? min(p + (int)stepExtra, srcEnd) : -1;
        for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
        if (mx >= dWriteEnd) {          // only the last visit of a segment can jump across its end
          int lo = (mine && min(p + (int)stepExtra, srcEnd) >= dWriteEnd) ? max(p + 1, dWriteEnd) : 0x7FFFFFFF;
          for (int o = 16; o > 0; o >>= 1) lo = min(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, o));
          overLo = (overHi < 0) ? lo : min(overLo, lo); overHi = max(overHi, mx);
        }
        lastSkip = max(lastSkip, min(mx, dWriteEnd - 1));
        __threadfence_block();
      }
      const int lastP = __shfl_sync(0xFFFFFFFFu, p, nMiss - 1);
      srcIdx = lastP + 1 + lastEx;
      srcInc += nMiss;
      repIdx = 0;
      __syncwarp();
    }
    if (!stopMask || nMiss >= nValid) continue;
    if (srcIdx >= limit) break;
    // ---- one iteration of the reference loop at srcIdx (:366-566), exact; lane `f` already holds this position's loads ----
    const int f = nMiss;
    const int evL0 = __shfl_sync(0xFFFFFFFFu, l0, f);
    const int evPv = __shfl_sync(0xFFFFFFFFu, pv, f);
    const int evRepSmall = __shfl_sync(0xFFFFFFFFu, repSmall, f);
    const int evRepRef = __shfl_sync(0xFFFFFFFFu, repRef, f);
    const int evCnd = __shfl_sync(0xFFFFFFFFu, cnd, f);
    // lane f+1 holds position srcIdx + 1 when lane f's step is 1
    const int nextOk = (f + 1 < 32) && (((validMask >> (f + 1)) & 1u) != 0) && (__shfl_sync(0xFFFFFFFFu, (int)stepExtra, f) == 0);
    const int nxL0 = __shfl_sync(0xFFFFFFFFu, l0, (f + 1) & 31);
    const int nxPv = __shfl_sync(0xFFFFFFFFu, pv, (f + 1) & 31);
    const int nxCnd = __shfl_sync(0xFFFFFFFFu, cnd, (f + 1) & 31);
    int bestLen = 0;
    const int srcIdx1 = srcIdx + 1;
    const int minRef = max(srcIdx - maxDist, 0);
    int ref = evRepRef;
    if (ref > 0) bestLen = (evRepSmall < 8) ? evRepSmall : lzf_find_match_warp(src, srcIdx1, ref, min(srcEnd - srcIdx1, LZ_MAX_MATCH), lane);
    if (bestLen < minMatch) {
      // check match at position in hash table (:389-395): table content = first entry of the prev chain that was inserted
      const int ref0 = (evCnd >= 0) ? evCnd : cand(srcIdx);
      const bool firstHop = (ref0 == evPv);
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
          if (lane == 0) markSkipped(srcIdx + 1, srcIdx + ex);
          const int mx = min(srcIdx + ex, srcEnd);
          if (mx >= dWriteEnd) { const int lo = max(srcIdx + 1, dWriteEnd); overLo = (overHi < 0) ? lo : min(overLo, lo); overHi = max(overHi, mx); }
          lastSkip = max(lastSkip, min(mx, dWriteEnd - 1));
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
        if (nextOk && nxCnd >= 0) { ref1 = nxCnd; hop1 = (nxCnd == nxPv); if (hop1) l01 = nxL0; }
        else ref1 = cand(srcIdx1);
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
          const int ref2 = cand(srcIdx2);
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
        const bool wide = (srcIdx >= 8) && (ref >= 8);
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
        for (int q = max(srcIdx + 1, ownStart) + lane; q <= hi; q += 32) atomicAnd(&D[q >> 5], ~(1u << (q & 31)));
        __threadfence_block();
        __syncwarp();
      }
    } else {
      if ((bestLen >= LZ_MAX_MATCH) || (src[srcIdx] != src[ref - 1])) srcIdx++;
      else { bestLen++; ref--; }
    }
    // log the match (:467-538 turn it into a token; lzf_emit_kernel does that from the log)
    srcInc = 0;
    const int dist = srcIdx - ref;
    repd1 = repd0; repd0 = dist; repIdx = 1;
    if (nEv >= evCap) { fail = 1; break; }
    if (lane == 0) ev[nEv] = make_uint4((u32)srcIdx, (u32)bestLen, (u32)dist, 0u);
    nEv++;
    const int start = srcIdx;
    anchor = srcIdx + bestLen;
    srcIdx = anchor;
    st.srcIdx = srcIdx; st.anchor = anchor; st.srcInc = 0; st.repd0 = repd0; st.repd1 = repd1; st.repIdx = 1;
    st.lastSkip = lastSkip; st.overLo = overLo; st.overHi = overHi;
    if (onMatch(start, bestLen, dist, st)) return;
  }
  st.srcIdx = srcIdx; st.anchor = anchor; st.srcInc = srcInc; st.repd0 = repd0; st.repd1 = repd1; st.repIdx = repIdx;
  st.lastSkip = lastSkip; st.overLo = overLo; st.overHi = overHi;
}

// ---- the whole-block walk against a real table (sparse blocks) -------------------------------------------------------------------
// Where nearly every position is jumped over (srcInc >> 6 large for long stretches: noise, PCM) speculative segments never
// meet the true parse (srcInc is part of the state and only a match resets it) and the chains of hs[] consist of
// jumped-over entries, so each lookup walks dozens of them.  Such a block is cheap to parse in order instead: the visits
// are few, 32 of them are evaluated per step, and the table is the reference's own (`hashes`, LZCodec.java:372-376): T[h] =
// last inserted position, kept in global memory (L2 resident), updated for every visit and every match interior.
__device__ __forceinline__ u32 lzf_hash_of(u64 w, bool extra) { return (u32)(((w << 24) * LZ_HASH_SEED) >> (extra ? (64 - 19) : (64 - 16))); }

template <bool EXTRA>
__device__ __forceinline__ void lzf_direct_core(const LzfBlock& L, const u8* __restrict__ src, u32* __restrict__ T, u32* __restrict__ Kn, LzfState& st,
                                                uint4* __restrict__ ev, int& nEv, const int evCap, int& fail, const int lane) {
  const int srcEnd = L.srcEnd, maxDist = L.maxDist, minMatch = L.minMatch;
  int srcIdx = st.srcIdx, anchor = st.anchor, srcInc = st.srcInc, repd0 = st.repd0, repd1 = st.repd1, repIdx = st.repIdx;
  int lastSkip = st.lastSkip;
  const u32 lower = (1u << lane) - 1;
  int spanSeg = srcIdx / L.segLen, spanMatches = 0;
  auto insertRange = [&](int lo, int hi) {            // T[hash(q)] = q for q in [lo, hi), in order (max wins)
    for (int q = lo + lane; q < hi; q += 32) atomicMax(&T[lzf_hash_of(lzf_ld64(src + q), EXTRA)], (u32)q);
    __threadfence_block(); __syncwarp();
  };
  auto markSkipped = [&](int lo, int hi) {            // one lane: positions lo..hi were jumped over (the bitmap later segments and the check read)
    hi = min(hi, srcEnd);
    if (lo > hi) return;
    const int w0 = lo >> 5, w1 = hi >> 5;
    for (int w = w0; w <= w1; w++) {
      u32 mask = 0xFFFFFFFFu;
      if (w == w0) mask &= 0xFFFFFFFFu << (lo & 31);
      if (w == w1) mask &= 0xFFFFFFFFu >> (31 - (hi & 31));
      atomicOr(&Kn[w], mask);
    }
  };
  auto save = [&]() { st.srcIdx = srcIdx; st.anchor = anchor; st.srcInc = srcInc; st.repd0 = repd0; st.repd1 = repd1; st.repIdx = repIdx;
                      st.lastSkip = lastSkip; st.overLo = 0; st.overHi = -1; };
  while (srcIdx < srcEnd) {
    // back to the segment logs once the data turns dense again (a segment's worth of positions with a match per KiB)
    const int sg = srcIdx / L.segLen;
    if (sg != spanSeg) {
      // the segments entered are not adopted this round; should the block run another one they are parsed again from the
      // state recorded here and adopted at once
      if (lane == 0) {
        for (int k = spanSeg + 1; k <= min(sg, L.nSeg - 1); k++) {
          LzfSeg& S = L.seg[k];
          S.trueEntry.srcIdx = srcIdx; S.trueEntry.anchor = anchor; S.trueEntry.srcInc = srcInc; S.trueEntry.repd0 = repd0; S.trueEntry.repd1 = repd1;
          S.trueEntry.repIdx = repIdx; S.trueEntry.lastSkip = lastSkip; S.trueEntry.overLo = 0; S.trueEntry.overHi = -1;
          S.haveTrue = 1; S.adopted = 0;
        }
      }
      if (spanMatches >= (sg - spanSeg) * (L.segLen >> 10)) { save(); return; }
      spanSeg = sg; spanMatches = 0;
    }
    // ---- the next 32 visits, assuming the ones before each are misses ----
    const u32 stepExtra = (u32)((srcInc + lane) >> 6);
    u32 exIncl = stepExtra;
    if (srcInc + 31 >= 64) {
      for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, exIncl, o); if (lane >= o) exIncl += t; }
    }
    const int p = srcIdx + lane + (int)(exIncl - stepExtra);
    const bool valid = p < min(srcEnd, (sg + 1) * L.segLen);      // (a batch ends with its segment: the state recorded above is the stitcher's)
    if (srcInc >= 64) {                              // next two batches of a run of misses: pull their lines in (see lzf_core)
      int nb = __shfl_sync(0xFFFFFFFFu, p + (int)stepExtra + 1, 31);
      #pragma unroll
      for (int d = 1; d <= 2; d++) {
        const int inc = srcInc + 32 * d;
        const int e0 = inc >> 6, kc = 64 - (inc & 63);
        const int q = nb + lane * (1 + e0) + max(0, lane - kc);
        if (q < srcEnd) {
          lzf_prefetch(src + q);
          if (q + 1 - repd0 > 0) lzf_prefetch(src + q + 1 - repd0);
          if (q + 1 - repd1 > 0) lzf_prefetch(src + q + 1 - repd1);
        }
        nb += 32 * (1 + e0) + max(0, 32 - kc);
      }
    }
    bool hit = false;
    int repSmall = 0, repRef = 0, cnd = 0;
    u32 h = 0x80000000u | (u32)lane;                 // (lanes without a position never match anybody)
    if (valid) {
      const int p1 = p + 1;
      const int minRef = max(p - maxDist, 0);
      const int maxM = min(srcEnd - p1, LZ_MAX_MATCH);
      const int rA = (lane == 0 && repIdx) ? repd1 : repd0, rB = (lane == 0 && repIdx) ? repd0 : repd1;
      const int refA = p1 - rA, refB = p1 - rB;
      const u64 w = lzf_ld64(src + p);
      const u64 n8 = lzf_ld64(src + p1);
      const u64 a8 = (refA > minRef) ? lzf_ld64(src + refA) : ~n8;
      const u64 b8 = (refB > minRef) ? lzf_ld64(src + refB) : ~n8;
      h = lzf_hash_of(w, EXTRA);
      u64 diff;
      if ((u32)(a8 ^ n8) == 0) { diff = a8 ^ n8; repRef = refA; }
      else if ((u32)(b8 ^ n8) == 0) { diff = b8 ^ n8; repRef = refB; }
      else { diff = 1; repRef = 0; }
      if (repRef > 0) repSmall = (maxM < 8) ? 0 : ((diff == 0) ? 8 : ((__ffsll((long long)diff) - 1) >> 3));
      cnd = (int)__ldcg(T + h);
    }
    // an earlier visit of this batch with the same hash is what the table holds by the time this one looks
    const u32 peers = __match_any_sync(0xFFFFFFFFu, h) & lower;
    const int peerP = __shfl_sync(0xFFFFFFFFu, p, peers ? (31 - __clz(peers)) : 0);
    if (valid) {
      if (peers) cnd = peerP;
      const int minRef = max(p - maxDist, 0);
      const bool tableHit = (cnd > minRef) && (lzf_ld32(src + cnd) == lzf_ld32(src + p));
      hit = (repSmall >= minMatch) || tableHit;
    }
    const u32 stopMask = __ballot_sync(0xFFFFFFFFu, hit);
    const u32 validMask = __ballot_sync(0xFFFFFFFFu, valid);
    const int nValid = __popc(validMask);
    const int nMiss = stopMask ? min(__ffs(stopMask) - 1, nValid) : nValid;
    // ---- commit the misses: every visit enters the table ----
    if (nMiss > 0) {
      if (lane < nMiss) atomicMax(&T[h], (u32)p);
      const int lastEx = __shfl_sync(0xFFFFFFFFu, (int)stepExtra, nMiss - 1);
      const int lastP = __shfl_sync(0xFFFFFFFFu, p, nMiss - 1);
      if (lastEx > 0) {
        if (lane < nMiss && stepExtra > 0) markSkipped(p + 1, p + (int)stepExtra);
        lastSkip = max(lastSkip, min(lastP + lastEx, srcEnd));
      }
      srcIdx = lastP + 1 + lastEx;
      srcInc += nMiss;
      repIdx = 0;
      __threadfence_block(); __syncwarp();
    }
    if (!stopMask || nMiss >= nValid) continue;
    if (srcIdx >= srcEnd) break;
    // ---- one iteration of the reference loop at srcIdx (:366-566), exact ----
    const int f = nMiss;
    const int evRepSmall = __shfl_sync(0xFFFFFFFFu, repSmall, f);
    const int evRepRef = __shfl_sync(0xFFFFFFFFu, repRef, f);
    const int evCnd = __shfl_sync(0xFFFFFFFFu, cnd, f);
    if (lane == f) atomicMax(&T[h], (u32)p);         // hashes[h0] = srcIdx (:372-373)
    __threadfence_block(); __syncwarp();
    int bestLen = 0;
    const int srcIdx1 = srcIdx + 1;
    const int minRef = max(srcIdx - maxDist, 0);
    int ref = evRepRef;
    int insLo;                                       // first position the match inserts again (:553-565 and :456-458)
    if (ref > 0) bestLen = (evRepSmall < 8) ? evRepSmall : lzf_find_match_warp(src, srcIdx1, ref, min(srcEnd - srcIdx1, LZ_MAX_MATCH), lane);
    if (bestLen < minMatch) {
      ref = evCnd;
      int hl = 0;
      if ((ref > minRef) && (lzf_ld32(src + ref) == lzf_ld32(src + srcIdx))) hl = lzf_find_match_warp(src, srcIdx, ref, min(srcEnd - srcIdx, LZ_MAX_MATCH), lane);
      bestLen = (hl >= minMatch) ? hl : 0;
      if (bestLen < minMatch) {       // no good match
        const int ex = srcInc >> 6;
        if (ex > 0) {
          if (lane == 0) markSkipped(srcIdx + 1, srcIdx + ex);
          lastSkip = max(lastSkip, min(srcIdx + ex, srcEnd));
          __threadfence_block(); __syncwarp();
        }
        srcIdx = srcIdx1 + ex;
        srcInc++;
        repIdx = 0;
        continue;
      }
      if ((ref != srcIdx - repd0) && (ref != srcIdx - repd1)) {
        // check if better match at next position (:405-422): the table already holds srcIdx
        const u32 h1 = lzf_hash_of(lzf_ld64(src + srcIdx1), EXTRA);
        const int ref1 = __shfl_sync(0xFFFFFFFFu, (int)__ldcg(T + h1), 0);
        if (lane == 0) atomicMax(&T[h1], (u32)srcIdx1);
        __threadfence_block(); __syncwarp();
        if ((ref1 > minRef + 1) && (lzf_ld32(src + ref1 + bestLen - 3) == lzf_ld32(src + srcIdx1 + bestLen - 3))) {
          const int bestLen1 = lzf_find_match_warp(src, srcIdx1, ref1, min(srcEnd - srcIdx1, LZ_MAX_MATCH), lane);
          if (bestLen1 >= bestLen) { ref = ref1; bestLen = bestLen1; srcIdx = srcIdx1; }
        }
        if (EXTRA) {
          const int srcIdx2 = srcIdx1 + 1;
          const u32 h2 = lzf_hash_of(lzf_ld64(src + srcIdx2), EXTRA);
          const int ref2 = __shfl_sync(0xFFFFFFFFu, (int)__ldcg(T + h2), 0);
          if (lane == 0) atomicMax(&T[h2], (u32)srcIdx2);
          __threadfence_block(); __syncwarp();
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
        const bool wide = (srcIdx >= 8) && (ref >= 8);
        int e;
        if (wide) {
          const u64 d = lzf_ld64(src + srcIdx - 8) ^ lzf_ld64(src + ref - 8);
          e = (d == 0) ? 8 : (__clzll((long long)d) >> 3);
        } else {
          e = (src[srcIdx - 1] == src[ref - 1]) ? 1 : 0;
        }
        const int lim = wide ? 8 : 1;
        const int take = min(e, room);
        bestLen += take; ref -= take; srcIdx -= take;
        if (take < lim) break;
      }
      if (bestLen > LZ_MAX_MATCH) { ref += (bestLen - LZ_MAX_MATCH); srcIdx += (bestLen - LZ_MAX_MATCH); bestLen = LZ_MAX_MATCH; }
      insLo = srcIdx + 1;
      if (lastSkip > srcIdx && srcIdx < visited) {   // jumped-over positions the match covers are inserted again (:553-565)
        const int hi = min(visited, lastSkip);
        for (int q = srcIdx + 1 + lane; q <= hi; q += 32) atomicAnd(&Kn[q >> 5], ~(1u << (q & 31)));
        __threadfence_block(); __syncwarp();
      }
    } else {
      if ((bestLen >= LZ_MAX_MATCH) || (src[srcIdx] != src[ref - 1])) { srcIdx++; insLo = srcIdx; }   // hashes[h1] = srcIdx (:456-458)
      else { bestLen++; ref--; insLo = srcIdx + 1; }
    }
    srcInc = 0;
    const int dist = srcIdx - ref;
    repd1 = repd0; repd0 = dist; repIdx = 1;
    if (nEv >= evCap) { fail = 1; break; }
    if (lane == 0) ev[nEv] = make_uint4((u32)srcIdx, (u32)bestLen, (u32)dist, 0u);
    nEv++;
    anchor = srcIdx + bestLen;
    insertRange(insLo, anchor);                      // (:553-565)
    srcIdx = anchor;
    spanMatches++;
  }
  save();
}

__device__ __forceinline__ int lzf_seg_end(const LzfBlock& L, int s) { return (s + 1 >= L.nSeg) ? L.srcEnd : (s + 1) * L.segLen; }

// Head of every block, parsed for real before the first round.  A block starts with srcInc climbing until the first match, so
// its first few hundred bytes nearly always hold jumped-over positions; rare 4-grams whose only earlier occurrence sits
// there make later segments depend on them, and with nothing assumed about them the block needed a second round for a
// handful of lookups.  The exact bits of [0, LZF_HEAD) go straight into the assumed bitmap A.
#define LZF_HEAD 2048
template <bool EXTRA>
__global__ void __launch_bounds__(32) lzf_head_kernel(const KzgBlock* __restrict__ blocks, LzfBlock* __restrict__ lb, const int* __restrict__ bmap) {
  const int b = bmap[blockIdx.x], lane = threadIdx.x;
  const LzfBlock L = lb[b];
  if (L.n <= 0 || !L.active || L.nSeg < 2) return;
  const u8* __restrict__ src = blocks[b].cur;
  LzfState st;
  st.srcIdx = 0; st.anchor = 0; st.srcInc = 0; st.repd0 = L.count; st.repd1 = L.count; st.repIdx = 0;
  st.lastSkip = -1; st.overLo = 0; st.overHi = -1;
  int nEv = 0, fail = 0;
  LzfNoSync ns;
  lzf_core<EXTRA>(L, src, st, min(LZF_HEAD, L.segLen), nullptr, -1, 0, L.A, nullptr, 0, 0x7FFFFFFF, L.patchEv, nEv, L.patchCap, fail, lane, ns);
  if (lane == 0) lb[b].aMax = st.lastSkip;
}

// (re)start of a round.  Round 0 clears every per-round bitmap.  Later rounds parse again only the segments the check flagged:
// their own jumped-over bits (D) are cleared, everybody else's D, log and dependency marks (C) stay; Kn is rebuilt by the stitch.
__global__ void lzf_round_init_kernel(LzfBlock* __restrict__ lb, const int* __restrict__ bmap, int first) {
  LzfBlock& L = lb[bmap[blockIdx.y]];
  if (L.n <= 0 || !L.active) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) { L.chkDiff = 0; L.chkMax = -1; L.chkFlag = 0; }
  if (first) for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L.nSeg; i += gridDim.x * blockDim.x) { L.seg[i].haveTrue = 0; L.seg[i].rerun = 1; L.seg[i].adopted = 0; }
  const int nW = (L.n + 31) / 32 + 2;
  const int wps = L.segLen >> 5;                     // bitmap words per segment (segLen is a multiple of 4096)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nW; i += gridDim.x * blockDim.x) {
    L.Kn[i] = 0;
    if (first) { L.D[i] = 0; L.C[i] = 0; }
    else if (L.seg[min(i / wps, L.nSeg - 1)].rerun) L.D[i] = 0;
  }
}

template <bool EXTRA>
// One segment per warp, one warp per CTA (72 registers, 28 CTAs per SM).  Measured alternative (round 2): two segments per
// CTA at <= 48 registers (42+ warps per SM) — every launch got slower by about as much as there were more warps (5-8 ms
// instead of 3.5-5 ms per group), no gain in throughput: the parse waits on L2 / DRAM round trips that more warps only queue up.
#define LZF_SPEC_WARPS 1
__global__ void __launch_bounds__(32 * LZF_SPEC_WARPS) lzf_spec_kernel(const KzgBlock* __restrict__ blocks, LzfBlock* __restrict__ lb, const int* __restrict__ bmap) {
  const int b = bmap[blockIdx.y], s = blockIdx.x * LZF_SPEC_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const LzfBlock L = lb[b];
  if (L.n <= 0 || !L.active || s >= L.nSeg) return;
  if (!L.seg[s].rerun) return;                 // its log still stands (no lookup of it resolved differently last round)
  const u8* __restrict__ src = blocks[b].cur;
  const int segStart = s * L.segLen, segEnd = lzf_seg_end(L, s);
  LzfState st;
  st.srcIdx = segStart; st.anchor = segStart; st.srcInc = 0; st.repd0 = L.count; st.repd1 = L.count; st.repIdx = 0;
  st.lastSkip = -1; st.overLo = 0; st.overHi = -1;
  int nEv = 0, fail = 0;
  LzfNoSync ns;
  const int dEnd = (s + 1 >= L.nSeg) ? 0x7FFFFFFF : segEnd;
  uint4* ev = L.specEv + (size_t)s * L.evStride;
  if (s > 0 && L.seg[s].haveTrue) {
    // a previous round's stitch walked up to this segment: start from the state it arrived in, with the positions its last
    // visit before the segment jumped over (they are part of A by now)
    const LzfState te = L.seg[s].trueEntry;
    st.srcIdx = te.srcIdx; st.anchor = te.anchor; st.srcInc = te.srcInc; st.repd0 = te.repd0; st.repd1 = te.repd1; st.repIdx = te.repIdx;
    const int hi = min(st.srcIdx, segEnd);
    if (hi > segStart) {
      bool any = false;
      for (int w = (segStart >> 5) + lane; w <= ((hi - 1) >> 5); w += 32) {
        u32 mask = 0xFFFFFFFFu;
        if (w == ((hi - 1) >> 5) && (hi & 31)) mask &= (1u << (hi & 31)) - 1;
        const u32 v = L.A[w] & mask;
        if (v) { atomicOr(&L.D[w], v); any = true; }
      }
      if (__ballot_sync(0xFFFFFFFFu, any)) st.lastSkip = hi - 1;
      __threadfence_block(); __syncwarp();
    }
  } else if (s > 0) {
    // warm-up: parse the tail of the previous segment so that the state at the segment start is, most of the time,
    // already the reference's (greedy parses re-synchronise within a few matches); nothing of it is kept but the state
    // (where matches are sparse — few positions of the last 2 KiB have a table candidate — two parses meet less often:
    // take a longer run-up there)
    int cnt = 0;
    for (int q = segStart - LZF_WARMUP + lane * 4; q < segStart; q += 128) {
      const u32 w = *reinterpret_cast<const u32*>(L.len0 + q);
      cnt += __popc(__vcmpgeu4(w, 0x04040404u)) >> 3;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    const int warm = (cnt < LZF_WARMUP / 5) ? 4 * LZF_WARMUP : LZF_WARMUP;
    st.srcIdx = st.anchor = max(segStart - warm, 0);
    lzf_core<EXTRA>(L, src, st, segStart, L.A, min(L.aMax, segStart - 1), segStart, L.D, L.C, segStart, dEnd, ev, nEv, L.evStride, fail, lane, ns);
    nEv = 0;
  }
  const LzfState entry = st;
  lzf_core<EXTRA>(L, src, st, segEnd, L.A, min(L.aMax, segStart - 1), segStart, L.D, L.C, segStart, dEnd, ev, nEv, L.evStride, fail, lane, ns);
  int nSkip = 0;
  if (st.lastSkip >= segStart) for (int w = (segStart >> 5) + lane; w < (segEnd >> 5); w += 32) nSkip += __popc(__ldcg(L.D + w));
  for (int o = 16; o > 0; o >>= 1) nSkip += __shfl_xor_sync(0xFFFFFFFFu, nSkip, o);
  if (lane == 0) { LzfSeg& S = L.seg[s]; S.entry = entry; S.end = st; S.nEv = nEv; S.fail = fail; S.rerun = 0; S.nSkip = nSkip; }
}

// D == Kn on bit positions [lo, hi)?
__device__ __forceinline__ bool lzf_bits_equal(const u32* D, const u32* Kn, int lo, int hi, int lane) {
  if (hi <= lo) return true;
  const int w0 = lo >> 5, w1 = (hi - 1) >> 5;
  bool diff = false;
  for (int w = w0 + lane; w <= w1; w += 32) {
    u32 mask = 0xFFFFFFFFu;
    if (w == w0) mask &= 0xFFFFFFFFu << (lo & 31);
    if (w == w1 && (hi & 31)) mask &= (1u << (hi & 31)) - 1;
    if ((__ldcg(D + w) ^ __ldcg(Kn + w)) & mask) diff = true;
  }
  return __ballot_sync(0xFFFFFFFFu, diff) == 0;
}
// Kn[lo, hi) := D[lo, hi)
__device__ __forceinline__ void lzf_bits_copy(const u32* D, u32* Kn, int lo, int hi, int lane) {
  if (hi <= lo) return;
  const int w0 = lo >> 5, w1 = (hi - 1) >> 5;
  for (int w = w0 + lane; w <= w1; w += 32) {
    u32 mask = 0xFFFFFFFFu;
    if (w == w0) mask &= 0xFFFFFFFFu << (lo & 31);
    if (w == w1 && (hi & 31)) mask &= (1u << (hi & 31)) - 1;
    const u32 v = (__ldcg(Kn + w) & ~mask) | (__ldcg(D + w) & mask);
    __stcg(Kn + w, v);
  }
  __threadfence_block();
  __syncwarp();
}

struct LzfSync {           // stitcher side of the re-synchronisation test
  const uint4* spec; int nSpec; int j; int entryRepd0; int segStart, segEnd; const u32* D; const u32* Kn; int lane;
  bool synced; bool dead;
  __device__ __forceinline__ bool operator()(int start, int len, int dist, const LzfState& st) {
    if (dead || start + len >= segEnd) return false;
    // advance to the first logged match at or after `start`
    while (j < nSpec) {
      const int k = j + lane;
      const u32 x = (k < nSpec) ? spec[k].x : 0xFFFFFFFFu;
      const u32 m = __ballot_sync(0xFFFFFFFFu, x >= (u32)start);
      if (m == 0) { j += 32; continue; }
      j += __ffs(m) - 1;
      break;
    }
    if (j >= nSpec) { dead = true; return false; }
    const uint4 e = spec[j];
    if ((int)e.x != start || (int)e.y != len || (int)e.z != dist) return false;
    const int rp1 = (j > 0) ? (int)spec[j - 1].z : entryRepd0;
    if (rp1 != st.repd1) return false;
    if (st.overHi >= 0 || !lzf_bits_equal(D, Kn, segStart, start + len, lane)) { dead = true; return false; }
    synced = true;
    return true;
  }
};

template <bool EXTRA>
__global__ void __launch_bounds__(32) lzf_stitch_kernel(const KzgBlock* __restrict__ blocks, LzfBlock* __restrict__ lb, const int* __restrict__ bmap, int dbg) {
  const int b = bmap[blockIdx.x], lane = threadIdx.x;
  const LzfBlock L = lb[b];
  if (L.n <= 0 || !L.active) return;
  const u8* __restrict__ src = blocks[b].cur;
  int nR = 0, nPatch = 0, fail = 0;
  int nSync = 0, nDead = 0, nOver = 0, nEnd = 0, nAtOnce = 0, lostRun = 0, nDirect = 0, nDbg = 0; long long t0 = clock64();
  int nFin = 0;
  auto addRange = [&](const uint4* p, int cnt) { if (cnt > 0) { if (lane == 0) { L.rng[nR].ev = p; L.rng[nR].count = cnt; L.rng[nR].start = nFin; } nR++; nFin += cnt; } };
  auto applyOver = [&](const LzfState& e) {          // positions a segment jumped over beyond its own end
    if (e.overHi >= 0 && lane == 0) for (int q = e.overLo; q <= e.overHi; q++) atomicOr(&L.Kn[q >> 5], 1u << (q & 31));
    __threadfence_block(); __syncwarp();
  };
  // segment 0 started from the reference's initial state: exact as it stands
  LzfState st = L.seg[0].end;
  fail |= L.seg[0].fail;
  addRange(L.specEv, L.seg[0].nEv);
  if (st.lastSkip >= 0) lzf_bits_copy(L.D, L.Kn, 0, lzf_seg_end(L, 0) + ((L.nSeg == 1) ? 2 : 0), lane);
  applyOver(st);
  st.lastSkip = max(st.lastSkip, st.overHi); st.overLo = 0; st.overHi = -1;
  for (int s = 1; s < L.nSeg; s++) {
    const int segStart = s * L.segLen, segEnd = lzf_seg_end(L, s);
    LzfSeg& S = L.seg[s];
    if (lane == 0) { S.trueEntry = st; S.haveTrue = 1; S.adopted = 0; }
    if (st.srcIdx >= segEnd) { nOver++; continue; }    // a match ran over the whole segment
    LzfSync sy;
    sy.spec = L.specEv + (size_t)s * L.evStride; sy.nSpec = S.nEv; sy.j = 0; sy.entryRepd0 = S.entry.repd0;
    sy.segStart = segStart; sy.segEnd = segEnd; sy.D = L.D; sy.Kn = L.Kn; sy.lane = lane; sy.synced = false;
    sy.dead = (S.fail != 0);
    const LzfState& en = S.entry;
    const bool atOnce = !sy.dead && st.srcIdx == en.srcIdx && st.anchor == en.anchor && st.srcInc == en.srcInc && st.repd0 == en.repd0 &&
                        st.repd1 == en.repd1 && st.repIdx == en.repIdx && lzf_bits_equal(L.D, L.Kn, segStart, st.srcIdx, lane);
    if ((dbg & 32) && !atOnce && lane == 0 && nDbg < 6 && s > 0) {
      nDbg++;
      printf("lzf stitch block %d seg %d not at once: true (%d %d %d %d %d %d) entry (%d %d %d %d %d %d) dead %d nEv %d bitsEq %d\n", b, s, st.srcIdx, st.anchor, st.srcInc,
             st.repd0, st.repd1, st.repIdx, en.srcIdx, en.anchor, en.srcInc, en.repd0, en.repd1, en.repIdx, (int)sy.dead, S.nEv, -1);
    }
    if (atOnce) {                                      // the warm-up already was in the reference's state at the segment start
      nAtOnce++;
      sy.synced = true; sy.j = -1;
      if (S.nEv > 0) {                                 // a first match that reaches back over the segment start re-inserts what it covers
        const uint4 e0 = sy.spec[0];
        const int hi = min(segStart, (int)(e0.x + e0.y)) - 1;
        if ((int)e0.x < segStart && st.lastSkip > (int)e0.x) {
          for (int q = (int)e0.x + 1 + lane; q <= hi; q += 32) atomicAnd(&L.Kn[q >> 5], ~(1u << (q & 31)));
          __threadfence_block(); __syncwarp();
        }
      }
    } else {
      const int from = nPatch;
      lzf_core<EXTRA>(L, src, st, segEnd, L.A, -1, 0, L.Kn, nullptr, 0, 0x7FFFFFFF, L.patchEv, nPatch, L.patchCap, fail, lane, sy);
      addRange(L.patchEv + from, nPatch - from);
      if (fail) break;
      if (sy.dead) nDead++; else if (!sy.synced) nEnd++;
    }
    if (!sy.synced && !fail) {
      // Several segments in a row that never met the true parse, deep inside a run of misses: the data is sparse (noise,
      // PCM ...).  Segment logs are useless there (srcInc never agrees) and the hs[] chains consist of jumped-over entries.
      // Continue in order against a real table until the data turns dense again (lzf_direct_core).
      // (the table is rebuilt from everything before srcIdx by this one warp: only worth it near the start of a block)
      if (++lostRun >= 4 && st.srcInc >= 256 && s + 1 < L.nSeg && st.srcIdx < L.srcEnd && st.srcIdx <= (1 << 18)) {
        u32* T = reinterpret_cast<u32*>(L.kb);                  // (the radix buffers are dead by now)
        for (int i = lane; i < (EXTRA ? (1 << 19) : (1 << 16)); i += 32) T[i] = 0;
        __threadfence_block(); __syncwarp();
        // every position below srcIdx that was not jumped over is in the reference's table: rebuild it from the exact bitmap
        for (int q0 = 0; q0 < st.srcIdx; q0 += 128) {           // (a streaming pass: four positions per lane in flight, lines pulled in ahead)
          if (q0 + 4096 + 128 * lane < st.srcIdx) lzf_prefetch(src + q0 + 4096 + 128 * lane);
          u64 w[4]; u32 km[4];
          #pragma unroll
          for (int k = 0; k < 4; k++) {
            const int q = q0 + 32 * k + lane;
            const bool in = q > 0 && q < st.srcIdx;
            km[k] = in ? __ldcg(L.Kn + (q >> 5)) : 0xFFFFFFFFu;
            w[k] = in ? lzf_ld64(src + q) : 0ull;
          }
          #pragma unroll
          for (int k = 0; k < 4; k++) {
            const int q = q0 + 32 * k + lane;
            if (!((km[k] >> (q & 31)) & 1u)) atomicMax(&T[lzf_hash_of(w[k], EXTRA)], (u32)q);
          }
        }
        __threadfence_block(); __syncwarp();
        const int from = nPatch;
        if (lane == 0) for (int k = s + 1; k <= min(st.srcIdx / L.segLen, L.nSeg - 1); k++) { L.seg[k].trueEntry = st; L.seg[k].haveTrue = 1; L.seg[k].adopted = 0; }
        lzf_direct_core<EXTRA>(L, src, T, L.Kn, st, L.patchEv, nPatch, L.patchCap, fail, lane);
        addRange(L.patchEv + from, nPatch - from);
        if (fail) break;
        nDirect++;
        lostRun = 0;
        // the segments walked this way were not adopted; the loop resumes at the segment holding srcIdx
        const int sTo = min(st.srcIdx / L.segLen, L.nSeg);
        s = sTo - 1;
        continue;
      }
    }
    if (sy.synced) {
      nSync++;
      lostRun = 0;
      if (lane == 0) S.adopted = 1;
      addRange(sy.spec + sy.j + 1, S.nEv - sy.j - 1);
      if (S.end.lastSkip >= segStart)                  // (nothing to merge when the segment never jumped)
        lzf_bits_copy(L.D, L.Kn, atOnce ? segStart : st.anchor, segEnd + ((s + 1 >= L.nSeg) ? 2 : 0), lane);
      const int ls = st.lastSkip;
      st = S.end;
      applyOver(st);
      st.lastSkip = max(max(ls, st.lastSkip), st.overHi); st.overLo = 0; st.overHi = -1;
    }
  }
  if (lane == 0) { lb[b].nRng = nR; lb[b].nFin = nFin; lb[b].giveUpIdx = 0x7FFFFFFF; if (fail) lb[b].needSerial = 1; }
  if ((dbg & 1) && lane == 0) printf("lzf stitch block %d: %d segs, %d synced (%d at once), %d dead, %d unsynced, %d covered, %d own matches, %d in-order stretches, fail %d, %lld cycles\n", b, L.nSeg, nSync, nAtOnce, nDead, nEnd, nOver, nPatch, nDirect, fail, clock64() - t0);
}

// Did any lookup depend on a position whose assumed state (A) is not the produced one (Kn)?
// A lookup that entered the assumed range at chain entry q resolved to the first entry from q on that A does not mark;
// it would have found the same thing under Kn iff that resolution is the same.  If so for every such q of every segment
// whose log the stitched parse adopted, parsing with Kn assumed reproduces this very parse, so the parse is the
// reference's (the stitcher's own matches never look at A).  A mark q that resolves differently can only come from the
// segment holding the next entry y of q's hash class (q was the last entry below that segment's start; a lazy step may
// look one or two positions past a segment's end, hence seg(y) - 1 when y opens its segment): those segments are parsed
// again next round with A := Kn, everybody else keeps log, D bits and marks (stale marks only cost a spurious re-run).
__device__ __forceinline__ int lzf_resolve(const LzfBlock& L, const u32* __restrict__ X, int q) {
  while (q > 0 && ((X[q >> 5] >> (q & 31)) & 1u)) q = (int)(L.prev[q] & ~LZF_NOCAND);
  return q;
}
__global__ void __launch_bounds__(256) lzf_check_marks_kernel(LzfBlock* __restrict__ lb, const int* __restrict__ bmap) {
  LzfBlock& L = lb[bmap[blockIdx.y]];
  if (L.n <= 0 || !L.active || L.needSerial) return;
  const int nW = (L.n + 31) / 32 + 1;
  int differ = 0, mx = -1, flagged = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nW; i += gridDim.x * blockDim.x) {
    const u32 k = L.Kn[i], a = L.A[i];
    if (k != a) differ = 1;
    if (k) mx = i * 32 + 31 - __clz(k);
    u32 c = L.C[i] & (a | k);          // an entry neither bitmap marks resolves to itself under both
    while (c) {
      const int q = i * 32 + __ffs(c) - 1;
      c &= c - 1;
      if (lzf_resolve(L, L.A, q) == lzf_resolve(L, L.Kn, q)) continue;
      const u32 r = L.rank[q] + 1;
      if ((int)r >= L.n) continue;
      const u32 e = L.hs[r];
      if (e & LZF_RUNSTART) continue;                      // q is the last entry of its class: nobody looked it up
      const int y = (int)e, sq = q / L.segLen;
      const int sy = min(y / L.segLen, L.nSeg - 1);
      if (sy > sq && L.seg[sy].adopted) { L.seg[sy].rerun = 1; flagged = 1; }
      if (y - sy * L.segLen <= 2 && sy - 1 > sq && L.seg[sy - 1].adopted) { L.seg[sy - 1].rerun = 1; flagged = 1; }
    }
  }
  if (__syncthreads_or(differ) && threadIdx.x == 0) L.chkDiff = 1;
  if (__syncthreads_or(flagged) && threadIdx.x == 0) L.chkFlag = 1;
  if (mx >= 0) atomicMax(&L.chkMax, mx);
}
__global__ void lzf_check_final_kernel(LzfBlock* __restrict__ lb, const int* __restrict__ bmap, int nBlocks, int lastRound, int* __restrict__ nActive, int dbg) {
  const int i = blockIdx.x;
  if (i >= nBlocks) return;
  const int b = bmap[i];
  LzfBlock& L = lb[b];
  if (L.n <= 0 || !L.active) return;
  __shared__ int go;
  if (threadIdx.x == 0) {
    if ((dbg & 1) && L.chkDiff) {
      int nr = 0; for (int s = 0; s < L.nSeg; s++) nr += L.seg[s].rerun;
      printf("lzf check block %d: bitmaps differ, %d segments to parse again, Kn max %d\n", b, nr, L.chkMax);
    }
    go = 0;
    if (L.needSerial) { L.active = 0; }
    else if (!L.chkDiff || !L.chkFlag) { L.active = 0; }
    else if (lastRound) { L.active = 0; L.needSerial = 1; }
    else { u32* t = L.A; L.A = L.Kn; L.Kn = t; L.aMax = L.chkMax; atomicAdd(nActive, 1); go = 1; }
    if (L.needSerial) atomicAdd(nActive + 1, 1);
  }
  __syncthreads();
  // a block that runs another round also parses again the segments the stitcher could not adopt: started from the state the
  // stitcher reached them in (srcInc included) they are adopted at once next time instead of being walked by one warp again
  if (go) for (int s = 1 + threadIdx.x; s < L.nSeg; s += blockDim.x) if (!L.seg[s].adopted) L.seg[s].rerun = 1;
}

// ---- phase 4: tokens from the match list (:467-538, :568-596) ----------------------------------------------------------------------
__device__ __forceinline__ int lzf_len_size(int length) { return (length < 254) ? 1 : ((length < 65536 + 254) ? 3 : 4); }
__device__ __forceinline__ void lzf_put_length(u8* p, int length) {      // emitLength (:211-231)
  if (length < 254) { p[0] = (u8)length; return; }
  if (length < 65536 + 254) { length -= 254; p[0] = 254; p[1] = (u8)(length >> 8); p[2] = (u8)length; return; }
  length -= 255; p[0] = 255; p[1] = (u8)(length >> 16); p[2] = (u8)(length >> 8); p[3] = (u8)length;
}
#define LZF_ET 1024
struct LzfTok { int litLen, token, nd, mlSize, mlExt, leSize; };
// token of match e given the match before it (p1: previous end and distance) and the distance two matches back
__device__ __forceinline__ LzfTok lzf_tok(const uint4 e, const uint4 p1, const u32 p2z, const int minMatch) {
  LzfTok k;
  k.nd = 0; k.mlSize = 0; k.mlExt = 0; k.leSize = 0;
  k.litLen = (int)e.x - (int)(p1.x + p1.y);
  const int dist = (int)e.z;
  int th;
  if (dist == (int)p1.z) { k.token = 0x00; th = 3; }
  else if (dist == (int)p2z) { k.token = 0x04; th = 3; }
  else { k.nd = 1 + (dist >= 256 ? 1 : 0) + (dist >= 65536 ? 1 : 0); k.token = k.nd << 3; th = 7; }
  const int mLen = (int)e.y - minMatch;
  if (mLen >= th) { k.token += th; k.mlExt = mLen - th; k.mlSize = lzf_len_size(k.mlExt); } else k.token += mLen;
  if (k.litLen >= 7) { k.token |= (7 << 5); k.leSize = lzf_len_size(k.litLen - 7); } else k.token |= (k.litLen << 5);
  return k;
}
__device__ __forceinline__ void lzf_neighbours(const uint4* __restrict__ fin, int i, int count, uint4& p1, u32& p2z) {
  p1 = (i > 0) ? fin[i - 1] : make_uint4(0, 0, (u32)count, 0);
  p2z = (i > 1) ? fin[i - 2].z : (u32)count;
}
// E1: the match list as one array (ranges of the segment logs and of the stitcher's own log, in order)
__global__ void __launch_bounds__(LZF_ET) lzf_emit_gather_kernel(LzfBlock* __restrict__ lb, const int* __restrict__ bmap) {
  const LzfBlock& L = lb[bmap[blockIdx.y]];
  if (L.n <= 0 || L.needSerial) return;
  const int i = blockIdx.x * LZF_ET + threadIdx.x;
  if (i >= L.nFin) return;
  int lo = 0, hi = L.nRng - 1;                          // last range whose start <= i
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (L.rng[mid].start <= i) lo = mid; else hi = mid - 1; }
  L.fin[i] = L.rng[lo].ev[i - L.rng[lo].start];
}
// E2: bytes every tile of 1024 matches adds to the literal area, the distance bytes and the length bytes
__global__ void __launch_bounds__(LZF_ET) lzf_emit_size_kernel(LzfBlock* __restrict__ lb, const int* __restrict__ bmap) {
  __shared__ u32 sA[32], sB[32];
  LzfBlock& L = lb[bmap[blockIdx.y]];
  if (L.n <= 0 || L.needSerial) return;
  const int base = blockIdx.x * LZF_ET;
  if (base >= L.nFin) return;
  const int i = base + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u32 vA = 0, vB = 0;
  if (i < L.nFin) {
    uint4 p1; u32 p2z;
    lzf_neighbours(L.fin, i, L.count, p1, p2z);
    const LzfTok k = lzf_tok(L.fin[i], p1, p2z, L.minMatch);
    vA = (u32)(k.leSize + k.litLen); vB = ((u32)k.nd << 16) | (u32)k.mlSize;
    if (k.litLen >= (1 << 24)) atomicMin(&L.giveUpIdx, i);
  }
  for (int o = 16; o > 0; o >>= 1) { vA += __shfl_xor_sync(0xFFFFFFFFu, vA, o); vB += __shfl_xor_sync(0xFFFFFFFFu, vB, o); }
  if (lane == 0) { sA[warp] = vA; sB[warp] = vB; }
  __syncthreads();
  if (warp == 0) {
    vA = sA[lane]; vB = sB[lane];
    for (int o = 16; o > 0; o >>= 1) { vA += __shfl_xor_sync(0xFFFFFFFFu, vA, o); vB += __shfl_xor_sync(0xFFFFFFFFu, vB, o); }
    if (lane == 0) { L.tileSum[3 * blockIdx.x] = vA; L.tileSum[3 * blockIdx.x + 1] = vB >> 16; L.tileSum[3 * blockIdx.x + 2] = vB & 0xFFFFu; }
  }
}
// E3 (one CTA per block): tile offsets, the reference's end-of-block decisions (:568-596), header and last literals
__global__ void __launch_bounds__(LZF_ET) lzf_emit_scan_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P, LzfBlock* __restrict__ lb, const int* __restrict__ bmap) {
  __shared__ u32 ws[3][32];
  __shared__ u32 carry[3];
  const int b = bmap[blockIdx.x], tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  LzfBlock& L = lb[b];
  if (L.n <= 0 || L.needSerial) return;
  KzgBlock& B = blocks[b];
  int* res = P.result + 2 * b;
  const u8* __restrict__ src = B.cur;
  u8* __restrict__ dst = B.alt;
  const int count = L.count, m = L.nFin;
  const int nT = (m + LZF_ET - 1) / LZF_ET;
  if (tid == 0) { carry[0] = 13; carry[1] = 0; carry[2] = 0; L.emitGo = 0; }
  __syncthreads();
  for (int base = 0; base < nT; base += LZF_ET) {
    const int t = base + tid;
    u32 v[3], inc[3];
    for (int k = 0; k < 3; k++) { v[k] = (t < nT) ? L.tileSum[3 * t + k] : 0u; inc[k] = v[k]; }
    for (int o = 1; o < 32; o <<= 1)
      for (int k = 0; k < 3; k++) { const u32 x = __shfl_up_sync(0xFFFFFFFFu, inc[k], o); if (lane >= o) inc[k] += x; }
    if (lane == 31) for (int k = 0; k < 3; k++) ws[k][warp] = inc[k];
    __syncthreads();
    if (warp == 0) {
      for (int k = 0; k < 3; k++) {
        const u32 a = ws[k][lane]; u32 ia = a;
        for (int o = 1; o < 32; o <<= 1) { const u32 x = __shfl_up_sync(0xFFFFFFFFu, ia, o); if (lane >= o) ia += x; }
        ws[k][lane] = ia - a;
      }
    }
    __syncthreads();
    u32 tot[3];
    for (int k = 0; k < 3; k++) {
      const u32 off = carry[k] + ws[k][warp] + inc[k] - v[k];
      if (t < nT) L.tileSum[3 * t + k] = off;                     // now the exclusive offset of the tile
      tot[k] = off + v[k];
    }
    __syncthreads();
    if (tid == LZF_ET - 1) for (int k = 0; k < 3; k++) carry[k] = tot[k];
    __syncthreads();
  }
  int dstIdx = (int)carry[0]; const int mIdx = (int)carry[1], mLenIdx = (int)carry[2];
  const bool litOverflow = (long long)carry[0] > (long long)B.cap;
  // (error ordering follows the serial loop: the first offending match decides)
  const int gvIdx = L.giveUpIdx;
  // tkBuf is never grown (:324-333): match number tkCap (0-based) is the first whose token does not fit -> block error
  if (gvIdx != 0x7FFFFFFF && !(m > L.tkCap && L.tkCap < gvIdx)) return;                    // forward returns false (:523-524)
  if (m > L.tkCap || litOverflow || mIdx + 3 > L.mCap || mLenIdx + 4 > L.mlCap) { if (tid == 0) atomicExch(&B.status, -KZG_ERR_PROCESS_BLOCK); return; }
  const int prevEnd = (m > 0) ? (int)(L.fin[m - 1].x + L.fin[m - 1].y) : 0;
  const int litLen = count - prevEnd;
  int tkIdx = m;
  if (dstIdx + litLen + tkIdx + mIdx + mLenIdx >= count) return;                           // forward returns false
  if (m == L.tkCap) { if (tid == 0) atomicExch(&B.status, -KZG_ERR_PROCESS_BLOCK); return; }   // the last token does not fit
  const int litBase = dstIdx;
  int lastTok;
  if (litLen >= 7) { lastTok = 7 << 5; if (tid == 0) lzf_put_length(dst + dstIdx, litLen - 7); dstIdx += lzf_len_size(litLen - 7); }
  else lastTok = litLen << 5;
  (void)litBase;
  for (int i = tid; i < litLen; i += LZF_ET) dst[dstIdx + i] = src[prevEnd + i];
  dstIdx += litLen;
  tkIdx = m + 1;
  if (tid == 0) {
    const u32 a = (u32)dstIdx, t = (u32)tkIdx, mm = (u32)mIdx;
    dst[0] = (u8)a; dst[1] = (u8)(a >> 8); dst[2] = (u8)(a >> 16); dst[3] = (u8)(a >> 24);
    dst[4] = (u8)t; dst[5] = (u8)(t >> 8); dst[6] = (u8)(t >> 16); dst[7] = (u8)(t >> 24);
    dst[8] = (u8)mm; dst[9] = (u8)(mm >> 8); dst[10] = (u8)(mm >> 16); dst[11] = (u8)(mm >> 24);
    dst[12] = (u8)(((L.maxDist == LZ_MAX_DISTANCE1) ? 0 : 1) | (((L.minMatch - 2) & 0x07) << 1));
    dst[dstIdx + m] = (u8)lastTok;
    L.tkBase = dstIdx; L.mBase = dstIdx + tkIdx; L.mlBase = dstIdx + tkIdx + mIdx;
    const int total = dstIdx + tkIdx + mIdx + mLenIdx;
    res[1] = total; res[0] = (total <= count - (count / 100)) ? 1 : 0;
    L.emitGo = 1;
  }
}
// E4: every match writes its token, distance bytes and length bytes where the prefix sums put them; the literal area of the
// tile (length-extension bytes + literal runs, contiguous in dst) is then written output-centred: every thread takes four
// consecutive destination bytes, finds the match they belong to in the tile's shared prefix table and fetches the bytes, so
// stores are aligned words and consecutive lanes read consecutive source bytes inside a run.
__global__ void __launch_bounds__(LZF_ET) lzf_emit_write_kernel(KzgBlock* __restrict__ blocks, LzfBlock* __restrict__ lb, const int* __restrict__ bmap) {
  __shared__ u32 wsA[32], wsB[32];
  __shared__ u32 sOff[LZF_ET + 1];                  // literal-area offset of every match of the tile, relative to the tile's start
  __shared__ int sSrc[LZF_ET];                      // source position of area byte 0 (= literal start - leSize)
  __shared__ u8 sLe[LZF_ET];                        // length-extension bytes that open the area
  const LzfBlock& L = lb[bmap[blockIdx.y]];
  if (L.n <= 0 || L.needSerial || !L.emitGo) return;
  const int base = blockIdx.x * LZF_ET;
  if (base >= L.nFin) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i = base + tid;
  const bool on = i < L.nFin;
  const KzgBlock& B = blocks[bmap[blockIdx.y]];
  const u8* __restrict__ src = B.cur;
  u8* __restrict__ dst = B.alt;
  LzfTok k; k.litLen = 0; k.token = 0; k.nd = 0; k.mlSize = 0; k.mlExt = 0; k.leSize = 0;
  uint4 e = make_uint4(0, 0, 0, 0), p1 = e; u32 p2z = 0;
  if (on) { e = L.fin[i]; lzf_neighbours(L.fin, i, L.count, p1, p2z); k = lzf_tok(e, p1, p2z, L.minMatch); }
  const u32 vA = (u32)(k.leSize + k.litLen), vB = ((u32)k.nd << 16) | (u32)k.mlSize;
  u32 iA = vA, iB = vB;
  for (int o = 1; o < 32; o <<= 1) {
    const u32 tA = __shfl_up_sync(0xFFFFFFFFu, iA, o), tB = __shfl_up_sync(0xFFFFFFFFu, iB, o);
    if (lane >= o) { iA += tA; iB += tB; }
  }
  if (lane == 31) { wsA[warp] = iA; wsB[warp] = iB; }
  __syncthreads();
  if (warp == 0) {
    const u32 a = wsA[lane], bb = wsB[lane];
    u32 ia = a, ib = bb;
    for (int o = 1; o < 32; o <<= 1) {
      const u32 tA = __shfl_up_sync(0xFFFFFFFFu, ia, o), tB = __shfl_up_sync(0xFFFFFFFFu, ib, o);
      if (lane >= o) { ia += tA; ib += tB; }
    }
    wsA[lane] = ia - a; wsB[lane] = ib - bb;
    if (lane == 31) sOff[LZF_ET] = ia;              // literal-area bytes of the whole tile
  }
  __syncthreads();
  const u32 relLit = wsA[warp] + iA - vA;
  const u32 offB = wsB[warp] + iB - vB;
  const u32 offM = L.tileSum[3 * blockIdx.x + 1] + (offB >> 16), offML = L.tileSum[3 * blockIdx.x + 2] + (offB & 0xFFFFu);
  sOff[tid] = relLit; sSrc[tid] = (int)(p1.x + p1.y) - k.leSize; sLe[tid] = (u8)k.leSize;
  if (on) {
    const int dist = (int)e.z;
    dst[L.tkBase + i] = (u8)k.token;
    if (k.nd) { u8* q = dst + L.mBase + offM; int j = 0; if (k.nd == 3) q[j++] = (u8)(dist >> 16); if (k.nd >= 2) q[j++] = (u8)(dist >> 8); q[j] = (u8)dist; }
    if (k.mlSize) lzf_put_length(dst + L.mlBase + offML, k.mlExt);
  }
  __syncthreads();
  const u32 total = sOff[LZF_ET];
  const u32 t0 = L.tileSum[3 * blockIdx.x];         // absolute dst offset of the tile's literal area
  const u32 t1 = t0 + total;
  for (u32 g = (t0 & ~3u) + 4u * tid; g < t1; g += 4u * LZF_ET) {
    const u32 lo = max(g, t0), hi = min(g + 4u, t1);
    // last match whose area starts at or before byte lo (areas of length 0 share their start with the next one)
    const u32 o0 = lo - t0;
    int a = 0, b = LZF_ET - 1;
    while (a < b) { const int mid = (a + b + 1) >> 1; if (sOff[mid] <= o0) a = mid; else b = mid - 1; }
    u32 w = 0;
    for (u32 x = lo; x < hi; x++) {
      const u32 o = x - t0;
      while (sOff[a + 1] <= o) a++;
      const int j = (int)(o - sOff[a]);
      const int le = sLe[a];
      u32 v;
      if (j >= le) v = src[sSrc[a] + j];
      else {                                        // byte j of emitLength(litLen - 7) (:211-231)
        const int litLen = (int)(sOff[a + 1] - sOff[a]) - le;
        int len = litLen - 7;
        if (len < 254) v = (u32)len;
        else if (len < 65536 + 254) { len -= 254; v = (j == 0) ? 254u : ((j == 1) ? (u32)(len >> 8) : (u32)len); }
        else { len -= 255; v = (j == 0) ? 255u : ((j == 1) ? (u32)(len >> 16) : ((j == 2) ? (u32)(len >> 8) : (u32)len)); }
        v &= 0xFFu;
      }
      w |= v << (8 * (x - g));
    }
    if (hi - lo == 4u) *reinterpret_cast<u32*>(dst + g) = w;
    else for (u32 x = lo; x < hi; x++) dst[x] = (u8)(w >> (8 * (x - g)));
  }
}

// ---- host ------------------------------------------------------------------------------------------------------------------------------
static size_t lzf_al(size_t v) { return (v + 255) / 256 * 256; }
struct LzfSizes { size_t key, keyB, sa, prev, len0, skipped, hist, tk, m, ml, spec, patch, seg, rng, total; int segLen, maxSeg, evStride, patchCap; };
static LzfSizes lzf_sizes(i32 maxLen) {
  LzfSizes z;
  const size_t n = (size_t)maxLen + 64;
  const size_t nT = (n + LZF_WT - 1) / LZF_WT;
  z.key = lzf_al(8 * n); z.keyB = lzf_al(std::max(8 * n, (size_t)4 << 19));   // kb doubles as the real hash table of the sparse walk (2^19 entries for LZX)
  z.sa = lzf_al(4 * n); z.prev = lzf_al(4 * n); z.len0 = lzf_al(n); z.skipped = lzf_al(n / 8 + 64);
  z.hist = lzf_al(4 * 256 * nT);
  z.tk = lzf_al(std::max<size_t>(n / 5, 256) + 64); z.m = lzf_al(n + 64); z.ml = lzf_al(n / 2 + 64);
  static const int segMin = getenv("KZG_LZ_SEG") ? std::max(8192, atoi(getenv("KZG_LZ_SEG")) / 4096 * 4096) : 8192;   // developer knob
  z.segLen = std::max(segMin, (int)((n / 512 + 4095) / 4096 * 4096));
  z.maxSeg = (int)((n + z.segLen - 1) / z.segLen);
  z.evStride = z.segLen / 4 + 16;
  z.patchCap = (int)(n / 4 + 64);
  z.spec = lzf_al((size_t)z.maxSeg * z.evStride * sizeof(uint4)); z.patch = lzf_al((size_t)z.patchCap * sizeof(uint4));
  z.seg = lzf_al((size_t)z.maxSeg * sizeof(LzfSeg)); z.rng = lzf_al((size_t)(2 * z.maxSeg + 4) * sizeof(LzfRange));
  z.total = z.key + z.keyB + 2 * z.sa + z.prev + z.len0 + 5 * z.skipped + 2 * z.hist + z.tk + z.m + z.ml + z.spec + z.patch + z.seg + z.rng + 1024;
  return z;
}
void kzg_lzf_scratch(i32 maxLen, size_t* perBlockBytes) { *perBlockBytes = std::max(*perBlockBytes, lzf_sizes(maxLen).total + sizeof(LzfBlock) + 256 + 1024); }

// developer aid (KZG_DEBUG bit 4): per-phase times of one launch
struct LzfTimer {
  cudaStream_t s; bool on; std::vector<cudaEvent_t> ev; std::vector<const char*> name;
  LzfTimer(cudaStream_t st, bool o) : s(st), on(o) { mark("start"); }
  void mark(const char* n) { if (!on) return; cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); ev.push_back(e); name.push_back(n); }
  void report() {
    if (!on) return;
    cudaStreamSynchronize(s);
    for (size_t i = 1; i < ev.size(); i++) { float ms = 0; cudaEventElapsedTime(&ms, ev[i - 1], ev[i]); fprintf(stderr, "  lzf %-10s %8.3f ms\n", name[i], ms); }
    for (auto e : ev) cudaEventDestroy(e);
  }
};

// How repetitive is each eighth of a block?  One CTA hashes the 4-grams of a 4 KiB sample into a 64 Kibit set and counts the
// positions whose 4-gram was already there.  The sparsest eighth decides how early the block's group is dealt (the order
// must be known before anything else runs: every group runs its whole pipeline on a stream of its own).
__global__ void __launch_bounds__(256) lzf_sample_kernel(const KzgBlock* __restrict__ blocks, const LzfBlock* __restrict__ lb, int* __restrict__ key) {
  __shared__ u32 seen[2048];
  __shared__ int dup;
  const int b = blockIdx.y, e = blockIdx.x;
  const LzfBlock& L = lb[b];
  if (L.n <= 0) return;
  const int eighth = blocks[b].curLen / LZF_SAMPLES;          // sample e = the first LZF_SAMPLE bytes of eighth e (the layout a
  const int len = min(eighth, LZF_SAMPLE) - 8;                 // host-buffer encode uploads ahead of the blocks themselves)
  if (len < 60) return;
  const u8* __restrict__ src = blocks[b].cur + (size_t)e * eighth;
  for (int i = threadIdx.x; i < 2048; i += 256) seen[i] = 0;
  if (threadIdx.x == 0) dup = 0;
  __syncthreads();
  int d = 0;
  for (int i = threadIdx.x; i < len; i += 256) {
    const u32 hsh = (lzf_ld32(src + i) * 0x9E3779B1u) >> 16;
    const u32 bit = 1u << (hsh & 31);
    if (atomicOr(&seen[hsh >> 5], bit) & bit) d++;
  }
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xFFFFFFFFu, d, o);
  if ((threadIdx.x & 31) == 0 && d) atomicAdd(&dup, d);
  __syncthreads();
  if (threadIdx.x == 0) atomicMin(&key[b], (int)(((long long)dup << 16) / max(len, 1)));
}

// LZF_NOCAND only pays where lookups walk chains of jumped-over entries, i.e. in sparse blocks; dense blocks skip the search
__global__ void lzf_lookback_kernel(LzfBlock* __restrict__ lb, const int* __restrict__ key, int nBlocks, int threshold) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nBlocks) lb[b].fpLookback = (key[b] < threshold) ? LZF_FP_LOOKBACK : 0;
}

// per-thread pool of side streams for the grouped rounds (created once; the calling thread's codec stream forks into them)
#define LZF_MAXG 64
// (LZF_SAMPLES / LZF_SAMPLE: include/.. kzg_transforms.cuh users upload exactly these ranges ahead of the blocks)
struct LzfStreams {
  cudaStream_t st[LZF_MAXG]; cudaEvent_t fork; int n = 0; int* hCnt = nullptr;
  int init(int g) {
    if (!hCnt) { if (cudaMallocHost(&hCnt, 2 * LZF_MAXG * sizeof(int)) != cudaSuccess) return -1; if (cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) != cudaSuccess) return -1; }
    int lo = 0, hi = 0;                               // (numerically lower = more urgent)
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    while (n < LZF_MAXG && n < g) {                   // the groups dealt first (sparse blocks, blocks that need more rounds) get the urgent streams
      const int levels = lo - hi + 1;
      const int pr = hi + std::min(levels - 1, n * levels / 32);
      if (cudaStreamCreateWithPriority(&st[n], cudaStreamNonBlocking, pr) != cudaSuccess) return -1;
      n++;
    }
    return 0;
  }
};
static LzfStreams& lzf_streams() { static thread_local LzfStreams S; return S; }
// frees the calling thread's side streams (kzg_set_device moving the thread to another GPU)
void kzg_lzf_release() {
  LzfStreams& S = lzf_streams();
  for (int i = 0; i < S.n; i++) { cudaStreamSynchronize(S.st[i]); cudaStreamDestroy(S.st[i]); }
  if (S.hCnt) { cudaFreeHost(S.hCnt); cudaEventDestroy(S.fork); }
  S.n = 0; S.hCnt = nullptr;
}

int kzg_lz_forward2_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, bool extra, i32 maxLen) {
  const LzfSizes z = lzf_sizes(maxLen);
  const size_t nb = (size_t)nBlocks;
  if (nb * (z.total + 256) + 1024 > nb * (size_t)P.scratchStride) { kzg_set_error("lz forward: scratch pool too small"); return -KZG_ERR_CREATE_CODEC; }
  // flat pool: [LzfBlock descriptors][per-block areas]
  LzfBlock* dlb = (LzfBlock*)P.scratch;
  u8* base = P.scratch + lzf_al(nb * sizeof(LzfBlock));
  std::vector<LzfBlock> hl(nBlocks);
  for (int b = 0; b < nBlocks; b++) {
    u8* o = base + (size_t)b * z.total;
    LzfBlock& L = hl[b];
    memset(&L, 0, sizeof(L));
    L.ka = (u64*)o; o += z.key; L.kb = (u64*)o; o += z.keyB; L.hs = (u32*)o; o += z.sa; L.rank = (u32*)o; o += z.sa; L.prev = (u32*)o; o += z.prev;
    L.len0 = o; o += z.len0; L.skipped = (u32*)o; o += z.skipped; L.hist = (u32*)o; o += z.hist; L.offs = (u32*)o; o += z.hist;
    L.tk = o; o += z.tk; L.m = o; o += z.m; L.ml = o; o += z.ml;
    L.mCap = (i32)z.m - 16; L.mlCap = (i32)z.ml - 16;
    L.A = (u32*)o; o += z.skipped; L.Kn = (u32*)o; o += z.skipped; L.D = (u32*)o; o += z.skipped; L.C = (u32*)o; o += z.skipped;
    L.specEv = (uint4*)o; o += z.spec; L.patchEv = (uint4*)o; o += z.patch;
    L.seg = (LzfSeg*)o; o += z.seg; L.rng = (LzfRange*)o; o += z.rng;
    L.evStride = z.evStride; L.patchCap = z.patchCap;
    L.fin = (uint4*)L.ka;                       // (the radix buffers are dead once prev[] and its flags exist)
    L.tileSum = L.hist;
  }
  CUDA_TRY(cudaMemcpyAsync(dlb, hl.data(), sizeof(LzfBlock) * nb, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaStreamSynchronize(s));          // hl is stack-owned
  const int dbg = P.flags >> 12;
  LzfTimer tm(s, (dbg & 16) != 0);
  lzf_setup_kernel<<<(nBlocks + 63) / 64, 64, 0, s>>>(d_blocks, nBlocks, P, dlb, z.segLen, (dbg & 8) ? 1 : 0);
  const int gx = std::max(1, std::min((maxLen + 255) / 256, 8 * KZG_SM_COUNT));
  const int nT = (maxLen + LZF_WT - 1) / LZF_WT;
  const int bits = extra ? 19 : 16;
  int launches = 2;
  int* dCnt = (int*)(base + nb * z.total);                 // [2 * LZF_MAXG] counters, the block order, the sampled keys
  int* dMap = dCnt + 2 * LZF_MAXG;
  int* dKey = dMap + nb;
  // phase 1 of `cnt` blocks (bm = their indices) on stream q: hashes, hash sort, prev / len0, fingerprint sort, flags, block head
  auto sortPhase = [&](cudaStream_t q, const int* bm, int cnt, bool head) {
    const dim3 gridT((nT + LZF_WARPS - 1) / LZF_WARPS, cnt);
    if (extra) lzf_hash_kernel<true><<<dim3(gx, cnt), 256, 0, q>>>(d_blocks, dlb, bm);
    else lzf_hash_kernel<false><<<dim3(gx, cnt), 256, 0, q>>>(d_blocks, dlb, bm);
    int pass = 0;
    for (int shift = 0; shift < bits; shift += 8, pass++) {
      const int mask = (1 << std::min(8, bits - shift)) - 1;
      KZG_PROF("lzf_hist_kernel", q, (lzf_hist_kernel<<<gridT, 32 * LZF_WARPS, 0, q>>>(dlb, bm, 30 + shift, mask, pass)));
      lzf_scan_kernel<<<cnt, 1024, 0, q>>>(dlb, bm);
      KZG_PROF("lzf_scatter_kernel", q, (lzf_scatter_kernel<<<dim3(nT, cnt), 32 * LZF_WARPS, 0, q>>>(dlb, bm, 30 + shift, mask, pass)));
    }
    KZG_PROF("lzf_prev_kernel", q, (lzf_prev_kernel<<<dim3(gx, cnt), 256, 0, q>>>(dlb, bm, pass, bits)));
    KZG_PROF("lzf_cand_kernel", q, (lzf_cand_kernel<<<dim3(gx, cnt), 256, 0, q>>>(d_blocks, dlb, bm)));
    launches += 3 + 3 * pass;
    if (head) {
      if (extra) lzf_head_kernel<true><<<cnt, 32, 0, q>>>(d_blocks, dlb, bm);
      else lzf_head_kernel<false><<<cnt, 32, 0, q>>>(d_blocks, dlb, bm);
      launches++;
    }
  };
  const int evTiles = (maxLen / 4 + 64 + LZF_ET - 1) / LZF_ET;
  auto emit = [&](cudaStream_t q, const int* bm, int cnt) {   // phase 4 of `cnt` blocks
    lzf_emit_gather_kernel<<<dim3(evTiles, cnt), LZF_ET, 0, q>>>(dlb, bm);
    lzf_emit_size_kernel<<<dim3(evTiles, cnt), LZF_ET, 0, q>>>(dlb, bm);
    lzf_emit_scan_kernel<<<cnt, LZF_ET, 0, q>>>(d_blocks, P, dlb, bm);
    lzf_emit_write_kernel<<<dim3(evTiles, cnt), LZF_ET, 0, q>>>(d_blocks, dlb, bm);
    launches += 4;
  };
  bool emitted = false;
  // Every group of blocks runs its whole pipeline on a stream of its own: the sort passes of one group (HBM bound) overlap
  // the segment parses of another (latency / issue bound), and a group's stitch (one warp per block) overlaps everything.
  // Blocks with a sparse stretch go first on the urgent streams: the stitcher walks those stretches itself.
  int hCnt[2] = {0, (dbg & 8) ? nBlocks : 0};
  int rounds = 0;
  std::vector<int> order(nBlocks);
  for (int b = 0; b < nBlocks; b++) order[b] = b;
  if (dbg & 8) {                                          // developer switch: everything by the serial walker
    if (P.lazyHost) CUDA_TRY(cudaMemcpyAsync(P.lazyDev, P.lazyHost, (size_t)P.lazyN, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(dMap, order.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, s));
    sortPhase(s, dMap, nBlocks, false);
    CUDA_TRY(cudaStreamSynchronize(s));                   // order is stack-owned
  } else {
    std::vector<int> key(nBlocks, 0x7FFFFFFF);
    CUDA_TRY(cudaMemcpyAsync(dKey, key.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, s));
    lzf_sample_kernel<<<dim3(LZF_SAMPLES, nBlocks), 256, 0, s>>>(d_blocks, dlb, dKey);
    lzf_lookback_kernel<<<(nBlocks + 63) / 64, 64, 0, s>>>(dlb, dKey, nBlocks, (int)(0.08 * 65536));      // sparsest eighth repeats < 8 % of its sampled 4-grams
    CUDA_TRY(cudaMemcpyAsync(key.data(), dKey, sizeof(int) * nb, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return key[x] < key[y]; });
    if (dbg & 1) for (int b = 0; b < nBlocks; b++) fprintf(stderr, "lzf block %d: sparsest eighth repeats %.2f %% of its sampled 4-grams\n", b, key[b] == 0x7FFFFFFF ? -1.0 : 100.0 * key[b] / 65536.0);
    const int gEnv = getenv("KZG_LZ_GROUPS") ? atoi(getenv("KZG_LZ_GROUPS")) : 0;   // developer knob (read per call: the tests flip it)
    const int G = std::max(1, std::min(std::min(gEnv > 0 ? gEnv : 32, LZF_MAXG), nBlocks));
    LzfStreams& ST = lzf_streams();
    if (ST.init(G) < 0) return -KZG_ERR_CREATE_CODEC;
    CUDA_TRY(cudaMemcpyAsync(dMap, order.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemsetAsync(dCnt, 0, 2 * LZF_MAXG * sizeof(int), s));
    CUDA_TRY(cudaEventRecord(ST.fork, s));
    // after maxRounds fixed-point rounds a block goes to the exact serial walker; KZG_LZ_MAXROUNDS (1..8) lowers the cap so the
    // tests can drive blocks down that path (tests/test_gpu_fuzz.py)
    const int maxRounds = getenv("KZG_LZ_MAXROUNDS") ? std::max(1, std::min(8, atoi(getenv("KZG_LZ_MAXROUNDS")))) : 8;
    int gBeg[LZF_MAXG + 1];
    for (int g = 0; g <= G; g++) gBeg[g] = (int)((long long)nBlocks * g / G);
    auto enqueue = [&](int g, int round) {
      cudaStream_t q = ST.st[g];
      const int* bm = dMap + gBeg[g];
      const int cnt = gBeg[g + 1] - gBeg[g];
      if (round > 0) cudaMemsetAsync(dCnt + 2 * g, 0, 2 * sizeof(int), q);
      lzf_round_init_kernel<<<dim3(std::min(gx, 64), cnt), 256, 0, q>>>(dlb, bm, round == 0 ? 1 : 0);
      kzg_prof_begin("lzf_spec_kernel", q);
      if (extra) lzf_spec_kernel<true><<<dim3((z.maxSeg + LZF_SPEC_WARPS - 1) / LZF_SPEC_WARPS, cnt), 32 * LZF_SPEC_WARPS, 0, q>>>(d_blocks, dlb, bm);
      else lzf_spec_kernel<false><<<dim3((z.maxSeg + LZF_SPEC_WARPS - 1) / LZF_SPEC_WARPS, cnt), 32 * LZF_SPEC_WARPS, 0, q>>>(d_blocks, dlb, bm);
      kzg_prof_end(q);
      kzg_prof_begin("lzf_stitch_kernel", q);
      if (extra) lzf_stitch_kernel<true><<<cnt, 32, 0, q>>>(d_blocks, dlb, bm, dbg);
      else lzf_stitch_kernel<false><<<cnt, 32, 0, q>>>(d_blocks, dlb, bm, dbg);
      kzg_prof_end(q);
      KZG_PROF("lzf_check_marks_kernel", q, (lzf_check_marks_kernel<<<dim3(16, cnt), 256, 0, q>>>(dlb, bm)));
      lzf_check_final_kernel<<<cnt, 128, 0, q>>>(dlb, bm, cnt, round == maxRounds - 1 ? 1 : 0, dCnt + 2 * g, dbg);
      cudaMemcpyAsync(ST.hCnt + 2 * g, dCnt + 2 * g, 2 * sizeof(int), cudaMemcpyDeviceToHost, q);
      launches += 5;
    };
    int groupRound[LZF_MAXG];
    bool live[LZF_MAXG], emittedGroup[LZF_MAXG];
    for (int g = 0; g < LZF_MAXG; g++) emittedGroup[g] = false;
    for (int g = 0; g < G; g++) {
      CUDA_TRY(cudaStreamWaitEvent(ST.st[g], ST.fork, 0));
      groupRound[g] = 0; live[g] = true;
      if (P.lazyHost) {                                  // this group's blocks are still on the host: upload them on its stream
        for (int i = gBeg[g]; i < gBeg[g + 1]; i++) {
          const i64 off = (i64)order[i] * P.lazyBlock, len = std::min<i64>(P.lazyBlock, P.lazyN - off);
          if (len > 0) CUDA_TRY(cudaMemcpyAsync(P.lazyDev + off, P.lazyHost + off, (size_t)len, cudaMemcpyHostToDevice, ST.st[g]));
        }
      }
      sortPhase(ST.st[g], dMap + gBeg[g], gBeg[g + 1] - gBeg[g], true);
      enqueue(g, 0);
    }
    int nLive = G;
    const auto tHost0 = std::chrono::steady_clock::now();
    while (nLive > 0) {                                    // whichever group has finished its round gets the next one
      bool progressed = false;
      for (int g = 0; g < G; g++) {
        if (!live[g]) continue;
        const cudaError_t qe = cudaStreamQuery(ST.st[g]);
        if (qe == cudaErrorNotReady) continue;
        if (qe != cudaSuccess) { kzg_set_error("lz forward: %s", cudaGetErrorString(qe)); return -KZG_ERR_PROCESS_BLOCK; }
        progressed = true;
        if (dbg & 64) fprintf(stderr, "lzf t=%7.3f ms: group %d (blocks %d..%d of the order, first block %d) finished round %d, %d still active\n",
                              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tHost0).count(), g, gBeg[g], gBeg[g + 1] - 1, order[gBeg[g]], groupRound[g], ST.hCnt[2 * g]);
        rounds = std::max(rounds, groupRound[g] + 1);
        if (ST.hCnt[2 * g] == 0 || groupRound[g] + 1 >= maxRounds) {
          live[g] = false; nLive--; hCnt[1] += ST.hCnt[2 * g + 1];
          if (ST.hCnt[2 * g + 1] == 0) { emit(ST.st[g], dMap + gBeg[g], gBeg[g + 1] - gBeg[g]); emittedGroup[g] = true; }   // the group's tokens, while the others still parse
          continue;
        }
        groupRound[g]++;
        enqueue(g, groupRound[g]);
      }
      // (rounds last milliseconds: between sweeps the thread sleeps instead of spinning through 32 driver calls — with one process per
      //  GPU on an 8-GPU box the spinning ranks slowed each other's launches: LZ forward 32 -> 41 ms at N = 8)
      if (!progressed) std::this_thread::sleep_for(std::chrono::microseconds(25));
    }
    // join: the groups' emit kernels still run; blocks left to the serial walker (never on the test corpora) are emitted after it
    for (int g = 0; g < G; g++) {
      if (emittedGroup[g]) CUDA_TRY(cudaStreamSynchronize(ST.st[g]));
    }
    if (hCnt[1] == 0) emitted = true;
    else {
      if (extra) lzf_walk_kernel<true><<<nBlocks, 64, 0, s>>>(d_blocks, P, dlb);
      else lzf_walk_kernel<false><<<nBlocks, 64, 0, s>>>(d_blocks, P, dlb);
      launches++;
      for (int g = 0; g < G; g++) if (!emittedGroup[g]) emit(s, dMap + gBeg[g], gBeg[g + 1] - gBeg[g]);
      emitted = true;
    }
    tm.mark("pipeline");
  }
  if (dbg & 1) fprintf(stderr, "lzf: %d blocks, %d rounds, %d serial\n", nBlocks, rounds, hCnt[1]);
  if (!emitted && hCnt[1] > 0) {                          // (developer switch)
    if (extra) lzf_walk_kernel<true><<<nBlocks, 64, 0, s>>>(d_blocks, P, dlb);
    else lzf_walk_kernel<false><<<nBlocks, 64, 0, s>>>(d_blocks, P, dlb);
    launches++;
  }
  if (!emitted) { emit(s, dMap, nBlocks); }                // (serial-walker blocks somewhere, or the developer switch: everything at the end)
  tm.mark("emit");
  tm.report();
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(launches);
  return 0;
}
