// lz_inverse.cu — parallel LZ inverse (sm_100a): replaces the serial token loop of
// K/transform/LZCodec.java LZXCodec.inverseV6 (:626-756).
//
// The reference walks tokens one by one and copies bytes as it goes; its output is a pure function of
// the token stream, so the walk is split into data-parallel passes:
//   1. token parse (one warp per block, 32 tokens per step): literal / match lengths, extension bytes,
//      distance bytes, the repeat-offset state (a warp scan over composable "select" maps) and the
//      output offset of every token (prefix sums).  The only serial part, O(tokens/32) steps.
//   2. pointer fill (one thread per 4 output bytes): ptr[pos] = LIT | literal-stream offset for literal
//      bytes, pos - dist for match bytes (the byte the reference would copy from).
//   3. pointer jumping: ptr[pos] = ptr[ptr[pos]] until every byte points at a literal (up to 16 hops per
//      round with path compression; chains only run backwards, so a handful of rounds suffice).
//   4. gather: dst[pos] = src[ptr[pos]].
// Passes 2-4 are HBM-bound streaming / gather passes over 4 bytes of scratch per output byte.
#include "kzg_common.cuh"
#include "kzg_transforms.cuh"
#include <algorithm>

#define LZI_LIT 0x80000000u
#define LZ_MAX_DISTANCE1 ((1 << 16) - 2)
#define LZ_MAX_DISTANCE2 ((1 << 24) - 2)

struct LziTok { u32 outPos, litSrc, litLen, mLen, dist; };
struct LziHdr { i32 nTok, outLen, ok, done[8]; };

__device__ __forceinline__ u32 lzi_scan_u32(u32 v, int lane, u32& total) {
  u32 incl = v;
  for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
  total = __shfl_sync(0xFFFFFFFFu, incl, 31);
  return incl - v;
}

// repeat-offset state maps: a map sends (r0, r1) to (out0, out1); each output is CONST(d) or IN0 or IN1.
// encoding: bit 31..30 = 0 const (value in low bits), 1 = IN0, 2 = IN1
#define LZI_IN0 0x40000000u
#define LZI_IN1 0x80000000u
__device__ __forceinline__ u32 lzi_apply(u32 sel, u32 f0, u32 f1) {      // g.sel evaluated on the outputs (f0, f1) of the earlier map
  return (sel == LZI_IN0) ? f0 : ((sel == LZI_IN1) ? f1 : sel);
}

// ---- pass 1: token parse -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) lzi_tokens_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P, LziTok* __restrict__ toks, i64 tokStride,
                                                       LziHdr* __restrict__ hdrs) {
  const int lane = threadIdx.x, b = blockIdx.x;
  KzgBlock& B = blocks[b];
  int* res = P.result + 2 * b;
  LziHdr& H = hdrs[b];
  if (lane == 0) { res[0] = 0; res[1] = 0; H.nTok = 0; H.outLen = 0; H.ok = 0; for (int i = 0; i < 8; i++) H.done[i] = 0; }
  __syncwarp();
  if (B.status != 0 || !P.enabled[b]) return;
  const int count = B.curLen;
  const u8* __restrict__ src = B.cur;
  const int dstEnd = min(P.dstLimit[b], B.cap);
  if (count < 13) return;
  auto le32 = [&](int o) { return (i32)((u32)src[o] | ((u32)src[o + 1] << 8) | ((u32)src[o + 2] << 16) | ((u32)src[o + 3] << 24)); };
  const i32 tkLen = le32(0), mIdxLen = le32(4), mLenLen = le32(8);
  if ((tkLen < 0) || (mIdxLen < 0) || (mLenLen < 0)) return;
  if ((tkLen < 13) || (tkLen > count) || (mIdxLen > count - tkLen) || (mLenLen > count - tkLen - mIdxLen)) return;
  const int tkBase = tkLen;                 // tokens start where the literal area ends
  int mIdx = tkBase + mIdxLen;              // NB: Java names: tkIdx = A, mIdx = tkIdx + mIdxLen (B), mLenIdx = mIdx + mLenLen (C)
  // (the header's second field is the token byte count, the third the distance byte count)
  const int nTokBytes = mIdxLen;
  const int distBase = tkBase + nTokBytes;
  const int mLenBase = distBase + mLenLen;
  (void)mIdx;
  const int srcEndLit = tkBase - 13;        // `srcIdx >= srcEnd` ends the walk (:673-674)
  const int litEnd = tkBase;
  const int maxDist = ((src[12] & 1) == 0) ? LZ_MAX_DISTANCE1 : LZ_MAX_DISTANCE2;
  const int minMatch = ((src[12] >> 1) & 0x07) + 2;
  LziTok* T = toks + (i64)b * tokStride;
  if ((i64)nTokBytes > tokStride) { if (lane == 0) atomicExch(&B.status, -KZG_ERR_PROCESS_BLOCK); return; }

  u32 litCur = 13, distCur = (u32)distBase, mLenCur = (u32)mLenBase, outCur = 0;
  u32 rep0 = (u32)count, rep1 = (u32)count;
  int nTok = 0;
  bool fail = false, finished = false;
  for (int base = 0; base < nTokBytes && !finished; base += 32) {
    const int t = base + lane;
    const bool on = t < nTokBytes;
    const int token = on ? src[tkBase + t] : 0;
    const bool hasLit = on && token >= 32;
    const int f = token & 0x18;
    const bool isRep = (f == 0);
    // distance bytes
    const u32 nd = (on && !isRep) ? (u32)(f >> 3) : 0u;
    u32 ndTot; const u32 ndOff = lzi_scan_u32(nd, lane, ndTot);
    // match length extensions, resolved in lane order (each needs the cursor left by the previous one)
    const bool mExt = on && (isRep ? ((token & 3) == 3) : ((token & 7) == 7));
    u32 mLen = on ? (u32)(isRep ? (token & 3) : (token & 7)) + (u32)minMatch : 0u;
    u32 em = __ballot_sync(0xFFFFFFFFu, mExt);
    while (em) {
      const int l = __ffs(em) - 1; em &= em - 1;
      u32 sz = 0;
      if (lane == l) {
        u32 c = mLenCur;
        if (c + 4 > (u32)count + 8) { fail = true; }
        else {
          u32 r = src[c];
          if (r < 254) sz = 1;
          else if (r == 254) { r += ((u32)src[c + 1] << 8) + (u32)src[c + 2]; sz = 3; }
          else { r += ((u32)src[c + 1] << 16) + ((u32)src[c + 2] << 8) + (u32)src[c + 3]; sz = 4; }
          mLen += r;
        }
      }
      mLenCur += __shfl_sync(0xFFFFFFFFu, sz, l);
    }
    // literal lengths: tokens with LLL == 7 carry an extension at the head of their literal run
    const bool lExt = hasLit && token >= 0xE0;
    u32 litLen = hasLit ? (u32)(token >> 5) : 0u;           // 7 for extended ones until resolved
    u32 known = lExt ? 0u : litLen;
    u32 kTot; const u32 kOff = lzi_scan_u32(known, lane, kTot);
    u32 extra = 0;                                            // literal-area bytes of resolved extended tokens in lower lanes
    u32 myExtra = 0; u32 extSz = 0;
    u32 el = __ballot_sync(0xFFFFFFFFu, lExt);
    while (el) {
      const int l = __ffs(el) - 1; el &= el - 1;
      u32 add = 0;
      if (lane == l) {
        const u32 c = litCur + kOff + extra;
        if (c + 4 > (u32)count + 8) { fail = true; }
        else {
          u32 r = src[c];
          if (r < 254) extSz = 1;
          else if (r == 254) { r += ((u32)src[c + 1] << 8) + (u32)src[c + 2]; extSz = 3; }
          else { r += ((u32)src[c + 1] << 16) + ((u32)src[c + 2] << 8) + (u32)src[c + 3]; extSz = 4; }
          litLen = 7 + r;
          add = litLen + extSz;
          myExtra = extra;
        }
      }
      const u32 a = __shfl_sync(0xFFFFFFFFu, add, l);
      if (lane > l) extra += a;
    }
    if (!lExt) myExtra = extra;
    // where this token's literal bytes start (after its own extension bytes) and end
    const u32 litSrc = litCur + kOff + myExtra + extSz;
    const u32 litAfter = litSrc + litLen;
    // the walk ends at the first literal-carrying token whose literals reach srcEnd (:673-674)
    const bool isLast = hasLit && ((i32)litAfter >= srcEndLit);
    const u32 lastMask = __ballot_sync(0xFFFFFFFFu, isLast);
    int nValid = on ? 32 : 0;
    nValid = __popc(__ballot_sync(0xFFFFFFFFu, on));
    if (lastMask) { nValid = __ffs(lastMask); finished = true; }
    const bool live = lane < nValid;
    const bool hasMatch = live && !(finished && lane == nValid - 1);
    if (!hasMatch) mLen = 0;
    // bounds of the literal run (:657-661)
    if (live && hasLit && (litAfter > (u32)litEnd)) fail = true;
    // distances: explicit bytes, then the repeat-offset scan
    u32 dExp = 0;
    if (hasMatch && !isRep) {
      const u32 c = distCur + ndOff;
      if (c + nd > (u32)count + 8) fail = true;
      else { dExp = src[c]; if (nd >= 2) dExp = (dExp << 8) | src[c + 1]; if (nd == 3) dExp = (dExp << 8) | src[c + 2]; }
    }
    // map of this token: NEW(d): (d, IN0); REP0: (IN0, IN0); REP1: (IN1, IN0); no match: identity
    u32 m0, m1;
    if (!hasMatch) { m0 = LZI_IN0; m1 = LZI_IN1; }
    else if (!isRep) { m0 = dExp; m1 = LZI_IN0; }
    else if ((token & 0x04) == 0) { m0 = LZI_IN0; m1 = LZI_IN0; }
    else { m0 = LZI_IN1; m1 = LZI_IN0; }
    // inclusive scan of map composition (later o earlier)
    u32 s0 = m0, s1 = m1;
    for (int o = 1; o < 32; o <<= 1) {
      const u32 p0 = __shfl_up_sync(0xFFFFFFFFu, s0, o), p1 = __shfl_up_sync(0xFFFFFFFFu, s1, o);
      if (lane >= o) { const u32 n0 = lzi_apply(s0, p0, p1), n1 = lzi_apply(s1, p0, p1); s0 = n0; s1 = n1; }
    }
    // state after this token = inclusive map applied to the carried state; the distance used is out0
    const u32 a0 = lzi_apply(s0, rep0, rep1), a1 = lzi_apply(s1, rep0, rep1);
    const u32 dist = a0;
    // output offsets
    const u32 span = live ? (litLen + mLen) : 0u;
    u32 spanTot; const u32 spanOff = lzi_scan_u32(span, lane, spanTot);
    const u32 outPos = outCur + spanOff;
    if (live) {
      // sanity checks of the reference (:657-661, 706-711)
      if (hasLit && (litLen > (u32)dstEnd - min(outPos, (u32)dstEnd))) fail = true;
      if (hasMatch) {
        const u32 mStart = outPos + litLen;
        if (dist > mStart || dist == 0 || dist > (u32)maxDist || mStart + mLen > (u32)dstEnd) fail = true;
      }
      LziTok tk; tk.outPos = outPos; tk.litSrc = litSrc; tk.litLen = litLen; tk.mLen = mLen; tk.dist = dist;
      T[nTok + lane] = tk;
    }
    if (__any_sync(0xFFFFFFFFu, fail)) { fail = true; break; }
    // carry
    const int lastLane = nValid - 1;
    rep0 = __shfl_sync(0xFFFFFFFFu, a0, lastLane); rep1 = __shfl_sync(0xFFFFFFFFu, a1, lastLane);
    litCur = __shfl_sync(0xFFFFFFFFu, litAfter, lastLane);
    // cursors of lanes without literals keep the running value: litAfter of such a lane equals the cursor before it
    distCur += __shfl_sync(0xFFFFFFFFu, ndOff + nd, lastLane);
    outCur += __shfl_sync(0xFFFFFFFFu, spanOff + span, lastLane);
    nTok += nValid;
  }
  if (fail || !finished) return;            // inverse returns false (res[0] stays 0)
  if (lane == 0) {
    H.nTok = nTok; H.outLen = (i32)outCur; H.ok = (litCur == (u32)litEnd) ? 1 : 0;      // `return srcIdx == srcEnd + 13`
    res[1] = (int)outCur;                   // res[0] is set by the gather pass
  }
}

// ---- pass 2: pointer fill -----------------------------------------------------------------------------------------------
#define LZI_TILE 1024
__global__ void __launch_bounds__(256) lzi_fill_kernel(const KzgBlock* __restrict__ blocks, const LziTok* __restrict__ toks, i64 tokStride,
                                                     const LziHdr* __restrict__ hdrs, u32* __restrict__ ptrs, i64 ptrStride) {
  __shared__ u32 sOut[LZI_TILE / 2 + 8];
  __shared__ int sT0, sCnt;
  const int b = blockIdx.y;
  const LziHdr& H = hdrs[b];
  if (H.nTok <= 0) return;
  const int outLen = H.outLen;
  const int tileBeg = blockIdx.x * LZI_TILE;
  if (tileBeg >= outLen) return;
  const int tileEnd = min(tileBeg + LZI_TILE, outLen);
  const LziTok* T = toks + (i64)b * tokStride;
  u32* ptr = ptrs + (i64)b * ptrStride;
  if (threadIdx.x == 0) {
    // last token with outPos <= tileBeg
    int lo = 0, hi = H.nTok - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (T[mid].outPos <= (u32)tileBeg) lo = mid; else hi = mid - 1; }
    sT0 = lo;
  }
  __syncthreads();
  const int t0 = sT0;
  // tokens overlapping the tile: t0 .. first token starting at or after tileEnd (exclusive); at most TILE/2 + 2 (every match is >= 2 bytes)
  for (int i = threadIdx.x; i < LZI_TILE / 2 + 8; i += 256) {
    const int t = t0 + i;
    sOut[i] = (t < H.nTok) ? T[t].outPos : 0xFFFFFFFFu;
  }
  __syncthreads();
  const int pos0 = tileBeg + threadIdx.x * 4;
  if (pos0 >= tileEnd) return;
  // token of pos0: last i with sOut[i] <= pos0
  int lo = 0, hi = LZI_TILE / 2 + 7;
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (sOut[mid] <= (u32)pos0) lo = mid; else hi = mid - 1; }
  int ti = lo;
  LziTok tk = T[t0 + ti];
  u32 v[4];
  #pragma unroll
  for (int k = 0; k < 4; k++) {
    const u32 pos = (u32)pos0 + k;
    if ((int)pos >= tileEnd) { v[k] = LZI_LIT; continue; }
    while (pos >= tk.outPos + tk.litLen + tk.mLen) { ti++; tk = T[t0 + ti]; }
    const u32 off = pos - tk.outPos;
    v[k] = (off < tk.litLen) ? (LZI_LIT | (tk.litSrc + off)) : (pos - tk.dist);
  }
  *reinterpret_cast<uint4*>(ptr + pos0) = make_uint4(v[0], v[1], v[2], v[3]);
}

// ---- pass 3: pointer jumping with path compression ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lzi_jump_kernel(LziHdr* __restrict__ hdrs, u32* __restrict__ ptrs, i64 ptrStride, int round) {
  const int b = blockIdx.y;
  LziHdr& H = hdrs[b];
  if (H.nTok <= 0) return;
  if (round > 0 && H.done[round - 1] == 0) return;            // previous round left nothing unresolved
  const int outLen = H.outLen;
  const int pos0 = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (pos0 >= outLen) return;
  u32* ptr = ptrs + (i64)b * ptrStride;
  uint4 q = *reinterpret_cast<uint4*>(ptr + pos0);
  u32 v[4] = {q.x, q.y, q.z, q.w};
  bool changed = false, open = false;
  #pragma unroll
  for (int k = 0; k < 4; k++) {
    u32 p = v[k];
    if (p & LZI_LIT) continue;
    for (int hop = 0; hop < 16 && !(p & LZI_LIT); hop++) p = ptr[p];
    if (p != v[k]) { v[k] = p; changed = true; }
    if (!(p & LZI_LIT)) open = true;
  }
  if (changed) *reinterpret_cast<uint4*>(ptr + pos0) = make_uint4(v[0], v[1], v[2], v[3]);
  if (open) H.done[round] = 1;                                // benign race: any writer stores 1
}

// ---- pass 4: gather ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lzi_gather_kernel(KzgBlock* __restrict__ blocks, const LziHdr* __restrict__ hdrs, const u32* __restrict__ ptrs,
                                                        i64 ptrStride, int* __restrict__ result) {
  const int b = blockIdx.y;
  const LziHdr& H = hdrs[b];
  if (H.nTok <= 0) return;
  const int outLen = H.outLen;
  const int pos0 = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (blockIdx.x == 0 && threadIdx.x == 0) result[2 * b] = H.ok;
  if (pos0 >= outLen) return;
  const KzgBlock& B = blocks[b];
  const u8* __restrict__ src = B.cur;
  u8* __restrict__ dst = B.alt;
  const u32* ptr = ptrs + (i64)b * ptrStride;
  const uint4 q = *reinterpret_cast<const uint4*>(ptr + pos0);
  u32 v[4] = {q.x, q.y, q.z, q.w};
  u32 outw = 0;
  #pragma unroll
  for (int k = 0; k < 4; k++) {
    u32 p = v[k];
    while (!(p & LZI_LIT)) p = ptr[p];                        // leftovers of very deep chains
    outw |= (u32)src[p & ~LZI_LIT] << (8 * k);
  }
  if (pos0 + 4 <= outLen) *reinterpret_cast<u32*>(dst + pos0) = outw;
  else for (int k = 0; pos0 + k < outLen; k++) dst[pos0 + k] = (u8)(outw >> (8 * k));
}

// scratch: per block tokens (20 B each, up to maxLen/4 + 1024) + 4 bytes per output byte + header
void kzg_lzi_scratch(i32 maxLen, size_t* perBlockBytes, size_t* aux32) {
  const size_t toks = (size_t)maxLen / 4 + 1024;
  *perBlockBytes = std::max(*perBlockBytes, toks * sizeof(LziTok) + 256 + 512);
  *aux32 = std::max(*aux32, (size_t)maxLen + 64);
}

int kzg_lz_inverse_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, i32 maxLen) {
  // flat scratch pool (nBlocks * scratchStride bytes): [dense headers, 256 B reserved per block][token arrays]; pointers in aux32
  const i64 tokStride = (i64)maxLen / 4 + 1024;
  const size_t need = (size_t)nBlocks * (256 + (size_t)tokStride * sizeof(LziTok));
  if (need > (size_t)nBlocks * (size_t)P.scratchStride) { kzg_set_error("lz inverse: scratch pool too small"); return -KZG_ERR_CREATE_CODEC; }
  LziHdr* hdrs = (LziHdr*)P.scratch;
  LziTok* toks = (LziTok*)(P.scratch + (size_t)nBlocks * 256);
  u32* ptrs = (u32*)P.aux32;
  const i64 ptrStride = P.aux32Stride;
  lzi_tokens_kernel<<<nBlocks, 32, 0, s>>>(d_blocks, P, toks, tokStride, hdrs);
  const int tiles = (maxLen + LZI_TILE - 1) / LZI_TILE;
  lzi_fill_kernel<<<dim3(tiles, nBlocks), 256, 0, s>>>(d_blocks, toks, tokStride, hdrs, ptrs, ptrStride);
  for (int r = 0; r < 6; r++) lzi_jump_kernel<<<dim3(tiles, nBlocks), 256, 0, s>>>(hdrs, ptrs, ptrStride, r);
  lzi_gather_kernel<<<dim3(tiles, nBlocks), 256, 0, s>>>(d_blocks, hdrs, ptrs, ptrStride, P.result);
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(9);
  return 0;
}
