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
// One CTA per block, 1024 tokens per tile.  Everything is a prefix sum over tokens except two cursor chains whose steps
// are data dependent: the extended literal lengths (a record at the head of each literal run of 7+ bytes, so the place
// of a record depends on every length before it) and the extended match lengths (1, 3 or 4 byte records in their own
// stream).  Tile loop 1 gathers, for every extended-literal token, the literal bytes of the ordinary tokens before it;
// then one thread chases the literal records (a short dependent-load chain, one step per record) while a warp decodes
// the match-length records 32 at a time; tile loop 2 derives lengths, cursors, the repeat-offset state (a scan over
// composable maps) and output offsets for all tokens.
#define LZI_TT 1024
#define LZI_WIN 16384
#define LZI_KB 1024
struct LziShared {
  u32 wsA[32], wsB[32];
  u32 wm0[32], wm1[32], ws0[32], ws1[32];
  u32 carryK, carryL, carryM, carryD, carryOut, rep0, rep1, totA, totB;
  int nExtL, nExtM, nExtLValid, nExtMValid, lastIdx, fail;
  u32 winBase; int winJ, chaseDone;
  u32 winK[LZI_KB];
  u8 win[LZI_WIN + 16];
};

// exclusive scan of two values over the CTA (1024 threads); totals left in S.totA / S.totB
__device__ __forceinline__ void lzi_cta_scan2(u32 vA, u32 vB, u32& offA, u32& offB, LziShared& S, int lane, int warp) {
  u32 iA = vA, iB = vB;
  for (int o = 1; o < 32; o <<= 1) {
    const u32 tA = __shfl_up_sync(0xFFFFFFFFu, iA, o), tB = __shfl_up_sync(0xFFFFFFFFu, iB, o);
    if (lane >= o) { iA += tA; iB += tB; }
  }
  __syncthreads();                                   // previous users of wsA/wsB are done
  if (lane == 31) { S.wsA[warp] = iA; S.wsB[warp] = iB; }
  __syncthreads();
  if (warp == 0) {
    const u32 a = S.wsA[lane], bb = S.wsB[lane];
    u32 ia = a, ib = bb;
    for (int o = 1; o < 32; o <<= 1) {
      const u32 tA = __shfl_up_sync(0xFFFFFFFFu, ia, o), tB = __shfl_up_sync(0xFFFFFFFFu, ib, o);
      if (lane >= o) { ia += tA; ib += tB; }
    }
    S.wsA[lane] = ia - a; S.wsB[lane] = ib - bb;
    if (lane == 31) { S.totA = ia; S.totB = ib; }
  }
  __syncthreads();
  offA = S.wsA[warp] + iA - vA; offB = S.wsB[warp] + iB - vB;
}

__global__ void __launch_bounds__(LZI_TT) lzi_tokens_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P, LziTok* __restrict__ toks, i64 tokStride,
                                                           LziHdr* __restrict__ hdrs, u32* __restrict__ extPool) {
  __shared__ LziShared S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x;
  KzgBlock& B = blocks[b];
  int* res = P.result + 2 * b;
  LziHdr& H = hdrs[b];
  if (tid == 0) { res[0] = 0; res[1] = 0; H.nTok = 0; H.outLen = 0; H.ok = 0; for (int i = 0; i < 8; i++) H.done[i] = 0; }
  if (B.status != 0 || !P.enabled[b]) return;
  const int count = B.curLen;
  const u8* __restrict__ src = B.cur;
  const int dstEnd = min(P.dstLimit[b], B.cap);
  if (count < 13) return;
  auto le32 = [&](int o) { return (i32)((u32)src[o] | ((u32)src[o + 1] << 8) | ((u32)src[o + 2] << 16) | ((u32)src[o + 3] << 24)); };
  const i32 tkLen = le32(0), mIdxLen = le32(4), mLenLen = le32(8);
  if ((tkLen < 0) || (mIdxLen < 0) || (mLenLen < 0)) return;
  if ((tkLen < 13) || (tkLen > count) || (mIdxLen > count - tkLen) || (mLenLen > count - tkLen - mIdxLen)) return;
  // (Java names: the first header field is where the tokens start, the second the token byte count, the third the distance byte count)
  const int tkBase = tkLen;
  const int nTokBytes = mIdxLen;
  const int distBase = tkBase + nTokBytes;
  const int mLenBase = distBase + mLenLen;
  const int srcEndLit = tkBase - 13;        // `srcIdx >= srcEnd` ends the walk (:673-674)
  const int litEnd = tkBase;
  const int maxDist = ((src[12] & 1) == 0) ? LZ_MAX_DISTANCE1 : LZ_MAX_DISTANCE2;
  const int minMatch = ((src[12] >> 1) & 0x07) + 2;
  LziTok* T = toks + (i64)b * tokStride;
  if ((i64)nTokBytes > tokStride) { if (tid == 0) atomicExch(&B.status, -KZG_ERR_PROCESS_BLOCK); return; }
  // per-block lists: extK[j] literal bytes of ordinary tokens before extended-literal token j; extE[j] literal-area bytes of
  // the extended tokens before j (extE[nExtL] = all of them); extLS[j] = its length | record size << 28; mxVal[k]
  u32* extK = extPool + (i64)b * 4 * tokStride;
  u32* extE = extK + tokStride;
  u32* extLS = extE + tokStride;
  u32* mxVal = extLS + tokStride;

  if (tid == 0) { S.carryK = 0; S.carryL = 0; S.carryM = 0; S.fail = 0; }
  __syncthreads();
  // ---- tile loop 1: ranks of the extended tokens and the ordinary literal bytes before each extended-literal token ----
  for (int base = 0; base < nTokBytes; base += LZI_TT) {
    const int t = base + tid;
    const bool on = t < nTokBytes;
    const int token = on ? src[tkBase + t] : 0;
    const bool hasLit = on && token >= 32;
    const bool lExt = hasLit && token >= 0xE0;
    const bool isRep = (token & 0x18) == 0;
    const bool mExt = on && (isRep ? ((token & 3) == 3) : ((token & 7) == 7));
    const u32 known = (hasLit && !lExt) ? (u32)(token >> 5) : 0u;
    u32 kOff, cOff;
    lzi_cta_scan2(known, ((u32)lExt << 16) | (u32)mExt, kOff, cOff, S, lane, warp);
    if (lExt) extK[S.carryL + (cOff >> 16)] = S.carryK + kOff;
    __syncthreads();
    if (tid == 0) { S.carryK += S.totA; S.carryL += S.totB >> 16; S.carryM += S.totB & 0xFFFFu; }
    __syncthreads();
  }
  if (tid == 0) { S.nExtL = (int)S.carryL; S.nExtM = (int)S.carryM; S.nExtLValid = 0; S.nExtMValid = 0; }
  __syncthreads();
  // ---- the two cursor chains ----
  // warp 1 decodes the match-length records on its own; the other 31 warps stage windows of the literal area and of
  // extK in shared memory so that thread 0's chase runs at shared-memory latency (named barrier 1, 992 threads)
  if (warp == 1) {
    const int nExtM = S.nExtM;
    u32 c = (u32)mLenBase; int k = 0;
    while (k < nExtM) {
      if (c + 32 + 4 > (u32)count + 8) {                         // tail: one record at a time
        if (c + 4 > (u32)count + 8) break;
        u32 r = src[c], sz;
        if (r < 254) sz = 1;
        else if (r == 254) { r += ((u32)src[c + 1] << 8) + (u32)src[c + 2]; sz = 3; }
        else { r += ((u32)src[c + 1] << 16) + ((u32)src[c + 2] << 8) + (u32)src[c + 3]; sz = 4; }
        if (lane == 0) mxVal[k] = r;
        k++; c += sz;
        continue;
      }
      const u32 v = src[c + lane];
      const u32 big = __ballot_sync(0xFFFFFFFFu, v >= 254);
      const int nSmall = big ? (__ffs(big) - 1) : 32;           // one-byte records before the first long one
      const int take = min(nSmall, nExtM - k);
      if (lane < take) mxVal[k + lane] = v;
      k += take; c += take;
      if (big && k < nExtM && take == nSmall) {
        u32 r = src[c], sz;
        if (r == 254) { r += ((u32)src[c + 1] << 8) + (u32)src[c + 2]; sz = 3; }
        else { r += ((u32)src[c + 1] << 16) + ((u32)src[c + 2] << 8) + (u32)src[c + 3]; sz = 4; }
        if (lane == 0) mxVal[k] = r;
        k++; c += sz;
      }
    }
    if (lane == 0) S.nExtMValid = k;
  } else {
    const int wtid = (warp == 0) ? tid : tid - 32;             // 0..991
    const int nExtL = S.nExtL;
    u32 E = 0; int j = 0;                                       // (thread 0)
    bool stop = (nExtL == 0);
    for (;;) {
      if (tid == 0) {
        if (!stop) {
          const u32 c = 13u + extK[j] + E;
          if (c + 4 > (u32)count + 8) stop = true;              // beyond the block: whatever follows cannot be a live token
          S.winBase = c; S.winJ = j;
        }
        S.chaseDone = stop ? 1 : 0;
      }
      asm volatile("bar.sync 1, 992;" ::: "memory");
      if (S.chaseDone) break;
      const u32 wb = S.winBase; const int j0 = S.winJ;
      for (int i = wtid; i < LZI_WIN + 8; i += 992) S.win[i] = (wb + i < (u32)count + 8) ? src[wb + i] : (u8)0;
      for (int i = wtid; i < LZI_KB; i += 992) S.winK[i] = (j0 + i < nExtL) ? extK[j0 + i] : 0u;
      asm volatile("bar.sync 1, 992;" ::: "memory");
      if (tid == 0) {
        while (j < nExtL && j - j0 < LZI_KB) {
          const u32 c = 13u + S.winK[j - j0] + E;
          if (c + 4 > wb + LZI_WIN + 8) break;                  // next window
          if (c + 4 > (u32)count + 8) { stop = true; break; }
          const u8* w = S.win + (c - wb);
          u32 r = w[0], sz;
          if (r < 254) sz = 1;
          else if (r == 254) { r += ((u32)w[1] << 8) + (u32)w[2]; sz = 3; }
          else { r += ((u32)w[1] << 16) + ((u32)w[2] << 8) + (u32)w[3]; sz = 4; }
          const u32 len = 7 + r;
          extE[j] = E; extLS[j] = len | (sz << 28);
          E += len + sz;
          j++;
        }
        if (j >= nExtL) stop = true;
      }
    }
    if (tid == 0) { extE[j] = E; S.nExtLValid = j; }
  }
  __syncthreads();
  if (tid == 0) { S.carryK = 0; S.carryL = 0; S.carryM = 0; S.carryD = (u32)distBase; S.carryOut = 0; S.rep0 = (u32)count; S.rep1 = (u32)count; }
  __syncthreads();
  const int nExtLValid = S.nExtLValid, nExtMValid = S.nExtMValid;
  // ---- tile loop 2: lengths, cursors, repeat offsets, output offsets ----
  int nTok = 0;
  bool finished = false;
  u32 litCurEnd = 13;
  for (int base = 0; base < nTokBytes && !finished; base += LZI_TT) {
    const int t = base + tid;
    const bool on = t < nTokBytes;
    const int token = on ? src[tkBase + t] : 0;
    const bool hasLit = on && token >= 32;
    const bool lExt = hasLit && token >= 0xE0;
    const int f = token & 0x18;
    const bool isRep = (f == 0);
    const bool mExt = on && (isRep ? ((token & 3) == 3) : ((token & 7) == 7));
    const u32 nd = (on && !isRep) ? (u32)(f >> 3) : 0u;
    const u32 known = (hasLit && !lExt) ? (u32)(token >> 5) : 0u;
    u32 kOff, cOff;
    u32 aOff;
    lzi_cta_scan2(known | (nd << 16), ((u32)lExt << 16) | (u32)mExt, aOff, cOff, S, lane, warp);   // (sums stay below 2^16 per tile)
    kOff = aOff & 0xFFFFu;
    const u32 totA = S.totA, totC = S.totB;
    const u32 jb = S.carryL + (cOff >> 16), kb = S.carryM + (cOff & 0xFFFFu), ndOff = aOff >> 16;
    bool bad = false;
    u32 litLen = known, extSz = 0;
    u32 Eb = extE[min(jb, (u32)nExtLValid)];
    if (lExt) {
      if ((int)jb >= nExtLValid) bad = true;
      else { const u32 ls = extLS[jb]; litLen = ls & 0x0FFFFFFFu; extSz = ls >> 28; }
    } else if ((int)jb > nExtLValid) bad = true;
    const u32 litSrc = 13u + S.carryK + kOff + Eb + extSz;
    const u32 litAfter = litSrc + litLen;
    u32 mLen = on ? (u32)(isRep ? (token & 3) : (token & 7)) + (u32)minMatch : 0u;
    bool mBad = false;
    if (mExt) { if ((int)kb >= nExtMValid) mBad = true; else mLen += mxVal[kb]; }
    // the walk ends at the first literal-carrying token whose literals reach srcEnd (:673-674)
    const bool isLast = hasLit && (bad || ((i32)litAfter >= srcEndLit));
    __syncthreads();
    if (tid == 0) S.lastIdx = 0x7FFFFFFF;
    __syncthreads();
    if (isLast) atomicMin(&S.lastIdx, tid);
    __syncthreads();
    const int lastIdx = S.lastIdx;
    int nValid = min(LZI_TT, nTokBytes - base);
    if (lastIdx != 0x7FFFFFFF) { nValid = lastIdx + 1; finished = true; }
    const bool live = tid < nValid;
    const bool hasMatch = live && !(finished && tid == nValid - 1);
    if (!hasMatch) mLen = 0;
    bool fail = live && (bad || (hasMatch && mBad));
    if (live && hasLit && (litAfter > (u32)litEnd)) fail = true;
    // distances: explicit bytes, then the repeat-offset scan
    u32 dExp = 0;
    if (hasMatch && !isRep) {
      const u32 c = S.carryD + ndOff;
      if (c + nd > (u32)count + 8) fail = true;
      else { dExp = src[c]; if (nd >= 2) dExp = (dExp << 8) | src[c + 1]; if (nd == 3) dExp = (dExp << 8) | src[c + 2]; }
    }
    // map of this token: NEW(d): (d, IN0); REP0: (IN0, IN0); REP1: (IN1, IN0); no match: identity
    u32 m0, m1;
    if (!hasMatch) { m0 = LZI_IN0; m1 = LZI_IN1; }
    else if (!isRep) { m0 = dExp; m1 = LZI_IN0; }
    else if ((token & 0x04) == 0) { m0 = LZI_IN0; m1 = LZI_IN0; }
    else { m0 = LZI_IN1; m1 = LZI_IN0; }
    u32 s0 = m0, s1 = m1;                       // inclusive scan of map composition (later o earlier) inside the warp
    for (int o = 1; o < 32; o <<= 1) {
      const u32 p0 = __shfl_up_sync(0xFFFFFFFFu, s0, o), p1 = __shfl_up_sync(0xFFFFFFFFu, s1, o);
      if (lane >= o) { const u32 n0 = lzi_apply(s0, p0, p1), n1 = lzi_apply(s1, p0, p1); s0 = n0; s1 = n1; }
    }
    if (lane == 31) { S.wm0[warp] = s0; S.wm1[warp] = s1; }
    const u32 span = live ? (litLen + mLen) : 0u;
    u32 spanOff, dummy;
    lzi_cta_scan2(span, 0u, spanOff, dummy, S, lane, warp);     // (its barriers also publish wm0/wm1)
    const u32 totSpan = S.totA;
    if (tid == 0) {                              // state at the start of every warp
      u32 r0 = S.rep0, r1 = S.rep1;
      for (int w = 0; w < 32; w++) {
        S.ws0[w] = r0; S.ws1[w] = r1;
        const u32 n0 = lzi_apply(S.wm0[w], r0, r1), n1 = lzi_apply(S.wm1[w], r0, r1);
        r0 = n0; r1 = n1;
      }
      S.rep0 = r0; S.rep1 = r1;
    }
    __syncthreads();
    const u32 dist = lzi_apply(s0, S.ws0[warp], S.ws1[warp]);
    const u32 outPos = S.carryOut + spanOff;
    if (live) {
      // sanity checks of the reference (:657-661, 706-711)
      if (hasLit && (litLen > (u32)dstEnd - min(outPos, (u32)dstEnd))) fail = true;
      if (hasMatch) {
        const u32 mStart = outPos + litLen;
        if (dist > mStart || dist == 0 || dist > (u32)maxDist || mStart + mLen > (u32)dstEnd) fail = true;
      }
      LziTok tk; tk.outPos = outPos; tk.litSrc = litSrc; tk.litLen = litLen; tk.mLen = mLen; tk.dist = dist;
      T[nTok + tid] = tk;
      if (tid == nValid - 1) { S.totA = litAfter; }             // literal cursor after the last live token (tokens without literals keep it)
    }
    if (fail) S.fail = 1;
    __syncthreads();
    litCurEnd = S.totA;
    if (tid == 0) {
      S.carryK += totA & 0xFFFFu; S.carryD += totA >> 16; S.carryL += totC >> 16; S.carryM += totC & 0xFFFFu; S.carryOut += totSpan;
    }
    nTok += nValid;
    __syncthreads();
    if (S.fail) return;                         // inverse returns false (res[0] stays 0)
  }
  if (!finished) return;
  if (tid == 0) {
    H.nTok = nTok; H.outLen = (i32)S.carryOut; H.ok = (litCurEnd == (u32)litEnd) ? 1 : 0;      // `return srcIdx == srcEnd + 13`
    res[1] = (int)S.carryOut;                   // res[0] is set by the gather pass
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
  *perBlockBytes = std::max(*perBlockBytes, toks * (sizeof(LziTok) + 16) + 256 + 512);
  *aux32 = std::max(*aux32, (size_t)maxLen + 64);
}

int kzg_lz_inverse_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, i32 maxLen) {
  // flat scratch pool (nBlocks * scratchStride bytes): [dense headers, 256 B reserved per block][token arrays]; pointers in aux32
  const i64 tokStride = (i64)maxLen / 4 + 1024;
  const size_t need = (size_t)nBlocks * (256 + (size_t)tokStride * (sizeof(LziTok) + 16));
  if (need > (size_t)nBlocks * (size_t)P.scratchStride) { kzg_set_error("lz inverse: scratch pool too small"); return -KZG_ERR_CREATE_CODEC; }
  LziHdr* hdrs = (LziHdr*)P.scratch;
  LziTok* toks = (LziTok*)(P.scratch + (size_t)nBlocks * 256);
  u32* extPool = (u32*)(toks + (size_t)nBlocks * tokStride);
  u32* ptrs = (u32*)P.aux32;
  const i64 ptrStride = P.aux32Stride;
  lzi_tokens_kernel<<<nBlocks, LZI_TT, 0, s>>>(d_blocks, P, toks, tokStride, hdrs, extPool);
  const int tiles = (maxLen + LZI_TILE - 1) / LZI_TILE;
  lzi_fill_kernel<<<dim3(tiles, nBlocks), 256, 0, s>>>(d_blocks, toks, tokStride, hdrs, ptrs, ptrStride);
  for (int r = 0; r < 6; r++) lzi_jump_kernel<<<dim3(tiles, nBlocks), 256, 0, s>>>(hdrs, ptrs, ptrStride, r);
  lzi_gather_kernel<<<dim3(tiles, nBlocks), 256, 0, s>>>(d_blocks, hdrs, ptrs, ptrStride, P.result);
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(9);
  return 0;
}
