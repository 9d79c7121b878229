// lz_inverse.cu — parallel LZ inverse (sm_100a): replaces the serial token loop of
// K/transform/LZCodec.java LZXCodec.inverseV6 (:626-756).
//
// The reference walks tokens one by one and copies bytes as it goes; its output is a pure function of
// the token stream, so the walk is split into data-parallel passes:
//   1. token parse (one warp per block, 32 tokens per step): literal / match lengths, extension bytes,
//      distance bytes, the repeat-offset state (a warp scan over composable "select" maps) and the
//      output offset of every token (prefix sums).  The only serial part, O(tokens/32) steps.
//   2. copy resolution (see "pass 2" below): pointers built and resolved per 8 KiB tile in shared memory, literal-resolved
//      bytes written at once; only bytes whose chain leaves the tile go through the global pointer array.
// Pass 2 moves 4 bytes of pointer per output byte once, plus the open tiles' re-reads.
#include "kzg_common.cuh"
#include "kzg_transforms.cuh"
#include <algorithm>

#define LZI_LIT 0x80000000u
#define LZ_MAX_DISTANCE1 ((1 << 16) - 2)
#define LZ_MAX_DISTANCE2 ((1 << 24) - 2)

struct LziTok { u32 outPos, litSrc, litLen, mLen, dist, flags; };
struct LziHdr { i32 nTok, outLen, ok, done[8]; i32 valid, lastTok, fail, nTokRaw, outLenRaw, okRaw, nExtL, nExtM, nExtLValid, nExtMValid; };

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

// block layout shared by the token kernels (header checks of :636-660)
struct LziLayout { int count, tkBase, nTokBytes, distBase, mLenBase, srcEndLit, litEnd, maxDist, minMatch, dstEnd; };
__device__ __forceinline__ bool lzi_layout(const KzgBlock& B, const KzgXfParams& P, int b, LziLayout& Y) {
  if (B.status != 0 || !P.enabled[b]) return false;
  const int count = B.curLen;
  const u8* __restrict__ src = B.cur;
  if (count < 13) return false;
  auto le32 = [&](int o) { return (i32)((u32)src[o] | ((u32)src[o + 1] << 8) | ((u32)src[o + 2] << 16) | ((u32)src[o + 3] << 24)); };
  const i32 tkLen = le32(0), mIdxLen = le32(4), mLenLen = le32(8);
  if ((tkLen < 0) || (mIdxLen < 0) || (mLenLen < 0)) return false;
  if ((tkLen < 13) || (tkLen > count) || (mIdxLen > count - tkLen) || (mLenLen > count - tkLen - mIdxLen)) return false;
  // (Java names: the first header field is where the tokens start, the second the token byte count, the third the distance byte count)
  Y.count = count; Y.tkBase = tkLen; Y.nTokBytes = mIdxLen; Y.distBase = Y.tkBase + Y.nTokBytes; Y.mLenBase = Y.distBase + mLenLen;
  Y.srcEndLit = Y.tkBase - 13;             // `srcIdx >= srcEnd` ends the walk (:673-674)
  Y.litEnd = Y.tkBase;
  Y.maxDist = ((src[12] & 1) == 0) ? LZ_MAX_DISTANCE1 : LZ_MAX_DISTANCE2;
  Y.minMatch = ((src[12] >> 1) & 0x07) + 2;
  Y.dstEnd = min(kzg_dst_limit(B, P.dstLimit[b]), B.cap);
  return true;
}
// per-block scratch of the token passes
struct LziArrays { LziTok* T; u32 *extK, *extE, *extLS, *mxVal; uint2* tileSum; uint4* tileOff; uint4* tileSpan; u32* tileOut; uint2* tileRep; };
__device__ __forceinline__ LziArrays lzi_arrays(LziTok* toks, u32* extPool, u32* tilePool, i64 tokStride, i64 tileStride, int b) {
  LziArrays A;
  A.T = toks + (i64)b * tokStride;
  A.extK = extPool + (i64)b * 4 * tokStride; A.extE = A.extK + tokStride; A.extLS = A.extE + tokStride; A.mxVal = A.extLS + tokStride;
  u32* t = tilePool + (i64)b * 16 * tileStride;
  A.tileSum = (uint2*)t; A.tileOff = (uint4*)(t + 2 * tileStride); A.tileSpan = (uint4*)(t + 6 * tileStride);
  A.tileOut = t + 10 * tileStride; A.tileRep = (uint2*)(t + 12 * tileStride);
  return A;
}
struct LziFlags { bool on, hasLit, lExt, isRep, mExt; u32 known, nd; };
__device__ __forceinline__ LziFlags lzi_flags(int token, bool on) {
  LziFlags F;
  F.on = on;
  F.hasLit = on && token >= 32;
  F.lExt = F.hasLit && token >= 0xE0;
  const int f = token & 0x18;
  F.isRep = (f == 0);
  F.mExt = on && (F.isRep ? ((token & 3) == 3) : ((token & 7) == 7));
  F.nd = (on && !F.isRep) ? (u32)(f >> 3) : 0u;
  F.known = (F.hasLit && !F.lExt) ? (u32)(token >> 5) : 0u;
  return F;
}

// T1: per tile of 1024 tokens: ordinary literal bytes, explicit distance bytes, extended-literal and extended-match tokens
__global__ void __launch_bounds__(LZI_TT) lzi_tok_sums_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P, LziTok* toks, i64 tokStride, LziHdr* __restrict__ hdrs,
                                                             u32* extPool, u32* tilePool, i64 tileStride) {
  __shared__ LziShared S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.y;
  KzgBlock& B = blocks[b];
  LziHdr& H = hdrs[b];
  LziLayout Y;
  bool ok = lzi_layout(B, P, b, Y);
  if (ok && (i64)Y.nTokBytes > tokStride) { if (blockIdx.x == 0 && tid == 0) atomicExch(&B.status, -KZG_ERR_PROCESS_BLOCK); ok = false; }
  if (blockIdx.x == 0 && tid == 0) {
    int* res = P.result + 2 * b;
    res[0] = 0; res[1] = 0; H.nTok = 0; H.outLen = 0; H.ok = 0; for (int i = 0; i < 8; i++) H.done[i] = 0;
    H.valid = ok ? 1 : 0; H.lastTok = 0x7FFFFFFF; H.fail = 0; H.nTokRaw = 0; H.outLenRaw = 0; H.okRaw = 0;
    H.nExtL = 0; H.nExtM = 0; H.nExtLValid = 0; H.nExtMValid = 0;
  }
  if (!ok) return;
  const LziArrays A = lzi_arrays(toks, extPool, tilePool, tokStride, tileStride, b);
  // (the grid is a fixed number of CTAs per block: the token count is only known on the device)
  for (int tile = blockIdx.x; tile * LZI_TT < Y.nTokBytes; tile += gridDim.x) {
    const int t = tile * LZI_TT + tid;
    const bool on = t < Y.nTokBytes;
    const LziFlags F = lzi_flags(on ? B.cur[Y.tkBase + t] : 0, on);
    u32 o0, o1;
    lzi_cta_scan2(F.known | (F.nd << 16), ((u32)F.lExt << 16) | (u32)F.mExt, o0, o1, S, lane, warp);
    if (tid == 0) A.tileSum[tile] = make_uint2(S.totA, S.totB);
    __syncthreads();
  }
}
// T2 (one CTA per block): exclusive scan of the tile sums
__global__ void __launch_bounds__(LZI_TT) lzi_tok_scan1_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P, LziTok* toks, i64 tokStride, LziHdr* __restrict__ hdrs,
                                                              u32* extPool, u32* tilePool, i64 tileStride) {
  __shared__ LziShared S;
  __shared__ u32 carry[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x;
  LziHdr& H = hdrs[b];
  if (!H.valid) return;
  LziLayout Y;
  if (!lzi_layout(blocks[b], P, b, Y)) return;
  const LziArrays A = lzi_arrays(toks, extPool, tilePool, tokStride, tileStride, b);
  const int nTiles = (Y.nTokBytes + LZI_TT - 1) / LZI_TT;
  if (tid < 4) carry[tid] = 0;
  __syncthreads();
  for (int base = 0; base < nTiles; base += LZI_TT) {
    const int t = base + tid;
    const uint2 v = (t < nTiles) ? A.tileSum[t] : make_uint2(0u, 0u);
    u32 oK, oD, oL, oM;
    lzi_cta_scan2(v.x & 0xFFFFu, v.x >> 16, oK, oD, S, lane, warp);
    const u32 tK = S.totA, tD = S.totB;
    lzi_cta_scan2(v.y >> 16, v.y & 0xFFFFu, oL, oM, S, lane, warp);
    const u32 tL = S.totA, tM = S.totB;
    if (t < nTiles) A.tileOff[t] = make_uint4(carry[0] + oK, carry[1] + oD, carry[2] + oL, carry[3] + oM);
    __syncthreads();
    if (tid == 0) { carry[0] += tK; carry[1] += tD; carry[2] += tL; carry[3] += tM; }
    __syncthreads();
  }
  if (tid == 0) { H.nExtL = (int)carry[2]; H.nExtM = (int)carry[3]; }
}
// T3: extK[j] = ordinary literal bytes before extended-literal token j
__global__ void __launch_bounds__(LZI_TT) lzi_tok_extk_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P, LziTok* toks, i64 tokStride, LziHdr* __restrict__ hdrs,
                                                             u32* extPool, u32* tilePool, i64 tileStride) {
  __shared__ LziShared S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.y;
  if (!hdrs[b].valid) return;
  LziLayout Y;
  if (!lzi_layout(blocks[b], P, b, Y)) return;
  const LziArrays A = lzi_arrays(toks, extPool, tilePool, tokStride, tileStride, b);
  for (int tile = blockIdx.x; tile * LZI_TT < Y.nTokBytes; tile += gridDim.x) {
    const int t = tile * LZI_TT + tid;
    const bool on = t < Y.nTokBytes;
    const LziFlags F = lzi_flags(on ? blocks[b].cur[Y.tkBase + t] : 0, on);
    const uint4 off = A.tileOff[tile];
    u32 kOff, cOff;
    lzi_cta_scan2(F.known, (u32)F.lExt, kOff, cOff, S, lane, warp);
    if (F.lExt) A.extK[off.z + cOff] = off.x + kOff;
    __syncthreads();
  }
}
// T4 (one CTA per block): the two cursor chains
#define LZI_GCAP 8190                                   // largest advance a table entry can hold (stored doubled in 16 bits)
#define LZI_GZERO (LZI_WIN + LZI_GCAP + 2)              // zero region behind the window: reach of four unchecked steps
struct LziChaseSmem {
  u32 winBase, winEbase; int winJ, chaseDone;
  u32 winKaddr[LZI_KB];                                 // shared-memory address of G[13 + extK[j] - window base], clamped to the window end
  u32 winKrel[LZI_KB];                                  // the same, unclamped and as a position
  u32 winE[LZI_KB];                                     // 2 * (E_j - E at the window start)
  u16 G[LZI_WIN + LZI_GZERO];
};
extern __shared__ __align__(16) u8 lzi_chase_smem[];
__global__ void __launch_bounds__(LZI_TT) lzi_tok_chase_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P, LziTok* toks, i64 tokStride, LziHdr* __restrict__ hdrs,
                                                              u32* extPool, u32* tilePool, i64 tileStride) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x;
  LziHdr& H = hdrs[b];
  if (!H.valid) return;
  LziLayout Y;
  if (!lzi_layout(blocks[b], P, b, Y)) return;
  const LziArrays A = lzi_arrays(toks, extPool, tilePool, tokStride, tileStride, b);
  const u8* __restrict__ src = blocks[b].cur;
  const int count = Y.count, mLenBase = Y.mLenBase;
  u32* extK = A.extK; u32* extE = A.extE; u32* extLS = A.extLS; u32* mxVal = A.mxVal;
  // warp 1 decodes the match-length records on its own; the other 31 warps stage windows of the literal area and of
  // extK in shared memory so that thread 0's chase runs at shared-memory latency (named barrier 1, 992 threads)
  const bool dbg = ((P.flags >> 12) & 1) != 0;
  const long long tStart = clock64();
  if (warp == 1) {
    const int nExtM = H.nExtM;
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
    if (lane == 0) H.nExtMValid = k;
    if (dbg && lane == 0) printf("lzi chase block %d: %d match-length records, warp 1 took %lld cycles\n", b, nExtM, clock64() - tStart);
  } else {
    // Record j sits at c_j = 13 + extK[j] + E_j with E_{j+1} = E_j + (record bytes + literal bytes it announces): a chain of
    // dependent loads by construction (no earlier byte says where record j is).  The 31 warps take everything else off that
    // chain: for every byte position x of a 16 KiB window they precompute G[x] = 2 * (what a record starting at x adds), and
    // for the window's next 1024 records the shared-memory ADDRESS of G[13 + extK[j] - window base]; thread 0's step is then
    //     address = kaddr[j] + e2;  g2 = G[address];  e2 += g2                      (two additions and one shared load)
    // with e2 = 2 * (E - E at the window start).  No range test sits on the chain: G is followed by a zero region long
    // enough for any four steps, kaddr is clamped to the window end, positions past the block and runs too long for 16 bits
    // hold 0, and "0" (a real record adds at least 8) sends the step to the careful path, once per group of four.
    // The E values left behind are turned into extE / extLS by the helper warps afterwards.
    LziChaseSmem& C = *reinterpret_cast<LziChaseSmem*>(lzi_chase_smem);
    const int wtid = (warp == 0) ? tid : tid - 32;             // 0..991
    const int nExtL = H.nExtL;
    const u32 lim = (u32)count + 8;
    const u32 gAddr = (u32)__cvta_generic_to_shared(&C.G[0]);
    u32 E = 0; int j = 0;                                       // (thread 0)
    bool stop = (nExtL == 0);
    u32 wbOld = 0, ebOld = 0; int pj0 = 0, pj1 = 0;
    int nWin = 0; long long tChase = 0;
    for (int x = wtid; x < LZI_GZERO; x += 992) C.G[LZI_WIN + x] = 0;      // the zero region (never rewritten)
    for (;;) {
      if (tid == 0) {
        if (!stop) {
          const u32 c = 13u + extK[j] + E;
          if (c + 4 > lim) stop = true;                         // beyond the block: whatever follows cannot be a live token
          C.winBase = c; C.winEbase = E;
        }
        C.winJ = j;
        C.chaseDone = stop ? 1 : 0;
      }
      asm volatile("bar.sync 1, 992;" ::: "memory");
      pj1 = C.winJ;
      // records the chase placed in the previous window: their lengths and sizes
      for (int i = pj0 + wtid; i < pj1; i += 992) {
        const u32 e = ebOld + (C.winE[i - pj0] >> 1);
        const u32 c = wbOld + C.winKrel[i - pj0] + (C.winE[i - pj0] >> 1);
        u32 r = src[c], sz;
        if (r < 254) sz = 1;
        else if (r == 254) { r += ((u32)src[c + 1] << 8) + (u32)src[c + 2]; sz = 3; }
        else { r += ((u32)src[c + 1] << 16) + ((u32)src[c + 2] << 8) + (u32)src[c + 3]; sz = 4; }
        extE[i] = e; extLS[i] = (7 + r) | (sz << 28);
      }
      if (C.chaseDone) break;
      const u32 wb = C.winBase, eb = C.winEbase; const int j0 = pj1;
      asm volatile("bar.sync 1, 992;" ::: "memory");           // winKrel / winE of the previous window are consumed
      {
        const u32 k0 = extK[j0];
        for (int i = wtid; i < LZI_KB; i += 992) {
          const u32 kr = (j0 + i < nExtL) ? (extK[j0 + i] - k0) : (u32)LZI_WIN;       // relative to the window base
          C.winKrel[i] = kr;
          C.winKaddr[i] = gAddr + 2u * min(kr, (u32)LZI_WIN);
        }
      }
      for (int x = wtid; x < LZI_WIN; x += 992) {
        const u32 c = wb + (u32)x;
        u32 g = 0;
        if (c + 4 <= lim) {
          u32 r = src[c], sz;
          if (r < 254) sz = 1;
          else if (r == 254) { r += ((u32)src[c + 1] << 8) + (u32)src[c + 2]; sz = 3; }
          else { r += ((u32)src[c + 1] << 16) + ((u32)src[c + 2] << 8) + (u32)src[c + 3]; sz = 4; }
          g = 7u + r + sz;
          if (g > (u32)LZI_GCAP) g = 0;                         // a literal run too long for the table: careful path
        }
        C.G[x] = (u16)(2u * g);
      }
      if (wtid < 128 && wb + LZI_WIN + 128u * wtid < lim) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + wb + LZI_WIN + 128u * wtid));
      asm volatile("bar.sync 1, 992;" ::: "memory");
      if (tid == 0) {
        nWin++; const long long tc = clock64();
        const int qMax = min(nExtL - j0, LZI_KB);
        int q = 0;
        u32 e2 = 0;
        const u32* ka = C.winKaddr;
        for (;;) {
          if ((e2 >> 1) >= (u32)LZI_WIN) break;                 // (a long run of the careful path can leave the window: the unchecked loads below must not)
          while (q + 4 <= qMax) {
            u32 g0, g1, g2, g3;
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(g0) : "r"(ka[q] + e2));
            const u32 e1 = e2 + g0;
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(g1) : "r"(ka[q + 1] + e1));
            const u32 e2b = e1 + g1;
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(g2) : "r"(ka[q + 2] + e2b));
            const u32 e3 = e2b + g2;
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(g3) : "r"(ka[q + 3] + e3));
            const u32 e4 = e3 + g3;
            if (g0 != 0 && g1 != 0 && g2 != 0 && g3 != 0) { C.winE[q] = e2; C.winE[q + 1] = e1; C.winE[q + 2] = e2b; C.winE[q + 3] = e3; e2 = e4; q += 4; continue; }
            if (g0 == 0) break;
            C.winE[q] = e2; e2 = e1; q++;
            if (g1 == 0) break;
            C.winE[q] = e2; e2 = e2b; q++;
            if (g2 == 0) break;
            C.winE[q] = e2; e2 = e3; q++;
            break;
          }
          // one careful step: the group's odd record, the last records of a window, a long run
          if (q >= qMax) break;
          const u32 cRel = C.winKrel[q] + (e2 >> 1);
          if (cRel >= (u32)LZI_WIN) break;                      // next window
          const u32 c = wb + cRel;
          if (c + 4 > lim) { stop = true; break; }
          u32 g = (u32)C.G[cRel] >> 1;
          if (g == 0) {                                         // a literal run beyond the table's range
            u32 r = src[c], sz;
            if (r < 254) sz = 1;
            else if (r == 254) { r += ((u32)src[c + 1] << 8) + (u32)src[c + 2]; sz = 3; }
            else { r += ((u32)src[c + 1] << 16) + ((u32)src[c + 2] << 8) + (u32)src[c + 3]; sz = 4; }
            g = 7u + r + sz;
          }
          C.winE[q] = e2;
          e2 += 2u * g;
          q++;
        }
        E = eb + (e2 >> 1);
        j = j0 + q;
        if (j >= nExtL) stop = true;
        tChase += clock64() - tc;
      }
      wbOld = wb; ebOld = eb; pj0 = j0;
    }
    (void)wbOld; (void)ebOld;
    if (tid == 0) { extE[j] = E; H.nExtLValid = j; }
    if (dbg && tid == 0) printf("lzi chase block %d: %d literal records, %d windows, %lld cycles (%lld in the chain)\n", b, nExtL, nWin, clock64() - tStart, tChase);
  }
}
// inclusive scan of repeat-offset maps over the CTA: on return (s0, s1) is the map from the CTA's entry state to the state
// after this thread's element; S.totA / S.totB hold the map of the whole CTA tile
__device__ __forceinline__ void lzi_cta_mapscan(u32& s0, u32& s1, LziShared& S, int tid, int lane, int warp) {
  for (int o = 1; o < 32; o <<= 1) {
    const u32 p0 = __shfl_up_sync(0xFFFFFFFFu, s0, o), p1 = __shfl_up_sync(0xFFFFFFFFu, s1, o);
    if (lane >= o) { const u32 n0 = lzi_apply(s0, p0, p1), n1 = lzi_apply(s1, p0, p1); s0 = n0; s1 = n1; }
  }
  __syncthreads();
  if (lane == 31) { S.wm0[warp] = s0; S.wm1[warp] = s1; }
  __syncthreads();
  if (tid == 0) {                                // map from the tile entry to the start of every warp
    u32 r0 = LZI_IN0, r1 = LZI_IN1;
    for (int w = 0; w < 32; w++) {
      S.ws0[w] = r0; S.ws1[w] = r1;
      const u32 n0 = lzi_apply(S.wm0[w], r0, r1), n1 = lzi_apply(S.wm1[w], r0, r1);
      r0 = n0; r1 = n1;
    }
    S.rep0 = r0; S.rep1 = r1;
  }
  __syncthreads();
  const u32 n0 = lzi_apply(s0, S.ws0[warp], S.ws1[warp]), n1 = lzi_apply(s1, S.ws0[warp], S.ws1[warp]);
  s0 = n0; s1 = n1;
}
#define LZI_F_FAIL 1u
#define LZI_F_LIT 2u
#define LZI_F_MATCH 4u
// T5: lengths, cursors and, relative to the tile, output offsets and repeat-offset selectors of every token
__global__ void __launch_bounds__(LZI_TT) lzi_tok_span_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P, LziTok* toks, i64 tokStride, LziHdr* __restrict__ hdrs,
                                                             u32* extPool, u32* tilePool, i64 tileStride) {
  __shared__ LziShared S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.y;
  LziHdr& H = hdrs[b];
  if (!H.valid) return;
  LziLayout Y;
  if (!lzi_layout(blocks[b], P, b, Y)) return;
  const LziArrays A = lzi_arrays(toks, extPool, tilePool, tokStride, tileStride, b);
  const u8* __restrict__ src = blocks[b].cur;
  const int nExtLValid = H.nExtLValid, nExtMValid = H.nExtMValid;
  for (int tile = blockIdx.x; tile * LZI_TT < Y.nTokBytes; tile += gridDim.x) {
  const int base = tile * LZI_TT;
  const int t = base + tid;
  const bool on = t < Y.nTokBytes;
  const int token = on ? src[Y.tkBase + t] : 0;
  const LziFlags F = lzi_flags(token, on);
  const uint4 off = A.tileOff[tile];
  u32 aOff, cOff;
  lzi_cta_scan2(F.known | (F.nd << 16), ((u32)F.lExt << 16) | (u32)F.mExt, aOff, cOff, S, lane, warp);   // (sums stay below 2^16 per tile)
  const u32 kOff = aOff & 0xFFFFu, ndOff = aOff >> 16;
  const u32 jb = off.z + (cOff >> 16), kb = off.w + (cOff & 0xFFFFu);
  bool bad = false;
  u32 litLen = F.known, extSz = 0;
  const u32 Eb = A.extE[min(jb, (u32)nExtLValid)];
  if (F.lExt) {
    if ((int)jb >= nExtLValid) bad = true;
    else { const u32 ls = A.extLS[jb]; litLen = ls & 0x0FFFFFFFu; extSz = ls >> 28; }
  } else if ((int)jb > nExtLValid) bad = true;
  const u32 litSrc = 13u + off.x + kOff + Eb + extSz;
  const u32 litAfter = litSrc + litLen;
  u32 mLen = on ? (u32)(F.isRep ? (token & 3) : (token & 7)) + (u32)Y.minMatch : 0u;
  bool mBad = false;
  if (F.mExt) { if ((int)kb >= nExtMValid) mBad = true; else mLen += A.mxVal[kb]; }
  // the walk ends at the first literal-carrying token whose literals reach srcEnd (:673-674)
  const bool isLast = F.hasLit && (bad || ((i32)litAfter >= Y.srcEndLit));
  __syncthreads();
  if (tid == 0) S.lastIdx = 0x7FFFFFFF;
  __syncthreads();
  if (isLast) atomicMin(&S.lastIdx, tid);
  __syncthreads();
  const int lastIdx = S.lastIdx;
  int nValid = min(LZI_TT, Y.nTokBytes - base);
  bool finished = false;
  if (lastIdx != 0x7FFFFFFF) { nValid = lastIdx + 1; finished = true; }
  const bool live = tid < nValid;
  const bool hasMatch = live && !(finished && tid == nValid - 1);
  if (!hasMatch) mLen = 0;
  bool fail = live && (bad || (hasMatch && mBad));
  if (live && F.hasLit && (litAfter > (u32)Y.litEnd)) fail = true;
  // distances: explicit bytes, then the repeat-offset scan
  u32 dExp = 0;
  if (hasMatch && !F.isRep) {
    const u32 c = (u32)Y.distBase + off.y + ndOff;
    if (c + F.nd > (u32)Y.count + 8) fail = true;
    else { dExp = src[c]; if (F.nd >= 2) dExp = (dExp << 8) | src[c + 1]; if (F.nd == 3) dExp = (dExp << 8) | src[c + 2]; }
  }
  // map of this token: NEW(d): (d, IN0); REP0: (IN0, IN0); REP1: (IN1, IN0); no match: identity
  u32 s0, s1;
  if (!hasMatch) { s0 = LZI_IN0; s1 = LZI_IN1; }
  else if (!F.isRep) { s0 = dExp; s1 = LZI_IN0; }
  else if ((token & 0x04) == 0) { s0 = LZI_IN0; s1 = LZI_IN0; }
  else { s0 = LZI_IN1; s1 = LZI_IN0; }
  lzi_cta_mapscan(s0, s1, S, tid, lane, warp);
  const u32 tile0 = S.rep0, tile1 = S.rep1;
  const u32 span = live ? (litLen + mLen) : 0u;
  u32 spanOff, dummy;
  lzi_cta_scan2(span, 0u, spanOff, dummy, S, lane, warp);
  if (live) {
    LziTok tk; tk.outPos = spanOff; tk.litSrc = litSrc; tk.litLen = litLen; tk.mLen = mLen; tk.dist = s0;
    tk.flags = (fail ? LZI_F_FAIL : 0u) | (F.hasLit ? LZI_F_LIT : 0u) | (hasMatch ? LZI_F_MATCH : 0u);
    A.T[t] = tk;
  }
  if (tid == 0) {
    A.tileSpan[tile] = make_uint4(S.totA, tile0, tile1, 0u);
    if (finished) atomicMin(&H.lastTok, base + lastIdx);
  }
  __syncthreads();
  }
}
// T6 (one CTA per block): output offset and repeat-offset state at the start of every tile
__global__ void __launch_bounds__(LZI_TT) lzi_tok_scan2_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P, LziTok* toks, i64 tokStride, LziHdr* __restrict__ hdrs,
                                                              u32* extPool, u32* tilePool, i64 tileStride) {
  __shared__ LziShared S;
  __shared__ u32 carryOut, c0, c1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x;
  LziHdr& H = hdrs[b];
  if (!H.valid || H.lastTok == 0x7FFFFFFF) return;
  LziLayout Y;
  if (!lzi_layout(blocks[b], P, b, Y)) return;
  const LziArrays A = lzi_arrays(toks, extPool, tilePool, tokStride, tileStride, b);
  const int nTiles = H.lastTok / LZI_TT + 1;
  if (tid == 0) { carryOut = 0; c0 = (u32)Y.count; c1 = (u32)Y.count; }
  __syncthreads();
  for (int base = 0; base < nTiles; base += LZI_TT) {
    const int t = base + tid;
    const uint4 v = (t < nTiles) ? A.tileSpan[t] : make_uint4(0u, LZI_IN0, LZI_IN1, 0u);
    u32 s0 = v.y, s1 = v.z;
    lzi_cta_mapscan(s0, s1, S, tid, lane, warp);          // inclusive: state after tile t as a function of the chunk's entry state
    const u32 all0 = S.rep0, all1 = S.rep1;
    // exclusive = inclusive of the element before
    u32 e0 = __shfl_up_sync(0xFFFFFFFFu, s0, 1), e1 = __shfl_up_sync(0xFFFFFFFFu, s1, 1);
    __syncthreads();
    if (lane == 31) { S.wm0[warp] = s0; S.wm1[warp] = s1; }
    __syncthreads();
    if (lane == 0) { if (warp == 0) { e0 = LZI_IN0; e1 = LZI_IN1; } else { e0 = S.wm0[warp - 1]; e1 = S.wm1[warp - 1]; } }
    u32 oOff, dummy;
    lzi_cta_scan2(v.x, 0u, oOff, dummy, S, lane, warp);
    const u32 totOut = S.totA;
    if (t < nTiles) { A.tileOut[t] = carryOut + oOff; A.tileRep[t] = make_uint2(lzi_apply(e0, c0, c1), lzi_apply(e1, c0, c1)); }
    __syncthreads();
    if (tid == 0) { carryOut += totOut; const u32 n0 = lzi_apply(all0, c0, c1), n1 = lzi_apply(all1, c0, c1); c0 = n0; c1 = n1; }
    __syncthreads();
  }
}
// T7: absolute output offsets and distances, the reference's sanity checks (:657-661, 706-711)
__global__ void __launch_bounds__(256) lzi_tok_final_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P, LziTok* toks, i64 tokStride, LziHdr* __restrict__ hdrs,
                                                           u32* extPool, u32* tilePool, i64 tileStride) {
  const int b = blockIdx.y;
  LziHdr& H = hdrs[b];
  if (!H.valid) return;
  const int lastTok = H.lastTok;
  if (lastTok == 0x7FFFFFFF) return;
  LziLayout Y;
  if (!lzi_layout(blocks[b], P, b, Y)) return;
  const LziArrays A = lzi_arrays(toks, extPool, tilePool, tokStride, tileStride, b);
  for (int t = blockIdx.x * 256 + threadIdx.x; t <= lastTok; t += gridDim.x * 256) {
  const int tile = t / LZI_TT;
  LziTok tk = A.T[t];
  const uint2 rep = A.tileRep[tile];
  const u32 outPos = A.tileOut[tile] + tk.outPos;
  const u32 dist = lzi_apply(tk.dist, rep.x, rep.y);
  bool fail = (tk.flags & LZI_F_FAIL) != 0;
  const u32 dstEnd = (u32)Y.dstEnd;
  if ((tk.flags & LZI_F_LIT) && (tk.litLen > dstEnd - min(outPos, dstEnd))) fail = true;
  if (tk.flags & LZI_F_MATCH) {
    const u32 mStart = outPos + tk.litLen;
    if (dist > mStart || dist == 0 || dist > (u32)Y.maxDist || mStart + tk.mLen > dstEnd) fail = true;
  }
  tk.outPos = outPos; tk.dist = dist;
  A.T[t] = tk;
  if (fail) H.fail = 1;
  if (t == lastTok) {
    H.nTokRaw = lastTok + 1; H.outLenRaw = (i32)(outPos + tk.litLen + tk.mLen);
    H.okRaw = (tk.litSrc + tk.litLen == (u32)Y.litEnd) ? 1 : 0;                               // `return srcIdx == srcEnd + 13`
  }
  }
}
// T8: commit (a failed check anywhere leaves nTok = 0: inverse returns false, res[0] stays 0)
__global__ void lzi_tok_commit_kernel(KzgXfParams P, LziHdr* __restrict__ hdrs, int nBlocks) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nBlocks) return;
  LziHdr& H = hdrs[b];
  if (!H.valid || H.lastTok == 0x7FFFFFFF || H.fail) return;
  H.nTok = H.nTokRaw; H.outLen = H.outLenRaw; H.ok = H.okRaw;
  P.result[2 * b + 1] = H.outLenRaw;              // res[0] is set by the gather pass
}

// ---- pass 2: copy resolution ---------------------------------------------------------------------------------------------------
// Every output byte is a literal (a byte of the literal area) or a copy of an earlier output byte.  ptr[pos] = LIT | literal
// offset, or the output position it copies.  A match that overlaps itself (dist < length) points straight into the bytes
// before it: pos' = mStart - dist + (pos - mStart) mod dist, so runs do not build chains of their own.
//   lzi_resolve_kernel : one CTA per tile of 8192 output bytes.  Pointers are built in shared memory, pointers that stay
//                        inside the tile are resolved there by pointer doubling (no DRAM traffic), literal-resolved bytes are
//                        gathered and written at once; what is left points before the tile.  The tile's pointers go to
//                        global memory (targets of later tiles' lookups), plus a per-tile count of open bytes.
//   lzi_global_kernel  : open tiles only: up to 16 hops through the global pointer array (every hop lands on a pointer its own
//                        tile has already resolved, so a hop crosses a whole tile), writing every byte whose chain ends; a
//                        second launch walks whatever is still open to the end of its chain.
#define LZI_TILE 8192
#define LZI_RT 256                      // threads per resolve CTA
#define LZI_MAXTOK (LZI_TILE / 4 + 8)   // tokens overlapping a tile: every token but a block's last covers >= minMatch (>= 4) bytes
#define LZI_OUT 0x40000000u             // shared-memory pointers only: points before the tile (global position in the low 30 bits)
__global__ void __launch_bounds__(LZI_RT) lzi_resolve_kernel(KzgBlock* __restrict__ blocks, const LziTok* __restrict__ toks, i64 tokStride,
                                                            const LziHdr* __restrict__ hdrs, u32* __restrict__ ptrs, i64 ptrStride,
                                                            int* __restrict__ open, int tilesPerBlock, int* __restrict__ result) {
  __shared__ u32 sPtr[LZI_TILE];
  __shared__ u32 sOut[LZI_MAXTOK];
  __shared__ int sT0;
  const int b = blockIdx.y, tid = threadIdx.x;
  const LziHdr& H = hdrs[b];
  if (blockIdx.x == 0 && tid == 0) result[2 * b] = (H.nTok > 0) ? H.ok : 0;
  if (tid == 0) open[(i64)b * tilesPerBlock + blockIdx.x] = 0;
  if (H.nTok <= 0) return;
  const int outLen = H.outLen;
  const int tileBeg = blockIdx.x * LZI_TILE;
  if (tileBeg >= outLen) return;
  const int tileEnd = min(tileBeg + LZI_TILE, outLen);
  const LziTok* T = toks + (i64)b * tokStride;
  u32* ptr = ptrs + (i64)b * ptrStride;
  const KzgBlock& B = blocks[b];
  const u8* __restrict__ src = B.cur;
  u8* __restrict__ dst = B.alt;
  if (tid == 0) {                       // last token with outPos <= tileBeg
    int lo = 0, hi = H.nTok - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (T[mid].outPos <= (u32)tileBeg) lo = mid; else hi = mid - 1; }
    sT0 = lo;
  }
  __syncthreads();
  const int t0 = sT0;
  for (int i = tid; i < LZI_MAXTOK; i += LZI_RT) {
    const int t = t0 + i;
    sOut[i] = (t < H.nTok) ? T[t].outPos : 0xFFFFFFFFu;
  }
  __syncthreads();
  // build: thread owns 4 consecutive positions per 1024-position slab
  for (int slab = 0; slab < LZI_TILE; slab += 4 * LZI_RT) {
    const int pos0 = tileBeg + slab + 4 * tid;
    if (pos0 >= tileEnd) break;
    int lo = 0, hi = LZI_MAXTOK - 1;    // token of pos0: last i with sOut[i] <= pos0
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (sOut[mid] <= (u32)pos0) lo = mid; else hi = mid - 1; }
    int ti = lo;
    LziTok tk = T[t0 + ti];
    #pragma unroll
    for (int k = 0; k < 4; k++) {
      const u32 pos = (u32)pos0 + k;
      u32 v = LZI_LIT;
      if ((int)pos < tileEnd) {
        while (pos >= tk.outPos + tk.litLen + tk.mLen) { ti++; tk = T[t0 + ti]; }
        const u32 off = pos - tk.outPos;
        if (off < tk.litLen) v = LZI_LIT | (tk.litSrc + off);
        else {
          const u32 k2 = off - tk.litLen;                     // byte of the match
          const u32 from = (k2 < tk.dist) ? pos - tk.dist : (pos - k2) - tk.dist + (k2 % tk.dist);
          v = (from >= (u32)tileBeg) ? (from - (u32)tileBeg) : (LZI_OUT | from);
        }
      }
      sPtr[slab + 4 * tid + k] = v;
    }
  }
  __syncthreads();
  // pointer doubling inside the tile (in place: a reader sees the old or the new pointer of its target, both lead to the same byte)
  for (int round = 0; round < 16; round++) {
    int local = 0;
    for (int i = tid; i < tileEnd - tileBeg; i += LZI_RT) {      // (entries beyond the block's last byte were never built)
      u32 v = sPtr[i];
      if (!(v & (LZI_LIT | LZI_OUT))) {
        v = sPtr[v];
        if (!(v & (LZI_LIT | LZI_OUT))) { v = sPtr[v]; }
        sPtr[i] = v;
        if (!(v & (LZI_LIT | LZI_OUT))) local = 1;
      }
    }
    if (!__syncthreads_or(local)) break;
  }
  // (a chain deeper than 2^32 hops cannot exist in 8192 positions: every local pointer is resolved now)
  int nOpen = 0;
  for (int slab = 0; slab < LZI_TILE; slab += 4 * LZI_RT) {
    const int i0 = slab + 4 * tid;
    const int pos0 = tileBeg + i0;
    if (pos0 >= tileEnd) break;
    u32 v[4]; u32 outw = 0; int nLit = 0;
    #pragma unroll
    for (int k = 0; k < 4; k++) {
      u32 x = sPtr[i0 + k];
      if (x & LZI_LIT) { if (pos0 + k < tileEnd) { outw |= (u32)src[x & ~LZI_LIT] << (8 * k); nLit++; } }
      else { x &= ~LZI_OUT; nOpen++; }
      v[k] = x;
    }
    *reinterpret_cast<uint4*>(ptr + pos0) = make_uint4(v[0], v[1], v[2], v[3]);
    if (nLit == 4) *reinterpret_cast<u32*>(dst + pos0) = outw;
    else {
      #pragma unroll
      for (int k = 0; k < 4; k++) if ((v[k] & LZI_LIT) && pos0 + k < tileEnd) dst[pos0 + k] = (u8)(outw >> (8 * k));
    }
  }
  if (__syncthreads_or(nOpen)) { if (tid == 0) open[(i64)b * tilesPerBlock + blockIdx.x] = 1; }
}

// open tiles only: hop through the global pointers
__global__ void __launch_bounds__(LZI_RT) lzi_global_kernel(KzgBlock* __restrict__ blocks, const LziHdr* __restrict__ hdrs, u32* __restrict__ ptrs, i64 ptrStride,
                                                           int* __restrict__ open, int tilesPerBlock, int finish) {
  const int b = blockIdx.y, tid = threadIdx.x;
  int* flag = open + (i64)b * tilesPerBlock + blockIdx.x;
  if (*flag == 0) return;
  const LziHdr& H = hdrs[b];
  const int outLen = H.outLen;
  const int tileBeg = blockIdx.x * LZI_TILE;
  const int tileEnd = min(tileBeg + LZI_TILE, outLen);
  u32* ptr = ptrs + (i64)b * ptrStride;
  const KzgBlock& B = blocks[b];
  const u8* __restrict__ src = B.cur;
  u8* __restrict__ dst = B.alt;
  int still = 0;
  for (int slab = 0; slab < LZI_TILE; slab += 4 * LZI_RT) {
    const int pos0 = tileBeg + slab + 4 * tid;
    if (pos0 >= tileEnd) break;
    const uint4 q = *reinterpret_cast<const uint4*>(ptr + pos0);
    u32 v[4] = {q.x, q.y, q.z, q.w};
    if ((v[0] & v[1] & v[2] & v[3]) & LZI_LIT) continue;
    u32 p[4] = {v[0], v[1], v[2], v[3]};
    const int hops = finish ? (1 << 30) : 16;
    for (int hop = 0; hop < hops; hop++) {          // the four chains advance together: four independent loads in flight per step
      const bool o0 = !(p[0] & LZI_LIT), o1 = !(p[1] & LZI_LIT), o2 = !(p[2] & LZI_LIT), o3 = !(p[3] & LZI_LIT);
      if (!(o0 | o1 | o2 | o3)) break;
      const u32 n0 = o0 ? __ldcg(ptr + p[0]) : p[0], n1 = o1 ? __ldcg(ptr + p[1]) : p[1], n2 = o2 ? __ldcg(ptr + p[2]) : p[2], n3 = o3 ? __ldcg(ptr + p[3]) : p[3];
      p[0] = n0; p[1] = n1; p[2] = n2; p[3] = n3;
    }
    bool changed = false;
    #pragma unroll
    for (int k = 0; k < 4; k++) {
      if (p[k] != v[k]) {
        changed = true;
        if ((p[k] & LZI_LIT) && pos0 + k < tileEnd) dst[pos0 + k] = src[p[k] & ~LZI_LIT];
      }
      if (!(p[k] & LZI_LIT)) still = 1;
    }
    if (changed) *reinterpret_cast<uint4*>(ptr + pos0) = make_uint4(p[0], p[1], p[2], p[3]);
  }
  still = __syncthreads_or(still);
  if (tid == 0 && !still) *flag = 0;
}

// scratch: per block tokens (24 B each, up to maxLen/4 + 1024) + 16 B of record lists per token + 64 B per tile of 1024 tokens
// + 4 bytes per output byte + header
static i64 lzi_tok_stride(i32 maxLen) { return (((i64)maxLen / 4 + 1024) + 15) & ~(i64)15; }
static i64 lzi_tile_stride(i64 tokStride) { return (((tokStride + LZI_TT - 1) / LZI_TT + 8) + 3) & ~(i64)3; }
void kzg_lzi_scratch(i32 maxLen, size_t* perBlockBytes, size_t* aux32) {
  const size_t toks = (size_t)lzi_tok_stride(maxLen);
  *perBlockBytes = std::max(*perBlockBytes, toks * (sizeof(LziTok) + 16) + (size_t)lzi_tile_stride((i64)toks) * 64 + 256 + 512 + 4 * ((size_t)maxLen / LZI_TILE + 2));
  *aux32 = std::max(*aux32, (size_t)maxLen + 64);
}

int kzg_lz_inverse_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, i32 maxLen) {
  // flat scratch pool (nBlocks * scratchStride bytes): [dense headers, 256 B reserved per block][token arrays]; pointers in aux32
  const i64 tokStride = lzi_tok_stride(maxLen);
  const i64 tileStride = lzi_tile_stride(tokStride);
  const size_t need = (size_t)nBlocks * (256 + (size_t)tokStride * (sizeof(LziTok) + 16) + (size_t)tileStride * 64 + 4 * ((size_t)maxLen / LZI_TILE + 2));
  if (need > (size_t)nBlocks * (size_t)P.scratchStride) { kzg_set_error("lz inverse: scratch pool too small"); return -KZG_ERR_CREATE_CODEC; }
  static_assert(sizeof(LziHdr) <= 256, "LziHdr must fit its 256-byte slot");
  LziHdr* hdrs = (LziHdr*)P.scratch;
  LziTok* toks = (LziTok*)(P.scratch + (size_t)nBlocks * 256);
  u32* extPool = (u32*)(toks + (size_t)nBlocks * tokStride);
  u32* tilePool = extPool + (size_t)nBlocks * 4 * tokStride;
  u32* ptrs = (u32*)P.aux32;
  const i64 ptrStride = P.aux32Stride;
  const int tokTiles = (int)((tokStride + LZI_TT - 1) / LZI_TT);
  const dim3 gT(std::min(tokTiles, 96), nBlocks);     // CTAs loop over the token tiles that exist (the count is device-side knowledge)
  KZG_PROF("lzi_tok_sums_kernel", s, (lzi_tok_sums_kernel<<<gT, LZI_TT, 0, s>>>(d_blocks, P, toks, tokStride, hdrs, extPool, tilePool, tileStride)));
  KZG_PROF("lzi_tok_scan1_kernel", s, (lzi_tok_scan1_kernel<<<nBlocks, LZI_TT, 0, s>>>(d_blocks, P, toks, tokStride, hdrs, extPool, tilePool, tileStride)));
  KZG_PROF("lzi_tok_extk_kernel", s, (lzi_tok_extk_kernel<<<gT, LZI_TT, 0, s>>>(d_blocks, P, toks, tokStride, hdrs, extPool, tilePool, tileStride)));
  // The chase is one dependent chain per block on one warp: 4.5 ms alone for cfg2's longest block, but up to 7.8 ms while the other
  // block groups' throughput kernels (span, resolve, global: CTAs on every SM) share its SM — the chain thread's own loop slows from 54
  // to 97 cycles per record, because it has to share its scheduler's issue slots.  The CTA therefore asks for ALL of the SM's shared
  // memory (every CTA needs at least its 1 KiB system slice, so nothing else can become resident next to it): the chains get an SM to
  // themselves, the throughput kernels the other SMs (only when the batch has no more blocks than the GPU has SMs: KZG_XF_LZI_EXCLUSIVE).  (200 KiB was not enough — small-footprint kernels still fitted — and 32 KiB
  // windows or a barrier between all groups' chains and the rest of the stage were slower.)
  int chaseSmem = (int)sizeof(LziChaseSmem);
  {
    int dev = 0, optin = 0;
    cudaFuncAttributes fa;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess &&
        cudaFuncGetAttributes(&fa, lzi_tok_chase_kernel) == cudaSuccess && optin - (int)fa.sharedSizeBytes > chaseSmem && (P.flags & KZG_XF_LZI_EXCLUSIVE))
      chaseSmem = optin - (int)fa.sharedSizeBytes;
    else cudaGetLastError();
  }
  CUDA_TRY(cudaFuncSetAttribute(lzi_tok_chase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, chaseSmem));
  CUDA_TRY(cudaFuncSetAttribute(lzi_tok_chase_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  KZG_PROF("lzi_tok_chase_kernel", s, (lzi_tok_chase_kernel<<<nBlocks, LZI_TT, chaseSmem, s>>>(d_blocks, P, toks, tokStride, hdrs, extPool, tilePool, tileStride)));
  KZG_PROF("lzi_tok_span_kernel", s, (lzi_tok_span_kernel<<<gT, LZI_TT, 0, s>>>(d_blocks, P, toks, tokStride, hdrs, extPool, tilePool, tileStride)));
  KZG_PROF("lzi_tok_scan2_kernel", s, (lzi_tok_scan2_kernel<<<nBlocks, LZI_TT, 0, s>>>(d_blocks, P, toks, tokStride, hdrs, extPool, tilePool, tileStride)));
  KZG_PROF("lzi_tok_final_kernel", s, (lzi_tok_final_kernel<<<dim3(std::min((int)((tokStride + 255) / 256), 256), nBlocks), 256, 0, s>>>(d_blocks, P, toks, tokStride, hdrs, extPool, tilePool, tileStride)));
  KZG_PROF("lzi_tok_commit_kernel", s, (lzi_tok_commit_kernel<<<(nBlocks + 63) / 64, 64, 0, s>>>(P, hdrs, nBlocks)));
  const int tiles = (maxLen + LZI_TILE - 1) / LZI_TILE;
  int* open = (int*)(tilePool + (size_t)nBlocks * 16 * tileStride);
  KZG_PROF("lzi_resolve_kernel", s, (lzi_resolve_kernel<<<dim3(tiles, nBlocks), LZI_RT, 0, s>>>(d_blocks, toks, tokStride, hdrs, ptrs, ptrStride, open, tiles, P.result)));
  KZG_PROF("lzi_global_kernel", s, (lzi_global_kernel<<<dim3(tiles, nBlocks), LZI_RT, 0, s>>>(d_blocks, hdrs, ptrs, ptrStride, open, tiles, 0)));
  KZG_PROF("lzi_global_kernel", s, (lzi_global_kernel<<<dim3(tiles, nBlocks), LZI_RT, 0, s>>>(d_blocks, hdrs, ptrs, ptrStride, open, tiles, 1)));
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(11);
  return 0;
}
