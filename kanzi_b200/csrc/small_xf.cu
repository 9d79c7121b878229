// small_xf.cu — ZRLT, SBRT (RANK / MTFT) and SRT kernels (sm_100a).
//
// Replaces K/transform/ZRLT.java, SBRT.java, SRT.java.
// ZRLT is a scan: 4096-byte tiles, one warp each, classify bytes with ballots (zero runs, 0xFF escapes), size each lane's
// output, prefix-sum the sizes and scatter; what crosses a tile edge (run in progress, digit sequence, escape parity) is carried by
// a scan over per-tile summaries.
// SBRT / SRT forward are tile-parallel too (the emitted rank is a function of the symbols' last occurrences); their inverses are
// list-update state machines (the output at i depends on the list after i-1): one warp per block, the head of the list in
// registers.
#include "kzg_common.cuh"
#include "kzg_transforms.cuh"
#include "kzg_xf_kernels.cuh"

__device__ __forceinline__ u32 warp_excl_scan(u32 v, int lane, u32& total) {
  u32 incl = v;
  for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
  total = __shfl_sync(0xFFFFFFFFu, incl, 31);
  return incl - v;
}

// ================================================================================================================
// ZRLT (ZRLT.java:54-136 forward, 146-233 inverse), tile-parallel: a block is cut into tiles of ZR_TILE input bytes, one
// warp per tile.  What a tile emits depends on the tiles before it only through a few carried values (forward: the zero run
// in progress; inverse: the digit sequence in progress and whether an escape is waiting for its payload), so a summary pass
// (one warp per tile), a scan over the summaries (one warp per block) and an emit pass (one warp per tile) reproduce the
// sequential result, including every "no room" test of the reference (each one is evaluated with the same dstIdx).
// ================================================================================================================
#define ZR_TILE 4096
struct ZrSum { u64 a, b; u32 c, d, e, f; };     // 32 bytes per tile (fields documented at the kernels)

// digit sequence value `val` (leading 1 included; 1 = none) extended by n more digits `bits` (MSB first), in Java int arithmetic
// (`runLength += runLength + val` wraps, ZRLT.java:186): after 32 or more digits only the last 32 are left
__device__ __forceinline__ u32 zr_append(u32 val, int n, u32 bits) {
  if (n == 0) return val;
  if (n >= 32) return bits;
  return (val << n) | bits;
}
// zeros a digit sequence stands for: runLength - 1 if that is positive as a Java int (:193-196)
__device__ __forceinline__ u32 zr_zeros(u32 val) { const i32 z = (i32)(val - 1u); return z > 0 ? (u32)z : 0u; }
// digits of lanes [lo, lo + m) of a ballot of ones, first lane most significant
__device__ __forceinline__ u32 zr_bits(u32 onemask, int lo, int m) {
  if (m <= 0) return 0u;
  const u32 w = (onemask >> lo) & ((m >= 32) ? 0xFFFFFFFFu : ((1u << m) - 1));
  return __brev(w) >> (32 - m);
}

// ---- forward: one 32-byte step of the original scan; `carry` = zeros pending before the step -------------------------
template <bool EMIT>
__device__ __forceinline__ void zrf_tile(const u8* __restrict__ src, u8* __restrict__ dst, int beg, int end, int dstEnd, u32 carry, u32 dstIdx,
                                         int lane, ZrSum* sum, int* res) {
  const u32 lowMask = (1u << lane) - 1;
  bool fail = false, reach = true;          // reach: no non-zero byte of this tile seen yet
  u32 size0 = 0, lead = 0;
  for (int base = beg; base < end; base += 32) {
    const int i = base + lane;
    const bool valid = i < end;
    const int v = valid ? src[i] : 1;
    const bool z = valid && (v == 0);
    const u32 zmask = __ballot_sync(0xFFFFFFFFu, z);
    const u32 vmask = __ballot_sync(0xFFFFFFFFu, valid);
    const u32 nzBelow = ~zmask & lowMask;
    u32 run;
    if (nzBelow == 0) run = (u32)lane + carry;
    else run = (u32)(lane - 1 - (31 - __clz(nzBelow)));
    u32 size = 0; int lg = 0;
    const bool first = reach && nzBelow == 0;             // this lane's run reaches the start of the tile
    if (valid && !z) {
      if (run > 0) lg = ilog2(run + 1);
      size = (u32)((EMIT || !first) ? lg : 0) + ((v >= 0xFE) ? 2u : 1u);
    }
    u32 total;
    const u32 off = warp_excl_scan(size, lane, total);
    if (EMIT) {
      if (valid && !z) {
        u32 o = dstIdx + off;
        if (run > 0) {
          if ((i64)o >= (i64)dstEnd - lg) fail = true;        // :76
          else { const u32 rl = run + 1; for (int k = lg - 1; k >= 0; k--) dst[o++] = (u8)((rl >> k) & 1); }
        }
        if (!fail) {
          if (v >= 0xFE) {
            if ((i64)o >= (i64)dstEnd - 1) fail = true;       // :94
            else { dst[o] = 0xFF; dst[o + 1] = (u8)(v - 0xFE); }
          } else {
            if ((i64)o >= (i64)dstEnd) fail = true;           // :111
            else dst[o] = (u8)(v + 1);
          }
        }
      }
      if (__any_sync(0xFFFFFFFFu, fail)) { fail = true; break; }
      dstIdx += total;
    } else size0 += total;
    const u32 nzAll = ~zmask & vmask;
    if (!EMIT && reach && nzAll != 0) lead = carry + (u32)(__ffs(nzAll) - 1);
    if (nzAll == 0) carry += __popc(vmask);
    else { carry = (u32)__popc(vmask) - 1 - (31 - __clz(nzAll)); reach = false; }
  }
  if (EMIT) { if (fail && lane == 0) res[0] = 0; }
  else if (lane == 0) { sum->a = size0; sum->c = reach ? carry : lead; sum->d = carry; sum->e = reach ? 1u : 0u; }
}
// summary: a = bytes emitted except the run digits of the first non-zero byte, c = zeros at the start of the tile (all of it when
// e = 1), d = zeros at its end.  After the scan: a = dstIdx at the tile, c = zeros pending at its start.
__global__ void __launch_bounds__(32) zrlt_fwd_sum_kernel(const KzgBlock* __restrict__ blocks, KzgXfParams P) {
  const int lane = threadIdx.x, b = blockIdx.y, tile = blockIdx.x;
  const KzgBlock& B = blocks[b];
  if (B.status != 0 || !P.enabled[b]) return;
  const int count = B.curLen, beg = tile * ZR_TILE;
  if (beg >= count) return;
  ZrSum* sums = reinterpret_cast<ZrSum*>(P.scratch + (i64)b * P.scratchStride);
  zrf_tile<false>(B.cur, nullptr, beg, min(beg + ZR_TILE, count), count, 0u, 0u, lane, sums + tile, nullptr);
}
__global__ void __launch_bounds__(32) zrlt_fwd_scan_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P) {
  const int lane = threadIdx.x, b = blockIdx.x;
  KzgBlock& B = blocks[b];
  int* res = P.result + 2 * b;
  if (lane == 0) { res[0] = 0; res[1] = 0; }
  if (B.status != 0 || !P.enabled[b]) return;
  const int count = B.curLen;
  const int nTiles = (count + ZR_TILE - 1) / ZR_TILE;
  ZrSum* sums = reinterpret_cast<ZrSum*>(P.scratch + (i64)b * P.scratchStride);
  u32 carry = 0, off = 0;
  for (int t0 = 0; t0 < nTiles; t0 += 32) {
    const int t = t0 + lane;
    u32 a = 0, c = 0, d = 0, e = 1;
    if (t < nTiles) { a = (u32)sums[t].a; c = sums[t].c; d = sums[t].d; e = sums[t].e; }
    u32 myOff = 0, myCarry = 0;
    const int n = min(32, nTiles - t0);
    for (int k = 0; k < n; k++) {           // the carried pair walks the 32 summaries (a few shuffles each)
      const u32 ak = __shfl_sync(0xFFFFFFFFu, a, k), ck = __shfl_sync(0xFFFFFFFFu, c, k);
      const u32 dk = __shfl_sync(0xFFFFFFFFu, d, k), ek = __shfl_sync(0xFFFFFFFFu, e, k);
      if (lane == k) { myOff = off; myCarry = carry; }
      u32 size = ak;
      if (!ek) { const u32 run = ck + carry; if (run > 0) size += (u32)ilog2(run + 1); carry = dk; }
      else carry += ck;
      off += size;
    }
    if (t < nTiles) { sums[t].a = myOff; sums[t].c = myCarry; }
  }
  bool fail = false;
  u8* __restrict__ dst = B.alt;
  if (carry > 0) {                          // input ends inside a zero run
    const u32 rl = carry + 1;
    const int lg = ilog2(rl);
    if ((i64)off >= (i64)count - lg) fail = true;
    else if (lane == 0) { for (int k = lg - 1; k >= 0; k--) dst[off + (lg - 1 - k)] = (u8)((rl >> k) & 1); }
    off += lg;
  }
  if (lane == 0 && !fail) { res[0] = 1; res[1] = (int)off; }     // (the emit pass clears res[0] when one of its tests fails)
}
__global__ void __launch_bounds__(32) zrlt_fwd_emit_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P) {
  const int lane = threadIdx.x, b = blockIdx.y, tile = blockIdx.x;
  KzgBlock& B = blocks[b];
  if (B.status != 0 || !P.enabled[b]) return;
  const int count = B.curLen, beg = tile * ZR_TILE;
  if (beg >= count) return;
  const ZrSum* sums = reinterpret_cast<const ZrSum*>(P.scratch + (i64)b * P.scratchStride);
  zrf_tile<true>(B.cur, B.alt, beg, min(beg + ZR_TILE, count), count, sums[tile].c, (u32)sums[tile].a, lane, nullptr, P.result + 2 * b);
}

// ---- inverse -----------------------------------------------------------------------------------------------------------
// 0xFF bytes at the end of every tile (an escape pairs with the byte after it: whether a tile starts with a payload byte is the
// parity of the 0xFF run that ends at its first byte)
__global__ void __launch_bounds__(32) zrlt_inv_ff_kernel(const KzgBlock* __restrict__ blocks, KzgXfParams P, int maxTiles) {
  const int lane = threadIdx.x, b = blockIdx.y, tile = blockIdx.x;
  const KzgBlock& B = blocks[b];
  if (B.status != 0 || !P.enabled[b]) return;
  const int count = B.curLen, beg = tile * ZR_TILE;
  if (beg >= count) return;
  const int end = min(beg + ZR_TILE, count);
  const u8* __restrict__ src = B.cur;
  int cnt = 0;
  for (int top = end; top > beg; top -= 32) {         // 32 bytes at a time from the end
    const int i = top - 1 - lane;
    const bool valid = i >= beg;
    const u32 nonFF = __ballot_sync(0xFFFFFFFFu, valid && src[i] != 0xFF);
    if (nonFF == 0) { cnt += min(32, top - beg); continue; }
    cnt += __ffs(nonFF) - 1;
    break;
  }
  u32* ff = reinterpret_cast<u32*>(P.scratch + (i64)b * P.scratchStride + (i64)maxTiles * sizeof(ZrSum));
  if (lane == 0) ff[tile] = (u32)cnt;
}
__device__ __forceinline__ bool zri_pending_esc(const u32* ff, int tile) {
  u32 run = 0;
  for (int u = tile - 1; u >= 0; u--) { const u32 c = ff[u]; run += c; if (c != ZR_TILE) break; }
  return (run & 1u) != 0;
}

// one tile of the original scan.  EMIT = false: sizes only (a digit sequence entering the tile counts as absent: the scan corrects)
template <bool EMIT>
__device__ __forceinline__ void zri_tile(const u8* __restrict__ src, u8* __restrict__ dst, int beg, int end, i64 dstEnd, u32 carryVal, bool pendingEsc,
                                         i64 dstIdx, int lane, ZrSum* sum, int* res) {
  const u32 lowMask = (1u << lane) - 1;
  bool fail = false, lead = true;       // lead: every byte of the tile so far is a digit
  u64 sizeLocal = 0;
  int nL = 0; u32 bitsL = 0;
  for (int base = beg; base < end; base += 32) {
    const int i = base + lane;
    const bool valid = i < end;
    const int v = valid ? src[i] : 2;
    const u32 vmask = __ballot_sync(0xFFFFFFFFu, valid);
    const u32 ffmask = __ballot_sync(0xFFFFFFFFu, valid && v == 0xFF);
    const u32 nonFFBelow = ~ffmask & lowMask;
    bool payload;
    if (nonFFBelow == 0) payload = (((lane & 1) != 0) != pendingEsc);
    else payload = (((lane - 1 - (31 - __clz(nonFFBelow))) & 1) != 0);
    payload = payload && valid;
    const bool digit = valid && !payload && v <= 1;
    const bool esc = valid && !payload && v == 0xFF;
    const u32 dmask = __ballot_sync(0xFFFFFFFFu, digit);
    const u32 onemask = __ballot_sync(0xFFFFFFFFu, digit && v == 1);
    const u32 ndBelow = ~dmask & lowMask;
    int k; bool reachesStart;
    if (ndBelow == 0) { k = lane; reachesStart = true; }
    else { k = lane - 1 - (31 - __clz(ndBelow)); reachesStart = false; }
    u32 zeros = 0;
    const bool closer = valid && !digit && !payload;          // literal or escape: closes a digit run if one precedes it
    if (closer) zeros = zr_zeros(zr_append(reachesStart ? carryVal : 1u, k, zr_bits(onemask, lane - k, k)));
    const bool lit = closer && !esc;
    const int nv = __popc(vmask);
    const u32 ndAll = ~dmask & vmask;
    if (EMIT) {
      if ((i64)zeros >= dstEnd) { fail = true; zeros = 0; }
      if (__any_sync(0xFFFFFFFFu, fail)) { fail = true; break; }
      const u32 size = zeros + ((lit || payload) ? 1u : 0u);
      u32 total;
      const u32 off = warp_excl_scan(size, lane, total);
      // bounds: a run followed by its closer needs dstIdx + run < dstEnd (:172); every byte needs dstIdx < dstEnd
      const i64 o = dstIdx + off;
      if (closer && zeros > 0 && o + (i64)zeros >= dstEnd) fail = true;
      if ((lit || payload) && o + (i64)zeros >= dstEnd) fail = true;
      if (__any_sync(0xFFFFFFFFu, fail)) { fail = true; break; }
      if (zeros > 0 && zeros <= 64) for (u32 t = 0; t < (u32)zeros; t++) dst[o + t] = 0;
      if (lit) dst[o + zeros] = (u8)(v - 1);
      else if (payload) dst[o] = (u8)(0xFE + v);
      u32 longMask = __ballot_sync(0xFFFFFFFFu, zeros > 64);       // long runs: the whole warp fills them
      while (longMask) {
        const int l = __ffs(longMask) - 1;
        longMask &= longMask - 1;
        const i64 o2 = __shfl_sync(0xFFFFFFFFu, o, l);
        const u32 z2 = __shfl_sync(0xFFFFFFFFu, zeros, l);
        for (u32 t = lane; t < z2; t += 32) dst[o2 + t] = 0;
      }
      dstIdx += total;
    } else {
      u64 size = (u64)zeros + ((lit || payload) ? 1u : 0u);
      for (int o = 16; o > 0; o >>= 1) size += __shfl_xor_sync(0xFFFFFFFFu, size, o);
      sizeLocal += size;
      if (lead) {                             // digits before the tile's first non-digit byte
        const int m = (ndAll == 0) ? nv : (__ffs(ndAll) - 1);
        bitsL = zr_append(bitsL, m, zr_bits(onemask, 0, m)); nL = min(nL + m, 64);
        if (ndAll != 0) lead = false;
      }
    }
    if (ndAll == 0) carryVal = zr_append(carryVal, nv, zr_bits(onemask, 0, nv));      // the whole step is digits
    else {
      const int top = 31 - __clz(ndAll);            // highest non-digit lane
      const int t = nv - 1 - top;                   // trailing digits
      carryVal = zr_append(1u, t, zr_bits(onemask, top + 1, t));
    }
    pendingEsc = (__shfl_sync(0xFFFFFFFFu, (int)esc, nv - 1) != 0);
  }
  if (EMIT) { if (fail && lane == 0) res[0] = 0; }
  else {
    // a = bytes emitted with no digit sequence entering, b = digit sequence in progress at the end, c / d = number (capped at 64) and
    // value (last 32) of the digits before the first non-digit byte, e = 1: the tile is digits only
    if (lane == 0) { sum->a = sizeLocal; sum->b = carryVal; sum->c = (u32)nL; sum->d = bitsL; sum->e = lead ? 1u : 0u; sum->f = 0; }
  }
}
__global__ void __launch_bounds__(32) zrlt_inv_sum_kernel(const KzgBlock* __restrict__ blocks, KzgXfParams P, int maxTiles) {
  const int lane = threadIdx.x, b = blockIdx.y, tile = blockIdx.x;
  const KzgBlock& B = blocks[b];
  if (B.status != 0 || !P.enabled[b]) return;
  const int count = B.curLen, beg = tile * ZR_TILE;
  if (beg >= count) return;
  ZrSum* sums = reinterpret_cast<ZrSum*>(P.scratch + (i64)b * P.scratchStride);
  const u32* ff = reinterpret_cast<const u32*>(P.scratch + (i64)b * P.scratchStride + (i64)maxTiles * sizeof(ZrSum));
  zri_tile<false>(B.cur, nullptr, beg, min(beg + ZR_TILE, count), 0, 1u, zri_pending_esc(ff, tile), 0, lane, sums + tile, nullptr);
}
__global__ void __launch_bounds__(32) zrlt_inv_scan_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P) {
  const int lane = threadIdx.x, b = blockIdx.x;
  KzgBlock& B = blocks[b];
  int* res = P.result + 2 * b;
  if (lane == 0) { res[0] = 0; res[1] = 0; }
  if (B.status != 0 || !P.enabled[b]) return;
  const int count = B.curLen;
  const i64 dstEnd = min(kzg_dst_limit(B, P.dstLimit[b]), B.cap);
  const int nTiles = (count + ZR_TILE - 1) / ZR_TILE;
  ZrSum* sums = reinterpret_cast<ZrSum*>(P.scratch + (i64)b * P.scratchStride);
  u32 carry = 1; u64 off = 0;
  bool fail = false;
  if (lane == 0) {
    for (int t = 0; t < nTiles; t++) {
      const ZrSum S = sums[t];
      sums[t].a = off; sums[t].b = carry;
      if (S.e) { carry = zr_append(carry, (int)S.c, S.d); continue; }
      // the digit sequence that enters the tile changes the run its first non-digit byte closes
      u64 size = S.a - zr_zeros(zr_append(1u, (int)S.c, S.d)) + zr_zeros(zr_append(carry, (int)S.c, S.d));
      if (size > (1ull << 32)) { fail = true; size = 0; }
      off += size;
      if (off > (1ull << 32)) { fail = true; off = 0; }
      carry = (u32)S.b;
    }
  }
  carry = __shfl_sync(0xFFFFFFFFu, carry, 0); off = __shfl_sync(0xFFFFFFFFu, off, 0);
  fail = __shfl_sync(0xFFFFFFFFu, (int)fail, 0) != 0;
  u8* __restrict__ dst = B.alt;
  if (!fail && (i32)carry > 0) {      // the input ends inside a digit sequence: runLength - 1 trailing zeros (:221-229)
    const u64 zeros = carry - 1;
    if ((i64)off + (i64)zeros > dstEnd) fail = true;
    else { for (u64 t = lane; t < zeros; t += 32) dst[off + t] = 0; off += zeros; }
  }
  if (lane == 0 && !fail) { res[0] = 1; res[1] = (int)off; }      // (the emit pass clears res[0] when one of its tests fails)
}
__global__ void __launch_bounds__(32) zrlt_inv_emit_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P, int maxTiles) {
  const int lane = threadIdx.x, b = blockIdx.y, tile = blockIdx.x;
  KzgBlock& B = blocks[b];
  if (B.status != 0 || !P.enabled[b]) return;
  const int count = B.curLen, beg = tile * ZR_TILE;
  if (beg >= count) return;
  const ZrSum* sums = reinterpret_cast<const ZrSum*>(P.scratch + (i64)b * P.scratchStride);
  const u32* ff = reinterpret_cast<const u32*>(P.scratch + (i64)b * P.scratchStride + (i64)maxTiles * sizeof(ZrSum));
  int* res = P.result + 2 * b;
  if (res[0] == 0) return;                 // the scan already failed the block
  const i64 dstEnd = min(kzg_dst_limit(B, P.dstLimit[b]), B.cap);
  zri_tile<true>(B.cur, B.alt, beg, min(beg + ZR_TILE, count), dstEnd, (u32)sums[tile].b, zri_pending_esc(ff, tile), (i64)sums[tile].a, lane, nullptr, res);
}

// ================================================================================================================
// SBRT forward / inverse (SBRT.java:87-151, 154-214); mode 1 = MTF, 2 = RANK, 3 = TIMESTAMP
// ================================================================================================================
// ---- SBRT forward, tile-parallel -----------------------------------------------------------------------------------------
// The list of SBRT is always sorted by (q, time of the last move) descending (a symbol that gets key qc moves above everything
// with q <= qc, SBRT.java:138-146; q never decreases), with the symbols never seen yet at the bottom in index order.  So the
// rank the forward transform emits at position i for symbol c is a pure function of every symbol's last two occurrences before
// i:  rank = #{ d : key_d > key_c },  key = (q << 31) | (last occurrence + 256)  for seen symbols, 255 - d for unseen ones, with
// q = ((t1 & m1) + (t2 & m2)) >> s of the last occurrence t1 and the one before it t2 (0 if none: `p[]` starts at 0).
// Three kernels: last two occurrences of every symbol per tile of 4096 positions; a scan over the tiles (one thread per
// symbol) turns them into every tile's entry state; one warp per tile then replays its 4096 positions from that state, every
// lane holding 8 of the 256 keys in registers (8 compares + one warp reduction per position).
#define SBRT_TILE 4096
__global__ void __launch_bounds__(32) sbrt_fwd_last2_kernel(const KzgBlock* __restrict__ blocks, KzgXfParams P) {
  __shared__ int l1[256], l2[256];
  const int lane = threadIdx.x, b = blockIdx.y, tile = blockIdx.x;
  const KzgBlock& B = blocks[b];
  if (B.status != 0 || !P.enabled[b]) return;
  const int count = B.curLen;
  const int beg = tile * SBRT_TILE;
  if (beg >= count || count > B.cap) return;
  const int end = min(beg + SBRT_TILE, count);
  const u8* __restrict__ src = B.cur;
  for (int i = lane; i < 256; i += 32) { l1[i] = -1; l2[i] = -1; }
  __syncwarp();
  for (int base = beg; base < end; base += 32) {
    const int p = base + lane;
    const bool on = p < end;
    const int c = on ? (int)src[p] : 256 + lane;
    const u32 peers = __match_any_sync(0xFFFFFFFFu, c);
    if (on && (peers >> lane) == 1u) {                      // highest lane of its symbol group: the last occurrence of this chunk
      const u32 below = peers & ((1u << lane) - 1);
      l2[c] = below ? base + (31 - __clz(below)) : l1[c];
      l1[c] = p;
    }
    __syncwarp();
  }
  int* info = reinterpret_cast<int*>(P.scratch + (i64)b * P.scratchStride) + (i64)tile * 512;
  for (int i = lane; i < 256; i += 32) { info[2 * i] = l1[i]; info[2 * i + 1] = l2[i]; }
}
__global__ void __launch_bounds__(256) sbrt_fwd_scan_kernel(const KzgBlock* __restrict__ blocks, KzgXfParams P) {
  const int b = blockIdx.x, sym = threadIdx.x;
  const KzgBlock& B = blocks[b];
  if (B.status != 0 || !P.enabled[b] || B.curLen > B.cap) return;
  const int nTiles = (B.curLen + SBRT_TILE - 1) / SBRT_TILE;
  int* info = reinterpret_cast<int*>(P.scratch + (i64)b * P.scratchStride);
  int t1 = -1, t2 = -1;
  for (int t = 0; t < nTiles; t++) {
    int* e = info + (i64)t * 512 + 2 * sym;
    const int a1 = e[0], a2 = e[1];
    e[0] = t1; e[1] = t2;                                   // the tile's entry state replaces its own summary
    if (a1 >= 0) { t2 = (a2 >= 0) ? a2 : t1; t1 = a1; }
  }
}
__global__ void __launch_bounds__(32) sbrt_fwd_rank_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P, int mode) {
  __shared__ u64 K[256];
  const int lane = threadIdx.x, b = blockIdx.y, tile = blockIdx.x;
  KzgBlock& B = blocks[b];
  int* res = P.result + 2 * b;
  if (B.status != 0 || !P.enabled[b]) { if (tile == 0 && lane == 0) { res[0] = 0; res[1] = 0; } return; }
  const int count = B.curLen;
  if (count > B.cap) { if (tile == 0 && lane == 0) { res[0] = 0; res[1] = 0; } return; }
  if (tile == 0 && lane == 0) { res[0] = 1; res[1] = count; }
  const int beg = tile * SBRT_TILE;
  if (beg >= count) return;
  const int end = min(beg + SBRT_TILE, count);
  const u8* __restrict__ src = B.cur;
  u8* __restrict__ dst = B.alt;
  const int m1 = (mode == 3) ? 0 : -1, m2 = (mode == 1) ? 0 : -1, s = (mode == 2) ? 1 : 0;
  const int* info = reinterpret_cast<const int*>(P.scratch + (i64)b * P.scratchStride) + (i64)tile * 512;
  u64 kr[8];
  #pragma unroll
  for (int k = 0; k < 8; k++) {
    const int d = lane + 32 * k;
    const int t1 = info[2 * d], t2 = info[2 * d + 1];
    u64 key = (u64)(255 - d);
    if (t1 >= 0) { const int q = ((t1 & m1) + (max(t2, 0) & m2)) >> s; key = ((u64)(u32)q << 31) | (u64)(u32)(t1 + 256); }
    kr[k] = key; K[d] = key;
  }
  __syncwarp();
  for (int base = beg; base < end; base += 32) {
    const int nIn = min(32, end - base);
    const int mine = (lane < nIn) ? (int)src[base + lane] : 0;
    int outv = 0;
    for (int t = 0; t < nIn; t++) {
      const int i = base + t;
      const int c = __shfl_sync(0xFFFFFFFFu, mine, t);
      const u64 kc = K[c];
      int cnt = 0;
      #pragma unroll
      for (int k = 0; k < 8; k++) cnt += (kr[k] > kc) ? 1 : 0;
      const int r = (int)__reduce_add_sync(0xFFFFFFFFu, (unsigned)cnt);
      if (lane == t) outv = r;
      const u32 low = (u32)(kc & 0x7FFFFFFFull);
      const int pOld = (low >= 256u) ? (int)(low - 256u) : 0;
      const int qc = ((i & m1) + (pOld & m2)) >> s;
      const u64 nk = ((u64)(u32)qc << 31) | (u64)(u32)(i + 256);
      __syncwarp();
      if (lane == 0) K[c] = nk;
      if (lane == (c & 31)) {
        #pragma unroll
        for (int k = 0; k < 8; k++) if (k == (c >> 5)) kr[k] = nk;
      }
      __syncwarp();
    }
    if (lane < nIn) dst[base + lane] = (u8)outv;
  }
}

// ---- SBRT inverse: one warp per block (the list state depends on every symbol decoded so far).  The list is kept in rank order
// (symbol + key per rank): ranks 0-31 live in registers, one per lane (a lone warp issues an instruction every ~5 cycles, so the
// step is written for the fewest instructions: three shuffles fetch the hit entry, one ballot finds its new rank, three
// shuffle-ups move the entries in between); ranks 32-255 stay in shared memory and are touched only when a deep rank is hit.
template <int MODE>
__global__ void __launch_bounds__(32) sbrt_inv_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P) {
  // list entry = {q, last occurrence + 256 (255 - symbol while never seen), symbol}.  A hit entry gets the newest occurrence, so a
  // tie in q always goes to it: "key > new key" is a comparison of the q fields alone.
  __shared__ u32 KQ[256 + 32], KT[256 + 32];
  __shared__ u8 r2s[256 + 32];
  const int lane = threadIdx.x, b = blockIdx.x;
  KzgBlock& B = blocks[b];
  int* res = P.result + 2 * b;
  if (lane == 0) { res[0] = 0; res[1] = 0; }
  if (B.status != 0 || !P.enabled[b]) return;
  const int count = B.curLen;
  const u8* __restrict__ src = B.cur;
  u8* __restrict__ dst = B.alt;
  if (count > min(kzg_dst_limit(B, P.dstLimit[b]), B.cap)) return;
  constexpr int m1 = (MODE == 3) ? 0 : -1, m2 = (MODE == 1) ? 0 : -1, s = (MODE == 2) ? 1 : 0;      // (compile-time: the step is instruction bound)
  for (int i = lane; i < 256; i += 32) { KQ[i] = 0; KT[i] = (u32)(255 - i); r2s[i] = (u8)i; }
  __syncwarp();
  u32 kq = 0, kt = (u32)(255 - lane);   // rank `lane`
  int sy = lane;
  for (int base = 0; base < count; base += 32) {
    const int nIn = min(32, count - base);
    const int mine = (lane < nIn) ? (int)src[base + lane] : 0;
    int outv = 0;
    const u32 deep = __ballot_sync(0xFFFFFFFFu, mine >= 32);
    for (int t = 0; t < nIn; t++) {
      const int i = base + t;
      const int r = __shfl_sync(0xFFFFFFFFu, mine, t);
      if (!((deep >> t) & 1u)) {
        const int c = __shfl_sync(0xFFFFFFFFu, sy, r);
        const u32 low = __shfl_sync(0xFFFFFFFFu, kt, r);
        const u32 upQ = __shfl_up_sync(0xFFFFFFFFu, kq, 1), upT = __shfl_up_sync(0xFFFFFFFFu, kt, 1);
        const int upS = __shfl_up_sync(0xFFFFFFFFu, sy, 1);
        if (lane == t) outv = c;
        const int pOld = (low >= 256u) ? (int)(low - 256u) : 0;
        const u32 qc = (u32)(((i & m1) + (pOld & m2)) >> s);
        // entries 0..r-1: those with a greater q keep their place (q descends with the rank), the rest move down one slot
        const int rn = __popc(__ballot_sync(0xFFFFFFFFu, (lane < r) && (kq > qc)));
        const bool mv = (lane > rn) && (lane <= r), at = (lane == rn);      // (selects, not branches: the lanes must not diverge here)
        kq = mv ? upQ : (at ? qc : kq);
        kt = mv ? upT : (at ? (u32)(i + 256) : kt);
        sy = mv ? upS : (at ? c : sy);
      } else {
        // deep rank: through the shared-memory list (the registers are its first 32 entries)
        KQ[lane] = kq; KT[lane] = kt; r2s[lane] = (u8)sy;
        __syncwarp();
        const int c = r2s[r];
        const u32 low = KT[r];
        if (lane == t) outv = c;
        const int pOld = (low >= 256u) ? (int)(low - 256u) : 0;
        const u32 qc = (u32)(((i & m1) + (pOld & m2)) >> s);
        int rn = 0;
        for (int lo = 0; lo < r; lo += 32) {                  // q descends with the rank: count the entries above the new key
          const int k = lo + lane;
          const u32 m = __ballot_sync(0xFFFFFFFFu, (k < r) && (KQ[k] > qc));
          rn += __popc(m);
          if (m != 0xFFFFFFFFu) break;
        }
        for (int top = r - 1; top >= rn; top -= 32) {         // entries [rn, r) move down one slot, highest chunk first
          const int k = top - lane;
          const bool on = k >= rn;
          u32 q2 = 0, t2 = 0; int s2 = 0;
          if (on) { q2 = KQ[k]; t2 = KT[k]; s2 = r2s[k]; }
          __syncwarp();
          if (on) { KQ[k + 1] = q2; KT[k + 1] = t2; r2s[k + 1] = (u8)s2; }
          __syncwarp();
        }
        if (lane == 0) { KQ[rn] = qc; KT[rn] = (u32)(i + 256); r2s[rn] = (u8)c; }
        __syncwarp();
        kq = KQ[lane]; kt = KT[lane]; sy = r2s[lane];
      }
    }
    if (lane < nIn) dst[base + lane] = (u8)outv;
  }
  if (lane == 0) { res[0] = 1; res[1] = count; }
}

// ================================================================================================================
// SRT (SRT.java:73-168 forward, 178-257 inverse)
// ================================================================================================================
struct SrtSmem { i32 freqs[256]; i32 buckets[256]; i32 bucketEnds[256]; i32 firstPos[256]; u8 r2s[256]; u8 s2r[256]; u8 symbols[256]; int nbSymbols; int hdrLen; };

// SRT.preprocess (:266-302): present symbols ordered by (freq desc, symbol asc) — a strict total order, so any sort gives it
__device__ void srt_preprocess(SrtSmem& S, int lane) {
  if (lane == 0) {
    int n = 0;
    for (int i = 0; i < 256; i++) if (S.freqs[i] > 0) S.symbols[n++] = (u8)i;
    S.nbSymbols = n;
    for (int i = 1; i < n; i++) {           // insertion sort
      const int t = S.symbols[i];
      int k = i - 1;
      while (k >= 0 && ((S.freqs[S.symbols[k]] < S.freqs[t]) || ((S.freqs[t] == S.freqs[S.symbols[k]]) && (t < S.symbols[k])))) {
        S.symbols[k + 1] = S.symbols[k]; k--;
      }
      S.symbols[k + 1] = (u8)t;
    }
  }
  __syncwarp();
}

// ---- SRT forward, tile-parallel ------------------------------------------------------------------------------------------
// The rank SRT emits for position i is the move-to-front rank of its symbol c in a list that starts in order of first
// appearance (:92-106): the number of symbols whose last occurrence is later than c's, a symbol never seen counting as earliest
// (the first occurrence of the b-th distinct symbol is emitted as b).  The byte goes to slot (occurrences of c before i) of c's
// bucket; buckets are ordered by (frequency desc, symbol asc) (preprocess :270-302).  Three kernels, tiles of 4096 positions:
// per tile and symbol {occurrences, last occurrence}; one CTA per block (a thread per symbol) turns them into every tile's entry
// state {occurrences before the tile, last occurrence before the tile}, the bucket starts and the header; one warp per tile
// replays its positions from that state with the 256 keys in registers (8 per lane).
#define SRT_TILE 4096
__global__ void __launch_bounds__(32) srt_fwd_tile_kernel(const KzgBlock* __restrict__ blocks, KzgXfParams P) {
  __shared__ int cnt[256], last[256];
  const int lane = threadIdx.x, b = blockIdx.y, tile = blockIdx.x;
  const KzgBlock& B = blocks[b];
  if (B.status != 0 || !P.enabled[b]) return;
  const int count = B.curLen;
  const int beg = tile * SRT_TILE;
  if (beg >= count || count + 1024 > B.cap) return;
  const int end = min(beg + SRT_TILE, count);
  const u8* __restrict__ src = B.cur;
  for (int i = lane; i < 256; i += 32) { cnt[i] = 0; last[i] = -1; }
  __syncwarp();
  for (int base = beg; base < end; base += 32) {
    const int p = base + lane;
    const bool on = p < end;
    const int c = on ? (int)src[p] : 256 + lane;
    const u32 peers = __match_any_sync(0xFFFFFFFFu, c);
    if (on && (peers >> lane) == 1u) { cnt[c] += __popc(peers); last[c] = p; }      // highest lane of its symbol group
    __syncwarp();
  }
  int* info = reinterpret_cast<int*>(P.scratch + (i64)b * P.scratchStride) + 512 + (i64)tile * 512;
  for (int i = lane; i < 256; i += 32) { info[i] = cnt[i]; info[256 + i] = last[i]; }
}
__global__ void __launch_bounds__(256) srt_fwd_scan_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P) {
  __shared__ int freqs[256];
  __shared__ int hdrLen;
  const int b = blockIdx.x, sym = threadIdx.x;
  KzgBlock& B = blocks[b];
  int* res = P.result + 2 * b;
  if (sym == 0) { res[0] = 0; res[1] = 0; }
  if (B.status != 0 || !P.enabled[b]) return;
  const int count = B.curLen;
  if (count == 0) { if (sym == 0) res[0] = 1; return; }       // :74-75
  if (count + 1024 > B.cap) return;
  const int nTiles = (count + SRT_TILE - 1) / SRT_TILE;
  int* head = reinterpret_cast<int*>(P.scratch + (i64)b * P.scratchStride);       // [0,256) bucket starts, [256] header length
  int* info = head + 512;
  int occ = 0, lst = -1;
  for (int t0 = 0; t0 < nTiles; t0 += 8) {       // eight tiles' loads in flight, then their entry states replace them
    int c[8], l[8];
    #pragma unroll
    for (int k = 0; k < 8; k++) if (t0 + k < nTiles) { c[k] = info[(i64)(t0 + k) * 512 + sym]; l[k] = info[(i64)(t0 + k) * 512 + 256 + sym]; }
    #pragma unroll
    for (int k = 0; k < 8; k++) if (t0 + k < nTiles) {
      info[(i64)(t0 + k) * 512 + sym] = occ; info[(i64)(t0 + k) * 512 + 256 + sym] = lst;
      occ += c[k]; if (l[k] >= 0) lst = l[k];
    }
  }
  freqs[sym] = occ;
  __syncthreads();
  int start = 0;                                 // bytes in the buckets ordered before this symbol's (preprocess :270-302)
  for (int o = 0; o < 256; o++) { const int f = freqs[o]; if (f > occ || (f == occ && o < sym)) start += f; }
  head[sym] = start;
  if (sym == 0) {
    u8* __restrict__ dst = B.alt;
    int k = 0;                                   // encodeHeader (:312-325)
    for (int i = 0; i < 256; i++) {
      u32 f = (u32)freqs[i];
      while (f >= 128) { dst[k++] = (u8)(0x80 | f); f >>= 7; }
      dst[k++] = (u8)f;
    }
    head[256] = k;
    res[0] = 1; res[1] = k + count;
  }
}
__global__ void __launch_bounds__(32) srt_fwd_rank_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P) {
  __shared__ u32 K[256];
  __shared__ int slot[256];
  const int lane = threadIdx.x, b = blockIdx.y, tile = blockIdx.x;
  KzgBlock& B = blocks[b];
  if (B.status != 0 || !P.enabled[b]) return;
  const int count = B.curLen;
  const int beg = tile * SRT_TILE;
  if (beg >= count || count + 1024 > B.cap) return;
  const int end = min(beg + SRT_TILE, count);
  const u8* __restrict__ src = B.cur;
  const int* head = reinterpret_cast<const int*>(P.scratch + (i64)b * P.scratchStride);
  const int* info = head + 512 + (i64)tile * 512;
  u8* __restrict__ out = B.alt + head[256];
  u32 kr[8];
  #pragma unroll
  for (int k = 0; k < 8; k++) {
    const int d = lane + 32 * k;
    kr[k] = (u32)(info[256 + d] + 1);            // key = last occurrence + 1, 0 = never seen
    K[d] = kr[k];
    slot[d] = head[d] + info[d];                 // where this symbol's next rank goes
  }
  __syncwarp();
  for (int base = beg; base < end; base += 32) {
    const int nIn = min(32, end - base);
    const int mine = (lane < nIn) ? (int)src[base + lane] : 0;
    for (int t = 0; t < nIn; t++) {
      const int c = __shfl_sync(0xFFFFFFFFu, mine, t);
      const u32 kc = K[c];
      int cnt = 0;
      #pragma unroll
      for (int k = 0; k < 8; k++) cnt += (kr[k] > kc) ? 1 : 0;
      const int r = (int)__reduce_add_sync(0xFFFFFFFFu, (unsigned)cnt);
      const u32 nk = (u32)(base + t + 1);
      __syncwarp();
      if (lane == 0) { K[c] = nk; const int sl = slot[c]; slot[c] = sl + 1; out[sl] = (u8)r; }
      if (lane == (c & 31)) {
        #pragma unroll
        for (int k = 0; k < 8; k++) if (k == (c >> 5)) kr[k] = nk;
      }
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(32) srt_inverse_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P) {
  __shared__ SrtSmem S;
  const int lane = threadIdx.x, b = blockIdx.x;
  KzgBlock& B = blocks[b];
  int* res = P.result + 2 * b;
  if (lane == 0) { res[0] = 0; res[1] = 0; }
  if (B.status != 0 || !P.enabled[b]) return;
  const int length = B.curLen;
  const u8* __restrict__ src = B.cur;
  u8* __restrict__ dst = B.alt;
  if (lane == 0) {                          // decodeHeader (:335-353)
    int k = 0; bool bad = false;
    for (int i = 0; i < 256; i++) {
      if (k >= length) { bad = true; break; }
      int val = src[k++];
      int r = val & 0x7F, shift = 7;
      while (val >= 128) {
        if (k >= length) { bad = true; break; }
        val = src[k++];
        r |= ((val & 0x7F) << shift);
        if (shift > 21) break;
        shift += 7;
      }
      S.freqs[i] = r;
    }
    S.hdrLen = bad ? -1 : k;
  }
  __syncwarp();
  const int hdr = S.hdrLen;
  if (hdr < 0) return;
  const int count = length - hdr;
  if (count > min(kzg_dst_limit(B, P.dstLimit[b]), B.cap) || count < 0) return;
  const u8* __restrict__ in = src + hdr;
  for (int i = lane; i < 256; i += 32) { S.r2s[i] = 0; S.buckets[i] = 0; S.bucketEnds[i] = 0; if (S.freqs[i] < 0) S.freqs[i] = 0; }
  __syncwarp();
  srt_preprocess(S, lane);
  int bad = 0;
  if (lane == 0) {
    int bucketPos = 0;
    for (int i = 0; i < S.nbSymbols; i++) {
      const int c = S.symbols[i];
      if ((hdr + bucketPos < 0) || (hdr + bucketPos >= length)) { bad = 1; break; }   // :206-207
      S.r2s[in[bucketPos]] = (u8)c;
      S.buckets[c] = bucketPos + 1;
      bucketPos += S.freqs[c];
      S.bucketEnds[c] = bucketPos;
    }
  }
  bad = __shfl_sync(0xFFFFFFFFu, bad, 0);
  __syncwarp();
  if (bad) return;
  int nbSymbols = S.nbSymbols;
  int c = S.r2s[0];
  for (int base = 0; base < count; base += 32) {
    const int nOut = min(32, count - base);
    int outv = 0;
    for (int t = 0; t < nOut; t++) {
      if (lane == t) outv = c;
      const int bk = S.buckets[c], be = S.bucketEnds[c];
      if (bk < be) {
        if (bk >= count) { bad = 1; break; }
        const int r = in[bk];
        __syncwarp();
        if (lane == 0) S.buckets[c] = bk + 1;
        if (r != 0) {
          // for (s < r) r2s[s] = r2s[s+1]; r2s[r] = c   (entries (0, r] move down one slot)
          for (int lo = 0; lo < r; lo += 32) {
            const int k = lo + lane;
            u8 v = 0;
            if (k < r) v = S.r2s[k + 1];
            __syncwarp();
            if (k < r) S.r2s[k] = v;
            __syncwarp();
          }
          if (lane == 0) S.r2s[r] = (u8)c;
          __syncwarp();
          c = S.r2s[0];
        }
        __syncwarp();
      } else {
        if (nbSymbols == 1) continue;
        nbSymbols--;
        for (int lo = 0; lo < nbSymbols; lo += 32) {
          const int k = lo + lane;
          u8 v = 0;
          if (k < nbSymbols) v = S.r2s[k + 1];
          __syncwarp();
          if (k < nbSymbols) S.r2s[k] = v;
          __syncwarp();
        }
        c = S.r2s[0];
      }
    }
    if (bad) break;
    if (lane < nOut) dst[base + lane] = (u8)outv;
  }
  if (lane == 0 && !bad) { res[0] = 1; res[1] = count; }
}

// ---- launchers -------------------------------------------------------------------------------------------------------
void kzg_small_scratch(int type, i32 maxLen, bool forward, size_t* perBlockBytes, size_t*) {
  // SRT forward: per tile of 4096 positions and symbol {occurrences, last occurrence} (2 KiB), 2 KiB of bucket starts in front
  if (forward && type == KZG_T_SRT) *perBlockBytes = std::max(*perBlockBytes, ((size_t)maxLen / SRT_TILE + 3) * 2048);
  // ZRLT: a 32-byte summary per tile of 4096 input bytes (+ the inverse's 0xFF counts)
  if (type == KZG_T_ZRLT) *perBlockBytes = std::max(*perBlockBytes, ((size_t)maxLen / ZR_TILE + 2) * (sizeof(ZrSum) + 4) + 64);
  // SBRT forward: per tile of 4096 positions the last two occurrences of every symbol (2 KiB), then the tile's entry state
  if (forward && (type == KZG_T_RANK || type == KZG_T_MTFT)) *perBlockBytes = std::max(*perBlockBytes, ((size_t)maxLen / SBRT_TILE + 2) * 2048);
}

int kzg_zrlt_launch(cudaStream_t s, bool forward, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, i32 maxLen) {
  const int tiles = std::max(1, (maxLen + ZR_TILE - 1) / ZR_TILE);
  if (((size_t)maxLen / ZR_TILE + 2) * (sizeof(ZrSum) + 4) + 64 > (size_t)P.scratchStride) { kzg_set_error("zrlt: scratch pool too small"); return -KZG_ERR_CREATE_CODEC; }
  const dim3 grid(tiles, nBlocks);
  if (forward) {
    KZG_PROF("zrlt_fwd_sum_kernel", s, (zrlt_fwd_sum_kernel<<<grid, 32, 0, s>>>(d_blocks, P)));
    KZG_PROF("zrlt_fwd_scan_kernel", s, (zrlt_fwd_scan_kernel<<<nBlocks, 32, 0, s>>>(d_blocks, P)));
    KZG_PROF("zrlt_fwd_emit_kernel", s, (zrlt_fwd_emit_kernel<<<grid, 32, 0, s>>>(d_blocks, P)));
    kzg_count_launch(3);
  } else {
    KZG_PROF("zrlt_inv_ff_kernel", s, (zrlt_inv_ff_kernel<<<grid, 32, 0, s>>>(d_blocks, P, tiles + 1)));
    KZG_PROF("zrlt_inv_sum_kernel", s, (zrlt_inv_sum_kernel<<<grid, 32, 0, s>>>(d_blocks, P, tiles + 1)));
    KZG_PROF("zrlt_inv_scan_kernel", s, (zrlt_inv_scan_kernel<<<nBlocks, 32, 0, s>>>(d_blocks, P)));
    KZG_PROF("zrlt_inv_emit_kernel", s, (zrlt_inv_emit_kernel<<<grid, 32, 0, s>>>(d_blocks, P, tiles + 1)));
    kzg_count_launch(4);
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}
int kzg_sbrt_launch(cudaStream_t s, bool forward, int mode, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, i32 maxLen) {
  if (forward) {
    const int tiles = (maxLen + SBRT_TILE - 1) / SBRT_TILE;
    if (((size_t)maxLen / SBRT_TILE + 2) * 2048 > (size_t)P.scratchStride) { kzg_set_error("sbrt: scratch pool too small"); return -KZG_ERR_CREATE_CODEC; }
    KZG_PROF("sbrt_fwd_last2_kernel", s, (sbrt_fwd_last2_kernel<<<dim3(tiles, nBlocks), 32, 0, s>>>(d_blocks, P)));
    KZG_PROF("sbrt_fwd_scan_kernel", s, (sbrt_fwd_scan_kernel<<<nBlocks, 256, 0, s>>>(d_blocks, P)));
    KZG_PROF("sbrt_fwd_rank_kernel", s, (sbrt_fwd_rank_kernel<<<dim3(tiles, nBlocks), 32, 0, s>>>(d_blocks, P, mode)));
    kzg_count_launch(2);
  } else if (mode == 1) KZG_PROF("sbrt_inv_kernel", s, (sbrt_inv_kernel<1><<<nBlocks, 32, 0, s>>>(d_blocks, P)));
  else if (mode == 2) KZG_PROF("sbrt_inv_kernel", s, (sbrt_inv_kernel<2><<<nBlocks, 32, 0, s>>>(d_blocks, P)));
  else KZG_PROF("sbrt_inv_kernel", s, (sbrt_inv_kernel<3><<<nBlocks, 32, 0, s>>>(d_blocks, P)));
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(1);
  return 0;
}
int kzg_srt_launch(cudaStream_t s, bool forward, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, i32 maxLen) {
  if (forward) {
    const int tiles = std::max(1, (maxLen + SRT_TILE - 1) / SRT_TILE);
    if (((size_t)maxLen / SRT_TILE + 3) * 2048 > (size_t)P.scratchStride) { kzg_set_error("srt: scratch pool too small"); return -KZG_ERR_CREATE_CODEC; }
    KZG_PROF("srt_fwd_tile_kernel", s, (srt_fwd_tile_kernel<<<dim3(tiles, nBlocks), 32, 0, s>>>(d_blocks, P)));
    KZG_PROF("srt_fwd_scan_kernel", s, (srt_fwd_scan_kernel<<<nBlocks, 256, 0, s>>>(d_blocks, P)));
    KZG_PROF("srt_fwd_rank_kernel", s, (srt_fwd_rank_kernel<<<dim3(tiles, nBlocks), 32, 0, s>>>(d_blocks, P)));
    kzg_count_launch(2);
  } else KZG_PROF("srt_inverse_kernel", s, (srt_inverse_kernel<<<nBlocks, 32, 0, s>>>(d_blocks, P)));
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(1);
  return 0;
}
