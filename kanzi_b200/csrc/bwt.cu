// bwt.cu — BWT forward / inverse and BWTBlockCodec (sm_100a).
//
// Replaces K/transform/BWT.java, BWTBlockCodec.java and the DivSufSort.java suffix sorter (SURVEY.md §8
// rows a10-a12).  The BWT is mathematically defined (Appendix B-5), so the serial induced sort is replaced
// by an on-device radix suffix sort: ranks from the first 3 bytes, then prefix doubling — every round is
// an LSD radix sort of the suffix indices on (rank[i], rank[i+h]) (8-bit digits, stable, per-warp tiles
// ranked with __match_any) followed by a flag/scan re-ranking.  All blocks of a batch advance together
// (grid.y = block); rounds stop when every block has n distinct ranks.
// Inverse: stable counting sort of the last column (same radix pass) gives the LF links; the n-step
// pointer chase of BWT.inverseMergeTPSI / inverseBiPSIv2 (8 chains of n/8 dependent loads) becomes a
// splitter list ranking: every 256-th row is a chain head, all heads walk to the next head in parallel,
// the head list is ranked by pointer jumping, a second parallel walk writes the bytes in place.
#include "kzg_common.cuh"
#include "kzg_transforms.cuh"
#include "kzg_xf_kernels.cuh"
#include "kzg_stages.cuh"
#include <vector>
#include <algorithm>

#define BW_WT 4096                 // elements per warp tile
#define BW_WARPS 8                 // warps per CTA
#define BW_SPLIT 256               // inverse: one chain head every BW_SPLIT rows

struct BwBlock {                   // per-block device state of the suffix sort / inverse
  const u8* T; u8* out; i32 n; i32 active; i32 groups; i32 pad;
  u32 *sa, *sa2, *rk, *rk2, *hist, *offs, *tsum;
  i32 pidx[8];
};

__device__ __forceinline__ int bw_tiles(int n) { return (n + BW_WT - 1) / BW_WT; }
static inline size_t rnd256(size_t v) { return (v + 255) / 256 * 256; }

// ---- initial ranks: first three bytes, (byte+1) per position, 0 past the end (shorter suffix sorts first) ----------
__global__ void bw_init_kernel(BwBlock* __restrict__ bb) {
  BwBlock& B = bb[blockIdx.y];
  if (!B.active) return;
  const int n = B.n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const u32 b0 = B.T[i] + 1u, b1 = (i + 1 < n) ? B.T[i + 1] + 1u : 0u, b2 = (i + 2 < n) ? B.T[i + 2] + 1u : 0u;
    B.rk[i] = (b0 * 257u + b1) * 257u + b2 + 1u;      // >= 1; 0 is reserved for "past the end"
    B.sa[i] = (u32)i;
  }
}

// key of suffix s for a radix pass: which = 0 -> rank[s + h] (0 past the end), 1 -> rank[s], 2 -> the byte T[s] (inverse BWT)
__device__ __forceinline__ u32 bw_key(const BwBlock& B, u32 s, int which, int h) {
  if (which == 1) return B.rk[s];
  if (which == 0) { const u32 t = s + (u32)h; return (t < (u32)B.n) ? B.rk[t] : 0u; }
  return B.T[s];
}

// ---- radix pass 1/3: per-tile digit histogram, hist[digit][tile] ----------------------------------------------------------
__global__ void __launch_bounds__(32 * BW_WARPS) bw_hist_kernel(BwBlock* __restrict__ bb, int which, int h, int shift) {
  __shared__ u32 cnt[BW_WARPS][256];
  BwBlock& B = bb[blockIdx.y];
  if (!B.active) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x * BW_WARPS + warp;
  const int nT = bw_tiles(B.n);
  if (tile >= nT) return;
  for (int i = lane; i < 256; i += 32) cnt[warp][i] = 0;
  __syncwarp();
  const int beg = tile * BW_WT, end = min(beg + BW_WT, B.n);
  for (int i = beg + lane; i < end; i += 32) {
    const u32 s = (which == 2) ? (u32)i : B.sa[i];
    atomicAdd(&cnt[warp][(bw_key(B, s, which, h) >> shift) & 255], 1u);
  }
  __syncwarp();
  for (int d = lane; d < 256; d += 32) B.hist[(size_t)d * nT + tile] = cnt[warp][d];
}

// ---- radix pass 2/3: exclusive scan over (digit-major, tile-minor); one CTA per block ------------------------------------
__global__ void __launch_bounds__(1024) bw_scan_kernel(BwBlock* __restrict__ bb) {
  __shared__ u32 wsum[32];
  __shared__ u32 carry;
  BwBlock& B = bb[blockIdx.x];
  if (!B.active) return;
  const int total = 256 * bw_tiles(B.n);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < total; base += 1024) {
    const int i = base + threadIdx.x;
    const u32 v = (i < total) ? B.hist[i] : 0u;
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
    if (i < total) B.offs[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
}

// ---- radix pass 3/3: stable scatter.  mode 0: sa -> sa2 (suffix indices).  mode 1 (inverse BWT): row i -> next/fcol -----
__global__ void __launch_bounds__(32 * BW_WARPS) bw_scatter_kernel(BwBlock* __restrict__ bb, int which, int h, int shift, int mode, int pIdxArg) {
  __shared__ u32 pos[BW_WARPS][256];
  BwBlock& B = bb[blockIdx.y];
  if (!B.active) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x * BW_WARPS + warp;
  const int nT = bw_tiles(B.n);
  if (tile >= nT) return;
  for (int d = lane; d < 256; d += 32) pos[warp][d] = B.offs[(size_t)d * nT + tile];
  __syncwarp();
  const int beg = tile * BW_WT, end = min(beg + BW_WT, B.n);
  const u32 lower = (1u << lane) - 1;
  for (int base = beg; base < end; base += 32) {
    const int i = base + lane;
    const bool on = i < end;
    u32 s = 0; int d = 256 + lane;
    if (on) { s = (which == 2) ? (u32)i : B.sa[i]; d = (int)((bw_key(B, s, which, h) >> shift) & 255); }
    const u32 peers = __match_any_sync(0xFFFFFFFFu, d);
    if (on) {
      const u32 p = pos[warp][d] + __popc(peers & lower);
      if (mode == 0) B.sa2[p] = s;
      else {
        // LF links (BWT.java:268-287): row i of the last column lands at sorted row p; following it yields
        // text position "i-1 if i < pIdx else i" (the row whose sorted slot continues the text); row 0 ends the walk
        const int pIdx = B.pidx[0];
        B.sa2[p] = (i == 0) ? 0xFFFFFFFFu : (u32)((i < pIdx) ? i - 1 : i);
        ((u8*)B.rk2)[p] = (u8)d;
      }
    }
    __syncwarp();
    if (on && (peers >> lane) <= 1u) pos[warp][d] += __popc(peers);     // highest lane of each digit group advances the counter
    __syncwarp();
  }
}

// ---- re-ranking: flags, device-wide inclusive scan (3 kernels), scatter of the new dense ranks -------------------------------
__global__ void bw_flag_kernel(BwBlock* __restrict__ bb, int h) {
  BwBlock& B = bb[blockIdx.y];
  if (!B.active) return;
  const int n = B.n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    u32 f = 1;
    if (i > 0) {
      const u32 a = B.sa[i], p = B.sa[i - 1];
      f = (B.rk[a] != B.rk[p]) || (bw_key(B, a, 0, h) != bw_key(B, p, 0, h));
    }
    B.sa2[i] = f;
  }
}
// tile sums of sa2 (flags) -> tsum[tile]; tile = 4096 elements per CTA of 256 threads
__global__ void __launch_bounds__(256) bw_tilesum_kernel(BwBlock* __restrict__ bb) {
  __shared__ u32 ws[8];
  BwBlock& B = bb[blockIdx.y];
  if (!B.active) return;
  const int nT = bw_tiles(B.n);
  if ((int)blockIdx.x >= nT) return;
  const int beg = blockIdx.x * BW_WT, end = min(beg + BW_WT, B.n);
  u32 s = 0;
  for (int i = beg + threadIdx.x; i < end; i += 256) s += B.sa2[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { u32 t = 0; for (int w = 0; w < 8; w++) t += ws[w]; B.tsum[blockIdx.x] = t; }
}
// exclusive scan of tsum (one CTA per block), total -> groups
__global__ void __launch_bounds__(1024) bw_tilescan_kernel(BwBlock* __restrict__ bb) {
  __shared__ u32 wsum[32];
  __shared__ u32 carry;
  BwBlock& B = bb[blockIdx.x];
  if (!B.active) return;
  const int total = bw_tiles(B.n);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < total; base += 1024) {
    const int i = base + threadIdx.x;
    const u32 v = (i < total) ? B.tsum[i] : 0u;
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
    if (i < total) B.tsum[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) B.groups = (i32)carry;
}
// new rank of sa[i] = inclusive prefix sum of the flags (dense, 1..groups); one warp-tile walks its 4096 elements in order
__global__ void __launch_bounds__(32 * BW_WARPS) bw_rerank_kernel(BwBlock* __restrict__ bb) {
  BwBlock& B = bb[blockIdx.y];
  if (!B.active) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x * BW_WARPS + warp;
  if (tile >= bw_tiles(B.n)) return;
  const int beg = tile * BW_WT, end = min(beg + BW_WT, B.n);
  u32 run = B.tsum[tile];
  for (int base = beg; base < end; base += 32) {
    const int i = base + lane;
    const u32 f = (i < end) ? B.sa2[i] : 0u;
    u32 incl = f;
    for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
    if (i < end) B.rk2[B.sa[i]] = run + incl;
    run += __shfl_sync(0xFFFFFFFFu, incl, 31);
  }
}
__global__ void bw_swap_kernel(BwBlock* __restrict__ bb, int nBlocks, int what) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nBlocks || !bb[b].active) return;
  if (what == 0) { u32* t = bb[b].sa; bb[b].sa = bb[b].sa2; bb[b].sa2 = t; }
  else { u32* t = bb[b].rk; bb[b].rk = bb[b].rk2; bb[b].rk2 = t; if (bb[b].groups >= bb[b].n) bb[b].active = 0; }
}

// ---- output of the forward transform (DivSufSort.computeBWT :216-226 and the primary indexes, Appendix B-5) --------------------
__global__ void bw_emit_kernel(BwBlock* __restrict__ bb) {
  BwBlock& B = bb[blockIdx.y];
  const int n = B.n;
  if (n < 2) return;
  const int pIdx = (int)B.rk[0] - 1;        // rank of suffix 0 (ranks are dense 1..n once the sort is done)
  const int chunks = (n < 256) ? 1 : 8;
  const int st = n / chunks;
  const int step = (st * chunks != n) ? st + 1 : st;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const u32 s = B.sa[i];
    if (s == 0) { B.out[0] = B.T[n - 1]; B.pidx[0] = pIdx + 1; continue; }
    if ((int)(s % (u32)step) == 0 && (int)(s / (u32)step) < 8) B.pidx[s / step] = i + 1;
    B.out[(i < pIdx) ? i + 1 : i] = B.T[s - 1];
  }
}

// ---- inverse: splitter walk ------------------------------------------------------------------------------------------------------------
// heads: rows t with t % BW_SPLIT == 0, plus the start row t0 (id = nHeadsRegular).  next[] = sa2, fcol = (u8*)rk2.
__global__ void bw_walk1_kernel(BwBlock* __restrict__ bb) {
  BwBlock& B = bb[blockIdx.y];
  if (!B.active) return;
  const int n = B.n;
  const int nReg = (n + BW_SPLIT - 1) / BW_SPLIT;
  const int t0 = B.pidx[0] - 1;
  u32* segLen = B.hist; u32* segNext = B.offs;
  for (int hId = blockIdx.x * blockDim.x + threadIdx.x; hId <= nReg; hId += gridDim.x * blockDim.x) {
    u32 t = (hId == nReg) ? (u32)t0 : (u32)hId * BW_SPLIT;
    if (hId < nReg && (int)t == t0) { segLen[hId] = 0; segNext[hId] = 0xFFFFFFFFu; continue; }   // served by the start head
    u32 len = 0, nx = 0xFFFFFFFFu;
    while (true) {
      len++;
      const u32 nt = B.sa2[t];
      if (nt == 0xFFFFFFFFu || nt >= (u32)n) break;                 // end of the text (row 0 of the last column) or corrupt link
      if (len >= (u32)n) break;
      if (nt == (u32)t0) { nx = (u32)nReg; break; }
      if ((nt % BW_SPLIT) == 0) { nx = nt / BW_SPLIT; break; }
      t = nt;
    }
    segLen[hId] = len; segNext[hId] = nx;
  }
}
// list ranking of the heads by pointer jumping: dist[h] = bytes from h's segment start to the end of the text; one CTA per block
__global__ void __launch_bounds__(1024) bw_rank_kernel(BwBlock* __restrict__ bb) {
  BwBlock& B = bb[blockIdx.x];
  if (!B.active) return;
  const int n = B.n;
  const int nH = (n + BW_SPLIT - 1) / BW_SPLIT + 1;
  u32* dist = B.hist; u32* nxt = B.offs;            // in place: dist starts as segLen
  u32* dist2 = B.tsum; u32* nxt2 = B.tsum + nH;
  u32 *d0 = dist, *n0 = nxt, *d1 = dist2, *n1 = nxt2;
  for (int round = 0; round < 32; round++) {
    for (int h = threadIdx.x; h < nH; h += 1024) {
      const u32 nx = n0[h];
      if (nx == 0xFFFFFFFFu) { d1[h] = d0[h]; n1[h] = nx; }
      else { d1[h] = d0[h] + d0[nx]; n1[h] = n0[nx]; }
    }
    __syncthreads();
    u32* t = d0; d0 = d1; d1 = t; t = n0; n0 = n1; n1 = t;
    if ((1 << round) >= nH) break;
  }
  __syncthreads();
  // result sits in d0; copy to dist (= B.hist) if needed
  if (d0 != dist) { for (int h = threadIdx.x; h < nH; h += 1024) dist[h] = d0[h]; }
}
__global__ void bw_walk2_kernel(BwBlock* __restrict__ bb, const u32* __restrict__ segLenCopy, int segStride) {
  BwBlock& B = bb[blockIdx.y];
  if (!B.active) return;
  const int n = B.n;
  const int nReg = (n + BW_SPLIT - 1) / BW_SPLIT;
  const int t0 = B.pidx[0] - 1;
  const u32* dist = B.hist;
  const u32* segLen = segLenCopy + (size_t)blockIdx.y * segStride;
  const u8* fcol = (const u8*)B.rk2;
  const u32 total = dist[nReg];                     // bytes reachable from the start row (n for a valid BWT)
  for (int hId = blockIdx.x * blockDim.x + threadIdx.x; hId <= nReg; hId += gridDim.x * blockDim.x) {
    const u32 len = segLen[hId];
    if (len == 0) continue;
    const u32 d = dist[hId];
    if (d > total) continue;                        // not on the path of the start row
    u32 o = total - d;                              // output offset of this segment
    u32 t = (hId == nReg) ? (u32)t0 : (u32)hId * BW_SPLIT;
    for (u32 k = 0; k < len && o < (u32)n; k++, o++) {
      B.out[o] = fcol[t];
      t = B.sa2[t];
      if (t >= (u32)n) break;
    }
  }
}

// ---- BWTBlockCodec header (BWTBlockCodec.java:98-127 forward, 152-180 inverse) --------------------------------------------------
__global__ void bw_header_kernel(KzgBlock* __restrict__ blocks, BwBlock* __restrict__ bb, int nBlocks, int* __restrict__ result, int forward) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nBlocks) return;
  KzgBlock& K = blocks[b];
  BwBlock& B = bb[b];
  if (!B.active && forward != 2) {}
  if (B.n <= 0) return;
  if (forward == 1) {
    const int n = B.n;
    int lg = ilog2((u32)n); if ((n & (n - 1)) != 0) lg++;
    const int pIndexSize = (lg + 7) >> 3;
    const int chunks = (n < 256) ? 1 : 8;
    u8* o = K.alt;
    int idx = 1;
    for (int i = 0; i < chunks; i++) {
      const int pi = ((n == 1) ? 0 : B.pidx[i]) - 1;
      for (int sh = (pIndexSize - 1) << 3; sh >= 0; sh -= 8) o[idx++] = (u8)(pi >> sh);
    }
    o[0] = (u8)((ilog2((u32)chunks) << 2) | (pIndexSize - 1));
    result[2 * b] = 1; result[2 * b + 1] = idx + n;
  }
}

// ================================================================================================================================
// host side
// ================================================================================================================================
static size_t bw_aux_words(i32 maxLen) {
  const size_t n = (size_t)maxLen + 64;
  const size_t nT = (n + BW_WT - 1) / BW_WT;
  return 4 * n + 2 * 256 * nT + 4 * nT + 4 * (n / BW_SPLIT + 8) + 256;
}
void kzg_bwt_scratch(i32 maxLen, bool, size_t* perBlockBytes, size_t* aux32) {
  *aux32 = std::max(*aux32, bw_aux_words(maxLen));
  *perBlockBytes = std::max(*perBlockBytes, (size_t)sizeof(BwBlock) + 256 + 4 * ((size_t)maxLen / BW_SPLIT + 8));
}

struct BwHostPlan { std::vector<BwBlock> hb; int maxN = 0; };

static void bw_layout(BwBlock& B, u32* aux, i32 maxLen) {
  const size_t n = (size_t)maxLen + 64;
  const size_t nT = (n + BW_WT - 1) / BW_WT;
  B.sa = aux; B.sa2 = aux + n; B.rk = aux + 2 * n; B.rk2 = aux + 3 * n;
  B.hist = aux + 4 * n; B.offs = B.hist + 256 * nT; B.tsum = B.offs + 256 * nT;
}

static int bw_radix_pass(cudaStream_t s, BwBlock* dbb, int nBlocks, int maxN, int which, int h, int shift, int mode) {
  const int nT = (maxN + BW_WT - 1) / BW_WT;
  dim3 grid((nT + BW_WARPS - 1) / BW_WARPS, nBlocks);
  bw_hist_kernel<<<grid, 32 * BW_WARPS, 0, s>>>(dbb, which, h, shift);
  bw_scan_kernel<<<nBlocks, 1024, 0, s>>>(dbb);
  bw_scatter_kernel<<<grid, 32 * BW_WARPS, 0, s>>>(dbb, which, h, shift, mode, 0);
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(3);
  if (mode == 0) { bw_swap_kernel<<<(nBlocks + 63) / 64, 64, 0, s>>>(dbb, nBlocks, 0); kzg_count_launch(1); }
  return 0;
}

// suffix sort + BWT output for every active block in hb (device copies in dbb)
static int bw_forward_run(cudaStream_t s, std::vector<BwBlock>& hb, BwBlock* dbb, int maxN) {
  const int nBlocks = (int)hb.size();
  const int nT = (maxN + BW_WT - 1) / BW_WT;
  dim3 gridE(std::min((maxN + 255) / 256, 4 * KZG_SM_COUNT), nBlocks);
  CUDA_TRY(cudaMemcpyAsync(dbb, hb.data(), sizeof(BwBlock) * nBlocks, cudaMemcpyHostToDevice, s));
  bw_init_kernel<<<gridE, 256, 0, s>>>(dbb);
  kzg_count_launch(1);
  // initial sort by the 25-bit 3-byte key
  for (int shift = 0; shift < 25; shift += 8) { int r = bw_radix_pass(s, dbb, nBlocks, maxN, 1, 0, shift, 0); if (r < 0) return r; }
  int h = 0;                 // h = 0: re-rank on rk alone (bw_key(.,0,0) adds rank[s+0] == rk[s], harmless)
  int maxGroups = 0;
  for (int round = 0; round < 40; round++) {
    bw_flag_kernel<<<gridE, 256, 0, s>>>(dbb, h);
    bw_tilesum_kernel<<<dim3(nT, nBlocks), 256, 0, s>>>(dbb);
    bw_tilescan_kernel<<<nBlocks, 1024, 0, s>>>(dbb);
    bw_rerank_kernel<<<dim3((nT + BW_WARPS - 1) / BW_WARPS, nBlocks), 32 * BW_WARPS, 0, s>>>(dbb);
    bw_swap_kernel<<<(nBlocks + 63) / 64, 64, 0, s>>>(dbb, nBlocks, 1);
    CUDA_TRY(cudaGetLastError());
    kzg_count_launch(5);
    CUDA_TRY(cudaMemcpyAsync(hb.data(), dbb, sizeof(BwBlock) * nBlocks, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    bool any = false; maxGroups = 0;
    for (auto& B : hb) { if (B.active) { any = true; maxGroups = std::max(maxGroups, B.groups); } }
    if (!any) break;
    h = (h == 0) ? 3 : 2 * h;
    if (h >= 2 * maxN + 8) { kzg_set_error("suffix sort did not converge"); return -KZG_ERR_PROCESS_BLOCK; }
    int bits = 1; while ((1 << bits) <= maxGroups) bits++;
    for (int shift = 0; shift < bits; shift += 8) { int r = bw_radix_pass(s, dbb, nBlocks, maxN, 0, h, shift, 0); if (r < 0) return r; }
    for (int shift = 0; shift < bits; shift += 8) { int r = bw_radix_pass(s, dbb, nBlocks, maxN, 1, h, shift, 0); if (r < 0) return r; }
  }
  bw_emit_kernel<<<gridE, 256, 0, s>>>(dbb);
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(1);
  return 0;
}

static int bw_inverse_run(cudaStream_t s, std::vector<BwBlock>& hb, BwBlock* dbb, int maxN, u32* dSegCopy, int segStride) {
  const int nBlocks = (int)hb.size();
  CUDA_TRY(cudaMemcpyAsync(dbb, hb.data(), sizeof(BwBlock) * nBlocks, cudaMemcpyHostToDevice, s));
  const int nT = (maxN + BW_WT - 1) / BW_WT;
  dim3 grid((nT + BW_WARPS - 1) / BW_WARPS, nBlocks);
  bw_hist_kernel<<<grid, 32 * BW_WARPS, 0, s>>>(dbb, 2, 0, 0);
  bw_scan_kernel<<<nBlocks, 1024, 0, s>>>(dbb);
  bw_scatter_kernel<<<grid, 32 * BW_WARPS, 0, s>>>(dbb, 2, 0, 0, 1, 0);
  const int nH = maxN / BW_SPLIT + 2;
  dim3 gridW(std::max(1, std::min((nH + 127) / 128, 8 * KZG_SM_COUNT)), nBlocks);
  bw_walk1_kernel<<<gridW, 128, 0, s>>>(dbb);
  CUDA_TRY(cudaGetLastError());
  // keep the segment lengths (the ranking overwrites them)
  for (int b = 0; b < nBlocks; b++)
    if (hb[b].active) CUDA_TRY(cudaMemcpyAsync(dSegCopy + (size_t)b * segStride, hb[b].hist, sizeof(u32) * (hb[b].n / BW_SPLIT + 2), cudaMemcpyDeviceToDevice, s));
  bw_rank_kernel<<<nBlocks, 1024, 0, s>>>(dbb);
  bw_walk2_kernel<<<gridW, 128, 0, s>>>(dbb, dSegCopy, segStride);
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(6);
  return 0;
}

// BWTBlockCodec.forward / inverse for a batch.  Block lengths are read back first (one sync): the sort rounds
// need host-side loop control anyway.
int kzg_bwtblock_launch(cudaStream_t s, bool forward, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, i32 maxLen) {
  std::vector<KzgBlock> hk(nBlocks);
  std::vector<u8> en(nBlocks);
  std::vector<int> lim(nBlocks);
  CUDA_TRY(cudaMemcpyAsync(hk.data(), d_blocks, sizeof(KzgBlock) * nBlocks, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(en.data(), P.enabled, nBlocks, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(lim.data(), P.dstLimit, sizeof(int) * nBlocks, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  CUDA_TRY(cudaMemsetAsync(P.result, 0, sizeof(int) * 2 * nBlocks, s));
  const bool asref = (P.flags & KZG_FLAG_BWT_ASREF) != 0;
  std::vector<BwBlock> hb(nBlocks);
  std::vector<int> hres(2 * nBlocks, 0);
  int maxN = 0;
  const size_t auxStride = (size_t)P.aux32Stride;
  for (int b = 0; b < nBlocks; b++) {
    BwBlock& B = hb[b];
    memset(&B, 0, sizeof(B));
    bw_layout(B, (u32*)P.aux32 + (size_t)b * auxStride, maxLen);
    const KzgBlock& K = hk[b];
    if (K.status != 0 || !en[b] || K.curLen <= 0) continue;
    if (forward) {
      const int n = K.curLen;
      int lg = 0; while ((2 << lg) <= n) lg++; if ((n & (n - 1)) != 0) lg++;
      const int pIndexSize = (lg + 7) >> 3;
      if (pIndexSize <= 0 || pIndexSize >= 5) continue;
      const int chunks = (n < 256) ? 1 : 8;
      const int hdr = 1 + chunks * pIndexSize;
      if (hdr + n > K.cap) continue;
      // as written, BWT.forward rejects dst.index + dst.length > dst.array.length (BWT.java:152-156): in the stream
      // path the destination slice is the scratch buffer with length == array.length (DESIGN.md "E-1")
      if (asref) {
        int succeeded = 0; for (int i = 0; i < 8; i++) if (!(K.skipFlags & (1 << (7 - i)))) succeeded++;
        if ((succeeded & 1) == 0) continue;       // destination = `buffer`: always rejected
      }
      B.T = K.cur; B.out = K.alt + hdr; B.n = n; B.active = (n >= 2) ? 1 : 0;
      if (n == 1) { hres[2 * b] = 1; }
      maxN = std::max(maxN, n);
    } else {
      if (asref) continue;                        // BWT.java:211 as written: count > src.length - src.index once the header is consumed
      B.n = -1;                                   // filled after the header parse below
    }
  }
  BwBlock* dbb = (BwBlock*)(P.scratch);           // per-block scratch area is large enough for the descriptors (nBlocks * stride)
  if ((size_t)nBlocks * sizeof(BwBlock) > (size_t)P.scratchStride * nBlocks) return -KZG_ERR_CREATE_CODEC;
  if (forward) {
    if (maxN >= 2) { int r = bw_forward_run(s, hb, dbb, maxN); if (r < 0) return r; }
    else CUDA_TRY(cudaMemcpyAsync(dbb, hb.data(), sizeof(BwBlock) * nBlocks, cudaMemcpyHostToDevice, s));
    // single-byte blocks: BWT copies the byte (BWT.java:172-175)
    for (int b = 0; b < nBlocks; b++) if (hb[b].n == 1) CUDA_TRY(cudaMemcpyAsync(hb[b].out, hb[b].T, 1, cudaMemcpyDeviceToDevice, s));
    bw_header_kernel<<<(nBlocks + 63) / 64, 64, 0, s>>>(d_blocks, dbb, nBlocks, P.result, 1);
    CUDA_TRY(cudaGetLastError());
    kzg_count_launch(1);
    return 0;
  }
  // ---- inverse: parse the headers on the host (they are a handful of bytes per block) ----
  std::vector<u8> hdrBytes(40);
  for (int b = 0; b < nBlocks; b++) {
    BwBlock& B = hb[b];
    const KzgBlock& K = hk[b];
    if (B.n != -1) { B.n = 0; continue; }
    B.n = 0;
    const int blockSize = K.curLen;
    if (blockSize < 1) continue;
    CUDA_TRY(cudaMemcpyAsync(hdrBytes.data(), K.cur, std::min(blockSize, 33), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    const int mode = (int)(int8_t)hdrBytes[0];
    const int logNbChunks = (mode >> 2) & 0x07;
    const int pIndexSize = (mode & 0x03) + 1;
    const int chunks = 1 << logNbChunks;
    const int headerSize = 1 + chunks * pIndexSize;
    if (blockSize < headerSize || chunks > 8) continue;
    const int n = blockSize - headerSize;
    if (chunks != ((n < 256) ? 1 : 8)) continue;
    bool ok = true;
    int idx = 1;
    for (int i = 0; i < chunks; i++) {
      i64 pi = 0;
      for (int sh = (pIndexSize - 1) << 3; sh >= 0; sh -= 8) pi = (pi << 8) | hdrBytes[idx++];
      if (pi >= 0x7FFFFFFFLL) { ok = false; break; }
      B.pidx[i] = (int)pi + 1;
    }
    if (!ok || n <= 0) continue;
    if (n > kzg_dst_limit(K, lim[b]) || n > K.cap) continue;
    if (n == 1) { CUDA_TRY(cudaMemcpyAsync(K.alt, K.cur + headerSize, 1, cudaMemcpyDeviceToDevice, s)); hres[2 * b] = 1; hres[2 * b + 1] = 1; continue; }
    // BWT.inverse guards (BWT.java:260-262, 300-306): primary indexes in range
    if (B.pidx[0] <= 0 || B.pidx[0] > n) continue;
    bool inRange = true;
    if (chunks == 8) for (int i = 0; i < 8; i++) if (B.pidx[i] - 1 < 0 || B.pidx[i] - 1 >= n) inRange = false;
    if (!inRange) continue;
    B.T = K.cur + headerSize; B.out = K.alt; B.n = n; B.active = 1;
    hres[2 * b] = 1; hres[2 * b + 1] = n;
    maxN = std::max(maxN, n);
  }
  if (maxN >= 2) {
    u32* dSegCopy = (u32*)(P.scratch + rnd256((size_t)nBlocks * sizeof(BwBlock)));
    const int segStride = maxLen / BW_SPLIT + 8;
    // compact: run only the active blocks
    int r = bw_inverse_run(s, hb, dbb, maxN, dSegCopy, segStride); if (r < 0) return r;
  }
  CUDA_TRY(cudaMemcpyAsync(P.result, hres.data(), sizeof(int) * 2 * nBlocks, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaStreamSynchronize(s));       // hres is a stack-owned vector
  return 0;
}

// raw BWT of include/kzg.h (what T/test/TestBWT.java drives): host buffers, index 0
int kzg_bwt_raw(cudaStream_t s, bool forward, const u8* src, i32 n, u8* dst, i32* primaryIndexes8) {
  if (n <= 0) return 1;
  if (n == 1) { dst[0] = src[0]; return 1; }
  if (!forward) { if (primaryIndexes8[0] <= 0 || primaryIndexes8[0] > n) return 0; }
  u8 *dT = nullptr, *dOut = nullptr; u32* aux = nullptr; BwBlock* dbb = nullptr; u32* dSeg = nullptr;
  const size_t words = bw_aux_words(n);
  int rc = 1;
  if (cudaMalloc((void**)&dT, (size_t)n + 64) != cudaSuccess || cudaMalloc((void**)&dOut, (size_t)n + 64) != cudaSuccess ||
      cudaMalloc((void**)&aux, words * 4) != cudaSuccess || cudaMalloc((void**)&dbb, sizeof(BwBlock)) != cudaSuccess ||
      cudaMalloc((void**)&dSeg, 4 * ((size_t)n / BW_SPLIT + 8)) != cudaSuccess) {
    cudaGetLastError(); rc = -KZG_ERR_CREATE_CODEC;
  } else {
    std::vector<BwBlock> hb(1);
    memset(&hb[0], 0, sizeof(BwBlock));
    bw_layout(hb[0], aux, n);
    hb[0].T = dT; hb[0].out = dOut; hb[0].n = n; hb[0].active = 1;
    if (!forward) for (int i = 0; i < 8; i++) hb[0].pidx[i] = primaryIndexes8[i];
    cudaMemcpyAsync(dT, src, n, cudaMemcpyHostToDevice, s);
    int r = forward ? bw_forward_run(s, hb, dbb, n) : bw_inverse_run(s, hb, dbb, n, dSeg, n / BW_SPLIT + 8);
    if (r < 0) rc = r;
    else {
      cudaMemcpyAsync(dst, dOut, n, cudaMemcpyDeviceToHost, s);
      if (forward) { cudaMemcpyAsync(hb.data(), dbb, sizeof(BwBlock), cudaMemcpyDeviceToHost, s); }
      if (cudaStreamSynchronize(s) != cudaSuccess) rc = -KZG_ERR_PROCESS_BLOCK;
      if (forward && rc == 1) for (int i = 0; i < 8; i++) primaryIndexes8[i] = hb[0].pidx[i];
    }
  }
  cudaFree(dT); cudaFree(dOut); cudaFree(aux); cudaFree(dbb); cudaFree(dSeg);
  return rc;
}
