// ans.cu — rANS order-0 / order-1 encode and decode kernels (sm_100a).
//
// Replaces K/entropy/ANSRangeEncoder.java and ANSRangeDecoder.java with the
// same bitstream: per chunk a 3-bit logRange, per-context alphabet + frequencies (encodeHeader :211-252),
// varint(byte count), four 32-bit states, bytes.  The unit of parallel work is what the format fixes:
// one chunk (16 KiB for order 0, 4 MiB for order 1) = one coupled 4-state stream.  Four lanes run the
// four states; the shared read/write cursor is resolved per step with a warp ballot (the reference's
// st0..st3 / st3..st0 order becomes a popcount of the lower lanes' renormalisation flags).
// Order 0 keeps all per-chunk tables in shared memory (32 chunks per CTA); order 1 tables (512 KiB per
// chunk) live in global memory and stay L2-resident.
//
// HBM traffic per chunk (algorithmic): encode reads n bytes twice (histogram + code) and writes the
// coded bytes; decode reads coded bytes once and writes n bytes.  Both are bound by the dependent
// table-lookup chain of the rANS recurrence, not by bandwidth.
#include "kzg_common.cuh"
#include "kzg_entropy.cuh"
#include "ans_scan.cuh"

#define ANS_TOP (1 << 15)

// ---- EntropyUtils.normalizeFrequencies (K/entropy/EntropyUtils.java:141-250), one thread ------------------
// freqs: 256 counts (in/out), alphabet: out.  Order-sensitive, restated literally (tie-breaks matter).
__device__ int ans_normalize(u32* freqs, u8* alphabet, int totalFreq, int scale) {
  if (totalFreq == 0) return 0;
  int alphabetSize = 0;
  if (totalFreq == scale) {
    for (int i = 0; i < 256; i++)
      if (freqs[i] != 0) alphabet[alphabetSize++] = (u8)i;
    return alphabetSize;
  }
  int sumScaledFreq = 0, sumFreq = 0, idxMax = 0;
  const bool small = ((u64)totalFreq * (u64)scale) < 0x7FFFFFFFull;
  for (int i = 0; i < 256; i++) {
    const int f = (int)freqs[i];
    if (f == 0) continue;
    int scaledFreq;
    if (small) {
      const u32 sf = (u32)f * (u32)scale;
      scaledFreq = (sf <= (u32)totalFreq) ? 1 : (int)((sf + ((u32)totalFreq >> 1)) / (u32)totalFreq);
    } else {
      const u64 sf = (u64)f * (u64)scale;
      scaledFreq = (sf <= (u64)totalFreq) ? 1 : (int)((sf + ((u64)totalFreq >> 1)) / (u64)totalFreq);
    }
    alphabet[alphabetSize++] = (u8)i;
    sumScaledFreq += scaledFreq;
    freqs[i] = (u32)scaledFreq;
    sumFreq += f;
    if (scaledFreq > (int)freqs[idxMax]) idxMax = i;
    if (sumFreq >= totalFreq) break;
  }
  if (alphabetSize == 0) return 0;
  if (alphabetSize == 1) { freqs[alphabet[0]] = (u32)scale; return 1; }
  if (sumScaledFreq == scale) return alphabetSize;
  int delta = sumScaledFreq - scale;
  const int errThr = (int)freqs[idxMax] >> 4;
  if (abs(delta) <= errThr) { freqs[idxMax] -= delta; return alphabetSize; }
  if (delta < 0) { delta += errThr; freqs[idxMax] += errThr; }
  else { delta -= errThr; freqs[idxMax] -= errThr; }
  const int inc = (delta > 0) ? -1 : 1;
  delta = abs(delta);
  int round = 0;
  while ((++round < 6) && (delta > 0)) {
    int adjustments = 0;
    for (int i = 0; i < alphabetSize; i++) {
      const int idx = alphabet[i];
      if ((int)freqs[idx] <= 2) continue;
      freqs[idx] += inc;
      adjustments++;
      delta--;
      if (delta == 0) break;
    }
    if (adjustments == 0) break;
  }
  freqs[idxMax] = (u32)max((int)freqs[idxMax] - delta, 1);
  return alphabetSize;
}

// EntropyUtils.encodeAlphabet (K/entropy/EntropyUtils.java:38-75)
__device__ void ans_encode_alphabet(BitWriterD& bw, const u8* alphabet, int count) {
  if (count == 0) { bw.write(0, 1); bw.write(1, 1); return; }
  if (count == 256) { bw.write(0, 1); bw.write(0, 1); return; }
  bw.write(1, 1);
  const int lastMask = alphabet[count - 1] >> 3;
  bw.write((u32)lastMask, 5);
  int k = 0;
  for (int i = 0; i <= lastMask; i++) {
    u32 m = 0;
    while (k < count && (alphabet[k] >> 3) == i) { m |= 1u << (alphabet[k] & 7); k++; }
    bw.write(m, 8);
  }
}

// ANSRangeEncoder.encodeHeader (K/entropy/ANSRangeEncoder.java:211-252)
__device__ void ans_encode_header(BitWriterD& bw, int alphabetSize, const u8* alphabet, const u32* freqs, int lr) {
  ans_encode_alphabet(bw, alphabet, alphabetSize);
  if (alphabetSize <= 1) return;
  const int chkSize = (alphabetSize >= 64) ? 8 : 6;
  int llr = 3;
  while ((1 << llr) <= lr) llr++;
  for (int i = 1; i < alphabetSize; i += chkSize) {
    int mx = (int)freqs[alphabet[i]] - 1;
    int logMax = 0;
    const int endj = (i + chkSize < alphabetSize) ? i + chkSize : alphabetSize;
    for (int j = i + 1; j < endj; j++) mx = max(mx, (int)freqs[alphabet[j]] - 1);
    while ((1 << logMax) <= mx) logMax++;
    bw.write((u32)logMax, llr);
    if (logMax == 0) continue;
    for (int j = i; j < endj; j++) bw.write(freqs[alphabet[j]] - 1, logMax);
  }
}

// ANSRangeEncoder.Symbol.reset (K/entropy/ANSRangeEncoder.java:473-496) packed into two words:
//   a = invFreq (32 bits);  b = bias (13+ bits) | cmplFreq << 14 | (invShift - 32) << 28
// xMax is recomputed as freq << (31 - lr) with freq = scale - cmplFreq.
__device__ __forceinline__ void ans_symbol_reset(u32& a, u32& b, int cumFreq, int freq, int lr) {
  if (freq >= (1 << lr)) freq = (1 << lr) - 1;
  const u32 cmpl = (u32)((1 << lr) - freq);
  if (freq < 2) {
    a = 0xFFFFFFFFu;
    b = (u32)(cumFreq + (1 << lr) - 1) | (cmpl << 14) | (0u << 28);
  } else {
    int shift = 0;
    while (freq > (1 << shift)) shift++;
    a = (u32)((((1ull << (shift + 31)) + (u64)freq - 1) / (u64)freq) & 0xFFFFFFFFull);
    b = (u32)cumFreq | (cmpl << 14) | ((u32)(shift - 1) << 28);
  }
}

// ANSRangeEncoder.encodeSymbol (:315-328) minus the byte emission
__device__ __forceinline__ u32 ans_enc_step(u32 st, u32 a, u32 b) {
  const u32 q = __umulhi(st, a) >> (b >> 28);
  return st + (b & 0x3FFF) + q * ((b >> 14) & 0x3FFF);
}

// ================================================================================================================
// order 0 encode: grid (ceil(maxChunks/32), nBlocks), 128 threads; group of 4 lanes = one chunk
// ================================================================================================================
#define A0_GROUPS 32
struct A0EncSmem {          // rows padded by one entry: the eight chunks of a warp look up the same symbols, which must not share a bank
  u32 freq[A0_GROUPS][257];
  uint2 sym[A0_GROUPS][257];   // {invFreq, packed} (ans_symbol_reset); .x holds the cumulative frequency until the reset
  u8 alpha[A0_GROUPS][256];
};

__global__ void __launch_bounds__(128) ans0_encode_kernel(const KzgBlock* __restrict__ blocks, KzgEntParams P) {
  extern __shared__ __align__(16) u8 smem_raw[];
  A0EncSmem& S = *reinterpret_cast<A0EncSmem*>(smem_raw);
  const int g = threadIdx.x >> 2, j = threadIdx.x & 3;
  const int b = blockIdx.y;
  const int c = blockIdx.x * A0_GROUPS + g;
  const KzgBlock& B = blocks[b];
  if (!(B.status == 0 && B.entropy == P.entropy)) return;     // not this launch's codec (segments stay as the caller zeroed them)
  const int len = B.curLen;
  const u8* __restrict__ data = B.cur;
  const int chunkSize = P.chunkSize;
  const int lr = 12;
  const i64 gidx = (P.slotBase ? (i64)P.slotBase[b] : (i64)b * P.maxChunks) + c;
  const bool rawAll = (len <= 32);          // ANSRangeEncoder.encode :267-270: count <= 32 -> raw bytes
  const int start = c * chunkSize;
  const bool active = (!rawAll) && (c < P.maxChunks) && (start < len);
  const int end = active ? min(start + chunkSize, len) : 0;
  u8* hdr = P.hdrBuf + gidx * (i64)P.hdrStride;
  u8* pay = P.payBuf + gidx * (i64)P.payStride;
  const int bufLen = P.payStride - 16;      // scratch emulating `this.buffer` (filled from its end)
  KzgSeg* segs = P.segs + (i64)b * P.segsPerBlock + 1 + (i64)c * 2;

  if (rawAll && c == 0 && j == 0 && len > 0) {
    // whole call stored raw: one segment pointing at the data itself
    segs[0] = KzgSeg{data, 0, 0, (u64)len * 8};
    segs[1] = KzgSeg{nullptr, 0, 0, 0};
  } else if (!active && c < P.maxChunks && j == 0) {
    segs[0] = KzgSeg{nullptr, 0, 0, 0};
    segs[1] = KzgSeg{nullptr, 0, 0, 0};
  }

#ifdef KZG_A1_TIMING
  long long tq0 = clock64(), tq1, tq2, tq3;
#endif
  // ---- histogram (Global.computeHistogramOrder0 :274-330) ----
  for (int k = j; k < 256; k += 4) S.freq[g][k] = 0;
  __syncwarp();
  if (active) {
    int i = start + j;
    for (; i + 28 < end; i += 32) {          // eight loads in flight per lane (a byte load that waits for the atomic before it costs an L1 round trip each)
      u8 v[8];
      #pragma unroll
      for (int k = 0; k < 8; k++) v[k] = data[i + 4 * k];
      #pragma unroll
      for (int k = 0; k < 8; k++) atomicAdd(&S.freq[g][v[k]], 1u);
    }
    for (; i < end; i += 4) atomicAdd(&S.freq[g][data[i]], 1u);
  }
  __syncwarp();

#ifdef KZG_A1_TIMING
  tq1 = clock64();
#endif
  // ---- statistics + header (rebuildStatistics :419-449, updateFrequencies :171-200), lane 0 of the group ----
  int alphabetSize = 0;
  i64 hdrBits = 0;
  if (active && j == 0) {
    BitWriterD bw(hdr);
    bw.write((u32)(lr - 8), 3);
    alphabetSize = ans_normalize(S.freq[g], S.alpha[g], end - start, 1 << lr);
    int sum = 0;
    for (int k = 0; k < alphabetSize; k++) { const int sy = S.alpha[g][k]; S.sym[g][sy].x = (u32)sum; sum += (int)S.freq[g][sy]; }
    ans_encode_header(bw, alphabetSize, S.alpha[g], S.freq[g], lr);
    if (alphabetSize <= 1) {     // chunk skipped after its header (:296-299)
      bw.flush();
      segs[0] = KzgSeg{hdr, 0, 0, (u64)bw.bits()};
      segs[1] = KzgSeg{nullptr, 0, 0, 0};
    } else {
      hdrBits = bw.bits();
      bw.flush();
    }
  }
  alphabetSize = __shfl_sync(0xFFFFFFFFu, alphabetSize, (threadIdx.x & 31) & ~3);
  __syncwarp();
  const bool coding = active && (alphabetSize > 1);
  // Symbol.reset (:473-496) for every symbol of the alphabet: a 64-bit division each, shared by the group's four lanes
  if (coding) {
    for (int k = j; k < alphabetSize; k += 4) {
      const int sy = S.alpha[g][k];
      u32 sa, sb;
      ans_symbol_reset(sa, sb, (int)S.sym[g][sy].x, (int)S.freq[g][sy], lr);
      S.sym[g][sy] = make_uint2(sa, sb);
    }
  }
  __syncwarp();

#ifdef KZG_A1_TIMING
  tq2 = clock64();
#endif
  // ---- encodeChunk (:337-407): backwards, 4 interleaved states ----
  const int end4 = start + ((end - start) & -4);
  int n = bufLen - 1;
  if (coding) {
    const int tail = end - end4;
    if (j == 0) for (int i = end - 1; i >= end4; i--) pay[n - (end - 1 - i)] = data[i];
    n -= tail;
  }
  int steps = coding ? ((end4 - start) >> 2) : 0;
  int maxSteps = steps;
  for (int o = 16; o > 0; o >>= 1) maxSteps = max(maxSteps, __shfl_xor_sync(0xFFFFFFFFu, maxSteps, o));
  u32 st = ANS_TOP;
  int idx = n;
  const int gshift = (threadIdx.x & 31) & ~3;
  const u32 lowerMask = (1u << j) - 1;
  for (int s0 = 0; s0 < maxSteps; s0 += 8) {
    // the symbols of eight steps are requested together: they do not depend on the state, and a lone load per step would put an L1
    // round trip on every step of the chain
    int syms[8];
    #pragma unroll
    for (int k = 0; k < 8; k++) syms[k] = (s0 + k < steps) ? (int)data[end4 - 1 - 4 * (s0 + k) - j] : 0;
    #pragma unroll
    for (int k = 0; k < 8; k++) {
      const int s = s0 + k;
      if (s >= maxSteps) break;
      const bool on = s < steps;
      u32 a = 0, bb = 0;
      bool x = false;
      if (on) {
        const int sym = syms[k];
        const uint2 e = S.sym[g][sym];
        a = e.x; bb = e.y;
        const u32 xMax = ((u32)(1 << lr) - ((bb >> 14) & 0x3FFF)) << (31 - lr);
        x = (st >= xMax);
      }
      const u32 m = (__ballot_sync(0xFFFFFFFFu, x) >> gshift) & 0xFu;
      if (on) {
        if (x) {
          const int pos = idx - 2 * __popc(m & lowerMask);
          pay[pos] = (u8)st;
          pay[pos - 1] = (u8)(st >> 8);
          st >>= 16;
        }
        idx -= 2 * __popc(m);
        st = ans_enc_step(st, a, bb);
      }
    }
  }
#ifdef KZG_A1_TIMING
  tq3 = clock64();
  if (b == 3 && blockIdx.x == 0 && threadIdx.x == 0) printf("ans0 enc warp: hist %lld stats %lld code %lld cycles (%d steps)\n", tq1 - tq0, tq2 - tq1, tq3 - tq2, maxSteps);
#endif
  // gather the four final states in lane 0 of the group
  const u32 s1 = __shfl_sync(0xFFFFFFFFu, st, gshift + 1);
  const u32 s2 = __shfl_sync(0xFFFFFFFFu, st, gshift + 2);
  const u32 s3 = __shfl_sync(0xFFFFFFFFu, st, gshift + 3);
  if (coding && j == 0) {
    n = idx + 1;
    BitWriterD bw(hdr);
    // continue the header bit string: re-open at hdrBits (bytes already flushed; keep partial byte)
    bw.nbytes = hdrBits >> 3; bw.nacc = (int)(hdrBits & 7);
    bw.acc = (bw.nacc > 0) ? ((u64)hdr[bw.nbytes] >> (8 - bw.nacc)) : 0;
    write_varint(bw, bufLen - n);
    bw.write(st, 32); bw.write(s1, 32); bw.write(s2, 32); bw.write(s3, 32);
    bw.flush();
    segs[0] = KzgSeg{hdr, 0, 0, (u64)bw.bits()};
    segs[1] = KzgSeg{pay + n, 0, 0, (u64)(bufLen - n) * 8};
  }
}

// ================================================================================================================
// chunk scan for decode (order 0 and 1): one warp per block walks the chunk headers to find where each chunk starts
// (the only serial part of decoding: chunk k+1 starts where chunk k's byte count says).  The header bytes are staged
// 512 at a time into shared memory by the whole warp; every lane then runs the same parse out of the window
// (warp-uniform control flow, broadcast reads), so a field costs a shared-memory access instead of a trip to L2/HBM;
// the alphabet bitmap is counted 32 mask bytes at a time.
// ================================================================================================================
#define ANS_SCAN_WARPS 4
#define ANS_SCAN_WIN 512            // bytes
struct WarpBits {                   // MSB-first reader (format of BitReaderD) over a sliding shared window; all lanes in lock step
  const u8* stream; u64 pos, end; u32* win; u64 winBit0; int lane;
  __device__ __forceinline__ void stage() {
    __syncwarp();
    const u64 base4 = (pos >> 3) & ~3ull;
    const u64 lastByte = (end + 7) >> 3;
    #pragma unroll
    for (int k = 0; k < ANS_SCAN_WIN / 128; k++) {
      const u64 o = base4 + 4ull * (u64)(lane + 32 * k);
      win[lane + 32 * k] = (o < lastByte) ? *reinterpret_cast<const u32*>(stream + o) : 0u;
    }
    winBit0 = base4 * 8;
    __syncwarp();
  }
  __device__ __forceinline__ u32 read(int n) {      // n in 1..32; reads past `end` yield zeros
    u32 v = 0;
    if (pos + (u64)n <= end) {
      if (pos + 40 > winBit0 + 8ull * ANS_SCAN_WIN) stage();
      const u8* p = reinterpret_cast<const u8*>(win) + ((pos - winBit0) >> 3);
      const int sh = (int)(pos & 7);
      const u64 w = ((u64)p[0] << 32) | ((u64)p[1] << 24) | ((u64)p[2] << 16) | ((u64)p[3] << 8) | (u64)p[4];
      v = (u32)((w >> (40 - sh - n)) & ((n == 32) ? 0xFFFFFFFFull : ((1ull << n) - 1)));
    }
    pos += (u64)n;
    return v;
  }
  __device__ __forceinline__ void skip(u64 bits) { pos += bits; }
  __device__ __forceinline__ bool overrun() const { return pos > end; }
};
__device__ __forceinline__ int ans_skip_alphabet_warp(WarpBits& br) {   // EntropyUtils.decodeAlphabet sizes only
  if (br.read(1) == 0) return (br.read(1) == 1) ? 0 : 256;
  const int lastMask = (int)br.read(5);
  // lastMask + 1 mask bytes: lane i counts byte i
  const u64 p0 = br.pos;
  if (p0 + 8ull * (lastMask + 1) + 40 > br.winBit0 + 8ull * ANS_SCAN_WIN) br.stage();
  int c = 0;
  if (br.lane <= lastMask && p0 + 8ull * (br.lane + 1) <= br.end) {
    const u64 q = p0 + 8ull * br.lane;
    const u8* p = reinterpret_cast<const u8*>(br.win) + ((q - br.winBit0) >> 3);
    const int sh = (int)(q & 7);
    c = __popc(((((u32)p[0] << 8) | (u32)p[1]) >> (8 - sh)) & 0xFFu);
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
  br.pos = p0 + 8ull * (lastMask + 1);
  return c;
}
__global__ void __launch_bounds__(32 * ANS_SCAN_WARPS) ans_scan_kernel(KzgBlock* __restrict__ blocks, int nBlocks, KzgEntParams P, int order) {
  __shared__ u32 wins[ANS_SCAN_WARPS][ANS_SCAN_WIN / 4 + 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * ANS_SCAN_WARPS + warp;
  if (b >= nBlocks) return;
  KzgBlock& B = blocks[b];
  if (B.status != 0 || B.entropy != P.entropy) return;
  const int len = B.preLen;
  if (len <= 32) {   // raw (ANSRangeDecoder.decode :193-196)
    if (lane == 0) { if ((u64)len * 8 > (u64)B.srcBits) B.status = -KZG_ERR_PROCESS_BLOCK; else B.entBits = (i64)len * 8; }
    return;
  }
  WarpBits br; br.stream = P.stream; br.pos = (u64)B.srcBit; br.end = (u64)(B.srcBit + B.srcBits); br.win = wins[warp]; br.lane = lane;
  br.stage();
  KzgChunkInfo* ci = P.chunks + (i64)b * P.maxChunks;
  const int chunkSize = P.chunkSize;
  const int nChunks = (len + chunkSize - 1) / chunkSize;
  const int dim = 255 * order + 1;
  int status = 0;
  for (int c = 0; c < nChunks && status == 0; c++) {
    KzgChunkInfo info;
    info.hdrBit = (i64)br.pos;
    const int lr = 8 + (int)br.read(3);
    int llr = 3;
    while ((1 << llr) <= lr) llr++;
    int res = 0;
    for (int k = 0; k < dim; k++) {
      const int alphabetSize = ans_skip_alphabet_warp(br);
      if (alphabetSize == 0) continue;
      const int chkSize = (alphabetSize >= 64) ? 8 : 6;
      for (int i = 1; i < alphabetSize; i += chkSize) {
        const int logMax = (int)br.read(llr);
        const int endj = (i + chkSize < alphabetSize) ? i + chkSize : alphabetSize;
        br.skip((u64)(logMax * (endj - i)));
      }
      res += alphabetSize;
      if (br.overrun()) break;
    }
    info.alphabetSize = res;
    info.sz = 0; info.payBit = 0;
    info.st[0] = info.st[1] = info.st[2] = info.st[3] = 0;
    if (res == 0 || br.overrun()) { status = -KZG_ERR_PROCESS_BLOCK; break; }   // decode returns early (:218-219)
    if (!(order == 0 && res == 1)) {
      u32 value = br.read(8), sz = value & 0x7F;          // EntropyUtils.readVarInt (:283-300)
      int shift = 7;
      while (value >= 128) { value = br.read(8); sz |= ((value & 0x7F) << shift); if (shift == 28) break; shift += 7; }
      if ((i32)sz < 0 || sz >= (1u << 27)) { status = -KZG_ERR_PROCESS_BLOCK; break; }
      info.st[0] = br.read(32); info.st[1] = br.read(32); info.st[2] = br.read(32); info.st[3] = br.read(32);
      info.sz = (i32)sz;
      info.payBit = (i64)br.pos;
      br.skip((u64)sz * 8);
    }
    if (br.overrun()) { status = -KZG_ERR_PROCESS_BLOCK; break; }
    if (lane == 0) ci[c] = info;
  }
  if (lane == 0) { if (status < 0) B.status = status; else B.entBits = (i64)br.pos - B.srcBit; }
}

// 16 bits at byte offset `off` of a payload that starts at absolute bit `payBit`; bytes at or beyond `sz`
// read as zero (ANSRangeDecoder zero-fills its buffer, :372-375)
__device__ __forceinline__ u32 ans_pay16(const u8* __restrict__ stream, i64 payBit, int off, int sz) {
  const u64 pos = (u64)payBit + (u64)off * 8;
  const u8* p = stream + (pos >> 3);
  const int sh = (int)(pos & 7);
  const u32 w = ((u32)p[0] << 16) | ((u32)p[1] << 8) | (u32)p[2];
  u32 v = (w >> (8 - sh)) & 0xFFFFu;
  if (off + 2 > sz) v = (off >= sz) ? 0 : (v & 0xFF00u);
  return v;
}
__device__ __forceinline__ u32 ans_pay8(const u8* __restrict__ stream, i64 payBit, int off) {
  return get_bits(stream, (u64)payBit + (u64)off * 8, 8);
}

// ANSRangeDecoder.decodeHeader (:452-544) for one context: fills freq[256] (0 for absent symbols).
// Returns alphabetSize, or -1 on an invalid header.
template <bool CLEAR = true>
__device__ int ans_decode_ctx_header(BitReaderD& br, int lr, int llr, u16* freq, u8* alphabet, bool& cleared) {
  int alphabetSize = 0;
  if (br.read(1) == 0) {
    if (br.read(1) == 1) return 0;
    alphabetSize = 256;
    for (int i = 0; i < 256; i++) alphabet[i] = (u8)i;
  } else {
    const int lastMask = (int)br.read(5);
    for (int i = 0; i <= lastMask; i++) {
      const u32 mask = br.read(8);
      for (int jj = 0; jj < 8; jj++)
        if (mask & (1u << jj)) alphabet[alphabetSize++] = (u8)((i << 3) + jj);
    }
  }
  if (alphabetSize == 0) return 0;
  const int scale = 1 << lr;
  if (CLEAR && alphabetSize != 256) { for (int i = 0; i < 256; i++) freq[i] = 0; cleared = true; }
  const int chkSize = (alphabetSize >= 64) ? 8 : 6;
  int sum = 0;
  for (int i = 1; i < alphabetSize; i += chkSize) {
    const int logMax = (int)br.read(llr);
    if ((1 << logMax) > scale) return -1;
    const int endj = (i + chkSize < alphabetSize) ? i + chkSize : alphabetSize;
    for (int jj = i; jj < endj; jj++) {
      const int f = (logMax == 0) ? 1 : (int)(1 + br.read(logMax));
      if (f <= 0 || f >= scale) return -1;
      freq[alphabet[jj]] = (u16)f;
      sum += f;
    }
  }
  if (scale <= sum) return -1;
  freq[alphabet[0]] = (u16)(scale - sum);   // scale - sum may be 1<<lr when alphabetSize == 1 (stored mod 2^16: lr <= 15)
  return alphabetSize;
}

// ================================================================================================================
// order 0 decode: one warp = 8 chunks of 4 lanes (one lane per interleaved state), one warp per CTA, 5 CTAs per SM
// (42 KiB of tables each).
//   * headers are parsed by the whole warp out of a shared-memory copy (alphabet bitmap: one mask byte per lane; the
//     frequency groups: one group per lane once the 32-step chain of group offsets is known), cumulative frequencies by a
//     warp scan, the slot -> symbol table filled slot-parallel (every lane owns 1/32 of the slots);
//   * the coded bytes stream through a 256-byte shared-memory ring per chunk, filled by cp.async (LDGSTS) 64 bytes at a
//     time on a fixed schedule (every 8 steps, the most a chunk can consume), so no lane ever waits on a global load and
//     the eight chunks of a warp never diverge; a lane that renormalises reads its 16 bits straight from the ring;
//     the whole payload is pulled into L2 while the headers are parsed;
//   * output: the four symbols of a step are packed with two shuffles, every lane stores one 32-bit word per four
//     steps (16 contiguous bytes per group).
// ANSRangeDecoder.decodeChunkV2 (:357-440): symbol i+j is decoded with state st(3-j); states below TOP read 16 bits
// each in lane order (the ballot / popcount below).
// ================================================================================================================
#define A0D_CHUNKS 8
#define A0D_HDR_WORDS 144
#define A0D_RING 256
struct A0DecSmem {
  u8 f2s[A0D_CHUNKS][4096];
  u32 sym[A0D_CHUNKS][256];      // freq | cumFreq << 16
  u32 hdr[A0D_HDR_WORDS + 2];    // header bytes of the chunk being parsed, as big-endian words
  u16 freq[256];
  u8 alpha[256];
  uint4 ring[A0D_CHUNKS][A0D_RING / 16];   // per chunk: a 256-byte window of its coded bytes, filled by cp.async
};
__device__ __forceinline__ u64 kzg_shl64(u64 x, u32 n) { u64 r; asm("shl.b64 %0, %1, %2;" : "=l"(r) : "l"(x), "r"(n)); return r; }   // n >= 64 -> 0
__device__ __forceinline__ u64 kzg_shr64(u64 x, u32 n) { u64 r; asm("shr.u64 %0, %1, %2;" : "=l"(r) : "l"(x), "r"(n)); return r; }
__device__ __forceinline__ u32 a0d_bits(const u32* w, u32 pos, int n) {           // n in 1..32, MSB first
  const u32 i = pos >> 5;
  return __funnelshift_l(w[i + 1], w[i], pos & 31) >> (32 - n);
}

// ANSRangeDecoder.decodeHeader (:452-544) for chunk slot k of this warp, all 32 lanes: fills S.sym[k] / S.f2s[k].
// Returns the alphabet size (1: `single` is the symbol, no tables), or -1 for a header the reference rejects.
__device__ int a0d_parse_header(A0DecSmem& S, int k, const u8* __restrict__ stream, u64 hdrBit, u64 endBit, int lane, int& lrOut, int& single) {
  const u64 byte0 = (hdrBit >> 3) & ~3ull;
  const u64 lastByte = (endBit + 7) >> 3;
  __syncwarp();
  for (int i = lane; i < A0D_HDR_WORDS + 2; i += 32) {
    const u64 o = byte0 + 4ull * i;
    S.hdr[i] = (o < lastByte) ? __byte_perm(__ldg(reinterpret_cast<const u32*>(stream + o)), 0, 0x0123) : 0u;
  }
  __syncwarp();
  u32 pos = (u32)(hdrBit - 8 * byte0);
  const int lr = 8 + (int)a0d_bits(S.hdr, pos, 3); pos += 3;
  lrOut = lr;
  single = 0;
  if (lr > 12) return -1;                    // order-0 tables are sized for the reference's logRange 12 (ANSRangeEncoder :40)
  const int llr = 4;                         // while ((1 << llr) <= lr) llr++ from 3: 4 for every lr in 8..15
  const int scale = 1 << lr;
  int as;
  if (a0d_bits(S.hdr, pos, 1) == 0) {        // EntropyUtils.decodeAlphabet (:84-131)
    if (a0d_bits(S.hdr, pos + 1, 1) == 1) return -1;      // empty alphabet: decode() stops (:218-219)
    pos += 2; as = 256;
    for (int i = lane; i < 256; i += 32) S.alpha[i] = (u8)i;
  } else {
    const int lastMask = (int)a0d_bits(S.hdr, pos + 1, 5); pos += 6;
    const u32 m = (lane <= lastMask) ? a0d_bits(S.hdr, pos + 8 * lane, 8) : 0u;
    const int cnt = __popc(m);
    int incl = cnt;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
    as = __shfl_sync(0xFFFFFFFFu, incl, 31);
    int at = incl - cnt;
    for (u32 mm = m; mm; mm &= mm - 1) S.alpha[at++] = (u8)((lane << 3) + (__ffs(mm) - 1));
    pos += 8 * (lastMask + 1);
    if (as == 0) return -1;
  }
  if (as != 256) for (int i = lane; i < 256; i += 32) S.freq[i] = 0;
  __syncwarp();
  // the chain of group offsets (every lane runs it; lane gi keeps group gi)
  const int chk = (as >= 64) ? 8 : 6;
  int myPos = 0, myLog = 0, myCnt = 0, gi = 0;
  bool bad = false;
  for (int i = 1; i < as; i += chk, gi++) {
    const int logMax = (int)a0d_bits(S.hdr, pos, llr);
    const int cnt = min(chk, as - i);
    if (lane == gi) { myPos = (int)pos + llr; myLog = logMax; myCnt = cnt; }
    if ((1 << logMax) > scale) bad = true;
    pos += llr + logMax * cnt;
  }
  int sum = 0;
  if (!bad) {
    for (int t = 0; t < myCnt; t++) {
      const int f = (myLog == 0) ? 1 : 1 + (int)a0d_bits(S.hdr, (u32)(myPos + t * myLog), myLog);
      if (f >= scale) bad = true;
      S.freq[S.alpha[1 + lane * chk + t]] = (u16)f;
      sum += f;
    }
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
  if (__any_sync(0xFFFFFFFFu, bad) || scale <= sum) return -1;
  if (lane == 0) S.freq[S.alpha[0]] = (u16)(scale - sum);
  __syncwarp();
  if (as == 1) { single = S.alpha[0]; return 1; }
  // cumulative frequencies: lane owns symbols 8*lane .. 8*lane+7
  int f8[8], tot = 0;
  #pragma unroll
  for (int t = 0; t < 8; t++) { f8[t] = S.freq[8 * lane + t]; tot += f8[t]; }
  int incl = tot;
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
  int cum = incl - tot;
  #pragma unroll
  for (int t = 0; t < 8; t++) { S.sym[k][8 * lane + t] = (u32)f8[t] | ((u32)cum << 16); cum += f8[t]; }
  __syncwarp();
  // slot -> symbol, slot-parallel: lane owns slots [lane * per, (lane + 1) * per)
  const int per = scale >> 5;                  // 8 .. 128
  const int slot0 = lane * per;
  int lo = 0, hi = 255;
  while (lo < hi) {                            // first symbol whose interval ends beyond slot0
    const int mid = (lo + hi) >> 1;
    const u32 e = S.sym[k][mid];
    if ((int)((e >> 16) + (e & 0xFFFFu)) > slot0) hi = mid; else lo = mid + 1;
  }
  int sy = lo;
  u32 e = S.sym[k][sy];
  int endOf = (int)((e >> 16) + (e & 0xFFFFu));
  u32* row = reinterpret_cast<u32*>(&S.f2s[k][slot0]);
  for (int i = 0; i < per; i += 4) {
    u32 w = 0;
    #pragma unroll
    for (int t = 0; t < 4; t++) {
      while (slot0 + i + t >= endOf) { sy++; e = S.sym[k][sy]; endOf = (int)((e >> 16) + (e & 0xFFFFu)); }
      w |= (u32)sy << (8 * t);
    }
    row[i >> 2] = w;
  }
  __syncwarp();
  return as;
}

// Chunk scan, order 0: one CTA per block.  Warp 0 walks the chunk headers (chunk k+1 starts where chunk k's byte count says:
// the one serial chain of the decode) out of a shared-memory copy of each header, every lane running the same parse; the other
// three warps pull the block's whole payload into L2 meanwhile, so the walk (and the decode kernel after it) never waits for DRAM.
// Per chunk: one staging round trip to L2 + the 32-step chain of frequency-group offsets.
__global__ void __launch_bounds__(128) ans0_scan_kernel(KzgBlock* __restrict__ blocks, int nBlocks, KzgEntParams P) {
  __shared__ u32 hdr[A0D_HDR_WORDS + 2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  KzgBlock& B = blocks[b];
  if (B.status != 0 || B.entropy != P.entropy) return;
  const int len = B.preLen;
  if (len <= 32) {   // raw (ANSRangeDecoder.decode :193-196)
    if (threadIdx.x == 0) { if ((u64)len * 8 > (u64)B.srcBits) B.status = -KZG_ERR_PROCESS_BLOCK; else B.entBits = (i64)len * 8; }
    return;
  }
  const u8* __restrict__ stream = P.stream;
  const u64 endBit = (u64)(B.srcBit + B.srcBits);
  const u64 lastByte = (endBit + 7) >> 3;
  if (warp > 0) {
    const u64 first = ((u64)B.srcBit >> 3) & ~127ull;
    for (u64 o = first + 128ull * (threadIdx.x - 32); o < lastByte; o += 128ull * 96) asm volatile("prefetch.global.L2 [%0];" ::"l"(stream + o));
    return;
  }
  KzgChunkInfo* ci = P.chunks + (i64)b * P.maxChunks;
  const int chunkSize = P.chunkSize;
  const int nChunks = (len + chunkSize - 1) / chunkSize;
  u64 at = (u64)B.srcBit;
  int status = 0;
  for (int c = 0; c < nChunks; c++) {
    if (at + 3 > endBit) { status = -KZG_ERR_PROCESS_BLOCK; break; }
    const u64 byte0 = (at >> 3) & ~3ull;
    __syncwarp();
    for (int i = lane; i < A0D_HDR_WORDS + 2; i += 32) {
      const u64 o = byte0 + 4ull * i;
      hdr[i] = (o < lastByte) ? __byte_perm(__ldcg(reinterpret_cast<const u32*>(stream + o)), 0, 0x0123) : 0u;
    }
    __syncwarp();
    u32 pos = (u32)(at - 8 * byte0);
    const u32 pos00 = pos;
    pos += 3;                                // logRange
    int as;
    if (a0d_bits(hdr, pos, 1) == 0) { as = (a0d_bits(hdr, pos + 1, 1) == 1) ? 0 : 256; pos += 2; }
    else {
      const int lastMask = (int)a0d_bits(hdr, pos + 1, 5); pos += 6;
      int cnt = (lane <= lastMask) ? __popc(a0d_bits(hdr, pos + 8 * lane, 8)) : 0;
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
      as = cnt;
      pos += 8 * (lastMask + 1);
    }
    if (as == 0) { status = -KZG_ERR_PROCESS_BLOCK; break; }       // decode returns early (:218-219)
    const int chk = (as >= 64) ? 8 : 6;
    for (int i = 1; i < as; i += chk) pos += 4 + (int)a0d_bits(hdr, pos, 4) * min(chk, as - i);
    KzgChunkInfo info;
    info.hdrBit = (i64)at; info.alphabetSize = as; info.sz = 0; info.payBit = 0;
    info.st[0] = info.st[1] = info.st[2] = info.st[3] = 0;
    if (as != 1) {
      u32 value = a0d_bits(hdr, pos, 8), sz = value & 0x7F; pos += 8;          // EntropyUtils.readVarInt (:283-300)
      int shift = 7;
      while (value >= 128) { value = a0d_bits(hdr, pos, 8); pos += 8; sz |= ((value & 0x7F) << shift); if (shift == 28) break; shift += 7; }
      if ((i32)sz < 0 || sz >= (1u << 27)) { status = -KZG_ERR_PROCESS_BLOCK; break; }
      info.st[0] = a0d_bits(hdr, pos, 32); info.st[1] = a0d_bits(hdr, pos + 32, 32); info.st[2] = a0d_bits(hdr, pos + 64, 32); info.st[3] = a0d_bits(hdr, pos + 96, 32);
      pos += 128;
      info.sz = (i32)sz;
      info.payBit = (i64)(at + (pos - pos00));
      at += (u64)(pos - pos00) + (u64)sz * 8;
    } else at += (u64)(pos - pos00);
    if (at > endBit) { status = -KZG_ERR_PROCESS_BLOCK; break; }
    if (lane == 0) ci[c] = info;
  }
  if (lane == 0) { if (status < 0) B.status = status; else B.entBits = (i64)at - B.srcBit; }
}

__global__ void __launch_bounds__(32) ans0_decode_kernel(KzgBlock* __restrict__ blocks, KzgEntParams P) {
  __shared__ A0DecSmem S;
  const int lane = threadIdx.x, g = lane >> 2, j = lane & 3, gl = lane & ~3;
  const int b = blockIdx.y;
  KzgBlock& B = blocks[b];
  const bool blockOk = (B.status == 0 && B.entropy == P.entropy);
  const int len = blockOk ? B.preLen : 0;
  u8* __restrict__ out = B.cur;
  const u8* __restrict__ stream = P.stream;
  if (len <= 32) {   // raw (ANSRangeDecoder.decode :193-196)
    if (blockOk && blockIdx.x == 0 && lane < len) out[lane] = (u8)get_bits(stream, (u64)B.srcBit + 8ull * lane, 8);
    return;
  }
  const int chunkSize = P.chunkSize;
  const int nChunks = min((len + chunkSize - 1) / chunkSize, P.maxChunks);
  const int c0 = blockIdx.x * A0D_CHUNKS;
  if (c0 >= nChunks) return;
  const int c = c0 + g;
  const bool active = c < nChunks;
  const int start = c * chunkSize;
  const int end = active ? min(start + chunkSize, len) : 0;
  const u64 endBit = (u64)(B.srcBit + B.srcBits);
  const u8* limit = reinterpret_cast<const u8*>((reinterpret_cast<uintptr_t>(stream + ((endBit + 7) >> 3)) + 7) & ~(uintptr_t)7);
  KzgChunkInfo info;
  info.hdrBit = 0; info.payBit = 0; info.sz = 0; info.alphabetSize = 0; info.st[0] = info.st[1] = info.st[2] = info.st[3] = 0;
  if (active) info = P.chunks[(i64)b * P.maxChunks + c];
  const u8* payByte = stream + ((u64)info.payBit >> 3);
  if (active && info.sz > 0) {       // pull the coded bytes into L2 while the headers are parsed
    for (int o = 128 * j; o < info.sz + 64; o += 512) if (payByte + o < limit) asm volatile("prefetch.global.L2 [%0];" ::"l"(payByte + o));
  }
  // ---- headers -> tables, one chunk after the other, the whole warp on each ----
  int lr = 12, bad = 0, single = -1;
  for (int k = 0; k < A0D_CHUNKS; k++) {
    if (c0 + k >= nChunks) break;
    const u64 hb = (u64)__shfl_sync(0xFFFFFFFFu, (unsigned long long)info.hdrBit, 4 * k);
    int lrK, sgl;
    const int as = a0d_parse_header(S, k, stream, hb, endBit, lane, lrK, sgl);
    if (g == k) { lr = lrK; bad = (as <= 0) ? 1 : 0; single = (as == 1) ? sgl : -1; }
  }
  __syncwarp();
  const bool coding = active && !bad && single < 0;
  const bool aligned = ((reinterpret_cast<uintptr_t>(out) + (uintptr_t)start) & 3) == 0;
  if (active && single >= 0) {     // shortcut for chunks with only one symbol (:221-224)
    for (int i = start + j; i < end; i += 4) out[i] = (u8)single;
  }

  // ---- decodeChunkV2 (:357-440) ----
  const int end4 = start + ((end - start) & -4);
  const int steps = coding ? ((end4 - start) >> 2) : 0;
  int maxSteps = steps;
  for (int o = 16; o > 0; o >>= 1) maxSteps = max(maxSteps, __shfl_xor_sync(0xFFFFFFFFu, maxSteps, o));
  u32 st = (j == 0) ? info.st[3] : ((j == 1) ? info.st[2] : ((j == 2) ? info.st[1] : info.st[0]));   // lane j decodes symbol i+j with state st(3-j) (:392-405)
  if (!coding) st = 0u;
  const u32 mask = (1u << lr) - 1;
  const u32 lowerMask = (1u << j) - 1;
  // the chunk's coded bytes: ring byte x holds stream byte base16 + x (mod 256); lane j fetches the j-th 16 bytes of a batch
  const u8* base16 = reinterpret_cast<const u8*>(reinterpret_cast<uintptr_t>(payByte) & ~(uintptr_t)15);
  u32 cbits = (u32)((payByte - base16) * 8) + (u32)((u64)info.payBit & 7);      // bit position of the cursor, relative to base16
  const u32 cbits0 = cbits;
  u32 fetched = 0;                                                               // bytes requested so far (multiple of 64)
  const u32 ringAddr = (u32)__cvta_generic_to_shared(&S.ring[g][0]);
  const u32* ring32 = reinterpret_cast<const u32*>(&S.ring[g][0]);
  auto fetch = [&]() {                         // 64 more bytes into the ring (reads beyond the block's payload are skipped)
    const u8* srcp = base16 + fetched + 16 * j;
    if (coding && srcp < limit) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ringAddr + ((fetched + 16 * j) & (A0D_RING - 1))), "l"(srcp) : "memory");
    fetched += 64;
  };
  fetch(); fetch(); fetch();
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
  const u8* f2s = S.f2s[g];
  const u32* symt = S.sym[g];
  u32* out32 = reinterpret_cast<u32*>(out + (aligned ? start : 0));
  u32 keep = 0;
  for (int s = 0; s < maxSteps; s++) {
    const bool on = s < steps;
    bool need = false;
    u32 symv = 0;
    if (on) {
      const u32 slot = st & mask;
      symv = f2s[slot];
      const u32 fc = symt[symv];
      st = (fc & 0xFFFFu) * (st >> lr) + slot - (fc >> 16);
      need = st < (u32)ANS_TOP;
    }
    const u32 m = (__ballot_sync(0xFFFFFFFFu, need) >> gl) & 0xFu;
    u32 pk = symv << (8 * j);
    pk |= __shfl_xor_sync(0xFFFFFFFFu, pk, 1);
    pk |= __shfl_xor_sync(0xFFFFFFFFu, pk, 2);
    if (need) {                                // 16 bits at bit p of the ring (big-endian bit order inside the byte stream)
      const u32 p = cbits + 16u * __popc(m & lowerMask);
      const u32 w = (p >> 5) & (A0D_RING / 4 - 1);
      const u32 hi = __byte_perm(ring32[w], 0, 0x0123), lo = __byte_perm(ring32[(w + 1) & (A0D_RING / 4 - 1)], 0, 0x0123);
      st = (st << 16) | (__funnelshift_l(lo, hi, p & 31u) >> 16);
    }
    cbits += 16u * __popc(m);
    if (on) {
      if (aligned) {
        if ((s & 3) == j) keep = pk;
        if ((s & 3) == 3) out32[4 * (s >> 2) + j] = keep;
      } else out[start + 4 * s + j] = (u8)symv;
    }
    if ((s & 7) == 7) {                        // every chunk tops its ring up on the same steps: at most 64 bytes go per 8 steps
      // invariant: fetched - cursor >= 72 bytes at every top-up (64 to consume + 8 of read slack), and < 200 (the ring keeps the cursor's word)
      if (fetched - (cbits >> 3) < 136u) fetch();
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");      // everything but the batch just issued has landed
      __syncwarp();
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (coding) {
    if (aligned && j < (steps & 3)) out32[(steps & ~3) + j] = keep;
    const int consumed = (int)((cbits - cbits0) >> 3);
    const int tail = end - end4;
    const int sz = info.sz;
    if (j == 0) {
      for (int i = 0; i < tail; i++) out[end4 + i] = (consumed + i < sz) ? (u8)get_bits(stream, (u64)info.payBit + 8ull * (u64)(consumed + i), 8) : 0;
      if (consumed + tail != sz) atomicExch(&B.status, -KZG_ERR_PROCESS_BLOCK);   // decodeChunkV2 returns n == sz
    }
  }
  if (active && bad && j == 0) atomicExch(&B.status, -KZG_ERR_PROCESS_BLOCK);
}

// ================================================================================================================
// order 1: one CTA per chunk (4 MiB); the 256-context tables live in global memory (L2-resident)
// ================================================================================================================
// encode, per-chunk global scratch (u32 units): sym[256][256] (uint4 {xMax, invFreq, bias | cmplFreq << 16, invShift - 32}), freq[256][257],
// pad, hdrPriv[256][120], alpha[256][256 bytes].  256 threads: histogram (one red.global per byte), one context per thread (normalizeFrequencies, Symbol
// tables, its header bits into a private buffer), the headers merged bit-granularly into the chunk header; then warp 0 codes:
// ONE lane runs the four states as four interleaved dependency chains (a lone warp issues a dependent instruction only every 5-7
// cycles: instruction-level parallelism inside one thread is worth more than four lanes in lock step) while all 32 lanes fetch the
// Symbol entries two batches of 32 steps ahead (data byte -> table entry are dependent global loads; the rANS recurrence itself
// never waits for memory).
// decode, per-chunk global scratch: f2s[256][2048] (u8), sym[256][256] (u32), freq[256] (u16), alpha[256]
#define A1_THREADS 256
#ifdef KZG_A1_TIMING
#define A1_CLK(i) clk[i] = clock64();
#else
#define A1_CLK(i)
#endif
#define A1_HDR_WORDS 120          // private header buffer per context: 2 + 5 + 256 alphabet bits, 43 groups x 4 + 255 x 11 frequency bits < 480 bytes
#define A1_BATCH 32

__global__ void __launch_bounds__(A1_THREADS) ans1_encode_kernel(const KzgBlock* __restrict__ blocks, KzgEntParams P) {
  __shared__ uint4 ring[2][4][A1_BATCH];
#ifdef KZG_A1_TIMING
  long long clk[5];
#endif
  __shared__ u32 scanBuf[A1_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const int c = blockIdx.x;
  const KzgBlock& B = blocks[b];
  if (!(B.status == 0 && B.entropy == P.entropy)) return;
  const int len = B.curLen;
  const u8* __restrict__ data = B.cur;
  const int chunkSize = P.chunkSize;     // 4 MiB
  const int lr = 11;
  const i64 gidx = (P.slotBase ? (i64)P.slotBase[b] : (i64)b * P.maxChunks) + c;
  KzgSeg* segs = P.segs + (i64)b * P.segsPerBlock + 1 + (i64)c * 2;
  if (len <= 32) {
    if (tid == 0) {
      segs[0] = (c == 0 && len > 0) ? KzgSeg{data, 0, 0, (u64)len * 8} : KzgSeg{nullptr, 0, 0, 0};
      segs[1] = KzgSeg{nullptr, 0, 0, 0};
    }
    return;
  }
  if ((i64)c * chunkSize >= (i64)len) {          // (64-bit: the grid may be far wider than this block's chunk count)
    if (tid == 0) { segs[0] = KzgSeg{nullptr, 0, 0, 0}; segs[1] = KzgSeg{nullptr, 0, 0, 0}; }
    return;
  }
  const int start = c * chunkSize;
  const int end = min(start + chunkSize, len);
  u8* hdr = P.hdrBuf + gidx * (i64)P.hdrStride;
  u8* pay = P.payBuf + gidx * (i64)P.payStride;
  const int bufLen = P.payStride - 16;
  u32* tab = P.tabBuf + gidx * (i64)P.tabStride;       // u32 units
  uint4* sym = reinterpret_cast<uint4*>(tab);          // [256 contexts][256 symbols]
  u32* freq = tab + 4 * 65536;                         // [256][257]
  u32* hdrPriv = tab + 4 * 65536 + 256 * 257 + 64;     // [256][A1_HDR_WORDS]
  u8* alphaAll = (u8*)(hdrPriv + 256 * A1_HDR_WORDS);  // [256][256]

  A1_CLK(0)
  // Symbol objects are re-created per encode() call (:277-282): zero = "new Symbol()"
  for (int k = tid; k < (4 * 65536 + 256 * 257) / 4; k += A1_THREADS) reinterpret_cast<uint4*>(tab)[k] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  A1_CLK(1)
  // ---- order-1 histogram: first byte of each quarter in context 0 (rebuildStatistics :430-446) ----
  {
    const int quarter = (end - start) >> 2;
    if (quarter == 0) {
      if (tid == 0) {
        int prv = 0;
        for (int i = start; i < end; i++) { freq[prv * 257 + data[i]]++; prv = data[i]; }
      }
    } else {
      const int n4 = 4 * quarter;
      for (int p0 = tid * 4; p0 < n4; p0 += A1_THREADS * 4) {
        int prv = (p0 > 0) ? (int)data[start + p0 - 1] : 0;
        #pragma unroll
        for (int r = 0; r < 4; r++) {
          const int p = p0 + r;
          if (p >= n4) break;
          const int cur = data[start + p];
          if (p == quarter || p == 2 * quarter || p == 3 * quarter) prv = 0;
          atomicAdd(&freq[prv * 257 + cur], 1u);
          prv = cur;
        }
      }
    }
  }
  __syncthreads();
  A1_CLK(2)

  // ---- one context per thread: statistics, Symbol table, header bits ----
  u32 myBits;
  {
    const int k = tid;
    u32* f = freq + k * 257;
    u8* alpha = alphaAll + k * 256;
    u32 total = 0;
    for (int i = 0; i < 256; i++) total += f[i];
    const int alphabetSize = ans_normalize(f, alpha, (int)total, 1 << lr);
    if (alphabetSize > 0) {
      int sum = 0;
      for (int i = 0; i < alphabetSize; i++) {
        const int sy = alpha[i];
        u32 sa, sb;
        ans_symbol_reset(sa, sb, sum, (int)f[sy], lr);
        // unpacked for the coding loop: xMax = freq << (31 - lr) (:479), invFreq, bias | cmplFreq << 16, invShift - 32
        sym[k * 256 + sy] = make_uint4(((u32)(1 << lr) - ((sb >> 14) & 0x3FFF)) << (31 - lr), sa, (sb & 0x3FFF) | (((sb >> 14) & 0x3FFF) << 16), sb >> 28);
        sum += (int)f[sy];
      }
    }
    BitWriterD bw((u8*)(hdrPriv + k * A1_HDR_WORDS));
    if (k == 0) bw.write((u32)(lr - 8), 3);
    ans_encode_header(bw, alphabetSize, alpha, f, lr);
    myBits = (u32)bw.bits();
    bw.flush();
  }
  // exclusive scan of the header lengths over the 256 contexts
  u32 incl = myBits;
  for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) scanBuf[warp] = incl;
  __syncthreads();
  u32 wbase = 0, hdrBits32 = 0;
  for (int w = 0; w < A1_THREADS / 32; w++) { const u32 t = scanBuf[w]; if (w < warp) wbase += t; hdrBits32 += t; }
  const u32 myOff = wbase + incl - myBits;
  u32* hdrW = reinterpret_cast<u32*>(hdr);
  for (u32 w = tid; w < (hdrBits32 >> 5) + 2; w += A1_THREADS) hdrW[w] = 0;
  __syncthreads();
  {
    const u32* src = hdrPriv + tid * A1_HDR_WORDS;
    for (u32 w = 0; w * 32 < myBits; w++) {
      u32 v = __byte_perm(src[w], 0, 0x0123);                   // the 32 bits in stream order, first bit = MSB
      const u32 left = myBits - w * 32;
      if (left < 32) v &= ~(0xFFFFFFFFu >> left);
      const u32 d = myOff + w * 32, wi = d >> 5, sh = d & 31;
      atomicOr(&hdrW[wi], __byte_perm(v >> sh, 0, 0x0123));
      if (sh) atomicOr(&hdrW[wi + 1], __byte_perm(v << (32 - sh), 0, 0x0123));
    }
  }
  __threadfence_block();
  __syncthreads();
  A1_CLK(3)
  if (warp != 0) return;
  const i64 hdrBits = (i64)hdrBits32;

  // ---- encodeChunk order 1 (:359-390): state q codes quarter q backwards ----
  const int end4 = start + ((end - start) & -4);
  int n = bufLen - 1;
  const int tail = end - end4;
  if (lane == 0) for (int i = end - 1; i >= end4; i--) pay[n - (end - 1 - i)] = data[i];
  n -= tail;
  const int quarter = (end4 - start) >> 2;
  u32 st0 = ANS_TOP, st1 = ANS_TOP, st2 = ANS_TOP, st3 = ANS_TOP;      // lane 0 runs all four states
  int idx = n;
  // Java: i_q = start + (q+1)*quarter - 2, prv_q = block[i_q + 1]; loop while i0 >= start (quarter-1 steps), then "last symbols" in ctx 0.
  // Step s of quarter q codes symbol block[i_q + 1 - s] in context block[i_q - s] (context 0 at the last step, s == steps).
  const int steps = (quarter > 0) ? quarter - 1 : 0;
  const int fq = lane >> 3, fsub = lane & 7;                   // fetch role: quarter fq, steps [4 * fsub, 4 * fsub + 4) of a batch
  const i64 fiq = (i64)start + (i64)(fq + 1) * quarter - 2;
  const int nBatches = steps / A1_BATCH + 1;
  u32 dat[2] = {0, 0};       // the five bytes of a batch (four symbols + contexts): dat[0] = bytes p-3..p, dat[1] = byte p+1, p = i_q - S - 4 fsub
  uint4 tv[4];
  auto loadData = [&](int bt, u32* d) {
    d[0] = 0; d[1] = 0;
    const i64 s0 = (i64)bt * A1_BATCH + 4 * fsub;
    if (bt >= nBatches || s0 > steps) return;
    const i64 p = fiq - s0;              // context position of the batch's first step for this lane
    #pragma unroll
    for (int r = 0; r < 4; r++) { const i64 q = p - 3 + r; if (q >= 0 && q >= (i64)start - 1) d[0] |= (u32)data[q] << (8 * r); }
    if (p + 1 >= 0) d[1] = data[p + 1];
  };
  auto loadTab = [&](int bt, const u32* d, uint4* t) {
    const i64 s0 = (i64)bt * A1_BATCH + 4 * fsub;
    #pragma unroll
    for (int r = 0; r < 4; r++) {
      t[r] = make_uint4(0, 0, 0, 0);
      const i64 s = s0 + r;
      if (bt >= nBatches || s > steps) continue;
      // step s: context byte at p - r (byte 3 - r of d[0]), symbol at p - r + 1
      const u32 sy = (r == 0) ? d[1] : ((d[0] >> (8 * (4 - r))) & 0xFF);
      const u32 cx = (s == steps) ? 0u : ((d[0] >> (8 * (3 - r))) & 0xFF);
      t[r] = sym[cx * 256 + sy];
    }
  };
  auto storeTab = [&](int bt, const uint4* t) {
    #pragma unroll
    for (int r = 0; r < 4; r++) ring[bt & 1][fq][4 * fsub + r] = t[r];
  };
  u32 dnext[2];
  uint4 tnext[4];
  loadData(0, dat); loadTab(0, dat, tv); storeTab(0, tv);
  loadData(1, dat); loadTab(1, dat, tv);
  loadData(2, dat);
  __syncwarp();
  for (int bt = 0; bt < nBatches; bt++) {
    loadTab(bt + 2, dat, tnext);           // in flight while this batch is coded
    loadData(bt + 3, dnext);
    const int cnt = min(A1_BATCH, steps + 1 - bt * A1_BATCH);
    if (__shfl_sync(0xFFFFFFFFu, idx, 0) < 8 * A1_BATCH + 8) {      // the coded bytes would run off the front of the chunk's buffer (cannot happen for statistics taken from the data itself)
      if (lane == 0) atomicExch((int*)&blocks[b].status, -KZG_ERR_PROCESS_BLOCK);
      break;
    }
    if (lane == 0) {
      // encodeSymbol (:315-328) for st0..st3 in this order (the bytes of st0 land at the higher address); an uninitialised Symbol
      // (all zero: Java's default object) has xMax 0 -> always emits, and leaves the state alone
      // (branch-free: the two bytes are stored whether or not the state renormalises; when it does not, idx stays and the next
      //  emission overwrites them; a lone warp pays ~20 cycles for every taken branch)
      #define A1_ENC(ST, E) { \
        const bool x = ((i32)ST >= (i32)E.x); \
        pay[idx] = (u8)ST; pay[idx - 1] = (u8)(ST >> 8); \
        idx -= x ? 2 : 0; ST = x ? (u32)((i32)ST >> 16) : ST; \
        ST += (E.z & 0xFFFFu) + (__umulhi(ST, E.y) >> E.w) * (E.z >> 16); }
      #pragma unroll 2
      for (int s = 0; s < cnt; s++) {
        const uint4 e0 = ring[bt & 1][0][s], e1 = ring[bt & 1][1][s], e2 = ring[bt & 1][2][s], e3 = ring[bt & 1][3][s];
        A1_ENC(st0, e0) A1_ENC(st1, e1) A1_ENC(st2, e2) A1_ENC(st3, e3)
      }
      #undef A1_ENC
    }
    __syncwarp();
    storeTab(bt + 1, tv);
    __syncwarp();
    #pragma unroll
    for (int r = 0; r < 4; r++) tv[r] = tnext[r];
    dat[0] = dnext[0]; dat[1] = dnext[1];
  }
  A1_CLK(4)
  if (lane == 0) {
#ifdef KZG_A1_TIMING
    printf("ans1 enc chunk %d/%d (%d bytes): zero %lld hist %lld ctx %lld code %lld cycles\n", b, c, end - start, clk[1] - clk[0], clk[2] - clk[1], clk[3] - clk[2], clk[4] - clk[3]);
#endif
    n = idx + 1;
    BitWriterD bw(hdr);
    bw.nbytes = hdrBits >> 3; bw.nacc = (int)(hdrBits & 7);
    bw.acc = (bw.nacc > 0) ? ((u64)hdr[bw.nbytes] >> (8 - bw.nacc)) : 0;
    write_varint(bw, bufLen - n);
    bw.write(st0, 32); bw.write(st1, 32); bw.write(st2, 32); bw.write(st3, 32);
    bw.flush();
    segs[0] = KzgSeg{hdr, 0, 0, (u64)bw.bits()};
    segs[1] = KzgSeg{pay + n, 0, 0, (u64)(bufLen - n) * 8};
  }
}

// decode, per-chunk global scratch (u32 units): sym[256][256] (freq | cum << 16), f2s[256][2048] (u8), freqAll[256][256] (u16).
// 256 threads: thread 0 walks the 256 context headers (a serial bit stream), every thread then builds one context's cumulative
// frequencies and slot -> symbol table; warp 0 decodes: ONE lane runs the four states as four interleaved dependency chains (see
// the encoder), the coded bytes stream through a shared-memory ring (big-endian words) that all 32 lanes refill a batch of 32
// steps ahead; a step's renormalisation reads come out of one 64-bit window of the ring, no read waits for the one before.
#define A1D_RING 1024
__global__ void __launch_bounds__(A1_THREADS) ans1_decode_kernel(KzgBlock* __restrict__ blocks, KzgEntParams P) {
  __shared__ __align__(16) u32 ring[A1D_RING / 4];        // coded bytes as big-endian words
  __shared__ u8 declared[256];
  __shared__ u8 alphaS[256];
  __shared__ int hdrState[2];          // lr, bad
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const int c = blockIdx.x;
  KzgBlock& B = blocks[b];
  const bool blockOk = (B.status == 0 && B.entropy == P.entropy);
  const int len = blockOk ? B.preLen : 0;
  u8* __restrict__ out = B.cur;
  const u8* __restrict__ stream = P.stream;
  if (len <= 32) {
    if (blockOk && c == 0 && tid == 0) for (int i = 0; i < len; i++) out[i] = (u8)get_bits(stream, (u64)B.srcBit + 8ull * i, 8);
    return;
  }
  const int chunkSize = P.chunkSize;
  if ((i64)c * chunkSize >= (i64)len) return;
  const int start = c * chunkSize;
  const int end = min(start + chunkSize, len);
  const i64 gidx = (i64)b * P.maxChunks + c;
  const KzgChunkInfo info = P.chunks[gidx];
  u32* tab = P.tabBuf + ((P.slotBase ? (i64)P.slotBase[b] : (i64)b * P.maxChunks) + c) * (i64)P.tabStride;
  u32* sym = tab;                               // [256][256] freq | cum << 16
  u8* f2s = (u8*)(tab + 65536);                 // [256][2048]  (logRange <= 11 for order 1)
  u16* freqAll = (u16*)(f2s + 256 * 2048);      // [256][256]

#ifdef KZG_A1_TIMING
  long long clk[5];
#endif
  A1_CLK(0)
  // frequencies do not persist across contexts here: a context's header either declares a symbol or leaves it absent
  for (int k = tid; k < 256 * 256 / 2; k += A1_THREADS) reinterpret_cast<u32*>(freqAll)[k] = 0;
  declared[tid] = 0;
  __syncthreads();
  if (tid == 0) {
    BitReaderD br(stream, (u64)info.hdrBit, (u64)(B.srcBit + B.srcBits));
    int lr = 8 + (int)br.read(3), bad = 0;
    int llr = 3;
    while ((1 << llr) <= lr) llr++;
    if (lr > 11) bad = 1;   // order-1 tables sized for the reference's logRange 11 (ANSRangeEncoder :112,135)
    for (int k = 0; k < 256 && !bad; k++) {
      bool cleared = true;      // (the row is zero already)
      const int as = ans_decode_ctx_header<false>(br, lr, llr, freqAll + k * 256, alphaS, cleared);
      if (as < 0) { bad = 1; break; }
      if (as > 0) declared[k] = 1;
    }
    hdrState[0] = lr; hdrState[1] = bad;
  }
  __syncthreads();
  A1_CLK(1)
  const int lr = hdrState[0];
  if (hdrState[1]) { if (tid == 0) atomicExch(&B.status, -KZG_ERR_PROCESS_BLOCK); return; }
  if (!declared[tid]) {       // a context the header does not declare: all-zero tables, frequency 0 tells the decoder it was used
    for (int i = 0; i < 2048 / 4; i++) reinterpret_cast<u32*>(f2s + tid * 2048)[i] = 0;
    for (int i = 0; i < 256; i++) sym[tid * 256 + i] = 0;
  } else {
    const int k = tid;
    const u16* fr = freqAll + k * 256;
    int sum = 0;
    for (int i = 0; i < 256; i++) {
      const int f = fr[i];
      if (f == 0) continue;
      for (int t = f - 1; t >= 0; t--) f2s[k * 2048 + sum + t] = (u8)i;
      const int fe = (f >= (1 << lr)) ? (1 << lr) - 1 : f;
      sym[k * 256 + i] = (u32)fe | ((u32)sum << 16);
      sum += f;
    }
  }
  __threadfence_block();
  __syncthreads();
  A1_CLK(2)
  if (warp != 0) return;

  // ---- decodeChunkV2 order 1 (:406-432): lane j walks quarter j with state st_j; read order st3, st2, st1, st0 ----
  const int end4 = start + ((end - start) & -4);
  const int quarter = (end4 - start) >> 2;
  u32 st0 = info.st[0], st1 = info.st[1], st2 = info.st[2], st3 = info.st[3];      // lane 0 runs all four states
  const u32 mask = (1u << lr) - 1;
  int cursor = 0;
  u32 prv0 = 0, prv1 = 0, prv2 = 0, prv3 = 0;
  u8* o0 = out + start; u8* o1 = o0 + quarter; u8* o2 = o1 + quarter; u8* o3 = o2 + quarter;
  const i64 payBit = info.payBit;
  const int sz = info.sz;
  u32 minFreq = 1;                        // becomes 0 when a symbol is decoded in a context without a table
  // coded bytes [fill, fill + 256): lane l fetches bytes fill + 8 l .. + 7 (zero beyond sz, as ans_pay16 reads them)
  auto fetch8 = [&](int fill, u32* w) {
    w[0] = 0; w[1] = 0;
    const int o = fill + 8 * lane;
    if (o >= sz) return;
    const u64 bp = (u64)payBit + 8ull * (u64)o;
    const u8* p = stream + (bp >> 3);
    const int sh = (int)(bp & 7);
    u32 prevB = p[0];
    #pragma unroll
    for (int k = 0; k < 8; k++) {
      if (o + k >= sz) break;
      const u32 nextB = p[k + 1];
      w[k >> 2] |= (((prevB << sh) | (nextB >> (8 - sh))) & 0xFFu) << (8 * (k & 3));
      prevB = nextB;
    }
  };
  auto put8 = [&](int fill, const u32* w) {
    *reinterpret_cast<uint2*>(ring + (((fill + 8 * lane) & (A1D_RING - 1)) >> 2)) = make_uint2(__byte_perm(w[0], 0, 0x0123), __byte_perm(w[1], 0, 0x0123));
  };
  u32 w8[2];
  int fill = 0;
  fetch8(0, w8); put8(0, w8); fetch8(256, w8); put8(256, w8);
  fill = 512;
  __syncwarp();
  for (int s0 = 0; s0 < quarter; s0 += 32) {
    const bool doFill = (fill - cursor) <= A1D_RING - 256;     // uniform: cursor is the same in every coding lane... (lanes >= 4 keep it too)
    if (doFill) fetch8(fill, w8);
    const int cnt = min(32, quarter - s0);
    if (lane == 0) {
      // decodeSymbol (:333-347) for the four states, then the reads in the order st3, st2, st1, st0 (:419-430)
      // (all table loads of a step are issued before its output stores: the compiler may not move a load across a store that
      //  could alias it, and a lone warp has nothing else to hide a load behind)
      const u8* __restrict__ f2sR = f2s;
      const u32* __restrict__ symR = sym;
      #pragma unroll 2
      for (int s = 0; s < cnt; s++) {
        // the next eight coded bytes (a step reads at most four 16-bit words)
        const u32 wi = ((u32)cursor & (A1D_RING - 1)) >> 2;
        const u32 w0 = ring[wi], w1 = ring[(wi + 1) & (A1D_RING / 4 - 1)], w2 = ring[(wi + 2) & (A1D_RING / 4 - 1)];
        const u32 sh = ((u32)cursor & 2u) * 8u;
        const u64 win = ((u64)__funnelshift_l(w1, w0, sh) << 32) | (u64)__funnelshift_l(w2, w1, sh);
        const u32 sl0 = st0 & mask, sl1 = st1 & mask, sl2 = st2 & mask, sl3 = st3 & mask;
        const u32 c0 = f2sR[prv0 * 2048 + sl0], c1 = f2sR[prv1 * 2048 + sl1], c2 = f2sR[prv2 * 2048 + sl2], c3 = f2sR[prv3 * 2048 + sl3];
        const u32 f0 = symR[prv0 * 256 + c0], f1 = symR[prv1 * 256 + c1], f2 = symR[prv2 * 256 + c2], f3 = symR[prv3 * 256 + c3];
        o0[s0 + s] = (u8)c0; o1[s0 + s] = (u8)c1; o2[s0 + s] = (u8)c2; o3[s0 + s] = (u8)c3;
        minFreq = min(min(minFreq, f0 & 0xFFFFu), min(f1 & 0xFFFFu, min(f2 & 0xFFFFu, f3 & 0xFFFFu)));
        st0 = (f0 & 0xFFFFu) * (st0 >> lr) + sl0 - (f0 >> 16); st1 = (f1 & 0xFFFFu) * (st1 >> lr) + sl1 - (f1 >> 16);
        st2 = (f2 & 0xFFFFu) * (st2 >> lr) + sl2 - (f2 >> 16); st3 = (f3 & 0xFFFFu) * (st3 >> lr) + sl3 - (f3 >> 16);
        prv0 = c0; prv1 = c1; prv2 = c2; prv3 = c3;
        const u32 n3 = ((i32)st3 < ANS_TOP) ? 16u : 0u, n2 = ((i32)st2 < ANS_TOP) ? 16u : 0u;
        const u32 n1 = ((i32)st1 < ANS_TOP) ? 16u : 0u, n0 = ((i32)st0 < ANS_TOP) ? 16u : 0u;
        const u32 b2 = n3, b1 = n3 + n2, b0 = b1 + n1;            // bits of the window consumed before each state's read
        if (n3) st3 = (st3 << 16) | (u32)(win >> 48);
        if (n2) st2 = (st2 << 16) | ((u32)(win >> (48 - b2)) & 0xFFFFu);
        if (n1) st1 = (st1 << 16) | ((u32)(win >> (48 - b1)) & 0xFFFFu);
        if (n0) st0 = (st0 << 16) | ((u32)(win >> (48 - b0)) & 0xFFFFu);
        cursor += (int)((b0 + n0) >> 3);
      }
      #undef A1_DEC
    }
    cursor = __shfl_sync(0xFFFFFFFFu, cursor, 0);
    __syncwarp();
    if (doFill) { put8(fill, w8); fill += 256; }
    __syncwarp();
  }
  A1_CLK(3)
  if (lane == 0) {
#ifdef KZG_A1_TIMING
    printf("ans1 dec chunk %d/%d (%d bytes): headers %lld tables %lld decode %lld cycles\n", b, c, end - start, clk[1] - clk[0], clk[2] - clk[1], clk[3] - clk[2]);
#endif
    const int tail = end - end4;
    for (int i = 0; i < tail; i++) out[end4 + i] = (cursor + i < sz) ? (u8)ans_pay8(stream, payBit, cursor + i) : 0;
    if (cursor + tail != sz || minFreq == 0) atomicExch(&B.status, -KZG_ERR_PROCESS_BLOCK);
  }
}

// ================================================================================================================
// host launchers
// ================================================================================================================
int kzg_ans_encode_launch(cudaStream_t s, const KzgBlock* d_blocks, int nBlocks, const KzgEntParams& P, int order) {
  if (order == 0) {
    CUDA_TRY(cudaFuncSetAttribute(ans0_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(A0EncSmem)));   // per device: set on every launch
    dim3 grid((P.maxChunks + A0_GROUPS - 1) / A0_GROUPS, nBlocks);
    KZG_PROF("ans0_encode_kernel", s, (ans0_encode_kernel<<<grid, 128, sizeof(A0EncSmem), s>>>(d_blocks, P)));
  } else {
    dim3 grid(P.maxChunks, nBlocks);
    KZG_PROF("ans1_encode_kernel", s, (ans1_encode_kernel<<<grid, A1_THREADS, 0, s>>>(d_blocks, P)));
  }
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(1);
  return 0;
}

int kzg_ans_decode_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, const KzgEntParams& P, int order, bool withScan) {
  if (withScan) {
    if (order == 0) KZG_PROF("ans0_scan_kernel", s, (ans0_scan_kernel<<<nBlocks, 128, 0, s>>>(d_blocks, nBlocks, P)));
    else ans_scan_kernel<<<(nBlocks + ANS_SCAN_WARPS - 1) / ANS_SCAN_WARPS, 32 * ANS_SCAN_WARPS, 0, s>>>(d_blocks, nBlocks, P, order);
    CUDA_TRY(cudaGetLastError());
  }
  if (order == 0) {
    dim3 grid((P.maxChunks + A0D_CHUNKS - 1) / A0D_CHUNKS, nBlocks);
    CUDA_TRY(cudaFuncSetAttribute(ans0_decode_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));   // five 42 KiB CTAs per SM
    KZG_PROF("ans0_decode_kernel", s, (ans0_decode_kernel<<<grid, 32, 0, s>>>(d_blocks, P)));
  } else {
    dim3 grid(P.maxChunks, nBlocks);
    KZG_PROF("ans1_decode_kernel", s, (ans1_decode_kernel<<<grid, A1_THREADS, 0, s>>>(d_blocks, P)));
  }
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(2);
  return 0;
}

size_t kzg_ans1_enc_tab_u32() { return 4 * 65536 + 256 * 257 + 64 + 256 * A1_HDR_WORDS + 256 * 256 / 4 + 16; }
size_t kzg_ans1_dec_tab_u32() { return 65536 + (256 * 2048) / 4 + (256 * 256) / 2 + 64 + 16; }
