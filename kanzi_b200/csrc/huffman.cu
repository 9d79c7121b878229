// huffman.cu — canonical Huffman encode / decode kernels (sm_100a).
//
// Replaces K/entropy/HuffmanEncoder.java, HuffmanDecoder.java, HuffmanCommon.java:
// 16 KiB chunks, code lengths <= 12, per chunk: alphabet bitmap, signed Exp-Golomb length deltas, four
// varint bit counts, four independently coded fragments of count/4 symbols, count%4 raw bytes.
// Unit of parallel work fixed by the format: chunk x 4 fragments -> 4 lanes per chunk, 16 chunks per CTA;
// code/decoding tables live in shared memory.  The length computation (Moffat-Katajainen in place,
// limitCodeLengths) is order-sensitive and is restated literally on one lane per chunk.
#include "kzg_common.cuh"
#include "kzg_entropy.cuh"

#define HF_GROUPS 16
#define HF_MAXLEN 12
#define HF_CHUNK 16384
#define HF_FRAG_STRIDE 6160     // 4096 symbols x 12 bits = 6144 bytes + slack

__device__ void hf_encode_alphabet(BitWriterD& bw, const u8* alphabet, int count);   // below
__device__ int hf_normalize(u32* freqs, u8* alphabet, int n, int totalFreq, int scale, int* err);

// signed Exp-Golomb of a code-length delta (K/entropy/ExpGolombEncoder.java:123-132; the table there holds
// '0' x log2, '1', then log2+1 bits of ((|v|+1-2^log2) << 1 | sign))
__device__ __forceinline__ void hf_expgolomb(BitWriterD& bw, int val) {
  if (val == 0) { bw.write(1, 1); return; }
  const int a = (val < 0) ? -val : val;
  const int lg = ilog2((u32)(a + 1));
  const u32 tail = ((u32)(a + 1 - (1 << lg)) << 1) | (u32)(val < 0);
  bw.write((1u << (lg + 1)) | tail, 2 * lg + 2);
}
// HuffmanEncoder.computeInPlaceSizesPhase1/2 (:317-376)
__device__ void hf_phase1(int* data, int n) {
  for (int s = 0, r = 0, t = 0; t < n - 1; t++) {
    int sum = 0;
    for (int i = 0; i < 2; i++) {
      if ((s >= n) || ((r < t) && (data[r] < data[s]))) { sum += data[r]; data[r] = t; r++; continue; }
      sum += data[s];
      if (s > t) data[s] = 0;
      s++;
    }
    data[t] = sum;
  }
}
__device__ int hf_phase2(int* data, int n) {
  if (n < 2) return 0;
  int levelTop = n - 2, depth = 1, i = n, totalNodesAtLevel = 2;
  while (i > 0) {
    int k = levelTop;
    while ((k > 0) && (data[k - 1] >= levelTop)) k--;
    const int internalNodesAtLevel = levelTop - k;
    const int leavesAtLevel = totalNodesAtLevel - internalNodesAtLevel;
    for (int j = 0; j < leavesAtLevel; j++) data[--i] = depth;
    totalNodesAtLevel = internalNodesAtLevel << 1;
    levelTop = k;
    depth++;
  }
  return depth - 1;
}

// HuffmanEncoder.computeCodeLengths (:285-308).  ranks in: (freq << 8) | symbol; out: symbols sorted by
// (freq, symbol).  work: `count` ints.  sizes indexed by symbol.
__device__ int hf_code_lengths(u8* sizes, int* ranks, int* work, int count, bool presorted = false) {
  // Arrays.sort(ranks, 0, count): shell sort (keys are distinct); the encode kernel sorts with all its threads beforehand
  for (int gap = presorted ? 0 : (1 << 7); gap > 0; gap >>= 1) {
    for (int i = gap; i < count; i++) {
      const int t = ranks[i];
      int k = i;
      while (k >= gap && ranks[k - gap] > t) { ranks[k] = ranks[k - gap]; k -= gap; }
      ranks[k] = t;
    }
  }
  for (int i = 0; i < count; i++) {
    work[i] = (int)((u32)ranks[i] >> 8);
    ranks[i] &= 0xFF;
    if (work[i] == 0) return 0;
  }
  hf_phase1(work, count);
  const int maxCodeLen = hf_phase2(work, count);
  for (int i = 0; i < count; i++) sizes[ranks[i]] = (u8)min(work[i], 255);
  return maxCodeLen;
}

// HuffmanEncoder.limitCodeLengths (:191-273).  lists: 6 x 256 bytes (FIFO of symbols per size delta).
__device__ int hf_limit_lengths(const u8* alphabet, u32* freqs, u8* sizes, int* ranks, int* work, u8* lists, int count, int* err) {
  int n = 0, debt = 0;
  while (n < 256 && sizes[ranks[n]] >= HF_MAXLEN) {
    debt += (sizes[ranks[n]] - HF_MAXLEN);
    sizes[ranks[n]] = HF_MAXLEN;
    n++;
  }
  int head[6] = {0, 0, 0, 0, 0, 0}, tail[6] = {0, 0, 0, 0, 0, 0};
  while (n < count) {
    const int idx = HF_MAXLEN - 1 - sizes[ranks[n]];
    if ((idx >= 6) || (debt < (1 << idx))) break;
    lists[idx * 256 + tail[idx]++] = (u8)ranks[n];
    n++;
  }
  int idx = 5;
  while ((debt > 0) && (idx >= 0)) {
    if ((head[idx] >= tail[idx]) || (debt < (1 << idx))) { idx--; continue; }
    const int r = lists[idx * 256 + head[idx]++];
    sizes[r]++;
    debt -= (1 << idx);
  }
  idx = 0;
  while ((debt > 0) && (idx < 6)) {
    if (head[idx] >= tail[idx]) { idx++; continue; }
    const int r = lists[idx * 256 + head[idx]++];
    sizes[r]++;
    debt -= (1 << idx);
  }
  if (debt > 0) {
    // slow path: renormalise the frequencies to scale 2048 and recompute (:247-269).  f[] is compacted
    // (index i = i-th alphabet symbol); reuse `work` + 256 as f and lists as the index alphabet.
    u32* f = (u32*)(work + 256);
    int totalFreq = 0;
    for (int i = 0; i < count; i++) { f[i] = freqs[alphabet[i]]; totalFreq += (int)f[i]; }
    hf_normalize(f, lists, count, totalFreq, HF_CHUNK >> 3, err);
    if (*err) return 0;
    for (int i = 0; i < count; i++) {
      freqs[alphabet[i]] = f[i];
      ranks[i] = (int)((f[i] << 8) | alphabet[i]);
    }
    return hf_code_lengths(sizes, ranks, work, count);
  }
  return HF_MAXLEN;
}

// EntropyUtils.normalizeFrequencies (K/entropy/EntropyUtils.java:141-250) over an n-entry frequency array
// (n = alphabet.length in the Java call).  The totalFreq == scale shortcut scans 256 entries in Java and
// would throw for n < 256: reported through *err.
__device__ int hf_normalize(u32* freqs, u8* alphabet, int n, int totalFreq, int scale, int* err) {
  if (n == 0 || totalFreq == 0) return 0;
  int alphabetSize = 0;
  if (totalFreq == scale) {
    if (n < 256) { *err = 1; return 0; }
    for (int i = 0; i < 256; i++) if (freqs[i] != 0) alphabet[alphabetSize++] = (u8)i;
    return alphabetSize;
  }
  int sumScaledFreq = 0, sumFreq = 0, idxMax = 0;
  for (int i = 0; i < n; i++) {
    const int f = (int)freqs[i];
    if (f == 0) continue;
    const u64 sf = (u64)f * (u64)scale;
    const int scaledFreq = (sf <= (u64)totalFreq) ? 1 : (int)((sf + ((u64)totalFreq >> 1)) / (u64)totalFreq);
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
__device__ void hf_encode_alphabet(BitWriterD& bw, const u8* alphabet, int count) {
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

// HuffmanCommon.generateCanonicalCodes (HuffmanCommon.java:71-111): symbols sorted by (size, symbol), codes counted up, shifted left at
// every size step.  All threads of a CTA at once: in (size, symbol) order the code of a symbol is the Kraft sum of the symbols before
// it, sum 2^(12 - size_t), shifted down to its own size (sizes ascend in that order, so the sum is a multiple of 2^(12 - size)).
// codes[s] = the sum itself (= the symbol's first slot in the 4096-entry decoding table).  Returns false for a present symbol
// whose size is outside 1..12 (generateCanonicalCodes would not place it).
__device__ __forceinline__ bool hf_kraft_prefix(const u8* sizes, const u8* present, u32* kraft, int tid, int nthreads) {
  bool ok = true;
  for (int sy = tid; sy < 256; sy += nthreads) {
    if (!present[sy]) continue;
    const int sz = sizes[sy];
    if (sz < 1 || sz > HF_MAXLEN) { ok = false; continue; }
    u32 k = 0;
    for (int t = 0; t < 256; t++) {
      const int st = sizes[t];
      if (present[t] && st >= 1 && st <= HF_MAXLEN && (st < sz || (st == sz && t < sy))) k += 1u << (HF_MAXLEN - st);
    }
    kraft[sy] = k;
  }
  return ok;
}

// ================================================================================================================
// encode: one CTA of 128 threads per chunk (grid (maxChunks, nBlocks)); 6 segments per chunk.
//   1. the chunk is staged in shared memory while it is counted (per-warp histograms, merged);
//   2. thread 0 computes the code lengths and canonical codes (order-sensitive, restated literally) and writes the header;
//   3. warp j packs fragment j: every lane sums the code lengths of its 1/32 of the fragment, a warp scan gives its first bit,
//      then it packs its symbols into the fragment's bit string in shared memory (whole 32-bit words with plain stores, the two
//      boundary words with atomicOr);
//   4. the four bit strings go out with coalesced word stores.
// ================================================================================================================
#define HFE_THREADS 128
struct HfEncSmem {
  u32 freq[4][256];               // per-warp histograms; freq[0] = merged
  u32 codes[256];                 // (len << 24) | code
  int ranks[256];
  int work[512];
  u8 sizes[256];
  u8 alpha[256];
  u8 present[256];
  u8 lists[6 * 256];
  __align__(16) u8 data[HF_CHUNK];
  u32 bits[4][HF_FRAG_STRIDE / 4];
  int nsym, err, direct; i64 hdrBits; u32 fragBits[4];
  u64 bar;                        // mbarrier of the chunk's bulk copy
};

__global__ void __launch_bounds__(HFE_THREADS) huff_encode_kernel(const KzgBlock* __restrict__ blocks, KzgEntParams P) {
  extern __shared__ __align__(16) u8 smem_raw[];
  HfEncSmem& S = *reinterpret_cast<HfEncSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const int c = blockIdx.x;
  const KzgBlock& B = blocks[b];
  const int len = (B.status == 0 && B.entropy == P.entropy) ? B.curLen : 0;
  const u8* __restrict__ data = B.cur;
  const i64 gidx = (i64)b * P.maxChunks + c;
  KzgSeg* segs = P.segs + (i64)b * P.segsPerBlock + 1 + (i64)c * 6;
  const int start = c * HF_CHUNK;
  if (start >= len) {
    if (tid < 6) segs[tid] = KzgSeg{nullptr, 0, 0, 0};
    return;
  }
  const int count = min(HF_CHUNK, len - start);
  if (count < 32) {     // small chunk stored raw (:400-402)
    if (tid == 0) {
      segs[0] = KzgSeg{data + start, 0, 0, (u64)count * 8};
      for (int k = 1; k < 6; k++) segs[k] = KzgSeg{nullptr, 0, 0, 0};
    }
    return;
  }
  u8* hdr = P.hdrBuf + gidx * (i64)P.hdrStride;
  u8* pay = P.payBuf + gidx * (i64)P.payStride;

  // the chunk comes in with ONE bulk asynchronous copy (cp.async.bulk, completion on an mbarrier) when its address allows (16-byte
  // aligned: always, for buffers of this library and for caller blocks on 16-byte boundaries); the histograms are cleared meanwhile
  const int bulk = (((uintptr_t)(data + start) & 15) == 0) ? (count & ~15) : 0;
  if (tid == 0) {
    kzg_mbar_init(&S.bar, 1);
    if (bulk) kzg_bulk_g2s(S.data, data + start, (u32)bulk, &S.bar);
  }
  for (int k = tid; k < 4 * 256; k += HFE_THREADS) (&S.freq[0][0])[k] = 0;
  for (int i = bulk + tid; i < count; i += HFE_THREADS) S.data[i] = data[start + i];
  __syncthreads();
  if (bulk) kzg_mbar_wait(&S.bar, 0);
  for (int i = tid * 4; i < count; i += HFE_THREADS * 4) {        // four symbols per thread and step out of shared memory
    const u32 w = *reinterpret_cast<const u32*>(S.data + i);
    const int nv = min(4, count - i);
    #pragma unroll
    for (int k = 0; k < 4; k++) if (k < nv) atomicAdd(&S.freq[warp][(w >> (8 * k)) & 0xFF], 1u);
  }
  __syncthreads();
  for (int k = tid; k < 256; k += HFE_THREADS) S.freq[0][k] += S.freq[1][k] + S.freq[2][k] + S.freq[3][k];
  __syncthreads();

  // Arrays.sort of (freq << 8 | symbol) over the present symbols (HuffmanEncoder.java:288): keys are distinct, so a key's place is the
  // number of smaller keys (every thread places two symbols)
  for (int sy = tid; sy < 256; sy += HFE_THREADS) {
    const u32 f = S.freq[0][sy];
    S.present[sy] = f > 0;
    if (f == 0) continue;
    const u32 key = (f << 8) | (u32)sy;
    int rk = 0;
    for (int t = 0; t < 256; t++) { const u32 ft = S.freq[0][t]; rk += (ft > 0 && ((ft << 8) | (u32)t) < key) ? 1 : 0; }
    S.ranks[rk] = (int)key;
  }
  __syncthreads();
  if (tid == 0) {   // updateFrequencies (:103-178)
    int nsym = 0, err = 0, direct = 0;
    BitWriterD bw(hdr);
    u8* alphabet = S.alpha; u8* sizes = S.sizes; u32* codes = S.codes; int* ranks = S.ranks;
    for (int i = 0; i < 256; i++) {
      codes[i] = 0; sizes[i] = 0;
      if (S.present[i]) alphabet[nsym++] = (u8)i;
    }
    hf_encode_alphabet(bw, alphabet, nsym);
    for (int i = nsym; i < 256; i++) ranks[i] = 0;       // (the reference's array is zero beyond the alphabet: limitCodeLengths scans it)
    if (nsym == 1) {
      sizes[alphabet[0]] = 1;       // (code 0)
    } else {
      int maxCodeLen = hf_code_lengths(sizes, ranks, S.work, nsym, true);
      if (maxCodeLen == 0) err = 1;
      if (!err && maxCodeLen > HF_MAXLEN) {
        maxCodeLen = hf_limit_lengths(alphabet, S.freq[0], sizes, ranks, S.work, S.lists, nsym, &err);
        if (maxCodeLen == 0) err = 1;
      }
      if (!err && maxCodeLen > HF_MAXLEN) {      // unlikely fallback (:146-155): 8-bit codes = the symbol's rank in the alphabet
        for (int i = 0; i < nsym; i++) { codes[alphabet[i]] = (u32)i; sizes[alphabet[i]] = 8; }
        direct = 1;
      }
    }
    if (!err) {
      int prevSize = 2;
      for (int i = 0; i < nsym; i++) {
        const int currSize = sizes[alphabet[i]];
        hf_expgolomb(bw, currSize - prevSize);
        prevSize = currSize;
      }
    }
    S.hdrBits = bw.bits();
    bw.flush();
    S.nsym = nsym; S.err = err; S.direct = direct;
  }
  __syncthreads();
  // canonical codes (HuffmanCommon.generateCanonicalCodes :71-111), all threads; codes[] = (size << 24) | code
  if (!S.err && S.nsym > 1) {
    if (!S.direct) {
      if (!hf_kraft_prefix(S.sizes, S.present, S.codes, tid, HFE_THREADS)) S.err = 1;
      __syncthreads();
    }
    for (int sy = tid; sy < 256; sy += HFE_THREADS) {
      if (!S.present[sy]) continue;
      const u32 sz = S.sizes[sy];
      S.codes[sy] = (sz << 24) | (S.direct ? S.codes[sy] : (S.codes[sy] >> (HF_MAXLEN - sz)));
    }
  }
  __syncthreads();
  const int nsym = S.nsym;
  if (S.err) {
    if (tid == 0) {
      atomicExch((int*)&blocks[b].status, -KZG_ERR_PROCESS_BLOCK);
      for (int k = 0; k < 6; k++) segs[k] = KzgSeg{nullptr, 0, 0, 0};
    }
    return;
  }
  if (nsym <= 1) {      // chunk skipped after its header (:407-409)
    if (tid == 0) {
      segs[0] = KzgSeg{hdr, 0, 0, (u64)S.hdrBits};
      for (int k = 1; k < 6; k++) segs[k] = KzgSeg{nullptr, 0, 0, 0};
    }
    return;
  }

  // ---- encodeChunk (:419-493): warp j packs fragment j ----
  const int szFrag = count / 4;
  {
    const int j = warp;
    u32* fb = S.bits[j];
    for (int k = lane; k < HF_FRAG_STRIDE / 4; k += 32) fb[k] = 0;
    __syncwarp();
    const u8* p = S.data + j * szFrag;
    const int per = (szFrag + 31) >> 5;
    const int i0 = min(lane * per, szFrag), i1 = min(i0 + per, szFrag);
    u32 myBits = 0;
    for (int i = i0; i < i1; i++) myBits += S.codes[p[i]] >> 24;
    u32 incl = myBits;
    for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
    const u32 total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    u32 pos = incl - myBits;                       // first bit of this lane's symbols (bit 0 = MSB of byte 0)
    // pack: `acc` holds `nacc` pending bits that belong at bit `pos - nacc`
    u64 acc = 0; int nacc = 0;
    u32 wpos = pos;                                // bit position of the first pending bit
    for (int i = i0; i < i1; i++) {
      const u32 code = S.codes[p[i]];
      const int cl = (int)(code >> 24);
      acc = (acc << cl) | (u64)(code & 0xFFFFFF);
      nacc += cl;
      if (nacc >= 32) {
        // emit the 32 oldest pending bits at bit wpos: two words when wpos is not word aligned
        const u32 v = (u32)(acc >> (nacc - 32));
        const u32 wi = wpos >> 5, sh = wpos & 31;
        if (sh == 0) fb[wi] = __byte_perm(v, 0, 0x0123);       // a lane's interior words are its own
        else { atomicOr(&fb[wi], __byte_perm(v >> sh, 0, 0x0123)); atomicOr(&fb[wi + 1], __byte_perm(v << (32 - sh), 0, 0x0123)); }
        nacc -= 32; wpos += 32;
      }
    }
    if (nacc > 0) {
      const u32 v = (u32)(acc << (32 - nacc));     // left-aligned remainder
      const u32 wi = wpos >> 5, sh = wpos & 31;
      atomicOr(&fb[wi], __byte_perm(v >> sh, 0, 0x0123));
      if (sh && sh + nacc > 32) atomicOr(&fb[wi + 1], __byte_perm(v << (32 - sh), 0, 0x0123));
    }
    __syncwarp();
    if (lane == 0) S.fragBits[j] = total;
    u32* out = reinterpret_cast<u32*>(pay + j * HF_FRAG_STRIDE);
    for (u32 k = lane; k < (total + 31) / 32; k += 32) out[k] = fb[k];
  }
  __syncthreads();
  if (tid == 0) {
    const u32 b0 = S.fragBits[0], b1 = S.fragBits[1], b2 = S.fragBits[2], b3 = S.fragBits[3];
    const i64 hdrBits = S.hdrBits;
    BitWriterD bw(hdr);
    bw.nbytes = hdrBits >> 3; bw.nacc = (int)(hdrBits & 7);
    bw.acc = (bw.nacc > 0) ? ((u64)hdr[bw.nbytes] >> (8 - bw.nacc)) : 0;
    write_varint(bw, (i32)b0); write_varint(bw, (i32)b1); write_varint(bw, (i32)b2); write_varint(bw, (i32)b3);
    bw.flush();
    segs[0] = KzgSeg{hdr, 0, 0, (u64)bw.bits()};
    segs[1] = KzgSeg{pay + 0 * HF_FRAG_STRIDE, 0, 0, b0};
    segs[2] = KzgSeg{pay + 1 * HF_FRAG_STRIDE, 0, 0, b1};
    segs[3] = KzgSeg{pay + 2 * HF_FRAG_STRIDE, 0, 0, b2};
    segs[4] = KzgSeg{pay + 3 * HF_FRAG_STRIDE, 0, 0, b3};
    segs[5] = KzgSeg{data + start + 4 * szFrag, 0, 0, (u64)(count - 4 * szFrag) * 8};
  }
}

// ================================================================================================================
// decode
// ================================================================================================================
// chunk scan: one warp per block, lane 0 walks the chunk headers (HuffmanDecoder.decodeV6 :353-390 walks chunks the same way; a
// chunk starts where the one before it ends, and only its header says where that is).  The walk is instruction bound (a lone warp),
// so the code lengths are skipped out of a 32-bit register window: a signed Exp-Golomb code is `1`, or z zeros, `1`, z + 1 bits —
// one count-leading-zeros per code instead of a bit read per bit.
__device__ __forceinline__ u32 hf_peek32(const u8* __restrict__ stream, u64 pos, u64 end) {      // the next 32 bits, zeros beyond `end`
  if (pos + 32 <= end) return get_bits(stream, pos, 32);
  if (pos >= end) return 0u;
  const int rem = (int)(end - pos);
  return get_bits(stream, pos, rem) << (32 - rem);
}
__global__ void __launch_bounds__(32) huff_scan_kernel(KzgBlock* __restrict__ blocks, int nBlocks, KzgEntParams P) {
  const int b = blockIdx.x;
  if (b >= nBlocks || threadIdx.x != 0) return;
  KzgBlock& B = blocks[b];
  if (B.status != 0 || B.entropy != P.entropy) return;
  const int len = B.preLen;
  KzgChunkInfo* ci = P.chunks + (i64)b * P.maxChunks;
  const u8* __restrict__ stream = P.stream;
  const u64 end = (u64)(B.srcBit + B.srcBits);
  BitReaderD br(stream, (u64)B.srcBit, end);
  const int nChunks = (len + HF_CHUNK - 1) / HF_CHUNK;
  for (int c = 0; c < nChunks; c++) {
    const int count = min(HF_CHUNK, len - c * HF_CHUNK);
    KzgChunkInfo info;
    info.hdrBit = (i64)br.pos; info.payBit = 0; info.sz = 0; info.alphabetSize = 0;
    info.st[0] = info.st[1] = info.st[2] = info.st[3] = 0;
    if (count < 32) {
      br.pos += (u64)count * 8;
    } else {
      int n = 0;
      if (br.read(1) == 0) n = (br.read(1) == 1) ? 0 : 256;
      else {
        const int nb = (int)br.read(5) + 1;                     // mask bytes
        for (int i = 0; i < nb; i += 4) { const int k = min(4, nb - i); n += __popc(br.read(8 * k)); }
      }
      if (n == 0) { B.status = -KZG_ERR_PROCESS_BLOCK; return; }      // readLengths() <= 0 -> decode returns early
      u64 pos = br.pos;
      u32 w = 0; int avail = 0;
      for (int i = 0; i < n; i++) {
        if (avail < 24) { w = hf_peek32(stream, pos, end); avail = 32; }
        const int z = __clz(w);
        if (z > 10) { B.status = -KZG_ERR_PROCESS_BLOCK; return; }    // (a length delta never needs more; all-zero = past the end)
        const int cl = z ? 2 * z + 2 : 1;
        pos += (u64)cl; w <<= cl; avail -= cl;
      }
      br.pos = pos;
      info.alphabetSize = n;
      if (n > 1) {
        u64 total = 0;
        for (int k = 0; k < 4; k++) {
          const i32 v = read_varint(br);
          if (v < 0 || v > 8 * HF_FRAG_STRIDE) { B.status = -KZG_ERR_PROCESS_BLOCK; return; }
          info.st[k] = (u32)v; total += (u64)v;
        }
        info.payBit = (i64)br.pos;
        br.pos += total + (u64)(count - 4 * (count / 4)) * 8;
      }
    }
    if (br.overrun()) { B.status = -KZG_ERR_PROCESS_BLOCK; return; }
    ci[c] = info;
  }
  B.entBits = (i64)br.pos - B.srcBit;
}

// decode: one CTA of 128 threads per chunk.  Thread 0 reads the code lengths (a serial bit stream) while the other warps stage the
// four fragments' bit strings in shared memory (word-aligned at the fragment's first bit, big-endian words); all threads fill the
// 4096-entry table; lanes 0-3 of warp 0 decode one fragment each out of shared memory into shared memory (no global access
// on the dependent chain); the chunk goes out with coalesced stores.
#define HFD_THREADS 128
struct HfDecSmem {
  u16 table[1 << HF_MAXLEN];
  u32 codes[256];
  u8 sizes[256];
  u8 alpha[256];
  u8 present[256];
  u32 bits[4][HF_FRAG_STRIDE / 4];
  u8 out[HF_CHUNK];
  int nsym, bad;
};

__global__ void __launch_bounds__(HFD_THREADS) huff_decode_kernel(KzgBlock* __restrict__ blocks, KzgEntParams P) {
  extern __shared__ __align__(16) u8 smem_raw[];
  HfDecSmem& S = *reinterpret_cast<HfDecSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int c = blockIdx.x;
  KzgBlock& B = blocks[b];
  if (!(B.status == 0 && B.entropy == P.entropy)) return;
  const int len = B.preLen;
  const int start = c * HF_CHUNK;
  if (start >= len) return;
  const int count = min(HF_CHUNK, len - start);
  u8* __restrict__ out = B.cur + start;
  const u8* __restrict__ stream = P.stream;
  const KzgChunkInfo info = P.chunks[(i64)b * P.maxChunks + c];
  if (count < 32) {
    for (int i = tid; i < count; i += HFD_THREADS) out[i] = (u8)get_bits(stream, (u64)info.hdrBit + 8ull * i, 8);
    return;
  }
  const int szFrag = count / 4;
  // ---- readLengths (:115-154) on thread 0; the fragments' bits are staged meanwhile ----
  if (tid == 0) {
    int nsym = 0, bad = 0;
    BitReaderD br(stream, (u64)info.hdrBit, (u64)(B.srcBit + B.srcBits));
    u8* alphabet = S.alpha;
    for (int i = 0; i < 256; i++) S.present[i] = 0;
    if (br.read(1) == 0) { if (br.read(1) == 0) { nsym = 256; for (int i = 0; i < 256; i++) alphabet[i] = (u8)i; } }
    else {
      const int lastMask = (int)br.read(5);
      for (int i = 0; i <= lastMask; i++) {
        const u32 m = br.read(8);
        for (int k = 0; k < 8; k++) if (m & (1u << k)) alphabet[nsym++] = (u8)((i << 3) + k);
      }
    }
    int curSize = 2;
    {   // ExpGolombDecoder.decodeByte (:41-60) out of a 32-bit register window (see huff_scan_kernel)
      const u64 end = (u64)(B.srcBit + B.srcBits);
      u64 pos = br.pos;
      u32 w = 0; int avail = 0;
      for (int i = 0; i < nsym; i++) {
        const int sy = alphabet[i];
        if (avail < 24) { w = hf_peek32(stream, pos, end); avail = 32; }
        const int z = __clz(w);
        if (z > 10) { bad = 1; break; }
        int delta = 0, cl = 1;
        if (z) {
          cl = 2 * z + 2;
          i64 res = (i64)((w >> (32 - cl)) & ((2u << z) - 1));       // the z + 1 bits behind the `1`
          const i64 sgn = res & 1;
          res = (res >> 1) + (1 << z) - 1;
          delta = (int)(int8_t)((res - sgn) ^ -sgn);
        }
        pos += (u64)cl; w <<= cl; avail -= cl;
        curSize += delta;
        if ((curSize <= 0) || (curSize > HF_MAXLEN)) { bad = 1; break; }
        S.sizes[sy] = (u8)curSize;
        S.present[sy] = 1;
      }
      if (pos > end) bad = 1;
    }
    S.nsym = nsym; S.bad = bad;
  } else if (tid >= 32 && info.alphabetSize > 1) {
    // fragment j: bits [pos_j, pos_j + st_j) of the stream; bits at or beyond the fragment's end read as zero (the Java buffer
    // is zero-filled, :414-415)
    u64 pos = (u64)info.payBit;
    for (int j = 0; j < 4; j++) {
      const u32 nb = info.st[j];
      const u32 nw = (nb + 31) >> 5;
      for (u32 k = tid - 32; k < nw; k += HFD_THREADS - 32) {
        u32 w = get_bits(stream, pos + 32ull * k, 32);
        const u32 left = nb - 32 * k;
        if (left < 32) w &= ~(0xFFFFFFFFu >> left);
        S.bits[j][k] = w;
      }
      for (u32 k = nw + (tid - 32); k < nw + 2 && k < HF_FRAG_STRIDE / 4; k += HFD_THREADS - 32) S.bits[j][k] = 0;
      pos += nb;
    }
  }
  __syncthreads();
  const int nsym = S.nsym;
  if (S.bad || nsym == 0) { if (tid == 0) atomicExch(&B.status, -KZG_ERR_PROCESS_BLOCK); return; }
  if (nsym == 1) {
    const u8 v = S.alpha[0];
    for (int i = tid; i < count; i += HFD_THREADS) out[i] = v;
    return;
  }
  for (int i = tid; i < (1 << HF_MAXLEN); i += HFD_THREADS) S.table[i] = 7;
  // canonical codes, all threads: codes[] = Kraft prefix = the symbol's first table slot (see hf_kraft_prefix)
  if (!hf_kraft_prefix(S.sizes, S.present, S.codes, tid, HFD_THREADS)) S.bad = 1;
  __syncthreads();
  if (S.bad) { if (tid == 0) atomicExch(&B.status, -KZG_ERR_PROCESS_BLOCK); return; }
  // buildDecodingTables (:162-191) walks the alphabet re-sorted by (size, symbol) (generateCanonicalCodes sorts it in place), so
  // its running `length` is the symbol's own size: idx = code << (12 - size).
  for (int i = tid; i < nsym; i += HFD_THREADS) {
    const int sy = S.alpha[i];
    const int sz = S.sizes[sy];
    const u16 val = (u16)((sz << 8) | sy);
    const int idx0 = (int)S.codes[sy];
    const int cnt = 1 << (HF_MAXLEN - sz);
    for (int k = 0; k < cnt; k++) if (idx0 + k < (1 << HF_MAXLEN)) S.table[idx0 + k] = val;
  }
  __syncthreads();

  // ---- decodeChunk (:404-587): lane j decodes fragment j ----
  if (tid < 4) {
    const int j = tid;
    const u32* __restrict__ fb = S.bits[j];
    u8* o = S.out + j * szFrag;
    u64 win = 0; int avail = 0;        // `avail` valid bits at the bottom of win
    int consumed = 0, wi = 0;
    for (int i = 0; i < szFrag; i++) {
      if (avail < HF_MAXLEN) { win = (win << 32) | (u64)fb[min(wi, HF_FRAG_STRIDE / 4 - 1)]; wi++; avail += 32; }
      const u32 idx = (u32)(win >> (avail - HF_MAXLEN)) & ((1u << HF_MAXLEN) - 1);
      const u32 val = S.table[idx];
      const int cl = (int)(val >> 8);
      avail -= cl;
      consumed += cl;
      o[i] = (u8)val;
    }
    if (consumed != (int)info.st[j]) atomicExch(&B.status, -KZG_ERR_PROCESS_BLOCK);
  }
  __syncthreads();
  for (int i = tid; i < 4 * szFrag; i += HFD_THREADS) out[i] = S.out[i];
  if (tid == 0) {
    u64 tpos = (u64)info.payBit + info.st[0] + info.st[1] + info.st[2] + info.st[3];
    for (int i = 4 * szFrag; i < count; i++, tpos += 8) out[i] = (u8)get_bits(stream, tpos, 8);
  }
}

int kzg_huff_encode_launch(cudaStream_t s, const KzgBlock* d_blocks, int nBlocks, const KzgEntParams& P) {
  CUDA_TRY(cudaFuncSetAttribute(huff_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HfEncSmem)));   // per device: set on every launch
  dim3 grid(P.maxChunks, nBlocks);
  KZG_PROF("huff_encode_kernel", s, (huff_encode_kernel<<<grid, HFE_THREADS, sizeof(HfEncSmem), s>>>(d_blocks, P)));
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(1);
  return 0;
}

int kzg_huff_decode_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, const KzgEntParams& P) {
  KZG_PROF("huff_scan_kernel", s, (huff_scan_kernel<<<nBlocks, 32, 0, s>>>(d_blocks, nBlocks, P)));
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaFuncSetAttribute(huff_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HfDecSmem)));
  dim3 grid(P.maxChunks, nBlocks);
  KZG_PROF("huff_decode_kernel", s, (huff_decode_kernel<<<grid, HFD_THREADS, sizeof(HfDecSmem), s>>>(d_blocks, P)));
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(2);
  return 0;
}
