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
// K/entropy/ExpGolombDecoder.java:41-60
__device__ __forceinline__ int hf_expgolomb_dec(BitReaderD& br) {
  if (br.read(1) == 1) return 0;
  int lg = 1;
  while (br.read(1) == 0) { lg++; if (lg > 30 || br.overrun()) return 127; }
  i64 res = (i64)br.read(lg + 1);
  const i64 sgn = res & 1;
  res = (res >> 1) + (1 << lg) - 1;
  return (int)(int8_t)((res - sgn) ^ -sgn);
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
__device__ int hf_code_lengths(u8* sizes, int* ranks, int* work, int count) {
  // Arrays.sort(ranks, 0, count): shell sort (keys are distinct)
  for (int gap = 1 << 7; gap > 0; gap >>= 1) {
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

// HuffmanCommon.generateCanonicalCodes (HuffmanCommon.java:71-111): symbols sorted by (size, symbol)
__device__ int hf_canonical_codes(const u8* sizes, u32* codes, const u8* present, int count, u8* order) {
  int n = 0;
  for (int len = 1; len <= HF_MAXLEN && n < count; len++)
    for (int s = 0; s < 256 && n < count; s++)
      if (present[s] && sizes[s] == len) order[n++] = (u8)s;
  if (n != count) return -1;
  int code = 0, curLen = sizes[order[0]];
  for (int i = 0; i < count; i++) {
    const int s = order[i];
    code <<= (sizes[s] - curLen);
    curLen = sizes[s];
    codes[s] = (u32)code;
    code++;
  }
  return count;
}

// ================================================================================================================
// encode: grid (ceil(maxChunks/16), nBlocks), 64 threads; 4 lanes (= 4 fragments) per chunk; 6 segments per chunk
// ================================================================================================================
struct HfEncSmem {
  u32 freq[HF_GROUPS][256];
  u32 codes[HF_GROUPS][256];      // (len << 24) | code
  int ranks[HF_GROUPS][256];
  int work[HF_GROUPS][512];
  u8 sizes[HF_GROUPS][256];
  u8 alpha[HF_GROUPS][256];
  u8 present[HF_GROUPS][256];
  u8 lists[HF_GROUPS][6 * 256];
};

__global__ void __launch_bounds__(64) huff_encode_kernel(const KzgBlock* __restrict__ blocks, KzgEntParams P) {
  extern __shared__ __align__(16) u8 smem_raw[];
  HfEncSmem& S = *reinterpret_cast<HfEncSmem*>(smem_raw);
  const int g = threadIdx.x >> 2, j = threadIdx.x & 3;
  const int b = blockIdx.y;
  const int c = blockIdx.x * HF_GROUPS + g;
  if (c >= P.maxChunks) return;     // whole 4-lane groups leave; no warp-wide collectives below
  const KzgBlock& B = blocks[b];
  const int len = (B.status == 0 && B.entropy == P.entropy) ? B.curLen : 0;
  const u8* __restrict__ data = B.cur;
  const i64 gidx = (i64)b * P.maxChunks + c;
  KzgSeg* segs = P.segs + (i64)b * P.segsPerBlock + 1 + (i64)c * 6;
  const int start = c * HF_CHUNK;
  const u32 gmask = 0xFu << ((threadIdx.x & 31) & ~3);
  if (start >= len) {
    if (j == 0) for (int k = 0; k < 6; k++) segs[k] = KzgSeg{nullptr, 0, 0, 0};
    return;
  }
  const int count = min(HF_CHUNK, len - start);
  if (count < 32) {     // small chunk stored raw (:400-402)
    if (j == 0) {
      segs[0] = KzgSeg{data + start, 0, 0, (u64)count * 8};
      for (int k = 1; k < 6; k++) segs[k] = KzgSeg{nullptr, 0, 0, 0};
    }
    return;
  }
  u8* hdr = P.hdrBuf + gidx * (i64)P.hdrStride;
  u8* pay = P.payBuf + gidx * (i64)P.payStride;

  for (int k = j; k < 256; k += 4) S.freq[g][k] = 0;
  __syncwarp(gmask);
  for (int i = start + j; i < start + count; i += 4) atomicAdd(&S.freq[g][data[i]], 1u);
  __syncwarp(gmask);

  int nsym = 0, err = 0;
  i64 hdrBits = 0;
  if (j == 0) {   // updateFrequencies (:103-178)
    BitWriterD bw(hdr);
    u8* alphabet = S.alpha[g]; u8* sizes = S.sizes[g]; u32* codes = S.codes[g]; int* ranks = S.ranks[g];
    for (int i = 0; i < 256; i++) {
      codes[i] = 0; sizes[i] = 0; S.present[g][i] = 0;
      if (S.freq[g][i] > 0) { alphabet[nsym++] = (u8)i; S.present[g][i] = 1; }
    }
    hf_encode_alphabet(bw, alphabet, nsym);
    if (nsym == 1) {
      codes[alphabet[0]] = 0;     // code value of (1 << 24) masks to 0; the length goes in below
      sizes[alphabet[0]] = 1;
    } else {
      for (int i = 0; i < 256; i++) ranks[i] = 0;
      for (int i = 0; i < nsym; i++) ranks[i] = (int)((S.freq[g][alphabet[i]] << 8) | alphabet[i]);
      int maxCodeLen = hf_code_lengths(sizes, ranks, S.work[g], nsym);
      if (maxCodeLen == 0) err = 1;
      if (!err && maxCodeLen > HF_MAXLEN) {
        maxCodeLen = hf_limit_lengths(alphabet, S.freq[g], sizes, ranks, S.work[g], S.lists[g], nsym, &err);
        if (maxCodeLen == 0) err = 1;
      }
      if (!err) {
        if (maxCodeLen > HF_MAXLEN) {      // unlikely fallback (:146-155)
          for (int i = 0; i < nsym; i++) { codes[alphabet[i]] = (u32)i; sizes[alphabet[i]] = 8; }
        } else {
          if (hf_canonical_codes(sizes, codes, S.present[g], nsym, S.lists[g]) < 0) err = 1;
        }
      }
    }
    if (!err) {
      int prevSize = 2;
      for (int i = 0; i < nsym; i++) {
        const int s = alphabet[i];
        const int currSize = sizes[s];
        codes[s] |= ((u32)currSize << 24);
        hf_expgolomb(bw, currSize - prevSize);
        prevSize = currSize;
      }
    }
    hdrBits = bw.bits();
    bw.flush();
  }
  nsym = __shfl_sync(gmask, nsym, (threadIdx.x & 31) & ~3);
  err = __shfl_sync(gmask, err, (threadIdx.x & 31) & ~3);
  __syncwarp(gmask);
  if (err) {
    if (j == 0) {
      atomicExch((int*)&blocks[b].status, -KZG_ERR_PROCESS_BLOCK);
      for (int k = 0; k < 6; k++) segs[k] = KzgSeg{nullptr, 0, 0, 0};
    }
    return;
  }
  if (nsym <= 1) {      // chunk skipped after its header (:407-409)
    if (j == 0) {
      segs[0] = KzgSeg{hdr, 0, 0, (u64)hdrBits};
      for (int k = 1; k < 6; k++) segs[k] = KzgSeg{nullptr, 0, 0, 0};
    }
    return;
  }

  // ---- encodeChunk (:419-493): lane j packs fragment j ----
  const int szFrag = count / 4;
  u8* fb = pay + j * HF_FRAG_STRIDE;
  u64 acc = 0; int nacc = 0; int nb = 0;
  const u8* __restrict__ p = data + start + j * szFrag;
  for (int i = 0; i < szFrag; i++) {
    const u32 code = S.codes[g][p[i]];
    const int cl = (int)(code >> 24);
    acc = (acc << cl) | (u64)(code & 0xFFFFFF);
    nacc += cl;
    while (nacc >= 8) { nacc -= 8; fb[nb++] = (u8)(acc >> nacc); }
  }
  const u32 myBits = (u32)(nb * 8 + nacc);
  if (nacc > 0) fb[nb] = (u8)(acc << (8 - nacc));
  const int gl = (threadIdx.x & 31) & ~3;
  const u32 b0 = __shfl_sync(gmask, myBits, gl + 0), b1 = __shfl_sync(gmask, myBits, gl + 1);
  const u32 b2 = __shfl_sync(gmask, myBits, gl + 2), b3 = __shfl_sync(gmask, myBits, gl + 3);
  if (j == 0) {
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
// chunk scan: one thread per block (HuffmanDecoder.decodeV6 :353-390 walks chunks the same way)
__global__ void huff_scan_kernel(KzgBlock* __restrict__ blocks, int nBlocks, KzgEntParams P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nBlocks) return;
  KzgBlock& B = blocks[b];
  if (B.status != 0 || B.entropy != P.entropy) return;
  const int len = B.preLen;
  KzgChunkInfo* ci = P.chunks + (i64)b * P.maxChunks;
  BitReaderD br(P.stream, (u64)B.srcBit, (u64)(B.srcBit + B.srcBits));
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
      else { const int lastMask = (int)br.read(5); for (int i = 0; i <= lastMask; i++) n += __popc(br.read(8)); }
      if (n == 0) { B.status = -KZG_ERR_PROCESS_BLOCK; return; }      // readLengths() <= 0 -> decode returns early
      for (int i = 0; i < n; i++) {
        if (br.read(1) == 1) continue;
        int lg = 1;
        while (br.read(1) == 0) { lg++; if (br.overrun() || lg > 30) { B.status = -KZG_ERR_PROCESS_BLOCK; return; } }
        br.pos += (u64)(lg + 1);
      }
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

struct HfDecSmem {
  u16 table[HF_GROUPS][1 << HF_MAXLEN];
  u32 codes[HF_GROUPS][256];
  u8 sizes[HF_GROUPS][256];
  u8 alpha[HF_GROUPS][256];
  u8 present[HF_GROUPS][256];
  u8 order[HF_GROUPS][256];
};

__global__ void __launch_bounds__(64) huff_decode_kernel(KzgBlock* __restrict__ blocks, KzgEntParams P) {
  extern __shared__ __align__(16) u8 smem_raw[];
  HfDecSmem& S = *reinterpret_cast<HfDecSmem*>(smem_raw);
  const int g = threadIdx.x >> 2, j = threadIdx.x & 3;
  const int b = blockIdx.y;
  const int c = blockIdx.x * HF_GROUPS + g;
  if (c >= P.maxChunks) return;
  KzgBlock& B = blocks[b];
  if (!(B.status == 0 && B.entropy == P.entropy)) return;
  const int len = B.preLen;
  const int start = c * HF_CHUNK;
  if (start >= len) return;
  const int count = min(HF_CHUNK, len - start);
  u8* __restrict__ out = B.cur + start;
  const u8* __restrict__ stream = P.stream;
  const KzgChunkInfo info = P.chunks[(i64)b * P.maxChunks + c];
  const u32 gmask = 0xFu << ((threadIdx.x & 31) & ~3);
  const int gl = (threadIdx.x & 31) & ~3;
  if (count < 32) {
    for (int i = j; i < count; i += 4) out[i] = (u8)get_bits(stream, (u64)info.hdrBit + 8ull * i, 8);
    return;
  }
  // ---- readLengths (:115-154) + buildDecodingTables (:162-191), lane 0; table fill shared by the 4 lanes ----
  int nsym = 0, bad = 0;
  if (j == 0) {
    BitReaderD br(stream, (u64)info.hdrBit, (u64)(B.srcBit + B.srcBits));
    u8* alphabet = S.alpha[g];
    for (int i = 0; i < 256; i++) S.present[g][i] = 0;
    if (br.read(1) == 0) { if (br.read(1) == 0) { nsym = 256; for (int i = 0; i < 256; i++) alphabet[i] = (u8)i; } }
    else {
      const int lastMask = (int)br.read(5);
      for (int i = 0; i <= lastMask; i++) {
        const u32 m = br.read(8);
        for (int k = 0; k < 8; k++) if (m & (1u << k)) alphabet[nsym++] = (u8)((i << 3) + k);
      }
    }
    int curSize = 2;
    for (int i = 0; i < nsym; i++) {
      const int s = alphabet[i];
      curSize += hf_expgolomb_dec(br);
      if ((curSize <= 0) || (curSize > HF_MAXLEN)) { bad = 1; break; }
      S.sizes[g][s] = (u8)curSize;
      S.present[g][s] = 1;
    }
    if (!bad && nsym > 1) {
      if (hf_canonical_codes(S.sizes[g], S.codes[g], S.present[g], nsym, S.order[g]) < 0) bad = 1;
    }
  }
  nsym = __shfl_sync(gmask, nsym, gl);
  bad = __shfl_sync(gmask, bad, gl);
  __syncwarp(gmask);
  if (bad || nsym == 0) { if (j == 0) atomicExch(&B.status, -KZG_ERR_PROCESS_BLOCK); return; }
  if (nsym == 1) {
    const u8 v = S.alpha[g][0];
    for (int i = j; i < count; i += 4) out[i] = v;
    return;
  }
  for (int i = j; i < (1 << HF_MAXLEN); i += 4) S.table[g][i] = 7;
  __syncwarp(gmask);
  {
    // buildDecodingTables (:162-191) walks the alphabet re-sorted by (size, symbol) (generateCanonicalCodes
    // sorts it in place), so its running `length` is the symbol's own size: idx = code << (12 - size).
    for (int i = j; i < nsym; i += 4) {
      const int s = S.alpha[g][i];
      const int sz = S.sizes[g][s];
      const u16 val = (u16)((sz << 8) | s);
      const int idx0 = (int)(S.codes[g][s] << (HF_MAXLEN - sz));
      const int cnt = 1 << (HF_MAXLEN - sz);
      for (int k = 0; k < cnt; k++) if (idx0 + k < (1 << HF_MAXLEN)) S.table[g][idx0 + k] = val;
    }
  }
  __syncwarp(gmask);

  // ---- decodeChunk (:404-587): lane j decodes fragment j ----
  const int szFrag = count / 4;
  u64 pos = (u64)info.payBit;
  for (int k = 0; k < j; k++) pos += info.st[k];
  const u64 fragEnd = pos + info.st[j];
  u8* o = out + j * szFrag;
  u64 win = 0; int avail = 0;        // `avail` valid bits at the bottom of win
  int consumed = 0;
  for (int i = 0; i < szFrag; i++) {
    if (avail < HF_MAXLEN) {
      // refill 32 bits; bits at or beyond fragEnd read as zero (the Java buffer is zero-filled, :414-415)
      u32 w = 0;
      if (pos < fragEnd) {
        w = get_bits(stream, pos, 32);
        const u64 left = fragEnd - pos;
        if (left < 32) w &= ~((1u << (32 - (int)left)) - 1);
      }
      win = (win << 32) | (u64)w;
      avail += 32;
      pos += 32;
    }
    const u32 idx = (u32)(win >> (avail - HF_MAXLEN)) & ((1u << HF_MAXLEN) - 1);
    const u32 val = S.table[g][idx];
    const int cl = (int)(val >> 8);
    avail -= cl;
    consumed += cl;
    o[i] = (u8)val;
  }
  const int okFrag = (consumed == (int)info.st[j]);
  if (!okFrag) atomicExch(&B.status, -KZG_ERR_PROCESS_BLOCK);
  if (j == 0) {
    u64 tpos = (u64)info.payBit + info.st[0] + info.st[1] + info.st[2] + info.st[3];
    for (int i = 4 * szFrag; i < count; i++, tpos += 8) out[i] = (u8)get_bits(stream, tpos, 8);
  }
}

int kzg_huff_encode_launch(cudaStream_t s, const KzgBlock* d_blocks, int nBlocks, const KzgEntParams& P) {
  CUDA_TRY(cudaFuncSetAttribute(huff_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HfEncSmem)));   // per device: set on every launch
  dim3 grid((P.maxChunks + HF_GROUPS - 1) / HF_GROUPS, nBlocks);
  huff_encode_kernel<<<grid, 64, sizeof(HfEncSmem), s>>>(d_blocks, P);
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(1);
  return 0;
}

int kzg_huff_decode_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, const KzgEntParams& P) {
  huff_scan_kernel<<<(nBlocks + 31) / 32, 32, 0, s>>>(d_blocks, nBlocks, P);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaFuncSetAttribute(huff_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HfDecSmem)));
  dim3 grid((P.maxChunks + HF_GROUPS - 1) / HF_GROUPS, nBlocks);
  huff_decode_kernel<<<grid, 64, sizeof(HfDecSmem), s>>>(d_blocks, P);
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(2);
  return 0;
}
