// ORACLE — TEST INFRASTRUCTURE ONLY (see kz_core.hpp header).  extern "C" surface of the CPU
// restatement for ctypes (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference).
// parity unpinned (no JVM in the image; reference holds no golden bitstreams).
#include "kz_stream.hpp"
#include <thread>
#include <atomic>
#include <mutex>

using namespace kzo;

extern "C" {

// ---- entropy stage ---------------------------------------------------------------------------------
// returns n on success, <0 on error; *outBits = bit length of the MSB-first string written at out[0..]
int64_t kzo_entropy_encode(int type, const uint8_t* src, int32_t n, uint8_t* out, int64_t outCap, int64_t* outBits) {
  try {
    BitWriter bs;
    const int r = entropyEncode(type, bs, src, n);
    const i64 bits = (i64)bs.written();
    bs.close();
    if ((i64)bs.buf.size() > outCap) return -2;
    memcpy(out, bs.buf.data(), bs.buf.size());
    *outBits = bits;
    return r;
  } catch (std::exception&) { return -1; }
}

// returns the codec's return value (n on success); *bitsUsed = bits consumed
int32_t kzo_entropy_decode(int type, const uint8_t* in, int64_t inBits, uint8_t* dst, int32_t n, int64_t* bitsUsed) {
  try {
    std::vector<u8> padded((size_t)((inBits + 7) >> 3) + 16, 0);
    memcpy(padded.data(), in, (size_t)((inBits + 7) >> 3));
    BitReader bs(padded.data(), (u64)inBits);
    std::vector<u8> outv((size_t)n + 16, 0);
    const int r = entropyDecode(type, bs, outv.data(), n);
    memcpy(dst, outv.data(), n);
    if (bitsUsed) *bitsUsed = (i64)bs.read();
    return r;
  } catch (std::exception&) { return -1; }
}

// ---- transform stage -------------------------------------------------------------------------------
// One ByteTransform call with SliceByteArray semantics: src slice = (array of srcCap bytes, length
// srcLen, index 0), dst slice = (array of dstCap bytes, length dstLen, index 0).
// Returns 1 (true) / 0 (false) / <0 (Java exception).  ctxv = {bsVersion, blockSize, size, jobs, dataType, bwtBounds}
int kzo_transform(int type, int inverse, int32_t* ctxv, const uint8_t* src, int32_t srcLen, int32_t srcCap,
                  uint8_t* dst, int32_t dstLen, int32_t dstCap, int32_t* srcUsed, int32_t* dstUsed) {
  try {
    Ctx ctx; ctx.bsVersion = ctxv[0]; ctx.blockSize = ctxv[1]; ctx.size = ctxv[2]; ctx.jobs = ctxv[3]; ctx.dataType = ctxv[4]; ctx.bwtBounds = ctxv[5] & 0xFF; ctx.entropyType = (ctxv[5] >> 8) & 0xFF;
    std::vector<u8> sv(src, src + srcCap), dv((size_t)dstCap, 0);
    Slice s(&sv, srcLen, 0), d(&dv, dstLen, 0);
    std::unique_ptr<Transform> t = newTransform(ctx, type);
    const bool ok = inverse ? t->inverse(s, d) : t->forward(s, d);
    memcpy(dst, dv.data(), std::min((size_t)dstCap, dv.size()));
    *srcUsed = s.index; *dstUsed = d.index;
    ctxv[4] = ctx.dataType;
    return ok ? 1 : 0;
  } catch (std::exception&) { return -1; }
}

int32_t kzo_transform_max_encoded_len(int type, int32_t n) {
  try { Ctx ctx; return newTransform(ctx, type)->getMaxEncodedLength(n); } catch (std::exception&) { return -1; }
}

// whole Sequence (chain of <= 8 ids) the way EncodingTask / DecodingTask drive it:
// forward: returns post-transform length, *skipFlags out.  inverse: skipFlags in.
int32_t kzo_sequence_forward(const int32_t* ids, int nIds, int32_t blockSize, int bwtBounds, const uint8_t* src, int32_t n,
                             uint8_t* dst, int32_t dstCap, int32_t* skipFlags) {
  try {
    Ctx ctx; ctx.blockSize = blockSize; ctx.size = n; ctx.bwtBounds = bwtBounds & 0xFF; ctx.entropyType = (bwtBounds >> 8) & 0xFF;
    int idv[8]; for (int i = 0; i < nIds; i++) idv[i] = ids[i];
    Sequence seq(ctx, transformTypeOf(idv, nIds));
    EncodeBuffers eb(std::max(blockSize, n));
    memcpy(eb.data.p(), src, n);
    const int req = seq.getMaxEncodedLength(n);
    eb.buffer.length = req; eb.bufArr.assign(req, 0);
    eb.data.length = n;
    seq.forward(eb.data, eb.buffer);
    const int post = eb.buffer.index;
    if (post > dstCap) return -2;
    memcpy(dst, eb.bufArr.data(), post);
    *skipFlags = seq.skipFlags;
    return post;
  } catch (std::exception&) { return -1; }
}

// ---- raw BWT (as TestBWT drives it: index 0, dst sized exactly n) --------------------------------------
int kzo_bwt_forward(const uint8_t* src, int32_t n, uint8_t* dst, int32_t* primaryIndexes8) {
  try {
    std::vector<u8> sv(src, src + n), dv((size_t)n, 0);
    Slice s(&sv, n, 0), d(&dv, n, 0);
    BWT bwt;
    const bool ok = bwt.forward(s, d);
    memcpy(dst, dv.data(), n);
    for (int i = 0; i < 8; i++) primaryIndexes8[i] = bwt.primaryIndexes[i];
    return ok ? 1 : 0;
  } catch (std::exception&) { return -1; }
}
// algo: 0 = as BWT.inverse picks (mergeTPSI <= 8 MiB < biPSIv2), 1 = force mergeTPSI, 2 = force biPSIv2
int kzo_bwt_inverse(const uint8_t* src, int32_t n, uint8_t* dst, const int32_t* primaryIndexes8, int algo) {
  try {
    // dst array carries slack like the stream decoder's buffers do (CIS:694-695): biPSIv2 writes up to
    // 8*ceil(n/8) bytes when n % 8 != 0 (BWT.java:612-640)
    std::vector<u8> sv(src, src + n), dv((size_t)n + 16, 0);
    Slice s(&sv, n, 0), d(&dv, n, 0);
    BWT bwt; bwt.asref = false;
    for (int i = 0; i < 8; i++) bwt.primaryIndexes[i] = primaryIndexes8[i];
    bool ok;
    if (algo == 1) ok = bwt.check(s, d) && bwt.inverseMergeTPSI(s, d, n);
    else if (algo == 2) ok = bwt.check(s, d) && bwt.inverseBiPSIv2(s, d, n);
    else ok = bwt.inverse(s, d);
    memcpy(dst, dv.data(), n);
    return ok ? 1 : 0;
  } catch (std::exception&) { return -1; }
}

// ---- whole streams ---------------------------------------------------------------------------------------
static StreamParams mkParams(const int32_t* ids, int nIds, int entropyType, int32_t blockSize, int64_t inputSize, int bwtBounds) {
  StreamParams sp;
  int idv[8]; for (int i = 0; i < nIds; i++) idv[i] = ids[i];
  sp.transformType = transformTypeOf(idv, nIds);
  sp.entropyType = entropyType; sp.blockSize = blockSize; sp.inputSize = inputSize; sp.bwtBounds = bwtBounds & 0xFF;
  sp.checksum = (bwtBounds >> 8) & 0xFF;       // (callers pass bwtBounds | checksum << 8: 0, 32 or 64)
  return sp;
}

// single-threaded, exactly the emulated CompressedOutputStream; returns byte length or <0
int64_t kzo_compress_stream(const uint8_t* in, int64_t n, const int32_t* ids, int nIds, int entropyType, int32_t blockSize,
                            int64_t inputSize, int bwtBounds, uint8_t* out, int64_t outCap) {
  try {
    StreamParams sp = mkParams(ids, nIds, entropyType, blockSize, inputSize, bwtBounds);
    std::vector<u8> r = compressStream(in, n, sp);
    if ((i64)r.size() > outCap) return -2;
    memcpy(out, r.data(), r.size());
    return (i64)r.size();
  } catch (std::exception&) { return -1; }
}

int64_t kzo_decompress_stream(const uint8_t* in, int64_t nBytes, int bwtBounds, uint8_t* out, int64_t outCap) {
  try {
    std::vector<u8> r = decompressStream(in, nBytes, bwtBounds);
    if ((i64)r.size() > outCap) return -2;
    memcpy(out, r.data(), r.size());
    return (i64)r.size();
  } catch (std::exception&) { return -1; }
}

// Block records only (what one EncodingTask appends to the shared bitstream), block-parallel on
// `nthreads` host threads the way the Java host runs EncodingTasks on its pool (COS:537-573).
// recBits[b] = bit length of block b's record (5-bit lw + length + payload); records are written
// byte-aligned at recOff[b] in `out`.  Returns number of blocks or <0.
int32_t kzo_encode_blocks_mt(const uint8_t* in, int64_t n, const int32_t* ids, int nIds, int entropyType, int32_t blockSize,
                             int bwtBounds, int nthreads, uint8_t* out, int64_t outCap, int64_t* recOff, int64_t* recBits, int32_t maxBlocks) {
  const int nb = (int)((n + blockSize - 1) / blockSize);
  if (nb > maxBlocks) return -2;
  StreamParams sp = mkParams(ids, nIds, entropyType, blockSize, n, bwtBounds);
  std::vector<std::vector<u8>> recs(nb);
  std::vector<i64> bits(nb, 0);
  std::atomic<int> next(0); std::atomic<int> err(0);
  auto work = [&]() {
    EncodeBuffers eb(blockSize);
    for (;;) {
      const int b = next.fetch_add(1);
      if (b >= nb) break;
      try {
        const i64 off = (i64)b * blockSize;
        const int len = (int)std::min<i64>(blockSize, n - off);
        memcpy(eb.data.p(), in + off, len);
        eb.data.index = 0;
        BitWriter obs;
        encodeBlock(obs, eb, len, sp);
        bits[b] = (i64)obs.written();
        obs.close();
        recs[b].swap(obs.buf);
      } catch (std::exception&) { err = 1; }
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < std::max(1, nthreads); t++) th.emplace_back(work);
  for (auto& t : th) t.join();
  if (err) return -1;
  i64 off = 0;
  for (int b = 0; b < nb; b++) {
    if (off + (i64)recs[b].size() > outCap) return -2;
    memcpy(out + off, recs[b].data(), recs[b].size());
    recOff[b] = off; recBits[b] = bits[b];
    off += (i64)recs[b].size();
  }
  return nb;
}

// Decode block records produced above (each byte-aligned at recOff[b]) block-parallel; out gets the
// blocks back to back at b*blockSize.  Returns total decoded bytes or <0.
int64_t kzo_decode_blocks_mt(const uint8_t* in, const int64_t* recOff, const int64_t* recBits, int32_t nb, const int32_t* ids, int nIds,
                             int entropyType, int32_t blockSize, int bwtBounds, int nthreads, uint8_t* out, int64_t outCap) {
  StreamHeader h; h.bsVersion = 7; h.entropyType = entropyType; h.blockSize = blockSize;
  int idv[8]; for (int i = 0; i < nIds; i++) idv[i] = ids[i];
  h.transformType = transformTypeOf(idv, nIds);
  std::atomic<int> next(0); std::atomic<int> err(0); std::atomic<i64> total(0);
  auto work = [&]() {
    for (;;) {
      const int b = next.fetch_add(1);
      if (b >= nb) break;
      try {
        std::vector<u8> padded((size_t)((recBits[b] + 7) >> 3) + 16, 0);
        memcpy(padded.data(), in + recOff[b], (size_t)((recBits[b] + 7) >> 3));
        BitReader ibs(padded.data(), (u64)recBits[b]);
        std::vector<u8> o;
        const int d = decodeBlock(ibs, h, o, bwtBounds);
        if ((i64)b * blockSize + d > outCap) { err = 1; continue; }
        memcpy(out + (i64)b * blockSize, o.data(), d);
        total += d;
      } catch (std::exception&) { err = 1; }
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < std::max(1, nthreads); t++) th.emplace_back(work);
  for (auto& t : th) t.join();
  return err ? -1 : (i64)total;
}

// stream header only (Appendix D KATs)
int32_t kzo_stream_header(const int32_t* ids, int nIds, int entropyType, int32_t blockSize, int64_t inputSize, uint8_t* out, int32_t outCap) {
  StreamParams sp = mkParams(ids, nIds, entropyType, blockSize, inputSize, 1);
  BitWriter obs; writeStreamHeader(obs, sp); obs.close();
  if ((int)obs.buf.size() > outCap) return -2;
  memcpy(out, obs.buf.data(), obs.buf.size());
  return (int)obs.buf.size();
}

// small helpers exposed for unit tests
int32_t kzo_normalize_frequencies(int32_t* freqs256, int32_t* alphabet256, int32_t totalFreq, int32_t scale) {
  try { return normalizeFrequencies(freqs256, alphabet256, 256, totalFreq, scale); } catch (std::exception&) { return -1; }
}
int32_t kzo_expgolomb_signed(int8_t v, uint32_t* bits) {
  BitWriter bs; expGolombEncodeSigned(bs, v); const int nb = (int)bs.written(); bs.close();
  u32 r = 0; for (size_t i = 0; i < bs.buf.size(); i++) r = (r << 8) | bs.buf[i];
  r >>= (bs.buf.size() * 8 - nb);
  *bits = r; return nb;
}
uint32_t kzo_xxhash32(const uint8_t* d, int32_t n, uint32_t seed) { return xxhash32(d, n, seed); }
uint64_t kzo_xxhash64(const uint8_t* d, int32_t n, uint64_t seed) { return kanzi_xxhash64(d, n, seed); }
int kzo_abi_version() { return 1; }

}  // extern "C"
