// ORACLE — TEST INFRASTRUCTURE ONLY (see kz_core.hpp header).  parity unpinned for whole-stream bytes
// beyond the hand-derived container KATs of SURVEY Appendix D.
//
// Host emulator: what CompressedOutputStream / CompressedInputStream and their EncodingTask /
// DecodingTask do around the hot path, restated for jobs == 1, no listeners, checksum none,
// so whole .knz streams can be produced and consumed without a JVM.
#pragma once
#include "kz_core.hpp"
#include "kz_entropy.hpp"
#include "kz_transforms.hpp"

namespace kzo {

enum { BITSTREAM_TYPE = 0x4B414E5A, BITSTREAM_FORMAT_VERSION = 7, COPY_BLOCK_MASK = 0x80, TRANSFORMS_MASK = 0x10,
       MIN_BITSTREAM_BLOCK_SIZE = 1024, MAX_BITSTREAM_BLOCK_SIZE = 1 << 30, SMALL_BLOCK_SIZE = 15, EXTRA_BUFFER_SIZE = 512 };

// mix32, CompressedOutputStream.java:89-93
static inline i32 mix32(i32 checksum, i32 hash, i32 value) {
  checksum ^= jmul(hash, ~value);
  checksum = rotl32(checksum, 13);
  return jadd(jmul(checksum, 5), 0x52DCE729);
}

// Magic.getType / isCompressed / isMultimedia / isExecutable, Magic.java:154-258
namespace magic {
enum : i32 { NO_MAGIC = 0, JPG = (i32)0xFFD8FFE0, GIF = 0x47494638, PDF = 0x25504446, ZIP = 0x504B0304, LZMA = 0x377ABCAF,
             PNG = (i32)0x89504E47, ELF = 0x7F454C46, MAC32 = (i32)0xFEEDFACE, CIGAM32 = (i32)0xCEFAEDFE, MAC64 = (i32)0xFEEDFACF,
             CIGAM64 = (i32)0xCFFAEDFE, ZSTD = 0x28B52FFD, BROTLI = (i32)0x81CFB2CE, RIFF = 0x52494646, CAB = 0x4D534346,
             FLAC = 0x664C6143, XZ = (i32)0xFD377A58, RAR = 0x52617221, KNZ = 0x4B414E5A, BZIP2 = 0x425A68, MP3_ID3 = 0x494433,
             GZIP = 0x1F8B, BMP = 0x424D, WIN = 0x4D5A, PBM = 0x5034, PGM = 0x5035, PPM = 0x5036 };
static inline i32 getType(const u8* src, int srcLen, int start) {
  if (srcLen < 4) return NO_MAGIC;
  const i32 key = (i32)be32(src + start);
  if ((key & ~0x0F) == JPG) return key;
  if (((key >> 8) == BZIP2) || ((key >> 8) == MP3_ID3)) return key >> 8;
  static const i32 KEYS32[] = {GIF, PDF, ZIP, LZMA, PNG, ELF, MAC32, CIGAM32, MAC64, CIGAM64, ZSTD, BROTLI, CAB, RIFF, FLAC, XZ, KNZ, RAR};
  for (i32 k : KEYS32) if (key == k) return key;
  const i32 key16 = key >> 16;
  static const i32 KEYS16[] = {GZIP, BMP, WIN};
  for (i32 k : KEYS16) if (key16 == k) return key16;
  if ((key16 == PBM) || (key16 == PGM) || (key16 == PPM)) {
    const int subkey = (key >> 8) & 0xFF;
    if ((subkey == 0x07) || (subkey == 0x0A) || (subkey == 0x0D) || (subkey == 0x20)) return key16;
  }
  return NO_MAGIC;
}
static inline bool isCompressed(i32 m) {
  switch (m) { case JPG: case GIF: case PNG: case LZMA: case ZSTD: case BROTLI: case CAB: case ZIP: case GZIP: case BZIP2:
               case FLAC: case MP3_ID3: case XZ: case KNZ: case RAR: return true; default: return false; }
}
static inline bool isMultimedia(i32 m) {
  switch (m) { case JPG: case GIF: case PNG: case RIFF: case FLAC: case MP3_ID3: case BMP: case PBM: case PGM: case PPM: return true;
               default: return false; }
}
static inline bool isExecutable(i32 m) {
  switch (m) { case ELF: case WIN: case MAC32: case CIGAM32: case MAC64: case CIGAM64: return true; default: return false; }
}
}  // namespace magic

// TransformFactory.getType (TransformFactory.java:132-164): first transform in the top 6-bit slot
static inline u64 transformTypeOf(const int* ids, int n) {      // TransformFactory.getType (:140-153): NONE tokens are skipped
  u64 res = 0; int shift = 42;
  for (int i = 0; i < n && i < 8; i++) { if (ids[i] == T_NONE) continue; res |= ((u64)ids[i] << shift); shift -= 6; }
  return res;
}

// ---- block checksums: K/util/hash/XXHash32.java:94-142 and XXHash64.java (seed = BITSTREAM_TYPE, COS:195-200) -----------------
// XXHash32 is the published XXH32.  Kanzi's XXHash64 is NOT the published XXH64: the four accumulators are folded with the
// 32-bit rotation idiom `(v << 1) | (v >>> 31)` applied to 64-bit values (XXHash64.java:127-128), restated as written.
static inline u32 xxh32_round(u32 acc, u32 val) { acc += val * 2246822519u; return ((acc << 13) | (acc >> 19)) * 2654435761u; }
static inline u32 xxhash32(const u8* data, int length, u32 seed) {
  const u32 P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
  const int end = length;
  u32 h32; int idx = 0;
  if (length >= 16) {
    const int end16 = end - 16;
    u32 v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
    do {
      v1 = xxh32_round(v1, le32(data + idx)); v2 = xxh32_round(v2, le32(data + idx + 4));
      v3 = xxh32_round(v3, le32(data + idx + 8)); v4 = xxh32_round(v4, le32(data + idx + 12));
      idx += 16;
    } while (idx <= end16);
    h32 = ((v1 << 1) | (v1 >> 31)) + ((v2 << 7) | (v2 >> 25)) + ((v3 << 12) | (v3 >> 20)) + ((v4 << 18) | (v4 >> 14));
  } else h32 = seed + P5;
  h32 += (u32)length;
  while (idx <= end - 4) { h32 += le32(data + idx) * P3; h32 = ((h32 << 17) | (h32 >> 15)) * P4; idx += 4; }
  while (idx < end) { h32 += (u32)data[idx] * P5; h32 = ((h32 << 11) | (h32 >> 21)) * P1; idx++; }
  h32 ^= h32 >> 15; h32 *= P2; h32 ^= h32 >> 13; h32 *= P3;
  return h32 ^ (h32 >> 16);
}
static inline u64 xxh64_round(u64 acc, u64 val) { acc += val * 0xC2B2AE3D27D4EB4Full; return ((acc << 31) | (acc >> 33)) * 0x9E3779B185EBCA87ull; }
static inline u64 xxh64_merge(u64 acc, u64 val) { acc ^= xxh64_round(0, val); return acc * 0x9E3779B185EBCA87ull + 0x85EBCA77C2B2AE63ull; }
static inline u64 kanzi_xxhash64(const u8* data, int length, u64 seed) {
  const u64 P1 = 0x9E3779B185EBCA87ull, P2 = 0xC2B2AE3D27D4EB4Full, P3 = 0x165667B19E3779F9ull, P4 = 0x85EBCA77C2B2AE63ull, P5 = 0x27D4EB2F165667C5ull;
  const int end = length;
  u64 h64; int idx = 0;
  if (length >= 32) {
    const int end32 = end - 32;
    u64 v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
    do {
      v1 = xxh64_round(v1, le64(data + idx)); v2 = xxh64_round(v2, le64(data + idx + 8));
      v3 = xxh64_round(v3, le64(data + idx + 16)); v4 = xxh64_round(v4, le64(data + idx + 24));
      idx += 32;
    } while (idx <= end32);
    h64 = ((v1 << 1) | (v1 >> 31)) + ((v2 << 7) | (v2 >> 25)) + ((v3 << 12) | (v3 >> 20)) + ((v4 << 18) | (v4 >> 14));     // as written (:127-128)
    h64 = xxh64_merge(h64, v1); h64 = xxh64_merge(h64, v2); h64 = xxh64_merge(h64, v3); h64 = xxh64_merge(h64, v4);
  } else h64 = seed + P5;
  h64 += (u64)(i64)length;
  while (idx + 8 <= end) { h64 ^= xxh64_round(0, le64(data + idx)); h64 = ((h64 << 27) | (h64 >> 37)) * P1 + P4; idx += 8; }
  while (idx + 4 <= end) { h64 ^= (u64)(i64)(i32)le32(data + idx) * P1; h64 = ((h64 << 23) | (h64 >> 41)) * P2 + P3; idx += 4; }   // readInt32 is a signed int widened to long
  while (idx < end) { h64 ^= (u64)data[idx] * P5; h64 = ((h64 << 11) | (h64 >> 53)) * P1; idx++; }
  h64 ^= h64 >> 33; h64 *= P2; h64 ^= h64 >> 29; h64 *= P3;
  return h64 ^ (h64 >> 32);
}

struct StreamParams {
  u64 transformType = 0;     // 48 bits, 8 x 6
  int entropyType = E_NONE;
  int blockSize = 4 << 20;
  i64 inputSize = 0;         // ctx["fileSize"]; 0 = unknown
  int bwtBounds = 1;         // 1 = as the reference is written (SURVEY E-1), 0 = "fixed"
  int checksum = 0;          // ctx["checksum"]: 0, 32 or 64 (COS:193-204)
};

// writeHeader, CompressedOutputStream.java:236-313
static inline void writeStreamHeader(BitWriter& obs, const StreamParams& sp) {
  obs.writeBits((u32)BITSTREAM_TYPE, 32);
  obs.writeBits(BITSTREAM_FORMAT_VERSION, 4);
  const int chkSize = (sp.checksum == 32) ? 1 : ((sp.checksum == 64) ? 2 : 0);     // COS:248-254
  obs.writeBits(chkSize, 2);
  obs.writeBits((u64)sp.entropyType, 5);
  obs.writeBits(sp.transformType, 48);
  obs.writeBits((u64)((u32)sp.blockSize >> 4), 28);
  int szMask = 0;
  if ((sp.inputSize != 0) && (sp.inputSize < (1LL << 48))) {
    if (sp.inputSize >= (1LL << 32)) szMask = 3;
    else {
      i64 isz = sp.inputSize;
      if (isz > (1LL << 30)) { isz >>= 4; szMask++; }
      szMask += ((log2i((u32)isz) >> 4) + 1);
    }
  }
  obs.writeBits((u64)szMask, 2);
  if (szMask > 0) obs.writeBits((u64)sp.inputSize, 16 * szMask);
  obs.writeBits(0, 15);
  const i32 seed = jmul(0x01030507, BITSTREAM_FORMAT_VERSION);
  const i32 HASH = 0x1E35A7BD;
  i32 cksum = jmul(HASH, seed);
  cksum = mix32(cksum, HASH, chkSize);
  cksum = mix32(cksum, HASH, sp.entropyType);
  cksum = mix32(cksum, HASH, (i32)(sp.transformType >> 32));
  cksum = mix32(cksum, HASH, (i32)sp.transformType);
  cksum = mix32(cksum, HASH, sp.blockSize);
  if (szMask > 0) {
    cksum = mix32(cksum, HASH, (i32)((u64)sp.inputSize >> 32));
    cksum = mix32(cksum, HASH, (i32)sp.inputSize);
  }
  cksum = (i32)(((u32)cksum >> 23) ^ ((u32)cksum >> 3));
  obs.writeBits((u64)(u32)cksum, 24);
}

// Per-"task" persistent slices, as CompressedOutputStream keeps buffers[0] / buffers[jobs] (COS:213-222)
struct EncodeBuffers {
  std::vector<u8> dataArr, bufArr;
  Slice data, buffer;
  explicit EncodeBuffers(int blockSize) {
    const int bufSize = std::max(blockSize + (blockSize >> 3), 256 * 1024);
    dataArr.assign(bufSize, 0);
    data = Slice(&dataArr, bufSize, 0);
    buffer = Slice(&bufArr, 0, 0);
  }
};

struct BlockInfo { int mode = 0, skipFlags = 0, postTransformLength = 0; i64 written = 0; };

// EncodingTask.encodeBlock, CompressedOutputStream.java:733-1054.  The block bytes are already in
// eb.data.arr[0..blockLength).  Appends the block record to `obs`.  Throws JavaException where Java throws.
static inline void encodeBlock(BitWriter& obs, EncodeBuffers& eb, int blockLength, const StreamParams& sp, BlockInfo* info = nullptr) {
  if (blockLength == 0) return;
  Slice& data = eb.data; Slice& buffer = eb.buffer;
  Ctx ctx; ctx.bsVersion = BITSTREAM_FORMAT_VERSION; ctx.blockSize = sp.blockSize; ctx.jobs = 1; ctx.bwtBounds = sp.bwtBounds;
  ctx.entropyType = sp.entropyType;             // the stream's ctx["entropy"] (COS:147), also for copy blocks
  int mode = 0;
  u64 blockTransformType = sp.transformType; int blockEntropyType = sp.entropyType;
  if (blockLength <= SMALL_BLOCK_SIZE) { blockTransformType = 0; blockEntropyType = E_NONE; mode |= COPY_BLOCK_MASK; }
  u64 checksum = 0;                              // of the original bytes, before anything else touches them (COS:745-755)
  if (sp.checksum == 32) checksum = (u64)xxhash32(data.p(), blockLength, (u32)BITSTREAM_TYPE);
  else if (sp.checksum == 64) checksum = kanzi_xxhash64(data.p(), blockLength, (u64)(i64)BITSTREAM_TYPE);
  ctx.size = blockLength;
  Sequence transform(ctx, blockTransformType);
  const int requiredSize = transform.getMaxEncodedLength(blockLength);
  if (blockLength >= 4) {
    const i32 m = magic::getType(data.p(), data.cap(), 0);
    if (magic::isCompressed(m)) ctx.dataType = DT_BIN;
    else if (magic::isMultimedia(m)) ctx.dataType = DT_MULTIMEDIA;
    else if (magic::isExecutable(m)) ctx.dataType = DT_EXE;
  }
  if (buffer.length < requiredSize) {
    buffer.length = requiredSize;
    if (buffer.cap() < buffer.length) buffer.arr->assign(buffer.length, 0);
  }
  buffer.index = 0;
  data.length = blockLength;
  transform.forward(data, buffer);
  const int postTransformLength = buffer.index;
  if (postTransformLength < 0) throw JavaException("Invalid transform size");
  ctx.size = postTransformLength;
  const int dataSize = (postTransformLength < 256) ? 1 : (log2i((u32)postTransformLength) >> 3) + 1;
  if (dataSize > 4) throw JavaException("Invalid block data length");
  const int skipFlags = transform.skipFlags & 0xFF;
  const int nbFunctions = transform.getNbFunctions();
  mode |= (((dataSize - 1) & 0x03) << 5);
  const int bufSize = std::max(256 * 1024, std::max(postTransformLength, blockLength + (blockLength >> 3)));
  if (data.length < bufSize) {
    data.length = bufSize;
    if (data.cap() < data.length) data.arr->assign(data.length, 0);
  }
  data.index = 0;
  BitWriter os;
  int headerSkipFlags = skipFlags;
  if (((mode & COPY_BLOCK_MASK) != 0) || (nbFunctions <= 4)) {
    mode |= (skipFlags >> 4);
    if ((mode & COPY_BLOCK_MASK) != 0) headerSkipFlags = 0;
    else headerSkipFlags = ((mode << 4) | 0x0F) & 0xFF;
    os.writeBits((u64)(mode & 0xFF), 8);
  } else {
    mode |= TRANSFORMS_MASK;
    os.writeBits((u64)(mode & 0xFF), 8);
    os.writeBits((u64)skipFlags, 8);
  }
  os.writeBits((u64)(u32)postTransformLength, 8 * dataSize);
  int headerChecksumIndex = 1 + dataSize;
  if (((mode & COPY_BLOCK_MASK) == 0) && (nbFunctions > 4)) headerChecksumIndex++;
  os.writeBits(0, 8);
  if (sp.checksum == 32) os.writeBits(checksum, 32); else if (sp.checksum == 64) os.writeBits(checksum, 64);       // COS:892-895
  if (entropyEncode(blockEntropyType, os, buffer.p(), postTransformLength) != postTransformLength)
    throw JavaException("Entropy coding failed");
  i64 written = (i64)os.written();
  os.close();
  std::vector<u8>* payload = &os.buf;
  BitWriter copyOs;
  if ((mode & COPY_BLOCK_MASK) == 0) {
    const i64 rawPayloadBytes = postTransformLength;
    const i64 entropyPayloadBytes = (written + 7) >> 3;
    if (rawPayloadBytes < entropyPayloadBytes) {       // "transformed copy" block, COS:926-973
      const int copyMode = mode | COPY_BLOCK_MASK | TRANSFORMS_MASK;
      copyOs.writeBits((u64)(copyMode & 0xFF), 8);
      if (nbFunctions > 4) copyOs.writeBits((u64)skipFlags, 8);
      copyOs.writeBits((u64)(u32)postTransformLength, 8 * dataSize);
      headerChecksumIndex = 1 + dataSize;
      if (nbFunctions > 4) { headerChecksumIndex++; headerSkipFlags = skipFlags; }
      else headerSkipFlags = ((copyMode << 4) | 0x0F) & 0xFF;
      copyOs.writeBits(0, 8);
      if (sp.checksum == 32) copyOs.writeBits(checksum, 32); else if (sp.checksum == 64) copyOs.writeBits(checksum, 64);
      copyOs.writeBytesBits(buffer.p(), (i64)postTransformLength << 3);
      written = (i64)copyOs.written();
      copyOs.close();
      payload = &copyOs.buf;
      mode = copyMode;
    }
  }
  const i32 HASH = 0x1E35A7BD;
  i32 cksum = jmul(HASH, 0x01030507);
  cksum = mix32(cksum, HASH, mode & 0xFF);
  cksum = mix32(cksum, HASH, headerSkipFlags & 0xFF);
  cksum = mix32(cksum, HASH, postTransformLength);
  cksum = mix32(cksum, HASH, (i32)((u64)written >> 32));
  cksum = mix32(cksum, HASH, (i32)written);
  cksum = (i32)(((u32)cksum >> 23) ^ ((u32)cksum >> 3));
  (*payload)[headerChecksumIndex] = (u8)cksum;
  const int lw = (written < 8) ? 3 : log2i((u32)(written >> 3)) + 4;
  obs.writeBits((u64)(lw - 3), 5);
  obs.writeBits((u64)written, lw);
  obs.writeBytesBits(payload->data(), written);
  // CustomByteArrayOutputStream keeps/grows data.array (COS:918-920): capacity can only grow
  if ((int)payload->size() > data.cap()) { data.arr->assign(payload->size(), 0); }
  data.length = data.cap();
  if (info) { info->mode = mode & 0xFF; info->skipFlags = skipFlags; info->postTransformLength = postTransformLength; info->written = written; }
}

// whole-stream compression: CompressedOutputStream.write/processBlock/close (COS:359-504) with jobs = 1
static inline std::vector<u8> compressStream(const u8* in, i64 n, const StreamParams& sp, std::vector<BlockInfo>* infos = nullptr) {
  if (sp.blockSize > MAX_BITSTREAM_BLOCK_SIZE || sp.blockSize < MIN_BITSTREAM_BLOCK_SIZE || (sp.blockSize & -16) != sp.blockSize)
    throw JavaException("invalid block size");
  BitWriter obs;
  writeStreamHeader(obs, sp);
  EncodeBuffers eb(sp.blockSize);
  for (i64 off = 0; off < n; off += sp.blockSize) {
    const int len = (int)std::min<i64>(sp.blockSize, n - off);
    memcpy(eb.data.p(), in + off, len);
    eb.data.index = 0;
    BlockInfo bi;
    encodeBlock(obs, eb, len, sp, &bi);
    if (infos) infos->push_back(bi);
  }
  obs.writeBits(0, 5); obs.writeBits(0, 3);
  obs.close();
  return obs.buf;
}

struct StreamHeader { int bsVersion = 0, chkSize = 0, entropyType = 0, blockSize = 0, szMask = 0; u64 transformType = 0; i64 outputSize = 0; };

// readHeader, CompressedInputStream.java:359-478 (bsVersion 7 only)
static inline StreamHeader readStreamHeader(BitReader& ibs) {
  StreamHeader h;
  if ((i32)ibs.readBits(32) != BITSTREAM_TYPE) throw JavaException("Invalid stream type");
  h.bsVersion = (int)ibs.readBits(4);
  if (h.bsVersion != BITSTREAM_FORMAT_VERSION) throw JavaException("oracle reads bitstream version 7 only");
  h.chkSize = (int)ibs.readBits(2);
  if (h.chkSize == 3) throw JavaException("Invalid bitstream, incorrect block checksum size");
  h.entropyType = (int)ibs.readBits(5);
  h.transformType = ibs.readBits(48);
  h.blockSize = (int)ibs.readBits(28) << 4;
  if ((h.blockSize < MIN_BITSTREAM_BLOCK_SIZE) || (h.blockSize > MAX_BITSTREAM_BLOCK_SIZE)) throw JavaException("incorrect block size");
  h.szMask = (int)ibs.readBits(2);
  if (h.szMask != 0) h.outputSize = (i64)ibs.readBits(16 * h.szMask);
  ibs.readBits(15);
  const i32 cksum1 = (i32)ibs.readBits(24);
  const i32 HASH = 0x1E35A7BD;
  i32 c = jmul(HASH, jmul(0x01030507, h.bsVersion));
  c = mix32(c, HASH, h.chkSize);
  c = mix32(c, HASH, h.entropyType);
  c = mix32(c, HASH, (i32)(h.transformType >> 32));
  c = mix32(c, HASH, (i32)h.transformType);
  c = mix32(c, HASH, h.blockSize);
  if (h.szMask > 0) { c = mix32(c, HASH, (i32)((u64)h.outputSize >> 32)); c = mix32(c, HASH, (i32)h.outputSize); }
  c = (i32)(((u32)c >> 23) ^ ((u32)c >> 3));
  if (cksum1 != (c & ((1 << 24) - 1))) throw JavaException("Invalid bitstream, checksum mismatch");
  return h;
}

// DecodingTask.decodeBlock, CompressedInputStream.java:1106-1378 (+ readBlockHeader :1025-1095).
// Returns decoded byte count appended to `out`; 0 at the end-of-stream marker.
static inline int decodeBlock(BitReader& ibs, const StreamHeader& h, std::vector<u8>& out, int bwtBounds) {
  const int lr = (int)ibs.readBits(5) + 3;
  i64 read = (i64)ibs.readBits(lr);
  if (read == 0) return 0;
  const i64 encodedBlockBytes = (read + 7) >> 3;
  const i64 encodedBlockLength = read;
  const int maxTransformLength = std::min(std::max(h.blockSize + h.blockSize / 2, 2048), (int)MAX_BITSTREAM_BLOCK_SIZE);
  // readBlockHeader
  if (encodedBlockLength < 8) throw JavaException("Invalid block size");
  Ctx ctx; ctx.bsVersion = h.bsVersion; ctx.blockSize = h.blockSize; ctx.jobs = 1; ctx.bwtBounds = bwtBounds;
  const int mode = (int)(int8_t)ibs.readBits(8);
  int skipFlags = 0; bool hasSkipFlags = false, transformedCopy = false;
  const bool copyBlock = (mode & COPY_BLOCK_MASK) != 0;
  if (copyBlock) {
    if ((mode & TRANSFORMS_MASK) != 0) {
      transformedCopy = true;
      Ctx tmp = ctx;
      const int nbFunctions = Sequence(tmp, h.transformType).getNbFunctions();
      if (nbFunctions > 4) hasSkipFlags = true; else skipFlags = ((mode << 4) | 0x0F) & 0xFF;
    }
  } else if ((mode & TRANSFORMS_MASK) != 0) hasSkipFlags = true;
  else skipFlags = ((mode << 4) | 0x0F) & 0xFF;
  const int dataSize = 1 + ((mode >> 5) & 0x03);
  const int headerSize = 1 + (hasSkipFlags ? 1 : 0) + dataSize + 1;
  if (encodedBlockLength < (headerSize << 3)) throw JavaException("Invalid block size");
  if (hasSkipFlags) skipFlags = (int)ibs.readBits(8);
  int preTransformLength = 0;
  for (int i = 0; i < dataSize; i++) preTransformLength = (int)(((u32)preTransformLength << 8) | (u32)ibs.readBits(8));
  const int headerChecksum = (int)ibs.readBits(8) & 0xFF;
  const i32 HASH = 0x1E35A7BD;
  i32 c = jmul(HASH, 0x01030507);
  c = mix32(c, HASH, mode & 0xFF);
  c = mix32(c, HASH, skipFlags & 0xFF);
  c = mix32(c, HASH, preTransformLength);
  c = mix32(c, HASH, (i32)((u64)encodedBlockLength >> 32));
  c = mix32(c, HASH, (i32)encodedBlockLength);
  c = (i32)(((u32)c >> 23) ^ ((u32)c >> 3));
  if (headerChecksum != (c & 0xFF)) throw JavaException("Invalid bitstream, block header checksum mismatch");
  const bool rawCopy = copyBlock && !transformedCopy;
  if ((preTransformLength < 0) || (preTransformLength > maxTransformLength)) throw JavaException("Invalid compressed block length");
  const int checksumSize = (h.chkSize == 2) ? 8 : ((h.chkSize == 1) ? 4 : 0);
  if (encodedBlockBytes > (i64)preTransformLength + headerSize + checksumSize) throw JavaException("Invalid block size");      // CIS:1158-1165
  read -= (i64)headerSize << 3;
  // payload bits -> private byte array (COS:1182-1187), then a per-block bit reader over it
  const int r = (int)encodedBlockBytes;
  std::vector<u8> payload((size_t)std::max(h.blockSize, r) + 8, 0);
  ibs.readBytesBits(payload.data(), read);
  BitReader is(payload.data(), (u64)(r - headerSize) * 8);
  u64 blockTransformType = h.transformType; int blockEntropyType = h.entropyType;
  if (rawCopy) { blockTransformType = 0; blockEntropyType = E_NONE; }
  else if (transformedCopy) blockEntropyType = E_NONE;
  if (preTransformLength == 0) return 0;
  u64 checksum1 = 0;                              // CIS:1247-1253
  if (h.chkSize == 1) checksum1 = is.readBits(32); else if (h.chkSize == 2) checksum1 = is.readBits(64);
  // buffers: `buffer` (entropy output) and `data` (final output), CIS:719-727, 1283-1288
  const int blkBuf = std::max(h.blockSize + EXTRA_BUFFER_SIZE, h.blockSize + (h.blockSize >> 4));
  std::vector<u8> dataArr((size_t)std::max(blkBuf, std::max(h.blockSize, r)), 0), bufArr;
  Slice data(&dataArr, blkBuf, 0), buffer(&bufArr, 0, 0);
  const int bufferSize = std::max(h.blockSize, preTransformLength + EXTRA_BUFFER_SIZE);
  buffer.length = bufferSize; bufArr.assign(bufferSize, 0);
  ctx.size = preTransformLength;
  if (transformedCopy) {
    is.readBytesBits(buffer.p(), (i64)preTransformLength << 3);
  } else {
    if (entropyDecode(blockEntropyType, is, buffer.p(), preTransformLength) != preTransformLength)
      throw JavaException("Entropy decoding failed");
  }
  Sequence transform(ctx, blockTransformType);
  transform.skipFlags = (u8)skipFlags;
  buffer.index = 0;
  buffer.length = preTransformLength;
  if (!transform.inverse(buffer, data)) throw JavaException("Transform inverse failed");
  const int decoded = data.index;
  if (decoded > h.blockSize) throw JavaException("Block incorrectly decompressed");
  if (h.chkSize == 1 && xxhash32(dataArr.data(), decoded, (u32)BITSTREAM_TYPE) != (u32)checksum1) throw JavaException("Corrupted bitstream: invalid checksum");     // CIS:1348-1370
  if (h.chkSize == 2 && kanzi_xxhash64(dataArr.data(), decoded, (u64)(i64)BITSTREAM_TYPE) != checksum1) throw JavaException("Corrupted bitstream: invalid checksum");
  out.insert(out.end(), dataArr.begin(), dataArr.begin() + decoded);
  return decoded;
}

static inline std::vector<u8> decompressStream(const u8* in, i64 nBytes, int bwtBounds = 1, StreamHeader* hdr = nullptr) {
  BitReader ibs(in, (u64)nBytes * 8);
  StreamHeader h = readStreamHeader(ibs);
  if (hdr) *hdr = h;
  std::vector<u8> out;
  while (true) { if (decodeBlock(ibs, h, out, bwtBounds) == 0) break; }
  return out;
}

}  // namespace kzo
