// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the flanglet/kanzi (Java, 2.5.0,
// bitstream v7) hot path.  Nothing under kanzi_b200/ may include, link or call this.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
//
// PARITY STATUS: "parity unpinned" for codec bitstreams — the reference's own tests hold no
// golden output bytes and no JVM exists in this image.  What IS pinned:
// BWT 'mississippi' KAT (BWT.java:45-50), the block/stream header checksum formula
// (TestCompressedStream.java:488-504) and the hand-derived container KATs of SURVEY Appendix D.
//
// Shared plumbing: Java integer semantics helpers, SliceByteArray, MSB-first bitstreams.
#pragma once
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <string>
#include <stdexcept>
#include <algorithm>

namespace kzo {

typedef uint8_t u8;
typedef int32_t i32;
typedef uint32_t u32;
typedef int64_t i64;
typedef uint64_t u64;

// ---- Java semantics -------------------------------------------------------------------------
static inline i32 jmul(i32 a, i32 b) { return (i32)((u32)a * (u32)b); }          // wrapping int *
static inline i32 jadd(i32 a, i32 b) { return (i32)((u32)a + (u32)b); }
static inline i32 jushr(i32 a, int s) { return (i32)((u32)a >> (s & 31)); }      // >>>
static inline i32 jshl(i32 a, int s) { return (i32)((u32)a << (s & 31)); }
static inline i32 rotl32(i32 x, int r) { u32 u = (u32)x; return (i32)((u << r) | (u >> (32 - r))); }

// Global.log2 (Global.java:207-212): floor(log2(x)), x > 0
static inline int log2i(u32 x) { return 31 - __builtin_clz(x); }

static inline u64 le64(const u8* p) { u64 v; memcpy(&v, p, 8); return v; }
static inline u32 le32(const u8* p) { u32 v; memcpy(&v, p, 4); return v; }
static inline u32 le16(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8); }
static inline void put_le32(u8* p, u32 v) { memcpy(p, &v, 4); }
static inline u32 be32(const u8* p) { return ((u32)p[0] << 24) | ((u32)p[1] << 16) | ((u32)p[2] << 8) | p[3]; }
static inline void put_be32(u8* p, u32 v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }

// ---- SliceByteArray (SliceByteArray.java:34-37): array + length + index ------------------------
// `arr` plays the role of the Java byte[] reference (identity = pointer equality, array.length =
// arr->size()).  Java code that does `slice.array = new byte[n]` becomes arr->assign(n, 0).
struct Slice {
  std::vector<u8>* arr; i32 length; i32 index;
  Slice() : arr(nullptr), length(0), index(0) {}
  Slice(std::vector<u8>* a, i32 l, i32 i) : arr(a), length(l), index(i) {}
  i32 cap() const { return (i32)arr->size(); }
  u8* p() const { return arr->data(); }
};

// Java exceptions that escape a codec (ArrayIndexOutOfBounds etc.) -> block error in the host.
struct JavaException : std::runtime_error { using std::runtime_error::runtime_error; };

struct BitStreamError : std::runtime_error { using std::runtime_error::runtime_error; };

// ---- DefaultOutputBitStream (bitstream/DefaultOutputBitStream.java:86-217): MSB-first appender --
class BitWriter {
 public:
  std::vector<u8> buf;
  u64 cur = 0; int avail = 64;  // bits free in cur
  void writeBit(int b) { writeBits((u64)(b & 1), 1); }
  // keeps only the `count` low bits of value (DefaultOutputBitStream.java:110)
  void writeBits(u64 value, int count) {
    if (count == 0) return;
    if (count < 64) value &= ((1ULL << count) - 1);
    if (count < avail) { avail -= count; cur |= value << avail; return; }
    int rem = count - avail;           // bits that do not fit
    cur |= (rem == 64) ? 0 : (value >> rem);
    push();
    if (rem) { avail = 64 - rem; cur = value << avail; }
  }
  // writeBits(byte[],start,count): count bits, MSB-first from bits[start] (…:139-206)
  void writeBytesBits(const u8* p, i64 count) {
    while (count >= 8) { writeBits(*p++, 8); count -= 8; }
    if (count > 0) writeBits((u64)(*p >> (8 - count)), (int)count);
  }
  u64 written() const { return (u64)buf.size() * 8 + (64 - avail); }
  // close(): pads the last byte with zeros (…:253-293)
  void close() {
    int used = 64 - avail;
    for (int sh = 56; used > 0; sh -= 8, used -= 8) buf.push_back((u8)(cur >> sh));
    cur = 0; avail = 64; closed_bits = true;
  }
 private:
  bool closed_bits = false;
  void push() { for (int sh = 56; sh >= 0; sh -= 8) buf.push_back((u8)(cur >> sh)); cur = 0; avail = 64; }
};

// ---- DefaultInputBitStream (bitstream/DefaultInputBitStream.java:77-192) -----------------------
class BitReader {
 public:
  const u8* p; u64 nbits; u64 pos = 0;
  BitReader(const u8* data, u64 bits) : p(data), nbits(bits) {}
  int readBit() { return (int)readBits(1); }
  u64 readBits(int count) {
    if (count == 0) return 0;
    if (pos + (u64)count > ((nbits + 7) & ~7ULL)) throw BitStreamError("end of stream");
    if (count > 56) { u64 hi = readBits(count - 32); return (hi << 32) | readBits(32); }
    const u64 byte = pos >> 3; const int off = (int)(pos & 7);
    if ((byte + 8) * 8 <= ((nbits + 7) & ~7ULL)) {   // fast path: 8 readable bytes
      u64 v = __builtin_bswap64(le64(p + byte));
      pos += count;
      return (v << off) >> (64 - count);
    }
    u64 r = 0; int c = count;
    while (c > 0) {
      int o = (int)(pos & 7); int take = std::min(8 - o, c);
      u32 b = p[pos >> 3];
      r = (r << take) | ((b >> (8 - o - take)) & ((1u << take) - 1));
      pos += take; c -= take;
    }
    return r;
  }
  void readBytesBits(u8* out, i64 count) {
    if ((pos & 7) == 0 && count >= 8) {
      i64 nb = count >> 3;
      if (pos + (u64)nb * 8 > ((nbits + 7) & ~7ULL)) throw BitStreamError("end of stream");
      memcpy(out, p + (pos >> 3), (size_t)nb); pos += (u64)nb * 8; out += nb; count -= nb * 8;
    }
    while (count >= 8) { *out++ = (u8)readBits(8); count -= 8; }
    if (count > 0) *out = (u8)(readBits((int)count) << (8 - count));
  }
  u64 read() const { return pos; }
};

// ---- ctx (the Map<String,Object> fields the hot-path codecs consume) ---------------
enum DataType { DT_UNDEFINED = 0, DT_TEXT, DT_MULTIMEDIA, DT_EXE, DT_NUMERIC, DT_BASE64, DT_DNA, DT_BIN,
                DT_UTF8, DT_SMALL_ALPHABET };
struct Ctx {
  i32 bsVersion = 7;
  i32 blockSize = 0;
  i32 size = 0;
  i32 jobs = 1;
  i32 dataType = DT_UNDEFINED;   // in/out
  i32 lzType = 3;                // TransformFactory.LZ_TYPE
  i32 sbrtMode = 0;
  i32 rolzExtra = 0;             // transform string is ROLZX
  i32 bwtBounds = 0;             // 0 = "fixed" (ignore the BWT.java:152-156 clause), 1 = "asref" 
  i32 entropyType = 0;           // ctx["entropy"] as an id (absent = "NONE"); RLT.forward reads it (RLT.java:101-107)
};

// transform ids (TransformFactory.java:36-112) and entropy ids (EntropyCodecFactory.java:38-74)
enum { T_NONE = 0, T_BWT = 1, T_BWTS = 2, T_LZ = 3, T_SNAPPY = 4, T_RLT = 5, T_ZRLT = 6, T_MTFT = 7, T_RANK = 8,
       T_EXE = 9, T_DICT = 10, T_ROLZ = 11, T_ROLZX = 12, T_SRT = 13, T_LZP = 14, T_MM = 15, T_LZX = 16,
       T_UTF = 17, T_PACK = 18, T_DNA = 19 };
enum { E_NONE = 0, E_HUFFMAN = 1, E_FPAQ = 2, E_PAQ = 3, E_RANGE = 4, E_ANS0 = 5, E_CM = 6, E_TPAQ = 7,
       E_ANS1 = 8, E_TPAQX = 9 };

}  // namespace kzo
