// kzg_common.cuh — shared device/host helpers for libkanzi_b200 (sm_100a).
// Product code: nothing here includes or links anything under oracle/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/kzg.h"

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;
typedef int64_t i64;

#define KZG_SM_COUNT 148   // B200: 2 dies x 74 SMs; persistent grids are sized in multiples of this

// ---- device-side block descriptor ------------------------------------------------------------------
// One per block of a batch (a "block" is Kanzi's unit of independent work, COS:512-586).
// cur/curLen track the live slice as Sequence.forward / inverse ping-pong two buffers
// (K/transform/Sequence.java:56-127,137-207): a stage that succeeds writes `alt`, then swaps.
struct KzgBlock {
  u8* cur;        // current data (input of the next stage)
  u8* alt;        // the other ping-pong buffer (output of the next stage)
  i32 curLen;     // bytes valid at cur
  i32 cap;        // capacity of both buffers
  i32 origLen;    // block length before transforms (encode) / decoded length (decode)
  i32 skipFlags;  // Sequence skip flags (bit 7-i set = transform i skipped)
  i32 dataType;   // ctx["dataType"] (KZG_DT_*), in/out
  i32 status;     // 0 ok, else -KZG_ERR_* raised by a kernel
  // entropy stage / container
  i64 entBits;    // bit length of the entropy payload
  i32 mode;       // block header mode byte
  i32 hdrBytes;   // block header length in bytes (mode [+skip] + len + cksum)
  i64 written;    // bit length of the block record payload (header + entropy payload)
  // decode side
  i64 srcBit;     // absolute bit offset of the entropy payload in the stream
  i64 srcBits;    // bits available from srcBit
  i32 preLen;     // preTransformLength (entropy decoder output size)
  i32 entropy;    // entropy codec id of this block (E_NONE for copy blocks)
  // buffer rotation: encode: aux0 = the read-only input block, aux1 = second scratch buffer;
  // decode: aux0 = final destination of the block, stagesLeft = inverse stages still to run
  u8* aux0; u8* aux1;
  i32 stagesLeft;
  i32 finalCap;   // decode: bytes the caller's output buffer holds at aux0 (0 = no clamp); see kzg_dst_limit
  u64 xxh;        // block checksum of the original bytes (XXHash32 / Kanzi's XXHash64), encode: computed, decode: expected
  i32 chkBytes;   // 0, 4 or 8: checksum bytes in the block record (COS:892-895)
  i32 pad1;
};

// dst limit of a stage as its kernels must see it: the Java-visible limit, clamped to the caller's buffer when the stage
// writes straight into the final destination (decode: alt == aux0), so no kernel ever writes past the caller's outCap
__host__ __device__ __forceinline__ int kzg_dst_limit(const KzgBlock& B, int limit) {
  return (B.finalCap > 0 && B.alt == B.aux0 && limit > B.finalCap) ? B.finalCap : limit;
}

// ---- MSB-first bit I/O on byte buffers (format of K/bitstream/Default{In,Out}putBitStream.java) -------
__device__ __forceinline__ u32 ld_be32_unaligned(const u8* p) {
  return ((u32)p[0] << 24) | ((u32)p[1] << 16) | ((u32)p[2] << 8) | (u32)p[3];
}

// read `n` (1..32) bits at absolute bit position `pos` of `base`, MSB first
__device__ __forceinline__ u32 get_bits(const u8* __restrict__ base, u64 pos, int n) {
  const u8* p = base + (pos >> 3);
  const int sh = (int)(pos & 7);
  u64 w = ((u64)p[0] << 32) | ((u64)p[1] << 24) | ((u64)p[2] << 16) | ((u64)p[3] << 8) | (u64)p[4];
  return (u32)((w >> (40 - sh - n)) & ((n == 32) ? 0xFFFFFFFFull : ((1ull << n) - 1)));
}

// sequential reader (single thread)
struct BitReaderD {
  const u8* base; u64 pos; u64 end;
  __device__ __forceinline__ BitReaderD(const u8* b, u64 p, u64 e) : base(b), pos(p), end(e) {}
  __device__ __forceinline__ u32 read(int n) {   // n in 1..32; reads past `end` yield zeros (flagged by caller via overrun())
    u32 v = 0;
    if (pos + (u64)n <= end) v = get_bits(base, pos, n);
    pos += (u64)n;
    return v;
  }
  __device__ __forceinline__ bool overrun() const { return pos > end; }
};

// sequential writer (single thread) into a zero-initialised or private byte buffer
struct BitWriterD {
  u8* base; u64 acc; int nacc; i64 nbytes;   // acc holds nacc (<64) pending bits, right-aligned
  __device__ __forceinline__ BitWriterD(u8* b) : base(b), acc(0), nacc(0), nbytes(0) {}
  __device__ __forceinline__ void write(u32 v, int n) {   // n in 0..32
    if (n == 0) return;
    const u64 m = (n == 32) ? 0xFFFFFFFFull : ((1ull << n) - 1);
    acc = (acc << n) | ((u64)v & m);
    nacc += n;
    while (nacc >= 8) { nacc -= 8; base[nbytes++] = (u8)(acc >> nacc); }
  }
  __device__ __forceinline__ i64 bits() const { return nbytes * 8 + nacc; }
  __device__ __forceinline__ void flush() {   // pad the last byte with zeros
    if (nacc > 0) { base[nbytes] = (u8)(acc << (8 - nacc)); }
  }
};

// EntropyUtils.writeVarInt (K/entropy/EntropyUtils.java:259-276)
__device__ __forceinline__ void write_varint(BitWriterD& bw, i32 value) {
  u32 v = (u32)value;
  if (value >= 128 || value < 0) {
    bw.write(0x80 | (v & 0x7F), 8); v >>= 7;
    while (v >= 128) { bw.write(0x80 | (v & 0x7F), 8); v >>= 7; }
  }
  bw.write(v, 8);
}
// EntropyUtils.readVarInt (K/entropy/EntropyUtils.java:284-300)
__device__ __forceinline__ i32 read_varint(BitReaderD& br) {
  u32 value = br.read(8);
  u32 res = value & 0x7F;
  int shift = 7;
  while (value >= 128) {
    value = br.read(8);
    res |= ((value & 0x7F) << shift);
    if (shift == 28) break;
    shift += 7;
  }
  return (i32)res;
}

__device__ __forceinline__ int ilog2(u32 x) { return 31 - __clz(x); }   // Global.log2, x > 0

// ---- bulk asynchronous copy global -> shared (the TMA engine's 1-D form, sm_90+: `cp.async.bulk`), completion on an mbarrier ----
// One thread issues the copy of a whole tile; the CTA's other instructions (and the issuing thread's) run while the bytes arrive.
// dst and src must be 16-byte aligned and `bytes` a multiple of 16.
__device__ __forceinline__ void kzg_mbar_init(u64* bar, int arrivals) {
  const u32 a = (u32)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(arrivals));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // make the initialised barrier visible to the async proxy
}
__device__ __forceinline__ void kzg_bulk_g2s(void* dstSmem, const void* srcGlobal, u32 bytes, u64* bar) {
  const u32 d = (u32)__cvta_generic_to_shared(dstSmem), b = (u32)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(d), "l"(srcGlobal), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void kzg_mbar_wait(u64* bar, u32 phase) {
  const u32 b = (u32)__cvta_generic_to_shared(bar);
  asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
               ::"r"(b), "r"(phase) : "memory");
}

// mix32 of the container checksums (COS:89-93)
__host__ __device__ __forceinline__ u32 kzg_mix32(u32 c, u32 h, u32 v) {
  c ^= h * ~v;
  c = (c << 13) | (c >> 19);
  return c * 5u + 0x52DCE729u;
}

// ---- bit-granular segment copy (container assembly, chunk concatenation) -------------------------------
// copies nBits bits from (src, srcBit) to (dst, dstBit); dst must be zero-filled beforehand; boundary
// words are merged with atomicOr so neighbouring segments may share a 32-bit word.
struct KzgSeg { const u8* src; u64 srcBit; u64 dstBit; u64 nBits; };

// per-kernel event timing (bench.py's roofline block): a no-op unless kzg_set_profiling(1) was called on this thread
void kzg_prof_begin(const char* name, cudaStream_t s);
void kzg_prof_end(cudaStream_t s);
#define KZG_PROF(name, s, launch) do { kzg_prof_begin(name, s); launch; kzg_prof_end(s); } while (0)

#define CUDA_TRY(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { kzg_set_error("%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e__)); return -KZG_ERR_PROCESS_BLOCK; } } while (0)
void kzg_set_error(const char* fmt, ...);
