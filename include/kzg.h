/*
 * kzg.h — C ABI of libkanzi_b200.so: the B200-native (sm_100a CUDA) replacement for the hot path
 * behind Kanzi's ByteTransform / EntropyEncoder / EntropyDecoder plugin interfaces.
 *
 * Reference = flanglet/kanzi (Java, 2.5.0, bitstream v7).  K/ = java/src/main/java/io/github/flanglet/kanzi/.
 * Each entry point names the reference interface it replaces.  Plain pointers and sizes only; no
 * C++/torch types; never throws; re-entrant (per-thread CUDA stream + workspace).
 *
 * There is NO CPU fallback: every compute entry point returns KZG_ERR_NO_DEVICE when no CUDA
 * device is usable.
 */
#ifndef KZG_H
#define KZG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- ids ------------------------------------------------------------------------------------ */
/* transform ids: K/transform/TransformFactory.java:36-58 */
enum {
  KZG_T_NONE = 0, KZG_T_BWT = 1, KZG_T_LZ = 3, KZG_T_RLT = 5, KZG_T_ZRLT = 6, KZG_T_MTFT = 7, KZG_T_RANK = 8,
  KZG_T_ROLZ = 11, KZG_T_ROLZX = 12, KZG_T_SRT = 13, KZG_T_LZP = 14, KZG_T_LZX = 16
};
/* entropy ids: K/entropy/EntropyCodecFactory.java:38-47 */
enum { KZG_E_NONE = 0, KZG_E_HUFFMAN = 1, KZG_E_FPAQ = 2, KZG_E_ANS0 = 5, KZG_E_ANS1 = 8 };

/* error codes mirror K/Error.java:29-124 (returned negated) */
enum {
  KZG_ERR_MISSING_PARAM = 1, KZG_ERR_BLOCK_SIZE = 2, KZG_ERR_INVALID_CODEC = 3, KZG_ERR_CREATE_COMPRESSOR = 4,
  KZG_ERR_CREATE_DECOMPRESSOR = 5, KZG_ERR_OUTPUT_IS_DIR = 6, KZG_ERR_OVERWRITE_FILE = 7, KZG_ERR_CREATE_FILE = 8,
  KZG_ERR_CREATE_BITSTREAM = 9, KZG_ERR_OPEN_FILE = 10, KZG_ERR_READ_FILE = 11, KZG_ERR_WRITE_FILE = 12,
  KZG_ERR_PROCESS_BLOCK = 13, KZG_ERR_CREATE_CODEC = 14, KZG_ERR_INVALID_FILE = 15, KZG_ERR_STREAM_VERSION = 16,
  KZG_ERR_CREATE_STREAM = 17, KZG_ERR_INVALID_PARAM = 18, KZG_ERR_CRC_CHECK = 19, KZG_ERR_UNKNOWN = 127,
  KZG_ERR_NO_DEVICE = 126   /* not in Error.java: CUDA device / driver missing (no CPU fallback exists) */
};

/* DataType ordinal: K/Global.java:40-90 */
enum { KZG_DT_UNDEFINED = 0, KZG_DT_TEXT, KZG_DT_MULTIMEDIA, KZG_DT_EXE, KZG_DT_NUMERIC, KZG_DT_BASE64, KZG_DT_DNA,
       KZG_DT_BIN, KZG_DT_UTF8, KZG_DT_SMALL_ALPHABET };

/* The Map<String,Object> ctx fields the hot-path codecs read or write:
 * bsVersion (COS:212), blockSize, size (COS:791,833), jobs, dataType (in/out: ROLZCodec.java:451-461),
 * flags bit 0: 1 = keep the bounds clauses of K/transform/BWT.java:152-156,199-203,211 exactly as
 * written (BWTBlockCodec then never fires, DESIGN.md "E-1"), 0 = "fixed". */
typedef struct kzg_ctx {
  int32_t bsVersion;
  int32_t blockSize;
  int32_t size;
  int32_t jobs;
  int32_t dataType;
  int32_t flags;
} kzg_ctx;
#define KZG_FLAG_BWT_ASREF 1
/* ctx.flags bits 8-11, per-block transform calls only: ctx["entropy"] as KZG_E_* id + 1 (0 = key absent, read as "NONE").  RLT.forward
 * chooses its escape byte by it (K/transform/RLT.java:101-107); kzg_compress* passes the stream's entropy codec itself. */
#define KZG_CTX_ENTROPY(e) ((((e) + 1) & 0xF) << 8)
/* kzg_compress* only: per-block checksum of the original bytes, ctx["checksum"] = 32 / 64 (CompressedOutputStream.java:193-204,
 * 745-755): K/util/hash/XXHash32.java (the published XXH32) / XXHash64.java (Kanzi's variant), seed 0x4B414E5A.  kzg_decompress*
 * reads the kind from the stream header and verifies every block (-KZG_ERR_CRC_CHECK on a mismatch). */
#define KZG_FLAG_XXH32 2
#define KZG_FLAG_XXH64 4

/* ---- library ---------------------------------------------------------------------------------- */
int kzg_abi_version(void);
/* number of usable CUDA devices (0 => every compute call fails with -KZG_ERR_NO_DEVICE) */
int kzg_device_count(void);
/* bind the calling thread to a device (default 0).  Returns 0 or a negated error. */
int kzg_set_device(int device);
/* last error text of the calling thread (static storage) */
const char* kzg_last_error(void);
/* kernels this thread launched since the last call with reset != 0 (bench.py's gpu_launches) */
int64_t kzg_launch_count(int reset);

/* ---- ByteTransform (K/ByteTransform.java:24-57) --------------------------------------------------
 * One call = one `forward(SliceByteArray src, SliceByteArray dst)` on host arrays: src slice =
 * (src, length srcLen, index 0), dst slice = (dst, length dstLen, index 0, array.length dstCap).
 * Returns 1 = true, 0 = false (transform skipped / recoverable), < 0 = -KZG_ERR_*.
 * On return *srcUsed / *dstUsed are the slices' new indexes (bytes consumed / produced).
 * getMaxEncodedLength: BWTBlockCodec.java:222-224 (n+33), LZCodec.java:961-964 (LZ, LZX), :1283-1285 (LZP), ROLZCodec.java:1001-1003,
 * :1417-1421 (ROLZX), SBRT/ZRLT n, SRT.java:364-366 (n+1024), RLT.java:355-357. */
int kzg_transform_forward(int type, kzg_ctx* ctx, const uint8_t* src, int32_t srcLen, uint8_t* dst, int32_t dstLen,
                          int32_t dstCap, int32_t* srcUsed, int32_t* dstUsed);
int kzg_transform_inverse(int type, kzg_ctx* ctx, const uint8_t* src, int32_t srcLen, uint8_t* dst, int32_t dstLen,
                          int32_t dstCap, int32_t* srcUsed, int32_t* dstUsed);
int32_t kzg_transform_max_encoded_len(int type, int32_t n);

/* raw BWT as K/transform/BWT.java forward/inverse with index 0 (what T/test/TestBWT.java drives):
 * primaryIndexes[8] out (forward) / in (inverse).  Returns 1 / 0 / < 0. */
int kzg_bwt_forward(const uint8_t* src, int32_t n, uint8_t* dst, int32_t* primaryIndexes8);
int kzg_bwt_inverse(const uint8_t* src, int32_t n, uint8_t* dst, const int32_t* primaryIndexes8);

/* ---- EntropyEncoder / EntropyDecoder (K/EntropyEncoder.java:23-49, K/EntropyDecoder.java:23-47) ----
 * encode: the bits `new XEncoder(bitstream, ctx).encode(src,0,n); dispose()` appends to its bitstream
 * (COS:907-916), returned as an MSB-first bit string starting at bit 0 of `out`; the JNI shim appends
 * it with bitstream.writeBits(out, 0, *outBits).  Returns n (Java's return value) or < 0.
 * decode: reads from bit 0 of `in` (inBits available); writes n bytes; *bitsUsed = bits consumed
 * (the shim advances the Java bitstream by that much).  Returns the Java return value (n on success). */
int64_t kzg_entropy_encode(int type, kzg_ctx* ctx, const uint8_t* src, int32_t n, uint8_t* out, int64_t outCap, int64_t* outBits);
int32_t kzg_entropy_decode(int type, kzg_ctx* ctx, const uint8_t* in, int64_t inBits, int64_t* bitsUsed, uint8_t* dst, int32_t n);

/* Coalescing of concurrent per-block calls.  Off by default: every call above runs alone on the
 * calling thread's stream.  kzg_set_coalescing(maxBatch > 1, windowMicros) starts one service thread (device = the calling
 * thread's): calls of the same kind/type arriving within windowMicros of each other, up to maxBatch of them, run as ONE
 * batch of blocks — what Kanzi's <= 64 EncodingTask / DecodingTask pool threads produce, one block each
 * (K/io/CompressedOutputStream.java:537-573).  Results are identical to uncoalesced calls.  maxBatch <= 1 stops the service.
 * kzg_coalescing_stats: requests served by the service so far (*batches = launches they were folded into). */
int kzg_set_coalescing(int maxBatch, int windowMicros);
int64_t kzg_coalescing_stats(int64_t* batches);

/* ---- batched whole-chain entries (SURVEY.md §8b "batched forms", §8f rank 1-2) ------------------------
 * What CompressedOutputStream / CompressedInputStream + EncodingTask / DecodingTask produce and
 * consume (COS:236-313,733-1054; CIS:359-515,1025-1378) for `nTransforms` chained ids + one entropy id,
 * with every block of the input in flight at once on the calling thread's device.  Host buffers.
 * kzg_compress: returns the .knz byte length (<0 on error).  kzg_decompress: returns decoded bytes.
 * flags: KZG_FLAG_BWT_ASREF | KZG_FLAG_XXH32 / KZG_FLAG_XXH64.  KZG_T_NONE entries of `transforms` are dropped before anything else, as
 * TransformFactory.getType drops the NONE tokens of a "-t A+NONE+B" list (K/transform/TransformFactory.java:140-153).
 * Capacities: kzg_compress needs outCap >= kzg_compress_bound(n, blockSize) to be sure of success (a smaller buffer
 * fails with -KZG_ERR_WRITE_FILE only if the stream really does not fit).  kzg_decompress needs outCap >= the decoded size
 * and never writes beyond `out + outCap`: every block but the last decodes to blockSize bytes, the last to what is left
 * (-KZG_ERR_WRITE_FILE when the stream holds more than outCap bytes).  Host libraries should export
 * CUDA_DEVICE_MAX_CONNECTIONS=32 before the CUDA context exists (the LZ stages run block groups on up to 32 streams). */
int64_t kzg_compress(const uint8_t* in, int64_t n, const int32_t* transforms, int32_t nTransforms, int32_t entropy,
                     int32_t blockSize, int32_t flags, uint8_t* out, int64_t outCap);
int64_t kzg_decompress(const uint8_t* in, int64_t nBytes, int32_t flags, uint8_t* out, int64_t outCap);
/* upper bound for kzg_compress output */
int64_t kzg_compress_bound(int64_t n, int32_t blockSize);

/* Same, device-resident (bench `value`: inputs already in HBM; pointers are device pointers, 16-byte
 * aligned; d_in of kzg_decompress_dev must be readable up to the next multiple of 8 bytes beyond nBytes, which every
 * cudaMalloc'ed buffer is).  The codec runs on the calling thread's stream; these calls synchronise before returning.
 * timing (optional, may be NULL): ms spent in [0] transforms, [1] entropy, [2] container assembly. */
int64_t kzg_compress_dev(const uint8_t* d_in, int64_t n, const int32_t* transforms, int32_t nTransforms, int32_t entropy,
                         int32_t blockSize, int32_t flags, uint8_t* d_out, int64_t outCap, float* timing3);
int64_t kzg_decompress_dev(const uint8_t* d_in, int64_t nBytes, const uint8_t* h_in, int32_t flags, uint8_t* d_out, int64_t outCap,
                           float* timing3);

/* ---- block sharding across GPUs (CompressedOutputStream.java:1024-1035) ---------------------------------
 * Blocks are independent: GPU r of G encodes blocks r, r+G, ... of a joint stream with kzg_compress[_dev] and reports each
 * record's bit length; the lengths are the one thing ranks exchange (NCCL all-gather), after which every rank knows the bit
 * offset of every record in the joint stream.
 * kzg_last_block_bits: record bit lengths (5 + lw + written) of the calling thread's last compress call; returns the count.
 * kzg_stream_index: host-side walk of a .knz: bit offset / bit length of every block record, *headerBits = first record. */
int32_t kzg_last_block_bits(int64_t* recBits, int32_t cap);
int32_t kzg_stream_index(const uint8_t* in, int64_t nBytes, int64_t* recBit, int64_t* recBits, int32_t cap, int64_t* headerBits);

/* per-kernel CUDA-event timing of the calling thread's calls (measurement aid: bench.py's per-kernel roofline).
 * kzg_set_profiling(1) starts a collection, kzg_profile_json() ends it: {"kernel name": [launches, total ms], ...}. */
void kzg_set_profiling(int on);
const char* kzg_profile_json(void);

/* the CUDA stream of the calling thread as a cudaStream_t cast to void* (for event timing by callers) */
void* kzg_stream(void);

#ifdef __cplusplus
}
#endif
#endif /* KZG_H */
