// kzg_api.cu — C ABI of libkanzi_b200 (include/kzg.h): per-thread workspace, the Sequence / block /
// stream host logic that surrounds the kernels, and the batched whole-chain entries.
//
// Host logic mirrors, for the calling Java code's benefit:
//   K/transform/Sequence.java:56-127,137-207   (stage order, skip flags, slice rotation)
//   K/transform/TransformFactory.java:240-351   (which codec an id names, nbFunctions)
//   K/io/CompressedOutputStream.java:236-313,733-1054 and CompressedInputStream.java:359-515,1025-1378
//   (stream header, block records) for kzg_compress / kzg_decompress.
// No CPU codec exists here: every byte of transform / entropy work runs in the CUDA kernels.
#include "kzg_common.cuh"
#include "kzg_entropy.cuh"
#include "kzg_transforms.cuh"
#include "kzg_container.cuh"
#include "kzg_stages.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <vector>
#include <string>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>

// ---- per-thread workspace --------------------------------------------------------------------------------
#define KZG_DEC_MAXG 16            // block groups of the decode (streams of their own)
struct Workspace {
  bool init = false;
  int device = 0;
  cudaStream_t stream = nullptr;
  u8* dArena = nullptr; size_t dCap = 0, dOff = 0;
  u8* hPinned = nullptr; size_t hCap = 0, hOff = 0;
  u8* ioBuf[2] = {nullptr, nullptr}; size_t ioCap[2] = {0, 0};      // device copies of the host-buffer entry points' streams (grow-only)
  i64 launches = 0;
  char err[512] = {0};
  cudaStream_t side[KZG_DEC_MAXG] = {}; cudaEvent_t sideEv[KZG_DEC_MAXG + 1] = {}; bool sideInit = false;
  bool lziExclusive = false;       // decode batch with no more blocks than SMs: the LZ inverse's record-chain CTAs take an SM each (lz_inverse.cu)
  std::vector<i64> lastRecBits;    // record bit lengths (5 + lw + written) of the blocks of this thread's last kzg_compress* call
};
static thread_local Workspace W;

void kzg_set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  vsnprintf(W.err, sizeof(W.err), fmt, ap);
  va_end(ap);
}
void kzg_count_launch(int n) { W.launches += n; }

// ---- per-kernel event timing --------------------------------------------------------------------------------------------
struct ProfRec { const char* name; cudaEvent_t e0, e1; };
static thread_local bool gProfOn = false;
static thread_local std::vector<ProfRec> gProf;
static thread_local std::string gProfJson;
void kzg_prof_begin(const char* name, cudaStream_t s) {
  if (!gProfOn) return;
  ProfRec r; r.name = name; r.e0 = nullptr; r.e1 = nullptr;
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) { cudaGetLastError(); return; }
  cudaEventRecord(r.e0, s);
  gProf.push_back(r);
}
void kzg_prof_end(cudaStream_t s) {
  if (!gProfOn || gProf.empty()) return;
  cudaEventRecord(gProf.back().e1, s);
}

static int ws_init() {
  if (W.init) return 0;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    kzg_set_error("no CUDA device: %s (libkanzi_b200 has no CPU fallback)", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    cudaGetLastError();
    return -KZG_ERR_NO_DEVICE;
  }
  if (W.device < 0 || W.device >= n) { kzg_set_error("device %d out of range (%d CUDA devices)", W.device, n); return -KZG_ERR_INVALID_PARAM; }
  CUDA_TRY(cudaSetDevice(W.device));
  CUDA_TRY(cudaStreamCreateWithFlags(&W.stream, cudaStreamNonBlocking));
  W.init = true;
  return 0;
}

// device staging for kzg_compress / kzg_decompress (host buffers): kept between calls, cudaMalloc/cudaFree cost milliseconds each
// side streams of the grouped decode (created once per calling thread; the first group gets the most urgent one)
static int ws_side_init() {
  if (W.sideInit) return 0;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  for (int g = 0; g < KZG_DEC_MAXG; g++) CUDA_TRY(cudaStreamCreateWithPriority(&W.side[g], cudaStreamNonBlocking, std::min(lo, hi + g)));
  for (int g = 0; g <= KZG_DEC_MAXG; g++) CUDA_TRY(cudaEventCreateWithFlags(&W.sideEv[g], cudaEventDisableTiming));
  W.sideInit = true;
  return 0;
}

static u8* ws_io(int which, size_t bytes) {
  if (bytes > W.ioCap[which]) {
    if (W.ioBuf[which]) { cudaStreamSynchronize(W.stream); cudaFree(W.ioBuf[which]); W.ioBuf[which] = nullptr; W.ioCap[which] = 0; }
    const size_t want = bytes + bytes / 16 + (1 << 20);
    if (cudaMalloc((void**)&W.ioBuf[which], want) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    W.ioCap[which] = want;
  }
  return W.ioBuf[which];
}

static int ws_reserve(size_t dBytes, size_t hBytes) {
  if (dBytes > W.dCap) {
    if (W.dArena) { CUDA_TRY(cudaStreamSynchronize(W.stream)); cudaFree(W.dArena); W.dArena = nullptr; W.dCap = 0; }
    const size_t want = dBytes + dBytes / 8 + (1 << 20);
    cudaError_t e = cudaMalloc((void**)&W.dArena, want);
    if (e != cudaSuccess) { kzg_set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e)); cudaGetLastError(); return -KZG_ERR_CREATE_CODEC; }
    W.dCap = want;
  }
  if (hBytes > W.hCap) {
    if (W.hPinned) { CUDA_TRY(cudaStreamSynchronize(W.stream)); cudaFreeHost(W.hPinned); W.hPinned = nullptr; W.hCap = 0; }
    const size_t want = hBytes + hBytes / 8 + (1 << 16);
    cudaError_t e = cudaMallocHost((void**)&W.hPinned, want);
    if (e != cudaSuccess) { kzg_set_error("cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e)); cudaGetLastError(); return -KZG_ERR_CREATE_CODEC; }
    W.hCap = want;
  }
  W.dOff = 0; W.hOff = 0;
  return 0;
}
static inline size_t rnd(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }
template <typename T> static T* dalloc(size_t count) {
  const size_t bytes = rnd(count * sizeof(T));
  if (W.dOff + bytes > W.dCap) { kzg_set_error("internal: device arena exhausted (%zu + %zu > %zu)", W.dOff, bytes, W.dCap); return nullptr; }
  T* p = (T*)(W.dArena + W.dOff); W.dOff += bytes; return p;
}
template <typename T> static T* halloc(size_t count) {
  const size_t bytes = rnd(count * sizeof(T));
  if (W.hOff + bytes > W.hCap) { kzg_set_error("internal: pinned arena exhausted"); return nullptr; }
  T* p = (T*)(W.hPinned + W.hOff); W.hOff += bytes; return p;
}
// timing events of one call: destroyed on every exit path
struct EvSet {
  cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};
  int create(int n) { for (int i = 0; i < n; i++) if (cudaEventCreate(&e[i]) != cudaSuccess) return -KZG_ERR_PROCESS_BLOCK; return 0; }
  cudaEvent_t& operator[](int i) { return e[i]; }
  ~EvSet() { for (int i = 0; i < 4; i++) if (e[i]) cudaEventDestroy(e[i]); }
};
#define NN(p) do { if ((p) == nullptr) return -KZG_ERR_CREATE_CODEC; } while (0)

// ---- ids / sizes (TransformFactory, getMaxEncodedLength of each codec) --------------------------------------
static bool xf_known(int t) {
  switch (t) { case KZG_T_NONE: case KZG_T_LZ: case KZG_T_LZX: case KZG_T_ROLZ: case KZG_T_BWT: case KZG_T_RANK: case KZG_T_MTFT:
               case KZG_T_SRT: case KZG_T_ZRLT: case KZG_T_LZP: case KZG_T_RLT: case KZG_T_ROLZX: return true; default: return false; }
}
static bool ent_known(int e) {
  switch (e) { case KZG_E_NONE: case KZG_E_HUFFMAN: case KZG_E_ANS0: case KZG_E_ANS1: case KZG_E_FPAQ: return true; default: return false; }
}
static i32 xf_max_len(int t, i32 n) {
  switch (t) {
    case KZG_T_LZ: case KZG_T_LZX: return ((n <= 1024) ? n + 16 : n + (n / 64)) + 2;   // LZCodec.java:961-964
    case KZG_T_LZP: return (n <= 1024) ? n + 16 : n + (n / 64);                       // LZCodec.java:1283-1285
    case KZG_T_RLT: return (n <= 512) ? n + 32 : n;                                   // RLT.java:355-357
    case KZG_T_ROLZX: return (n <= 16384) ? n + 1024 : n + (n / 32);                  // ROLZCodec.java:1417-1421
    case KZG_T_ROLZ: return (n <= 512) ? n + 64 : n;                                  // ROLZCodec.java:1001-1003
    case KZG_T_BWT: return n + 33;                                                    // BWTBlockCodec.java:222-224
    case KZG_T_SRT: return n + 1024;                                                  // SRT.java:364-366
    default: return n;                                                                // Null / SBRT / ZRLT
  }
}
// TransformFactory.newFunction (:240-270): ids of the Sequence's functions
static int seq_functions(const i32* transforms, int nT, int* out) {
  int slots[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < nT && i < 8; i++) slots[i] = transforms[i];
  int nbtr = 0;
  for (int i = 0; i < 8; i++) if (slots[i] != KZG_T_NONE) nbtr++;
  if (nbtr == 0) nbtr = 1;
  int k = 0;
  for (int i = 0; i < nbtr; i++) if (slots[i] != KZG_T_NONE || i == 0) out[k++] = slots[i];
  return k;
}
static i32 seq_max_len(const int* fn, int nf, i32 n) {      // Sequence.getMaxEncodedLength (:219-230)
  i32 req = n;
  for (int i = 0; i < nf; i++) req = std::max(req, xf_max_len(fn[i], req));
  return req;
}
static int ent_chunk_size(int e) {
  switch (e) { case KZG_E_HUFFMAN: case KZG_E_ANS0: return 16384; case KZG_E_ANS1: case KZG_E_FPAQ: return 4 << 20; default: return 1 << 30; }
}
static int ent_segs_per_chunk(int e) { return e == KZG_E_HUFFMAN ? 6 : 2; }

// ---- batch state ------------------------------------------------------------------------------------------
struct Batch {
  int nBlocks = 0;
  KzgBlock* hBlocks = nullptr;     // pinned
  KzgBlock* dBlocks = nullptr;
  int* dResult = nullptr;
  u8* dEnabled = nullptr; u8* hEnabled = nullptr;
  int* dDstLimit = nullptr; int* hDstLimit = nullptr;
  i32 maxLen = 0;                  // largest block length at any stage
  const u8* lazyHost = nullptr; u8* lazyDev = nullptr; i64 lazyN = 0; i32 lazyBlock = 0;   // see KzgXfParams (stage 0 of a host-buffer encode)
};

static int batch_upload(Batch& bt) {
  CUDA_TRY(cudaMemcpyAsync(bt.dBlocks, bt.hBlocks, sizeof(KzgBlock) * bt.nBlocks, cudaMemcpyHostToDevice, W.stream));
  return 0;
}
static int batch_download(Batch& bt) {
  CUDA_TRY(cudaMemcpyAsync(bt.hBlocks, bt.dBlocks, sizeof(KzgBlock) * bt.nBlocks, cudaMemcpyDeviceToHost, W.stream));
  CUDA_TRY(cudaStreamSynchronize(W.stream));
  return 0;
}

// scratch the transform stages of a chain need, per block of at most `maxLen` bytes
struct XfScratch { size_t perBlock = 0; int tk = 0, m = 0, ml = 0; size_t hashInts = 0; size_t aux32 = 0; };
static XfScratch xf_scratch_size(const int* fn, int nf, i32 maxLen, bool forward) {
  XfScratch s;
  for (int i = 0; i < nf; i++) {
    if (fn[i] == KZG_T_LZ || fn[i] == KZG_T_LZX) {
      if (forward) {
        s.tk = (int)rnd(std::max(maxLen / 5, 256) + 16, 16); s.m = (int)rnd((size_t)maxLen + 16, 16); s.ml = (int)rnd((size_t)maxLen / 2 + 64, 16);
        s.perBlock = std::max(s.perBlock, (size_t)s.tk + s.m + s.ml);
        kzg_lzf_scratch(maxLen, &s.perBlock);
      } else {
        kzg_lzi_scratch(maxLen, &s.perBlock, &s.aux32);
      }
    }
    kzg_stage_scratch(fn[i], maxLen, forward, &s.perBlock, &s.hashInts, &s.aux32);
  }
  s.perBlock = rnd(s.perBlock, 256);
  return s;
}

static int run_transform_stage(Batch& bt, int type, int stage, bool forward, const XfScratch& xs, u8* dScratch, i32* dHash, i32* dAux32, int flags) {
  KzgXfParams P;
  P.result = bt.dResult; P.enabled = bt.dEnabled; P.dstLimit = bt.dDstLimit;
  P.scratch = dScratch; P.scratchStride = (i64)xs.perBlock; P.tkStride = xs.tk; P.mStride = xs.m; P.mLenStride = xs.ml;
  P.hashBuf = dHash; P.aux32 = dAux32; P.aux32Stride = (i64)xs.aux32; P.flags = (type == KZG_T_RLT) ? flags : (flags & ~0xF00);   // bits 8-11: ctx["entropy"] + 1, RLT only
  const bool lazy = forward && stage == 0 && bt.lazyHost && (type == KZG_T_LZ || type == KZG_T_LZX);
  P.lazyHost = lazy ? bt.lazyHost : nullptr; P.lazyDev = bt.lazyDev; P.lazyN = bt.lazyN; P.lazyBlock = bt.lazyBlock;
  int r = 0;
  switch (type) {
    case KZG_T_LZ: case KZG_T_LZX:
      if (forward) {
        const char* dbgEnv = getenv("KZG_DEBUG");          // developer aid: bit 0 stats printf, bits 1-2 disable walker shortcuts, bit 3 = serial walker only
        if (dbgEnv) P.flags |= (atoi(dbgEnv) << 12);
        r = kzg_lz_forward2_launch(W.stream, bt.dBlocks, bt.nBlocks, P, type == KZG_T_LZX, bt.maxLen);
      }
      else {
        static const char* dbgEnvI = getenv("KZG_DEBUG");         // developer aid: bit 0 per-block statistics of the token chase
        if (dbgEnvI) P.flags |= (atoi(dbgEnvI) << 12);
        if (W.lziExclusive) P.flags |= KZG_XF_LZI_EXCLUSIVE;
        r = kzg_lz_inverse_launch(W.stream, bt.dBlocks, bt.nBlocks, P, bt.maxLen);
      }
      break;
    default:
      r = kzg_stage_launch(W.stream, type, forward, bt.dBlocks, bt.nBlocks, P, bt.maxLen);
      break;
  }
  if (r < 0) return r;
  return kzg_commit_launch(W.stream, bt.dBlocks, bt.nBlocks, bt.dResult, bt.dEnabled, stage, forward ? 1 : 0);
}

// entropy-stage scratch for encoding blocks of at most maxPost bytes
struct EntScratch { int maxChunks = 0, spc = 0, segsPerBlock = 0; size_t hdrStride = 0, payStride = 0, tabStride = 0; };
static EntScratch ent_scratch_size(int entropy, i32 maxPost, bool encode) {
  EntScratch e;
  const int cs = ent_chunk_size(entropy);
  e.maxChunks = std::max(1, (int)(((i64)maxPost + cs - 1) / cs));
  e.spc = ent_segs_per_chunk(entropy);
  e.segsPerBlock = 1 + e.spc * e.maxChunks;
  switch (entropy) {
    case KZG_E_ANS0: e.hdrStride = 512; e.payStride = encode ? 2 * 16384 + 16 + 16 : 0; break;
    case KZG_E_HUFFMAN: e.hdrStride = 1024; e.payStride = encode ? 4 * 6160 + 16 : 0; break;
    case KZG_E_ANS1:
      e.hdrStride = 256 * 480 + 64;
      e.payStride = encode ? (size_t)std::max(std::min(cs + (cs >> 3), 2 * maxPost), 65536) + 16 + 16 : 0;
      e.tabStride = encode ? kzg_ans1_enc_tab_u32() : kzg_ans1_dec_tab_u32();
      break;
    case KZG_E_FPAQ:
      e.hdrStride = 64; e.payStride = encode ? (size_t)cs + (cs >> 3) + 64 : 0;
      break;
    default: break;
  }
  if (!encode) e.hdrStride = 0;
  return e;
}

static int run_entropy_encode(Batch& bt, int entropy, const EntScratch& es, u8* dHdr, u8* dPay, u32* dTab, KzgSeg* dSegs) {
  if (entropy == KZG_E_NONE) return 0;
  KzgEntParams P;
  memset(&P, 0, sizeof(P));
  P.entropy = entropy; P.chunkSize = ent_chunk_size(entropy); P.maxChunks = es.maxChunks;
  P.hdrBuf = dHdr; P.hdrStride = (int)es.hdrStride; P.payBuf = dPay; P.payStride = (int)es.payStride;
  P.tabBuf = dTab; P.tabStride = (i64)es.tabStride; P.segs = dSegs; P.segsPerBlock = es.segsPerBlock;
  switch (entropy) {
    case KZG_E_ANS0: return kzg_ans_encode_launch(W.stream, bt.dBlocks, bt.nBlocks, P, 0);
    case KZG_E_ANS1: return kzg_ans_encode_launch(W.stream, bt.dBlocks, bt.nBlocks, P, 1);
    case KZG_E_HUFFMAN: return kzg_huff_encode_launch(W.stream, bt.dBlocks, bt.nBlocks, P);
    case KZG_E_FPAQ: return kzg_fpaq_encode_launch(W.stream, bt.dBlocks, bt.nBlocks, P);
    default: return -KZG_ERR_INVALID_CODEC;
  }
}

static int run_entropy_decode(Batch& bt, int entropy, const EntScratch& es, const u8* dStream, KzgChunkInfo* dChunks, u32* dTab) {
  if (entropy == KZG_E_NONE) return kzg_rawbits_launch(W.stream, bt.dBlocks, bt.nBlocks, dStream);
  KzgEntParams P;
  memset(&P, 0, sizeof(P));
  P.entropy = entropy; P.chunkSize = ent_chunk_size(entropy); P.maxChunks = es.maxChunks;
  P.stream = dStream; P.chunks = dChunks; P.tabBuf = dTab; P.tabStride = (i64)es.tabStride;
  switch (entropy) {
    case KZG_E_ANS0: return kzg_ans_decode_launch(W.stream, bt.dBlocks, bt.nBlocks, P, 0);
    case KZG_E_ANS1: return kzg_ans_decode_launch(W.stream, bt.dBlocks, bt.nBlocks, P, 1);
    case KZG_E_HUFFMAN: return kzg_huff_decode_launch(W.stream, bt.dBlocks, bt.nBlocks, P);
    case KZG_E_FPAQ: return kzg_fpaq_decode_launch(W.stream, bt.dBlocks, bt.nBlocks, P);
    default: return -KZG_ERR_INVALID_CODEC;
  }
}

// ============================================================================================================
// library entry points
// ============================================================================================================
// The LZ forward rounds run block groups on up to 32 side streams; with the default of 8 hardware work queues streams
// alias onto each other and a group's one-warp stitch kernel serialises behind another group's parse.  The host process
// decides: export CUDA_DEVICE_MAX_CONNECTIONS=32 before the CUDA context exists (the Python package and
// bench.py do).  The library itself never touches the environment.

extern "C" {

int kzg_abi_version(void) { return 1; }
int kzg_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }
// frees everything the calling thread's workspace owns (device of the workspace must be current)
static void ws_release() {
  if (!W.init) { W = Workspace(); return; }
  cudaSetDevice(W.device);
  cudaStreamSynchronize(W.stream);
  if (W.sideInit) {
    for (int g = 0; g < KZG_DEC_MAXG; g++) { cudaStreamSynchronize(W.side[g]); cudaStreamDestroy(W.side[g]); }
    for (int g = 0; g <= KZG_DEC_MAXG; g++) cudaEventDestroy(W.sideEv[g]);
  }
  kzg_lzf_release();
  if (W.dArena) cudaFree(W.dArena);
  if (W.hPinned) cudaFreeHost(W.hPinned);
  for (int i = 0; i < 2; i++) if (W.ioBuf[i]) cudaFree(W.ioBuf[i]);
  cudaStreamDestroy(W.stream);
  cudaGetLastError();
  W = Workspace();
}
int kzg_set_device(int device) {
  if (W.init && W.device != device) ws_release();
  W.device = device;
  return ws_init();
}
const char* kzg_last_error(void) { return W.err; }
int64_t kzg_launch_count(int reset) { const i64 v = W.launches; if (reset) W.launches = 0; return v; }
// Block sharding: the bit length of every block record of the calling thread's last kzg_compress /
// kzg_compress_dev call, in block order.  A rank that holds blocks b = r, r + G, ... of a joint stream all-gathers these and
// every rank knows every block's bit offset in the joint stream (CompressedOutputStream.java:1024-1035 appends records
// back to back, bit-granular).  Returns the block count.
int32_t kzg_last_block_bits(int64_t* recBits, int32_t cap) {
  const int n = (int)W.lastRecBits.size();
  for (int i = 0; i < n && i < cap; i++) recBits[i] = W.lastRecBits[i];
  return n;
}
// Walks a .knz stream on the host: bit offset and bit length (5 + lw + written) of every block record, *headerBits = where the
// first record starts.  Returns the block count (the end-of-stream marker is not a block) or < 0.
int32_t kzg_stream_index(const uint8_t* in, int64_t nBytes, int64_t* recBit, int64_t* recBits, int32_t cap, int64_t* headerBits) {
  if (in == nullptr || nBytes < 20) return -KZG_ERR_INVALID_FILE;
  auto rd = [&](u64 pos, int n) -> u64 { u64 v = 0; for (int i = 0; i < n; i++, pos++) v = (v << 1) | ((u64)(in[pos >> 3] >> (7 - (pos & 7))) & 1u); return v; };
  const u64 nbits = (u64)nBytes * 8;
  if ((u32)rd(0, 32) != 0x4B414E5Au) return -KZG_ERR_INVALID_FILE;
  const int szMask = (int)rd(32 + 4 + 2 + 5 + 48 + 28, 2);
  u64 pos = 32 + 4 + 2 + 5 + 48 + 28 + 2 + 16ull * szMask + 15 + 24;
  if (headerBits) *headerBits = (i64)pos;
  int nb = 0;
  for (;;) {
    if (pos + 8 > nbits) return -KZG_ERR_READ_FILE;
    const int lw = (int)rd(pos, 5) + 3;
    if (pos + 5 + lw > nbits) return -KZG_ERR_READ_FILE;
    const u64 written = rd(pos + 5, lw);
    if (written == 0) break;
    if (pos + 5 + lw + written > nbits) return -KZG_ERR_READ_FILE;
    if (nb < cap) { if (recBit) recBit[nb] = (i64)pos; if (recBits) recBits[nb] = (i64)(5 + lw + written); }
    nb++;
    pos += 5 + lw + written;
  }
  return nb;
}
// Per-kernel CUDA-event timing of the calling thread's next calls (bench.py: roofline of the top kernel).  on != 0 starts a
// fresh collection; kzg_profile_json() synchronises the device and returns {"kernel": [launches, total ms], ...} for the
// launch sites wrapped in KZG_PROF (the hot kernels of every stage), then clears the collection.
void kzg_set_profiling(int on) {
  for (auto& r : gProf) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  gProf.clear();
  gProfOn = on != 0;
}
const char* kzg_profile_json(void) {
  cudaDeviceSynchronize();
  std::vector<std::pair<std::string, std::pair<int, double>>> agg;
  for (auto& r : gProf) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) { cudaGetLastError(); continue; }
    bool found = false;
    for (auto& a : agg) if (a.first == r.name) { a.second.first++; a.second.second += ms; found = true; break; }
    if (!found) agg.push_back({r.name, {1, (double)ms}});
  }
  gProfJson = "{";
  for (size_t i = 0; i < agg.size(); i++) {
    char buf[160];
    snprintf(buf, sizeof(buf), "%s\"%s\": [%d, %.6f]", i ? ", " : "", agg[i].first.c_str(), agg[i].second.first, agg[i].second.second);
    gProfJson += buf;
  }
  if (getenv("KZG_TIMELINE") && !gProf.empty()) {       // developer aid: every launch's start / end in ms since the first one
    gProfJson += ", \"_timeline\": [";
    for (size_t i = 0; i < gProf.size(); i++) {
      float a = 0, b = 0;
      cudaEventElapsedTime(&a, gProf[0].e0, gProf[i].e0); cudaEventElapsedTime(&b, gProf[0].e0, gProf[i].e1);
      char buf[160];
      snprintf(buf, sizeof(buf), "%s[\"%s\", %.3f, %.3f]", i ? ", " : "", gProf[i].name, a, b);
      gProfJson += buf;
    }
    gProfJson += "]";
  }
  gProfJson += "}";
  for (auto& r : gProf) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  gProf.clear();
  return gProfJson.c_str();
}
void* kzg_stream(void) { if (ws_init() < 0) return nullptr; return (void*)W.stream; }
int32_t kzg_transform_max_encoded_len(int type, int32_t n) { return xf_known(type) ? xf_max_len(type, n) : -KZG_ERR_INVALID_CODEC; }
// Worst case of any supported chain: a block whose entropy stage does not pay is stored as a "transformed copy"
// (COS:926-973) = its transform output (at most Sequence.getMaxEncodedLength: ROLZX n + n/32, LZ n + n/64 + 2, BWT + 33, SRT + 1024) plus a
// record header (5 + 32 bits of length, mode, skip flags, 4 length bytes, checksum).
int64_t kzg_compress_bound(int64_t n, int32_t blockSize) {
  const i64 nb = (n + blockSize - 1) / std::max(blockSize, 1) + 1;
  return n + n / 32 + nb * (2 + 33 + 1024 + 16) + 64;
}

// ---- per-block calls: ByteTransform.forward / inverse, EntropyEncoder.encode, EntropyDecoder.decode -----------------------
// One request = one call of the Java interface.  Requests of the same kind run as ONE batch of blocks through the batched
// kernels (the kernels take block arrays anyway); a direct call is a batch of one.  With coalescing on
// (kzg_set_coalescing), concurrent callers — Kanzi's EncodingTask / DecodingTask pool threads, one block each,
// K/io/CompressedOutputStream.java:537-573 — are gathered by a service thread and launched together, so the path the
// Java host really drives is not one launch train per block.
enum { KZG_REQ_XF_FWD = 0, KZG_REQ_XF_INV = 1, KZG_REQ_ENT_ENC = 2, KZG_REQ_ENT_DEC = 3 };
struct KzgReq {
  int kind, type;
  kzg_ctx* ctx; int flags;
  const uint8_t* src; int32_t srcLen;        // transform input / entropy encoder input / decoder bit string
  uint8_t* dst; int32_t dstLen, dstCap;      // transform output slice; entropy: dstCap = encoder out capacity (bytes) or n to decode
  int64_t inBits;                            // decoder: bits available
  int32_t srcUsed = 0, dstUsed = 0; int64_t bits = 0;
  int64_t ret = 0;
  bool done = false;
};

static int transform_batch(std::vector<KzgReq*>& rq) {
  const int n = (int)rq.size();
  const int type = rq[0]->type; const bool forward = rq[0]->kind == KZG_REQ_XF_FWD;
  int r = ws_init(); if (r < 0) { for (auto q : rq) q->ret = r; return 0; }
  std::vector<int> live;                                   // requests that reach the device
  i32 maxLen = 0;
  for (int i = 0; i < n; i++) {
    KzgReq& Q = *rq[i];
    Q.srcUsed = 0; Q.dstUsed = 0;
    if (!xf_known(type)) { Q.ret = -KZG_ERR_INVALID_CODEC; continue; }
    if (Q.srcLen == 0) { Q.ret = 1; continue; }               // every codec: `if (input.length == 0) return true`
    if (Q.srcLen < 0 || Q.dstLen < 0 || Q.dstCap < 0 || Q.dstLen > Q.dstCap || Q.src == nullptr || Q.dst == nullptr) { Q.ret = 0; continue; }
    // host-evaluated slice preconditions (the parts of each codec's guard block that depend on lengths only)
    const int pre = kzg_stage_precheck(type, forward, Q.ctx, Q.srcLen, Q.dstLen, Q.dstCap);
    if (pre <= 0) { Q.ret = pre; continue; }
    live.push_back(i);
    maxLen = std::max(maxLen, std::max(Q.srcLen, Q.dstCap));
  }
  if (live.empty()) return 0;
  auto failAll = [&](int code) { for (int i : live) rq[i]->ret = code; return 0; };
  const size_t nl = live.size();
  const int fn[1] = {type};
  XfScratch xs = xf_scratch_size(fn, 1, maxLen, forward);
  const size_t cap = rnd((size_t)maxLen + 64);
  const size_t need = nl * (2 * cap + xs.perBlock + (xs.hashInts + xs.aux32) * 4 + sizeof(KzgBlock) + 64) + 65536;
  r = ws_reserve(need, nl * (sizeof(KzgBlock) + 32) + 4096); if (r < 0) return failAll(r);
  Batch bt; bt.nBlocks = (int)nl; bt.maxLen = maxLen;
  bt.hBlocks = halloc<KzgBlock>(nl); bt.hEnabled = halloc<u8>(nl); bt.hDstLimit = halloc<int>(nl);
  bt.dBlocks = dalloc<KzgBlock>(nl); bt.dResult = dalloc<int>(2 * nl); bt.dEnabled = dalloc<u8>(nl); bt.dDstLimit = dalloc<int>(nl);
  u8* dA = dalloc<u8>(nl * cap); u8* dB = dalloc<u8>(nl * cap);
  u8* dScratch = dalloc<u8>(nl * xs.perBlock + 16); i32* dHash = dalloc<i32>(nl * xs.hashInts + 4); i32* dAux = dalloc<i32>(nl * xs.aux32 + 4);
  if (!bt.hBlocks || !bt.hEnabled || !bt.hDstLimit || !bt.dBlocks || !bt.dResult || !bt.dEnabled || !bt.dDstLimit || !dA || !dB || !dScratch || !dHash || !dAux)
    return failAll(-KZG_ERR_CREATE_CODEC);
  auto cu = [&](cudaError_t e) { if (e != cudaSuccess) { kzg_set_error("transform batch: %s", cudaGetErrorString(e)); cudaGetLastError(); return false; } return true; };
  for (size_t k = 0; k < nl; k++) {
    KzgReq& Q = *rq[live[k]];
    KzgBlock& B = bt.hBlocks[k];
    memset(&B, 0, sizeof(B));
    B.cur = dA + k * cap; B.alt = dB + k * cap; B.curLen = Q.srcLen; B.cap = (forward && type != KZG_T_RLT) ? Q.dstLen : Q.dstCap; B.origLen = Q.srcLen; B.skipFlags = 0xFF;
    B.dataType = Q.ctx ? Q.ctx->dataType : 0; B.aux0 = nullptr; B.aux1 = B.cur; B.stagesLeft = 2;
    // dst.length of the slice, except where the codec looks at dst.array.length: every inverse but LZP's (LZCodec.java:1133) and RLT.forward (RLT.java:117)
    bt.hEnabled[k] = 1; bt.hDstLimit[k] = ((forward && type != KZG_T_RLT) || type == KZG_T_LZP || type == KZG_T_ROLZX) ? Q.dstLen : Q.dstCap;   // (ROLZX.inverse: szBlock > output.length, ROLZCodec.java:1306)
    if (!cu(cudaMemsetAsync(B.cur + Q.srcLen, 0, cap - Q.srcLen, W.stream)) || !cu(cudaMemcpyAsync(B.cur, Q.src, Q.srcLen, cudaMemcpyHostToDevice, W.stream)))
      return failAll(-KZG_ERR_PROCESS_BLOCK);
  }
  if (!cu(cudaMemcpyAsync(bt.dEnabled, bt.hEnabled, nl, cudaMemcpyHostToDevice, W.stream)) ||
      !cu(cudaMemcpyAsync(bt.dDstLimit, bt.hDstLimit, nl * sizeof(int), cudaMemcpyHostToDevice, W.stream))) return failAll(-KZG_ERR_PROCESS_BLOCK);
  r = batch_upload(bt); if (r < 0) return failAll(r);
  // (commit's inverse branch flags a failed inverse as a block error; here a false result is reported as 0)
  r = run_transform_stage(bt, type, 0, forward, xs, dScratch, dHash, dAux, rq[live[0]]->flags & ~KZG_FLAG_BWT_ASREF);   // the slice guards were evaluated by kzg_stage_precheck
  if (r < 0) return failAll(r);
  int* hres = halloc<int>(2 * nl); if (!hres) return failAll(-KZG_ERR_CREATE_CODEC);
  if (!cu(cudaMemcpyAsync(hres, bt.dResult, sizeof(int) * 2 * nl, cudaMemcpyDeviceToHost, W.stream))) return failAll(-KZG_ERR_PROCESS_BLOCK);
  r = batch_download(bt); if (r < 0) return failAll(r);
  for (size_t k = 0; k < nl; k++) {
    KzgReq& Q = *rq[live[k]];
    const KzgBlock& B = bt.hBlocks[k];
    if (Q.ctx) Q.ctx->dataType = B.dataType;
    if (B.status != 0 && !(B.status == -KZG_ERR_PROCESS_BLOCK && !forward && hres[2 * k] == 0)) { Q.ret = B.status; continue; }
    if (hres[2 * k] == 1) {
      if (hres[2 * k + 1] > Q.dstCap) { kzg_set_error("transform output %d exceeds dst capacity %d", hres[2 * k + 1], Q.dstCap); Q.ret = -KZG_ERR_PROCESS_BLOCK; continue; }
      if (!cu(cudaMemcpyAsync(Q.dst, B.cur, hres[2 * k + 1], cudaMemcpyDeviceToHost, W.stream))) { Q.ret = -KZG_ERR_PROCESS_BLOCK; continue; }
      Q.srcUsed = Q.srcLen; Q.dstUsed = hres[2 * k + 1]; Q.ret = 1;
    } else Q.ret = 0;
  }
  if (!cu(cudaStreamSynchronize(W.stream))) return failAll(-KZG_ERR_PROCESS_BLOCK);
  return 0;
}

static int entropy_encode_batch(std::vector<KzgReq*>& rq) {
  const int type = rq[0]->type;
  int r = ws_init(); if (r < 0) { for (auto q : rq) q->ret = r; return 0; }
  std::vector<int> live;
  i32 maxN = 0;
  for (int i = 0; i < (int)rq.size(); i++) {
    KzgReq& Q = *rq[i];
    Q.bits = 0;
    if (!ent_known(type)) { Q.ret = -KZG_ERR_INVALID_CODEC; continue; }
    if (Q.srcLen < 0 || Q.src == nullptr || Q.dst == nullptr) { Q.ret = -1; continue; }     // Java: encode returns -1 on bad arguments
    if (Q.srcLen == 0) { Q.ret = 0; continue; }
    live.push_back(i); maxN = std::max(maxN, Q.srcLen);
  }
  if (live.empty()) return 0;
  auto failAll = [&](int code) { for (int i : live) rq[i]->ret = code; return 0; };
  const size_t nl = live.size();
  EntScratch es = ent_scratch_size(type, maxN, true);
  const size_t cap = rnd((size_t)maxN + 64);
  const size_t outBytes = rnd(2 * (size_t)maxN + (256 << 10));     // ANS1 on noise: 256 context headers + up to 2 bytes per symbol
  const size_t need = nl * (cap + outBytes + (size_t)es.maxChunks * (es.hdrStride + es.payStride + es.tabStride * 4) +
                            (size_t)es.segsPerBlock * sizeof(KzgSeg) + sizeof(KzgBlock) + 64) + 65536;
  r = ws_reserve(need, nl * (sizeof(KzgBlock) + 16) + 4096); if (r < 0) return failAll(r);
  Batch bt; bt.nBlocks = (int)nl; bt.maxLen = maxN;
  bt.hBlocks = halloc<KzgBlock>(nl); bt.dBlocks = dalloc<KzgBlock>(nl);
  u8* dIn = dalloc<u8>(nl * cap); u8* dOut = dalloc<u8>(nl * outBytes);
  u8* dHdr = dalloc<u8>(nl * es.maxChunks * es.hdrStride + 16); u8* dPay = dalloc<u8>(nl * es.maxChunks * es.payStride + 16);
  u32* dTab = dalloc<u32>(nl * es.maxChunks * es.tabStride + 4); KzgSeg* dSegs = dalloc<KzgSeg>(nl * es.segsPerBlock);
  u8* dHdrBytes = dalloc<u8>(nl * KZG_HDR_STRIDE + 16); i64* dTotal = dalloc<i64>(2);
  if (!bt.hBlocks || !bt.dBlocks || !dIn || !dOut || !dHdr || !dPay || !dTab || !dSegs || !dHdrBytes || !dTotal) return failAll(-KZG_ERR_CREATE_CODEC);
  auto cu = [&](cudaError_t e) { if (e != cudaSuccess) { kzg_set_error("entropy batch: %s", cudaGetErrorString(e)); cudaGetLastError(); return false; } return true; };
  for (size_t k = 0; k < nl; k++) {
    KzgReq& Q = *rq[live[k]];
    KzgBlock& B = bt.hBlocks[k];
    memset(&B, 0, sizeof(B));
    B.cur = dIn + k * cap; B.alt = nullptr; B.curLen = Q.srcLen; B.cap = (i32)cap; B.origLen = Q.srcLen; B.entropy = type;
    B.srcBit = (i64)(k * outBytes) * 8;                    // container == 0: the payload lands at this bit of dOut
    if (!cu(cudaMemsetAsync(B.cur + Q.srcLen, 0, cap - Q.srcLen, W.stream)) || !cu(cudaMemcpyAsync(B.cur, Q.src, Q.srcLen, cudaMemcpyHostToDevice, W.stream)))
      return failAll(-KZG_ERR_PROCESS_BLOCK);
  }
  if (!cu(cudaMemsetAsync(dOut, 0, nl * outBytes, W.stream)) || !cu(cudaMemsetAsync(dSegs, 0, sizeof(KzgSeg) * nl * es.segsPerBlock, W.stream))) return failAll(-KZG_ERR_PROCESS_BLOCK);
  r = batch_upload(bt); if (r < 0) return failAll(r);
  r = run_entropy_encode(bt, type, es, dHdr, dPay, dTab, dSegs); if (r < 0) return failAll(r);
  r = kzg_assemble_launch(W.stream, bt.dBlocks, (int)nl, dSegs, es.segsPerBlock, dHdrBytes, 1, 0, dOut, 0, dTotal, (i64)(nl * outBytes) - 8); if (r < 0) return failAll(r);
  r = batch_download(bt); if (r < 0) return failAll(r);
  for (size_t k = 0; k < nl; k++) {
    KzgReq& Q = *rq[live[k]];
    const KzgBlock& B = bt.hBlocks[k];
    if (B.status != 0) { Q.ret = B.status; continue; }
    const i64 bits = B.entBits, bytes = (bits + 7) >> 3;
    if (bytes > (i64)Q.dstCap || bytes > (i64)outBytes - 8) { kzg_set_error("entropy output (%lld bytes) exceeds capacity", (long long)bytes); Q.ret = -KZG_ERR_PROCESS_BLOCK; continue; }
    if (!cu(cudaMemcpyAsync(Q.dst, dOut + k * outBytes, (size_t)bytes, cudaMemcpyDeviceToHost, W.stream))) { Q.ret = -KZG_ERR_PROCESS_BLOCK; continue; }
    Q.bits = bits; Q.ret = Q.srcLen;
  }
  if (!cu(cudaStreamSynchronize(W.stream))) return failAll(-KZG_ERR_PROCESS_BLOCK);
  return 0;
}

static int entropy_decode_batch(std::vector<KzgReq*>& rq) {
  const int type = rq[0]->type;
  int r = ws_init(); if (r < 0) { for (auto q : rq) q->ret = r; return 0; }
  std::vector<int> live;
  i32 maxN = 0; size_t maxIn = 0;
  for (int i = 0; i < (int)rq.size(); i++) {
    KzgReq& Q = *rq[i];
    Q.bits = 0;
    if (!ent_known(type)) { Q.ret = -KZG_ERR_INVALID_CODEC; continue; }
    if (Q.dstCap < 0 || Q.src == nullptr || Q.dst == nullptr || Q.inBits < 0) { Q.ret = -1; continue; }
    if (Q.dstCap == 0) { Q.ret = 0; continue; }
    live.push_back(i); maxN = std::max(maxN, Q.dstCap); maxIn = std::max(maxIn, (size_t)((Q.inBits + 7) >> 3));
  }
  if (live.empty()) return 0;
  auto failAll = [&](int code) { for (int i : live) rq[i]->ret = code; return 0; };
  const size_t nl = live.size();
  EntScratch es = ent_scratch_size(type, maxN, false);
  const size_t inCap = rnd(maxIn + 64), cap = rnd((size_t)maxN + 64);
  const size_t need = nl * (inCap + cap + (size_t)es.maxChunks * (sizeof(KzgChunkInfo) + es.tabStride * 4) + sizeof(KzgBlock) + 64) + 65536;
  r = ws_reserve(need, nl * (sizeof(KzgBlock) + 16) + 4096); if (r < 0) return failAll(r);
  Batch bt; bt.nBlocks = (int)nl; bt.maxLen = maxN;
  bt.hBlocks = halloc<KzgBlock>(nl); bt.dBlocks = dalloc<KzgBlock>(nl);
  u8* dIn = dalloc<u8>(nl * inCap); u8* dOut = dalloc<u8>(nl * cap);
  KzgChunkInfo* dChunks = dalloc<KzgChunkInfo>(nl * es.maxChunks + 1); u32* dTab = dalloc<u32>(nl * es.maxChunks * es.tabStride + 4);
  if (!bt.hBlocks || !bt.dBlocks || !dIn || !dOut || !dChunks || !dTab) return failAll(-KZG_ERR_CREATE_CODEC);
  auto cu = [&](cudaError_t e) { if (e != cudaSuccess) { kzg_set_error("entropy batch: %s", cudaGetErrorString(e)); cudaGetLastError(); return false; } return true; };
  for (size_t k = 0; k < nl; k++) {
    KzgReq& Q = *rq[live[k]];
    KzgBlock& B = bt.hBlocks[k];
    memset(&B, 0, sizeof(B));
    const size_t inBytes = (size_t)((Q.inBits + 7) >> 3);
    B.cur = dOut + k * cap; B.curLen = Q.dstCap; B.cap = (i32)cap; B.preLen = Q.dstCap; B.entropy = type;
    B.srcBit = (i64)(k * inCap) * 8; B.srcBits = Q.inBits;
    if (!cu(cudaMemsetAsync(dIn + k * inCap + inBytes, 0, inCap - inBytes, W.stream)) || !cu(cudaMemcpyAsync(dIn + k * inCap, Q.src, inBytes, cudaMemcpyHostToDevice, W.stream)))
      return failAll(-KZG_ERR_PROCESS_BLOCK);
  }
  r = batch_upload(bt); if (r < 0) return failAll(r);
  r = run_entropy_decode(bt, type, es, dIn, dChunks, dTab); if (r < 0) return failAll(r);
  r = batch_download(bt); if (r < 0) return failAll(r);
  for (size_t k = 0; k < nl; k++) {
    KzgReq& Q = *rq[live[k]];
    const KzgBlock& B = bt.hBlocks[k];
    if (B.status != 0) { Q.ret = 0; continue; }          // the Java decoders signal corrupt input by a short count / exception
    if (!cu(cudaMemcpyAsync(Q.dst, B.cur, Q.dstCap, cudaMemcpyDeviceToHost, W.stream))) { Q.ret = 0; continue; }
    Q.bits = B.entBits; Q.ret = Q.dstCap;
  }
  if (!cu(cudaStreamSynchronize(W.stream))) return failAll(-KZG_ERR_PROCESS_BLOCK);
  return 0;
}

static void run_batch(std::vector<KzgReq*>& rq) {
  switch (rq[0]->kind) {
    case KZG_REQ_XF_FWD: case KZG_REQ_XF_INV: transform_batch(rq); break;
    case KZG_REQ_ENT_ENC: entropy_encode_batch(rq); break;
    default: entropy_decode_batch(rq); break;
  }
}

// ---- the coalescing service: one thread per process, owns a workspace of its own on the device it was enabled for -----------------
struct Coalescer {
  std::mutex m; std::condition_variable cvWork, cvDone;
  std::vector<KzgReq*> pending;
  std::thread worker; bool running = false, stop = false;
  int maxBatch = 0, windowMicros = 0, device = 0;
  std::atomic<long long> batches{0}, requests{0};
  ~Coalescer() { if (running) { { std::lock_guard<std::mutex> g(m); stop = true; } cvWork.notify_all(); if (worker.joinable()) worker.join(); } }
  void loop() {
    W.device = device;
    std::unique_lock<std::mutex> lk(m);
    for (;;) {
      cvWork.wait(lk, [&] { return stop || !pending.empty(); });
      if (stop) return;
      // gather: wait for company until the batch is full or the window (counted from the first arrival) has passed
      const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(windowMicros);
      while ((int)pending.size() < maxBatch && !stop) { if (cvWork.wait_until(lk, deadline) == std::cv_status::timeout) break; }
      std::vector<KzgReq*> take, rest;
      const KzgReq* f = pending[0];
      for (KzgReq* q : pending) {
        if ((int)take.size() < maxBatch && q->kind == f->kind && q->type == f->type && q->flags == f->flags) take.push_back(q); else rest.push_back(q);
      }
      pending.swap(rest);
      lk.unlock();
      run_batch(take);
      batches++; requests += (long long)take.size();
      lk.lock();
      for (KzgReq* q : take) q->done = true;
      cvDone.notify_all();
    }
  }
  void submit(KzgReq& q) {
    std::unique_lock<std::mutex> lk(m);
    pending.push_back(&q);
    cvWork.notify_one();
    cvDone.wait(lk, [&] { return q.done; });
  }
};
static Coalescer gCo;

static void submit_or_run(KzgReq& q) {
  if (gCo.running) { gCo.submit(q); return; }
  std::vector<KzgReq*> one{&q};
  run_batch(one);
}

int kzg_set_coalescing(int maxBatch, int windowMicros) {
  std::unique_lock<std::mutex> lk(gCo.m);
  if (maxBatch <= 1) {                       // off: stop the service once its queue is empty
    if (gCo.running) {
      gCo.stop = true; gCo.cvWork.notify_all();
      lk.unlock(); gCo.worker.join(); lk.lock();
      gCo.running = false; gCo.stop = false;
    }
    return 0;
  }
  gCo.maxBatch = std::min(maxBatch, 1024); gCo.windowMicros = std::max(0, windowMicros);
  if (!gCo.running) {
    gCo.device = W.device;                   // the calling thread's device (kzg_set_device first for another one)
    gCo.running = true;
    gCo.worker = std::thread([] { gCo.loop(); });
  }
  return 0;
}
int64_t kzg_coalescing_stats(int64_t* batches) { if (batches) *batches = gCo.batches.load(); return gCo.requests.load(); }

int kzg_transform_forward(int type, kzg_ctx* ctx, const uint8_t* src, int32_t srcLen, uint8_t* dst, int32_t dstLen, int32_t dstCap,
                          int32_t* srcUsed, int32_t* dstUsed) {
  KzgReq q; q.kind = KZG_REQ_XF_FWD; q.type = type; q.ctx = ctx; q.flags = ctx ? ctx->flags : 0; q.src = src; q.srcLen = srcLen; q.dst = dst; q.dstLen = dstLen; q.dstCap = dstCap; q.inBits = 0;
  submit_or_run(q);
  if (srcUsed) *srcUsed = q.srcUsed; if (dstUsed) *dstUsed = q.dstUsed;
  return (int)q.ret;
}
int kzg_transform_inverse(int type, kzg_ctx* ctx, const uint8_t* src, int32_t srcLen, uint8_t* dst, int32_t dstLen, int32_t dstCap,
                          int32_t* srcUsed, int32_t* dstUsed) {
  KzgReq q; q.kind = KZG_REQ_XF_INV; q.type = type; q.ctx = ctx; q.flags = ctx ? ctx->flags : 0; q.src = src; q.srcLen = srcLen; q.dst = dst; q.dstLen = dstLen; q.dstCap = dstCap; q.inBits = 0;
  submit_or_run(q);
  if (srcUsed) *srcUsed = q.srcUsed; if (dstUsed) *dstUsed = q.dstUsed;
  return (int)q.ret;
}

int kzg_bwt_forward(const uint8_t* src, int32_t n, uint8_t* dst, int32_t* primaryIndexes8) {
  int r = ws_init(); if (r < 0) return r;
  return kzg_bwt_raw(W.stream, true, src, n, dst, primaryIndexes8);
}
int kzg_bwt_inverse(const uint8_t* src, int32_t n, uint8_t* dst, const int32_t* primaryIndexes8) {
  int r = ws_init(); if (r < 0) return r;
  return kzg_bwt_raw(W.stream, false, src, n, dst, (int32_t*)primaryIndexes8);
}

int64_t kzg_entropy_encode(int type, kzg_ctx* ctx, const uint8_t* src, int32_t n, uint8_t* out, int64_t outCap, int64_t* outBits) {
  KzgReq q; q.kind = KZG_REQ_ENT_ENC; q.type = type; q.ctx = ctx; q.flags = 0; q.src = src; q.srcLen = n; q.dst = out; q.dstLen = 0;
  q.dstCap = (int32_t)std::min<int64_t>(outCap, 0x7FFFFFFF); q.inBits = 0;
  submit_or_run(q);
  if (outBits) *outBits = q.bits;
  return q.ret;
}
int32_t kzg_entropy_decode(int type, kzg_ctx* ctx, const uint8_t* in, int64_t inBits, int64_t* bitsUsed, uint8_t* dst, int32_t n) {
  KzgReq q; q.kind = KZG_REQ_ENT_DEC; q.type = type; q.ctx = ctx; q.flags = 0; q.src = in; q.srcLen = 0; q.dst = dst; q.dstLen = 0; q.dstCap = n; q.inBits = inBits;
  submit_or_run(q);
  if (bitsUsed) *bitsUsed = q.bits;
  return (int32_t)q.ret;
}

// ---- whole streams ------------------------------------------------------------------------------------------------
// stream header, COS:236-313 (checksum kind 0).  Returns header byte count (whole bytes: 160 + 16*szMask bits).
static int stream_header(u8* h, int entropy, u64 transformType, i32 blockSize, i64 inputSize, int chkKind) {
  u64 acc = 0; int nacc = 0, nb = 0;
  auto put = [&](u64 v, int n) {
    for (int i = n - 1; i >= 0; i--) { acc = (acc << 1) | ((v >> i) & 1); if (++nacc == 8) { h[nb++] = (u8)acc; acc = 0; nacc = 0; } }
  };
  put(0x4B414E5A, 32); put(7, 4); put((u64)chkKind, 2); put((u64)entropy, 5); put(transformType, 48); put((u64)((u32)blockSize >> 4), 28);
  int szMask = 0;
  if (inputSize != 0 && inputSize < (1LL << 48)) {
    if (inputSize >= (1LL << 32)) szMask = 3;
    else {
      i64 isz = inputSize;
      if (isz > (1LL << 30)) { isz >>= 4; szMask++; }
      int lg = 0; while ((2LL << lg) <= isz) lg++;
      szMask += (lg >> 4) + 1;
    }
  }
  put((u64)szMask, 2);
  if (szMask > 0) put((u64)inputSize, 16 * szMask);
  put(0, 15);
  const u32 HASH = 0x1E35A7BDu;
  u32 c = HASH * (0x01030507u * 7u);
  c = kzg_mix32(c, HASH, (u32)chkKind);
  c = kzg_mix32(c, HASH, (u32)entropy);
  c = kzg_mix32(c, HASH, (u32)(transformType >> 32));
  c = kzg_mix32(c, HASH, (u32)transformType);
  c = kzg_mix32(c, HASH, (u32)blockSize);
  if (szMask > 0) { c = kzg_mix32(c, HASH, (u32)((u64)inputSize >> 32)); c = kzg_mix32(c, HASH, (u32)inputSize); }
  c = (c >> 23) ^ (c >> 3);
  put(c & 0xFFFFFF, 24);
  return nb;
}

// h_in (optional): the input is still on the host; it is uploaded here — whole, or, when the chain starts with LZ and the
// batch is large, a sample per eighth of every block now and the blocks themselves by the LZ stage on its group streams.
// Device memory one call may take for its arena: what is free now (plus what this thread's arena already holds), less a margin.
// The driver is only asked (cudaMemGetInfo costs milliseconds, more with many streams alive) when the whole batch does not fit the
// arena this thread already owns.  KZG_ARENA_MB (developer / test knob) lowers the budget so that small inputs exercise the slicing.
static size_t arena_budget(size_t wantAll) {
  const char* e = getenv("KZG_ARENA_MB");
  if (!e && wantAll <= W.dCap) return ~(size_t)0 >> 2;      // everything fits what is already reserved
  size_t freeB = 0, totalB = 0;
  if (cudaMemGetInfo(&freeB, &totalB) != cudaSuccess) { cudaGetLastError(); freeB = (size_t)8 << 30; }
  size_t budget = (size_t)((double)(freeB + W.dCap) * 0.85);
  if (e) budget = std::min(budget, (size_t)std::max(1, atoi(e)) << 20);
  return budget;
}

static int64_t compress_impl(const uint8_t* d_in, int64_t n, const int32_t* transforms, int32_t nTransforms, int32_t entropy,
                             int32_t blockSize, int32_t flags, uint8_t* d_out, int64_t outCap, float* timing3, const uint8_t* h_in) {
  if (n < 0 || nTransforms < 0 || nTransforms > 8 || !ent_known(entropy)) return -KZG_ERR_INVALID_PARAM;
  if (blockSize > (1 << 30) || blockSize < 1024 || (blockSize & -16) != blockSize) return -KZG_ERR_BLOCK_SIZE;    // COS:165-174
  for (int i = 0; i < nTransforms; i++) if (!xf_known(transforms[i])) return -KZG_ERR_INVALID_CODEC;
  int r = ws_init(); if (r < 0) return r;
  // TransformFactory.getType (:140-153) drops the NONE tokens of a "-t A+NONE+B" list before the type is formed; the Sequence and the
  // stream header only ever see the compacted list
  i32 tc[8]; int nc = 0;
  for (int i = 0; i < nTransforms; i++) if (transforms[i] != KZG_T_NONE) tc[nc++] = transforms[i];
  int fn[8]; const int nf = seq_functions(tc, nc, fn);
  u64 transformType = 0;
  for (int i = 0; i < nc; i++) transformType |= ((u64)tc[i] << (42 - 6 * i));
  W.lastRecBits.clear();
  const int nBlocks = (int)((n + blockSize - 1) / blockSize);
  const i32 maxBlock = (i32)std::min<i64>(n, blockSize);
  const i32 required = seq_max_len(fn, nf, maxBlock);
  const size_t cap = rnd((size_t)required + 64);
  bool anyXf = false; for (int i = 0; i < nf; i++) if (fn[i] != KZG_T_NONE) anyXf = true;
  XfScratch xs = xf_scratch_size(fn, nf, required, true);
  EntScratch es = ent_scratch_size(entropy, required, true);
  // The arena holds every block of a slice at once (descriptors, ping-pong buffers, codec scratch, entropy scratch).  An input
  // whose blocks do not all fit runs as several slices of whole blocks, one after the other, each appending its records where
  // the previous slice's ended (records are independent and bit-granular: COS:1024-1035) — the stream is the same.
  const size_t perBlock = (sizeof(KzgBlock) + 64) + (anyXf ? cap * (nf >= 2 ? 2 : 1) : 0) + xs.perBlock + (xs.hashInts + xs.aux32) * 4 +
                          (size_t)es.maxChunks * (es.hdrStride + es.payStride + es.tabStride * 4) + (size_t)es.segsPerBlock * sizeof(KzgSeg) + KZG_HDR_STRIDE;
  const size_t budget = arena_budget((size_t)std::max(nBlocks, 1) * perBlock + (1 << 20));
  const int sliceBlocks = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(nBlocks, 1), (budget > (2u << 20) ? budget - (2u << 20) : 0) / (perBlock + perBlock / 8)));
  const size_t nbMax = (size_t)std::max(std::min(nBlocks, sliceBlocks), 1);
  r = ws_reserve(nbMax * perBlock + (1 << 20), nbMax * (sizeof(KzgBlock) + 32) + 8192); if (r < 0) return r;
  if ((flags & KZG_FLAG_XXH32) && (flags & KZG_FLAG_XXH64)) return -KZG_ERR_INVALID_PARAM;
  const int chkBytes = (flags & KZG_FLAG_XXH32) ? 4 : ((flags & KZG_FLAG_XXH64) ? 8 : 0);       // ctx["checksum"] 32 / 64 (COS:193-204)
  u8 hdr[64];
  const int hdrLen = stream_header(hdr, entropy, transformType, blockSize, n, chkBytes / 4);
  if (outCap < hdrLen + 2) return -KZG_ERR_WRITE_FILE;
  EvSet ev;
  if (timing3) { r = ev.create(4); if (r < 0) return r; timing3[0] = timing3[1] = timing3[2] = 0; }
  CUDA_TRY(cudaMemsetAsync(d_out, 0, (size_t)outCap, W.stream));
  CUDA_TRY(cudaMemcpyAsync(d_out, hdr, hdrLen, cudaMemcpyHostToDevice, W.stream));
  i64 totalBits = (i64)hdrLen * 8 + 8;
  for (int b0 = 0; b0 < nBlocks; b0 += sliceBlocks) {
    const int cnt = std::min(sliceBlocks, nBlocks - b0);
    const size_t nb = (size_t)cnt;
    const i64 off0 = (i64)b0 * blockSize;
    const i64 sliceN = std::min<i64>((i64)cnt * blockSize, n - off0);
    const uint8_t* sIn = d_in + off0;
    const uint8_t* sHost = h_in ? h_in + off0 : nullptr;
    W.dOff = 0; W.hOff = 0;                        // the arena is reused slice after slice (the stream is in order)
    const bool lazyIn = sHost && nf >= 1 && (fn[0] == KZG_T_LZ || fn[0] == KZG_T_LZX) && cnt >= 8 && blockSize >= (1 << 18);
    if (sHost && sliceN > 0) {
      // the input is still on the host: uploaded here — whole, or, when the chain starts with LZ and the batch is large, a sample
      // per eighth of every block now and the blocks themselves by the LZ stage on its group streams
      if (!lazyIn) CUDA_TRY(cudaMemcpyAsync((u8*)sIn, sHost, (size_t)sliceN, cudaMemcpyHostToDevice, W.stream));
      else {
        for (int b = 0; b < cnt; b++) {       // the samples the LZ stage orders its blocks by (they also hold the block's magic bytes)
          const i64 off = (i64)b * blockSize;
          const i64 len = std::min<i64>(blockSize, sliceN - off);
          const i64 eighth = len / LZF_SAMPLES;
          if (eighth < LZF_SAMPLE) { CUDA_TRY(cudaMemcpyAsync((u8*)sIn + off, sHost + off, (size_t)len, cudaMemcpyHostToDevice, W.stream)); continue; }
          CUDA_TRY(cudaMemcpy2DAsync((u8*)sIn + off, (size_t)eighth, sHost + off, (size_t)eighth, LZF_SAMPLE, LZF_SAMPLES, cudaMemcpyHostToDevice, W.stream));
        }
      }
    }
    Batch bt; bt.nBlocks = cnt; bt.maxLen = required;
    if (lazyIn) { bt.lazyHost = sHost; bt.lazyDev = (u8*)sIn; bt.lazyN = sliceN; bt.lazyBlock = blockSize; }
    bt.hBlocks = halloc<KzgBlock>(nb); NN(bt.hBlocks);
    bt.hEnabled = halloc<u8>(nb); NN(bt.hEnabled);
    bt.hDstLimit = halloc<int>(nb); NN(bt.hDstLimit);
    bt.dBlocks = dalloc<KzgBlock>(nb); NN(bt.dBlocks);
    bt.dResult = dalloc<int>(2 * nb); NN(bt.dResult);
    bt.dEnabled = dalloc<u8>(nb); NN(bt.dEnabled);
    bt.dDstLimit = dalloc<int>(nb); NN(bt.dDstLimit);
    u8* dA = anyXf ? dalloc<u8>(nb * cap) : nullptr; if (anyXf) NN(dA);
    u8* dB = (anyXf && nf >= 2) ? dalloc<u8>(nb * cap) : nullptr; if (anyXf && nf >= 2) NN(dB);
    u8* dScratch = dalloc<u8>(nb * xs.perBlock + 16); NN(dScratch);
    i32* dHash = dalloc<i32>(nb * xs.hashInts + 4); NN(dHash);
    i32* dAux = dalloc<i32>(nb * xs.aux32 + 4); NN(dAux);
    u8* dHdr = dalloc<u8>(nb * es.maxChunks * es.hdrStride + 16); NN(dHdr);
    u8* dPay = dalloc<u8>(nb * es.maxChunks * es.payStride + 16); NN(dPay);
    u32* dTab = dalloc<u32>(nb * es.maxChunks * es.tabStride + 4); NN(dTab);
    KzgSeg* dSegs = dalloc<KzgSeg>(nb * es.segsPerBlock); NN(dSegs);
    u8* dHdrBytes = dalloc<u8>(nb * KZG_HDR_STRIDE + 16); NN(dHdrBytes);
    i64* dTotal = dalloc<i64>(2); NN(dTotal);
    for (int b = 0; b < cnt; b++) {
      KzgBlock& B = bt.hBlocks[b];
      memset(&B, 0, sizeof(B));
      const i64 off = (i64)b * blockSize;
      const i32 len = (i32)std::min<i64>(blockSize, sliceN - off);
      B.cur = (u8*)sIn + off; B.aux0 = B.cur;
      B.alt = dA ? dA + (size_t)b * cap : nullptr; B.aux1 = dB ? dB + (size_t)b * cap : B.alt;
      B.curLen = len; B.cap = (i32)cap - 64; B.origLen = len; B.skipFlags = 0xFF; B.entropy = entropy; B.chkBytes = chkBytes;
      const bool small = len <= 15;                             // COS:764-767: raw copy block
      if (small) { B.mode = 0x80; B.entropy = KZG_E_NONE; B.skipFlags = 0x7F; }   // NONE&NONE copy block: its NullTransform "succeeds" (COS:764-767, 792-817)
      bt.hEnabled[b] = small ? 0 : 1;
      // dst slice length of the transform stage: EncodingTask grows `buffer` to the Sequence's requiredSize and never shrinks it
      // (COS:806-811), so with the blocks of one stream handled in order every block after the first sees the full-block size
      bt.hDstLimit[b] = (b0 + b == 0) ? seq_max_len(fn, nf, len) : required;
    }
    CUDA_TRY(cudaMemcpyAsync(bt.dEnabled, bt.hEnabled, nb, cudaMemcpyHostToDevice, W.stream));
    CUDA_TRY(cudaMemcpyAsync(bt.dDstLimit, bt.hDstLimit, nb * sizeof(int), cudaMemcpyHostToDevice, W.stream));
    CUDA_TRY(cudaMemsetAsync(dSegs, 0, sizeof(KzgSeg) * nb * es.segsPerBlock, W.stream));
    r = batch_upload(bt); if (r < 0) return r;
    if (timing3) CUDA_TRY(cudaEventRecord(ev[0], W.stream));
    // ctx["dataType"] from the block's magic number (COS:795-804)
    r = kzg_magic_launch(W.stream, bt.dBlocks, cnt); if (r < 0) return r;
    if (chkBytes) {       // block checksums of the original bytes (COS:745-755): four dependent chains per block, on a stream of their own next to the transforms
      if (lazyIn) CUDA_TRY(cudaMemcpyAsync((u8*)sIn, sHost, (size_t)sliceN, cudaMemcpyHostToDevice, W.stream));      // (the hash needs the whole input now; the LZ stage's own uploads then rewrite the same bytes)
      r = ws_side_init(); if (r < 0) return r;
      CUDA_TRY(cudaEventRecord(W.sideEv[KZG_DEC_MAXG], W.stream));
      CUDA_TRY(cudaStreamWaitEvent(W.side[0], W.sideEv[KZG_DEC_MAXG], 0));
      r = kzg_xxh_launch(W.side[0], bt.dBlocks, cnt, 0); if (r < 0) return r;
      CUDA_TRY(cudaEventRecord(W.sideEv[0], W.side[0]));
    }
    // Sequence.forward: a NONE-only chain is a copy that always succeeds (NullTransform) -> skip bit 7 cleared
    for (int i = 0; i < nf; i++) {
      if (fn[i] == KZG_T_NONE) { r = kzg_null_forward_launch(W.stream, bt.dBlocks, cnt, bt.dEnabled, i); if (r < 0) return r; continue; }
      r = run_transform_stage(bt, fn[i], i, true, xs, dScratch, dHash, dAux, (flags & ~0xF00) | KZG_CTX_ENTROPY(entropy)); if (r < 0) return r;
    }
    if (timing3) CUDA_TRY(cudaEventRecord(ev[1], W.stream));
    r = run_entropy_encode(bt, entropy, es, dHdr, dPay, dTab, dSegs); if (r < 0) return r;
    if (timing3) CUDA_TRY(cudaEventRecord(ev[2], W.stream));
    if (chkBytes) CUDA_TRY(cudaStreamWaitEvent(W.stream, W.sideEv[0], 0));
    r = kzg_assemble_launch(W.stream, bt.dBlocks, cnt, dSegs, es.segsPerBlock, dHdrBytes, nf, 1, d_out, totalBits - 8, dTotal, outCap - 8);
    if (r < 0) return r;
    if (timing3) CUDA_TRY(cudaEventRecord(ev[3], W.stream));
    CUDA_TRY(cudaMemcpyAsync(&totalBits, dTotal, sizeof(i64), cudaMemcpyDeviceToHost, W.stream));
    r = batch_download(bt); if (r < 0) return r;
    for (int b = 0; b < cnt; b++)
      if (bt.hBlocks[b].status != 0) { kzg_set_error("block %d failed with status %d", b0 + b + 1, bt.hBlocks[b].status); return bt.hBlocks[b].status; }
    if (totalBits < 0) return -KZG_ERR_PROCESS_BLOCK;
    for (int b = 0; b < cnt; b++) {          // COS:1024-1035: 5 bits of (lw - 3), lw bits of `written`, then `written` bits
      const i64 written = bt.hBlocks[b].written;
      int lw = 3; if (written >= 8) { lw = 0; while ((2LL << lw) <= (written >> 3)) lw++; lw += 4; }
      W.lastRecBits.push_back(5 + lw + written);
    }
    if (timing3) for (int i = 0; i < 3; i++) { float ms = 0; CUDA_TRY(cudaEventElapsedTime(&ms, ev[i], ev[i + 1])); timing3[i] += ms; }
  }
  if (nBlocks == 0) CUDA_TRY(cudaStreamSynchronize(W.stream));
  const i64 bytes = (totalBits + 7) >> 3;
  if (bytes > outCap - 8) { kzg_set_error("compressed stream (%lld bytes) exceeds capacity %lld", (long long)bytes, (long long)outCap); return -KZG_ERR_WRITE_FILE; }
  return bytes;
}

int64_t kzg_compress_dev(const uint8_t* d_in, int64_t n, const int32_t* transforms, int32_t nTransforms, int32_t entropy,
                         int32_t blockSize, int32_t flags, uint8_t* d_out, int64_t outCap, float* timing3) {
  return compress_impl(d_in, n, transforms, nTransforms, entropy, blockSize, flags, d_out, outCap, timing3, nullptr);
}

// ---- decode side: the host walks the container (CIS:359-515, 1025-1095, 1127-1167), the device does the rest -------
struct HostBits {
  const u8* p; u64 nbits; u64 pos = 0; bool bad = false;
  u64 read(int n) {
    if (pos + (u64)n > nbits) { bad = true; pos += n; return 0; }
    u64 v = 0;
    for (int i = 0; i < n; i++, pos++) v = (v << 1) | ((p[pos >> 3] >> (7 - (pos & 7))) & 1);
    return v;
  }
};

// d_in / d_out: device stream and destination.  h_in: the stream on the host (the container walk runs there).  copyIn: the
// device copy of the stream is not there yet, every group uploads its own byte range first.  h_out (optional): every group
// downloads its blocks as soon as they are done.
static int64_t decompress_impl(const uint8_t* d_in, int64_t nBytes, const uint8_t* h_in, int32_t flags, uint8_t* d_out, int64_t outCap,
                               float* timing3, bool copyIn, uint8_t* h_out) {
  int r = ws_init(); if (r < 0) return r;
  if (nBytes < 20 || h_in == nullptr) return -KZG_ERR_INVALID_FILE;
  HostBits hb{h_in, (u64)nBytes * 8};
  if ((u32)hb.read(32) != 0x4B414E5Au) { kzg_set_error("Invalid stream type"); return -KZG_ERR_INVALID_FILE; }
  const int bsVersion = (int)hb.read(4);
  if (bsVersion != 7) { kzg_set_error("bitstream version %d not supported (7 only)", bsVersion); return -KZG_ERR_STREAM_VERSION; }
  const int chkSize = (int)hb.read(2);
  if (chkSize == 3) { kzg_set_error("Invalid bitstream, incorrect block checksum size"); return -KZG_ERR_INVALID_FILE; }
  const int chkBytes = 4 * chkSize;                 // 1 = XXHash32, 2 = XXHash64 (CIS:382-395)
  const int entropy = (int)hb.read(5);
  const u64 transformType = hb.read(48);
  const i32 blockSize = (i32)(hb.read(28) << 4);
  if (blockSize < 1024 || blockSize > (1 << 30)) return -KZG_ERR_BLOCK_SIZE;
  const int szMask = (int)hb.read(2);
  i64 outputSize = 0;
  if (szMask != 0) outputSize = (i64)hb.read(16 * szMask);
  hb.read(15);
  const u32 cksum1 = (u32)hb.read(24);
  {
    const u32 HASH = 0x1E35A7BDu;
    u32 c = HASH * (0x01030507u * (u32)bsVersion);
    c = kzg_mix32(c, HASH, (u32)chkSize); c = kzg_mix32(c, HASH, (u32)entropy);
    c = kzg_mix32(c, HASH, (u32)(transformType >> 32)); c = kzg_mix32(c, HASH, (u32)transformType);
    c = kzg_mix32(c, HASH, (u32)blockSize);
    if (szMask > 0) { c = kzg_mix32(c, HASH, (u32)((u64)outputSize >> 32)); c = kzg_mix32(c, HASH, (u32)outputSize); }
    c = (c >> 23) ^ (c >> 3);
    if (cksum1 != (c & 0xFFFFFF) || hb.bad) { kzg_set_error("Invalid bitstream, checksum mismatch"); return -KZG_ERR_CRC_CHECK; }
  }
  if (!ent_known(entropy)) return -KZG_ERR_INVALID_CODEC;
  i32 tr[8]; for (int i = 0; i < 8; i++) tr[i] = (i32)((transformType >> (42 - 6 * i)) & 63);
  for (int i = 0; i < 8; i++) if (!xf_known(tr[i])) { kzg_set_error("transform id %d not supported", tr[i]); return -KZG_ERR_INVALID_CODEC; }
  int fn[8]; const int nf = seq_functions(tr, 8, fn);

  // walk the block records
  struct Rec { i64 payBit, payBits; i32 preLen; int skipFlags; int entropy; bool rawCopy; u64 xxh; };
  std::vector<Rec> recs;
  const i32 maxTransformLength = std::min(std::max(blockSize + blockSize / 2, 2048), 1 << 30);
  while (true) {
    const int lr = (int)hb.read(5) + 3;
    const i64 written = (i64)hb.read(lr);
    if (hb.bad) { kzg_set_error("truncated stream"); return -KZG_ERR_READ_FILE; }
    if (written == 0) break;
    if (written < 8) return -KZG_ERR_BLOCK_SIZE;
    const u64 blockStart = hb.pos;
    const int mode = (int)hb.read(8);
    int skipFlags = 0; bool hasSkip = false, transformedCopy = false;
    const bool copyBlock = (mode & 0x80) != 0;
    if (copyBlock) {
      if (mode & 0x10) { transformedCopy = true; if (nf > 4) hasSkip = true; else skipFlags = ((mode << 4) | 0x0F) & 0xFF; }
    } else if (mode & 0x10) hasSkip = true;
    else skipFlags = ((mode << 4) | 0x0F) & 0xFF;
    const int dataSize = 1 + ((mode >> 5) & 0x03);
    const int headerSize = 1 + (hasSkip ? 1 : 0) + dataSize + 1;
    if (written < (i64)headerSize * 8) return -KZG_ERR_BLOCK_SIZE;
    if (hasSkip) skipFlags = (int)hb.read(8);
    u32 pre = 0;
    for (int i = 0; i < dataSize; i++) pre = (pre << 8) | (u32)hb.read(8);
    const u32 hck = (u32)hb.read(8);
    const u32 HASH = 0x1E35A7BDu;
    u32 c = HASH * 0x01030507u;
    c = kzg_mix32(c, HASH, (u32)(mode & 0xFF)); c = kzg_mix32(c, HASH, (u32)(skipFlags & 0xFF)); c = kzg_mix32(c, HASH, pre);
    c = kzg_mix32(c, HASH, (u32)((u64)written >> 32)); c = kzg_mix32(c, HASH, (u32)written);
    c = (c >> 23) ^ (c >> 3);
    if (hb.bad || hck != (c & 0xFF)) { kzg_set_error("Invalid bitstream, block header checksum mismatch"); return -KZG_ERR_CRC_CHECK; }
    if ((i32)pre < 0 || (i32)pre > maxTransformLength) { kzg_set_error("Invalid compressed block length: %d", (i32)pre); return -KZG_ERR_READ_FILE; }
    if (((written + 7) >> 3) > (i64)pre + headerSize + chkBytes) return -KZG_ERR_BLOCK_SIZE;      // CIS:1158-1165
    if (pre == 0) break;                                         // "last block is empty" (CIS:1223-1227)
    Rec rc;
    rc.xxh = 0;
    if (chkBytes) { if (written < (i64)(headerSize + chkBytes) * 8) return -KZG_ERR_BLOCK_SIZE; rc.xxh = (chkBytes == 4) ? hb.read(32) : ((hb.read(32) << 32) | hb.read(32)); }      // CIS:1247-1253
    rc.payBit = (i64)hb.pos; rc.payBits = written - (i64)(headerSize + chkBytes) * 8; rc.preLen = (i32)pre;
    rc.rawCopy = copyBlock && !transformedCopy;
    rc.skipFlags = rc.rawCopy ? 0xFF : (skipFlags & 0xFF);
    rc.entropy = copyBlock ? KZG_E_NONE : entropy;
    recs.push_back(rc);
    hb.pos = blockStart + (u64)written;
    if (hb.pos > hb.nbits) { kzg_set_error("truncated stream"); return -KZG_ERR_READ_FILE; }
  }
  const int nBlocks = (int)recs.size();
  if (nBlocks == 0) return 0;
  // every block but the last decodes to blockSize bytes; the last needs at least one byte of room (its kernels are clamped to
  // what is left: KzgBlock.finalCap, kzg_dst_limit)
  if ((i64)(nBlocks - 1) * blockSize >= outCap) { kzg_set_error("output capacity too small"); return -KZG_ERR_WRITE_FILE; }

  i32 maxPre = 0; for (auto& rc : recs) maxPre = std::max(maxPre, rc.preLen);
  const i32 blkBuf = std::max(blockSize + 512, blockSize + (blockSize >> 4));      // CIS:694-695
  const i32 maxLen = std::max(std::max(blockSize, maxPre + 512), blkBuf);
  const size_t cap = rnd((size_t)maxLen + 64);
  XfScratch xs = xf_scratch_size(fn, nf, maxLen, false);
  EntScratch es = ent_scratch_size(entropy, maxPre, false);
  // as in compress_impl: blocks that do not all fit the arena run as several slices of whole blocks, one after the other
  const size_t perBlock = (sizeof(KzgBlock) + 64) + cap * 2 + xs.perBlock + (xs.hashInts + xs.aux32) * 4 +
                          (size_t)es.maxChunks * (sizeof(KzgChunkInfo) + es.tabStride * 4) + nf + 16;
  const size_t budget = arena_budget((size_t)nBlocks * perBlock + (1 << 20));
  const int sliceBlocks = (int)std::max<size_t>(1, std::min<size_t>((size_t)nBlocks, (budget > (2u << 20) ? budget - (2u << 20) : 0) / (perBlock + perBlock / 8)));
  r = ws_reserve((size_t)std::min(nBlocks, sliceBlocks) * perBlock + (1 << 20), (size_t)std::min(nBlocks, sliceBlocks) * (sizeof(KzgBlock) + 16 + nf) + 4096); if (r < 0) return r;
  const int gDec = getenv("KZG_DEC_GROUPS") ? std::max(1, std::min(KZG_DEC_MAXG, atoi(getenv("KZG_DEC_GROUPS")))) : 12;   // developer knob (read per call)
  EvSet ev;
  if (timing3) { r = ev.create(3); if (r < 0) return r; timing3[0] = timing3[1] = timing3[2] = 0; }
  cudaStream_t const mainStream = W.stream;
  i64 total = 0;
  for (int s0 = 0; s0 < nBlocks; s0 += sliceBlocks) {
    const int sN = std::min(sliceBlocks, nBlocks - s0);
    const size_t nb = (size_t)sN;
    W.dOff = 0; W.hOff = 0;
    Batch bt; bt.nBlocks = sN; bt.maxLen = maxLen;
    bt.hBlocks = halloc<KzgBlock>(nb); NN(bt.hBlocks);
    bt.hEnabled = halloc<u8>(nb * nf + 16); NN(bt.hEnabled);
    bt.hDstLimit = halloc<int>(nb); NN(bt.hDstLimit);
    bt.dBlocks = dalloc<KzgBlock>(nb); NN(bt.dBlocks);
    bt.dResult = dalloc<int>(2 * nb); NN(bt.dResult);
    u8* dEnabledAll = dalloc<u8>(nb * nf + 16); NN(dEnabledAll);
    bt.dDstLimit = dalloc<int>(nb); NN(bt.dDstLimit);
    u8* dE = dalloc<u8>(nb * cap); NN(dE);
    u8* dF = dalloc<u8>(nb * cap); NN(dF);
    u8* dScratch = dalloc<u8>(nb * xs.perBlock + 16); NN(dScratch);
    i32* dHash = dalloc<i32>(nb * xs.hashInts + 4); NN(dHash);
    i32* dAux = dalloc<i32>(nb * xs.aux32 + 4); NN(dAux);
    KzgChunkInfo* dChunks = dalloc<KzgChunkInfo>(nb * es.maxChunks + 1); NN(dChunks);
    u32* dTab = dalloc<u32>(nb * es.maxChunks * es.tabStride + 4); NN(dTab);
    bool anyNone = false, anyEnt = false;
    for (int b = 0; b < sN; b++) {
      const int gb = s0 + b;                       // block index in the stream
      const Rec& rc = recs[gb];
      KzgBlock& B = bt.hBlocks[b];
      memset(&B, 0, sizeof(B));
      int k = 0;     // inverse stages this block runs (Sequence.inverse :160-166 skips flagged transforms)
      for (int i = 0; i < nf; i++) {
        const bool runs = (rc.skipFlags != 0xFF) && ((rc.skipFlags & (1 << (7 - i))) == 0) && fn[i] != KZG_T_NONE;
        bt.hEnabled[(size_t)i * nb + b] = runs ? 1 : 0;
        if (runs) k++;
      }
      u8* dest = d_out + (size_t)gb * blockSize;
      const i32 avail = (i32)std::min<i64>(blockSize, outCap - (i64)gb * blockSize);    // bytes of the caller's buffer this block owns
      B.aux0 = dest; B.stagesLeft = k; B.finalCap = avail;
      B.cur = (k == 0) ? dest : dE + (size_t)b * cap;
      B.alt = (k == 1) ? dest : dF + (size_t)b * cap;
      B.aux1 = dF + (size_t)b * cap;
      B.curLen = rc.preLen; B.preLen = rc.preLen; B.cap = (i32)cap - 64; B.skipFlags = rc.skipFlags;
      B.entropy = rc.entropy; B.srcBit = rc.payBit; B.srcBits = rc.payBits; B.xxh = rc.xxh; B.chkBytes = chkBytes;
      if (k == 0 && rc.preLen > avail) { kzg_set_error(rc.preLen > blockSize ? "Block %d incorrectly decompressed" : "output capacity too small for block %d", gb + 1); return rc.preLen > blockSize ? -KZG_ERR_PROCESS_BLOCK : -KZG_ERR_WRITE_FILE; }
      bt.hDstLimit[b] = blkBuf;
      if (rc.entropy == KZG_E_NONE) anyNone = true; else anyEnt = true;
    }
    CUDA_TRY(cudaMemcpyAsync(dEnabledAll, bt.hEnabled, nb * nf, cudaMemcpyHostToDevice, W.stream));
    CUDA_TRY(cudaMemcpyAsync(bt.dDstLimit, bt.hDstLimit, nb * sizeof(int), cudaMemcpyHostToDevice, W.stream));
    r = batch_upload(bt); if (r < 0) return r;
    // Blocks are independent: contiguous groups of them run the whole decode on streams of their own (most urgent first), so
    // that the latency-bound kernels of one group (chunk scan, the literal-record chain) overlap the streaming kernels of the
    // others and, for host buffers, a group's upload / download overlaps the other groups' kernels.
    {
      int nSM = 0;
      if (cudaDeviceGetAttribute(&nSM, cudaDevAttrMultiProcessorCount, W.device) != cudaSuccess) { cudaGetLastError(); nSM = 0; }
      // (more chains than SMs: exclusive SMs would only make them queue, 4 % slower at 400 blocks.  With host buffers the groups are
      //  staggered by their uploads: a chain CTA that needs a whole SM then waits for one to drain of the other groups' entropy-decode
      //  CTAs, and the decode as a whole was 4 % slower.)
      W.lziExclusive = sN <= nSM && (!copyIn || getenv("KZG_LZI_EXCL_E2E") != nullptr);
    }
    const int G = (sN >= 8 && (i64)sN * blockSize >= (16 << 20)) ? std::min(gDec, sN / 2) : 1;      // (small batches: one launch chain, the groups would only add launches)
    if (G > 1) { r = ws_side_init(); if (r < 0) return r; }
    if (timing3) CUDA_TRY(cudaEventRecord(ev[0], W.stream));
    if (G > 1) CUDA_TRY(cudaEventRecord(W.sideEv[KZG_DEC_MAXG], W.stream));
    int rc = 0;
    // Two passes over the groups.  The BWT launcher reads block headers back (a host wait for everything its stream holds): run in
    // the first pass it would stop the host from enqueueing the next group until this group's earlier stages are done, and the
    // groups would run one after the other.  Pass 0 enqueues every group up to its first such stage; pass 1 the rest, by which time
    // all groups are running.
    int resume[KZG_DEC_MAXG];                     // per group: the stage pass 1 resumes at (-1: the group was finished in pass 0)
    for (int pass = 0; pass < 2 && rc == 0; pass++) {
    for (int g = 0; g < G && rc == 0; g++) {
      if (pass == 1 && resume[g] < 0) continue;
      const int b0 = (int)((i64)sN * g / G), b1 = (int)((i64)sN * (g + 1) / G), cnt = b1 - b0;
      cudaStream_t q = (G > 1) ? W.side[g] : mainStream;
      if (pass == 0 && G > 1) CUDA_TRY(cudaStreamWaitEvent(q, W.sideEv[KZG_DEC_MAXG], 0));
      if (pass == 0 && copyIn) {                  // the bytes that hold this group's block records (+ the slack the bit readers touch)
        const i64 lo = (recs[s0 + b0].payBit >> 3) & ~(i64)63;
        const i64 hi = std::min<i64>(nBytes, ((recs[s0 + b1 - 1].payBit + recs[s0 + b1 - 1].payBits + 7) >> 3) + 128);
        if (hi > lo) CUDA_TRY(cudaMemcpyAsync((u8*)d_in + lo, h_in + lo, (size_t)(hi - lo), cudaMemcpyHostToDevice, q));
      }
      Batch sub = bt;
      sub.nBlocks = cnt; sub.hBlocks = bt.hBlocks + b0; sub.dBlocks = bt.dBlocks + b0; sub.dResult = bt.dResult + 2 * b0; sub.dDstLimit = bt.dDstLimit + b0;
      W.stream = q;                               // (the launch helpers enqueue on the calling thread's current stream)
      do {
        int first = nf - 1;
        if (pass == 0) {
          if (anyEnt) { rc = run_entropy_decode(sub, entropy, es, d_in, dChunks + (size_t)b0 * es.maxChunks, dTab + (size_t)b0 * es.maxChunks * es.tabStride); if (rc < 0) break; }
          if (anyNone) { rc = run_entropy_decode(sub, KZG_E_NONE, es, d_in, dChunks + (size_t)b0 * es.maxChunks, dTab + (size_t)b0 * es.maxChunks * es.tabStride); if (rc < 0) break; }
          if (timing3 && g == 0) { if (cudaEventRecord(ev[1], q) != cudaSuccess) { rc = -KZG_ERR_PROCESS_BLOCK; break; } }
          resume[g] = -1;
        } else first = resume[g];
        for (int i = first; i >= 0; i--) {
          if (fn[i] == KZG_T_NONE) continue;
          // a stage no block of this group runs (skip flags; BWT under the reference's bounds) is not launched at all
          bool anyOn = false;
          for (int b = b0; b < b1 && !anyOn; b++) anyOn = bt.hEnabled[(size_t)i * nb + b] != 0;
          if (!anyOn) continue;
          if (pass == 0 && fn[i] == KZG_T_BWT && G > 1) { resume[g] = i; break; }
          sub.dEnabled = dEnabledAll + (size_t)i * nb + b0;
          rc = run_transform_stage(sub, fn[i], i, false, xs, dScratch + (size_t)b0 * xs.perBlock, dHash + (size_t)b0 * xs.hashInts, dAux + (size_t)b0 * xs.aux32, flags);
          if (rc < 0) break;
        }
        // (a group with nothing left for pass 1 finishes in pass 0: its download overlaps the other groups' kernels from the start)
        if (rc >= 0 && (pass == 1 || resume[g] < 0) && chkBytes) rc = kzg_xxh_launch(q, sub.dBlocks, cnt, 1);       // verify the decoded bytes (CIS:1348-1370)
      } while (0);
      W.stream = mainStream;
      if (rc < 0) break;
      if (pass == 0 && resume[g] >= 0) continue;  // the rest of this group in pass 1
      if (h_out) {                                // every block but the stream's last is blockSize bytes; the last one follows below
        const int full = (s0 + b1 == nBlocks) ? cnt - 1 : cnt;
        if (full > 0) CUDA_TRY(cudaMemcpyAsync(h_out + (size_t)(s0 + b0) * blockSize, d_out + (size_t)(s0 + b0) * blockSize, (size_t)full * blockSize, cudaMemcpyDeviceToHost, q));
      }
      if (G > 1) { CUDA_TRY(cudaEventRecord(W.sideEv[g], q)); CUDA_TRY(cudaStreamWaitEvent(mainStream, W.sideEv[g], 0)); }
    }
    }
    if (rc < 0) { W.stream = mainStream; for (int g = 0; g < G && G > 1; g++) cudaStreamSynchronize(W.side[g]); return rc; }
    if (timing3) CUDA_TRY(cudaEventRecord(ev[2], W.stream));
    r = batch_download(bt); if (r < 0) return r;
    if (timing3) {      // [1] = entropy stage of the first group, [0] = everything else up to the join (groups overlap: a split, not a sum of kernels)
      float tot = 0, ent = 0;
      CUDA_TRY(cudaEventElapsedTime(&ent, ev[0], ev[1])); CUDA_TRY(cudaEventElapsedTime(&tot, ev[0], ev[2])); timing3[1] += ent; timing3[0] += tot - ent;
    }
    for (int b = 0; b < sN; b++) {
      const KzgBlock& B = bt.hBlocks[b];
      const int gb = s0 + b;
      if (B.status != 0) { kzg_set_error("block %d failed with status %d", gb + 1, B.status); return B.status; }
      if (B.curLen > blockSize) { kzg_set_error("Block %d incorrectly decompressed", gb + 1); return -KZG_ERR_PROCESS_BLOCK; }
      if (gb + 1 < nBlocks && B.curLen != blockSize) { kzg_set_error("short block %d (%d bytes) inside the stream", gb + 1, B.curLen); return -KZG_ERR_PROCESS_BLOCK; }
      if (B.cur != d_out + (size_t)gb * blockSize) { kzg_set_error("internal: block %d did not land in place", gb + 1); return -KZG_ERR_UNKNOWN; }
      total += B.curLen;
      if (gb + 1 == nBlocks && h_out && B.curLen > 0 &&
          cudaMemcpy(h_out + (size_t)gb * blockSize, d_out + (size_t)gb * blockSize, (size_t)B.curLen, cudaMemcpyDeviceToHost) != cudaSuccess) return -KZG_ERR_PROCESS_BLOCK;
    }
  }
  if (total > outCap) return -KZG_ERR_WRITE_FILE;
  return total;
}

int64_t kzg_decompress_dev(const uint8_t* d_in, int64_t nBytes, const uint8_t* h_in, int32_t flags, uint8_t* d_out, int64_t outCap,
                           float* timing3) {
  return decompress_impl(d_in, nBytes, h_in, flags, d_out, outCap, timing3, false, nullptr);
}

// ---- host-buffer forms: H2D, the device path, D2H ---------------------------------------------------------------------
int64_t kzg_compress(const uint8_t* in, int64_t n, const int32_t* transforms, int32_t nTransforms, int32_t entropy, int32_t blockSize,
                     int32_t flags, uint8_t* out, int64_t outCap) {
  int r = ws_init(); if (r < 0) return r;
  if (n < 0 || (n > 0 && in == nullptr) || out == nullptr) return -KZG_ERR_INVALID_PARAM;
  const i64 bound = std::min<i64>(outCap, kzg_compress_bound(n, blockSize)) + 16;
  const size_t inCap = rnd((size_t)n + 64);
  u8* dIn = ws_io(0, inCap); u8* dOut = ws_io(1, rnd((size_t)bound + 64));
  if (!dIn || !dOut) { kzg_set_error("cudaMalloc failed for stream buffers"); return -KZG_ERR_CREATE_CODEC; }
  i64 res;
  do {
    if (cudaMemsetAsync(dIn + n, 0, inCap - n, W.stream) != cudaSuccess) { res = -KZG_ERR_PROCESS_BLOCK; break; }
    res = compress_impl(dIn, n, transforms, nTransforms, entropy, blockSize, flags, dOut, bound, nullptr, in);   // (uploads the input itself)
    if (res < 0) break;
    if (res > outCap) { res = -KZG_ERR_WRITE_FILE; break; }
    if (cudaMemcpy(out, dOut, (size_t)res, cudaMemcpyDeviceToHost) != cudaSuccess) { res = -KZG_ERR_PROCESS_BLOCK; break; }
  } while (0);
  return res;
}

int64_t kzg_decompress(const uint8_t* in, int64_t nBytes, int32_t flags, uint8_t* out, int64_t outCap) {
  int r = ws_init(); if (r < 0) return r;
  if (in == nullptr || out == nullptr || nBytes < 0 || outCap < 0) return -KZG_ERR_INVALID_PARAM;
  const size_t inCap = rnd((size_t)nBytes + 64);
  // room for whole blocks: the last block may be short but kernels address by block
  u8* dIn = ws_io(0, inCap); u8* dOut = ws_io(1, rnd((size_t)outCap + 64));
  if (!dIn || !dOut) { kzg_set_error("cudaMalloc failed for stream buffers"); return -KZG_ERR_CREATE_CODEC; }
  i64 res;
  do {
    if (cudaMemsetAsync(dIn + nBytes, 0, inCap - nBytes, W.stream) != cudaSuccess) { res = -KZG_ERR_PROCESS_BLOCK; break; }
    // every block group uploads its own byte range of the stream and downloads its blocks (decompress_impl)
    res = decompress_impl(dIn, nBytes, in, flags, dOut, outCap, nullptr, true, out);
    if (res < 0) break;
  } while (0);
  return res;
}

}  // extern "C"
