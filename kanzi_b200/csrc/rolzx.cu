// rolzx.cu — ROLZX = ROLZCodec2 (K/transform/ROLZCodec.java:1016-1428; TransformFactory.ROLZX_TYPE = 12) as a transform stage.
// One warp per block: the lanes histogram the block for the data type (ROLZCodec.java:1203-1221), fill the probability cells,
// clear the counters and, before every 16 MiB chunk, the 8 MiB ring table; lane 0 runs the chain (rolzx_core.cuh says why it is a
// chain and what is done about its cost).  Blocks of any size up to 1 GiB: the chunks share the coder and the counters.
#include "kzg_transforms.cuh"
#include "kzg_xf_kernels.cuh"
#include "rolzx_core.cuh"

#define RZX_PROB_BYTES ((size_t)(RZX_LIT_CELLS + RZX_MATCH_CELLS) * 2)
#define RZX_HASH_INTS (RZX_MATCH_INTS + (size_t)RZX_HASH_SIZE)

// Global.detectSimpleType (K/Global.java:556-608): only DNA and "nothing special" matter to ROLZX, but ctx["dataType"] takes whatever it finds
__device__ int rzx_detect_type(int count, const u32* f) {
  if (count == 0) return KZG_DT_UNDEFINED;
  int sum = f['a'] + f['c'] + f['g'] + f['n'] + f['t'] + f['u'] + f['A'] + f['C'] + f['G'] + f['N'] + f['T'] + f['U'];
  if (sum > count - count / 12) return KZG_DT_DNA;
  sum = f['+'] + f['-'] + f['*'] + f['/'] + f['='] + f[','] + f['.'] + f[':'] + f[';'] + f[' '];
  for (int c = '0'; c <= '9'; c++) sum += f[c];
  if (sum == count) return KZG_DT_NUMERIC;
  sum = (f[0x3D] == 1) ? 1 : 0;
  for (int c = 'A'; c <= 'Z'; c++) sum += f[c];
  for (int c = 'a'; c <= 'z'; c++) sum += f[c];
  for (int c = '0'; c <= '9'; c++) sum += f[c];
  sum += f['+'] + f['/'];
  if (sum == count) return KZG_DT_BASE64;
  sum = 0;
  for (int i = 0; i < 256; i++) sum += (f[i] > 0) ? 1 : 0;
  if (sum == 256) return KZG_DT_BIN;
  if (sum <= 4) return KZG_DT_SMALL_ALPHABET;
  return KZG_DT_UNDEFINED;
}

template <bool FORWARD>
__global__ void __launch_bounds__(32) rolzx_kernel(KzgBlock* __restrict__ blocks, KzgXfParams P) {
  __shared__ u32 freqs[256];
  __shared__ int sType;
  const int lane = threadIdx.x, b = blockIdx.x;
  KzgBlock& B = blocks[b];
  int* res = P.result + 2 * b;
  if (lane == 0) { res[0] = 0; res[1] = 0; }
  if (B.status != 0 || !P.enabled[b]) return;
  const int count = B.curLen;
  const u8* src = B.cur;
  u8* dst = B.alt;
  u8* sc = P.scratch + (size_t)b * (size_t)P.scratchStride;
  sc = (u8*)(((uintptr_t)sc + 15) & ~(uintptr_t)15);
  uint16_t* lit = (uint16_t*)sc;
  uint16_t* mat = lit + RZX_LIT_CELLS;
  i32* matches = P.hashBuf + (size_t)b * RZX_HASH_INTS;
  i32* counters = matches + RZX_MATCH_INTS;
  int mm = 3, dt = 2, flags = 0;
  int total = 0, dstEnd = 0, dstCap = 0, limit = 0;      // forward: total = srcEnd; inverse: total = dstEnd = szBlock
  if (FORWARD) {
    const int maxEnc = (count <= 16384) ? count + 1024 : count + (count / 32);              // :1417-1421
    if (count < 64 || count > (1 << 30) || P.dstLimit[b] < maxEnc) return;                  // ROLZCodec.java:216-221, 1184-1185
    limit = min(P.dstLimit[b], B.cap);
    if (limit < 16) return;
    int dtp = B.dataType;
    if (dtp == KZG_DT_UNDEFINED) {
      for (int i = lane; i < 256; i += 32) freqs[i] = 0;
      __syncwarp();
      for (int i = lane; i < count; i += 32) atomicAdd(&freqs[src[i]], 1u);
      __syncwarp();
      if (lane == 0) { const int d = rzx_detect_type(count, freqs); if (d != KZG_DT_UNDEFINED) B.dataType = d; sType = d; }
      __syncwarp();
      dtp = sType;
    }
    if (dtp == KZG_DT_EXE) { dt = 3; flags |= 8; }
    else if (dtp == KZG_DT_DNA) { dt = 8; mm = 7; flags |= 4; }
    total = count - 4;
  } else {
    if (count > (1 << 30) || count < 13) return;                                            // (5 header bytes + the coder's 8: anything shorter makes the reference throw)
    const int szBlock = (int)(((u32)src[0] << 24) | ((u32)src[1] << 16) | ((u32)src[2] << 8) | (u32)src[3]);
    dstCap = min(kzg_dst_limit(B, P.dstLimit[b]), B.cap);
    if (szBlock <= 0 || szBlock > P.dstLimit[b] || szBlock > dstCap) return;                // :1306-1307 (output.length = dst.array.length in a Sequence)
    flags = src[4];
    if ((flags & 0x0E) == 8) dt = 3;                                                         // bsVersion >= 4 on this path (:1317-1327)
    else if ((flags & 0x0E) == 4) { dt = 8; mm = 7; }
    total = dstEnd = szBlock;
  }
  for (int i = lane; i < (RZX_LIT_CELLS + RZX_MATCH_CELLS) / 2; i += 32) ((u32*)lit)[i] = 0x7FFF7FFFu;
  for (int i = lane; i < RZX_HASH_SIZE; i += 32) counters[i] = 0;
  RzxCoder C;
  if (FORWARD) {
    if (lane == 0) {
      dst[0] = (u8)(count >> 24); dst[1] = (u8)(count >> 16); dst[2] = (u8)(count >> 8); dst[3] = (u8)count;
      dst[4] = (u8)flags;
    }
    rzx_coder_init(C, lit, mat, dst, 5, limit);
  } else {
    rzx_coder_init(C, lit, mat, const_cast<u8*>(src), 5, count);
    if (lane == 0) rzx_decoder_start(C);
  }
  const int sizeChunk = min(FORWARD ? count : total, RZX_CHUNK);
  int startChunk = 0, outIndex = 0;
  int okSoFar = 1;
  while (startChunk < total) {
    int4* m4 = (int4*)matches;
    for (size_t i = lane; i < RZX_MATCH_INTS / 4; i += 32) m4[i] = make_int4(0, 0, 0, 0);
    __syncwarp();
    const int endChunk = min(startChunk + sizeChunk, total);
    if (lane == 0 && okSoFar) {
      if (FORWARD) rzx_forward_chunk(src, startChunk, endChunk, total, C, matches, counters, mm, dt);
      else okSoFar = rzx_inverse_chunk(dst, startChunk, endChunk, dstEnd, dstCap, &outIndex, C, matches, counters, mm, dt) ? 1 : 0;
    }
    __syncwarp();
    okSoFar = __shfl_sync(0xFFFFFFFFu, okSoFar, 0);
    if (!okSoFar) break;
    startChunk = endChunk;
  }
  if (lane != 0) return;
  if (FORWARD) {
    rzx_forward_tail(src, total, C);
    if (C.overrun) { B.status = -KZG_ERR_PROCESS_BLOCK; return; }                           // the reference writes past its array here: an exception, "Error in block"
    res[0] = 1; res[1] = C.index;                                                           // (:1291-1293: the size clause compares an index with itself)
  } else {
    const bool ok = okSoFar && !C.overrun && C.index == count;                              // :1411-1413
    res[0] = ok ? 1 : 0;
    res[1] = ok ? outIndex : 0;
  }
}

void kzg_rolzx_scratch(i32 maxLen, bool forward, size_t* perBlockBytes, size_t* hashInts) {
  (void)maxLen; (void)forward;
  *perBlockBytes = std::max(*perBlockBytes, RZX_PROB_BYTES + 256);
  *hashInts = std::max(*hashInts, RZX_HASH_INTS);
}

int kzg_rolzx_launch(cudaStream_t s, bool forward, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, i32 maxLen) {
  (void)maxLen;
  if ((size_t)P.scratchStride < RZX_PROB_BYTES + 16) { kzg_set_error("rolzx: scratch pool too small"); return -KZG_ERR_CREATE_CODEC; }
  if (forward) KZG_PROF("rolzx_forward_kernel", s, (rolzx_kernel<true><<<nBlocks, 32, 0, s>>>(d_blocks, P)));
  else KZG_PROF("rolzx_inverse_kernel", s, (rolzx_kernel<false><<<nBlocks, 32, 0, s>>>(d_blocks, P)));
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(1);
  return 0;
}
