// stages.cu — stage dispatch, slice prechecks and the small helper kernels around the codecs (sm_100a).
#include "kzg_common.cuh"
#include "kzg_transforms.cuh"
#include "kzg_stages.cuh"
#include "kzg_xf_kernels.cuh"

// ---- NullEntropyDecoder (K/entropy/NullEntropyDecoder.java:44-58): copy preLen bytes from a bit offset -------------
__global__ void __launch_bounds__(256) kzg_rawbits_kernel(KzgBlock* __restrict__ blocks, const u8* __restrict__ stream) {
  KzgBlock& B = blocks[blockIdx.y];
  if (B.status != 0 || B.entropy != KZG_E_NONE) return;
  const int n = B.preLen;
  if ((i64)n * 8 > B.srcBits) { if (threadIdx.x == 0 && blockIdx.x == 0) B.status = -KZG_ERR_PROCESS_BLOCK; return; }
  u8* __restrict__ out = B.cur;
  const u64 base = (u64)B.srcBit;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = (u8)get_bits(stream, base + 8ull * i, 8);
  if (threadIdx.x == 0 && blockIdx.x == 0) B.entBits = (i64)n * 8;
}
int kzg_rawbits_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, const u8* d_stream) {
  kzg_rawbits_kernel<<<dim3(32, nBlocks), 256, 0, s>>>(d_blocks, d_stream);
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(1);
  return 0;
}

// ---- Magic.getType + the dataType rule of EncodingTask (K/Magic.java:154-186, COS:795-804) ---------------------------
__device__ int kzg_magic_datatype(const u8* p) {
  const i32 key = (i32)(((u32)p[0] << 24) | ((u32)p[1] << 16) | ((u32)p[2] << 8) | (u32)p[3]);
  i32 m = 0;
  const i32 k32[] = {0x47494638, 0x25504446, 0x504B0304, 0x377ABCAF, (i32)0x89504E47, 0x7F454C46, (i32)0xFEEDFACE, (i32)0xCEFAEDFE,
                     (i32)0xFEEDFACF, (i32)0xCFFAEDFE, 0x28B52FFD, (i32)0x81CFB2CE, 0x4D534346, 0x52494646, 0x664C6143, (i32)0xFD377A58,
                     0x4B414E5A, 0x52617221};
  if ((key & ~0x0F) == (i32)0xFFD8FFE0) m = key;
  else if (((key >> 8) == 0x425A68) || ((key >> 8) == 0x494433)) m = key >> 8;
  else {
    for (int i = 0; i < 18 && m == 0; i++) if (key == k32[i]) m = key;
    if (m == 0) {
      const i32 key16 = key >> 16;
      if (key16 == 0x1F8B || key16 == 0x424D || key16 == 0x4D5A) m = key16;
      else if (key16 == 0x5034 || key16 == 0x5035 || key16 == 0x5036) {
        const int sub = (key >> 8) & 0xFF;
        if (sub == 0x07 || sub == 0x0A || sub == 0x0D || sub == 0x20) m = key16;
      }
    }
  }
  switch (m) {     // isCompressed -> BIN, else isMultimedia -> MULTIMEDIA, else isExecutable -> EXE
    case (i32)0xFFD8FFE0: case 0x47494638: case (i32)0x89504E47: case 0x377ABCAF: case 0x28B52FFD: case (i32)0x81CFB2CE: case 0x4D534346:
    case 0x504B0304: case 0x1F8B: case 0x425A68: case 0x664C6143: case 0x494433: case (i32)0xFD377A58: case 0x4B414E5A: case 0x52617221:
      return KZG_DT_BIN;
    case 0x52494646: case 0x424D: case 0x5034: case 0x5035: case 0x5036: return KZG_DT_MULTIMEDIA;
    case 0x7F454C46: case 0x4D5A: case (i32)0xFEEDFACE: case (i32)0xCEFAEDFE: case (i32)0xFEEDFACF: case (i32)0xCFFAEDFE: return KZG_DT_EXE;
    default: return KZG_DT_UNDEFINED;
  }
}
__global__ void kzg_magic_kernel(KzgBlock* __restrict__ blocks, int nBlocks) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nBlocks) return;
  KzgBlock& B = blocks[b];
  if (B.origLen >= 4) { const int dt = kzg_magic_datatype(B.cur); if (dt != KZG_DT_UNDEFINED) B.dataType = dt; }
}
int kzg_magic_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks) {
  kzg_magic_kernel<<<(nBlocks + 127) / 128, 128, 0, s>>>(d_blocks, nBlocks);
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(1);
  return 0;
}

// ---- NullTransform.forward inside a Sequence (NullTransform.java:41-66): data stays where it is, the stage "succeeds" ----
__global__ void kzg_null_forward_kernel(KzgBlock* __restrict__ blocks, int nBlocks, const u8* __restrict__ enabled, int stage) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nBlocks) return;
  if (blocks[b].status == 0 && (enabled[b] || blocks[b].origLen <= 15)) blocks[b].skipFlags &= ~(1 << (7 - stage));
}
int kzg_null_forward_launch(cudaStream_t s, KzgBlock* d_blocks, int nBlocks, const u8* enabled, int stage) {
  kzg_null_forward_kernel<<<(nBlocks + 127) / 128, 128, 0, s>>>(d_blocks, nBlocks, enabled, stage);
  CUDA_TRY(cudaGetLastError());
  kzg_count_launch(1);
  return 0;
}

// ---- dispatch ------------------------------------------------------------------------------------------------------
void kzg_stage_scratch(int type, i32 maxLen, bool forward, size_t* perBlockBytes, size_t* hashInts, size_t* aux32) {
  switch (type) {
    case KZG_T_BWT: kzg_bwt_scratch(maxLen, forward, perBlockBytes, aux32); break;
    case KZG_T_ROLZ: kzg_rolz_scratch(maxLen, forward, perBlockBytes, hashInts, aux32); break;
    case KZG_T_LZP: kzg_lzp_scratch(maxLen, forward, perBlockBytes, hashInts); break;
    case KZG_T_ROLZX: kzg_rolzx_scratch(maxLen, forward, perBlockBytes, hashInts); break;
    case KZG_T_SRT: case KZG_T_RANK: case KZG_T_MTFT: case KZG_T_ZRLT: kzg_small_scratch(type, maxLen, forward, perBlockBytes, aux32); break;
    default: break;
  }
}

int kzg_stage_launch(cudaStream_t s, int type, bool forward, KzgBlock* d_blocks, int nBlocks, const KzgXfParams& P, i32 maxLen) {
  switch (type) {
    case KZG_T_ZRLT: return kzg_zrlt_launch(s, forward, d_blocks, nBlocks, P, maxLen);
    case KZG_T_RANK: return kzg_sbrt_launch(s, forward, 2, d_blocks, nBlocks, P, maxLen);
    case KZG_T_MTFT: return kzg_sbrt_launch(s, forward, 1, d_blocks, nBlocks, P, maxLen);
    case KZG_T_SRT: return kzg_srt_launch(s, forward, d_blocks, nBlocks, P, maxLen);
    case KZG_T_BWT: return kzg_bwtblock_launch(s, forward, d_blocks, nBlocks, P, maxLen);
    case KZG_T_ROLZ: return kzg_rolz_launch(s, forward, d_blocks, nBlocks, P, maxLen);
    case KZG_T_LZP: return kzg_lzp_launch(s, forward, d_blocks, nBlocks, P, maxLen);
    case KZG_T_RLT: return kzg_rlt_launch(s, forward, d_blocks, nBlocks, P, maxLen);
    case KZG_T_ROLZX: return kzg_rolzx_launch(s, forward, d_blocks, nBlocks, P, maxLen);
    default: kzg_set_error("transform id %d has no kernel", type); return -KZG_ERR_INVALID_CODEC;
  }
}

// The guard blocks of the codecs, evaluated for slices with index 0 (what the C ABI passes).
// src.array.length is taken as srcLen (the shim passes exactly the slice), dst.array.length = dstCap.
int kzg_stage_precheck(int type, bool forward, const kzg_ctx* ctx, i32 srcLen, i32 dstLen, i32 dstCap) {
  const bool asref = ctx && (ctx->flags & KZG_FLAG_BWT_ASREF);
  switch (type) {
    case KZG_T_NONE: return (dstLen < srcLen) ? 0 : 1;                                   // NullTransform.java:57-58
    case KZG_T_LZ: case KZG_T_LZX:
      if (forward) { if (dstLen < ((srcLen <= 1024 ? srcLen + 16 : srcLen + srcLen / 64) + 2)) return 0; }   // LZCodec.java:309-310
      return 1;
    case KZG_T_ROLZX:
      if (srcLen > (1 << 30)) return 0;                                                                                        // ROLZCodec.java:220-221, 252-253
      if (forward && (srcLen < 64 || dstLen < ((srcLen <= 16384) ? srcLen + 1024 : srcLen + srcLen / 32))) return 0;           // :216-217, 1184-1185
      return 1;
    case KZG_T_RLT:
      if (forward && (srcLen < 16 || dstLen < ((srcLen <= 512) ? srcLen + 32 : srcLen))) return 0;                          // RLT.java:71-80
      return 1;
    case KZG_T_LZP:
      if (forward) { if (dstLen < ((srcLen <= 1024) ? srcLen + 16 : srcLen + srcLen / 64) || srcLen < 128) return 0; }     // LZCodec.java:1013-1018
      else if (dstLen < srcLen) return 0;                                                                                    // :1140-1141
      return 1;
    case KZG_T_ROLZ:
      if (forward) { if (srcLen < 64 || srcLen > (1 << 30)) return 0; if (dstLen < ((srcLen <= 512) ? srcLen + 64 : srcLen)) return 0; }   // ROLZCodec.java:207-212,426-428
      else if (srcLen > (1 << 30)) return 0;
      return 1;
    case KZG_T_ZRLT:
      if (forward && dstLen < srcLen) return 0;                                          // ZRLT.java:65-66
      return 1;
    case KZG_T_RANK: case KZG_T_MTFT:
      if (dstLen > dstCap || srcLen > dstLen || srcLen > dstCap) return 0;               // SBRT.java:90-106
      return 1;
    case KZG_T_SRT:
      if (forward && dstLen < srcLen + 1024) return 0;                                   // SRT.java:83-84
      return 1;
    case KZG_T_BWT:
      if (forward) {                                                                     // BWTBlockCodec.java:82-96
        if (dstLen > dstCap || dstLen < srcLen + 33) return 0;
        if (asref) {             // BWT.java:152-156 as written: dst.index (= header size) + dst.length > dst.array.length
          int lg = 0; while ((2 << lg) <= srcLen) lg++;
          if ((srcLen & (srcLen - 1)) != 0) lg++;
          const int hdr = 1 + ((srcLen < 256) ? 1 : 8) * ((lg + 7) >> 3);
          if (hdr + dstLen > dstCap) return 0;
        }
      } else {
        if (asref) return 0;     // BWT.java:211 as written: count > src.length - src.index once the header is consumed
      }
      return 1;
    default: return -KZG_ERR_INVALID_CODEC;
  }
}
