/*
 * kzg_jni.c — JNI glue between reference Kanzi's plugin interfaces and libkanzi_b200.so (include/kzg.h).
 *
 * The five natives of integration/java/io/github/flanglet/kanzi/gpu/{GpuTransform,GpuEntropyEncoder,GpuEntropyDecoder}.java:
 *   GpuTransform.forward0 / inverse0 / maxLen0     -> ByteTransform.forward / inverse / getMaxEncodedLength
 *                                                     (K/ByteTransform.java:24-57, called from K/transform/Sequence.java:56-207)
 *   GpuEntropyEncoder.encode0                      -> EntropyEncoder.encode + dispose (K/EntropyEncoder.java:23-49, COS:907-916)
 *   GpuEntropyDecoder.decode0                      -> EntropyDecoder.decode (K/EntropyDecoder.java:23-47, CIS:1305-1316)
 * Compiled only where a JDK exists (this image has none): cc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux
 *   -Iinclude jni/kzg_jni.c -Lkanzi_b200 -lkanzi_b200 -o libkanzi_b200_jni.so.  Without <jni.h> the file compiles to an empty
 * object (the __has_include guard), so `make -C jni check` in a JDK-less tree still parses the non-JNI part.
 * Nothing here throws across JNI: every native returns Kanzi's own conventions (1/0 = true/false, < 0 = -Error.ERR_*).
 */
#include <stdint.h>
#include "../include/kzg.h"

#if defined(__has_include)
#if __has_include(<jni.h>)
#define KZG_HAVE_JNI 1
#endif
#endif

#ifdef KZG_HAVE_JNI
#include <jni.h>

/* ctx int[6] = { bsVersion, blockSize, size, jobs, dataType ordinal (in/out), flags } — the Map<String,Object> fields the hot
 * path reads or writes */
static void load_ctx(JNIEnv* env, jintArray jctx, kzg_ctx* ctx, jint* c) {
  (*env)->GetIntArrayRegion(env, jctx, 0, 6, c);
  ctx->bsVersion = c[0]; ctx->blockSize = c[1]; ctx->size = c[2]; ctx->jobs = c[3]; ctx->dataType = c[4]; ctx->flags = c[5];
}

static jint transform(JNIEnv* env, int forward, jint type, jintArray jctx, jbyteArray jsrc, jint srcIdx, jint srcLen,
                      jbyteArray jdst, jint dstIdx, jint dstLen, jintArray jio) {
  jint c[6];
  kzg_ctx ctx;
  load_ctx(env, jctx, &ctx, c);
  if (srcIdx < 0 || dstIdx < 0 || srcLen < 0 || dstLen < 0) return 0;
  const jint srcArr = (*env)->GetArrayLength(env, jsrc), dstArr = (*env)->GetArrayLength(env, jdst);
  if ((int64_t)srcIdx + srcLen > srcArr || dstIdx > dstArr) return 0;          /* the guard block every Java codec opens with */
  const jint dstCap = dstArr - dstIdx;
  int32_t used[2] = {0, 0};
  /* critical sections pin (or copy) the Java arrays; the library copies to the device and never keeps the pointers */
  jbyte* src = (jbyte*)(*env)->GetPrimitiveArrayCritical(env, jsrc, 0);
  if (src == 0) return -KZG_ERR_UNKNOWN;
  jbyte* dst = (jbyte*)(*env)->GetPrimitiveArrayCritical(env, jdst, 0);
  if (dst == 0) { (*env)->ReleasePrimitiveArrayCritical(env, jsrc, src, JNI_ABORT); return -KZG_ERR_UNKNOWN; }
  const int r = forward ? kzg_transform_forward(type, &ctx, (const uint8_t*)src + srcIdx, srcLen, (uint8_t*)dst + dstIdx, dstLen < dstCap ? dstLen : dstCap, dstCap, &used[0], &used[1])
                        : kzg_transform_inverse(type, &ctx, (const uint8_t*)src + srcIdx, srcLen, (uint8_t*)dst + dstIdx, dstLen < dstCap ? dstLen : dstCap, dstCap, &used[0], &used[1]);
  (*env)->ReleasePrimitiveArrayCritical(env, jdst, dst, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, jsrc, src, JNI_ABORT);
  c[4] = ctx.dataType;                                                          /* ROLZ may have sniffed a data type (ROLZCodec.java:451-461) */
  (*env)->SetIntArrayRegion(env, jctx, 0, 6, c);
  (*env)->SetIntArrayRegion(env, jio, 0, 2, (const jint*)used);
  return r;
}

JNIEXPORT jint JNICALL Java_io_github_flanglet_kanzi_gpu_GpuTransform_forward0(JNIEnv* env, jclass cls, jint type, jintArray jctx, jbyteArray jsrc,
    jint srcIdx, jint srcLen, jbyteArray jdst, jint dstIdx, jint dstLen, jintArray jio) {
  (void)cls;
  return transform(env, 1, type, jctx, jsrc, srcIdx, srcLen, jdst, dstIdx, dstLen, jio);
}

JNIEXPORT jint JNICALL Java_io_github_flanglet_kanzi_gpu_GpuTransform_inverse0(JNIEnv* env, jclass cls, jint type, jintArray jctx, jbyteArray jsrc,
    jint srcIdx, jint srcLen, jbyteArray jdst, jint dstIdx, jint dstLen, jintArray jio) {
  (void)cls;
  return transform(env, 0, type, jctx, jsrc, srcIdx, srcLen, jdst, dstIdx, dstLen, jio);
}

JNIEXPORT jint JNICALL Java_io_github_flanglet_kanzi_gpu_GpuTransform_maxLen0(JNIEnv* env, jclass cls, jint type, jint n) {
  (void)env; (void)cls;
  return kzg_transform_max_encoded_len(type, n);
}

/* encode0: block[blkptr, blkptr + count) -> MSB-first bit string in out[]; io[0..1] = bit count (low, high 32 bits).
 * Returns count (Java's return value) or < 0.  The Java side appends the bits with bitstream.writeBits(out, 0, bits). */
JNIEXPORT jint JNICALL Java_io_github_flanglet_kanzi_gpu_GpuEntropyEncoder_encode0(JNIEnv* env, jclass cls, jint type, jintArray jctx, jbyteArray jblock,
    jint blkptr, jint count, jbyteArray jout, jintArray jio) {
  (void)cls;
  jint c[6];
  kzg_ctx ctx;
  load_ctx(env, jctx, &ctx, c);
  if (blkptr < 0 || count < 0 || (int64_t)blkptr + count > (*env)->GetArrayLength(env, jblock)) return -1;
  const jint outCap = (*env)->GetArrayLength(env, jout);
  int64_t bits = 0;
  jbyte* src = (jbyte*)(*env)->GetPrimitiveArrayCritical(env, jblock, 0);
  if (src == 0) return -1;
  jbyte* out = (jbyte*)(*env)->GetPrimitiveArrayCritical(env, jout, 0);
  if (out == 0) { (*env)->ReleasePrimitiveArrayCritical(env, jblock, src, JNI_ABORT); return -1; }
  const int64_t r = kzg_entropy_encode(type, &ctx, (const uint8_t*)src + blkptr, count, (uint8_t*)out, outCap, &bits);
  (*env)->ReleasePrimitiveArrayCritical(env, jout, out, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, jblock, src, JNI_ABORT);
  jint io[2] = { (jint)(bits & 0xFFFFFFFF), (jint)(bits >> 32) };
  (*env)->SetIntArrayRegion(env, jio, 0, 2, io);
  return (jint)r;
}

/* decode0: reads from bit 0 of in[] (inBits available), writes count bytes to block[blkptr..); io[0..1] = bits consumed.
 * Returns the Java return value (count on success, 0 / short on a corrupt stream). */
JNIEXPORT jint JNICALL Java_io_github_flanglet_kanzi_gpu_GpuEntropyDecoder_decode0(JNIEnv* env, jclass cls, jint type, jintArray jctx, jbyteArray jin,
    jlong inBits, jbyteArray jblock, jint blkptr, jint count, jintArray jio) {
  (void)cls;
  jint c[6];
  kzg_ctx ctx;
  load_ctx(env, jctx, &ctx, c);
  if (blkptr < 0 || count < 0 || (int64_t)blkptr + count > (*env)->GetArrayLength(env, jblock)) return -1;
  if (inBits < 0 || ((inBits + 7) >> 3) > (*env)->GetArrayLength(env, jin)) return -1;
  int64_t used = 0;
  jbyte* in = (jbyte*)(*env)->GetPrimitiveArrayCritical(env, jin, 0);
  if (in == 0) return -1;
  jbyte* dst = (jbyte*)(*env)->GetPrimitiveArrayCritical(env, jblock, 0);
  if (dst == 0) { (*env)->ReleasePrimitiveArrayCritical(env, jin, in, JNI_ABORT); return -1; }
  const int32_t r = kzg_entropy_decode(type, &ctx, (const uint8_t*)in, inBits, &used, (uint8_t*)dst + blkptr, count);
  (*env)->ReleasePrimitiveArrayCritical(env, jblock, dst, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, jin, in, JNI_ABORT);
  jint io[2] = { (jint)(used & 0xFFFFFFFF), (jint)(used >> 32) };
  (*env)->SetIntArrayRegion(env, jio, 0, 2, io);
  return r;
}

/* library-level knobs the Java shim sets once (GpuTransform's static initialiser) */
JNIEXPORT jint JNICALL Java_io_github_flanglet_kanzi_gpu_GpuTransform_configure0(JNIEnv* env, jclass cls, jint device, jint maxBatch, jint windowMicros) {
  (void)env; (void)cls;
  const int r = kzg_set_device(device);
  if (r < 0) return r;
  return kzg_set_coalescing(maxBatch, windowMicros);
}
#endif /* KZG_HAVE_JNI */

/* keeps the translation unit non-empty (and the header parsed) where no JDK exists */
int kzg_jni_glue_abi(void) { return kzg_abi_version(); }
