/*
 * GpuTransform — Kanzi ByteTransform backed by libkanzi_b200 (B200 / sm_100a CUDA kernels) through JNI.
 * Drop-in for LZCodec / ROLZCodec / BWTBlockCodec / SBRT / SRT / ZRLT behind io.github.flanglet.kanzi.ByteTransform
 * (java/src/main/java/io/github/flanglet/kanzi/ByteTransform.java:24-57): same slices, same return values, same ctx side
 * effects.  TransformFactory.newFunctionToken (transform/TransformFactory.java:273-351) returns one of these per id when the
 * system property kanzi.gpu is set.  Not compiled in this repository (the build image has no JDK); see INTEGRATION.md.
 */
package io.github.flanglet.kanzi.gpu;

import io.github.flanglet.kanzi.ByteTransform;
import io.github.flanglet.kanzi.Global;
import io.github.flanglet.kanzi.SliceByteArray;
import java.util.Map;

public final class GpuTransform implements ByteTransform {
   // TransformFactory ids (transform/TransformFactory.java:36-58) == KZG_T_* of include/kzg.h
   public static final int BWT_TYPE = 1, LZ_TYPE = 3, ZRLT_TYPE = 6, MTFT_TYPE = 7, RANK_TYPE = 8, ROLZ_TYPE = 11, SRT_TYPE = 13, LZX_TYPE = 16;
   public static final int FLAG_BWT_ASREF = 1;      // keep BWT.java:152-156 as written (what today's streams contain)

   static {
      System.loadLibrary("kanzi_b200_jni");        // links libkanzi_b200.so; export CUDA_DEVICE_MAX_CONNECTIONS=32 before the JVM starts
      // one device per JVM here; a multi-GPU host calls configure0(blockId % deviceCount, ...) from each pool thread instead.
      // maxBatch 64 / 200 us: per-block calls of the <= 64 EncodingTask threads are coalesced into batched launches
      configure0(Integer.getInteger("kanzi.gpu.device", 0), Integer.getInteger("kanzi.gpu.batch", 64), Integer.getInteger("kanzi.gpu.windowMicros", 200));
   }

   private final int type;
   private final Map<String, Object> map;
   private final int[] ctx = new int[6];            // bsVersion, blockSize, size, jobs, dataType ordinal, flags
   private final int[] io = new int[2];             // bytes consumed, bytes produced

   public GpuTransform(int type, Map<String, Object> ctx) {
      this.type = type;
      this.map = ctx;
      this.ctx[0] = (ctx == null) ? 7 : (Integer) ctx.getOrDefault("bsVersion", 7);
      this.ctx[1] = (ctx == null) ? 0 : (Integer) ctx.getOrDefault("blockSize", 0);
      this.ctx[3] = (ctx == null) ? 1 : (Integer) ctx.getOrDefault("jobs", 1);
      this.ctx[5] = FLAG_BWT_ASREF;
      // ctx["entropy"] for RLT's choice of escape byte (RLT.java:101-107): KZG_CTX_ENTROPY(id) = (id + 1) << 8, 0 = key absent
      if ((ctx != null) && ctx.containsKey("entropy")) {
         int id = 14;                               // a codec this library has no id for (CM, TPAQ ...): "not one of the four plain ones"
         switch (String.valueOf(ctx.get("entropy")).toUpperCase()) {
            case "NONE": id = 0; break;
            case "HUFFMAN": id = 1; break;
            case "FPAQ": id = 2; break;
            case "RANGE": id = 4; break;
            case "ANS0": id = 5; break;
            case "ANS1": id = 8; break;
            default: break;
         }
         this.ctx[5] |= ((id + 1) & 0xF) << 8;
      }
   }

   static native int configure0(int device, int maxBatch, int windowMicros);
   private static native int forward0(int type, int[] ctx, byte[] src, int srcIdx, int srcLen, byte[] dst, int dstIdx, int dstLen, int[] io);
   private static native int inverse0(int type, int[] ctx, byte[] src, int srcIdx, int srcLen, byte[] dst, int dstIdx, int dstLen, int[] io);
   private static native int maxLen0(int type, int n);

   private void loadCtx(SliceByteArray src) {
      this.ctx[2] = src.length;
      Global.DataType dt = (this.map == null) ? Global.DataType.UNDEFINED
            : (Global.DataType) this.map.getOrDefault("dataType", Global.DataType.UNDEFINED);
      this.ctx[4] = dt.ordinal();                  // Global.java:40-90 ordinals == KZG_DT_*
   }

   private void storeCtx() {
      if ((this.map != null) && (this.ctx[4] != Global.DataType.UNDEFINED.ordinal()))
         this.map.put("dataType", Global.DataType.values()[this.ctx[4]]);       // ROLZCodec.java:451-461 writes it back
   }

   private boolean run(boolean forward, SliceByteArray src, SliceByteArray dst) {
      if (src.length == 0)
         return true;                              // every codec: `if (input.length == 0) return true`
      if (src.array == dst.array)
         return false;
      loadCtx(src);
      final int r = forward ? forward0(this.type, this.ctx, src.array, src.index, src.length, dst.array, dst.index, dst.length, this.io)
                            : inverse0(this.type, this.ctx, src.array, src.index, src.length, dst.array, dst.index, dst.length, this.io);
      if (r < 0)
         throw new IllegalStateException("libkanzi_b200: error " + (-r));     // surfaces as ERR_PROCESS_BLOCK in EncodingTask / DecodingTask
      if (r == 1) {
         src.index += this.io[0];
         dst.index += this.io[1];
         storeCtx();
      }
      return r == 1;                               // false: Sequence keeps the skip flag set (Sequence.java:95-105) / fails the block on inverse
   }

   @Override
   public boolean forward(SliceByteArray src, SliceByteArray dst) {
      return run(true, src, dst);
   }

   @Override
   public boolean inverse(SliceByteArray src, SliceByteArray dst) {
      return run(false, src, dst);
   }

   @Override
   public int getMaxEncodedLength(int srcLen) {
      return maxLen0(this.type, srcLen);
   }
}
