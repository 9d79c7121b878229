/*
 * GpuEntropyEncoder — Kanzi EntropyEncoder (EntropyEncoder.java:23-49) backed by libkanzi_b200 through JNI.
 * Drop-in for HuffmanEncoder / ANSRangeEncoder / FPAQEncoder: encode() appends exactly the bits
 * `new XEncoder(bitstream, ctx).encode(block, blkptr, count); dispose()` would (io/CompressedOutputStream.java:907-916).
 * EntropyCodecFactory.newEncoder (entropy/EntropyCodecFactory.java:76-128) returns one when kanzi.gpu is set.
 */
package io.github.flanglet.kanzi.gpu;

import io.github.flanglet.kanzi.EntropyEncoder;
import io.github.flanglet.kanzi.OutputBitStream;
import java.util.Map;

public final class GpuEntropyEncoder implements EntropyEncoder {
   public static final int HUFFMAN_TYPE = 1, FPAQ_TYPE = 2, ANS0_TYPE = 5, ANS1_TYPE = 8;      // == KZG_E_*

   private final OutputBitStream bitstream;
   private final int type;
   private final int[] ctx = new int[6];
   private final int[] io = new int[2];
   private byte[] out = new byte[0];

   public GpuEntropyEncoder(OutputBitStream bitstream, Map<String, Object> ctx, int type) {
      if (bitstream == null)
         throw new NullPointerException("Invalid null bitstream parameter");
      this.bitstream = bitstream;
      this.type = type;
      this.ctx[0] = (ctx == null) ? 7 : (Integer) ctx.getOrDefault("bsVersion", 7);
   }

   private static native int encode0(int type, int[] ctx, byte[] block, int blkptr, int count, byte[] out, int[] io);

   @Override
   public int encode(byte[] block, int blkptr, int count) {
      if ((block == null) || (blkptr + count > block.length) || (blkptr < 0) || (count < 0))
         return -1;
      if (count == 0)
         return 0;
      final int need = 2 * count + (300 << 10);         // ANS1 on noise: 256 context headers + up to 2 bytes per symbol
      if (this.out.length < need)
         this.out = new byte[need];
      final int r = encode0(this.type, this.ctx, block, blkptr, count, this.out, this.io);
      if (r != count)
         return r;
      long bits = (this.io[0] & 0xFFFFFFFFL) | ((long) this.io[1] << 32);
      int off = 0;
      while (bits > 0) {                                   // writeBits takes an int bit count: slices of at most 2^30 bits
         final int n = (int) Math.min(bits, 1L << 30);
         this.bitstream.writeBits(this.out, off, n);
         off += n >>> 3;
         bits -= n;
      }
      return count;
   }

   @Override
   public OutputBitStream getBitStream() {
      return this.bitstream;
   }

   @Override
   public void dispose() {
      // everything, including what the Java encoders flush in dispose() (FPAQEncoder.java:232-238), was written by encode()
   }
}
