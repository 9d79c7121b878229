/*
 * GpuEntropyDecoder — Kanzi EntropyDecoder (EntropyDecoder.java:23-47) backed by libkanzi_b200 through JNI.
 * Drop-in for HuffmanDecoder / ANSRangeDecoder / FPAQDecoder.  DecodingTask hands every decoder a private bitstream over
 * the block's own bytes (io/CompressedInputStream.java:1250-1316), so the shim reads what is left of it, lets the device
 * decode, and the bits the codec consumed are simply the ones read here (nothing follows the entropy payload in a block).
 */
package io.github.flanglet.kanzi.gpu;

import io.github.flanglet.kanzi.EntropyDecoder;
import io.github.flanglet.kanzi.InputBitStream;
import java.util.Map;

public final class GpuEntropyDecoder implements EntropyDecoder {
   private final InputBitStream bitstream;
   private final int type;
   private final long availableBits;                     // ctx "blockBits": bit length of the entropy payload (DecodingTask knows it)
   private final int[] ctx = new int[6];
   private final int[] io = new int[2];

   public GpuEntropyDecoder(InputBitStream bitstream, Map<String, Object> ctx, int type, long availableBits) {
      if (bitstream == null)
         throw new NullPointerException("Invalid null bitstream parameter");
      this.bitstream = bitstream;
      this.type = type;
      this.availableBits = availableBits;
      this.ctx[0] = (ctx == null) ? 7 : (Integer) ctx.getOrDefault("bsVersion", 7);
   }

   private static native int decode0(int type, int[] ctx, byte[] in, long inBits, byte[] block, int blkptr, int count, int[] io);

   @Override
   public int decode(byte[] block, int blkptr, int count) {
      if ((block == null) || (blkptr + count > block.length) || (blkptr < 0) || (count < 0))
         return -1;
      if (count == 0)
         return 0;
      final byte[] in = new byte[(int) ((this.availableBits + 7) >>> 3) + 8];
      long left = this.availableBits;
      int off = 0;
      while (left > 0) {
         final int n = (int) Math.min(left, 1L << 30);
         this.bitstream.readBits(in, off, n);
         off += n >>> 3;
         left -= n;
      }
      return decode0(this.type, this.ctx, in, this.availableBits, block, blkptr, count, this.io);
   }

   @Override
   public InputBitStream getBitStream() {
      return this.bitstream;
   }

   @Override
   public void dispose() {
   }
}
