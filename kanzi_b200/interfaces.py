"""Python mirror of the reference's plugin interfaces for the hot path, backed by the C ABI.

Same names, argument meaning and error behaviour as
  K/ByteTransform.java:24-57      forward(src, dst) / inverse(src, dst) / getMaxEncodedLength(n)
  K/EntropyEncoder.java:23-49     encode(block, blkptr, count) / getBitStream() / dispose()
  K/EntropyDecoder.java:23-47     decode(block, blkptr, count) / getBitStream() / dispose()
  K/SliceByteArray.java:34-37     array / length / index
  K/transform/TransformFactory.java:273-351, K/entropy/EntropyCodecFactory.java:113-203  (id -> codec)
so that the parity tests read like the reference's own (T/test/TestTransforms.java, TestEntropyCodec.java).
The bit streams are the host-side MSB-first containers the Java host owns (K/bitstream/Default*BitStream.java)."""
from . import binding as B


class SliceByteArray:
    def __init__(self, array=None, length=None, index=0):
        self.array = bytearray() if array is None else (array if isinstance(array, bytearray) else bytearray(array))
        self.length = len(self.array) if length is None else length
        self.index = index


class OutputBitStream:
    """MSB-first bit appender (DefaultOutputBitStream.java:103-125,139-206)."""

    def __init__(self):
        self._chunks = []      # (bytes, nbits) pieces, flushed lazily
        self._acc = 0
        self._nacc = 0
        self._bytes = bytearray()

    def writeBit(self, b):
        self.writeBits(b & 1, 1)

    def writeBits(self, value, count=None, nbits=None):
        if isinstance(value, (bytes, bytearray, memoryview)):      # writeBits(byte[], start, count)
            start, n = (count or 0), nbits
            data = bytes(value[start: start + (n + 7) // 8])
            if self._nacc == 0 and n % 8 == 0:
                self._bytes += data
            else:
                v = int.from_bytes(data, "big") >> (len(data) * 8 - n) if n else 0
                self._acc = (self._acc << n) | v
                self._nacc += n
                self._flush()
            return n
        self._acc = (self._acc << count) | (value & ((1 << count) - 1))
        self._nacc += count
        self._flush()
        return count

    def _flush(self):
        nb = self._nacc // 8
        if nb:
            rem = self._nacc - 8 * nb
            self._bytes += (self._acc >> rem).to_bytes(nb, "big")
            self._acc &= (1 << rem) - 1
            self._nacc = rem

    def written(self):
        return len(self._bytes) * 8 + self._nacc

    def close(self):
        if self._nacc:
            self._bytes.append((self._acc << (8 - self._nacc)) & 0xFF)
            self._acc = 0
            self._nacc = 0

    def toByteArray(self):
        out = bytearray(self._bytes)
        if self._nacc:
            out.append((self._acc << (8 - self._nacc)) & 0xFF)
        return bytes(out)


class InputBitStream:
    """MSB-first bit reader (DefaultInputBitStream.java:97-192)."""

    def __init__(self, data, nbits=None):
        self._data = bytes(data)
        self._nbits = len(self._data) * 8 if nbits is None else nbits
        self._pos = 0

    def readBit(self):
        return self.readBits(1)

    def readBits(self, count):
        if self._pos + count > len(self._data) * 8:
            raise EOFError("No more data to read in the bitstream")
        first, last = self._pos // 8, (self._pos + count + 7) // 8
        v = int.from_bytes(self._data[first:last], "big")
        v >>= (last * 8 - self._pos - count)
        self._pos += count
        return v & ((1 << count) - 1)

    def read(self):
        return self._pos

    def remaining_bytes_and_bits(self):
        """(bytes from the current position re-aligned to bit 0, bit count) for the native decoders."""
        n = len(self._data) * 8 - self._pos
        v = int.from_bytes(self._data, "big") & ((1 << n) - 1) if n else 0
        nb = (n + 7) // 8
        return (v << (nb * 8 - n)).to_bytes(nb, "big") if nb else b"", n

    def skip(self, nbits):
        self._pos += nbits


class ByteTransform:
    """A ByteTransform whose forward/inverse run in libkanzi_b200 (K/ByteTransform.java:24-57)."""

    def __init__(self, kind, ctx=None):
        self.kind = kind
        self.ctx = ctx if ctx is not None else {}

    def getMaxEncodedLength(self, srcLength):
        return B.transform_max_encoded_len(self.kind, srcLength)

    def _call(self, fn, src, dst):
        if src.length == 0:
            return True
        if src.index < 0 or dst.index < 0 or src.length < 0 or src.index + src.length > len(src.array) or dst.index > len(dst.array):
            return False
        if src.array is dst.array:
            return False
        data = bytes(src.array[src.index: src.index + src.length])
        kctx = {k: v for k, v in self.ctx.items() if k in ("bsVersion", "blockSize", "size", "jobs", "dataType", "flags")}
        ok, out, used = fn(self.kind, data, kctx, dst_len=dst.length - dst.index, dst_cap=len(dst.array) - dst.index)
        if "dataType" in kctx:
            self.ctx["dataType"] = kctx["dataType"]
        if ok:
            dst.array[dst.index: dst.index + len(out)] = out
            src.index += used
            dst.index += len(out)
        return ok

    def forward(self, src, dst):
        return self._call(B.transform_forward, src, dst)

    def inverse(self, src, dst):
        return self._call(B.transform_inverse, src, dst)


class EntropyEncoder:
    """K/EntropyEncoder.java:23-49; encode() appends to the host bit stream what the Java codec would."""

    def __init__(self, kind, bitstream, ctx=None):
        self.kind, self.bs, self.ctx = kind, bitstream, ctx

    def encode(self, block, blkptr, count):
        if block is None or blkptr + count > len(block) or blkptr < 0 or count < 0:
            return -1
        payload, nbits = B.entropy_encode(self.kind, bytes(block[blkptr: blkptr + count]))
        self.bs.writeBits(payload, 0, nbits)
        return count

    def getBitStream(self):
        return self.bs

    def dispose(self):      # FPAQ's final 56 bits are part of encode() here (COS:910-916 always calls both)
        pass


class EntropyDecoder:
    """K/EntropyDecoder.java:23-47."""

    def __init__(self, kind, bitstream, ctx=None):
        self.kind, self.bs, self.ctx = kind, bitstream, ctx

    def decode(self, block, blkptr, count):
        if block is None or blkptr + count > len(block) or blkptr < 0 or count < 0:
            return -1
        data, nbits = self.bs.remaining_bytes_and_bits()
        out, r, used = B.entropy_decode(self.kind, data, nbits, count)
        block[blkptr: blkptr + count] = out
        self.bs.skip(used)
        return r

    def getBitStream(self):
        return self.bs

    def dispose(self):
        pass


class TransformFactory:
    """K/transform/TransformFactory.java: name -> ByteTransform."""
    NAMES = ("NONE", "BWT", "LZ", "LZX", "LZP", "RLT", "ROLZ", "ROLZX", "RANK", "MTFT", "SRT", "ZRLT")

    @staticmethod
    def newFunction(ctx, name):
        if name not in TransformFactory.NAMES:
            raise ValueError(f"Unknown transform type: '{name}'")
        return ByteTransform(name, ctx)


class EntropyCodecFactory:
    """K/entropy/EntropyCodecFactory.java:113-203."""
    NAMES = ("NONE", "HUFFMAN", "ANS0", "ANS1", "FPAQ")

    @staticmethod
    def newEncoder(obs, ctx, name):
        if name not in EntropyCodecFactory.NAMES:
            raise ValueError(f"Unknown entropy codec type: '{name}'")
        return EntropyEncoder(name, obs, ctx)

    @staticmethod
    def newDecoder(ibs, ctx, name):
        if name not in EntropyCodecFactory.NAMES:
            raise ValueError(f"Unsupported entropy codec type: '{name}'")
        return EntropyDecoder(name, ibs, ctx)
