"""kanzi_b200 — B200-native (sm_100a CUDA) hot path behind Kanzi's ByteTransform / EntropyEncoder /
EntropyDecoder plugin interfaces.  The compute lives in libkanzi_b200.so (C ABI: include/kzg.h); this
package is the Python-side mirror of the reference interfaces used by the tests and the benchmark.
There is no CPU fallback: importing works anywhere, every compute call needs a CUDA device."""
import os as _os
# up to 32 block groups run on streams of their own (csrc/lz_forward2.cu); must be set before the CUDA context exists
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
from .binding import (lib, KzgError, T, E, DT, FLAG_BWT_ASREF, FLAG_XXH32, FLAG_XXH64, device_count, set_device, last_error, launch_count,
                      transform_forward, transform_inverse, transform_max_encoded_len, bwt_forward, bwt_inverse,
                      entropy_encode, entropy_decode, compress, decompress, compress_bound, last_block_bits, stream_index)
from .interfaces import SliceByteArray, ByteTransform, EntropyEncoder, EntropyDecoder, OutputBitStream, InputBitStream, \
    TransformFactory, EntropyCodecFactory

__all__ = [n for n in dir() if not n.startswith("_")]
