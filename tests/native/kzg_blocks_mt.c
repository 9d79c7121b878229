/*
 * kzg_blocks_mt.c — drives libkanzi_b200 the way Kanzi's Java host does: T pool threads, ONE block per call.
 *
 * EncodingTask.encodeBlock (K/io/CompressedOutputStream.java:792-916): transform.forward(data, buffer) then
 * entropyEncoder.encode(buffer, 0, len) + dispose(); DecodingTask.decodeBlock (K/io/CompressedInputStream.java:1305-1344):
 * entropyDecoder.decode(buffer, 0, len) then transform.inverse(buffer, data).  Every thread walks its own blocks
 * (b = t, t + T, ...) through kzg_transform_forward -> kzg_entropy_encode and back, checks the round trip, and the
 * program reports MB/s of the per-block path, with the library's call coalescing off and on.
 *   usage: kzg_blocks_mt <threads> <blocks> <blockBytes> <transform id> <entropy id> <maxBatch> <windowMicros>
 * Prints one JSON line.  Exit code 0 = every block round-tripped.
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "../../include/kzg.h"

static int T = 16, NB = 32, BS = 1 << 20, XF = KZG_T_LZ, ENT = KZG_E_ANS0;
static uint8_t *data, *enc, *back;
static int64_t* encBits; static int32_t* xfLen; static int* xfOk;
static size_t encStride;
static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
static int failures = 0, failKind[6] = {0, 0, 0, 0, 0, 0}, firstCode = 0; static char firstErr[256] = "";
static void fail(int kind, int code) { pthread_mutex_lock(&mu); failures++; failKind[kind]++; if (!firstCode) { firstCode = code ? code : -999; snprintf(firstErr, sizeof(firstErr), "kind %d code %d: %s", kind, code, kzg_last_error()); } pthread_mutex_unlock(&mu); }
static pthread_barrier_t bar;
static double tEnc, tDec;

static double now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

/* word-salad text: compressible by LZ the way prose is (vocabulary of 4096 words drawn with a skewed law) */
static void fill(uint8_t* p, size_t n, uint64_t seed) {
  static char vocab[4096][12]; static int vlen[4096]; static int init = 0;
  uint64_t s = seed * 0x9E3779B97F4A7C15ull + 1;
  if (!init) {
    uint64_t v = 12345;
    for (int i = 0; i < 4096; i++) { v = v * 6364136223846793005ull + 1442695040888963407ull; vlen[i] = 2 + (int)((v >> 33) % 9);
      for (int k = 0; k < vlen[i]; k++) { v = v * 6364136223846793005ull + 1442695040888963407ull; vocab[i][k] = (char)('a' + (v >> 33) % 26); } }
    init = 1;
  }
  size_t o = 0;
  while (o < n) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    uint32_t r = (uint32_t)(s >> 32);
    int w = (int)(((uint64_t)(r & 0xFFFF) * (r >> 16)) >> 20) & 4095;      /* product of two uniforms: skewed towards small ids */
    for (int k = 0; k < vlen[w] && o < n; k++) p[o++] = (uint8_t)vocab[w][k];
    if (o < n) p[o++] = ((r & 31) == 0) ? '\n' : ' ';
  }
}

static void* worker(void* arg) {
  const int t = (int)(intptr_t)arg;
  const int32_t cap = kzg_transform_max_encoded_len(XF, BS);
  uint8_t* buf = (uint8_t*)malloc((size_t)cap + 64);
  uint8_t* tmp = (uint8_t*)malloc((size_t)cap + 1024);
  pthread_barrier_wait(&bar);
  const double t0 = now();
  for (int b = t; b < NB; b += T) {                       /* EncodingTask */
    kzg_ctx ctx = {7, BS, BS, 1, 0, 0};
    int32_t su = 0, du = 0;
    const uint8_t* in = data + (size_t)b * BS;
    int r = kzg_transform_forward(XF, &ctx, in, BS, buf, cap, cap, &su, &du);
    const uint8_t* payload = in; int32_t plen = BS;
    if (r == 1) { payload = buf; plen = du; }             /* r == 0: Sequence keeps the block untransformed (skip flag) */
    else if (r < 0) { fail(0, r); continue; }
    xfOk[b] = (r == 1); xfLen[b] = plen;
    int64_t bits = 0;
    const int64_t e = kzg_entropy_encode(ENT, &ctx, payload, plen, enc + (size_t)b * encStride, (int64_t)encStride, &bits);
    if (e != plen) { fail(1, (int)e); continue; }
    encBits[b] = bits;
  }
  pthread_barrier_wait(&bar);
  const double t1 = now();
  for (int b = t; b < NB; b += T) {                       /* DecodingTask */
    kzg_ctx ctx = {7, BS, BS, 1, 0, 0};
    int64_t used = 0;
    const int32_t d = kzg_entropy_decode(ENT, &ctx, enc + (size_t)b * encStride, encBits[b], &used, tmp, xfLen[b]);
    if (d != xfLen[b]) { fail(2, d); continue; }
    if (used != encBits[b]) { fail(3, (int)(used - encBits[b])); continue; }
    uint8_t* out = back + (size_t)b * BS;
    if (xfOk[b]) {
      int32_t su = 0, du = 0;
      const int r = kzg_transform_inverse(XF, &ctx, tmp, xfLen[b], out, BS, BS, &su, &du);
      if (r != 1 || du != BS) { fail(4, r); continue; }
    } else memcpy(out, tmp, BS);
    if (memcmp(out, data + (size_t)b * BS, BS) != 0) fail(5, 0);
  }
  pthread_barrier_wait(&bar);
  const double t2 = now();
  if (t == 0) { tEnc = t1 - t0; tDec = t2 - t1; }
  free(buf); free(tmp);
  return 0;
}

static void run(double* encMBps, double* decMBps) {
  pthread_t th[256];
  pthread_barrier_init(&bar, 0, (unsigned)T);
  for (int t = 0; t < T; t++) pthread_create(&th[t], 0, worker, (void*)(intptr_t)t);
  for (int t = 0; t < T; t++) pthread_join(th[t], 0);
  pthread_barrier_destroy(&bar);
  const double mb = (double)NB * BS / 1e6;
  *encMBps = mb / tEnc; *decMBps = mb / tDec;
}

int main(int argc, char** argv) {
  int maxBatch = 64, window = 200;
  if (argc > 1) T = atoi(argv[1]);
  if (argc > 2) NB = atoi(argv[2]);
  if (argc > 3) BS = atoi(argv[3]);
  if (argc > 4) XF = atoi(argv[4]);
  if (argc > 5) ENT = atoi(argv[5]);
  if (argc > 6) maxBatch = atoi(argv[6]);
  if (argc > 7) window = atoi(argv[7]);
  if (T < 1 || T > 256 || NB < 1 || BS < 1024) { fprintf(stderr, "bad arguments\n"); return 2; }
  if (kzg_device_count() < 1) { printf("{\"error\": \"no CUDA device (libkanzi_b200 has no CPU fallback)\"}\n"); return 3; }
  encStride = 2 * (size_t)BS + (300 << 10);
  data = (uint8_t*)malloc((size_t)NB * BS); back = (uint8_t*)malloc((size_t)NB * BS); enc = (uint8_t*)malloc((size_t)NB * encStride);
  encBits = (int64_t*)calloc(NB, sizeof(int64_t)); xfLen = (int32_t*)calloc(NB, sizeof(int32_t)); xfOk = (int*)calloc(NB, sizeof(int));
  for (int b = 0; b < NB; b++) fill(data + (size_t)b * BS, BS, 1000 + b);
  double e0, d0, e1, d1, ew, dw;
  kzg_set_device(0);
  run(&ew, &dw);                                          /* warm-up: arenas, module load */
  run(&e0, &d0);                                          /* one launch train per block and thread */
  int64_t encSum = 0; for (int b = 0; b < NB; b++) encSum += encBits[b];
  kzg_set_coalescing(maxBatch, window);
  run(&ew, &dw);
  int64_t b0 = 0; const int64_t r0 = kzg_coalescing_stats(&b0);
  run(&e1, &d1);                                          /* concurrent callers folded into batches */
  int64_t b1 = 0; const int64_t r1 = kzg_coalescing_stats(&b1);
  int64_t encSum2 = 0; for (int b = 0; b < NB; b++) encSum2 += encBits[b];
  kzg_set_coalescing(0, 0);
  const int same = encSum == encSum2;
  printf("{\"threads\": %d, \"blocks\": %d, \"block_bytes\": %d, \"transform\": %d, \"entropy\": %d, \"failures\": %d, \"fail_kinds\": [%d, %d, %d, %d, %d, %d], \"first_error\": \"%s\", \"coalesced_bits_equal\": %s, "
         "\"per_block\": {\"encode_MBps\": %.1f, \"decode_MBps\": %.1f, \"MBps\": %.1f}, "
         "\"coalesced\": {\"max_batch\": %d, \"window_us\": %d, \"encode_MBps\": %.1f, \"decode_MBps\": %.1f, \"MBps\": %.1f, \"requests\": %lld, \"batches\": %lld}, \"ratio\": %.3f}\n",
         T, NB, BS, XF, ENT, failures, failKind[0], failKind[1], failKind[2], failKind[3], failKind[4], failKind[5], firstErr, same ? "true" : "false", e0, d0, 1.0 / (1.0 / e0 + 1.0 / d0),
         maxBatch, window, e1, d1, 1.0 / (1.0 / e1 + 1.0 / d1), (long long)(r1 - r0), (long long)(b1 - b0), (double)NB * BS * 8.0 / (double)(encSum ? encSum : 1));
  return (failures == 0 && same) ? 0 : 1;
}
