/*
 * kzg_sibling_check.c — TEST INFRASTRUCTURE: a torch-free parity check of the LZP stage on a real GPU, for when a box is only
 * available for a minute (no Python start-up): libkanzi_b200 (linked) against the oracle (dlopen of oracle/libkzoracle.so), on
 * word-salad text with long re-pasted passages, flag bytes sprinkled in, of the RLT stage on run-heavy bytes and of ROLZX on both.
 *   A  per-block LZP forward vs the oracle's bytes, inverse of those bytes vs the input     (kzg_transform_forward / _inverse)
 *   B  whole streams through chains with LZP vs the oracle's stream, then kzg_decompress    (kzg_compress / kzg_decompress)
 *   C  the same for LZ&ANS0 and ROLZ&ANS0: the chains that existed before must still match
 * Prints one line per check and "ALL PASS" / "FAILED n"; exit code = number of failed checks.
 *   usage: kzg_sibling_check <path to libkzoracle.so>
 */
#define _POSIX_C_SOURCE 200809L
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "../../include/kzg.h"

typedef int (*kzo_transform_fn)(int, int, int32_t*, const uint8_t*, int32_t, int32_t, uint8_t*, int32_t, int32_t, int32_t*, int32_t*);
typedef int64_t (*kzo_compress_fn)(const uint8_t*, int64_t, const int32_t*, int, int, int32_t, int64_t, int, uint8_t*, int64_t);
static kzo_transform_fn kzo_transform;
static kzo_compress_fn kzo_compress_stream;
static int failed = 0;
static double now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static void verdict(const char* what, int ok, const char* detail) { printf("%s %s %s\n", ok ? "PASS" : "FAIL", what, detail ? detail : ""); fflush(stdout); if (!ok) failed++; }

static uint64_t rs = 88172645463325252ull;
static uint32_t rnd(void) { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t)(rs >> 16); }
static void fill(uint8_t* p, size_t n) {
  static char vocab[2048][10]; static int vlen[2048]; static int init = 0;
  if (!init) { for (int i = 0; i < 2048; i++) { vlen[i] = 2 + (int)(rnd() % 8); for (int k = 0; k < vlen[i]; k++) vocab[i][k] = (char)('a' + rnd() % 26); } init = 1; }
  size_t o = 0;
  while (o < n) {
    const uint32_t r = rnd();
    if ((r & 63) == 0 && o > 4096) {                       /* re-paste an earlier passage: what LZP codes as matches */
      size_t from = rnd() % (o - 2048), len = 70 + rnd() % 1800;
      for (size_t k = 0; k < len && o < n; k++) p[o++] = p[from + k];
      continue;
    }
    if ((r & 1023) == 1) { p[o++] = (uint8_t)(0xFC + rnd() % 4); continue; }
    const int w = (int)((((uint64_t)(r & 0xFFFF)) * (r >> 16)) >> 21) & 2047;
    for (int k = 0; k < vlen[w] && o < n; k++) p[o++] = (uint8_t)vocab[w][k];
    if (o < n) p[o++] = ' ';
  }
}

static void fill_runs(uint8_t* p, size_t n) {
  size_t o = 0;
  while (o < n) {
    const uint32_t r = rnd();
    const uint8_t v = (r & 7) == 0 ? 0xFB : ((r & 7) == 1 ? 0 : (uint8_t)(r >> 8));
    size_t len = 1 + (r >> 16) % 12;
    if ((r & 0x300) == 0) len = 200 + rnd() % 9000;
    if ((r & 0x7F00) == 0) len = 60000 + rnd() % 30000;
    for (size_t k = 0; k < len && o < n; k++) p[o++] = v;
  }
}

/* ROLZX per-block: forward bytes and data type against the oracle, inverse of the oracle's bytes, a destination one byte short */
static void check_rolzx(const uint8_t* d, int32_t n, const char* tag) {
  char what[128], detail[256] = "";
  const int32_t cap = n <= 16384 ? n + 1024 : n + n / 32;
  uint8_t* ref = (uint8_t*)calloc((size_t)cap + 64, 1); uint8_t* got = (uint8_t*)calloc((size_t)cap + 64, 1); uint8_t* back = (uint8_t*)calloc((size_t)n + 64, 1);
  int32_t cv[6] = {7, n > 1024 ? n : 1024, n, 1, 0, 0}, su = 0, du = 0, gsu = 0, gdu = 0;
  const int okRef = kzo_transform(KZG_T_ROLZX, 0, cv, d, n, n, ref, cap, cap, &su, &du);
  kzg_ctx ctx = {7, n > 1024 ? n : 1024, n, 1, 0, 0};
  double t0 = now();
  const int ok = kzg_transform_forward(KZG_T_ROLZX, &ctx, d, n, got, cap, cap, &gsu, &gdu);
  const double tf = now() - t0;
  snprintf(what, sizeof(what), "A ROLZX forward %s n=%d", tag, n);
  int good = (ok == okRef) && ctx.dataType == cv[4] && (!ok || (gdu == du && gsu == su && memcmp(got, ref, (size_t)du) == 0));
  snprintf(detail, sizeof(detail), "(ok %d/%d, bytes %d/%d, type %d/%d, %.1f ms) %s", ok, okRef, gdu, du, ctx.dataType, cv[4], 1e3 * tf, good ? "" : kzg_last_error());
  verdict(what, good, detail);
  if (okRef == 1) {
    int32_t isu = 0, idu = 0;
    t0 = now();
    const int oki = kzg_transform_inverse(KZG_T_ROLZX, &ctx, ref, du, back, n, n, &isu, &idu);
    const double ti = now() - t0;
    snprintf(what, sizeof(what), "A ROLZX inverse %s n=%d", tag, n);
    good = (oki == 1) && idu == n && isu == du && memcmp(back, d, (size_t)n) == 0;
    snprintf(detail, sizeof(detail), "(ok %d, bytes %d, used %d/%d, %.1f ms) %s", oki, idu, isu, du, 1e3 * ti, good ? "" : kzg_last_error());
    verdict(what, good, detail);
    const int k2 = kzg_transform_inverse(KZG_T_ROLZX, &ctx, ref, du, back, n - 1, n - 1, &isu, &idu);
    snprintf(what, sizeof(what), "A ROLZX inverse short dst %s n=%d", tag, n);
    snprintf(detail, sizeof(detail), "(%d)", k2);
    verdict(what, k2 == 0, detail);
    if (du > 40) {                                          /* a truncated stream: refused by both */
      int32_t cv2[6] = {7, n, n, 1, 0, 0}, osu = 0, odu = 0;
      const int o3 = kzo_transform(KZG_T_ROLZX, 1, cv2, ref, du / 2, du / 2, got, n, n, &osu, &odu);
      const int k3 = kzg_transform_inverse(KZG_T_ROLZX, &ctx, ref, du / 2, back, n, n, &isu, &idu);
      snprintf(what, sizeof(what), "A ROLZX inverse truncated %s n=%d", tag, n);
      snprintf(detail, sizeof(detail), "(%d/%d)", k3, o3);
      verdict(what, k3 == (o3 < 0 ? 0 : o3), detail);
    }
  }
  free(ref); free(got); free(back);
}

/* RLT per-block: entropy id e decides the escape byte (ctx flags bits 8-11 on our side, ctxv[5] >> 8 on the oracle's) */
static void check_rlt(const uint8_t* d, int32_t n, int e) {
  char what[128], detail[256] = "";
  const int32_t cap = n <= 512 ? n + 32 : n;
  uint8_t* ref = (uint8_t*)calloc((size_t)cap + 64, 1); uint8_t* got = (uint8_t*)calloc((size_t)cap + 64, 1); uint8_t* back = (uint8_t*)calloc((size_t)n + 64, 1);
  int32_t cv[6] = {7, n > 1024 ? n : 1024, n, 1, 0, e << 8}, su = 0, du = 0, gsu = 0, gdu = 0;
  const int okRef = kzo_transform(KZG_T_RLT, 0, cv, d, n, n, ref, cap, cap, &su, &du);
  kzg_ctx ctx = {7, n > 1024 ? n : 1024, n, 1, 0, KZG_CTX_ENTROPY(e)};
  const int ok = kzg_transform_forward(KZG_T_RLT, &ctx, d, n, got, cap, cap, &gsu, &gdu);
  snprintf(what, sizeof(what), "A RLT forward n=%d entropy=%d", n, e);
  int good = (ok == okRef) && ctx.dataType == cv[4] && (!ok || (gdu == du && gsu == su && memcmp(got, ref, (size_t)du) == 0));
  snprintf(detail, sizeof(detail), "(ok %d/%d, bytes %d/%d, type %d/%d) %s", ok, okRef, gdu, du, ctx.dataType, cv[4], good ? "" : kzg_last_error());
  verdict(what, good, detail);
  if (okRef == 1) {
    int32_t isu = 0, idu = 0;
    const int oki = kzg_transform_inverse(KZG_T_RLT, &ctx, ref, du, back, n, n, &isu, &idu);
    snprintf(what, sizeof(what), "A RLT inverse n=%d", n);
    good = (oki == 1) && idu == n && isu == du && memcmp(back, d, (size_t)n) == 0;
    snprintf(detail, sizeof(detail), "(ok %d, bytes %d, used %d/%d) %s", oki, idu, isu, du, good ? "" : kzg_last_error());
    verdict(what, good, detail);
    /* one byte short: refused, unless the byte that does not fit is a literal escape at the very end (RLT.java:308-313 drops it and
     * still reports success); whatever the oracle says */
    int32_t cv2[6] = {7, n, n, 1, 0, 0}, osu = 0, odu = 0;
    const int o2 = kzo_transform(KZG_T_RLT, 1, cv2, ref, du, du, got, n - 1, n - 1, &osu, &odu);
    const int k2 = kzg_transform_inverse(KZG_T_RLT, &ctx, ref, du, back, n - 1, n - 1, &isu, &idu);
    snprintf(what, sizeof(what), "A RLT inverse short dst n=%d", n);
    snprintf(detail, sizeof(detail), "(%d/%d, bytes %d/%d)", k2, o2, idu, odu);
    verdict(what, k2 == (o2 < 0 ? 0 : o2) && (k2 != 1 || (idu == odu && memcmp(back, got, (size_t)idu) == 0)), detail);
  }
  free(ref); free(got); free(back);
}

static void check_block(const uint8_t* d, int32_t n) {
  char what[128], detail[256] = "";
  const int32_t cap = n + n / 64 + 1100;
  uint8_t* ref = (uint8_t*)calloc((size_t)cap + 64, 1); uint8_t* got = (uint8_t*)calloc((size_t)cap + 64, 1); uint8_t* back = (uint8_t*)calloc((size_t)n + 64, 1);
  int32_t cv[6] = {7, n > 1024 ? n : 1024, n, 1, 0, 0}, su = 0, du = 0, gsu = 0, gdu = 0;
  const int okRef = kzo_transform(KZG_T_LZP, 0, cv, d, n, n, ref, cap, cap, &su, &du);
  kzg_ctx ctx = {7, n > 1024 ? n : 1024, n, 1, 0, 0};
  double t0 = now();
  const int ok = kzg_transform_forward(KZG_T_LZP, &ctx, d, n, got, cap, cap, &gsu, &gdu);
  const double tf = now() - t0;
  snprintf(what, sizeof(what), "A forward n=%d", n);
  int good = (ok == okRef) && (!ok || (gdu == du && gsu == su && memcmp(got, ref, (size_t)du) == 0));
  snprintf(detail, sizeof(detail), "(ok %d/%d, bytes %d/%d, %.1f ms) %s", ok, okRef, gdu, du, 1e3 * tf, good ? "" : kzg_last_error());
  verdict(what, good, detail);
  if (okRef == 1) {
    int32_t isu = 0, idu = 0;
    t0 = now();
    const int oki = kzg_transform_inverse(KZG_T_LZP, &ctx, ref, du, back, n, n, &isu, &idu);
    const double ti = now() - t0;
    snprintf(what, sizeof(what), "A inverse n=%d", n);
    good = (oki == 1) && idu == n && isu == du && memcmp(back, d, (size_t)n) == 0;
    snprintf(detail, sizeof(detail), "(ok %d, bytes %d, used %d/%d, %.1f ms) %s", oki, idu, isu, du, 1e3 * ti, good ? "" : kzg_last_error());
    verdict(what, good, detail);
    if (n > 200) {                                         /* a destination one byte short must be refused, as the oracle refuses it */
      int32_t cv2[6] = {7, n, n, 1, 0, 0};
      const int o2 = kzo_transform(KZG_T_LZP, 1, cv2, ref, du, du, got, n - 1, n - 1, &su, &gsu);
      const int k2 = kzg_transform_inverse(KZG_T_LZP, &ctx, ref, du, back, n - 1, n - 1, &isu, &idu);
      snprintf(what, sizeof(what), "A inverse short dst n=%d", n);
      snprintf(detail, sizeof(detail), "(%d/%d)", k2, o2);
      verdict(what, k2 == o2 && k2 == 0, detail);
    }
  }
  free(ref); free(got); free(back);
}

static void check_stream(const char* name, const uint8_t* d, int64_t n, const int32_t* ids, int nIds, int ent, int32_t bs) {
  char what[160], detail[256];
  const int64_t cap = kzg_compress_bound(n, bs) + 4096;
  uint8_t* ref = (uint8_t*)calloc((size_t)cap, 1); uint8_t* got = (uint8_t*)calloc((size_t)cap, 1); uint8_t* back = (uint8_t*)calloc((size_t)n + 64, 1);
  const int64_t r = kzo_compress_stream(d, n, ids, nIds, ent, bs, n, 1, ref, cap);
  double t0 = now();
  const int64_t g = kzg_compress(d, n, ids, nIds, ent, bs, KZG_FLAG_BWT_ASREF, got, cap);
  const double tc = now() - t0;
  int good = r > 0 && g == r && memcmp(got, ref, (size_t)r) == 0;
  snprintf(what, sizeof(what), "%s compress n=%lld bs=%d", name, (long long)n, bs);
  snprintf(detail, sizeof(detail), "(bytes %lld/%lld, %.1f ms) %s", (long long)g, (long long)r, 1e3 * tc, good ? "" : kzg_last_error());
  verdict(what, good, detail);
  if (r > 0) {
    t0 = now();
    const int64_t b = kzg_decompress(ref, r, KZG_FLAG_BWT_ASREF, back, n);
    const double td = now() - t0;
    good = b == n && memcmp(back, d, (size_t)n) == 0;
    snprintf(what, sizeof(what), "%s decompress", name);
    snprintf(detail, sizeof(detail), "(bytes %lld, %.1f ms) %s", (long long)b, 1e3 * td, good ? "" : kzg_last_error());
    verdict(what, good, detail);
  }
  free(ref); free(got); free(back);
}

int main(int argc, char** argv) {
  void* h = dlopen(argc > 1 ? argv[1] : "oracle/libkzoracle.so", RTLD_NOW);
  if (!h) { printf("FAIL dlopen oracle: %s\n", dlerror()); return 99; }
  kzo_transform = (kzo_transform_fn)dlsym(h, "kzo_transform");
  kzo_compress_stream = (kzo_compress_fn)dlsym(h, "kzo_compress_stream");
  if (!kzo_transform || !kzo_compress_stream) { printf("FAIL dlsym\n"); return 98; }
  if (kzg_device_count() < 1) { printf("FAIL no CUDA device: %s\n", kzg_last_error()); return 97; }
  const size_t N = (size_t)3 << 20 | 12345;
  uint8_t* d = (uint8_t*)malloc(N + 64);
  fill(d, N);
  double t0 = now();
  { int32_t su, du; uint8_t tmp[600]; kzg_ctx c = {7, 1024, 300, 1, 0, 0}; kzg_transform_forward(KZG_T_ZRLT, &c, d, 300, tmp, 600, 600, &su, &du); }   /* context + arena warm-up */
  printf("warm-up %.1f ms\n", 1e3 * (now() - t0));
  const int32_t sizes[] = {127, 128, 129, 200, 5000, 70001, 1 << 20, (3 << 20) + 77};
  for (unsigned i = 0; i < sizeof(sizes) / sizeof(sizes[0]); i++) check_block(d + ((size_t)sizes[i] + 60000 < N ? (i * 4099) % 50000 : 0), sizes[i]);
  { uint8_t* z = (uint8_t*)calloc(300000, 1); check_block(z, 300000); memset(z, 0xFC, 300000); check_block(z, 300000); free(z); }
  uint8_t* rr = (uint8_t*)malloc(N + 64);
  fill_runs(rr, N);
  const int32_t rsizes[] = {15, 16, 600, 70001, 1 << 20};
  for (unsigned i = 0; i < sizeof(rsizes) / sizeof(rsizes[0]); i++) { check_rlt(rr + i * 1234, rsizes[i], KZG_E_NONE); check_rlt(rr + i * 1234, rsizes[i], KZG_E_FPAQ); }
  { uint8_t* z = (uint8_t*)malloc(40000); for (int i = 0; i < 40000; i++) z[i] = "ACGT"[(i * 7 + (i >> 3)) & 3]; check_rlt(z, 40000, KZG_E_ANS1); check_rlt(z, 40000, KZG_E_ANS0); free(z); }
  const int32_t rlt[] = {KZG_T_RLT}, rltlzp[] = {KZG_T_RLT, KZG_T_LZP};
  check_stream("B RLT&ANS0", rr, (int64_t)N, rlt, 1, KZG_E_ANS0, 1 << 18);
  check_stream("B RLT&FPAQ", rr, (int64_t)N, rlt, 1, KZG_E_FPAQ, 1 << 20);
  check_stream("B RLT+LZP&HUFFMAN", rr, (int64_t)1 << 21, rltlzp, 2, KZG_E_HUFFMAN, 1 << 19);
  check_stream("B RLT&NONE text", d, 700001, rlt, 1, KZG_E_NONE, 1 << 16);
  const int32_t xsizes[] = {63, 64, 5000, 70001, 1 << 20};
  for (unsigned i = 0; i < sizeof(xsizes) / sizeof(xsizes[0]); i++) check_rolzx(d + i * 777, xsizes[i], "text");
  check_rolzx(rr, 300000, "runs");
  { uint8_t* z = (uint8_t*)malloc(60000); for (int i = 0; i < 60000; i++) z[i] = "ACGT"[(rnd() >> 3) & 3]; memcpy(z + 30000, z + 1000, 20000); check_rolzx(z, 60000, "dna"); free(z); }
  { const int32_t big = (17 << 20) + 4321; uint8_t* z = (uint8_t*)calloc((size_t)big + 64, 1);                /* two 16 MiB chunks: mostly zeros, a short passage every 64 KiB */
    for (int32_t o = 0; o + 64 < big; o += 65536) memcpy(z + o, d + (o >> 10), 48);
    check_rolzx(z, big, "two chunks"); free(z); }
  const int32_t rx[] = {KZG_T_ROLZX}, rxz[] = {KZG_T_RLT, KZG_T_ROLZX};
  check_stream("B ROLZX&NONE", d, (int64_t)N, rx, 1, KZG_E_NONE, 1 << 18);
  check_stream("B ROLZX&ANS0", d, (int64_t)1 << 21, rx, 1, KZG_E_ANS0, 1 << 20);
  check_stream("B RLT+ROLZX&HUFFMAN", rr, (int64_t)1 << 21, rxz, 2, KZG_E_HUFFMAN, 1 << 19);
  const int32_t lzp[] = {KZG_T_LZP}, lzpz[] = {KZG_T_LZP, KZG_T_ZRLT}, lz[] = {KZG_T_LZ}, rolz[] = {KZG_T_ROLZ}, rl[] = {KZG_T_ROLZ, KZG_T_LZP};
  check_stream("B LZP&ANS0", d, (int64_t)N, lzp, 1, KZG_E_ANS0, 1 << 20);
  check_stream("B LZP+ZRLT&HUFFMAN", d, (int64_t)N, lzpz, 2, KZG_E_HUFFMAN, 1 << 18);
  check_stream("B LZP&NONE", d, 700001, lzp, 1, KZG_E_NONE, 1 << 16);
  check_stream("B ROLZ+LZP&ANS0", d, (int64_t)1 << 20, rl, 2, KZG_E_ANS0, 1 << 19);
  check_stream("C LZ&ANS0", d, (int64_t)N, lz, 1, KZG_E_ANS0, 1 << 20);
  check_stream("C ROLZ&ANS0", d, (int64_t)N, rolz, 1, KZG_E_ANS0, 1 << 20);
  if (failed) printf("FAILED %d\n", failed); else printf("ALL PASS\n");
  return failed;
}
