/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see fq.h).
 *
 * Keccak-f[1600] (FIPS 202), SHAKE256, STROBE-128 and the Merlin v1.0 transcript.
 * merlin ^3.0.0 and sha3 ^0.8.2 are third-party crates that are NOT under /root/reference
 * (Cargo.toml:19,22); this restates their published constructions and is pinned by the Merlin
 * conformance vector and hashlib (tests/test_oracle_transcript.py).  Call sites it serves:
 *   /root/reference/src/transcript.rs:13-37, random.rs:11-27, commitments.rs:15-33.
 */
#include <stdint.h>
#include <string.h>
#include <stddef.h>

static const uint64_t RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int ROTC[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
static const int PILN[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
#define ROL64(x, n) (((x) << (n)) | ((x) >> (64 - (n))))

void keccak_f1600(uint64_t st[25]) {
  uint64_t bc[5], t;
  for (int round = 0; round < 24; round++) {
    for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
    for (int i = 0; i < 5; i++) {
      t = bc[(i + 4) % 5] ^ ROL64(bc[(i + 1) % 5], 1);
      for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
    }
    t = st[1];
    for (int i = 0; i < 24; i++) { int j = PILN[i]; bc[0] = st[j]; st[j] = ROL64(t, ROTC[i]); t = bc[0]; }
    for (int j = 0; j < 25; j += 5) {
      for (int i = 0; i < 5; i++) bc[i] = st[j + i];
      for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
    }
    st[0] ^= RC[round];
  }
}

/* SHAKE256(input) -> outlen bytes (rate 136, domain 0x1f) */
void shake256(uint8_t *out, size_t outlen, const uint8_t *in, size_t inlen) {
  uint64_t st[25]; uint8_t *sb = (uint8_t *)st; const size_t rate = 136;
  memset(st, 0, sizeof st);
  while (inlen >= rate) { for (size_t i = 0; i < rate; i++) sb[i] ^= in[i]; keccak_f1600(st); in += rate; inlen -= rate; }
  for (size_t i = 0; i < inlen; i++) sb[i] ^= in[i];
  sb[inlen] ^= 0x1f; sb[rate - 1] ^= 0x80;
  keccak_f1600(st);
  while (outlen > 0) {
    size_t n = outlen < rate ? outlen : rate;
    memcpy(out, sb, n); out += n; outlen -= n;
    if (outlen) keccak_f1600(st);
  }
}

/* ---------------- STROBE-128 (the subset Merlin uses) ---------------- */
#define STROBE_R 166
#define FLAG_I 1
#define FLAG_A 2
#define FLAG_C 4
#define FLAG_T 8
#define FLAG_M 16
#define FLAG_K 32

typedef struct {
  uint64_t st[25]; /* viewed as 200 bytes */
  uint8_t pos, pos_begin, cur_flags;
} strobe_t; /* 208 bytes incl. padding */

static void strobe_run_f(strobe_t *s) {
  uint8_t *b = (uint8_t *)s->st;
  b[s->pos] ^= s->pos_begin;
  b[s->pos + 1] ^= 0x04;
  b[STROBE_R + 1] ^= 0x80;
  keccak_f1600(s->st);
  s->pos = 0; s->pos_begin = 0;
}
static void strobe_absorb(strobe_t *s, const uint8_t *d, size_t n) {
  uint8_t *b = (uint8_t *)s->st;
  for (size_t i = 0; i < n; i++) { b[s->pos] ^= d[i]; s->pos++; if (s->pos == STROBE_R) strobe_run_f(s); }
}
static void strobe_squeeze(strobe_t *s, uint8_t *d, size_t n) {
  uint8_t *b = (uint8_t *)s->st;
  for (size_t i = 0; i < n; i++) { d[i] = b[s->pos]; b[s->pos] = 0; s->pos++; if (s->pos == STROBE_R) strobe_run_f(s); }
}
static void strobe_begin_op(strobe_t *s, uint8_t flags, int more) {
  if (more) return; /* continuation of the same operation */
  uint8_t old_begin = s->pos_begin;
  s->pos_begin = s->pos + 1;
  s->cur_flags = flags;
  uint8_t hdr[2] = {old_begin, flags};
  strobe_absorb(s, hdr, 2);
  int force_f = (flags & (FLAG_C | FLAG_K)) != 0;
  if (force_f && s->pos != 0) strobe_run_f(s);
}
static void strobe_init(strobe_t *s, const uint8_t *proto, size_t n) {
  uint8_t *b = (uint8_t *)s->st;
  memset(s, 0, sizeof *s);
  const uint8_t hdr[6] = {1, STROBE_R + 2, 1, 0, 1, 96};
  memcpy(b, hdr, 6);
  memcpy(b + 6, "STROBEv1.0.2", 12);
  keccak_f1600(s->st);
  s->pos = 0; s->pos_begin = 0; s->cur_flags = 0;
  strobe_begin_op(s, FLAG_M | FLAG_A, 0); /* meta_ad(protocol label) */
  strobe_absorb(s, proto, n);
}
static void strobe_meta_ad(strobe_t *s, const uint8_t *d, size_t n, int more) { strobe_begin_op(s, FLAG_M | FLAG_A, more); strobe_absorb(s, d, n); }
static void strobe_ad(strobe_t *s, const uint8_t *d, size_t n, int more) { strobe_begin_op(s, FLAG_A, more); strobe_absorb(s, d, n); }
static void strobe_prf(strobe_t *s, uint8_t *d, size_t n, int more) { strobe_begin_op(s, FLAG_I | FLAG_A | FLAG_C, more); strobe_squeeze(s, d, n); }

/* ---------------- Merlin v1.0 ---------------- */
typedef strobe_t merlin_t;
size_t merlin_sizeof(void) { return sizeof(merlin_t); }

void merlin_append_message(merlin_t *t, const uint8_t *label, size_t llen, const uint8_t *msg, size_t mlen) {
  uint8_t len4[4] = {(uint8_t)mlen, (uint8_t)(mlen >> 8), (uint8_t)(mlen >> 16), (uint8_t)(mlen >> 24)};
  strobe_meta_ad(t, label, llen, 0);
  strobe_meta_ad(t, len4, 4, 1);
  strobe_ad(t, msg, mlen, 0);
}
void merlin_init(merlin_t *t, const uint8_t *label, size_t llen) {
  strobe_init(t, (const uint8_t *)"Merlin v1.0", 11);
  merlin_append_message(t, (const uint8_t *)"dom-sep", 7, label, llen);
}
void merlin_challenge_bytes(merlin_t *t, const uint8_t *label, size_t llen, uint8_t *out, size_t n) {
  uint8_t len4[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
  strobe_meta_ad(t, label, llen, 0);
  strobe_meta_ad(t, len4, 4, 1);
  strobe_prf(t, out, n, 0);
}
/* append n 32-byte items under the same label (transcript.rs:49-57 inner loop, dense_mlpoly.rs:295-297) */
void merlin_append_many32(merlin_t *t, const uint8_t *label, size_t llen, const uint8_t *items, size_t n) {
  for (size_t i = 0; i < n; i++) merlin_append_message(t, label, llen, items + 32 * i, 32);
}
