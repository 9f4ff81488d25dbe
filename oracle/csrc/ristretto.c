/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see ristretto.h / fq.h).
 * ristretto255 per RFC 9496 on top of edwards25519 extended coordinates.
 * Replaces (for the checker) curve25519-dalek's RistrettoPoint used through
 * /root/reference/src/group.rs:6-7,28-46,98-117 and commitments.rs:25 (from_uniform_bytes).
 */
#include "ristretto.h"
#include <string.h>
#include <stdlib.h>

typedef unsigned __int128 u128;
#define MASK51 ((1ULL << 51) - 1)

static const fe_t FE_ZERO = {{0, 0, 0, 0, 0}};
static const fe_t FE_ONE = {{1, 0, 0, 0, 0}};
static const fe_t FE_D = {{0x34dca135978a3ULL, 0x1a8283b156ebdULL, 0x5e7a26001c029ULL, 0x739c663a03cbbULL, 0x52036cee2b6ffULL}};
static const fe_t FE_2D = {{0x69b9426b2f159ULL, 0x35050762add7aULL, 0x3cf44c0038052ULL, 0x6738cc7407977ULL, 0x2406d9dc56dffULL}};
static const fe_t FE_SQRT_M1 = {{0x61b274a0ea0b0ULL, 0xd5a5fc8f189dULL, 0x7ef5e9cbd0c60ULL, 0x78595a6804c9eULL, 0x2b8324804fc1dULL}};
static const fe_t FE_SQRT_AD_MINUS_ONE = {{0x7f6a0497b2e1bULL, 0x1836f0a97afd2ULL, 0x7d747f6be7638ULL, 0x456079e7e6498ULL, 0x376931bf2b834ULL}};
static const fe_t FE_INVSQRT_A_MINUS_D = {{0xfdaa805d40eaULL, 0x2eb482e57d339ULL, 0x7610274bc58ULL, 0x6510b613dc8ffULL, 0x786c8905cfaffULL}};
static const fe_t FE_ONE_MINUS_D_SQ = {{0x409c1945fc176ULL, 0x719abc6a1fc4fULL, 0x1c37f90b20684ULL, 0x6bccca55eedfULL, 0x29072a8b2b3eULL}};
static const fe_t FE_D_MINUS_ONE_SQ = {{0x55aaa44ed4d20ULL, 0x59603c3332635ULL, 0x26d3baf4a7928ULL, 0x120a66e6997a9ULL, 0x5968b37af66c2ULL}};

const uint8_t RISTRETTO_BASEPOINT_COMPRESSED[32] = {
    0xe2, 0xf2, 0xae, 0x0a, 0x6a, 0xbc, 0x4e, 0x71, 0xa8, 0x84, 0xa9, 0x61, 0xc5, 0x00, 0x51, 0x5f,
    0x58, 0xe3, 0x0b, 0x6a, 0xa5, 0x82, 0xdd, 0x8d, 0xb6, 0xa6, 0x59, 0x45, 0xe0, 0x8d, 0x2d, 0x76};

/* ---------------- field 2^255-19, radix 2^51 ---------------- */
static inline void fe_weak_reduce(fe_t *r) {
  uint64_t c;
  c = r->v[0] >> 51; r->v[0] &= MASK51; r->v[1] += c;
  c = r->v[1] >> 51; r->v[1] &= MASK51; r->v[2] += c;
  c = r->v[2] >> 51; r->v[2] &= MASK51; r->v[3] += c;
  c = r->v[3] >> 51; r->v[3] &= MASK51; r->v[4] += c;
  c = r->v[4] >> 51; r->v[4] &= MASK51; r->v[0] += c * 19;
}
static inline void fe_add(fe_t *r, const fe_t *a, const fe_t *b) {
  for (int i = 0; i < 5; i++) r->v[i] = a->v[i] + b->v[i];
  fe_weak_reduce(r);
}
static inline void fe_sub(fe_t *r, const fe_t *a, const fe_t *b) {
  /* a + 16p - b keeps every limb non-negative for limbs < 2^54 */
  r->v[0] = a->v[0] + 36028797018963664ULL - b->v[0];
  r->v[1] = a->v[1] + 36028797018963952ULL - b->v[1];
  r->v[2] = a->v[2] + 36028797018963952ULL - b->v[2];
  r->v[3] = a->v[3] + 36028797018963952ULL - b->v[3];
  r->v[4] = a->v[4] + 36028797018963952ULL - b->v[4];
  fe_weak_reduce(r);
}
static inline void fe_neg(fe_t *r, const fe_t *a) { fe_sub(r, &FE_ZERO, a); }

static void fe_mul(fe_t *r, const fe_t *a, const fe_t *b) {
  const uint64_t *x = a->v, *y = b->v;
  uint64_t y1_19 = y[1] * 19, y2_19 = y[2] * 19, y3_19 = y[3] * 19, y4_19 = y[4] * 19;
  u128 c0 = (u128)x[0] * y[0] + (u128)x[4] * y1_19 + (u128)x[3] * y2_19 + (u128)x[2] * y3_19 + (u128)x[1] * y4_19;
  u128 c1 = (u128)x[1] * y[0] + (u128)x[0] * y[1] + (u128)x[4] * y2_19 + (u128)x[3] * y3_19 + (u128)x[2] * y4_19;
  u128 c2 = (u128)x[2] * y[0] + (u128)x[1] * y[1] + (u128)x[0] * y[2] + (u128)x[4] * y3_19 + (u128)x[3] * y4_19;
  u128 c3 = (u128)x[3] * y[0] + (u128)x[2] * y[1] + (u128)x[1] * y[2] + (u128)x[0] * y[3] + (u128)x[4] * y4_19;
  u128 c4 = (u128)x[4] * y[0] + (u128)x[3] * y[1] + (u128)x[2] * y[2] + (u128)x[1] * y[3] + (u128)x[0] * y[4];
  c1 += (uint64_t)(c0 >> 51); uint64_t o0 = (uint64_t)c0 & MASK51;
  c2 += (uint64_t)(c1 >> 51); uint64_t o1 = (uint64_t)c1 & MASK51;
  c3 += (uint64_t)(c2 >> 51); uint64_t o2 = (uint64_t)c2 & MASK51;
  c4 += (uint64_t)(c3 >> 51); uint64_t o3 = (uint64_t)c3 & MASK51;
  uint64_t carry = (uint64_t)(c4 >> 51); uint64_t o4 = (uint64_t)c4 & MASK51;
  o0 += carry * 19;
  o1 += o0 >> 51; o0 &= MASK51;
  r->v[0] = o0; r->v[1] = o1; r->v[2] = o2; r->v[3] = o3; r->v[4] = o4;
}
static inline void fe_sq(fe_t *r, const fe_t *a) { fe_mul(r, a, a); }
static void fe_sqn(fe_t *r, const fe_t *a, int n) {
  fe_sq(r, a);
  for (int i = 1; i < n; i++) fe_sq(r, r);
}

static void fe_tobytes(uint8_t s[32], const fe_t *a) {
  fe_t t = *a;
  fe_weak_reduce(&t); fe_weak_reduce(&t);
  /* compute q = floor((t + 19) / 2^255), then t + 19 q mod 2^255 is canonical */
  uint64_t q = (t.v[0] + 19) >> 51;
  q = (t.v[1] + q) >> 51; q = (t.v[2] + q) >> 51; q = (t.v[3] + q) >> 51; q = (t.v[4] + q) >> 51;
  t.v[0] += 19 * q;
  uint64_t c;
  c = t.v[0] >> 51; t.v[0] &= MASK51; t.v[1] += c;
  c = t.v[1] >> 51; t.v[1] &= MASK51; t.v[2] += c;
  c = t.v[2] >> 51; t.v[2] &= MASK51; t.v[3] += c;
  c = t.v[3] >> 51; t.v[3] &= MASK51; t.v[4] += c;
  t.v[4] &= MASK51;
  uint64_t w[4];
  w[0] = t.v[0] | (t.v[1] << 51);
  w[1] = (t.v[1] >> 13) | (t.v[2] << 38);
  w[2] = (t.v[2] >> 26) | (t.v[3] << 25);
  w[3] = (t.v[3] >> 39) | (t.v[4] << 12);
  memcpy(s, w, 32);
}
static void fe_frombytes(fe_t *r, const uint8_t s[32]) { /* ignores bit 255 */
  uint64_t w[4];
  memcpy(w, s, 32);
  r->v[0] = w[0] & MASK51;
  r->v[1] = ((w[0] >> 51) | (w[1] << 13)) & MASK51;
  r->v[2] = ((w[1] >> 38) | (w[2] << 26)) & MASK51;
  r->v[3] = ((w[2] >> 25) | (w[3] << 39)) & MASK51;
  r->v[4] = (w[3] >> 12) & MASK51;
}
static int fe_is_negative(const fe_t *a) { uint8_t s[32]; fe_tobytes(s, a); return s[0] & 1; }
static int fe_is_zero(const fe_t *a) {
  uint8_t s[32]; fe_tobytes(s, a);
  uint8_t acc = 0; for (int i = 0; i < 32; i++) acc |= s[i];
  return acc == 0;
}
static int fe_equal(const fe_t *a, const fe_t *b) {
  uint8_t s[32], t[32]; fe_tobytes(s, a); fe_tobytes(t, b);
  return memcmp(s, t, 32) == 0;
}
static void fe_abs(fe_t *r, const fe_t *a) { if (fe_is_negative(a)) fe_neg(r, a); else *r = *a; }

/* z^(2^250 - 1) shared prefix, then the two tails */
static void fe_pow_250m1(fe_t *t250, fe_t *z11, const fe_t *z) {
  fe_t z2, z9, z_5_0, z_10_0, z_20_0, z_40_0, z_50_0, z_100_0, t;
  fe_sq(&z2, z);
  fe_sqn(&t, &z2, 2);
  fe_mul(&z9, &t, z);
  fe_mul(z11, &z9, &z2);
  fe_sq(&t, z11);
  fe_mul(&z_5_0, &t, &z9);
  fe_sqn(&t, &z_5_0, 5);   fe_mul(&z_10_0, &t, &z_5_0);
  fe_sqn(&t, &z_10_0, 10); fe_mul(&z_20_0, &t, &z_10_0);
  fe_sqn(&t, &z_20_0, 20); fe_mul(&z_40_0, &t, &z_20_0);
  fe_sqn(&t, &z_40_0, 10); fe_mul(&z_50_0, &t, &z_10_0);
  fe_sqn(&t, &z_50_0, 50); fe_mul(&z_100_0, &t, &z_50_0);
  fe_sqn(&t, &z_100_0, 100); fe_mul(&t, &t, &z_100_0);
  fe_sqn(&t, &t, 50);      fe_mul(t250, &t, &z_50_0);
}
static void fe_invert(fe_t *r, const fe_t *z) { /* z^(p-2) = z^(2^255-21) */
  fe_t t250, z11, t;
  fe_pow_250m1(&t250, &z11, z);
  fe_sqn(&t, &t250, 5);
  fe_mul(r, &t, &z11);
}
static void fe_pow22523(fe_t *r, const fe_t *z) { /* z^((p-5)/8) = z^(2^252-3) */
  fe_t t250, z11, t;
  fe_pow_250m1(&t250, &z11, z);
  fe_sqn(&t, &t250, 2);
  fe_mul(r, &t, z);
}

/* RFC 9496 section 4.2 SQRT_RATIO_M1: returns was_square, r = sqrt(u/v) or sqrt(i*u/v), r non-negative */
static int fe_sqrt_ratio_i(fe_t *r, const fe_t *u, const fe_t *v) {
  fe_t v3, v7, t, rr, check, neg_u, neg_u_i, r_prime;
  fe_sq(&t, v); fe_mul(&v3, &t, v);
  fe_sq(&t, &v3); fe_mul(&v7, &t, v);
  fe_mul(&t, u, &v7); fe_pow22523(&t, &t);
  fe_mul(&rr, u, &v3); fe_mul(&rr, &rr, &t);
  fe_sq(&t, &rr); fe_mul(&check, v, &t);
  fe_neg(&neg_u, u);
  fe_mul(&neg_u_i, &neg_u, &FE_SQRT_M1);
  int correct_sign = fe_equal(&check, u);
  int flipped_sign = fe_equal(&check, &neg_u);
  int flipped_sign_i = fe_equal(&check, &neg_u_i);
  fe_mul(&r_prime, &rr, &FE_SQRT_M1);
  if (flipped_sign || flipped_sign_i) rr = r_prime;
  fe_abs(r, &rr);
  return correct_sign || flipped_sign;
}

/* ---------------- extended Edwards points ---------------- */
void ge_identity(ge_t *r) { r->X = FE_ZERO; r->Y = FE_ONE; r->Z = FE_ONE; r->T = FE_ZERO; }

void ge_add(ge_t *r, const ge_t *p, const ge_t *q) { /* add-2008-hwcd-3, a = -1 */
  fe_t A, B, C, D, E, F, G, H, t0, t1;
  fe_sub(&t0, &p->Y, &p->X); fe_sub(&t1, &q->Y, &q->X); fe_mul(&A, &t0, &t1);
  fe_add(&t0, &p->Y, &p->X); fe_add(&t1, &q->Y, &q->X); fe_mul(&B, &t0, &t1);
  fe_mul(&C, &p->T, &q->T); fe_mul(&C, &C, &FE_2D);
  fe_mul(&D, &p->Z, &q->Z); fe_add(&D, &D, &D);
  fe_sub(&E, &B, &A); fe_sub(&F, &D, &C); fe_add(&G, &D, &C); fe_add(&H, &B, &A);
  fe_mul(&r->X, &E, &F); fe_mul(&r->Y, &G, &H); fe_mul(&r->T, &E, &H); fe_mul(&r->Z, &F, &G);
}
void ge_neg(ge_t *r, const ge_t *p) { fe_neg(&r->X, &p->X); r->Y = p->Y; r->Z = p->Z; fe_neg(&r->T, &p->T); }
void ge_sub(ge_t *r, const ge_t *p, const ge_t *q) { ge_t n; ge_neg(&n, q); ge_add(r, p, &n); }
void ge_double(ge_t *r, const ge_t *p) { /* dbl-2008-hwcd, a = -1 */
  fe_t A, B, C, D, E, F, G, H, t;
  fe_sq(&A, &p->X); fe_sq(&B, &p->Y);
  fe_sq(&C, &p->Z); fe_add(&C, &C, &C);
  fe_neg(&D, &A);
  fe_add(&t, &p->X, &p->Y); fe_sq(&t, &t); fe_sub(&t, &t, &A); fe_sub(&E, &t, &B);
  fe_add(&G, &D, &B); fe_sub(&F, &G, &C); fe_sub(&H, &D, &B);
  fe_mul(&r->X, &E, &F); fe_mul(&r->Y, &G, &H); fe_mul(&r->T, &E, &H); fe_mul(&r->Z, &F, &G);
}
int ge_eq(const ge_t *p, const ge_t *q) { /* RFC 9496 4.3.3 */
  fe_t a, b, c, d;
  fe_mul(&a, &p->X, &q->Y); fe_mul(&b, &p->Y, &q->X);
  fe_mul(&c, &p->Y, &q->Y); fe_mul(&d, &p->X, &q->X);
  return fe_equal(&a, &b) || fe_equal(&c, &d);
}

/* ---------------- ristretto255 (RFC 9496 4.3) ---------------- */
int ristretto_decode(ge_t *r, const uint8_t sb[32]) {
  fe_t s, ss, u1, u2, u2_sqr, v, t, invsqrt, den_x, den_y, x, y, tt;
  uint8_t chk[32];
  fe_frombytes(&s, sb);
  fe_tobytes(chk, &s);
  if (memcmp(chk, sb, 32) != 0 || (sb[0] & 1)) return 0; /* non-canonical or negative */
  fe_sq(&ss, &s);
  fe_sub(&u1, &FE_ONE, &ss);
  fe_add(&u2, &FE_ONE, &ss);
  fe_sq(&u2_sqr, &u2);
  fe_sq(&t, &u1); fe_mul(&t, &t, &FE_D); fe_neg(&t, &t); fe_sub(&v, &t, &u2_sqr);
  fe_mul(&t, &v, &u2_sqr);
  int was_square = fe_sqrt_ratio_i(&invsqrt, &FE_ONE, &t);
  fe_mul(&den_x, &invsqrt, &u2);
  fe_mul(&den_y, &invsqrt, &den_x); fe_mul(&den_y, &den_y, &v);
  fe_add(&t, &s, &s); fe_mul(&t, &t, &den_x); fe_abs(&x, &t);
  fe_mul(&y, &u1, &den_y);
  fe_mul(&tt, &x, &y);
  if (!was_square || fe_is_negative(&tt) || fe_is_zero(&y)) return 0;
  r->X = x; r->Y = y; r->Z = FE_ONE; r->T = tt;
  return 1;
}

void ristretto_encode(uint8_t out[32], const ge_t *p) {
  fe_t u1, u2, t, invsqrt, den1, den2, z_inv, ix0, iy0, ench, x, y, den_inv, s;
  fe_add(&u1, &p->Z, &p->Y); fe_sub(&t, &p->Z, &p->Y); fe_mul(&u1, &u1, &t);
  fe_mul(&u2, &p->X, &p->Y);
  fe_sq(&t, &u2); fe_mul(&t, &t, &u1);
  (void)fe_sqrt_ratio_i(&invsqrt, &FE_ONE, &t);
  fe_mul(&den1, &invsqrt, &u1);
  fe_mul(&den2, &invsqrt, &u2);
  fe_mul(&z_inv, &den1, &den2); fe_mul(&z_inv, &z_inv, &p->T);
  fe_mul(&ix0, &p->X, &FE_SQRT_M1);
  fe_mul(&iy0, &p->Y, &FE_SQRT_M1);
  fe_mul(&ench, &den1, &FE_INVSQRT_A_MINUS_D);
  fe_mul(&t, &p->T, &z_inv);
  int rotate = fe_is_negative(&t);
  if (rotate) { x = iy0; y = ix0; den_inv = ench; } else { x = p->X; y = p->Y; den_inv = den2; }
  fe_mul(&t, &x, &z_inv);
  if (fe_is_negative(&t)) fe_neg(&y, &y);
  fe_sub(&t, &p->Z, &y); fe_mul(&s, &den_inv, &t); fe_abs(&s, &s);
  fe_tobytes(out, &s);
}
void ristretto_encode_batch(uint8_t *out, const ge_t *p, size_t n) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) ristretto_encode(out + 32 * i, &p[i]);
}

static void ristretto_map(ge_t *r, const fe_t *t) { /* RFC 9496 4.3.4 MAP */
  fe_t rr, u, v, c, s, s_prime, N, w0, w1, w2, w3, tmp;
  fe_sq(&rr, t); fe_mul(&rr, &rr, &FE_SQRT_M1);
  fe_add(&u, &rr, &FE_ONE); fe_mul(&u, &u, &FE_ONE_MINUS_D_SQ);
  fe_mul(&tmp, &rr, &FE_D); fe_neg(&v, &FE_ONE); fe_sub(&v, &v, &tmp); /* -1 - r*D */
  fe_add(&tmp, &rr, &FE_D); fe_mul(&v, &v, &tmp);
  int was_square = fe_sqrt_ratio_i(&s, &u, &v);
  fe_mul(&s_prime, &s, t); fe_abs(&s_prime, &s_prime); fe_neg(&s_prime, &s_prime);
  if (!was_square) { s = s_prime; c = rr; } else { fe_neg(&c, &FE_ONE); }
  fe_sub(&tmp, &rr, &FE_ONE); fe_mul(&N, &c, &tmp); fe_mul(&N, &N, &FE_D_MINUS_ONE_SQ); fe_sub(&N, &N, &v);
  fe_add(&w0, &s, &s); fe_mul(&w0, &w0, &v);
  fe_mul(&w1, &N, &FE_SQRT_AD_MINUS_ONE);
  fe_sq(&tmp, &s); fe_sub(&w2, &FE_ONE, &tmp); fe_add(&w3, &FE_ONE, &tmp);
  fe_mul(&r->X, &w0, &w3); fe_mul(&r->Y, &w2, &w1); fe_mul(&r->Z, &w1, &w3); fe_mul(&r->T, &w0, &w2);
}
void ristretto_from_uniform_bytes(ge_t *r, const uint8_t b[64]) {
  fe_t t1, t2; ge_t p1, p2;
  fe_frombytes(&t1, b); fe_frombytes(&t2, b + 32);
  ristretto_map(&p1, &t1); ristretto_map(&p2, &t2);
  ge_add(r, &p1, &p2);
}

/* ---------------- scalar multiplication / MSM ---------------- */
void ge_scalarmul_bytes(ge_t *r, const uint8_t k[32], const ge_t *p) {
  ge_t acc; ge_identity(&acc);
  int started = 0;
  for (int i = 255; i >= 0; i--) {
    if (started) ge_double(&acc, &acc);
    if ((k[i >> 3] >> (i & 7)) & 1) { ge_add(&acc, &acc, p); started = 1; }
  }
  *r = acc;
}
void ge_scalarmul(ge_t *r, const fq_t *k, const ge_t *p) {
  uint8_t kb[32]; fq_to_bytes(kb, k); ge_scalarmul_bytes(r, kb, p);
}

/* signed radix-2^w digits of a 256-bit little-endian integer (< 2^255): ndig = ceil(256/w) + 1 */
static int radix_2w(int16_t *digits, const uint8_t k[32], int w) {
  int ndig = (256 + w - 1) / w + 1;
  uint64_t limbs[5] = {0, 0, 0, 0, 0};
  memcpy(limbs, k, 32);
  int64_t carry = 0, radix = 1LL << w, mask = radix - 1;
  for (int i = 0; i < ndig; i++) {
    int bit = i * w, word = bit >> 6, off = bit & 63;
    uint64_t bits = 0;
    if (word < 4) {
      bits = limbs[word] >> off;
      if (off + w > 64 && word + 1 < 5) bits |= limbs[word + 1] << (64 - off);
    }
    int64_t coef = carry + (int64_t)(bits & (uint64_t)mask);
    carry = (coef + radix / 2) >> w;
    digits[i] = (int16_t)(coef - (carry << w));
  }
  return ndig;
}

/* dalek's vartime_multiscalar_mul picks Straus below 190 points and Pippenger with window
 * 6 / 7 / 8 at 190 / 500 / 800 points (third-party thresholds).  The result is
 * algorithm-independent; the structure only matters for the timed CPU baseline. */
static void msm_pippenger(ge_t *out, const uint8_t *kbytes, const ge_t *pts, size_t n, int w) {
  int ndig = (256 + w - 1) / w + 1;
  int16_t *digits = (int16_t *)malloc(sizeof(int16_t) * n * ndig);
  for (size_t i = 0; i < n; i++) radix_2w(digits + i * ndig, kbytes + 32 * i, w);
  size_t nb = (size_t)1 << (w - 1);
  ge_t *buckets = (ge_t *)malloc(sizeof(ge_t) * nb);
  ge_t total; ge_identity(&total);
  for (int d = ndig - 1; d >= 0; d--) {
    for (int s = 0; s < w; s++) ge_double(&total, &total);
    for (size_t b = 0; b < nb; b++) ge_identity(&buckets[b]);
    for (size_t i = 0; i < n; i++) {
      int dg = digits[i * ndig + d];
      if (dg > 0) ge_add(&buckets[dg - 1], &buckets[dg - 1], &pts[i]);
      else if (dg < 0) ge_sub(&buckets[-dg - 1], &buckets[-dg - 1], &pts[i]);
    }
    ge_t run, sum; ge_identity(&run); ge_identity(&sum);
    for (size_t b = nb; b-- > 0;) { ge_add(&run, &run, &buckets[b]); ge_add(&sum, &sum, &run); }
    ge_add(&total, &total, &sum);
  }
  *out = total;
  free(buckets); free(digits);
}

static void msm_straus(ge_t *out, const uint8_t *kbytes, const ge_t *pts, size_t n) {
  /* signed radix-16 interleaved (Straus): table of 1..8 multiples per point */
  const int w = 4; int ndig = (256 + w - 1) / w + 1;
  int16_t *digits = (int16_t *)malloc(sizeof(int16_t) * n * ndig);
  ge_t *tab = (ge_t *)malloc(sizeof(ge_t) * n * 8);
  for (size_t i = 0; i < n; i++) {
    radix_2w(digits + i * ndig, kbytes + 32 * i, w);
    tab[i * 8] = pts[i];
    for (int j = 1; j < 8; j++) ge_add(&tab[i * 8 + j], &tab[i * 8 + j - 1], &pts[i]);
  }
  ge_t acc; ge_identity(&acc);
  for (int d = ndig - 1; d >= 0; d--) {
    for (int s = 0; s < w; s++) ge_double(&acc, &acc);
    for (size_t i = 0; i < n; i++) {
      int dg = digits[i * ndig + d];
      if (dg > 0) ge_add(&acc, &acc, &tab[i * 8 + dg - 1]);
      else if (dg < 0) ge_sub(&acc, &acc, &tab[i * 8 - dg - 1]);
    }
  }
  *out = acc;
  free(tab); free(digits);
}

void ge_msm(ge_t *r, const fq_t *scalars, const ge_t *points, size_t n) {
  if (n == 0) { ge_identity(r); return; }
  /* group.rs:110-113: every scalar is converted out of Montgomery form first */
  uint8_t *kb = (uint8_t *)malloc(32 * n);
  for (size_t i = 0; i < n; i++) fq_to_bytes(kb + 32 * i, &scalars[i]);
  if (n < 190) msm_straus(r, kb, points, n);
  else msm_pippenger(r, kb, points, n, n < 500 ? 6 : (n < 800 ? 7 : 8));
  free(kb);
}
