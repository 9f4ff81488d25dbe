/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see fq.h).
 * Restates /root/reference/src/scalar/ristretto255.rs limb for limb:
 *   adc/sbb/mac helpers            ristretto255.rs:20-37
 *   constants MODULUS/INV/R/R2/R3  ristretto255.rs:248,304,307,315,323
 *   from_bytes :391  to_bytes :419  from_bytes_wide/from_u512 :435-466
 *   square :476  invert :541-595  batch_invert :597-639
 *   montgomery_reduce :642-686  mul :690-714  sub :718  add :736  neg :749
 */
#include "fq.h"
#include <string.h>
#include <stdlib.h>

typedef unsigned __int128 u128;

const fq_t FQ_MODULUS = {{0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0x0000000000000000ULL, 0x1000000000000000ULL}};
const fq_t FQ_R  = {{0xd6ec31748d98951dULL, 0xc6ef5bf4737dcf70ULL, 0xfffffffffffffffeULL, 0x0fffffffffffffffULL}};
const fq_t FQ_R2 = {{0xa40611e3449c0f01ULL, 0xd00e1ba768859347ULL, 0xceec73d217f5be65ULL, 0x0399411b7c309a3dULL}};
const fq_t FQ_R3 = {{0x2a9e49687b83a2dbULL, 0x278324e6aef7f3ecULL, 0x8065dc6c04ec5b65ULL, 0x0e530b773599cec7ULL}};

/* a + b + carry -> (lo, carry)                                  ristretto255.rs:20-24 */
static inline uint64_t adc(uint64_t a, uint64_t b, uint64_t carry, uint64_t *cout) {
  u128 ret = (u128)a + (u128)b + (u128)carry;
  *cout = (uint64_t)(ret >> 64);
  return (uint64_t)ret;
}
/* a - (b + borrow>>63) -> (lo, borrow mask)                     ristretto255.rs:27-31 */
static inline uint64_t sbb(uint64_t a, uint64_t b, uint64_t borrow, uint64_t *bout) {
  u128 ret = (u128)a - ((u128)b + (u128)(borrow >> 63));
  *bout = (uint64_t)(ret >> 64);
  return (uint64_t)ret;
}
/* a + b*c + carry -> (lo, carry)                                ristretto255.rs:34-37 */
static inline uint64_t mac(uint64_t a, uint64_t b, uint64_t c, uint64_t carry, uint64_t *cout) {
  u128 ret = (u128)a + (u128)b * (u128)c + (u128)carry;
  *cout = (uint64_t)(ret >> 64);
  return (uint64_t)ret;
}

void fq_sub(fq_t *r, const fq_t *a, const fq_t *b) {
  uint64_t borrow, carry, d0, d1, d2, d3;
  d0 = sbb(a->l[0], b->l[0], 0, &borrow);
  d1 = sbb(a->l[1], b->l[1], borrow, &borrow);
  d2 = sbb(a->l[2], b->l[2], borrow, &borrow);
  d3 = sbb(a->l[3], b->l[3], borrow, &borrow);
  /* borrow is an all-ones mask on underflow: conditionally add the modulus back */
  d0 = adc(d0, FQ_MODULUS.l[0] & borrow, 0, &carry);
  d1 = adc(d1, FQ_MODULUS.l[1] & borrow, carry, &carry);
  d2 = adc(d2, FQ_MODULUS.l[2] & borrow, carry, &carry);
  d3 = adc(d3, FQ_MODULUS.l[3] & borrow, carry, &carry);
  r->l[0] = d0; r->l[1] = d1; r->l[2] = d2; r->l[3] = d3;
}

void fq_add(fq_t *r, const fq_t *a, const fq_t *b) {
  uint64_t carry; fq_t t;
  t.l[0] = adc(a->l[0], b->l[0], 0, &carry);
  t.l[1] = adc(a->l[1], b->l[1], carry, &carry);
  t.l[2] = adc(a->l[2], b->l[2], carry, &carry);
  t.l[3] = adc(a->l[3], b->l[3], carry, &carry);
  fq_sub(r, &t, &FQ_MODULUS);
}

void fq_neg(fq_t *r, const fq_t *a) {
  uint64_t borrow, d0, d1, d2, d3;
  d0 = sbb(FQ_MODULUS.l[0], a->l[0], 0, &borrow);
  d1 = sbb(FQ_MODULUS.l[1], a->l[1], borrow, &borrow);
  d2 = sbb(FQ_MODULUS.l[2], a->l[2], borrow, &borrow);
  d3 = sbb(FQ_MODULUS.l[3], a->l[3], borrow, &borrow);
  uint64_t mask = (uint64_t)((a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0) - 1;
  r->l[0] = d0 & mask; r->l[1] = d1 & mask; r->l[2] = d2 & mask; r->l[3] = d3 & mask;
}

/* HAC 14.32, word by word                                     ristretto255.rs:642-686 */
void fq_montgomery_reduce(fq_t *out, const uint64_t t[8]) {
  uint64_t r0 = t[0], r1 = t[1], r2 = t[2], r3 = t[3], r4 = t[4], r5 = t[5], r6 = t[6], r7 = t[7];
  uint64_t k, carry, carry2;
  const uint64_t *m = FQ_MODULUS.l;

  k = r0 * FQ_INV;
  (void)mac(r0, k, m[0], 0, &carry);
  r1 = mac(r1, k, m[1], carry, &carry);
  r2 = mac(r2, k, m[2], carry, &carry);
  r3 = mac(r3, k, m[3], carry, &carry);
  r4 = adc(r4, 0, carry, &carry2);

  k = r1 * FQ_INV;
  (void)mac(r1, k, m[0], 0, &carry);
  r2 = mac(r2, k, m[1], carry, &carry);
  r3 = mac(r3, k, m[2], carry, &carry);
  r4 = mac(r4, k, m[3], carry, &carry);
  r5 = adc(r5, carry2, carry, &carry2);

  k = r2 * FQ_INV;
  (void)mac(r2, k, m[0], 0, &carry);
  r3 = mac(r3, k, m[1], carry, &carry);
  r4 = mac(r4, k, m[2], carry, &carry);
  r5 = mac(r5, k, m[3], carry, &carry);
  r6 = adc(r6, carry2, carry, &carry2);

  k = r3 * FQ_INV;
  (void)mac(r3, k, m[0], 0, &carry);
  r4 = mac(r4, k, m[1], carry, &carry);
  r5 = mac(r5, k, m[2], carry, &carry);
  r6 = mac(r6, k, m[3], carry, &carry);
  r7 = adc(r7, carry2, carry, &carry2);

  fq_t v = {{r4, r5, r6, r7}};
  fq_sub(out, &v, &FQ_MODULUS);
}

/* schoolbook 4x4 then reduce                                  ristretto255.rs:690-714 */
void fq_mul(fq_t *r, const fq_t *a, const fq_t *b) {
  uint64_t t[8] = {0}, carry;
  for (int i = 0; i < 4; i++) {
    carry = 0;
    for (int j = 0; j < 4; j++) t[i + j] = mac(t[i + j], a->l[i], b->l[j], carry, &carry);
    t[i + 4] = carry;
  }
  fq_montgomery_reduce(r, t);
}

void fq_square(fq_t *r, const fq_t *a) { fq_mul(r, a, a); }

int fq_is_zero(const fq_t *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
int fq_eq(const fq_t *a, const fq_t *b) { return memcmp(a, b, sizeof(fq_t)) == 0; }

void fq_from_u64(fq_t *r, uint64_t v) { /* From<u64>: Scalar([v,0,0,0]) * R2   ristretto255.rs:214-218 */
  fq_t t = {{v, 0, 0, 0}};
  fq_mul(r, &t, &FQ_R2);
}

int fq_from_bytes(fq_t *r, const uint8_t b[32]) {
  fq_t t; uint64_t borrow;
  memcpy(t.l, b, 32); /* little-endian host */
  (void)sbb(t.l[0], FQ_MODULUS.l[0], 0, &borrow);
  (void)sbb(t.l[1], FQ_MODULUS.l[1], borrow, &borrow);
  (void)sbb(t.l[2], FQ_MODULUS.l[2], borrow, &borrow);
  (void)sbb(t.l[3], FQ_MODULUS.l[3], borrow, &borrow);
  int is_some = (int)(borrow & 1);
  fq_mul(r, &t, &FQ_R2);
  return is_some;
}

void fq_to_bytes(uint8_t b[32], const fq_t *a) {
  uint64_t t[8] = {a->l[0], a->l[1], a->l[2], a->l[3], 0, 0, 0, 0};
  fq_t c; fq_montgomery_reduce(&c, t);
  memcpy(b, c.l, 32);
}

void fq_from_bytes_wide(fq_t *r, const uint8_t b[64]) { /* d0*R2 + d1*R3   ristretto255.rs:449-466 */
  fq_t d0, d1, x, y;
  memcpy(d0.l, b, 32); memcpy(d1.l, b + 32, 32);
  fq_mul(&x, &d0, &FQ_R2);
  fq_mul(&y, &d1, &FQ_R3);
  fq_add(r, &x, &y);
}

void fq_pow_vartime(fq_t *r, const fq_t *a, const uint64_t e[4]) {
  fq_t res = FQ_R;
  for (int w = 3; w >= 0; w--)
    for (int i = 63; i >= 0; i--) {
      fq_square(&res, &res);
      if ((e[w] >> i) & 1) fq_mul(&res, &res, a);
    }
  *r = res;
}

static void square_multiply(fq_t *y, int squarings, const fq_t *x) {
  for (int i = 0; i < squarings; i++) fq_square(y, y);
  fq_mul(y, y, x);
}

/* addition chain of ristretto255.rs:541-595 */
int fq_invert(fq_t *r, const fq_t *a) {
  fq_t _1 = *a, _10, _100, _11, _101, _111, _1001, _1011, _1111, y;
  fq_square(&_10, &_1);
  fq_square(&_100, &_10);
  fq_mul(&_11, &_10, &_1);
  fq_mul(&_101, &_10, &_11);
  fq_mul(&_111, &_10, &_101);
  fq_mul(&_1001, &_10, &_111);
  fq_mul(&_1011, &_10, &_1001);
  fq_mul(&_1111, &_100, &_1011);
  fq_mul(&y, &_1111, &_1);
  square_multiply(&y, 123 + 3, &_101);
  square_multiply(&y, 2 + 2, &_11);
  square_multiply(&y, 1 + 4, &_1111);
  square_multiply(&y, 1 + 4, &_1111);
  square_multiply(&y, 4, &_1001);
  square_multiply(&y, 2, &_11);
  square_multiply(&y, 1 + 4, &_1111);
  square_multiply(&y, 1 + 3, &_101);
  square_multiply(&y, 3 + 3, &_101);
  square_multiply(&y, 3, &_111);
  square_multiply(&y, 1 + 4, &_1111);
  square_multiply(&y, 2 + 3, &_111);
  square_multiply(&y, 2 + 2, &_11);
  square_multiply(&y, 1 + 4, &_1011);
  square_multiply(&y, 2 + 4, &_1011);
  square_multiply(&y, 6 + 4, &_1001);
  square_multiply(&y, 2 + 2, &_11);
  square_multiply(&y, 3 + 2, &_11);
  square_multiply(&y, 3 + 2, &_11);
  square_multiply(&y, 1 + 4, &_1001);
  square_multiply(&y, 1 + 3, &_111);
  square_multiply(&y, 2 + 4, &_1111);
  square_multiply(&y, 1 + 4, &_1011);
  square_multiply(&y, 3, &_101);
  square_multiply(&y, 2 + 4, &_1111);
  square_multiply(&y, 3, &_101);
  square_multiply(&y, 1 + 2, &_11);
  *r = y;
  return !fq_is_zero(a);
}

/* ristretto255.rs:597-639 */
void fq_batch_invert(fq_t *inputs, size_t n, fq_t *allinv) {
  fq_t *scratch = (fq_t *)malloc(sizeof(fq_t) * (n ? n : 1));
  fq_t acc = FQ_R, tmp;
  for (size_t i = 0; i < n; i++) { scratch[i] = acc; fq_mul(&acc, &acc, &inputs[i]); }
  fq_invert(&acc, &acc);
  if (allinv) *allinv = acc;
  for (size_t i = n; i-- > 0;) {
    fq_mul(&tmp, &acc, &inputs[i]);
    fq_mul(&inputs[i], &acc, &scratch[i]);
    acc = tmp;
  }
  free(scratch);
}
