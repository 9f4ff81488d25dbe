/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this code.  The product (spartan_b200/) never links or calls it.
 *
 * CPU restatement of the scalar field F_q of microsoft/Spartan:
 *   /root/reference/src/scalar/ristretto255.rs
 * q = 2^252 + 27742317777372353535851937790883648493, 4x64-bit little-endian limbs,
 * values always in Montgomery form (R = 2^256)              (ristretto255.rs:195-199).
 */
#ifndef ORACLE_FQ_H
#define ORACLE_FQ_H
#include <stdint.h>
#include <stddef.h>

typedef struct { uint64_t l[4]; } fq_t;

extern const fq_t FQ_MODULUS, FQ_R, FQ_R2, FQ_R3;
#define FQ_INV 0xd2b51da312547e1bULL

void fq_add(fq_t *r, const fq_t *a, const fq_t *b);
void fq_sub(fq_t *r, const fq_t *a, const fq_t *b);
void fq_neg(fq_t *r, const fq_t *a);
void fq_mul(fq_t *r, const fq_t *a, const fq_t *b);
void fq_square(fq_t *r, const fq_t *a);
void fq_montgomery_reduce(fq_t *r, const uint64_t t[8]);
int  fq_invert(fq_t *r, const fq_t *a);             /* returns 0 when a == 0 */
int  fq_from_bytes(fq_t *r, const uint8_t b[32]);   /* returns 0 when non-canonical */
void fq_to_bytes(uint8_t b[32], const fq_t *a);
void fq_from_bytes_wide(fq_t *r, const uint8_t b[64]);
void fq_from_u64(fq_t *r, uint64_t v);
int  fq_is_zero(const fq_t *a);
int  fq_eq(const fq_t *a, const fq_t *b);
void fq_pow_vartime(fq_t *r, const fq_t *a, const uint64_t e[4]);
void fq_batch_invert(fq_t *inputs, size_t n, fq_t *allinv);

#endif
