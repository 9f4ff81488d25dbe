/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see fq.h).
 *
 * The reference's group is curve25519-dalek ^4.1.1 `RistrettoPoint` (third-party crate, NOT under
 * /root/reference; /root/reference/src/group.rs:6-7, Cargo.toml:14-18).  This file restates the
 * published algorithms: RFC 9496 (ristretto255 encode / decode / one-way map) over the twisted
 * Edwards curve -x^2 + y^2 = 1 + d x^2 y^2 in extended coordinates (Hisil-Wong-Carter-Dawson 2008),
 * field 2^255-19 in 5x51-bit limbs (the layout of dalek's u64 backend).
 * Pinned against RFC 9496 appendix vectors and libsodium (tests/test_oracle_group.py).
 */
#ifndef ORACLE_RISTRETTO_H
#define ORACLE_RISTRETTO_H
#include <stdint.h>
#include <stddef.h>
#include "fq.h"

typedef struct { uint64_t v[5]; } fe_t;
typedef struct { fe_t X, Y, Z, T; } ge_t; /* 160 bytes */

void ge_identity(ge_t *r);
void ge_add(ge_t *r, const ge_t *p, const ge_t *q);
void ge_sub(ge_t *r, const ge_t *p, const ge_t *q);
void ge_neg(ge_t *r, const ge_t *p);
void ge_double(ge_t *r, const ge_t *p);
int  ge_eq(const ge_t *p, const ge_t *q); /* ristretto (quotient group) equality */

int  ristretto_decode(ge_t *r, const uint8_t s[32]);      /* 1 = ok, 0 = invalid encoding */
void ristretto_encode(uint8_t s[32], const ge_t *p);
void ristretto_from_uniform_bytes(ge_t *r, const uint8_t b[64]);
void ristretto_encode_batch(uint8_t *out, const ge_t *p, size_t n);

/* scalar given as canonical little-endian bytes (what Scalar::to_bytes yields) */
void ge_scalarmul_bytes(ge_t *r, const uint8_t k[32], const ge_t *p);
/* scalars given as Montgomery-form F_q limbs, converted first like group.rs:110-113 */
void ge_scalarmul(ge_t *r, const fq_t *k, const ge_t *p);
void ge_msm(ge_t *r, const fq_t *scalars, const ge_t *points, size_t n);

extern const uint8_t RISTRETTO_BASEPOINT_COMPRESSED[32];
#endif
