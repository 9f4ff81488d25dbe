// spartan_b200 — host-side edwards25519 / ristretto255 arithmetic in radix 2^51 (five 64-bit limbs, 128-bit products, lazy carries).
//
// The prover's host thread sits on the critical path between kernels: every sumcheck round and every inner-product round ends with a few
// small Pedersen commitments (sumcheck.rs:497-503,688-694; nizk/mod.rs:311-370; bullet.rs:83-97) and their ristretto encodings, which the
// transcript must absorb before the next challenge exists.  Those are single points, far too small for a kernel launch, so they are computed
// here from fixed-base tables copied off the device.  The device code keeps its 8x32-bit saturated limbs (field.cuh); this file is the same
// mathematics in the representation x86-64 is fastest at.  Encodings follow RFC 9496 and are bit-identical to the device's and to dalek's.
#pragma once
#include <cstdint>
#include <cstring>
#include "curve.cuh"

namespace sp {

struct fe51 { uint64_t v[5]; };
typedef unsigned __int128 u128_t;
static const uint64_t FE_M51 = (1ULL << 51) - 1;

inline fe51 fe_zero() { fe51 r = {{0, 0, 0, 0, 0}}; return r; }
inline fe51 fe_one() { fe51 r = {{1, 0, 0, 0, 0}}; return r; }
inline fe51 fe_from_u256(const u256& a) {   // any 256-bit value (not necessarily < p); limb 4 may hold 52 bits
  uint64_t x0 = (uint64_t)a.v[0] | ((uint64_t)a.v[1] << 32), x1 = (uint64_t)a.v[2] | ((uint64_t)a.v[3] << 32);
  uint64_t x2 = (uint64_t)a.v[4] | ((uint64_t)a.v[5] << 32), x3 = (uint64_t)a.v[6] | ((uint64_t)a.v[7] << 32);
  fe51 r;
  r.v[0] = x0 & FE_M51;
  r.v[1] = ((x0 >> 51) | (x1 << 13)) & FE_M51;
  r.v[2] = ((x1 >> 38) | (x2 << 26)) & FE_M51;
  r.v[3] = ((x2 >> 25) | (x3 << 39)) & FE_M51;
  r.v[4] = x3 >> 12;
  return r;
}
inline void fe_carry(fe51& a) {   // limbs -> < 2^51 + small
  uint64_t c;
  c = a.v[0] >> 51; a.v[0] &= FE_M51; a.v[1] += c;
  c = a.v[1] >> 51; a.v[1] &= FE_M51; a.v[2] += c;
  c = a.v[2] >> 51; a.v[2] &= FE_M51; a.v[3] += c;
  c = a.v[3] >> 51; a.v[3] &= FE_M51; a.v[4] += c;
  c = a.v[4] >> 51; a.v[4] &= FE_M51; a.v[0] += 19 * c;
}
inline fe51 fe_freeze(fe51 a) {   // the canonical representative in [0, p)
  fe_carry(a); fe_carry(a);
  uint64_t q = (a.v[0] + 19) >> 51;
  q = (a.v[1] + q) >> 51; q = (a.v[2] + q) >> 51; q = (a.v[3] + q) >> 51; q = (a.v[4] + q) >> 51;
  a.v[0] += 19 * q;
  uint64_t c;
  c = a.v[0] >> 51; a.v[0] &= FE_M51; a.v[1] += c;
  c = a.v[1] >> 51; a.v[1] &= FE_M51; a.v[2] += c;
  c = a.v[2] >> 51; a.v[2] &= FE_M51; a.v[3] += c;
  c = a.v[3] >> 51; a.v[3] &= FE_M51; a.v[4] += c;
  a.v[4] &= FE_M51;
  return a;
}
inline void fe_to_bytes(uint8_t out[32], const fe51& x) {
  fe51 a = fe_freeze(x);
  uint64_t w[4];
  w[0] = a.v[0] | (a.v[1] << 51);
  w[1] = (a.v[1] >> 13) | (a.v[2] << 38);
  w[2] = (a.v[2] >> 26) | (a.v[3] << 25);
  w[3] = (a.v[3] >> 39) | (a.v[4] << 12);
  memcpy(out, w, 32);   // little-endian host
}
inline fe51 fe_add(const fe51& a, const fe51& b) { fe51 r; for (int i = 0; i < 5; i++) r.v[i] = a.v[i] + b.v[i]; return r; }
inline fe51 fe_sub(const fe51& a, const fe51& b) {   // a + 4p - b, carried; needs b's limbs < 2^53
  fe51 r;
  r.v[0] = a.v[0] + 0x1fffffffffffb4ULL - b.v[0];
  r.v[1] = a.v[1] + 0x1ffffffffffffcULL - b.v[1];
  r.v[2] = a.v[2] + 0x1ffffffffffffcULL - b.v[2];
  r.v[3] = a.v[3] + 0x1ffffffffffffcULL - b.v[3];
  r.v[4] = a.v[4] + 0x1ffffffffffffcULL - b.v[4];
  fe_carry(r);
  return r;
}
inline fe51 fe_neg(const fe51& a) { return fe_sub(fe_zero(), a); }
inline fe51 fe_mul(const fe51& a, const fe51& b) {   // limbs of a, b < 2^55
  const uint64_t a0 = a.v[0], a1 = a.v[1], a2 = a.v[2], a3 = a.v[3], a4 = a.v[4];
  const uint64_t b0 = b.v[0], b1 = b.v[1], b2 = b.v[2], b3 = b.v[3], b4 = b.v[4];
  const uint64_t b1_19 = 19 * b1, b2_19 = 19 * b2, b3_19 = 19 * b3, b4_19 = 19 * b4;
  u128_t r0 = (u128_t)a0 * b0 + (u128_t)a1 * b4_19 + (u128_t)a2 * b3_19 + (u128_t)a3 * b2_19 + (u128_t)a4 * b1_19;
  u128_t r1 = (u128_t)a0 * b1 + (u128_t)a1 * b0 + (u128_t)a2 * b4_19 + (u128_t)a3 * b3_19 + (u128_t)a4 * b2_19;
  u128_t r2 = (u128_t)a0 * b2 + (u128_t)a1 * b1 + (u128_t)a2 * b0 + (u128_t)a3 * b4_19 + (u128_t)a4 * b3_19;
  u128_t r3 = (u128_t)a0 * b3 + (u128_t)a1 * b2 + (u128_t)a2 * b1 + (u128_t)a3 * b0 + (u128_t)a4 * b4_19;
  u128_t r4 = (u128_t)a0 * b4 + (u128_t)a1 * b3 + (u128_t)a2 * b2 + (u128_t)a3 * b1 + (u128_t)a4 * b0;
  fe51 r;
  r1 += (uint64_t)(r0 >> 51); r.v[0] = (uint64_t)r0 & FE_M51;
  r2 += (uint64_t)(r1 >> 51); r.v[1] = (uint64_t)r1 & FE_M51;
  r3 += (uint64_t)(r2 >> 51); r.v[2] = (uint64_t)r2 & FE_M51;
  r4 += (uint64_t)(r3 >> 51); r.v[3] = (uint64_t)r3 & FE_M51;
  uint64_t c = (uint64_t)(r4 >> 51); r.v[4] = (uint64_t)r4 & FE_M51;
  r.v[0] += 19 * c;
  c = r.v[0] >> 51; r.v[0] &= FE_M51; r.v[1] += c;
  return r;
}
inline fe51 fe_sqr(const fe51& a) {
  const uint64_t a0 = a.v[0], a1 = a.v[1], a2 = a.v[2], a3 = a.v[3], a4 = a.v[4];
  const uint64_t d0 = 2 * a0, d1 = 2 * a1, d2 = 2 * a2, a3_19 = 19 * a3, a4_19 = 19 * a4, d3_19 = 2 * a3_19;
  u128_t r0 = (u128_t)a0 * a0 + (u128_t)d1 * a4_19 + (u128_t)d2 * a3_19;
  u128_t r1 = (u128_t)d0 * a1 + (u128_t)d2 * a4_19 + (u128_t)a3 * a3_19;
  u128_t r2 = (u128_t)d0 * a2 + (u128_t)a1 * a1 + (u128_t)d3_19 * a4;
  u128_t r3 = (u128_t)d0 * a3 + (u128_t)d1 * a2 + (u128_t)a4 * a4_19;
  u128_t r4 = (u128_t)d0 * a4 + (u128_t)d1 * a3 + (u128_t)a2 * a2;
  fe51 r;
  r1 += (uint64_t)(r0 >> 51); r.v[0] = (uint64_t)r0 & FE_M51;
  r2 += (uint64_t)(r1 >> 51); r.v[1] = (uint64_t)r1 & FE_M51;
  r3 += (uint64_t)(r2 >> 51); r.v[2] = (uint64_t)r2 & FE_M51;
  r4 += (uint64_t)(r3 >> 51); r.v[3] = (uint64_t)r3 & FE_M51;
  uint64_t c = (uint64_t)(r4 >> 51); r.v[4] = (uint64_t)r4 & FE_M51;
  r.v[0] += 19 * c;
  c = r.v[0] >> 51; r.v[0] &= FE_M51; r.v[1] += c;
  return r;
}
inline bool fe_is_neg(const fe51& a) { return (fe_freeze(a).v[0] & 1) != 0; }
inline bool fe_eq(const fe51& a, const fe51& b) {
  fe51 x = fe_freeze(a), y = fe_freeze(b);
  return x.v[0] == y.v[0] && x.v[1] == y.v[1] && x.v[2] == y.v[2] && x.v[3] == y.v[3] && x.v[4] == y.v[4];
}
inline fe51 fe_abs(const fe51& a) { return fe_is_neg(a) ? fe_neg(a) : a; }

struct FeConsts { fe51 d2, sqrt_m1, invsqrt_a_minus_d; };
inline const FeConsts& fe_consts() {
  static const FeConsts c = {fe_from_u256(fp_2D()), fe_from_u256(fp_SQRT_M1()), fe_from_u256(fp_INVSQRT_A_MINUS_D())};
  return c;
}

// z^(2^252-3) for N independent inputs in lockstep (N = 1 or 2: two chains overlap in the CPU's pipelines)
template <int N>
inline void fe_pow22523_n(const fe51* z, fe51* out) {
  fe51 z2[N], z9[N], z11[N], t[N], x5[N], x10[N], x20[N], x40[N], x50[N], x100[N], x200[N];
  auto sqn = [](fe51* a, int n) { for (int i = 0; i < n; i++) for (int k = 0; k < N; k++) a[k] = fe_sqr(a[k]); };
  for (int k = 0; k < N; k++) { z2[k] = fe_sqr(z[k]); t[k] = z2[k]; }
  sqn(t, 2);
  for (int k = 0; k < N; k++) { z9[k] = fe_mul(t[k], z[k]); z11[k] = fe_mul(z9[k], z2[k]); x5[k] = fe_mul(fe_sqr(z11[k]), z9[k]); t[k] = x5[k]; }
  sqn(t, 5);   for (int k = 0; k < N; k++) { x10[k] = fe_mul(t[k], x5[k]); t[k] = x10[k]; }
  sqn(t, 10);  for (int k = 0; k < N; k++) { x20[k] = fe_mul(t[k], x10[k]); t[k] = x20[k]; }
  sqn(t, 20);  for (int k = 0; k < N; k++) { x40[k] = fe_mul(t[k], x20[k]); t[k] = x40[k]; }
  sqn(t, 10);  for (int k = 0; k < N; k++) { x50[k] = fe_mul(t[k], x10[k]); t[k] = x50[k]; }
  sqn(t, 50);  for (int k = 0; k < N; k++) { x100[k] = fe_mul(t[k], x50[k]); t[k] = x100[k]; }
  sqn(t, 100); for (int k = 0; k < N; k++) { x200[k] = fe_mul(t[k], x100[k]); t[k] = x200[k]; }
  sqn(t, 50);  for (int k = 0; k < N; k++) t[k] = fe_mul(t[k], x50[k]);     // 2^250 - 1
  sqn(t, 2);
  for (int k = 0; k < N; k++) out[k] = fe_mul(t[k], z[k]);
}

// ---- points
struct hge { fe51 X, Y, Z, T; };            // extended coordinates, a = -1
struct hniels { fe51 ypx, ymx, t2d; };      // affine (y+x, y-x, 2dxy)
inline hge hge_identity() { hge r; r.X = fe_zero(); r.Y = fe_one(); r.Z = fe_one(); r.T = fe_zero(); return r; }
inline hge to_hge(const ge& p) { hge r; r.X = fe_from_u256(p.X); r.Y = fe_from_u256(p.Y); r.Z = fe_from_u256(p.Z); r.T = fe_from_u256(p.T); return r; }
inline hniels to_hniels(const ge_niels& q) { hniels r; r.ypx = fe_from_u256(q.ypx); r.ymx = fe_from_u256(q.ymx); r.t2d = fe_from_u256(q.t2d); return r; }
inline hge hge_add(const hge& p, const hge& q) {   // add-2008-hwcd-3
  fe51 A = fe_mul(fe_sub(p.Y, p.X), fe_sub(q.Y, q.X));
  fe51 B = fe_mul(fe_add(p.Y, p.X), fe_add(q.Y, q.X));
  fe51 C = fe_mul(fe_mul(p.T, q.T), fe_consts().d2);
  fe51 D = fe_mul(p.Z, q.Z);
  D = fe_add(D, D);
  fe51 E = fe_sub(B, A), F = fe_sub(D, C), G = fe_add(D, C), H = fe_add(B, A);
  hge r;
  r.X = fe_mul(E, F); r.Y = fe_mul(G, H); r.T = fe_mul(E, H); r.Z = fe_mul(F, G);
  return r;
}
inline hge hge_madd(const hge& p, const hniels& q, bool neg) {   // 7M mixed addition; `neg` adds -q
  const fe51& qa = neg ? q.ypx : q.ymx;
  const fe51& qb = neg ? q.ymx : q.ypx;
  fe51 A = fe_mul(fe_sub(p.Y, p.X), qa);
  fe51 B = fe_mul(fe_add(p.Y, p.X), qb);
  fe51 C = fe_mul(p.T, q.t2d);
  fe51 D = fe_add(p.Z, p.Z);
  fe51 E = fe_sub(B, A), H = fe_add(B, A);
  fe51 F = neg ? fe_add(D, C) : fe_sub(D, C);
  fe51 G = neg ? fe_sub(D, C) : fe_add(D, C);
  hge r;
  r.X = fe_mul(E, F); r.Y = fe_mul(G, H); r.T = fe_mul(E, H); r.Z = fe_mul(F, G);
  return r;
}
inline hge hge_dbl(const hge& p) {   // dbl-2008-hwcd
  fe51 A = fe_sqr(p.X), B = fe_sqr(p.Y);
  fe51 C = fe_sqr(p.Z);
  C = fe_add(C, C);
  fe51 D = fe_neg(A);
  fe51 t = fe_add(p.X, p.Y);
  fe51 E = fe_sub(fe_sub(fe_sqr(t), A), B);
  fe51 G = fe_add(D, B), F = fe_sub(G, C), H = fe_sub(D, B);
  hge r;
  r.X = fe_mul(E, F); r.Y = fe_mul(G, H); r.T = fe_mul(E, H); r.Z = fe_mul(F, G);
  return r;
}
// k*P, k = canonical little-endian 256-bit integer (not Montgomery); signed 4-bit windows over {1..8}*P, vartime
inline hge hge_scalarmul(const u256& k, const hge& p) {
  hge tab[8];
  tab[0] = p;
  for (int i = 1; i < 8; i++) tab[i] = hge_add(tab[i - 1], p);
  int dig[65];
  int carry = 0;
  for (int i = 0; i < 64; i++) {
    int v = (int)((k.v[i >> 3] >> (4 * (i & 7))) & 15u) + carry;
    if (v > 8) { dig[i] = v - 16; carry = 1; } else { dig[i] = v; carry = 0; }
  }
  dig[64] = carry;
  hge acc = hge_identity();
  bool started = false;
  for (int i = 64; i >= 0; i--) {
    if (started) { acc = hge_dbl(acc); acc = hge_dbl(acc); acc = hge_dbl(acc); acc = hge_dbl(acc); }
    int d = dig[i];
    if (d > 0) { acc = hge_add(acc, tab[d - 1]); started = true; }
    else if (d < 0) { hge q = tab[-d - 1]; q.X = fe_neg(q.X); q.T = fe_neg(q.T); acc = hge_add(acc, q); started = true; }
  }
  return acc;
}

// RFC 9496 4.3.2 Encode for N points at once (the N exponentiations run in lockstep)
template <int N>
inline void hge_encode_n(const hge* p, uint8_t (*out)[32]) {
  const FeConsts& K = fe_consts();
  fe51 u1[N], u2[N], v[N], v3[N], base[N], pw[N];
  for (int k = 0; k < N; k++) {
    u1[k] = fe_mul(fe_add(p[k].Z, p[k].Y), fe_sub(p[k].Z, p[k].Y));
    u2[k] = fe_mul(p[k].X, p[k].Y);
    v[k] = fe_mul(u1[k], fe_sqr(u2[k]));
    // SQRT_RATIO_M1(1, v) (RFC 9496 4.2): r = v^3 * (v^7)^((p-5)/8)
    v3[k] = fe_mul(fe_sqr(v[k]), v[k]);
    base[k] = fe_mul(fe_sqr(v3[k]), v[k]);
  }
  fe_pow22523_n<N>(base, pw);
  for (int k = 0; k < N; k++) {
    fe51 r = fe_mul(v3[k], pw[k]);
    fe51 check = fe_mul(v[k], fe_sqr(r));
    fe51 one = fe_one(), neg_one = fe_neg(one);
    bool flipped = fe_eq(check, neg_one);
    bool flipped_i = fe_eq(check, fe_mul(neg_one, K.sqrt_m1));
    if (flipped || flipped_i) r = fe_mul(r, K.sqrt_m1);
    fe51 invsqrt = fe_abs(r);
    fe51 den1 = fe_mul(invsqrt, u1[k]), den2 = fe_mul(invsqrt, u2[k]);
    fe51 z_inv = fe_mul(fe_mul(den1, den2), p[k].T);
    fe51 ix0 = fe_mul(p[k].X, K.sqrt_m1), iy0 = fe_mul(p[k].Y, K.sqrt_m1);
    fe51 ench = fe_mul(den1, K.invsqrt_a_minus_d);
    bool rotate = fe_is_neg(fe_mul(p[k].T, z_inv));
    fe51 x = rotate ? iy0 : p[k].X;
    fe51 y = rotate ? ix0 : p[k].Y;
    fe51 den_inv = rotate ? ench : den2;
    if (fe_is_neg(fe_mul(x, z_inv))) y = fe_neg(y);
    fe51 s = fe_abs(fe_mul(den_inv, fe_sub(p[k].Z, y)));
    fe_to_bytes(out[k], s);
  }
}

}  // namespace sp
