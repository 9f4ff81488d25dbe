/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see fq.h).
 * O(n) loops of the reference's prover restated over arrays of Montgomery-form F_q limbs
 * (n x 4 x u64, exactly the memory layout of a Rust `&[Scalar]`).  Each function names the
 * reference loop it follows.  OpenMP is only used to let the timed CPU baseline use every host
 * thread (the reference itself is single-threaded apart from dense_mlpoly.rs:148-162); field
 * arithmetic is exact so any summation order gives the same canonical result.
 */
#include "fq.h"
#include "ristretto.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void oracle_set_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* DensePolynomial::bound_poly_var_top, dense_mlpoly.rs:215-223.  Z[i] += r*(Z[i+n]-Z[i]) */
void poly_bound_top(fq_t *Z, size_t len, const fq_t *r) {
  size_t n = len / 2;
#pragma omp parallel for schedule(static) if (n > 4096)
  for (size_t i = 0; i < n; i++) {
    fq_t d, m;
    fq_sub(&d, &Z[i + n], &Z[i]);
    fq_mul(&m, r, &d);
    fq_add(&Z[i], &Z[i], &m);
  }
}
/* DensePolynomial::bound_poly_var_bot, dense_mlpoly.rs:225-233 */
void poly_bound_bot(fq_t *Z, size_t len, const fq_t *r) {
  size_t n = len / 2;
  for (size_t i = 0; i < n; i++) {
    fq_t d, m;
    fq_sub(&d, &Z[2 * i + 1], &Z[2 * i]);
    fq_mul(&m, r, &d);
    fq_add(&Z[i], &Z[2 * i], &m);
  }
}

/* EqPolynomial::evals, dense_mlpoly.rs:68-84 */
void poly_eq_evals(fq_t *evals, const fq_t *r, size_t ell) {
  size_t size = 1;
  evals[0] = FQ_R;
  for (size_t j = 0; j < ell; j++) {
    size *= 2;
    if (size <= 8192) {   /* the reference's in-place descending loop */
      for (size_t i = size - 1;; i -= 2) {
        fq_t scalar = evals[i / 2];
        fq_mul(&evals[i], &scalar, &r[j]);
        fq_sub(&evals[i - 1], &scalar, &evals[i]);
        if (i == 1) break;
      }
    } else {
      /* same values, all host threads: pair p reads evals[p] and writes evals[2p], evals[2p+1]; only the upper half (2p >= size/2) can be
         written in place without clobbering an unread parent, so the parents are snapshotted first */
      size_t half = size / 2;
      fq_t *parent = (fq_t *)malloc(sizeof(fq_t) * half);
      memcpy(parent, evals, sizeof(fq_t) * half);
#pragma omp parallel for schedule(static)
      for (size_t p = 0; p < half; p++) {
        fq_mul(&evals[2 * p + 1], &parent[p], &r[j]);
        fq_sub(&evals[2 * p], &parent[p], &evals[2 * p + 1]);
      }
      free(parent);
    }
  }
}

/* DotProductProofLog::compute_dotproduct, nizk/mod.rs:435-438; bullet.rs:233-243 */
void poly_dot(fq_t *out, const fq_t *a, const fq_t *b, size_t n) {
  fq_t acc = {{0, 0, 0, 0}};
#pragma omp parallel if (n > 4096)
  {
    fq_t part = {{0, 0, 0, 0}}, m;
#pragma omp for schedule(static) nowait
    for (size_t i = 0; i < n; i++) { fq_mul(&m, &a[i], &b[i]); fq_add(&part, &part, &m); }
#pragma omp critical
    fq_add(&acc, &acc, &part);
  }
  *out = acc;
}

/* DensePolynomial::bound, dense_mlpoly.rs:206-213: out[i] = sum_j L[j] * Z[j*R + i] */
void poly_bound_rows(fq_t *out, const fq_t *Z, const fq_t *L, size_t L_size, size_t R_size) {
#pragma omp parallel for schedule(static) if (L_size * R_size > 4096)
  for (size_t i = 0; i < R_size; i++) {
    fq_t acc = {{0, 0, 0, 0}}, m;
    for (size_t j = 0; j < L_size; j++) { fq_mul(&m, &L[j], &Z[j * R_size + i]); fq_add(&acc, &acc, &m); }
    out[i] = acc;
  }
}

/* ---- sumcheck round evaluations: t = 0, 2, (3) using low/high halves ---- */
/* prove_quad loop, sumcheck.rs:460-469, comb = A*B (r1csproof.rs:122-123).  out = [e0, e2] */
void sc_eval_quad(fq_t out[2], const fq_t *A, const fq_t *B, size_t len) {
  size_t n = len / 2;
  fq_t e0 = {{0}}, e2 = {{0}};
#pragma omp parallel if (n > 4096)
  {
    fq_t p0 = {{0}}, p2 = {{0}}, a2, b2, m;
#pragma omp for schedule(static) nowait
    for (size_t i = 0; i < n; i++) {
      fq_mul(&m, &A[i], &B[i]); fq_add(&p0, &p0, &m);
      fq_add(&a2, &A[n + i], &A[n + i]); fq_sub(&a2, &a2, &A[i]);
      fq_add(&b2, &B[n + i], &B[n + i]); fq_sub(&b2, &b2, &B[i]);
      fq_mul(&m, &a2, &b2); fq_add(&p2, &p2, &m);
    }
#pragma omp critical
    { fq_add(&e0, &e0, &p0); fq_add(&e2, &e2, &p2); }
  }
  out[0] = e0; out[1] = e2;
}

/* which: 0 -> A*B*C (product_tree.rs:283-286 via sumcheck.rs:204-228 / :296-320 / :334-355)
 *        1 -> A*(B*C - D) (r1csproof.rs:87-91 via sumcheck.rs:625-652).  out = [e0, e2, e3] */
static inline void comb3(fq_t *r, const fq_t *a, const fq_t *b, const fq_t *c) { fq_t t; fq_mul(&t, a, b); fq_mul(r, &t, c); }
static inline void comb4(fq_t *r, const fq_t *a, const fq_t *b, const fq_t *c, const fq_t *d) {
  fq_t t; fq_mul(&t, b, c); fq_sub(&t, &t, d); fq_mul(r, a, &t);
}
void sc_eval_cubic(fq_t out[3], const fq_t *A, const fq_t *B, const fq_t *C, const fq_t *D, size_t len) {
  size_t n = len / 2;
  fq_t e0 = {{0}}, e2 = {{0}}, e3 = {{0}};
#pragma omp parallel if (n > 4096)
  {
    fq_t p0 = {{0}}, p2 = {{0}}, p3 = {{0}}, a, b, c, d = {{0}}, m;
#pragma omp for schedule(static) nowait
    for (size_t i = 0; i < n; i++) {
      if (D) comb4(&m, &A[i], &B[i], &C[i], &D[i]); else comb3(&m, &A[i], &B[i], &C[i]);
      fq_add(&p0, &p0, &m);
      fq_add(&a, &A[n + i], &A[n + i]); fq_sub(&a, &a, &A[i]);
      fq_add(&b, &B[n + i], &B[n + i]); fq_sub(&b, &b, &B[i]);
      fq_add(&c, &C[n + i], &C[n + i]); fq_sub(&c, &c, &C[i]);
      if (D) { fq_add(&d, &D[n + i], &D[n + i]); fq_sub(&d, &d, &D[i]); comb4(&m, &a, &b, &c, &d); }
      else comb3(&m, &a, &b, &c);
      fq_add(&p2, &p2, &m);
      fq_add(&a, &a, &A[n + i]); fq_sub(&a, &a, &A[i]);
      fq_add(&b, &b, &B[n + i]); fq_sub(&b, &b, &B[i]);
      fq_add(&c, &c, &C[n + i]); fq_sub(&c, &c, &C[i]);
      if (D) { fq_add(&d, &d, &D[n + i]); fq_sub(&d, &d, &D[i]); comb4(&m, &a, &b, &c, &d); }
      else comb3(&m, &a, &b, &c);
      fq_add(&p3, &p3, &m);
    }
#pragma omp critical
    { fq_add(&e0, &e0, &p0); fq_add(&e2, &e2, &p2); fq_add(&e3, &e3, &p3); }
  }
  out[0] = e0; out[1] = e2; out[2] = e3;
}

/* ---- sparse matrix ops on COO triples, sparse_mlpoly.rs:454-481 ---- */
void sparse_multiply_vec(fq_t *Mz, size_t num_rows, const uint64_t *row, const uint64_t *col, const fq_t *val, size_t nnz, const fq_t *z) {
  memset(Mz, 0, sizeof(fq_t) * num_rows);
  /* products on all host threads, the scatter-add (order-independent: exact field arithmetic) on one */
  fq_t *prod = (fq_t *)malloc(sizeof(fq_t) * (nnz ? nnz : 1));
#pragma omp parallel for schedule(static) if (nnz > 4096)
  for (size_t k = 0; k < nnz; k++) fq_mul(&prod[k], &val[k], &z[col[k]]);
  for (size_t k = 0; k < nnz; k++) fq_add(&Mz[row[k]], &Mz[row[k]], &prod[k]);
  free(prod);
}
void sparse_eval_table(fq_t *out, size_t num_cols, const uint64_t *row, const uint64_t *col, const fq_t *val, size_t nnz, const fq_t *rx) {
  memset(out, 0, sizeof(fq_t) * num_cols);
  fq_t *prod = (fq_t *)malloc(sizeof(fq_t) * (nnz ? nnz : 1));
#pragma omp parallel for schedule(static) if (nnz > 4096)
  for (size_t k = 0; k < nnz; k++) fq_mul(&prod[k], &rx[row[k]], &val[k]);
  for (size_t k = 0; k < nnz; k++) fq_add(&out[col[k]], &out[col[k]], &prod[k]);
  free(prod);
}
/* evaluate_with_tables, sparse_mlpoly.rs:426-438 */
void sparse_evaluate(fq_t *out, const uint64_t *row, const uint64_t *col, const fq_t *val, size_t nnz, const fq_t *trx, const fq_t *try_) {
  fq_t acc = {{0}};
#pragma omp parallel if (nnz > 4096)
  {
    fq_t local = {{0}}, m;
#pragma omp for schedule(static) nowait
    for (size_t k = 0; k < nnz; k++) { fq_mul(&m, &trx[row[k]], &try_[col[k]]); fq_mul(&m, &m, &val[k]); fq_add(&local, &local, &m); }
#pragma omp critical
    fq_add(&acc, &acc, &local);
  }
  *out = acc;
}

/* ---- SPARK helpers ---- */
/* AddrTimestamps::deref_mem, sparse_mlpoly.rs:256-265 */
void spark_deref(fq_t *out, const uint64_t *addr, size_t n, const fq_t *mem) {
#pragma omp parallel for schedule(static) if (n > 4096)
  for (size_t i = 0; i < n; i++) out[i] = mem[addr[i]];
}
/* DensePolynomial::from_usize, dense_mlpoly.rs:274-280 */
void poly_from_u64(fq_t *out, const uint64_t *v, size_t n) {
#pragma omp parallel for schedule(static) if (n > 4096)
  for (size_t i = 0; i < n; i++) fq_from_u64(&out[i], v[i]);
}
/* hash(addr,val,ts) - r_multiset = ts*r^2 + val*r + addr - r_multiset, sparse_mlpoly.rs:545-549.
 * addr_u64 != NULL: addr = Scalar::from(addr_u64[i]); ts_plus_one adds Scalar::one() to ts (write set :596) */
void spark_hash_layer(fq_t *out, size_t n, const fq_t *addr, const uint64_t *addr_identity_base, const fq_t *val, const fq_t *ts,
                      int ts_plus_one, const fq_t *r_hash, const fq_t *r_multiset) {
  fq_t r2; fq_mul(&r2, r_hash, r_hash);
#pragma omp parallel for schedule(static) if (n > 4096)
  for (size_t i = 0; i < n; i++) {
    fq_t a, t = {{0}}, h, m;
    if (addr) a = addr[i]; else fq_from_u64(&a, (uint64_t)i + (addr_identity_base ? *addr_identity_base : 0));
    if (ts) t = ts[i];
    if (ts_plus_one) fq_add(&t, &t, &FQ_R);
    fq_mul(&h, &t, &r2);
    fq_mul(&m, &val[i], r_hash);
    fq_add(&h, &h, &m);
    fq_add(&h, &h, &a);
    fq_sub(&out[i], &h, r_multiset);
  }
}
/* ProductCircuit::compute_layer, product_tree.rs:18-34: out[i] = left[i]*right[i], i < n */
void poly_hadamard(fq_t *out, const fq_t *a, const fq_t *b, size_t n) {
#pragma omp parallel for schedule(static) if (n > 4096)
  for (size_t i = 0; i < n; i++) fq_mul(&out[i], &a[i], &b[i]);
}
/* sum_i a[i]*b[i]*c[i], DotProductCircuit::evaluate product_tree.rs:83-87 */
void poly_dot3(fq_t *out, const fq_t *a, const fq_t *b, const fq_t *c, size_t n) {
  fq_t acc = {{0}}, m;
  for (size_t i = 0; i < n; i++) { fq_mul(&m, &a[i], &b[i]); fq_mul(&m, &m, &c[i]); fq_add(&acc, &acc, &m); }
  *out = acc;
}
/* out[i] = ra*A[i] + rb*B[i] + rc*C[i], r1csproof.rs:279-282 */
void poly_lincomb3(fq_t *out, const fq_t *A, const fq_t *B, const fq_t *C, const fq_t *ra, const fq_t *rb, const fq_t *rc, size_t n) {
#pragma omp parallel for schedule(static) if (n > 4096)
  for (size_t i = 0; i < n; i++) {
    fq_t x, y;
    fq_mul(&x, ra, &A[i]); fq_mul(&y, rb, &B[i]); fq_add(&x, &x, &y);
    fq_mul(&y, rc, &C[i]); fq_add(&out[i], &x, &y);
  }
}

/* ---- batch conversions ---- */
void fq_from_bytes_wide_batch(fq_t *out, const uint8_t *in, size_t n) {
#pragma omp parallel for schedule(static) if (n > 4096)
  for (size_t i = 0; i < n; i++) fq_from_bytes_wide(&out[i], in + 64 * i);
}
void fq_to_bytes_batch(uint8_t *out, const fq_t *in, size_t n) {
  for (size_t i = 0; i < n; i++) fq_to_bytes(out + 32 * i, &in[i]);
}
int fq_from_bytes_batch(fq_t *out, const uint8_t *in, size_t n) {
  int ok = 1;
  for (size_t i = 0; i < n; i++) ok &= fq_from_bytes(&out[i], in + 32 * i);
  return ok;
}

/* ---- Pedersen commitments ---- */
/* MultiCommitGens::new, commitments.rs:15-33: uniform = first 64*(n+1) bytes of SHAKE256(label || basepoint) */
void gens_from_uniform(ge_t *out, const uint8_t *uniform, size_t count) {
#pragma omp parallel for schedule(static) if (count > 16)
  for (size_t i = 0; i < count; i++) ristretto_from_uniform_bytes(&out[i], uniform + 64 * i);
}
/* DensePolynomial::commit_inner, dense_mlpoly.rs:148-177: C_i = (MSM(Z[iR..(i+1)R], G) + blinds[i]*h).compress() */
void poly_commit_rows(uint8_t *out32, const fq_t *Z, size_t L_size, size_t R_size, const fq_t *blinds, const ge_t *G, const ge_t *h) {
#pragma omp parallel for schedule(dynamic) if (L_size > 1)
  for (size_t i = 0; i < L_size; i++) {
    ge_t acc, bh;
    ge_msm(&acc, Z + i * R_size, G, R_size);
    ge_scalarmul(&bh, &blinds[i], h);
    ge_add(&acc, &acc, &bh);
    ristretto_encode(out32 + 32 * i, &acc);
  }
}

/* AddrTimestamps::new inner loop, sparse_mlpoly.rs:229-243: read_ts[i] = audit_ts[addr]; audit_ts[addr] += 1 */
void spark_timestamps(uint64_t *read_ts, uint64_t *audit_ts, const uint64_t *addr, size_t num_ops) {
  for (size_t i = 0; i < num_ops; i++) { uint64_t a = addr[i]; read_ts[i] = audit_ts[a]; audit_ts[a] += 1; }
}
