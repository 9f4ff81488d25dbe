// spartan_b200 — launch wrappers around the sm_100a kernels (kernels.cu).  Plain C++ signatures so the host
// prover (prover.cpp, compiled by g++) never sees CUDA syntax.  All pointers are DEVICE pointers unless named h_*.
// Every wrapper enqueues on `stream` and returns immediately; errors surface through sp::dev::check().
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string>
#include "field.cuh"
#include "curve.cuh"

struct CUstream_st;
typedef CUstream_st* cudaStream_t;

namespace sp {
namespace dev {

// ---- runtime plumbing
void check(const char* what);  // throws std::runtime_error on a pending CUDA error
int device_count();
void set_device(int dev);
cudaStream_t stream_create();
cudaStream_t stream_create_prio(int level);   // > 0: greatest priority of the device, < 0: least
// a stream confined to `sms` SMs (green context); nullptr when unsupported; *sms_granted = the SM count of the partition
cudaStream_t stream_create_partition(int sms, int level, int* sms_granted);
void stream_wait_event(cudaStream_t s, void* ev);
void stream_destroy(cudaStream_t s);
void stream_sync(cudaStream_t s);
void* dmalloc(size_t bytes);
void dfree(void* p);
// size-bucketed caching allocator on top of cudaMalloc: a prove call allocates the same multi-hundred-MB buffers every time, and
// cudaMalloc / cudaFree (which synchronises the device) would otherwise cost milliseconds per proof
void* pool_alloc(size_t bytes);
void pool_free(void* p);
void pool_trim();   // return every cached block to the driver
void* hmalloc_pinned(size_t bytes);
void hfree_pinned(void* p);
void h2d(void* d, const void* h, size_t bytes, cudaStream_t s);
void d2h(void* h, const void* d, size_t bytes, cudaStream_t s);
void d2d(void* dst, const void* src, size_t bytes, cudaStream_t s);
void dzero(void* d, size_t bytes, cudaStream_t s);
int sm_count();
void mem_info(size_t* free_bytes, size_t* total_bytes);
// event timing of the kernels launched through these wrappers (bench.py roofline leg)
void* event_create();
void event_record(void* ev, cudaStream_t s);
float event_elapsed_ms(void* a, void* b);
void event_destroy(void* ev);
unsigned long long launch_count();  // kernels launched by this library since load
void io_bytes(unsigned long long* h2d_bytes, unsigned long long* d2h_bytes);  // bytes moved through h2d()/d2h() since load
void prof_enable(bool on);          // CUDA-event timing around every kernel family (clears previous records)
std::string prof_report();          // "name:launches:total_ms:total_algorithmic_bytes;..."

// ---- sumcheck rounds (K1/K2 of SURVEY.md §2b)
enum ScKind { SC_QUAD = 0 /*A*B*/, SC_CUBIC3 = 1 /*A*B*C*/, SC_CUBIC4 = 2 /*A*(B*C-D)*/ };
struct ScInst {        // one sumcheck instance: up to four tables of the same current length
  u256* t[4];          // A, B, C, D (unused entries null)
  u256* c_out;         // where the folded C goes (== t[2] for private C; a ping-pong buffer when C is shared)
  int write_c;         // 1: this instance stores the folded C
};
// Optional host notification of a reduction kernel: results are also stored to `host_out` (mapped pinned memory) and `seq` is then
// published in `*flag`; `done` is a zero-initialised device counter (instances of a batched launch finish independently).
struct HostSig {
  u256* host_out = nullptr;
  unsigned int* flag = nullptr;
  unsigned int* done = nullptr;
  unsigned int seq = 0;
};
// ---- intra-proof sharding over NVLink peer memory (comm.cu).  One process per GPU; every rank owns a "window" (plain cudaMalloc memory
// exported with cudaIpcGetMemHandle and mapped by every peer), laid out as [control | data half 0 | data half 1]:
//   control: per-source flag words and mailboxes for the per-round exchange of partial sums, flag words for the bulk all-gathers, an error word
//   data   : destination of the bulk all-gathers (alternating halves), read in place by the kernels that follow
// All ranks execute the same sequence of collectives, so a sequence number identifies an operation on every rank.
#define SP_MAX_RANKS 8
#define SP_XR_SLOTS 4
#define SP_XR_CAP 80                                   // u256 per (slot, source): 24 instances x 3 evaluations fit
#define SP_WIN_CTRL_BYTES ((size_t)1 << 20)
#define SP_WIN_HALF_BYTES ((size_t)96 << 20)
struct WinCtrl {                                        // at offset 0 of every window
  unsigned int xflag[SP_MAX_RANKS];                     // [source]: sequence number of the last partial-sum message from `source`
  unsigned int bflag[SP_MAX_RANKS];                     // [source]: sequence number of the last bulk all-gather contribution from `source`
  unsigned int err;                                     // set by a kernel whose peer wait timed out
  unsigned int pad[15];
  u256 mbox[SP_XR_SLOTS][SP_MAX_RANKS][SP_XR_CAP];      // [seq % SP_XR_SLOTS][source][value]
};
// Cross-rank completion of a reduction kernel: the finishing warp stores this rank's totals into every peer's mailbox (NVLink stores), publishes
// the sequence number, waits for the peers' messages and adds them up in rank order (exact field arithmetic: every rank gets identical bytes),
// then hands the sums to the host.  world <= 1: no exchange.
struct XRank {
  int world = 1, rank = 0;
  unsigned int seq = 0;
  WinCtrl* win[SP_MAX_RANKS] = {};                      // this process's mapping of every rank's window (win[rank] = own)
};

// scratch: >= sc_scratch_bytes(); out: ninst*3 scalars [e0,e2,e3] (e3 = 0 for SC_QUAD), device memory
size_t sc_scratch_bytes(int ninst);
void dot_pairs(u256* out, const u256* const* a_list, const u256* const* b_list, int count, size_t n, void* scratch, cudaStream_t s, HostSig sig = HostSig());
void heads(u256* out, const u256* const* tables, int count, cudaStream_t s, HostSig sig = HostSig());
void sc_eval(ScKind kind, const ScInst* d_insts, int ninst, size_t len, u256* out, void* scratch, cudaStream_t s, HostSig sig = HostSig(), const XRank& xr = XRank());
// fold every table of every instance by r (len -> len/2, in place) and evaluate the next round on the result; r travels as a kernel argument
void sc_fold_eval(ScKind kind, const ScInst* d_insts, int ninst, size_t len, const u256& r, u256* out, void* scratch, cudaStream_t s,
                  HostSig sig = HostSig(), const XRank& xr = XRank());
// ---- eq-factored streaming rounds of a batched A*B*eq sumcheck (kernels_sc.cu: k_sc_eval_g): instances use t[0] = A, t[1] = B; E: the suffix eq table
// of the round (len/2 entries for sc_eval_g, len/4 for sc_fold_eval_g); out[3*i + 0] = q(0), out[3*i + 1] = q(inf) of instance i
void sc_eval_g(const ScInst* d_insts, int ninst, size_t len, const u256* E, u256* out, void* scratch, cudaStream_t s, HostSig sig = HostSig(), const XRank& xr = XRank());
void sc_fold_eval_g(const ScInst* d_insts, int ninst, size_t len, const u256& r, const u256* E, u256* out, void* scratch, cudaStream_t s, HostSig sig = HostSig(),
                    const XRank& xr = XRank());
// levels[k-1] (k = 1..K, back to back, n0 >> k entries each) = eq(tau[k+1..], .) from E0 = eq(tau[1..], .) of n0 entries
size_t eq_suffix_entries(size_t n0, int K);
void eq_suffix(u256* levels, const u256* E0, size_t n0, int K, cudaStream_t s);
// ---- persistent tail of a batched cubic sumcheck (prove_cubic_batched, sumcheck.rs:254-424, once the tables are small): ONE launch runs all the
// remaining rounds.  CTA i owns instance i (tables A, B and a private copy of C); after every bind it publishes the instance's evaluations to
// the host (HostSig, sequence numbers sig.seq, sig.seq+1, ...), then spins on a mailbox in mapped pinned host memory until the host has derived
// the next challenge from the transcript.  After the last bind it publishes the bound heads A[0], B[0], C[0] instead.  No launch, no cold
// caches and no kernel drain between rounds: a round costs the PCIe round trip plus a few microseconds of arithmetic.
struct PersistMail { u256 r; unsigned int seq; unsigned int pad[7]; };
// One CTA (one SM) per instance: worth it only while a round is latency, not throughput (<= 256 entries: one bind task per thread)
#define SC_PERSIST_MAX_LEN 256
// dmail: a device-memory copy of the mailbox (zero-initialised, reused across launches): only CTA 0 polls the host over PCIe and forwards the challenge
void sc_persist(const ScInst* insts, int ninst, int n_shared_c /* instances [0, n_shared_c) read the shared C table */, u256* c_scratch /* n_shared_c * len */,
                size_t len, const u256& r0, const PersistMail* mail, PersistMail* dmail, unsigned int mail_seq0, u256* out, cudaStream_t s, HostSig sig);
// fold only (bound_poly_var_top, dense_mlpoly.rs:215-223): tables[k][i] += r*(tables[k][i+len/2]-tables[k][i])
void fold_top(u256* const* d_tables, int ntables, size_t len, const u256& r, cudaStream_t s);
void fold_top_single(u256* table, size_t len, const u256& r, cudaStream_t s);

// ---- dense polynomial helpers (K7)
void eq_evals(u256* out, const u256* d_r, int ell, u256* scratch_small /* >= 2*2^ceil(ell/2) */, cudaStream_t s);
void dot(u256* out, const u256* a, const u256* b, size_t n, void* scratch, cudaStream_t s);
void dot_many(u256* out, const u256* const* a_list, int count, const u256* b, size_t n, void* scratch, cudaStream_t s);
void dot3(u256* out, const u256* a, const u256* b, const u256* c, size_t n, void* scratch, cudaStream_t s);
void bound_rows(u256* out, const u256* Z, const u256* L, size_t L_size, size_t R_size, u256* scratch /* >= 64*R_size */, cudaStream_t s);
void lincomb3(u256* out, const u256* A, const u256* B, const u256* C, const u256* d_rabc /*3*/, size_t n, cudaStream_t s);
void hadamard(u256* out, const u256* a, const u256* b, size_t n, cudaStream_t s);
void hadamard_many(u256* const* outs, const u256* const* as, const u256* const* bs, int count, size_t n, cudaStream_t s);
void from_u64(u256* out, const uint64_t* v, size_t n, cudaStream_t s);
void from_bytes_wide(u256* out, const uint8_t* in64, size_t n, cudaStream_t s);
void batch_invert_elems(u256* inout, size_t n, cudaStream_t s);
void to_canonical(u256* out, const u256* in, size_t n, cudaStream_t s);   // Montgomery -> canonical little-endian integers (transcript bytes)
void gather(u256* out, const u256* mem, const uint32_t* idx, size_t n, cudaStream_t s);
// hash layer of the SPARK memory check: out = ts*r^2 + val*r + addr - g  (sparse_mlpoly.rs:545-600)
// addr == null: addr = index i;  ts == null: ts = 0;  ts_plus_one adds one.   d_rg = [r_hash, r_multiset]
void spark_hash(u256* out, size_t n, const u256* addr, const u256* val, const u256* ts, int ts_plus_one, const u256* d_rg, cudaStream_t s);
// IPA vector folds (bullet.rs:105-107): a[i] = a[i]*u + uinv*a[i+n];  b[i] = b[i]*uinv + u*b[i+n]   d_u = [u, uinv]
void ipa_fold_ab(u256* a, u256* b, size_t n, const u256& u, const u256& uinv, cudaStream_t s);
// scalars for the L / R commitments against the UNFOLDED generators: see DESIGN.md "IPA without folding G"
void ipa_lr_scalars(u256* outL, u256* outR, const u256* a, const u256* svec, size_t n_cur, size_t n_full, cudaStream_t s);
void ipa_update_s(u256* svec, size_t n_cur_half, size_t n_full, const u256& u, const u256& uinv, cudaStream_t s);
void ipa_fold_update(u256* a, u256* b, u256* svec, size_t n_cur_half, size_t n_full, const u256& u, const u256& uinv, cudaStream_t s);   // ipa_fold_ab + ipa_update_s in one launch
void fill_one(u256* out, size_t n, cudaStream_t s);

// windows: IPC-exportable device allocations and their mappings in the peers
void* win_alloc(size_t bytes);
void win_free(void* p);
size_t ipc_handle_bytes();
void ipc_export(void* devptr, uint8_t* out);
void* ipc_open(const uint8_t* handle);
void ipc_close(void* p);
// ---- sharded table producers: rank `rank` of `world` holds the global indices i = j*world + rank (cyclic partition) at local index j
void scale(u256* inout, const u256& c, size_t n, cudaStream_t s);                                        // x[i] *= c
void take_cyclic(u256* out, const u256* full, size_t n_local, int rank, int world, cudaStream_t s);      // out[j] = full[j*world + rank]
void spmv_cyclic(u256* out, size_t nrows_local, int rank, int world, const uint32_t* ptr, const uint32_t* idx, const u256* val, const u256* x, cudaStream_t s);
void spark_hash_cyclic(u256* out, size_t n_local, int rank, int world, const u256* addr, const u256* val, const u256* ts, int ts_plus_one, const u256* d_rg, cudaStream_t s);
// ---- bulk all-gathers into every rank's window (comm.cu); `dst_off`: byte offset inside the window, the same on every rank
struct CommDev {                                        // device-side view of the communicator
  int world = 1, rank = 0;
  uint8_t* win[SP_MAX_RANKS] = {};
};
// every rank's `bytes` land at dst_off + source*bytes of every window
void push_block(const CommDev& c, const void* src, size_t bytes, size_t dst_off, unsigned int seq, unsigned int* ticket, cudaStream_t s);
// table t of rank `source`, element j  ->  window element (dst_off/32) + t*(n_local*world) + j*world + source
void push_cyclic(const CommDev& c, const u256* const* tables, int ntables, size_t n_local, size_t dst_off, unsigned int seq, unsigned int* ticket, cudaStream_t s);
void wait_peers(const CommDev& c, unsigned int seq, cudaStream_t s);   // returns (on the stream) once every peer's contribution `seq` has landed

// ---- sparse matrix-vector products on CSR (row-major) / CSC (column-major) copies of the COO triples
void spmv(u256* out, size_t nrows, const uint32_t* ptr, const uint32_t* idx, const u256* val, const u256* x, cudaStream_t s);
// sum_k trx[row_k]*try[col_k]*val_k
void sparse_eval3(u256* out, const uint32_t* row, const uint32_t* col, const u256* val, size_t nnz, const u256* trx, const u256* try_, void* scratch, cudaStream_t s);

// ---- group: generators, fixed-base window tables, multi-row MSM, compression (K3/K4/K6)
void gens_from_uniform(ge* out, const uint8_t* d_uniform64, size_t n, cudaStream_t s);
void decompress_batch(ge* out, int* ok, const uint8_t* in32, size_t n, cudaStream_t s);
void compress_batch(uint8_t* out32, const ge* in, size_t n, cudaStream_t s);
// table[(j*NWIN + w)*2^(W-1) + (d-1)] = d * 2^(W*w) * G_j in affine-niels form; W (wbits) is 8 or 13, NWIN = ceil(253 / W)
size_t table_entries(size_t nbases, int wbits);
void build_tables(ge_niels* table, const ge* G, size_t nbases, int wbits, cudaStream_t s);
// out[row] = sum_{j<R} scalars[row*stride + j] * G_j  (+ blinds[row] * G_{blind_base} when blinds != null)
// scalars are Montgomery-form; partial: scratch >= msm_scratch_bytes(L, R)
size_t msm_scratch_bytes(size_t L, size_t R);
// launch shape overrides for an MSM that runs in the background of latency-bound work: cpt = column passes per CTA (0: automatic, 1 or 4: short- or
// long-lived CTAs), smem_pad = bytes of unused dynamic shared memory per CTA (caps the resident CTAs per SM)
struct MsmTune { int cpt = 0; size_t smem_pad = 0; };
void msm_rows(ge* out, const ge_niels* table, int wbits, const u256* scalars, size_t stride, size_t L, size_t R, const u256* blinds, size_t blind_base,
              void* scratch, cudaStream_t s, const MsmTune& tune = MsmTune());

void add_points(ge* a, const ge* b, size_t n, cudaStream_t s);   // a[i] += b[i]
void sum_points(ge* out, const ge* in, int n, cudaStream_t s);   // out[0] = in[0] + ... + in[n-1], n <= 32
// both MSMs of an inner-product round (L, R -> out[0], out[1]) over unfolded generators: scalar of generator j is a[.]*svec[j].
// scratch >= ipa_msm_scratch_points(n_full, wbits) points; ticket: one zero-initialised word (self-resetting)
size_t ipa_msm_scratch_points(size_t n_full, int wbits);
void ipa_msm(ge* out, const ge_niels* table, int wbits, const u256* a, const u256* svec, size_t n_cur, size_t n_full, void* scratch, unsigned int* ticket,
             cudaStream_t s, HostSig sig = HostSig(), int max_ctas = 0 /* > 0: cap on the blocks of the quad-lane kernel (both sides together) */,
             // b != null (only when ipa_msm_fuses_dots()): the same launch also computes c_out[0] = <a_L, b_R>, c_out[1] = <a_R, b_L> of the round and
             // publishes them through sigc (its own flag word and completion counter), instead of a separate dot_pairs launch in front of the MSM
             const u256* b = nullptr, u256* c_out = nullptr, HostSig sigc = HostSig());
bool ipa_msm_fuses_dots();

// ---- variable-base MSM on arbitrary points (bucket method; kernels_pip.cu).  pts: affine-niels form of the caller's points.
struct PipPlan { size_t n = 0; int c = 0, nwin = 0, G = 32; uint32_t nb = 0; size_t tile = 0, ntiles = 0, S = 0, max_items = 0; };
PipPlan pip_plan(size_t n, int c_override);   // c_override = 0: pick the window width from n
size_t pip_scratch_bytes(const PipPlan& p);
void points_to_niels(ge_niels* out, const ge* in, size_t n, cudaStream_t s);
void niels_to_ge(ge* out, const ge_niels* in, size_t n, cudaStream_t s);
// out = sum_i scalars[i] * pts[i]   (scalars Montgomery-form, device-resident; scratch >= pip_scratch_bytes(p))
void msm_var(ge* out, const ge_niels* pts, const u256* scalars, const PipPlan& p, void* scratch, cudaStream_t s);

}  // namespace dev
}  // namespace sp
