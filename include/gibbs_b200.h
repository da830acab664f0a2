/*
 * gibbs_b200.h -- C-ABI of libgibbs_b200.so: the collapsed-Gibbs z-resampling sweep of
 * KenHBS/LDA_thesis (LabeledLDA / CascadeLDA.SubLDA / HSLDA) on one B200 (sm_100a).
 *
 * The reference is pure Python and has no FFI of its own; what this library replaces is the
 * body of three methods and the count matrices they own.  Each entry point cites the
 * reference statements it stands for (paths relative to the reference checkout):
 *
 *   gibbs_load           LabeledLDA.py:73-92, CascadeLDA.py:364-385, HSLDA.py:117-130 (state + histogram)
 *   gibbs_sweep          LabeledLDA.py:101-125 training_iteration, CascadeLDA.py:397-421
 *   gibbs_hslda_*        HSLDA.py:171-272 sample_z
 *   gibbs_emit_phi       LabeledLDA.py:231-234 get_phi, CascadeLDA.py:394-395 get_ph, HSLDA.py:151-152
 *   gibbs_emit_theta     LabeledLDA.py:236-239 get_theta, HSLDA.py:148-149 get_zbar
 *   gibbs_test_*         LabeledLDA.py:155-212 prep4test + run_test, CascadeLDA.py:186-247 cascade_test (frozen phi)
 *   gibbs_thin_*         LabeledLDA.py:138-145, CascadeLDA.py:430-434, HSLDA.py:327-333 (thinning mean)
 *   gibbs_perplexity     LabeledLDA.py:256-265
 *
 * Conventions
 *   - plain C, no torch/NumPy types; the caller owns every host buffer, the handle owns all device memory
 *   - every function returns 0 on success or a negative GIBBS_E_* code; gibbs_last_error() gives the text
 *   - calls are synchronous on return; a handle is not thread-safe; gibbs_load may be called again on a live handle
 *     (same D, V, K) and re-uses its device allocations
 *   - there is NO CPU fallback: without a CUDA device gibbs_create fails with GIBBS_E_CUDA
 *
 * Corpus layout (CSR over "draws"; a draw is one (document, unique word id) pair with weight f,
 * LabeledLDA.py:108; one raw token with f == 1 for HSLDA, HSLDA.py:232):
 *   doc_ptr  int64[D+1]   draws of document d are [doc_ptr[d], doc_ptr[d+1])
 *   word     int32[N]     word id  (LabeledLDA.py:82 self.docs)
 *   freq     int32[N]     weight f (LabeledLDA.py:83 self.freqs); NULL means all ones
 *   z        int32[N]     topic of each draw, global topic id (LabeledLDA.py:73 self.z_dn)
 *   lab_ptr  int64[D+1]   active-topic list of document d is lab_idx[lab_ptr[d] .. lab_ptr[d+1])
 *   lab_idx  int32[...]   ascending topic ids with lab == 1 (LabeledLDA.py:94-99 set_label; root = 0 first)
 * Count layout
 *   n_wk     int32[V][ldk]  word-major transpose of the reference's n_k_v[K][V] (LabeledLDA.py:76),
 *                           ldk = K rounded up to a multiple of 32 (one 128-byte line)
 *   n_dk_act int32[lab_ptr[D]]  n_d_k[d][lab_idx[..]] -- the only entries of LabeledLDA.py:75 that can be non-zero
 *   n_k      int32[K]       LabeledLDA.py:74 self.n_zk
 */
#ifndef GIBBS_B200_H
#define GIBBS_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GIBBS_OK          0
#define GIBBS_E_ARG      -1   /* bad argument / inconsistent corpus (e.g. z not in the document's label list) */
#define GIBBS_E_CUDA     -2   /* CUDA runtime error, including "no device" */
#define GIBBS_E_NOMEM    -3
#define GIBBS_E_STATE    -4   /* call out of order (sweep before load, ...) */

#define GIBBS_KIND_LLDA   0   /* LabeledLDA / SubLDA / batched CascadeLDA nodes / LocalLDA (full mask) */
#define GIBBS_KIND_HSLDA  1

#define GIBBS_MODE_EXACT     0  /* corpus-order sequential chain, fp64, live counts: bit-parity with the reference loop */
#define GIBBS_MODE_SNAPSHOT  1  /* document-parallel, fp32, counts frozen per refresh block, integer delta table */

/* Row fetch of the snapshot schedule (same arithmetic, same results):
 *   DENSE  -- a group of lanes per document; cp.async of the whole ldk-wide row (or the document's topic segment)
 *             into a shared-memory ring
 *   GATHER -- one 4-byte load per active topic (touches |label list| 32-byte sectors): thread-per-document kernel for
 *             label lists <= 8, group-per-document kernels for <= 32; longer lists always use DENSE
 *   AUTO   -- GATHER where it moves fewer bytes than the row, DENSE otherwise */
#define GIBBS_FETCH_AUTO     0
#define GIBBS_FETCH_DENSE    1
#define GIBBS_FETCH_GATHER   2

#define GIBBS_COMM_ID_BYTES  128

typedef struct gibbs_handle gibbs_t;

typedef struct {
    int32_t  kind;        /* GIBBS_KIND_*  */
    int32_t  mode;        /* GIBBS_MODE_*  */
    int64_t  D;           /* documents in this shard */
    int32_t  V;           /* vocabulary size; V*beta uses this (LabeledLDA.py:115, CascadeLDA.py:361,411) */
    int32_t  K;           /* topics (global topic index space) */
    double   alpha;       /* LabeledLDA.py:56 / HSLDA: unused (alpha*beta_k vector is set by gibbs_hslda_set) */
    double   beta;        /* LabeledLDA.py:57 / HSLDA.py:93 gamma */
    uint64_t seed;        /* Philox key */
    int32_t  device;      /* CUDA device ordinal */
    int32_t  n_refresh;   /* snapshot mode: refresh blocks per sweep (>= 1) */
    int64_t  doc_base;    /* global id of this shard's document 0: the RNG is addressed by (global document id, position in
                             the document) and a document's refresh block is ((doc_base + d) / tile_docs) % n_refresh, so a
                             chain does not depend on how documents are sharded */
    int64_t  reserved64;
    int32_t  tile_docs;   /* documents per tile (0 -> library default 256) */
    int32_t  row_fetch;   /* GIBBS_FETCH_*: how the snapshot kernels read n_wk[v][label list] */
} gibbs_desc;

typedef struct {
    int64_t  draws;            /* draws resampled since create */
    int64_t  sweeps;           /* completed sweeps */
    double   last_sweep_ms;    /* CUDA-event time of the sampling kernels of the LAST sweep of the last gibbs_sweep call */
    double   last_merge_ms;    /* same for the delta all-reduce + merge */
    double   last_call_ms;     /* CUDA-event time of the whole last gibbs_sweep call (all its sweeps) */
    int64_t  last_launches;    /* kernels launched per sweep by the last gibbs_sweep call */
    double   bytes_per_draw;   /* algorithmic bytes per draw of the row fetch in use */
    double   bytes_per_draw_dense;
    double   bytes_per_draw_gather;
    int32_t  row_fetch;        /* GIBBS_FETCH_DENSE or GIBBS_FETCH_GATHER: the one that carries most draws */
    int32_t  reserved;
    int32_t  ldk;
    int32_t  max_active;       /* max |label list| over documents */
    int64_t  changed;          /* draws whose topic changed in the last sweep */
    int64_t  device_bytes;     /* device memory owned by the handle */
} gibbs_stats_t;

const char *gibbs_last_error(void);
const char *gibbs_version(void);
/* Number of visible CUDA devices (0 and GIBBS_E_CUDA text in gibbs_last_error when there are none). */
int gibbs_device_count(void);

int  gibbs_create (gibbs_t **out, const gibbs_desc *desc);
void gibbs_destroy(gibbs_t *h);

/* Upload the corpus and build the three count arrays on the device.
 * z_init == NULL: z ~ Uniform(label list) from Philox stream 1 (device-side replacement of LabeledLDA.py:86-87).
 * seg    == NULL: every draw reads the whole ldk-wide n_wk row; otherwise seg[2*d], seg[2*d+1] = [lo, hi) topic
 *                 range document d needs (a CascadeLDA node's topic block); lo and hi multiples of 4. */
int gibbs_load(gibbs_t *h, const int64_t *doc_ptr, const int32_t *word, const int32_t *freq,
               const int32_t *z_init, const int64_t *lab_ptr, const int32_t *lab_idx, const int32_t *seg);

/* n_sweeps full sweeps, enqueued back to back on the handle's stream; synchronous on return.
 * Per refresh block: sampling kernels -> [all-reduce of the delta table over the communicator] -> merge
 * (n_wk += delta, n_k += column sums, delta = 0). */
int gibbs_sweep(gibbs_t *h, int32_t n_sweeps);

/* Multi-GPU (one process per GPU; SURVEY.md §8e -- the reference has no counterpart).  Documents are sharded by the
 * caller: each rank creates its handle with its shard's D and doc_base, calls gibbs_comm_init BEFORE
 * gibbs_load, and from then on gibbs_load all-reduces the initial histograms and every refresh block of gibbs_sweep
 * all-reduces the int32 delta table (ncclAllReduce, sum) on the handle's stream.  Every rank then holds identical
 * n_wk / n_k, and the chain is bit-identical to the single-GPU run of the concatenated corpus.
 * The 128-byte id comes from gibbs_comm_unique_id on one rank and is distributed by the caller. */
int gibbs_comm_unique_id(char *id /* [GIBBS_COMM_ID_BYTES] */);
int gibbs_comm_init(gibbs_t *h, int32_t nranks, int32_t rank, const char *id /* [GIBBS_COMM_ID_BYTES] */);
/* Stream all of the handle's work is enqueued on (cudaStream_t as void*). */
int gibbs_stream(gibbs_t *h, void **stream);

/* Copy state out.  Any pointer may be NULL.  n_wk is written word-major [V][K] (no padding). */
int gibbs_get_state(gibbs_t *h, int32_t *z, int32_t *n_wk, int32_t *n_dk_act, int32_t *n_k);
/* Replace z (global topic ids) and rebuild all counts. */
int gibbs_set_z(gibbs_t *h, const int32_t *z);

/* n_wk[word[i]][topic[i]] += count[i] for i < n; n_k and n_dk are left alone.  Exists for ONE reference quirk:
 * SubLDA.__init__ (CascadeLDA.py:382-385) iterates (id, freq) tuples, so `n_k_v[z, (id, freq)] += f` also credits
 * column `freq`; a drop-in CascadeLDA must start from the same table. */
int gibbs_add_counts(gibbs_t *h, int64_t n, const int32_t *word, const int32_t *topic, const int32_t *count);

/* phi[K][V] (C order, fp64).  smoothed != 0: (n_kv + beta) / (n_k + V*beta)  (LabeledLDA.py:231-234)
 *                             smoothed == 0: n_kv / sum_v n_kv, NaN rows for empty topics (CascadeLDA.py:394-395).
 * The thinning mean (LabeledLDA.py:138-145) is taken by the host class over these snapshots. */
int gibbs_emit_phi(gibbs_t *h, double *phi_KV, int32_t smoothed);
/* theta[D][K] dense fp64: (n_dk + lab*alpha) / rowsum (LabeledLDA.py:236-239); smoothed == 0: n_dk / rowsum
 * (HSLDA.py:148-149). */
int gibbs_emit_theta(gibbs_t *h, double *theta_DK, int32_t smoothed);

/* theta over the label lists only: theta_act[lab_ptr[d] + j] = theta[d][lab_idx[lab_ptr[d] + j]]; every other entry of
 * the dense matrix is 0.  Use this at scale (D*K doubles do not fit anywhere at 1M x 500). */
int gibbs_emit_theta_csr(gibbs_t *h, double *theta_act, int32_t smoothed);

int gibbs_stats(gibbs_t *h, gibbs_stats_t *out);
/* Free the handle's staging memory (upload / export scratch); it is re-allocated on demand. */
int gibbs_trim(gibbs_t *h);
int gibbs_set_sweep_counter(gibbs_t *h, uint32_t sweep);

/* HSLDA.sample_z state (HSLDA.py:222-231): eta[L][K], per-document label lists come from lab_ptr/lab_idx of
 * gibbs_load (labels, not topics, for this kind), a_act / mean_a_act aligned with lab_idx, alpha_beta[K]. */
int gibbs_hslda_set(gibbs_t *h, int32_t L, const double *eta, const double *a_act, const double *mean_a_act,
                    const double *alpha_beta);

/* Thinning running mean kept on the device (LabeledLDA.py:138-145, CascadeLDA.py:430-434, HSLDA.py:327-333):
 *   hat = c_old * hat + c_new * current     (the first call after gibbs_load stores `current`)
 * `what`: bit 0 = phi [K][V] (smoothed as in gibbs_emit_phi), bit 1 = theta over the label lists (gibbs_emit_theta_csr).
 * The host class passes the reference's own coefficients ((s-1)/s, 1/s) so every rounding matches NumPy's. */
int gibbs_thin_accumulate(gibbs_t *h, double c_old, double c_new, int32_t smoothed, int32_t what);
int gibbs_thin_get(gibbs_t *h, double *ph_hat_KV, double *th_hat_act);   /* either pointer may be NULL */

/* Training perplexity of LabeledLDA.py:256-265 from the live counts: *neg_log_sum = -sum over (doc, unique word) pairs of
 * log(phi[:, w] . theta_d), *n_pairs = number of pairs; perplexity = exp(neg_log_sum / n_pairs). */
int gibbs_perplexity(gibbs_t *h, double *neg_log_sum, int64_t *n_pairs);

/* Frozen-phi test chains: LabeledLDA.py:155-212 (prep4test + run_test), CascadeLDA.py:186-247 (prep4test + cascade_test).
 * The handle keeps a word-major device copy of phi_KV ([K][V] fp64, e.g. model.ph_hat / model.ph).
 * A chain is one (document, topic list) pair; chains are independent.  lab_ptr/lab_idx == NULL: every chain uses all K
 * topics and th_hat is [n_chains][K]; otherwise th_hat is aligned with lab_idx.
 *   init_mode 0: z holds the start state (global topic ids)
 *             1: z ~ phi[:, v] (LabeledLDA.py:162-175)
 *             2: z ~ (phi[:, v] + beta_fb) / sum with the first list entry's weight set to 1 / len(doc) (CascadeLDA.py:194-206)
 *   beta_fb > 0: when every weight of a draw is zero, redraw from (n_dk + alpha) * (phi + beta_fb) (CascadeLDA.py:225-230)
 * th_hat: thinning mean of n_dk / sum(n_dk) over iterations i with (i + 1) % thinning == 0; zeros if none.
 * z (optional unless init_mode 0) receives the final assignments.  RNG: stream 2 / 4, counter (iteration, chain_base + chain, position). */
typedef struct gibbs_test_handle gibbs_test_t;
int  gibbs_test_create(gibbs_test_t **out, int32_t device, int32_t K, int32_t V, const double *phi_KV);
void gibbs_test_destroy(gibbs_test_t *t);
int  gibbs_test_run(gibbs_test_t *t, double alpha, double beta_fb, int64_t n_chains, const int64_t *doc_ptr,
                    const int32_t *word, const int32_t *freq, const int64_t *lab_ptr, const int32_t *lab_idx,
                    int32_t *z, int32_t init_mode, int32_t it, int32_t thinning, uint64_t seed, int64_t chain_base,
                    double *th_hat);

/* Known-answer hook: Philox4x32-10 evaluated ON THE DEVICE for n (ctr,key) pairs. */
int gibbs_philox_kat(int32_t device, int32_t n, const uint32_t *ctr4, const uint32_t *key2, uint32_t *out4);

#ifdef __cplusplus
}
#endif
#endif /* GIBBS_B200_H */
