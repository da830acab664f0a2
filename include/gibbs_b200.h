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
 *   gibbs_test_*         LabeledLDA.py:179-212 run_test, CascadeLDA.py:210-247 cascade_test (frozen phi)
 *
 * Conventions
 *   - plain C, no torch/NumPy types; the caller owns every host buffer, the handle owns all device memory
 *   - every function returns 0 on success or a negative GIBBS_E_* code; gibbs_last_error() gives the text
 *   - calls are synchronous on return unless stated; a handle is not thread-safe
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
    int64_t  draw_base;   /* global index of this shard's first draw (RNG addressing across shards) */
    int64_t  tile_base;   /* global index of this shard's first tile (refresh-block assignment across shards) */
    int32_t  tile_docs;   /* documents per tile (0 -> library default) */
    int32_t  reserved;
} gibbs_desc;

typedef struct {
    int64_t  draws;            /* draws resampled since create */
    int64_t  sweeps;           /* completed sweeps */
    double   last_sweep_ms;    /* CUDA-event time of the sampling kernels of the last gibbs_sweep call, per sweep */
    double   last_merge_ms;    /* same for the delta merge */
    int64_t  last_launches;    /* kernels launched by the last gibbs_sweep / gibbs_sweep_begin+end */
    double   bytes_per_draw;   /* algorithmic bytes per draw of the dense-row model */
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

/* n_sweeps full sweeps (sampling kernels + delta merge).  Single shard only. */
int gibbs_sweep(gibbs_t *h, int32_t n_sweeps);

/* Multi-shard protocol, one call sequence per refresh block b = 0 .. n_refresh-1:
 *   gibbs_sweep_begin(h, b)   -- sample this shard's tiles of block b against the frozen table into the delta table
 *   (caller all-reduces the buffer returned by gibbs_delta_buffer over all shards, e.g. NCCL sum, int32)
 *   gibbs_sweep_end(h, b)     -- n_wk += delta, n_k += column sums, delta = 0; after the last block: sweep counter++ */
int gibbs_sweep_begin(gibbs_t *h, int32_t block);
int gibbs_sweep_end  (gibbs_t *h, int32_t block);
/* Device pointer + element count (int32) of the delta table, for the caller's collective. */
int gibbs_delta_buffer(gibbs_t *h, void **dev_ptr, int64_t *n_elems);
/* Stream all of the handle's work is enqueued on (cudaStream_t as void*). */
int gibbs_stream(gibbs_t *h, void **stream);

/* Copy state out.  Any pointer may be NULL.  n_wk is written word-major [V][K] (no padding). */
int gibbs_get_state(gibbs_t *h, int32_t *z, int32_t *n_wk, int32_t *n_dk_act, int32_t *n_k);
/* Replace z (global topic ids) and rebuild all counts. */
int gibbs_set_z(gibbs_t *h, const int32_t *z);

/* phi[K][V] (C order, fp64).  smoothed != 0: (n_kv + beta) / (n_k + V*beta)  (LabeledLDA.py:231-234)
 *                             smoothed == 0: n_kv / sum_v n_kv, NaN rows for empty topics (CascadeLDA.py:394-395).
 * The thinning mean (LabeledLDA.py:138-145) is taken by the host class over these snapshots. */
int gibbs_emit_phi(gibbs_t *h, double *phi_KV, int32_t smoothed);
/* theta[D][K] dense fp64: (n_dk + lab*alpha) / rowsum (LabeledLDA.py:236-239); smoothed == 0: n_dk / rowsum
 * (HSLDA.py:148-149). */
int gibbs_emit_theta(gibbs_t *h, double *theta_DK, int32_t smoothed);

int gibbs_stats(gibbs_t *h, gibbs_stats_t *out);
int gibbs_set_sweep_counter(gibbs_t *h, uint32_t sweep);

/* HSLDA.sample_z state (HSLDA.py:222-231): eta[L][K], per-document label lists come from lab_ptr/lab_idx of
 * gibbs_load (labels, not topics, for this kind), a_act / mean_a_act aligned with lab_idx, alpha_beta[K]. */
int gibbs_hslda_set(gibbs_t *h, int32_t L, const double *eta, const double *a_act, const double *mean_a_act,
                    const double *alpha_beta);

/* Frozen-phi test chains (LabeledLDA.py:179-212): independent documents, `it` sweeps, thinning mean of
 * n_dk / sum(n_dk) written to th_hat[D_test][K].  phi_KV is [K][V] fp64 as produced by gibbs_emit_phi. */
int gibbs_test_chains(int32_t device, int32_t K, int32_t V, double alpha, const double *phi_KV,
                      int64_t D_test, const int64_t *doc_ptr, const int32_t *word, const int32_t *freq,
                      const int32_t *z_init, int32_t it, int32_t thinning, uint64_t seed, double *th_hat);

/* Known-answer hook: Philox4x32-10 evaluated ON THE DEVICE for n (ctr,key) pairs. */
int gibbs_philox_kat(int32_t device, int32_t n, const uint32_t *ctr4, const uint32_t *key2, uint32_t *out4);

#ifdef __cplusplus
}
#endif
#endif /* GIBBS_B200_H */
