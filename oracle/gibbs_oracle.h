/*
 * gibbs_oracle.h -- CPU restatement of the collapsed-Gibbs hot path of KenHBS/LDA_thesis.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product library (lda_thesis_b200/csrc) never links or calls it and has no CPU path.
 *
 * Parity status: the reference ships no tests, golden vectors or seeds, so the
 * pins are (1) the Random123 Philox4x32-10 known-answer vectors and (2) golden fixtures under
 * tests/golden/ produced by running the UNMODIFIED reference classes in the build container with
 * only `multinom_draw` replaced by the counter-keyed inverse-CDF draw (oracle/patched_reference.py,
 * oracle/make_golden.py).  `oracle_llda_exact_sweep` is checked against those fixtures
 * bit-for-bit on the integer state.
 *
 * Layout conventions (shared with include/gibbs_b200.h):
 *   doc_ptr  int64[D+1]   CSR offsets of the draws of each document
 *   word     int32[N]     word id of each draw (LabeledLDA.py:82 `self.docs`)
 *   freq     int32[N]     weight f of each draw (LabeledLDA.py:83 `self.freqs`; HSLDA: all 1)
 *   z        int32[N]     current topic of each draw, global topic id (LabeledLDA.py:73 `z_dn`)
 *   lab_ptr  int64[D+1]   CSR offsets of the active topic list of each document
 *   lab_idx  int32[...]   ascending global topic ids with lab==1 (LabeledLDA.py:94-99 `set_label`)
 *   n_wk     int32[V*ldk] word-major transpose of the reference's n_k_v[K][V] (LabeledLDA.py:76)
 *   n_dk_act int32[lab_ptr[D]]  n_d_k[d][lab_idx[..]] -- the only non-zero entries of LabeledLDA.py:75
 *   n_k      int32[K]     LabeledLDA.py:74 `n_zk`
 */
#ifndef GIBBS_ORACLE_H
#define GIBBS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Philox4x32-10 (Random123).  */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* One 32-bit word per (seed, stream, sweep, global document id, position of the draw inside the document);
 * see oracle/philox.py for the addressing. */
uint32_t oracle_draw_word(uint64_t seed, uint32_t stream, uint32_t sweep, uint64_t doc, uint64_t pos);

/* z ~ Uniform(labels of the doc): z = lab_idx[lab_ptr[d] + ((word * A) >> 32)], stream 1.
 * Device-side replacement of LabeledLDA.py:86-87 for synthetic scale runs. */
int oracle_init_z(int64_t D, const int64_t *doc_ptr, const int64_t *lab_ptr, const int32_t *lab_idx,
                  int32_t *z, uint64_t seed, uint64_t doc_base);

/* Histogram z -> counts (LabeledLDA.py:89-92, CascadeLDA.py:382-385, HSLDA.py:127-130). */
int oracle_counts_build(int64_t D, const int64_t *doc_ptr, const int32_t *word, const int32_t *freq,
                        const int32_t *z, const int64_t *lab_ptr, const int32_t *lab_idx,
                        int32_t K, int32_t V, int32_t ldk,
                        int32_t *n_wk, int32_t *n_dk_act, int32_t *n_k);

/* One corpus-order sequential sweep in fp64 with live counts: LabeledLDA.py:101-125 /
 * CascadeLDA.py:397-421 with the draw at :119/:415 replaced by the inverse-CDF draw. */
int oracle_llda_exact_sweep(int64_t d_begin, int64_t d_end, const int64_t *doc_ptr,
                            const int32_t *word, const int32_t *freq, int32_t *z,
                            const int64_t *lab_ptr, const int32_t *lab_idx,
                            int32_t K, int32_t V, int32_t ldk, double alpha, double beta,
                            int32_t *n_wk, int32_t *n_dk_act, int32_t *n_k,
                            uint64_t seed, uint32_t sweep, uint64_t doc_base);

/* One sweep of the document-parallel `snapshot` schedule in fp32: every draw reads
 * n_wk frozen at the start of its refresh block, the document's own live n_dk, and
 * n_k(frozen) + (n_dk(live) - n_dk(block start)).  Tile i is the document range
 * [tile_rng[2i], tile_rng[2i+1]) (empty ranges allowed); tiles with i % n_blocks == b form block b.
 * Weights of a document with at most 8 active topics are added left to right; longer label lists in 32-wide
 * Kogge-Stone chunks -- the two summation orders of the device kernels.
 * n_threads > 1 runs the documents of a block concurrently (the result does not depend on it). */
/* 0 when built without OpenMP, else omp_get_max_threads(). */
int oracle_openmp_threads(void);
int oracle_llda_snapshot_sweep(int64_t n_tiles, const int64_t *tile_rng, int32_t n_blocks,
                               const int64_t *doc_ptr, const int32_t *word, const int32_t *freq,
                               int32_t *z, const int64_t *lab_ptr, const int32_t *lab_idx,
                               int32_t K, int32_t V, int32_t ldk, double alpha, double beta,
                               int32_t *n_wk, int32_t *n_dk_act, int32_t *n_k,
                               uint64_t seed, uint32_t sweep, uint64_t doc_base, int32_t n_threads);

/* Frozen-phi test chains in fp64: LabeledLDA.py:155-212, CascadeLDA.py:186-247 (see include/gibbs_b200.h
 * gibbs_test_run for the argument meaning; phi_KV is [K][V]). */
int oracle_test_chains(int32_t K, int32_t V, const double *phi_KV, double alpha, double beta_fb, int64_t n_chains,
                       const int64_t *doc_ptr, const int32_t *word, const int32_t *freq,
                       const int64_t *lab_ptr, const int32_t *lab_idx, int32_t *z, int32_t init_mode,
                       int32_t it, int32_t thinning, uint64_t seed, int64_t chain_base, double *th_hat);

#ifdef __cplusplus
}
#endif
#endif
