/*
 * gibbs_oracle.c -- CPU restatement of the collapsed-Gibbs hot path (TEST INFRASTRUCTURE ONLY).
 * See gibbs_oracle.h for the contract, the layout and the parity status.
 *
 * Build: oracle/Makefile  (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 * -ffp-contract=off matters: every floating-point statement below is one IEEE operation, and the
 * device kernels use the matching __fadd_rn/__fmul_rn/__fdiv_rn/__dmul_rn/... intrinsics.
 */
#include "gibbs_oracle.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ Philox4x32-10 (Random123) */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

uint32_t oracle_draw_word(uint64_t seed, uint32_t stream, uint32_t sweep, uint64_t doc, uint64_t pos) {
    uint32_t ctr[4] = {(uint32_t)(pos >> 2), (uint32_t)doc, sweep, stream};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    oracle_philox4x32_10(ctr, key, out);
    return out[pos & 3];
}

/* ------------------------------------------------------------------ init + histogram */
int oracle_init_z(int64_t D, const int64_t *doc_ptr, const int64_t *lab_ptr, const int32_t *lab_idx,
                  int32_t *z, uint64_t seed, uint64_t doc_base) {
    for (int64_t d = 0; d < D; ++d) {
        uint64_t A = (uint64_t)(lab_ptr[d + 1] - lab_ptr[d]);
        if (A == 0) return -1;
        for (int64_t n = doc_ptr[d]; n < doc_ptr[d + 1]; ++n) {
            uint32_t w = oracle_draw_word(seed, 1u, 0u, doc_base + (uint64_t)d, (uint64_t)(n - doc_ptr[d]));
            z[n] = lab_idx[lab_ptr[d] + (int64_t)(((uint64_t)w * A) >> 32)];
        }
    }
    return 0;
}

static int find_label(const int32_t *lab, int A, int32_t k) {
    for (int j = 0; j < A; ++j) if (lab[j] == k) return j;
    return -1;
}

int oracle_counts_build(int64_t D, const int64_t *doc_ptr, const int32_t *word, const int32_t *freq,
                        const int32_t *z, const int64_t *lab_ptr, const int32_t *lab_idx,
                        int32_t K, int32_t V, int32_t ldk,
                        int32_t *n_wk, int32_t *n_dk_act, int32_t *n_k) {
    memset(n_wk, 0, sizeof(int32_t) * (size_t)V * (size_t)ldk);
    memset(n_k, 0, sizeof(int32_t) * (size_t)K);
    memset(n_dk_act, 0, sizeof(int32_t) * (size_t)lab_ptr[D]);
    for (int64_t d = 0; d < D; ++d) {
        const int32_t *lab = lab_idx + lab_ptr[d];
        int A = (int)(lab_ptr[d + 1] - lab_ptr[d]);
        for (int64_t n = doc_ptr[d]; n < doc_ptr[d + 1]; ++n) {
            int32_t k = z[n], v = word[n], f = freq ? freq[n] : 1;
            int j = find_label(lab, A, k);
            if (j < 0 || v < 0 || v >= V || k < 0 || k >= K) return -1;
            n_wk[(size_t)v * ldk + k] += f;     /* LabeledLDA.py:92 */
            n_dk_act[lab_ptr[d] + j] += f;      /* LabeledLDA.py:91 */
            n_k[k] += f;                        /* LabeledLDA.py:90 */
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ exact (sequential, fp64) */
int oracle_llda_exact_sweep(int64_t d_begin, int64_t d_end, const int64_t *doc_ptr,
                            const int32_t *word, const int32_t *freq, int32_t *z,
                            const int64_t *lab_ptr, const int32_t *lab_idx,
                            int32_t K, int32_t V, int32_t ldk, double alpha, double beta,
                            int32_t *n_wk, int32_t *n_dk_act, int32_t *n_k,
                            uint64_t seed, uint32_t sweep, uint64_t doc_base) {
    double *cum = (double *)malloc(sizeof(double) * (size_t)(K > 0 ? K : 1));
    if (!cum) return -2;
    const double vbeta = (double)V * beta;                 /* LabeledLDA.py:115  self.V * self.beta */
    for (int64_t d = d_begin; d < d_end; ++d) {            /* LabeledLDA.py:106 */
        const int32_t *lab = lab_idx + lab_ptr[d];
        int32_t *ndk = n_dk_act + lab_ptr[d];              /* LabeledLDA.py:107 */
        int A = (int)(lab_ptr[d + 1] - lab_ptr[d]);
        for (int64_t n = doc_ptr[d]; n < doc_ptr[d + 1]; ++n) {   /* LabeledLDA.py:108 */
            int32_t v = word[n], f = freq ? freq[n] : 1, zo = z[n];
            int jo = find_label(lab, A, zo);
            if (jo < 0) { free(cum); return -1; }
            n_wk[(size_t)v * ldk + zo] -= f;               /* :109 */
            ndk[jo] -= f;                                  /* :110 */
            n_k[zo] -= f;                                  /* :111 */
            double run = 0.0;
            for (int j = 0; j < A; ++j) {
                int32_t k = lab[j];
                double a = (double)ndk[j] + alpha;                          /* :113 */
                double num = (double)n_wk[(size_t)v * ldk + k] + beta;      /* :114 */
                double den = (double)n_k[k] + vbeta;                        /* :115 */
                double q = num / den;
                double w = a * q;                                           /* :117  (lab*a)*(num_b/den_b), lab==1 */
                run = run + w;
                cum[j] = run;
            }
            /* :118-119 replaced: inverse CDF on the unnormalised weights (see patched_reference.py) */
            uint32_t x = oracle_draw_word(seed, 0u, sweep, doc_base + (uint64_t)d, (uint64_t)(n - doc_ptr[d]));
            double u = ((double)x + 0.5) * (1.0 / 4294967296.0);
            double thr = u * run;
            int jn = A - 1;
            for (int j = 0; j < A; ++j) if (cum[j] > thr) { jn = j; break; }
            int32_t zn = lab[jn];
            z[n] = zn;                                     /* :121 */
            n_wk[(size_t)v * ldk + zn] += f;               /* :123 */
            ndk[jn] += f;                                  /* :124 */
            n_k[zn] += f;                                  /* :125 */
        }
    }
    free(cum);
    return 0;
}

/* ------------------------------------------------------------------ snapshot (doc-parallel, fp32) */

/* Label lists of at most this many topics are summed serially (thread-per-document kernel); longer ones in
 * 32-lane Kogge-Stone chunks (group-per-document kernels).  Must equal GIBBS_SERIAL_MAX in llda_kernels.cuh. */
#define ORACLE_SERIAL_MAX 8

/* Inclusive scan of x[0..31] in the order a 32-lane Kogge-Stone __shfl_up scan produces. */
static void ks_scan32(float *x) {
    for (int off = 1; off < 32; off <<= 1) {
        float y[32];
        for (int j = 0; j < 32; ++j) y[j] = (j >= off) ? x[j] + x[j - off] : x[j];
        memcpy(x, y, sizeof(y));
    }
}

/* One document under the snapshot schedule.  frozen n_wk / n_k are read-only; updates go to delta. */
static int snapshot_doc(int64_t d, const int64_t *doc_ptr, const int32_t *word, const int32_t *freq,
                        int32_t *z, const int64_t *lab_ptr, const int32_t *lab_idx, int32_t ldk,
                        float alpha_f, float beta_f, float vbeta_f,
                        const int32_t *n_wk, int32_t *n_dk_act, const int32_t *n_k,
                        int32_t *delta_wk, int32_t *delta_k,
                        uint64_t seed, uint32_t sweep, uint64_t doc_base,
                        int32_t *nkb, float *cum) {
    const int32_t *lab = lab_idx + lab_ptr[d];
    int32_t *ndk = n_dk_act + lab_ptr[d];
    int A = (int)(lab_ptr[d + 1] - lab_ptr[d]);
    int nchunk = (A + 31) / 32;
    for (int j = 0; j < A; ++j) nkb[j] = n_k[lab[j]] - ndk[j];
    for (int64_t n = doc_ptr[d]; n < doc_ptr[d + 1]; ++n) {
        int32_t v = word[n], f = freq ? freq[n] : 1, zo = z[n];
        int jo = find_label(lab, A, zo);
        if (jo < 0) return -1;
        const int32_t *row = n_wk + (size_t)v * ldk;
        float carry = 0.0f, total = 0.0f;
        if (A <= ORACLE_SERIAL_MAX) {
            /* short label lists: one thread owns the document and adds the weights left to right */
            float run = 0.0f;
            for (int j = 0; j < A; ++j) {
                int32_t self = (j == jo) ? f : 0;
                int32_t nd = ndk[j] - self;
                int32_t nw = row[lab[j]] - self;
                float a = (float)nd + alpha_f;
                float b = (float)nw + beta_f;
                float cc = (float)(nkb[j] + nd) + vbeta_f;
                float ab = a * b;
                float w = ab / cc;
                run = run + w;
                cum[j] = run;
            }
            total = run;
        } else
        for (int c = 0; c < nchunk; ++c) {
            float x[32];
            for (int l = 0; l < 32; ++l) {
                int j = c * 32 + l;
                if (j < A) {
                    int32_t self = (j == jo) ? f : 0;
                    int32_t nd = ndk[j] - self;
                    int32_t nw = row[lab[j]] - self;
                    float a = (float)nd + alpha_f;
                    float b = (float)nw + beta_f;
                    float cc = (float)(nkb[j] + nd) + vbeta_f;
                    float ab = a * b;
                    x[l] = ab / cc;
                } else {
                    x[l] = 0.0f;
                }
            }
            ks_scan32(x);
            for (int l = 0; l < 32; ++l) {
                int j = c * 32 + l;
                if (j < A) cum[j] = carry + x[l];
            }
            if (c == nchunk - 1) total = cum[A - 1];
            carry = carry + x[31];
        }
        uint32_t xw = oracle_draw_word(seed, 0u, sweep, doc_base + (uint64_t)d, (uint64_t)(n - doc_ptr[d]));
        float u = (float)(xw >> 8) * (1.0f / 16777216.0f);
        float thr = u * total;
        int jn = A - 1;
        for (int j = 0; j < A; ++j) if (cum[j] > thr) { jn = j; break; }
        if (jn != jo) {
            int32_t zn = lab[jn];
            z[n] = zn;
            ndk[jo] -= f;
            ndk[jn] += f;
#pragma omp atomic
            delta_wk[(size_t)v * ldk + zo] -= f;
#pragma omp atomic
            delta_wk[(size_t)v * ldk + zn] += f;
#pragma omp atomic
            delta_k[zo] -= f;
#pragma omp atomic
            delta_k[zn] += f;
        }
    }
    return 0;
}

int oracle_openmp_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 0;
#endif
}

int oracle_llda_snapshot_sweep(int64_t n_tiles, const int64_t *tile_rng, int32_t n_blocks,
                               const int64_t *doc_ptr, const int32_t *word, const int32_t *freq,
                               int32_t *z, const int64_t *lab_ptr, const int32_t *lab_idx,
                               int32_t K, int32_t V, int32_t ldk, double alpha, double beta,
                               int32_t *n_wk, int32_t *n_dk_act, int32_t *n_k,
                               uint64_t seed, uint32_t sweep, uint64_t doc_base, int32_t n_threads) {
    if (n_blocks < 1) n_blocks = 1;
    if (n_threads < 1) n_threads = 1;
    const float alpha_f = (float)alpha, beta_f = (float)beta, vbeta_f = (float)((double)V * beta);
    size_t tab = (size_t)V * (size_t)ldk;
    int32_t *delta_wk = (int32_t *)calloc(tab, sizeof(int32_t));
    int32_t *delta_k = (int32_t *)calloc((size_t)K, sizeof(int32_t));
    if (!delta_wk || !delta_k) { free(delta_wk); free(delta_k); return -2; }
    int err = 0;
    for (int32_t b = 0; b < n_blocks; ++b) {
#pragma omp parallel num_threads(n_threads)
        {
            int32_t *nkb = (int32_t *)malloc(sizeof(int32_t) * (size_t)K);
            float *cum = (float *)malloc(sizeof(float) * (size_t)K);
#pragma omp for schedule(dynamic, 1)
            for (int64_t i = b; i < n_tiles; i += n_blocks) {
                for (int64_t d = tile_rng[2 * i]; d < tile_rng[2 * i + 1]; ++d) {
                    int r = snapshot_doc(d, doc_ptr, word, freq, z, lab_ptr, lab_idx, ldk,
                                         alpha_f, beta_f, vbeta_f, n_wk, n_dk_act, n_k,
                                         delta_wk, delta_k, seed, sweep, doc_base, nkb, cum);
                    if (r) {
#pragma omp atomic write
                        err = r;
                    }
                }
            }
            free(nkb); free(cum);
        }
        /* refresh: fold the block's deltas into the tables every later block reads */
        for (size_t i = 0; i < tab; ++i) { n_wk[i] += delta_wk[i]; delta_wk[i] = 0; }
        for (int32_t k = 0; k < K; ++k) { n_k[k] += delta_k[k]; delta_k[k] = 0; }
    }
    free(delta_wk); free(delta_k);
    return err;
}

/* ------------------------------------------------------------------ frozen-phi test chains (fp64) */

static void ks_scan32_f64(double *x) {
    for (int off = 1; off < 32; off <<= 1) {
        double y[32];
        for (int j = 0; j < 32; ++j) y[j] = (j >= off) ? x[j] + x[j - off] : x[j];
        memcpy(x, y, sizeof(y));
    }
}

/* cum[] = running sums of w[0..A) in 32-wide Kogge-Stone chunks with a carry (the device order); returns cum[A-1]. */
static double chunk_cumsum_f64(const double *w, int A, double *cum) {
    double carry = 0.0;
    int nch = (A + 31) / 32;
    for (int c = 0; c < nch; ++c) {
        double x[32];
        for (int l = 0; l < 32; ++l) { int j = c * 32 + l; x[l] = (j < A) ? w[j] : 0.0; }
        ks_scan32_f64(x);
        for (int l = 0; l < 32; ++l) { int j = c * 32 + l; if (j < A) cum[j] = carry + x[l]; }
        carry = carry + x[31];
    }
    return cum[A - 1];
}

static int pick_f64(const double *cum, int A, double thr) {
    for (int j = 0; j < A; ++j) if (cum[j] > thr) return j;
    return A - 1;
}

/* LabeledLDA.py:155-212 (prep4test + run_test) and CascadeLDA.py:186-247 (prep4test + cascade_test), one chain per
 * (document, topic list); the draw at :172/:194 (:203/:233) is the inverse-CDF draw on the unnormalised weights.
 * init_mode 0: z given (global topic ids); 1: LabeledLDA.prep4test; 2: CascadeLDA.prep4test.  Streams 4 (start
 * state) and 2 (chain), counter (iteration, chain_base + chain, position). */
int oracle_test_chains(int32_t K, int32_t V, const double *phi_KV, double alpha, double beta_fb, int64_t n_chains,
                       const int64_t *doc_ptr, const int32_t *word, const int32_t *freq,
                       const int64_t *lab_ptr, const int32_t *lab_idx, int32_t *z, int32_t init_mode,
                       int32_t it, int32_t thinning, uint64_t seed, int64_t chain_base, double *th_hat) {
    double *w = (double *)malloc(sizeof(double) * (size_t)K);
    double *cum = (double *)malloc(sizeof(double) * (size_t)K);
    int32_t *ndk = (int32_t *)malloc(sizeof(int32_t) * (size_t)K);
    int32_t *lab = (int32_t *)malloc(sizeof(int32_t) * (size_t)K);
    int rc = 0;
    if (!w || !cum || !ndk || !lab) { rc = -2; goto done; }
    for (int64_t c = 0; c < n_chains && !rc; ++c) {
        int64_t n0 = doc_ptr[c];
        int len = (int)(doc_ptr[c + 1] - n0);
        int64_t l0 = lab_ptr ? lab_ptr[c] : c * (int64_t)K;
        int A = lab_ptr ? (int)(lab_ptr[c + 1] - lab_ptr[c]) : K;
        for (int j = 0; j < A; ++j) { lab[j] = lab_ptr ? lab_idx[l0 + j] : j; ndk[j] = 0; th_hat[l0 + j] = 0.0; }
        int32_t *zc = z + n0;                                 /* list indices while the chain runs */
        for (int n = 0; n < len; ++n) {                       /* start state */
            int32_t v = word[n0 + n], f = freq ? freq[n0 + n] : 1;
            int jn;
            if (init_mode == 0) {
                jn = find_label(lab, A, zc[n]);
                if (jn < 0) { rc = -1; break; }
            } else {
                double total;
                if (init_mode == 1) {
                    for (int j = 0; j < A; ++j) w[j] = phi_KV[(size_t)lab[j] * V + v];          /* LabeledLDA.py:162,169 */
                    total = chunk_cumsum_f64(w, A, cum);
                } else {
                    for (int j = 0; j < A; ++j) w[j] = phi_KV[(size_t)lab[j] * V + v] + beta_fb; /* CascadeLDA.py:194-195 */
                    double colsum = chunk_cumsum_f64(w, A, cum);
                    double first = 1.0 / (double)len;                                             /* :198 */
                    for (int j = 0; j < A; ++j) w[j] = (j == 0) ? first : w[j] / colsum;          /* :196 */
                    total = chunk_cumsum_f64(w, A, cum);
                }
                uint32_t x = oracle_draw_word(seed, 4u, 0u, (uint64_t)(chain_base + c), (uint64_t)n);
                double u = ((double)x + 0.5) * (1.0 / 4294967296.0);
                jn = pick_f64(cum, A, u * total);
            }
            zc[n] = jn;
            ndk[jn] += f;                                     /* LabeledLDA.py:174-175 */
        }
        for (int i = 0; i < it && !rc; ++i) {                 /* LabeledLDA.py:184 */
            for (int n = 0; n < len; ++n) {
                int32_t v = word[n0 + n], f = freq ? freq[n0 + n] : 1;
                int jo = zc[n];
                ndk[jo] -= f;                                 /* :186 */
                for (int j = 0; j < A; ++j) {
                    double a = (double)ndk[j] + alpha;        /* :188 */
                    w[j] = a * phi_KV[(size_t)lab[j] * V + v];/* :189-190 */
                }
                double total = chunk_cumsum_f64(w, A, cum);
                if (total == 0.0 && beta_fb > 0.0) {          /* CascadeLDA.py:225-230 */
                    for (int j = 0; j < A; ++j) {
                        double a = (double)ndk[j] + alpha;
                        double b = phi_KV[(size_t)lab[j] * V + v] + beta_fb;
                        w[j] = a * b;
                    }
                    total = chunk_cumsum_f64(w, A, cum);
                }
                uint32_t x = oracle_draw_word(seed, 2u, (uint32_t)i, (uint64_t)(chain_base + c), (uint64_t)n);
                double u = ((double)x + 0.5) * (1.0 / 4294967296.0);
                int jn = pick_f64(cum, A, u * total);
                zc[n] = jn;                                   /* :196 */
                ndk[jn] += f;                                 /* :197 */
            }
            if ((i + 1) % thinning == 0) {                    /* :201-211 */
                int s2 = (i + 1) / thinning;
                long long sum = 0;
                for (int j = 0; j < A; ++j) sum += ndk[j];
                double den = (double)sum;
                double c_old = (double)(s2 - 1) / (double)s2, c_new = 1.0 / (double)s2;
                for (int j = 0; j < A; ++j) {
                    double cur = (double)ndk[j] / den;
                    if (s2 > 1) {
                        double o = c_old * th_hat[l0 + j];
                        double nw = c_new * cur;
                        th_hat[l0 + j] = o + nw;
                    } else th_hat[l0 + j] = cur;
                }
            }
        }
        for (int n = 0; n < len; ++n) zc[n] = lab[zc[n]];
    }
done:
    free(w); free(cum); free(ndk); free(lab);
    return rc;
}
