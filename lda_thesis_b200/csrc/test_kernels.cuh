// Frozen-phi test chains (LabeledLDA.py:155-212 prep4test + run_test, CascadeLDA.py:186-247 prep4test + cascade_test).
// sm_100a only.  One warp per chain, fp64 throughout (the reference computes these in float64 and the chains are
// short); every chain is independent, so there are no atomics and no exchange step.
//
// A chain is one (test document, topic list) pair: the topic list is all K topics for LabeledLDA.run_test and the
// `[parent, children...]` rows of one tree node for CascadeLDA.cascade_test.  Lane l owns entries l, l+32, ... of
// the list; per draw the weights (n_dk + alpha) * phi[topic][v] are prefix-summed in 32-lane Kogge-Stone chunks
// with a running carry (the order oracle/gibbs_oracle.c:oracle_test_chains restates) and the draw is the first
// entry whose running sum exceeds u * total.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "philox.cuh"

#define GIBBS_TEST_INIT_GIVEN    0   // z_init supplied by the caller (global topic ids)
#define GIBBS_TEST_INIT_LLDA     1   // LabeledLDA.py:162-175: z ~ phi[:, v] / sum
#define GIBBS_TEST_INIT_CASCADE  2   // CascadeLDA.py:194-206: z ~ (phi[:, v] + beta) / sum with entry 0 set to 1 / len(doc)

struct TestParams {
    const double *phiT;          // [V][ldk] word-major copy of ph_hat[K][V]
    int ldk, K;
    long long n_chains;
    const long long *doc_ptr;    // [n_chains + 1]
    const int *word, *freq;      // freq may be nullptr (all ones)
    const long long *lab_ptr;    // [n_chains + 1] or nullptr: every chain uses topics 0..K-1
    const int *lab_idx;
    int *z;                      // [N] in: global topic ids (INIT_GIVEN); out: global topic ids
    double *th_out;              // aligned with lab_idx, or [n_chains][K] when lab_ptr == nullptr
    double alpha, beta_fb;       // beta_fb > 0: CascadeLDA.py:225-230 zero-mass fallback
    int it, thinning, init_mode;
    unsigned seed_lo, seed_hi;
    long long chain_base;        // RNG address of chain 0
    int a_cap;                   // shared-memory slots per warp (multiple of 32, >= longest topic list)
    int *err;
};

__device__ __forceinline__ double ks_scan32_f64(double x, const int lane) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x = __dadd_rn(x, y);
    }
    return x;
}

// cum[j] for the whole list; returns cum[A-1].  `wfun(j)` gives the weight of entry j.
template <typename F>
__device__ __forceinline__ double test_scan_list(const int A, const int lane, double *cum, F wfun) {
    double carry = 0.0;
    const int nch = (A + 31) >> 5;
    for (int c = 0; c < nch; ++c) {
        const int j = c * 32 + lane;
        double x = (j < A) ? wfun(j) : 0.0;
        x = ks_scan32_f64(x, lane);
        if (j < A) cum[j] = __dadd_rn(carry, x);
        carry = __dadd_rn(carry, __shfl_sync(0xffffffffu, x, 31));
    }
    __syncwarp();
    return cum[A - 1];
}

// first j with cum[j] > thr, else A-1
__device__ __forceinline__ int test_pick(const int A, const int lane, const double *cum, const double thr) {
    const int nch = (A + 31) >> 5;
    for (int c = 0; c < nch; ++c) {
        const int j = c * 32 + lane;
        const unsigned bal = __ballot_sync(0xffffffffu, (j < A) && (cum[j] > thr));
        if (bal) return c * 32 + __ffs(bal) - 1;
    }
    return A - 1;
}

__global__ void __launch_bounds__(256) test_chain_kernel(const TestParams p) {
    extern __shared__ __align__(16) unsigned char test_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const size_t per_warp = (size_t)p.a_cap * 16;
    double *cum = reinterpret_cast<double *>(test_smem + warp * per_warp);
    int *ndk = reinterpret_cast<int *>(cum + p.a_cap);
    int *lab = ndk + p.a_cap;
    const uint2 key = make_uint2(p.seed_lo, p.seed_hi);
    const double alpha = p.alpha;

    for (long long chain = (long long)blockIdx.x * wpc + warp; chain < p.n_chains; chain += (long long)gridDim.x * wpc) {
        const long long n0 = p.doc_ptr[chain];
        const int len = (int)(p.doc_ptr[chain + 1] - n0);
        long long l0;
        int A;
        if (p.lab_ptr) { l0 = p.lab_ptr[chain]; A = (int)(p.lab_ptr[chain + 1] - l0); }
        else           { l0 = chain * (long long)p.K; A = p.K; }
        __syncwarp();
        for (int j = lane; j < A; j += 32) { lab[j] = p.lab_idx ? p.lab_idx[l0 + j] : j; ndk[j] = 0; }
        __syncwarp();
        const unsigned docg = (unsigned)(p.chain_base + chain);

        // ---- start state
        for (int n = 0; n < len; ++n) {
            const int v = p.word[n0 + n];
            const int f = p.freq ? p.freq[n0 + n] : 1;
            const double *col = p.phiT + (size_t)v * p.ldk;
            int jn;
            if (p.init_mode == GIBBS_TEST_INIT_GIVEN) {
                const int zg = p.z[n0 + n];
                jn = -1;
                const int nch = (A + 31) >> 5;
                for (int c = 0; c < nch && jn < 0; ++c) {
                    const int j = c * 32 + lane;
                    const unsigned bal = __ballot_sync(0xffffffffu, (j < A) && (lab[j] == zg));
                    if (bal) jn = c * 32 + __ffs(bal) - 1;
                }
                if (jn < 0) { if (lane == 0) atomicExch(p.err, 1); jn = 0; }
            } else {
                double total;
                if (p.init_mode == GIBBS_TEST_INIT_LLDA) {
                    total = test_scan_list(A, lane, cum, [&](int j) { return col[lab[j]]; });
                } else {
                    const double bfb = p.beta_fb;
                    const double colsum = test_scan_list(A, lane, cum, [&](int j) { return __dadd_rn(col[lab[j]], bfb); });
                    __syncwarp();
                    const double first = __ddiv_rn(1.0, (double)len);
                    total = test_scan_list(A, lane, cum, [&](int j) {
                        return j == 0 ? first : __ddiv_rn(__dadd_rn(col[lab[j]], bfb), colsum);
                    });
                }
                const uint32_t xw = philox_word(docg, (unsigned)n, 0u, GIBBS_STREAM_TEST_INIT, key);
                jn = test_pick(A, lane, cum, __dmul_rn(u01_f64(xw), total));
            }
            __syncwarp();
            if (lane == 0) { p.z[n0 + n] = jn; ndk[jn] += f; }       // z holds list indices while the chain runs
            __syncwarp();
        }

        // ---- the chain (LabeledLDA.py:184-197, CascadeLDA.py:216-236)
        for (int i = 0; i < p.it; ++i) {
            int nv = 0, nf = 1, nz = 0;                               // draw n + 1, requested one draw ahead
            if (len > 0) { nv = p.word[n0]; nf = p.freq ? p.freq[n0] : 1; nz = p.z[n0]; }
            for (int n = 0; n < len; ++n) {
                const int v = nv, f = nf, jo = nz;
                if (n + 1 < len) { nv = p.word[n0 + n + 1]; nf = p.freq ? p.freq[n0 + n + 1] : 1; nz = p.z[n0 + n + 1]; }
                const double *col = p.phiT + (size_t)v * p.ldk;
                if (lane == 0) ndk[jo] -= f;                          // :186
                __syncwarp();
                double total = test_scan_list(A, lane, cum, [&](int j) {
                    return __dmul_rn(__dadd_rn((double)ndk[j], alpha), col[lab[j]]);   // :188-190
                });
                if (total == 0.0 && p.beta_fb > 0.0) {                // CascadeLDA.py:225-230
                    const double bfb = p.beta_fb;
                    __syncwarp();
                    total = test_scan_list(A, lane, cum, [&](int j) {
                        return __dmul_rn(__dadd_rn((double)ndk[j], alpha), __dadd_rn(col[lab[j]], bfb));
                    });
                }
                const uint32_t xw = philox_word(docg, (unsigned)n, (unsigned)i, GIBBS_STREAM_TEST, key);
                const int jn = test_pick(A, lane, cum, __dmul_rn(u01_f64(xw), total));
                __syncwarp();
                if (lane == 0) { p.z[n0 + n] = jn; ndk[jn] += f; }    // :196-197
                __syncwarp();
            }
            // thinning mean of n_dk / sum(n_dk)  (LabeledLDA.py:201-211)
            if ((i + 1) % p.thinning == 0) {
                const int s2 = (i + 1) / p.thinning;
                long long part = 0;
                for (int j = lane; j < A; j += 32) part += ndk[j];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
                const double den = (double)part;
                const double c_old = __ddiv_rn((double)(s2 - 1), (double)s2), c_new = __ddiv_rn(1.0, (double)s2);
                for (int j = lane; j < A; j += 32) {
                    const double cur = __ddiv_rn((double)ndk[j], den);
                    double out = cur;
                    if (s2 > 1) out = __dadd_rn(__dmul_rn(c_old, p.th_out[l0 + j]), __dmul_rn(c_new, cur));
                    p.th_out[l0 + j] = out;
                }
            }
        }
        // z back to global topic ids
        __syncwarp();
        for (int n = lane; n < len; n += 32) p.z[n0 + n] = lab[p.z[n0 + n]];
    }
}

// phiT[v][k] = phi[k][v]  (32 x 32 tile transpose through shared memory)
__global__ void transpose_phi_kernel(const double *__restrict__ phi, double *__restrict__ phiT, int K, int V, int ldk) {
    __shared__ double tile[32][33];
    const int v0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int k = k0 + r, v = v0 + threadIdx.x;
        tile[r][threadIdx.x] = (k < K && v < V) ? phi[(size_t)k * V + v] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int v = v0 + r, k = k0 + threadIdx.x;
        if (v < V && k < ldk) phiT[(size_t)v * ldk + k] = tile[threadIdx.x][r];
    }
}
