// Device kernels for the Labeled-LDA family sweep (LabeledLDA.py:101-125, CascadeLDA.py:397-421).
// sm_100a only.  See DESIGN.md for the data layout and the per-kernel byte model.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "philox.cuh"

// ----------------------------------------------------------------------------------------------
// Draw record: 8 bytes per draw.  x = word id, y = (f << 16) | j  where j is the index of the
// draw's current topic inside its document's label list (z = lab_idx[lab_ptr[d] + j]).
// ----------------------------------------------------------------------------------------------
#define REC_F(y)  ((int)((unsigned)(y) >> 16))
#define REC_J(y)  ((int)((unsigned)(y) & 0xffffu))
#define REC_PACK(f, j) ((int)(((unsigned)(f) << 16) | (unsigned)(j)))

struct SweepParams {
    const long long *doc_ptr;     // [D+1]
    const long long *lab_ptr;     // [D+1]
    const int       *lab_idx;     // [lab_ptr[D]]
    int             *n_dk_act;    // [lab_ptr[D]]
    int2            *rec;         // [N]
    const int       *n_wk;        // [V][ldk] frozen for the refresh block
    int             *delta_wk;    // [V][ldk] +-f land here
    const int       *n_k;         // [K] frozen
    const int       *seg;         // [2*D] or nullptr
    const int       *doc_list;    // documents of this launch
    long long        n_list;
    unsigned long long *counter;  // work counter (zeroed before launch)
    unsigned long long *changed;  // draws whose topic changed
    int              ldk;
    int              row_ints;    // ints per ring slot (max segment length, multiple of 4)
    float            alpha, beta, vbeta;
    unsigned         seed_lo, seed_hi, sweep;
    long long        draw_base;
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ----------------------------------------------------------------------------------------------
// Snapshot sweep.  A group of G lanes owns one document at a time; lane gl (+32*c for chunk c when
// G == 32) owns one entry of the document's label list.  Per draw:
//   cp.async the draw's n_wk row (segment) into a shared-memory ring R-1 draws ahead,
//   gather the active entries from shared memory, weight = (n_dk+alpha)*(n_wk+beta)/(n_k+V*beta) with the
//   draw's own f removed, Kogge-Stone prefix sum over the group, Philox uniform, first cum > u*total,
//   +-f to the delta table with RED, n_dk stays in registers until the document ends.
// Arithmetic is restated operation for operation by oracle/gibbs_oracle.c:snapshot_doc.
// ----------------------------------------------------------------------------------------------
template <int G, int NCH, int R>
__global__ void __launch_bounds__(256) llda_snapshot_kernel(const SweepParams p) {
    static_assert(G == 32 || NCH == 1, "multi-chunk label lists need a full warp");
    static_assert((R & (R - 1)) == 0 && R >= 2, "ring depth must be a power of two");
    constexpr int P = 2 * R - 1;   // record prefetch distance
    constexpr int M = 2 * R;       // record ring slots
    constexpr int GPW = 32 / G;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);
    const int gbase = lane & ~(G - 1);
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gbase);
    const int grp = (threadIdx.x >> 5) * GPW + (lane / G);
    const int row_ints = p.row_ints;
    const size_t grp_bytes = (size_t)M * 8 + (size_t)R * row_ints * 4;
    unsigned char *gsm = smem_raw + (size_t)grp * grp_bytes;
    int2 *meta = reinterpret_cast<int2 *>(gsm);
    int *ring = reinterpret_cast<int *>(gsm + M * 8);

    const float alpha = p.alpha, beta = p.beta, vbeta = p.vbeta;
    const uint2 key = make_uint2(p.seed_lo, p.seed_hi);
    const int ldk = p.ldk;

    // group state
    long long n = 0, n_end = 0, lab0 = 0;
    int A = 0, i = 0, seg_lo = 0, seg_n16 = 0;
    int lab[NCH], ndk[NCH], nkb[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) { lab[c] = 0; ndk[c] = 0; nkb[c] = 0; }
    uint4 rw = make_uint4(0, 0, 0, 0);
    long long rB0 = -(1ll << 40);
    bool have = false, done = false;
    unsigned n_changed = 0;

    auto issue_row = [&](int slot, int v) {
        const int *src = p.n_wk + (size_t)v * ldk + seg_lo;
        int *dst = ring + (size_t)slot * row_ints;
        for (int c16 = gl; c16 < seg_n16; c16 += G) cp_async16(dst + c16 * 4, src + c16 * 4);
    };

    while (true) {
        if (!have && !done) {
            unsigned long long di = 0;
            if (gl == 0) di = atomicAdd(p.counter, 1ull);
            di = __shfl_sync(gmask, di, gbase);
            if (di >= (unsigned long long)p.n_list) {
                done = true;
            } else {
                const int d = p.doc_list[di];
                n = p.doc_ptr[d];
                n_end = p.doc_ptr[d + 1];
                lab0 = p.lab_ptr[d];
                A = (int)(p.lab_ptr[d + 1] - lab0);
                if (p.seg) { seg_lo = p.seg[2 * d]; seg_n16 = (p.seg[2 * d + 1] - seg_lo) >> 2; }
                else       { seg_lo = 0; seg_n16 = ldk >> 2; }
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const int j = c * G + gl;
                    if (j < A) {
                        lab[c] = p.lab_idx[lab0 + j];
                        ndk[c] = p.n_dk_act[lab0 + j];
                        nkb[c] = p.n_k[lab[c]] - ndk[c];
                    } else { lab[c] = seg_lo; ndk[c] = 0; nkb[c] = 0; }
                }
                if (n < n_end) {
                    have = true;
                    i = 0;
                    for (int q = gl; q < P; q += G)
                        if (n + q < n_end) cp_async8(&meta[q & (M - 1)], &p.rec[n + q]);
                    cp_async_commit();
                    cp_async_wait<0>();
                    __syncwarp(gmask);
#pragma unroll
                    for (int r = 0; r < R - 1; ++r) {
                        if (n + r < n_end) issue_row(r, meta[r].x);
                        cp_async_commit();
                    }
                }
            }
        }
        if (__all_sync(0xffffffffu, done)) break;
        if (have) {
            // -- prefetch: record of draw n+P, row of draw n+R-1
            if (gl == 0 && n + P < n_end) cp_async8(&meta[(i + P) & (M - 1)], &p.rec[n + P]);
            if (n + (R - 1) < n_end) issue_row((i + R - 1) & (R - 1), meta[(i + R - 1) & (M - 1)].x);
            cp_async_commit();
            cp_async_wait<R - 1>();
            __syncwarp(gmask);

            const int2 mt = meta[i & (M - 1)];
            const int v = mt.x, f = REC_F(mt.y), jo = REC_J(mt.y);
            const int *row = ring + (size_t)(i & (R - 1)) * row_ints - seg_lo;

            // -- weights + prefix sums (chunk by chunk; carry is the running total)
            float cum[NCH];
            float carry = 0.0f;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int j = c * G + gl;
                float x = 0.0f;
                if (j < A) {
                    const int self = (j == jo) ? f : 0;
                    const int nd = ndk[c] - self;
                    const int nw = row[lab[c]] - self;
                    const float a = __fadd_rn((float)nd, alpha);
                    const float b = __fadd_rn((float)nw, beta);
                    const float cc = __fadd_rn((float)(nkb[c] + nd), vbeta);
                    x = __fdiv_rn(__fmul_rn(a, b), cc);
                }
#pragma unroll
                for (int off = 1; off < G; off <<= 1) {
                    const float y = __shfl_up_sync(gmask, x, off, G);
                    if (gl >= off) x = __fadd_rn(x, y);
                }
                cum[c] = __fadd_rn(carry, x);
                carry = __fadd_rn(carry, __shfl_sync(gmask, x, G - 1, G));
            }
            // total = cum[A-1] (NOT the last lane: a Kogge-Stone lane past A-1 sums the same terms in another order)
            float cl = cum[0];
#pragma unroll
            for (int c = 1; c < NCH; ++c) if (((A - 1) / G) == c) cl = cum[c];
            const float total = __shfl_sync(gmask, cl, (A - 1) & (G - 1), G);

            // -- uniform
            const long long t = p.draw_base + n;
            const long long blk = t >> 2;
            if (blk < rB0 || blk >= rB0 + G) {
                rB0 = blk;
                rw = philox_block((uint64_t)(blk + gl), p.sweep, GIBBS_STREAM_SWEEP, key);
            }
            const unsigned mine = select_word(rw, (unsigned)(t & 3));
            const unsigned xw = __shfl_sync(gmask, mine, (int)(blk - rB0), G);
            const float thr = __fmul_rn(u01_f32(xw), total);

            // -- first index with cum > thr, else A-1
            int jn = A - 1;
            bool found = false;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int j = c * G + gl;
                unsigned bal = __ballot_sync(gmask, (j < A) && (cum[c] > thr));
                bal >>= gbase;
                if (!found && bal) { jn = c * G + __ffs(bal) - 1; found = true; }
            }

            if (jn != jo) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const int j = c * G + gl;
                    if (j == jo) { ndk[c] -= f; atomicAdd(&p.delta_wk[(size_t)v * ldk + lab[c]], -f); }
                    else if (j == jn) { ndk[c] += f; atomicAdd(&p.delta_wk[(size_t)v * ldk + lab[c]], f); }
                }
                if (gl == 0) { p.rec[n].y = REC_PACK(f, jn); ++n_changed; }
            }
            ++n; ++i;
            if (n == n_end) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const int j = c * G + gl;
                    if (j < A) p.n_dk_act[lab0 + j] = ndk[c];
                }
                have = false;
            }
        }
    }
    if (gl == 0 && n_changed) atomicAdd(p.changed, (unsigned long long)n_changed);
}

// ----------------------------------------------------------------------------------------------
// Snapshot sweep, masked-gather row fetch (label lists of at most G <= 32 topics).
//
// A group of G lanes owns one document; lane j owns entry j of its label list and keeps n_dk[j] in a
// register for the whole document.  The document is walked in chunks of G draws:
//   * lane i of the group holds the record of draw n0+i (one coalesced 8-byte load per lane per chunk,
//     the next chunk's records are requested one chunk ahead) and computes that draw's Philox word;
//   * per draw, lane j needs ONE count, n_wk[v][lab_j]: a 4-byte load that touches |label list| 32-byte
//     sectors of the row instead of the whole ldk-wide row.  The loads run R draws ahead in a register
//     ring, so R x (warps per SM) independent sectors are in flight per lane;
//   * weight, Kogge-Stone scan, threshold, ballot exactly as in the dense-row kernel -- the arithmetic
//     (and therefore the oracle restatement, oracle/gibbs_oracle.c:snapshot_doc) is the same;
//   * +-f go to the delta table with RED.ADD; changed records are written back once per chunk.
// No shared memory, no barriers: occupancy is bounded by registers only.
// ----------------------------------------------------------------------------------------------
template <int G, int R>
__global__ void __launch_bounds__(256, 3) llda_gather_kernel(const SweepParams p) {
    static_assert(G >= 4 && G <= 32 && (G & (G - 1)) == 0, "group width");
    static_assert(R >= 1 && R <= G && (G % R) == 0, "ring depth must divide the chunk");
    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);
    const int gbase = lane & ~(G - 1);
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gbase);
    const float alpha = p.alpha, beta = p.beta, vbeta = p.vbeta;
    const uint2 key = make_uint2(p.seed_lo, p.seed_hi);
    const int ldk = p.ldk;
    const int *__restrict__ n_wk = p.n_wk;
    unsigned n_changed = 0;

    while (true) {
        // ---- next document of this group
        unsigned long long di = 0;
        if (gl == 0) di = atomicAdd(p.counter, 1ull);
        di = __shfl_sync(gmask, di, 0, G);
        if (di >= (unsigned long long)p.n_list) break;
        const int d = p.doc_list[di];
        long long n0 = p.doc_ptr[d];
        const long long n_end = p.doc_ptr[d + 1];
        const long long lab0 = p.lab_ptr[d];
        const int A = (int)(p.lab_ptr[d + 1] - lab0);
        if (n0 >= n_end) continue;
        int lab = 0, ndk = 0, nkb = 0;
        if (gl < A) {
            lab = p.lab_idx[lab0 + gl];
            ndk = p.n_dk_act[lab0 + gl];
            nkb = p.n_k[lab] - ndk;
        }
        int2 cur = make_int2(0, 0), nxt = make_int2(0, 0);
        if (n0 + gl < n_end) cur = p.rec[n0 + gl];
        if (n0 + G + gl < n_end) nxt = p.rec[n0 + G + gl];
        int nwq[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int v = __shfl_sync(gmask, cur.x, r, G);
            nwq[r] = (gl < A && n0 + r < n_end) ? __ldg(n_wk + (size_t)v * ldk + lab) : 0;
        }

        // ---- chunks of G draws
        while (true) {
            const unsigned word = philox_word((uint64_t)(p.draw_base + n0 + gl), p.sweep, GIBBS_STREAM_SWEEP, key);
            int newy = cur.y;
            for (int i0 = 0; i0 < G; i0 += R) {
                if (n0 + i0 >= n_end) break;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int i = i0 + r;
                    const bool live = n0 + i < n_end;
                    const int v = __shfl_sync(gmask, cur.x, i, G);
                    const int y = __shfl_sync(gmask, cur.y, i, G);
                    const unsigned xw = __shfl_sync(gmask, word, i, G);
                    const int f = REC_F(y), jo = REC_J(y);
                    const int nw_raw = nwq[r];
                    {   // refill the ring slot with draw i + R (this chunk or the next one)
                        const int ii = i + R;
                        const int srcx = (ii < G) ? cur.x : nxt.x;
                        const int v2 = __shfl_sync(gmask, srcx, ii & (G - 1), G);
                        nwq[r] = (gl < A && n0 + ii < n_end) ? __ldg(n_wk + (size_t)v2 * ldk + lab) : 0;
                    }
                    float x = 0.0f;
                    const int self = (gl == jo) ? f : 0;
                    const int nd = ndk - self;
                    if (gl < A) {
                        const float a = __fadd_rn((float)nd, alpha);
                        const float b = __fadd_rn((float)(nw_raw - self), beta);
                        const float cc = __fadd_rn((float)(nkb + nd), vbeta);
                        x = __fdiv_rn(__fmul_rn(a, b), cc);
                    }
#pragma unroll
                    for (int off = 1; off < G; off <<= 1) {
                        const float t = __shfl_up_sync(gmask, x, off, G);
                        if (gl >= off) x = __fadd_rn(x, t);
                    }
                    const float total = __shfl_sync(gmask, x, A - 1, G);
                    const float thr = __fmul_rn(u01_f32(xw), total);
                    unsigned bal = __ballot_sync(gmask, (gl < A) && (x > thr));
                    bal = (bal >> gbase) & ((G == 32) ? 0xffffffffu : ((1u << G) - 1u));
                    const int jn = bal ? (__ffs(bal) - 1) : (A - 1);
                    if (live && jn != jo) {
                        if (gl == jo) { ndk -= f; atomicAdd(&p.delta_wk[(size_t)v * ldk + lab], -f); }
                        else if (gl == jn) { ndk += f; atomicAdd(&p.delta_wk[(size_t)v * ldk + lab], f); }
                        if (gl == i) { newy = REC_PACK(f, jn); ++n_changed; }
                    }
                }
            }
            if (newy != cur.y) p.rec[n0 + gl].y = newy;
            n0 += G;
            if (n0 >= n_end) break;
            cur = nxt;
            nxt = make_int2(0, 0);
            if (n0 + G + gl < n_end) nxt = p.rec[n0 + G + gl];
        }
        if (gl < A) p.n_dk_act[lab0 + gl] = ndk;
    }
    n_changed = __reduce_add_sync(0xffffffffu, n_changed);
    if (lane == 0 && n_changed) atomicAdd(p.changed, (unsigned long long)n_changed);
}

// ----------------------------------------------------------------------------------------------
// Exact sweep: one warp walks the corpus in order with live counts in fp64, the operation order of
// LabeledLDA.py:109-125 (restated by oracle/gibbs_oracle.c:oracle_llda_exact_sweep).
// ----------------------------------------------------------------------------------------------
struct ExactParams {
    const long long *doc_ptr, *lab_ptr;
    const int *lab_idx;
    int *n_dk_act;
    int2 *rec;
    int *n_wk;
    int *n_k;
    long long d_begin, d_end;
    int ldk;
    double alpha, beta, vbeta;
    unsigned seed_lo, seed_hi, sweep;
    long long draw_base;
    unsigned long long *changed;
};

__global__ void __launch_bounds__(32) llda_exact_kernel(const ExactParams p) {
    const int lane = threadIdx.x;
    volatile int *nwk = p.n_wk;
    volatile int *nk = p.n_k;
    volatile int *ndk = p.n_dk_act;
    const uint2 key = make_uint2(p.seed_lo, p.seed_hi);
    unsigned long long n_changed = 0;
    for (long long d = p.d_begin; d < p.d_end; ++d) {
        const long long lab0 = p.lab_ptr[d];
        const int A = (int)(p.lab_ptr[d + 1] - lab0);
        const int nch = (A + 31) >> 5;
        for (long long n = p.doc_ptr[d]; n < p.doc_ptr[d + 1]; ++n) {
            const int2 mt = p.rec[n];
            const int v = mt.x, f = REC_F(mt.y), jo = REC_J(mt.y);
            const int zo = p.lab_idx[lab0 + jo];
            if (lane == 0) {
                nwk[(size_t)v * p.ldk + zo] -= f;   // LabeledLDA.py:109
                ndk[lab0 + jo] -= f;                // :110
                nk[zo] -= f;                        // :111
            }
            __syncwarp();
            const uint32_t xw = philox_word((uint64_t)(p.draw_base + n), p.sweep, GIBBS_STREAM_SWEEP, key);
            const double u = u01_f64(xw);
            // pass 1: total; pass 2: pick.  Both passes run the same serial sum, so cum values are identical.
            double run = 0.0;
            for (int c = 0; c < nch; ++c) {
                const int j = c * 32 + lane;
                double w = 0.0;
                if (j < A) {
                    const int k = p.lab_idx[lab0 + j];
                    const double a = __dadd_rn((double)ndk[lab0 + j], p.alpha);                  // :113
                    const double num = __dadd_rn((double)nwk[(size_t)v * p.ldk + k], p.beta);    // :114
                    const double den = __dadd_rn((double)nk[k], p.vbeta);                        // :115
                    w = __dmul_rn(a, __ddiv_rn(num, den));                                       // :117
                }
                const int cnt = min(32, A - c * 32);
                for (int l = 0; l < cnt; ++l) run = __dadd_rn(run, __shfl_sync(0xffffffffu, w, l));
            }
            const double thr = __dmul_rn(u, run);
            int jn = A - 1;
            bool found = false;
            double run2 = 0.0;
            for (int c = 0; c < nch && !found; ++c) {
                const int j = c * 32 + lane;
                double w = 0.0;
                if (j < A) {
                    const int k = p.lab_idx[lab0 + j];
                    const double a = __dadd_rn((double)ndk[lab0 + j], p.alpha);
                    const double num = __dadd_rn((double)nwk[(size_t)v * p.ldk + k], p.beta);
                    const double den = __dadd_rn((double)nk[k], p.vbeta);
                    w = __dmul_rn(a, __ddiv_rn(num, den));
                }
                const int cnt = min(32, A - c * 32);
                double cum = 0.0;
                for (int l = 0; l < cnt; ++l) {
                    run2 = __dadd_rn(run2, __shfl_sync(0xffffffffu, w, l));
                    if (lane == l) cum = run2;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, (j < A) && (cum > thr));
                if (bal) { jn = c * 32 + __ffs(bal) - 1; found = true; }
            }
            const int zn = p.lab_idx[lab0 + jn];
            if (lane == 0) {
                if (jn != jo) { p.rec[n].y = REC_PACK(f, jn); ++n_changed; }   // :121
                nwk[(size_t)v * p.ldk + zn] += f;    // :123
                ndk[lab0 + jn] += f;                 // :124
                nk[zn] += f;                         // :125
            }
            __syncwarp();
        }
    }
    if (lane == 0 && n_changed) atomicAdd(p.changed, n_changed);
}

// ----------------------------------------------------------------------------------------------
// Corpus preparation, histogram, merge, export.
// ----------------------------------------------------------------------------------------------

// One warp per document: pack records; z from z_init (global topic ids -> label-list index) or
// Uniform(label list) from Philox stream 1.  err[0] is set when a z_init value is not in the list.
__global__ void prepare_records_kernel(long long D, const long long *doc_ptr, const int *word, const int *freq,
                                       const int *z_init, const long long *lab_ptr, const int *lab_idx,
                                       int2 *rec, int V, uint2 key, long long draw_base, int *err) {
    const long long d = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (d >= D) return;
    const long long lab0 = lab_ptr[d];
    const int A = (int)(lab_ptr[d + 1] - lab0);
    for (long long n = doc_ptr[d] + lane; n < doc_ptr[d + 1]; n += 32) {
        const int f = freq ? freq[n] : 1;
        int j = -1;
        if (z_init) {
            const int zg = z_init[n];
            for (int q = 0; q < A; ++q)
                if (lab_idx[lab0 + q] == zg) { j = q; break; }
        } else if (A > 0) {
            const uint32_t w = philox_word((uint64_t)(draw_base + n), 0u, GIBBS_STREAM_INIT, key);
            j = (int)(((uint64_t)w * (uint64_t)A) >> 32);
        }
        int v = word[n];
        if (j < 0 || f < 0 || f > 0xffff || v < 0 || v >= V) { atomicExch(err, 1); j = 0; v = 0; }
        rec[n] = make_int2(v, REC_PACK(f, j));
    }
}

// One warp per document: histogram (LabeledLDA.py:89-92).
__global__ void counts_build_kernel(long long D, const long long *doc_ptr, const long long *lab_ptr,
                                    const int *lab_idx, const int2 *rec, int ldk,
                                    int *n_wk, int *n_dk_act, int *n_k) {
    const long long d = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (d >= D) return;
    const long long lab0 = lab_ptr[d];
    for (long long n = doc_ptr[d] + lane; n < doc_ptr[d + 1]; n += 32) {
        const int2 r = rec[n];
        const int f = REC_F(r.y), j = REC_J(r.y);
        const int k = lab_idx[lab0 + j];
        atomicAdd(&n_wk[(size_t)r.x * ldk + k], f);
        atomicAdd(&n_dk_act[lab0 + j], f);
        atomicAdd(&n_k[k], f);
    }
}

// n_wk[word][topic] += count for a COO list (duplicates accumulate).
__global__ void add_counts_kernel(long long n, const int *word, const int *topic, const int *count, int ldk, int *n_wk) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&n_wk[(size_t)word[i] * ldk + topic[i]], count[i]);
}

// Replace z: global topic ids -> label-list index inside the existing records.
__global__ void set_z_kernel(long long D, const long long *doc_ptr, const long long *lab_ptr, const int *lab_idx,
                             const int *z, int2 *rec, int *err) {
    const long long d = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (d >= D) return;
    const long long lab0 = lab_ptr[d];
    const int A = (int)(lab_ptr[d + 1] - lab0);
    for (long long n = doc_ptr[d] + lane; n < doc_ptr[d + 1]; n += 32) {
        int j = -1;
        for (int q = 0; q < A; ++q)
            if (lab_idx[lab0 + q] == z[n]) { j = q; break; }
        if (j < 0) { atomicExch(err, 1); j = 0; }
        rec[n].y = REC_PACK(REC_F(rec[n].y), j);
    }
}

__global__ void export_z_kernel(long long D, const long long *doc_ptr, const long long *lab_ptr,
                                const int *lab_idx, const int2 *rec, int *z_out) {
    const long long d = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (d >= D) return;
    const long long lab0 = lab_ptr[d];
    for (long long n = doc_ptr[d] + lane; n < doc_ptr[d + 1]; n += 32)
        z_out[n] = lab_idx[lab0 + REC_J(rec[n].y)];
}

// n_wk += delta; delta = 0; n_k += column sums of delta.  blockDim = (ldk/4 capped to 256, rows per block).
// Thread (x, y) owns int4 column chunks x, x + blockDim.x, ... so its topic columns are fixed.
__global__ void merge_delta_kernel(int4 *__restrict__ n_wk, int4 *__restrict__ delta, int *__restrict__ n_k,
                                   long long V, int ldk4, int K) {
    for (int c4 = threadIdx.x; c4 < ldk4; c4 += blockDim.x) {
        int4 acc = make_int4(0, 0, 0, 0);
        for (long long v = (long long)blockIdx.x * blockDim.y + threadIdx.y; v < V;
             v += (long long)gridDim.x * blockDim.y) {
            const size_t idx = (size_t)v * ldk4 + c4;
            const int4 dl = delta[idx];
            if (dl.x | dl.y | dl.z | dl.w) {
                int4 t = n_wk[idx];
                t.x += dl.x; t.y += dl.y; t.z += dl.z; t.w += dl.w;
                n_wk[idx] = t;
                delta[idx] = make_int4(0, 0, 0, 0);
                acc.x += dl.x; acc.y += dl.y; acc.z += dl.z; acc.w += dl.w;
            }
        }
        const int k = c4 * 4;
        if (acc.x && k < K) atomicAdd(&n_k[k], acc.x);
        if (acc.y && k + 1 < K) atomicAdd(&n_k[k + 1], acc.y);
        if (acc.z && k + 2 < K) atomicAdd(&n_k[k + 2], acc.z);
        if (acc.w && k + 3 < K) atomicAdd(&n_k[k + 3], acc.w);
    }
}

// Column sums of the word-major table: out[k] = sum_v n_wk[v][k]  (= row sums of the reference's n_k_v).
__global__ void column_sums_kernel(const int *__restrict__ n_wk, int *__restrict__ out, int V, int ldk) {
    for (int k = threadIdx.x; k < ldk; k += blockDim.x) {
        int acc = 0;
        for (int v = blockIdx.x; v < V; v += gridDim.x) acc += n_wk[(size_t)v * ldk + k];
        if (acc) atomicAdd(&out[k], acc);
    }
}

// phi[k][v] from n_wk[v][k]: 32x32 tile transpose through shared memory.
// smoothed: (n + beta) / (den[k] + V*beta) with den = n_k (LabeledLDA.py:231-234);
// otherwise n / den[k] with den = column sums of the table (CascadeLDA.py:394-395; 0/0 -> NaN as in NumPy).
__global__ void emit_phi_kernel(const int *__restrict__ n_wk, const int *__restrict__ den_k, double *__restrict__ phi,
                                int V, int K, int ldk, double beta, double vbeta, int smoothed) {
    __shared__ int tile[32][33];
    const int v0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int v = v0 + r, k = k0 + threadIdx.x;
        tile[r][threadIdx.x] = (v < V && k < ldk) ? n_wk[(size_t)v * ldk + k] : 0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int k = k0 + r, v = v0 + threadIdx.x;
        if (k < K && v < V) {
            const int c = tile[threadIdx.x][r];
            const double den = smoothed ? __dadd_rn((double)den_k[k], vbeta) : (double)den_k[k];
            const double num = smoothed ? __dadd_rn((double)c, beta) : (double)c;
            phi[(size_t)k * V + v] = __ddiv_rn(num, den);
        }
    }
}

// theta[d][:] dense.  One warp per document; the row is zero-filled, then the active entries written.
__global__ void emit_theta_kernel(long long D, const long long *lab_ptr, const int *lab_idx, const int *n_dk_act,
                                  double *theta, int K, double alpha, int smoothed) {
    const long long d = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (d >= D) return;
    const long long lab0 = lab_ptr[d];
    const int A = (int)(lab_ptr[d + 1] - lab0);
    double *row = theta + (size_t)d * K;
    for (int k = lane; k < K; k += 32) row[k] = 0.0;
    double den = 0.0;
    for (int j = 0; j < A; ++j) {
        const double x = smoothed ? __dadd_rn((double)n_dk_act[lab0 + j], alpha) : (double)n_dk_act[lab0 + j];
        den = __dadd_rn(den, x);
    }
    __syncwarp();
    for (int j = lane; j < A; j += 32) {
        const double x = smoothed ? __dadd_rn((double)n_dk_act[lab0 + j], alpha) : (double)n_dk_act[lab0 + j];
        row[lab_idx[lab0 + j]] = __ddiv_rn(x, den);
    }
}

// theta over the label lists only (the non-zero entries of LabeledLDA.py:236-239), aligned with lab_idx.
__global__ void emit_theta_csr_kernel(long long D, const long long *lab_ptr, const int *n_dk_act, double *theta_act,
                                      double alpha, int smoothed) {
    const long long d = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (d >= D) return;
    const long long lab0 = lab_ptr[d];
    const int A = (int)(lab_ptr[d + 1] - lab0);
    double den = 0.0;
    for (int j = 0; j < A; ++j) {
        const double x = smoothed ? __dadd_rn((double)n_dk_act[lab0 + j], alpha) : (double)n_dk_act[lab0 + j];
        den = __dadd_rn(den, x);
    }
    for (int j = lane; j < A; j += 32) {
        const double x = smoothed ? __dadd_rn((double)n_dk_act[lab0 + j], alpha) : (double)n_dk_act[lab0 + j];
        theta_act[lab0 + j] = __ddiv_rn(x, den);
    }
}

__global__ void philox_kat_kernel(int n, const uint32_t *ctr4, const uint32_t *key2, uint32_t *out4) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 r = philox4x32_10(make_uint4(ctr4[4 * i], ctr4[4 * i + 1], ctr4[4 * i + 2], ctr4[4 * i + 3]),
                                  make_uint2(key2[2 * i], key2[2 * i + 1]));
    out4[4 * i] = r.x; out4[4 * i + 1] = r.y; out4[4 * i + 2] = r.z; out4[4 * i + 3] = r.w;
}
