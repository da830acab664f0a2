// libgibbs_b200.so -- C-ABI implementation (include/gibbs_b200.h).  sm_100a only, no CPU path.
#include "../../include/gibbs_b200.h"
#include "llda_kernels.cuh"
#include "hslda_kernels.cuh"
#include "test_kernels.cuh"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

// ---------------------------------------------------------------------------------------------- errors
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char buf_[512];                                                                        \
            snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),    \
                     __FILE__, __LINE__);                                                          \
            return fail(e_ == cudaErrorMemoryAllocation ? GIBBS_E_NOMEM : GIBBS_E_CUDA, buf_);     \
        }                                                                                          \
    } while (0)

extern "C" const char *gibbs_last_error(void) { return g_err.c_str(); }
extern "C" const char *gibbs_version(void) { return "gibbs_b200 0.1 (sm_100a)"; }
extern "C" int gibbs_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { g_err = std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e); return 0; }
    return n;
}

// ---------------------------------------------------------------------------------------------- handle
struct DocList { long long off = 0, len = 0; };   // range inside doc_list
static const int N_BINS = 7;                       // A<=8, <=16, <=32, <=64, <=128, <=256, <=512
static const int BIN_CAP[N_BINS] = {8, 16, 32, 64, 128, 256, 512};

struct gibbs_handle {
    gibbs_desc desc{};
    int ldk = 0, sm_count = 0, max_active = 0, row_ints = 0;
    long long N = 0, n_lab = 0;
    bool loaded = false;
    uint32_t sweep = 0;
    cudaStream_t stream = nullptr;
    // corpus
    long long *doc_ptr = nullptr, *lab_ptr = nullptr;
    int *lab_idx = nullptr, *seg = nullptr, *doc_list = nullptr;
    int2 *rec = nullptr;
    // counts
    int *n_wk = nullptr, *delta_wk = nullptr, *n_k = nullptr, *n_dk_act = nullptr;
    unsigned long long *counters = nullptr;   // [0] work counter, [1] changed
    int *err_flag = nullptr;
    std::vector<DocList> lists;               // [block * N_BINS + bin]
    std::vector<long long> h_doc_ptr, h_lab_ptr;
    // hslda
    HsldaState hs{};
    // stats
    gibbs_stats_t st{};
    std::vector<cudaEvent_t> ev;
    size_t dev_bytes = 0;
};

template <typename T>
static int dalloc(gibbs_handle *h, T **p, size_t n) {
    *p = nullptr;
    if (n == 0) n = 1;
    CK(cudaMalloc((void **)p, n * sizeof(T)));
    h->dev_bytes += n * sizeof(T);
    return 0;
}
#define TRY(x) do { int r_ = (x); if (r_) return r_; } while (0)

extern "C" int gibbs_create(gibbs_t **out, const gibbs_desc *desc) {
    if (!out || !desc) return fail(GIBBS_E_ARG, "gibbs_create: null argument");
    *out = nullptr;
    if (desc->D < 0 || desc->V <= 0 || desc->K <= 0) return fail(GIBBS_E_ARG, "gibbs_create: D, V, K must be positive");
    if (desc->kind != GIBBS_KIND_LLDA && desc->kind != GIBBS_KIND_HSLDA) return fail(GIBBS_E_ARG, "gibbs_create: unknown kind");
    if (desc->mode != GIBBS_MODE_EXACT && desc->mode != GIBBS_MODE_SNAPSHOT) return fail(GIBBS_E_ARG, "gibbs_create: unknown mode");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(GIBBS_E_CUDA, std::string("gibbs_create: no CUDA device (") + cudaGetErrorString(e) +
                                      "); this library has no CPU path");
    if (desc->device < 0 || desc->device >= ndev) return fail(GIBBS_E_ARG, "gibbs_create: bad device ordinal");
    CK(cudaSetDevice(desc->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, desc->device));
    if (prop.major < 10) return fail(GIBBS_E_CUDA, "gibbs_create: device is not sm_100 class; kernels are built for sm_100a only");
    gibbs_handle *h = new gibbs_handle();
    h->desc = *desc;
    if (h->desc.n_refresh < 1) h->desc.n_refresh = 1;
    if (h->desc.mode == GIBBS_MODE_EXACT) h->desc.n_refresh = 1;
    if (h->desc.tile_docs <= 0) h->desc.tile_docs = 256;
    h->ldk = (desc->K + 31) / 32 * 32;
    h->sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    *out = h;
    return 0;
}

extern "C" void gibbs_destroy(gibbs_t *h) {
    if (!h) return;
    cudaSetDevice(h->desc.device);
    cudaStreamSynchronize(h->stream);
    void *ptrs[] = {h->doc_ptr, h->lab_ptr, h->lab_idx, h->seg, h->doc_list, h->rec, h->n_wk, h->delta_wk,
                    h->n_k, h->n_dk_act, h->counters, h->err_flag};
    for (void *p : ptrs) if (p) cudaFree(p);
    hslda_free(&h->hs);
    for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
    cudaStreamDestroy(h->stream);
    delete h;
}

static int rebuild_counts(gibbs_handle *h) {
    CK(cudaMemsetAsync(h->n_wk, 0, sizeof(int) * (size_t)h->desc.V * h->ldk, h->stream));
    if (h->delta_wk) CK(cudaMemsetAsync(h->delta_wk, 0, sizeof(int) * (size_t)h->desc.V * h->ldk, h->stream));
    CK(cudaMemsetAsync(h->n_k, 0, sizeof(int) * (size_t)h->desc.K, h->stream));
    CK(cudaMemsetAsync(h->n_dk_act, 0, sizeof(int) * (size_t)std::max<long long>(h->n_lab, 1), h->stream));
    if (h->desc.D > 0) {
        const long long blocks = (h->desc.D * 32 + 255) / 256;
        counts_build_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(h->desc.D, h->doc_ptr, h->lab_ptr, h->lab_idx,
                                                                     h->rec, h->ldk, h->n_wk, h->n_dk_act, h->n_k);
        CK(cudaGetLastError());
    }
    return 0;
}

extern "C" int gibbs_load(gibbs_t *h, const int64_t *doc_ptr, const int32_t *word, const int32_t *freq,
                          const int32_t *z_init, const int64_t *lab_ptr, const int32_t *lab_idx, const int32_t *seg) {
    if (!h || !doc_ptr || !word || !lab_ptr || !lab_idx) return fail(GIBBS_E_ARG, "gibbs_load: null argument");
    if (h->loaded) return fail(GIBBS_E_STATE, "gibbs_load: handle already loaded");
    CK(cudaSetDevice(h->desc.device));
    const long long D = h->desc.D;
    const bool hs = h->desc.kind == GIBBS_KIND_HSLDA;
    // ---- host validation
    if (doc_ptr[0] != 0 || lab_ptr[0] != 0) return fail(GIBBS_E_ARG, "gibbs_load: CSR offsets must start at 0");
    int max_a = 0;
    for (long long d = 0; d < D; ++d) {
        if (doc_ptr[d + 1] < doc_ptr[d] || lab_ptr[d + 1] < lab_ptr[d]) return fail(GIBBS_E_ARG, "gibbs_load: CSR offsets must be non-decreasing");
        const long long a = lab_ptr[d + 1] - lab_ptr[d];
        if (!hs && (a < 1 || a > BIN_CAP[N_BINS - 1])) {
            char b[160]; snprintf(b, sizeof b, "gibbs_load: document %lld has %lld active topics (supported: 1..%d)", d, a, BIN_CAP[N_BINS - 1]);
            return fail(GIBBS_E_ARG, b);
        }
        max_a = std::max<int>(max_a, (int)a);
    }
    h->N = doc_ptr[D];
    h->n_lab = lab_ptr[D];
    h->max_active = max_a;
    const int label_space = hs ? h->hs.L_hint : h->desc.K;   // HSLDA: lab_idx are label ids, checked in hslda_set
    if (!hs)
        for (long long q = 0; q < h->n_lab; ++q)
            if (lab_idx[q] < 0 || lab_idx[q] >= label_space) return fail(GIBBS_E_ARG, "gibbs_load: lab_idx out of range");
    h->row_ints = h->ldk;
    if (seg) {
        int mx = 0;
        for (long long d = 0; d < D; ++d) {
            const int lo = seg[2 * d], hi = seg[2 * d + 1];
            if (lo < 0 || hi > h->ldk || lo >= hi || (lo & 3) || (hi & 3)) return fail(GIBBS_E_ARG, "gibbs_load: bad topic segment (need 0 <= lo < hi <= ldk, multiples of 4)");
            for (long long q = lab_ptr[d]; q < lab_ptr[d + 1]; ++q)
                if (lab_idx[q] < lo || lab_idx[q] >= hi) return fail(GIBBS_E_ARG, "gibbs_load: label outside the document's topic segment");
            mx = std::max(mx, hi - lo);
        }
        h->row_ints = std::max(mx, 4);
    }
    h->h_doc_ptr.assign(doc_ptr, doc_ptr + D + 1);
    h->h_lab_ptr.assign(lab_ptr, lab_ptr + D + 1);

    // ---- upload
    const size_t tab = (size_t)h->desc.V * h->ldk;
    TRY(dalloc(h, &h->doc_ptr, (size_t)D + 1));
    TRY(dalloc(h, &h->lab_ptr, (size_t)D + 1));
    TRY(dalloc(h, &h->lab_idx, (size_t)h->n_lab));
    TRY(dalloc(h, &h->rec, (size_t)h->N));
    TRY(dalloc(h, &h->n_wk, tab));
    if (h->desc.mode == GIBBS_MODE_SNAPSHOT) TRY(dalloc(h, &h->delta_wk, tab));
    TRY(dalloc(h, &h->n_k, (size_t)h->desc.K));
    TRY(dalloc(h, &h->n_dk_act, (size_t)(hs ? D * h->desc.K : h->n_lab)));
    TRY(dalloc(h, &h->counters, 4));
    TRY(dalloc(h, &h->err_flag, 1));
    CK(cudaMemcpyAsync(h->doc_ptr, doc_ptr, sizeof(long long) * (D + 1), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->lab_ptr, lab_ptr, sizeof(long long) * (D + 1), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->lab_idx, lab_idx, sizeof(int) * h->n_lab, cudaMemcpyHostToDevice, h->stream));
    if (seg) {
        TRY(dalloc(h, &h->seg, (size_t)2 * D));
        CK(cudaMemcpyAsync(h->seg, seg, sizeof(int) * 2 * D, cudaMemcpyHostToDevice, h->stream));
    }
    int *t_word = nullptr, *t_freq = nullptr, *t_z = nullptr;
    CK(cudaMalloc((void **)&t_word, sizeof(int) * std::max<long long>(h->N, 1)));
    CK(cudaMemcpyAsync(t_word, word, sizeof(int) * h->N, cudaMemcpyHostToDevice, h->stream));
    if (freq) {
        CK(cudaMalloc((void **)&t_freq, sizeof(int) * std::max<long long>(h->N, 1)));
        CK(cudaMemcpyAsync(t_freq, freq, sizeof(int) * h->N, cudaMemcpyHostToDevice, h->stream));
    }
    if (z_init) {
        CK(cudaMalloc((void **)&t_z, sizeof(int) * std::max<long long>(h->N, 1)));
        CK(cudaMemcpyAsync(t_z, z_init, sizeof(int) * h->N, cudaMemcpyHostToDevice, h->stream));
    }
    CK(cudaMemsetAsync(h->err_flag, 0, sizeof(int), h->stream));
    const uint2 key = make_uint2((uint32_t)h->desc.seed, (uint32_t)(h->desc.seed >> 32));
    if (D > 0) {
        const long long blocks = (D * 32 + 255) / 256;
        if (hs)
            hslda_prepare_records_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(D, h->doc_ptr, t_word, t_z, h->rec,
                                                                                  h->desc.K, h->desc.V, key, h->desc.draw_base, h->err_flag);
        else
            prepare_records_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(D, h->doc_ptr, t_word, t_freq, t_z, h->lab_ptr,
                                                                            h->lab_idx, h->rec, h->desc.V, key, h->desc.draw_base, h->err_flag);
        CK(cudaGetLastError());
    }
    int err = 0;
    CK(cudaMemcpyAsync(&err, h->err_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(t_word); if (t_freq) cudaFree(t_freq); if (t_z) cudaFree(t_z);
    if (err) return fail(GIBBS_E_ARG, "gibbs_load: inconsistent corpus (word id out of range, f outside 0..65535, or z not in the document's label list)");

    // ---- work lists: [refresh block][bin by label-list length], documents in corpus order
    const int nb = h->desc.n_refresh;
    std::vector<std::vector<int>> tmp((size_t)nb * N_BINS);
    for (long long d = 0; d < D; ++d) {
        const long long tile = h->desc.tile_base + d / h->desc.tile_docs;
        const int b = (int)(tile % nb);
        int bin = 0;
        if (!hs) { const int a = (int)(lab_ptr[d + 1] - lab_ptr[d]); while (BIN_CAP[bin] < a) ++bin; }
        tmp[(size_t)b * N_BINS + bin].push_back((int)d);
    }
    std::vector<int> flat; flat.reserve((size_t)D);
    h->lists.assign((size_t)nb * N_BINS, DocList());
    for (size_t q = 0; q < tmp.size(); ++q) {
        h->lists[q].off = (long long)flat.size();
        h->lists[q].len = (long long)tmp[q].size();
        flat.insert(flat.end(), tmp[q].begin(), tmp[q].end());
    }
    TRY(dalloc(h, &h->doc_list, flat.size()));
    CK(cudaMemcpyAsync(h->doc_list, flat.data(), sizeof(int) * flat.size(), cudaMemcpyHostToDevice, h->stream));

    if (hs) {
        TRY(hslda_rebuild_counts(h->stream, D, h->doc_ptr, h->rec, h->desc.K, h->ldk, h->desc.V, h->n_wk, h->delta_wk,
                                 h->n_dk_act, h->n_k));
    } else {
        TRY(rebuild_counts(h));
    }
    CK(cudaStreamSynchronize(h->stream));
    h->loaded = true;
    h->st.ldk = h->ldk;
    h->st.max_active = max_a;
    // dense-row byte model: record read 8 + z write 4 + row segment + 2 RED (8 B each way) + per-doc terms
    {
        const double nd = D > 0 ? (double)h->N / (double)D : 1.0;
        const double abar = D > 0 ? (double)h->n_lab / (double)D : 1.0;
        h->st.bytes_per_draw = 16.0 + 4.0 * h->row_ints + 16.0 + (8.0 * h->row_ints + 8.0 + 4.0 * abar) / std::max(nd, 1.0);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------- sweeps
template <int G, int NCH, int R>
static int launch_snapshot(gibbs_handle *h, const SweepParams &p) {
    auto kern = llda_snapshot_kernel<G, NCH, R>;
    const size_t grp_bytes = (size_t)(2 * R) * 8 + (size_t)R * p.row_ints * 4;
    const int gpw = 32 / G;
    const size_t warp_bytes = grp_bytes * gpw;
    int wpc = (int)std::min<size_t>(8, std::max<size_t>(1, (72 * 1024) / warp_bytes));
    const size_t smem = warp_bytes * wpc;
    if (smem > 227 * 1024) return -100;   // caller retries with a shallower ring
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, wpc * 32, smem));
    if (occ < 1) return -100;
    const long long groups_needed = p.n_list;
    const long long ctas_needed = (groups_needed + (long long)wpc * gpw - 1) / ((long long)wpc * gpw);
    const unsigned grid = (unsigned)std::min<long long>(ctas_needed, (long long)occ * h->sm_count);
    kern<<<grid, wpc * 32, smem, h->stream>>>(p);
    CK(cudaGetLastError());
    h->st.last_launches++;
    return 0;
}

template <int G, int NCH>
static int launch_snapshot_r(gibbs_handle *h, const SweepParams &p) {
    int r = launch_snapshot<G, NCH, 4>(h, p);
    if (r == -100) r = launch_snapshot<G, NCH, 2>(h, p);
    if (r == -100) return fail(GIBBS_E_ARG, "gibbs_sweep: n_wk row segment too long for the shared-memory ring");
    return r;
}

static int sample_block(gibbs_handle *h, int block) {
    if (h->desc.kind == GIBBS_KIND_HSLDA) {
        const DocList &dl = h->lists[(size_t)block * N_BINS];
        if (!dl.len) return 0;
        CK(cudaMemsetAsync(h->counters, 0, sizeof(unsigned long long), h->stream));
        int r = hslda_launch(h->stream, h->sm_count, &h->hs, h->doc_ptr, h->lab_ptr, h->lab_idx, h->rec, h->n_wk, h->delta_wk,
                             h->n_k, h->n_dk_act, h->doc_list + dl.off, dl.len, h->counters, h->counters + 1, h->ldk,
                             h->desc.K, (float)h->desc.beta, (float)((double)h->desc.V * h->desc.beta), h->desc.seed, h->sweep,
                             h->desc.draw_base);
        if (r) return fail(GIBBS_E_CUDA, std::string("hslda launch failed: ") + cudaGetErrorString(cudaGetLastError()));
        h->st.last_launches++;
        return 0;
    }
    if (h->desc.mode == GIBBS_MODE_EXACT) {
        ExactParams p{};
        p.doc_ptr = h->doc_ptr; p.lab_ptr = h->lab_ptr; p.lab_idx = h->lab_idx; p.n_dk_act = h->n_dk_act;
        p.rec = h->rec; p.n_wk = h->n_wk; p.n_k = h->n_k; p.d_begin = 0; p.d_end = h->desc.D; p.ldk = h->ldk;
        p.alpha = h->desc.alpha; p.beta = h->desc.beta; p.vbeta = (double)h->desc.V * h->desc.beta;
        p.seed_lo = (uint32_t)h->desc.seed; p.seed_hi = (uint32_t)(h->desc.seed >> 32); p.sweep = h->sweep;
        p.draw_base = h->desc.draw_base; p.changed = h->counters + 1;
        llda_exact_kernel<<<1, 32, 0, h->stream>>>(p);
        CK(cudaGetLastError());
        h->st.last_launches++;
        return 0;
    }
    for (int bin = 0; bin < N_BINS; ++bin) {
        const DocList &dl = h->lists[(size_t)block * N_BINS + bin];
        if (!dl.len) continue;
        SweepParams p{};
        p.doc_ptr = h->doc_ptr; p.lab_ptr = h->lab_ptr; p.lab_idx = h->lab_idx; p.n_dk_act = h->n_dk_act;
        p.rec = h->rec; p.n_wk = h->n_wk; p.delta_wk = h->delta_wk; p.n_k = h->n_k; p.seg = h->seg;
        p.doc_list = h->doc_list + dl.off; p.n_list = dl.len; p.counter = h->counters; p.changed = h->counters + 1;
        p.ldk = h->ldk; p.row_ints = h->row_ints;
        p.alpha = (float)h->desc.alpha; p.beta = (float)h->desc.beta; p.vbeta = (float)((double)h->desc.V * h->desc.beta);
        p.seed_lo = (uint32_t)h->desc.seed; p.seed_hi = (uint32_t)(h->desc.seed >> 32); p.sweep = h->sweep;
        p.draw_base = h->desc.draw_base;
        CK(cudaMemsetAsync(h->counters, 0, sizeof(unsigned long long), h->stream));
        switch (bin) {
            case 0: TRY((launch_snapshot_r<8, 1>(h, p))); break;
            case 1: TRY((launch_snapshot_r<16, 1>(h, p))); break;
            case 2: TRY((launch_snapshot_r<32, 1>(h, p))); break;
            case 3: TRY((launch_snapshot_r<32, 2>(h, p))); break;
            case 4: TRY((launch_snapshot_r<32, 4>(h, p))); break;
            case 5: TRY((launch_snapshot_r<32, 8>(h, p))); break;
            default: TRY((launch_snapshot_r<32, 16>(h, p))); break;
        }
    }
    return 0;
}

static int merge_block(gibbs_handle *h) {
    if (h->desc.mode != GIBBS_MODE_SNAPSHOT) return 0;
    const int ldk4 = h->ldk / 4;
    int bx = std::min(256, (ldk4 + 31) / 32 * 32);
    int by = std::max(1, 256 / bx);
    dim3 block(bx, by);
    const long long rows_per_pass = by;
    unsigned grid = (unsigned)std::min<long long>((h->desc.V + rows_per_pass - 1) / rows_per_pass, (long long)h->sm_count * 8);
    merge_delta_kernel<<<grid, block, 0, h->stream>>>(reinterpret_cast<int4 *>(h->n_wk), reinterpret_cast<int4 *>(h->delta_wk),
                                                      h->n_k, h->desc.V, ldk4, h->desc.K);
    CK(cudaGetLastError());
    h->st.last_launches++;
    return 0;
}

static cudaEvent_t get_event(gibbs_handle *h, size_t i) {
    while (h->ev.size() <= i) { cudaEvent_t e; cudaEventCreate(&e); h->ev.push_back(e); }
    return h->ev[i];
}

extern "C" int gibbs_sweep_begin(gibbs_t *h, int32_t block) {
    if (!h || !h->loaded) return fail(GIBBS_E_STATE, "gibbs_sweep_begin: corpus not loaded");
    if (block < 0 || block >= h->desc.n_refresh) return fail(GIBBS_E_ARG, "gibbs_sweep_begin: bad block");
    CK(cudaSetDevice(h->desc.device));
    if (block == 0) {
        h->st.last_launches = 0;
        CK(cudaMemsetAsync(h->counters + 1, 0, sizeof(unsigned long long), h->stream));
    }
    CK(cudaEventRecord(get_event(h, 0), h->stream));
    TRY(sample_block(h, block));
    CK(cudaEventRecord(get_event(h, 1), h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
    if (block == 0) h->st.last_sweep_ms = 0;
    h->st.last_sweep_ms += ms;
    return 0;
}

extern "C" int gibbs_sweep_end(gibbs_t *h, int32_t block) {
    if (!h || !h->loaded) return fail(GIBBS_E_STATE, "gibbs_sweep_end: corpus not loaded");
    if (block < 0 || block >= h->desc.n_refresh) return fail(GIBBS_E_ARG, "gibbs_sweep_end: bad block");
    CK(cudaSetDevice(h->desc.device));
    CK(cudaEventRecord(get_event(h, 0), h->stream));
    TRY(merge_block(h));
    CK(cudaEventRecord(get_event(h, 1), h->stream));
    if (block == h->desc.n_refresh - 1) {
        CK(cudaMemcpyAsync(&h->st.changed, h->counters + 1, sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
    if (block == 0) h->st.last_merge_ms = 0;
    h->st.last_merge_ms += ms;
    if (block == h->desc.n_refresh - 1) { h->sweep++; h->st.sweeps++; h->st.draws += h->N; }
    return 0;
}

extern "C" int gibbs_sweep(gibbs_t *h, int32_t n_sweeps) {
    if (!h || !h->loaded) return fail(GIBBS_E_STATE, "gibbs_sweep: corpus not loaded");
    if (n_sweeps < 0) return fail(GIBBS_E_ARG, "gibbs_sweep: n_sweeps < 0");
    if (n_sweeps == 0) return 0;
    CK(cudaSetDevice(h->desc.device));
    const int nb = h->desc.n_refresh;
    h->st.last_launches = 0;
    size_t e = 0;
    for (int s = 0; s < n_sweeps; ++s) {
        CK(cudaMemsetAsync(h->counters + 1, 0, sizeof(unsigned long long), h->stream));
        for (int b = 0; b < nb; ++b) {
            CK(cudaEventRecord(get_event(h, e++), h->stream));
            TRY(sample_block(h, b));
            CK(cudaEventRecord(get_event(h, e++), h->stream));
            TRY(merge_block(h));
            CK(cudaEventRecord(get_event(h, e++), h->stream));
        }
        h->sweep++;
    }
    CK(cudaMemcpyAsync(&h->st.changed, h->counters + 1, sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    double samp = 0, mrg = 0;
    for (size_t q = 0; q + 2 < e + 0 && q < e; q += 3) {
        float a = 0, b = 0;
        CK(cudaEventElapsedTime(&a, h->ev[q], h->ev[q + 1]));
        CK(cudaEventElapsedTime(&b, h->ev[q + 1], h->ev[q + 2]));
        samp += a; mrg += b;
    }
    h->st.last_sweep_ms = samp / n_sweeps;
    h->st.last_merge_ms = mrg / n_sweeps;
    h->st.last_launches /= n_sweeps;
    h->st.sweeps += n_sweeps;
    h->st.draws += (long long)n_sweeps * h->N;
    return 0;
}

extern "C" int gibbs_delta_buffer(gibbs_t *h, void **dev_ptr, int64_t *n_elems) {
    if (!h || !h->loaded || !dev_ptr || !n_elems) return fail(GIBBS_E_STATE, "gibbs_delta_buffer: corpus not loaded");
    if (!h->delta_wk) return fail(GIBBS_E_STATE, "gibbs_delta_buffer: exact mode has no delta table");
    *dev_ptr = h->delta_wk;
    *n_elems = (int64_t)h->desc.V * h->ldk;
    return 0;
}

extern "C" int gibbs_stream(gibbs_t *h, void **stream) {
    if (!h || !stream) return fail(GIBBS_E_ARG, "gibbs_stream: null argument");
    *stream = (void *)h->stream;
    return 0;
}

extern "C" int gibbs_set_sweep_counter(gibbs_t *h, uint32_t sweep) {
    if (!h) return fail(GIBBS_E_ARG, "gibbs_set_sweep_counter: null handle");
    h->sweep = sweep;
    return 0;
}

// ---------------------------------------------------------------------------------------------- state out / in
extern "C" int gibbs_get_state(gibbs_t *h, int32_t *z, int32_t *n_wk, int32_t *n_dk_act, int32_t *n_k) {
    if (!h || !h->loaded) return fail(GIBBS_E_STATE, "gibbs_get_state: corpus not loaded");
    CK(cudaSetDevice(h->desc.device));
    const long long D = h->desc.D;
    int *t_z = nullptr;
    if (z && h->N > 0) {
        CK(cudaMalloc((void **)&t_z, sizeof(int) * h->N));
        const long long blocks = (D * 32 + 255) / 256;
        if (h->desc.kind == GIBBS_KIND_HSLDA)
            hslda_export_z_kernel<<<(unsigned)((h->N + 255) / 256), 256, 0, h->stream>>>(h->N, h->rec, t_z);
        else
            export_z_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(D, h->doc_ptr, h->lab_ptr, h->lab_idx, h->rec, t_z);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(z, t_z, sizeof(int) * h->N, cudaMemcpyDeviceToHost, h->stream));
    }
    if (n_wk)
        CK(cudaMemcpy2DAsync(n_wk, sizeof(int) * h->desc.K, h->n_wk, sizeof(int) * h->ldk, sizeof(int) * h->desc.K,
                             (size_t)h->desc.V, cudaMemcpyDeviceToHost, h->stream));
    if (n_dk_act) {
        const long long cnt = h->desc.kind == GIBBS_KIND_HSLDA ? D * h->desc.K : h->n_lab;
        CK(cudaMemcpyAsync(n_dk_act, h->n_dk_act, sizeof(int) * cnt, cudaMemcpyDeviceToHost, h->stream));
    }
    if (n_k) CK(cudaMemcpyAsync(n_k, h->n_k, sizeof(int) * h->desc.K, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (t_z) cudaFree(t_z);
    return 0;
}

extern "C" int gibbs_set_z(gibbs_t *h, const int32_t *z) {
    if (!h || !h->loaded || !z) return fail(GIBBS_E_STATE, "gibbs_set_z: corpus not loaded");
    CK(cudaSetDevice(h->desc.device));
    const long long D = h->desc.D;
    int *t_z = nullptr;
    CK(cudaMalloc((void **)&t_z, sizeof(int) * std::max<long long>(h->N, 1)));
    CK(cudaMemcpyAsync(t_z, z, sizeof(int) * h->N, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->err_flag, 0, sizeof(int), h->stream));
    if (D > 0) {
        const long long blocks = (D * 32 + 255) / 256;
        if (h->desc.kind == GIBBS_KIND_HSLDA)
            hslda_set_z_kernel<<<(unsigned)((h->N + 255) / 256), 256, 0, h->stream>>>(h->N, t_z, h->rec, h->desc.K, h->err_flag);
        else
            set_z_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(D, h->doc_ptr, h->lab_ptr, h->lab_idx, t_z, h->rec, h->err_flag);
        CK(cudaGetLastError());
    }
    int err = 0;
    CK(cudaMemcpyAsync(&err, h->err_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(t_z);
    if (err) return fail(GIBBS_E_ARG, "gibbs_set_z: z not in the document's label list");
    if (h->desc.kind == GIBBS_KIND_HSLDA)
        TRY(hslda_rebuild_counts(h->stream, D, h->doc_ptr, h->rec, h->desc.K, h->ldk, h->desc.V, h->n_wk, h->delta_wk, h->n_dk_act, h->n_k));
    else
        TRY(rebuild_counts(h));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------- outputs
extern "C" int gibbs_emit_phi(gibbs_t *h, double *phi_KV, int32_t smoothed) {
    if (!h || !h->loaded || !phi_KV) return fail(GIBBS_E_STATE, "gibbs_emit_phi: corpus not loaded");
    CK(cudaSetDevice(h->desc.device));
    const int K = h->desc.K, V = h->desc.V;
    double *d_phi = nullptr;
    CK(cudaMalloc((void **)&d_phi, sizeof(double) * (size_t)K * V));
    dim3 grid((V + 31) / 32, (K + 31) / 32), block(32, 8);
    emit_phi_kernel<<<grid, block, 0, h->stream>>>(h->n_wk, h->n_k, d_phi, V, K, h->ldk, h->desc.beta,
                                                   (double)V * h->desc.beta, smoothed);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(phi_KV, d_phi, sizeof(double) * (size_t)K * V, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(d_phi);
    return 0;
}

extern "C" int gibbs_emit_theta(gibbs_t *h, double *theta_DK, int32_t smoothed) {
    if (!h || !h->loaded || !theta_DK) return fail(GIBBS_E_STATE, "gibbs_emit_theta: corpus not loaded");
    CK(cudaSetDevice(h->desc.device));
    const long long D = h->desc.D;
    const int K = h->desc.K;
    if (D == 0) return 0;
    double *d_th = nullptr;
    CK(cudaMalloc((void **)&d_th, sizeof(double) * (size_t)D * K));
    const long long blocks = (D * 32 + 255) / 256;
    if (h->desc.kind == GIBBS_KIND_HSLDA)
        hslda_emit_zbar_kernel<<<(unsigned)((D * K + 255) / 256), 256, 0, h->stream>>>(D, K, h->doc_ptr, h->n_dk_act, d_th);
    else
        emit_theta_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(D, h->lab_ptr, h->lab_idx, h->n_dk_act, d_th, K, h->desc.alpha, smoothed);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(theta_DK, d_th, sizeof(double) * (size_t)D * K, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(d_th);
    return 0;
}

extern "C" int gibbs_stats(gibbs_t *h, gibbs_stats_t *out) {
    if (!h || !out) return fail(GIBBS_E_ARG, "gibbs_stats: null argument");
    h->st.device_bytes = (int64_t)h->dev_bytes;
    *out = h->st;
    return 0;
}

// ---------------------------------------------------------------------------------------------- HSLDA state
extern "C" int gibbs_hslda_set(gibbs_t *h, int32_t L, const double *eta, const double *a_act, const double *mean_a_act,
                               const double *alpha_beta) {
    if (!h || !h->loaded) return fail(GIBBS_E_STATE, "gibbs_hslda_set: corpus not loaded");
    if (h->desc.kind != GIBBS_KIND_HSLDA) return fail(GIBBS_E_STATE, "gibbs_hslda_set: handle is not HSLDA");
    if (L <= 0 || !eta || !a_act || !mean_a_act || !alpha_beta) return fail(GIBBS_E_ARG, "gibbs_hslda_set: null argument");
    CK(cudaSetDevice(h->desc.device));
    int r = hslda_set(h->stream, &h->hs, L, h->desc.K, h->n_lab, eta, a_act, mean_a_act, alpha_beta, h->max_active);
    if (r) return fail(GIBBS_E_CUDA, std::string("gibbs_hslda_set: ") + cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// ---------------------------------------------------------------------------------------------- test chains
extern "C" int gibbs_test_chains(int32_t device, int32_t K, int32_t V, double alpha, const double *phi_KV,
                                 int64_t D_test, const int64_t *doc_ptr, const int32_t *word, const int32_t *freq,
                                 const int32_t *z_init, int32_t it, int32_t thinning, uint64_t seed, double *th_hat) {
    if (!phi_KV || !doc_ptr || !word || !z_init || !th_hat) return fail(GIBBS_E_ARG, "gibbs_test_chains: null argument");
    if (K <= 0 || V <= 0 || D_test < 0 || it < 0 || thinning <= 0) return fail(GIBBS_E_ARG, "gibbs_test_chains: bad sizes");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(GIBBS_E_CUDA, "gibbs_test_chains: no CUDA device; this library has no CPU path");
    CK(cudaSetDevice(device));
    int r = test_chains_run(K, V, alpha, phi_KV, D_test, doc_ptr, word, freq, z_init, it, thinning, seed, th_hat);
    if (r == -1) return fail(GIBBS_E_ARG, "gibbs_test_chains: K too large for the test kernel");
    if (r) return fail(GIBBS_E_CUDA, std::string("gibbs_test_chains: ") + cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// ---------------------------------------------------------------------------------------------- KAT
extern "C" int gibbs_philox_kat(int32_t device, int32_t n, const uint32_t *ctr4, const uint32_t *key2, uint32_t *out4) {
    if (n <= 0 || !ctr4 || !key2 || !out4) return fail(GIBBS_E_ARG, "gibbs_philox_kat: bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(GIBBS_E_CUDA, "gibbs_philox_kat: no CUDA device");
    CK(cudaSetDevice(device));
    uint32_t *d_c = nullptr, *d_k = nullptr, *d_o = nullptr;
    CK(cudaMalloc((void **)&d_c, sizeof(uint32_t) * 4 * n));
    CK(cudaMalloc((void **)&d_k, sizeof(uint32_t) * 2 * n));
    CK(cudaMalloc((void **)&d_o, sizeof(uint32_t) * 4 * n));
    CK(cudaMemcpy(d_c, ctr4, sizeof(uint32_t) * 4 * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_k, key2, sizeof(uint32_t) * 2 * n, cudaMemcpyHostToDevice));
    philox_kat_kernel<<<(n + 127) / 128, 128>>>(n, d_c, d_k, d_o);
    CK(cudaGetLastError());
    CK(cudaMemcpy(out4, d_o, sizeof(uint32_t) * 4 * n, cudaMemcpyDeviceToHost));
    cudaFree(d_c); cudaFree(d_k); cudaFree(d_o);
    return 0;
}
