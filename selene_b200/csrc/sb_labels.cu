// selene_b200 — kernel (b): interval-overlap target labels over sorted BED intervals.
//
// Reference: selene_sdk/targets/_genomic_features.pyx:12-41 (_fast_get_feature_data) fed by the
// tabix query of selene_sdk/targets/genomic_features.py:313-348.  The reference paints a
// (bin_len x n_features) int64 bitmap per query and sums its columns; the quantity it computes is,
// per feature, the length of the UNION of the clipped intervals, compared with an integer
// threshold.  Here intervals live in HBM sorted by (feature, chromosome, start) with a running
// maximum of `end` per segment, so that one thread per (query, feature):
//   1. binary-searches the running max for the first interval that can still reach the bin
//      (exact for any interval length mix — no "max interval length" heuristic),
//   2. sweeps the few overlapping intervals in start order accumulating union coverage,
//   3. compares with the host-computed integer threshold and stores the label.
// Algorithmic traffic per query: n_features x sizeof(out) written + 12 B per overlapping interval
// read (+ ~log2(segment) x 4 B of L2-resident index probes).
#include "sb_common.cuh"

#include <algorithm>
#include <numeric>
#include <vector>

struct sb_intervals {
    int device;
    int n_features;
    int n_chrom;
    int64_t n;
    int32_t* start;    // dev, sorted by (feature, chrom, start)
    int32_t* end;      // dev
    int32_t* cmax;     // dev, running max of end inside each (feature, chrom) segment
    int64_t* seg_off;  // dev, n_features * n_chrom + 1
};

template <typename OutT>
__global__ void __launch_bounds__(256) label_bins_kernel(
    const int32_t* __restrict__ iv_start, const int32_t* __restrict__ iv_end,
    const int32_t* __restrict__ iv_cmax, const int64_t* __restrict__ seg_off, int n_features,
    int n_chrom, const int32_t* __restrict__ q_chrom, const int32_t* __restrict__ q_start,
    const int32_t* __restrict__ q_end, int n_queries, const int32_t* __restrict__ thr_i,
    OutT* __restrict__ out) {
    const int64_t total = (int64_t)n_queries * n_features;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int q = (int)(idx / n_features);
        const int f = (int)(idx - (int64_t)q * n_features);
        const int c = q_chrom[q];
        const int qs = q_start[q];
        const int qe = q_end[q];
        const int qlen = qe - qs;
        int label = 0;
        if (c >= 0 && c < n_chrom && qlen > 0) {
            const int64_t lo0 = seg_off[(int64_t)f * n_chrom + c];
            const int64_t hi0 = seg_off[(int64_t)f * n_chrom + c + 1];
            // first interval whose running max end exceeds qs (running max is non-decreasing)
            int64_t lo = lo0, hi = hi0;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (iv_cmax[mid] > qs) hi = mid; else lo = mid + 1;
            }
            int cov = 0;
            int cur_end = 0;  // in bin coordinates
            bool any = false;
            for (int64_t i = lo; i < hi0; ++i) {
                const int fs = iv_start[i];
                if (fs >= qe) break;  // sorted by start: nothing later overlaps
                const int fe = iv_end[i];
                if (fe <= qs) continue;  // tabix semantics: start < q_end and end > q_start
                any = true;
                int is = fs - qs; if (is < 0) is = 0;               // _genomic_features.pyx:32
                int ie = fe - qs; if (ie > qlen) ie = qlen;          // :33
                if (is == ie) ie += 1;                               // :35-36
                if (ie > qlen) ie = qlen;                            // numpy slice clamps
                if (ie > cur_end) {
                    cov += ie - (is > cur_end ? is : cur_end);
                    cur_end = ie;
                }
            }
            const int t = thr_i[f];
            label = t < 0 ? (any ? 1 : 0) : (cov > t ? 1 : 0);       // :40 / genomic_features.py:406-415
        }
        out[idx] = (OutT)label;
    }
}

extern "C" int sb_intervals_create(const int32_t* chrom, const int32_t* start, const int32_t* end,
                                   const int32_t* feat, int64_t n, int n_features, int n_chrom,
                                   int device, sb_intervals** out) {
    SB_REQUIRE(out && n_features > 0 && n_chrom > 0 && n >= 0, "sb_intervals_create: bad sizes");
    SB_REQUIRE(n == 0 || (chrom && start && end && feat), "sb_intervals_create: null argument");
    for (int64_t i = 0; i < n; ++i) {
        SB_REQUIRE(feat[i] >= 0 && feat[i] < n_features,
                   "sb_intervals_create: interval %lld has feature index %d outside [0,%d)",
                   (long long)i, feat[i], n_features);
        SB_REQUIRE(chrom[i] >= 0 && chrom[i] < n_chrom,
                   "sb_intervals_create: interval %lld has chromosome index %d outside [0,%d)",
                   (long long)i, chrom[i], n_chrom);
    }
    std::vector<int64_t> order((size_t)n);
    std::iota(order.begin(), order.end(), (int64_t)0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
        if (feat[a] != feat[b]) return feat[a] < feat[b];
        if (chrom[a] != chrom[b]) return chrom[a] < chrom[b];
        return start[a] < start[b];
    });
    const int64_t n_seg = (int64_t)n_features * n_chrom;
    std::vector<int32_t> s((size_t)n), e((size_t)n), m((size_t)n);
    std::vector<int64_t> seg((size_t)n_seg + 1, 0);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t k = order[(size_t)i];
        s[(size_t)i] = start[k];
        e[(size_t)i] = end[k];
        seg[(size_t)((int64_t)feat[k] * n_chrom + chrom[k]) + 1] += 1;
    }
    for (int64_t k = 0; k < n_seg; ++k) seg[(size_t)k + 1] += seg[(size_t)k];
    for (int64_t k = 0; k < n_seg; ++k) {
        int32_t run = INT32_MIN;
        for (int64_t i = seg[(size_t)k]; i < seg[(size_t)k + 1]; ++i) {
            run = std::max(run, e[(size_t)i]);
            m[(size_t)i] = run;
        }
    }
    SB_CHECK_CUDA(cudaSetDevice(device));
    sb_intervals* iv = new sb_intervals();
    memset(iv, 0, sizeof(*iv));
    iv->device = device;
    iv->n_features = n_features;
    iv->n_chrom = n_chrom;
    iv->n = n;
    const size_t nb = sizeof(int32_t) * (size_t)std::max<int64_t>(n, 1);
    SB_CHECK_CUDA(cudaMalloc(&iv->start, nb));
    SB_CHECK_CUDA(cudaMalloc(&iv->end, nb));
    SB_CHECK_CUDA(cudaMalloc(&iv->cmax, nb));
    SB_CHECK_CUDA(cudaMalloc(&iv->seg_off, sizeof(int64_t) * (size_t)(n_seg + 1)));
    if (n > 0) {
        SB_CHECK_CUDA(cudaMemcpy(iv->start, s.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice));
        SB_CHECK_CUDA(cudaMemcpy(iv->end, e.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice));
        SB_CHECK_CUDA(cudaMemcpy(iv->cmax, m.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice));
    }
    SB_CHECK_CUDA(cudaMemcpy(iv->seg_off, seg.data(), sizeof(int64_t) * (size_t)(n_seg + 1),
                             cudaMemcpyHostToDevice));
    *out = iv;
    return SB_OK;
}

extern "C" int sb_intervals_destroy(sb_intervals* iv) {
    if (!iv) return SB_OK;
    cudaSetDevice(iv->device);
    cudaFree(iv->start);
    cudaFree(iv->end);
    cudaFree(iv->cmax);
    cudaFree(iv->seg_off);
    delete iv;
    return SB_OK;
}

extern "C" int sb_label_bins(const sb_intervals* iv, const int32_t* chrom, const int32_t* bin_start,
                             const int32_t* bin_end, int n, const int32_t* thr_i, int out_dtype,
                             void* out, void* stream) {
    SB_REQUIRE(iv, "sb_label_bins: null handle");
    SB_REQUIRE(n >= 0, "sb_label_bins: bad n");
    if (n == 0) return SB_OK;
    SB_REQUIRE(chrom && bin_start && bin_end && thr_i && out, "sb_label_bins: null argument");
    const int64_t total = (int64_t)n * iv->n_features;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)sb_sm_count(iv->device) * 8;
    if (blocks > cap) blocks = cap;
    cudaStream_t s = (cudaStream_t)stream;
#define SB_LAUNCH_LABEL(T)                                                                       \
    label_bins_kernel<T><<<(int)blocks, 256, 0, s>>>(iv->start, iv->end, iv->cmax, iv->seg_off,   \
                                                     iv->n_features, iv->n_chrom, chrom,          \
                                                     bin_start, bin_end, n, thr_i, (T*)out)
    switch (out_dtype) {
        case SB_U8: SB_LAUNCH_LABEL(uint8_t); break;
        case SB_F32: SB_LAUNCH_LABEL(float); break;
        case SB_I64: SB_LAUNCH_LABEL(long long); break;
        case SB_F64: SB_LAUNCH_LABEL(double); break;
        default:
            sb_set_error("sb_label_bins: bad out_dtype %d", out_dtype);
            return SB_ERR_ARG;
    }
#undef SB_LAUNCH_LABEL
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    return SB_OK;
}
