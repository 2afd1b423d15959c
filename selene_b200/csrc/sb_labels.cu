// selene_b200 — kernel (b): interval-overlap target labels over sorted BED intervals.
//
// Reference: selene_sdk/targets/_genomic_features.pyx:12-41 (_fast_get_feature_data) fed by the
// tabix query of selene_sdk/targets/genomic_features.py:313-348.  The reference paints a
// (bin_len x n_features) int64 bitmap per query and sums its columns; what it computes is, per
// feature, the length of the UNION of the clipped intervals, compared with an integer threshold.
//
// Layout in HBM: all intervals sorted by (chromosome, start) as four int32 arrays — start, end,
// feature, and a running maximum of `end` inside each chromosome — plus per-chromosome offsets
// (16 B per interval).  One WARP owns one query bin:
//   1. lanes binary-search the running max for the first interval that can still reach the bin
//      (exact for any interval-length mix; ~log2(n) L2-resident probes, done once per QUERY, not per
//      feature),
//   2. the warp sweeps the overlapping intervals 32 at a time (coalesced loads), replaying them in
//      start order so the per-feature union coverage (cov[f], covered_until[f] in shared memory) is
//      exact,
//   3. lanes compare coverage with the host-computed integer thresholds and write the F labels with
//      coalesced stores.
// Algorithmic traffic per query: F x sizeof(out) written + 12 B per overlapping interval read; for
// large batches the kernel is bound by those output writes.
#include "sb_common.cuh"

#include <algorithm>
#include <numeric>
#include <vector>

struct sb_intervals {
    int device;
    int n_features;
    int n_chrom;
    int64_t n;
    int32_t* start;      // dev, sorted by (chrom, start)
    int32_t* end;        // dev
    int32_t* feat;       // dev
    int32_t* cmax;       // dev, running max of end inside each chromosome
    int64_t* chrom_off;  // dev, n_chrom + 1
};

namespace {

constexpr int kWarpsPerBlock = 4;

template <typename OutT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) label_bins_kernel(
    const int32_t* __restrict__ iv_start, const int32_t* __restrict__ iv_end, const int32_t* __restrict__ iv_feat,
    const int32_t* __restrict__ iv_cmax, const int64_t* __restrict__ chrom_off, int n_features, int n_chrom,
    const int32_t* __restrict__ q_chrom, const int32_t* __restrict__ q_start, const int32_t* __restrict__ q_end,
    int n_queries, const int32_t* __restrict__ thr_i, OutT* __restrict__ out) {
    extern __shared__ int32_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int32_t* cov = smem + (size_t)warp * 2 * n_features;  // union coverage per feature
    int32_t* until = cov + n_features;                     // bin coordinate covered so far per feature
    const int warps_total = gridDim.x * kWarpsPerBlock;
    for (int q = blockIdx.x * kWarpsPerBlock + warp; q < n_queries; q += warps_total) {
        for (int f = lane; f < n_features; f += 32) { cov[f] = 0; until[f] = 0; }
        __syncwarp();
        const int c = q_chrom[q];
        const int qs = q_start[q];
        const int qe = q_end[q];
        const int qlen = qe - qs;
        if (c >= 0 && c < n_chrom && qlen > 0) {
            const long long lo0 = chrom_off[c], hi0 = chrom_off[c + 1];
            // first interval whose running max end exceeds qs (the running max is non-decreasing)
            long long lo = lo0, hi = hi0;
            while (lo < hi) {
                const long long mid = (lo + hi) >> 1;
                if (__ldg(iv_cmax + mid) > qs) hi = mid; else lo = mid + 1;
            }
            for (long long base = lo; base < hi0; base += 32) {
                const long long i = base + lane;
                int fs = 0x7fffffff, fe = 0, ff = 0;
                if (i < hi0) { fs = __ldg(iv_start + i); fe = __ldg(iv_end + i); ff = __ldg(iv_feat + i); }
                const unsigned live = __ballot_sync(0xffffffffu, fs < qe);   // sorted by start: a prefix
                const int cnt = __popc(live);
                for (int l = 0; l < cnt; ++l) {
                    const int s = __shfl_sync(0xffffffffu, fs, l);
                    const int e = __shfl_sync(0xffffffffu, fe, l);
                    const int f = __shfl_sync(0xffffffffu, ff, l);
                    if (lane == 0 && e > qs) {                 // tabix: start < q_end and end > q_start
                        int is = s - qs; if (is < 0) is = 0;              // _genomic_features.pyx:32
                        int ie = e - qs; if (ie > qlen) ie = qlen;        // :33
                        if (is == ie) ie += 1;                            // :35-36
                        if (ie > qlen) ie = qlen;                         // numpy slice assignment clamps
                        const int u = until[f];
                        if (ie > u) { cov[f] += ie - (is > u ? is : u); until[f] = ie; }
                        cov[f] |= 0x40000000;                             // "any row" marker for no-threshold mode
                    }
                }
                if (cnt < 32) break;
            }
        }
        __syncwarp();
        OutT* o = out + (long long)q * n_features;
        for (int f = lane; f < n_features; f += 32) {
            const int v = cov[f];
            const int t = __ldg(thr_i + f);
            // :39-40 with thresholds; genomic_features.py:406-415 (any overlapping row) when thr < 0
            const int label = t < 0 ? (v != 0) : ((v & 0x3fffffff) > t);
            o[f] = (OutT)label;
        }
        __syncwarp();
    }
}

}  // namespace

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
        if (chrom[a] != chrom[b]) return chrom[a] < chrom[b];
        return start[a] < start[b];
    });
    std::vector<int32_t> s((size_t)n), e((size_t)n), f((size_t)n), m((size_t)n);
    std::vector<int64_t> off((size_t)n_chrom + 1, 0);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t k = order[(size_t)i];
        s[(size_t)i] = start[k];
        e[(size_t)i] = end[k];
        f[(size_t)i] = feat[k];
        off[(size_t)chrom[k] + 1] += 1;
    }
    for (int c = 0; c < n_chrom; ++c) off[(size_t)c + 1] += off[(size_t)c];
    for (int c = 0; c < n_chrom; ++c) {
        int32_t run = INT32_MIN;
        for (int64_t i = off[(size_t)c]; i < off[(size_t)c + 1]; ++i) {
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
    SB_CHECK_CUDA(cudaMalloc(&iv->feat, nb));
    SB_CHECK_CUDA(cudaMalloc(&iv->cmax, nb));
    SB_CHECK_CUDA(cudaMalloc(&iv->chrom_off, sizeof(int64_t) * (size_t)(n_chrom + 1)));
    if (n > 0) {
        SB_CHECK_CUDA(cudaMemcpy(iv->start, s.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice));
        SB_CHECK_CUDA(cudaMemcpy(iv->end, e.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice));
        SB_CHECK_CUDA(cudaMemcpy(iv->feat, f.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice));
        SB_CHECK_CUDA(cudaMemcpy(iv->cmax, m.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice));
    }
    SB_CHECK_CUDA(cudaMemcpy(iv->chrom_off, off.data(), sizeof(int64_t) * (size_t)(n_chrom + 1),
                             cudaMemcpyHostToDevice));
    *out = iv;
    return SB_OK;
}

extern "C" int sb_intervals_destroy(sb_intervals* iv) {
    if (!iv) return SB_OK;
    cudaSetDevice(iv->device);
    cudaFree(iv->start);
    cudaFree(iv->end);
    cudaFree(iv->feat);
    cudaFree(iv->cmax);
    cudaFree(iv->chrom_off);
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
    const size_t smem = (size_t)kWarpsPerBlock * 2 * iv->n_features * sizeof(int32_t);
    SB_REQUIRE(smem <= 200 * 1024, "sb_label_bins: %d features need %zu bytes of shared memory", iv->n_features, smem);
    int64_t blocks = ((int64_t)n + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const int64_t cap = (int64_t)sb_sm_count(iv->device) * 6;
    if (blocks > cap) blocks = cap;
    cudaStream_t s = (cudaStream_t)stream;
#define SB_LAUNCH_LABEL(T)                                                                                  \
    do {                                                                                                    \
        if (smem > 48 * 1024)                                                                               \
            SB_CHECK_CUDA(cudaFuncSetAttribute(label_bins_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                               (int)smem));                                                 \
        label_bins_kernel<T><<<(int)blocks, kWarpsPerBlock * 32, smem, s>>>(                               \
            iv->start, iv->end, iv->feat, iv->cmax, iv->chrom_off, iv->n_features, iv->n_chrom, chrom, bin_start,  \
            bin_end, n, thr_i, (T*)out);                                                                    \
    } while (0)
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
