// selene_b200 — kernel (b): interval-overlap target labels over sorted BED intervals.
//
// Reference: selene_sdk/targets/_genomic_features.pyx:12-41 (_fast_get_feature_data) fed by the
// tabix query of selene_sdk/targets/genomic_features.py:313-348.  The reference paints a
// (bin_len x n_features) int64 bitmap per query and sums its columns; what it computes is, per
// feature, the length of the UNION of the clipped intervals, compared with an integer threshold.
//
// Layout in HBM: all intervals sorted by (chromosome, start) as four int32 arrays — start, end,
// feature, and a running maximum of `end` inside each chromosome — plus per-chromosome offsets
// (16 B per interval).  One WARP owns one query bin, 48 warps per SM; what bounds small outputs (uint8 rows) is the
// chain of DEPENDENT memory round trips per query, so the kernel is built to have two:
//   1. one read of a bucket table over positions (one int32 per 1024 bases, L2-resident) gives the first interval
//      whose running max end exceeds the start of the bin's bucket: everything before it ends at or before the bin
//      start (exact for any interval-length mix; once per QUERY, not per feature);
//   2. the warp loads the 32 intervals from there (three coalesced loads in flight together).  Intervals are sorted
//      by start, so if one of the 32 starts at or after the bin end the candidates are complete — the common case.
//      A feature that occurs once among the overlapping ones is labelled straight from registers; a feature with
//      several intervals has its union built by the lowest lane of its group from warp shuffles, in start order;
//   3. the row is zero-filled with 16-byte streaming stores (issued before the loads, off the chain) and only the
//      features of overlapping intervals are compared with the host-computed integer thresholds and set (labels
//      are sparse; thresholds are >= 0, so an untouched feature is always 0).
// Bins with more than 32 candidates or 2^15 bases and more take a general path over a per-warp union state
// (coverage and covered_until per feature) in a scratch buffer owned by the handle.
// Algorithmic traffic per query: F x sizeof(out) written + 12 B per overlapping interval read + the table entry;
// for int64 / float32 rows the kernel is bound by those output writes.
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
    // bucket table: for bucket b of chromosome c (positions [b << tab_shift, ...)), the first interval (relative to
    // chrom_off[c]) whose running max end exceeds b << tab_shift; the answer for any bin start inside the bucket lies
    // between two consecutive entries, so a query needs one table read and (almost always) one round of probes
    int32_t* tab;        // dev, tab_off[n_chrom] entries
    int64_t* tab_off;    // dev, n_chrom + 1
    int tab_shift;
    // general-path union state, one slice per resident warp, all-zero between queries (allocated by the first
    // sb_label_bins; launches that share a handle must be ordered on one stream)
    int32_t* scratch;
    size_t scratch_words;
};

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kBlocksPerSm = 6;

// Per-feature state of one query, one 32-bit word per feature while the bin is shorter than 2^15 bases
// (bits 0-14 covered_until, bits 15-29 union coverage, bit 30 "some row overlapped"); longer bins take the
// features in two passes with two words each in the same shared memory.
constexpr int kPackedMaxLen = 32767;
constexpr int32_t kAnyRow = 0x40000000;

// zero one output row with 16-byte stores (rows are only sizeof(OutT)-aligned: element stores for the ragged ends)
template <typename OutT>
__device__ __forceinline__ void zero_row(OutT* __restrict__ o, int n, int lane) {
    constexpr int kPer = 16 / (int)sizeof(OutT);
    const int mis = (int)((reinterpret_cast<uintptr_t>(o) & 15) / sizeof(OutT));
    int head = mis ? kPer - mis : 0;
    if (head > n) head = n;
    if (lane < head) o[lane] = (OutT)0;
    const int body = (n - head) / kPer;
    uint4* o4 = reinterpret_cast<uint4*>(o + head);
    for (int i = lane; i < body; i += 32) __stcs(o4 + i, make_uint4(0, 0, 0, 0));   // streaming: written once, never re-read here
    const int done = head + body * kPer;
    if (lane < n - done) o[done + lane] = (OutT)0;
}

struct LabelArgs {
    const int32_t *iv_start, *iv_end, *iv_feat, *iv_cmax;
    const int64_t* chrom_off;
    const int32_t* tab;
    const int64_t* tab_off;
    int tab_shift, n_features, n_chrom;
    int meta_in_smem;   // chrom_off / tab_off copied to shared memory (n_chrom small)
    int32_t* scratch;   // general-path state: (resident warps) x words
};

// clipped [is, ie) of an interval inside the bin, as the reference paints it
__device__ __forceinline__ void clip_interval(int s, int e, int qs, int qlen, int& is, int& ie) {
    is = s - qs; if (is < 0) is = 0;              // _genomic_features.pyx:32
    ie = e - qs; if (ie > qlen) ie = qlen;        // :33
    if (is == ie) ie += 1;                        // :35-36
    if (ie > qlen) ie = qlen;                     // numpy slice assignment clamps
}

// lane 0 replays the intervals of `hit` (lane order = start order) into the per-feature union state
__device__ __forceinline__ void replay(unsigned hit, int fs, int fe, int ff, int qs, int qlen, bool packed, int f_lo,
                                       int half, int32_t* st, int lane) {
    while (hit) {
        const int l = __ffs(hit) - 1;
        hit &= hit - 1;
        const int s = __shfl_sync(0xffffffffu, fs, l);
        const int e = __shfl_sync(0xffffffffu, fe, l);
        const int f = __shfl_sync(0xffffffffu, ff, l);
        if (lane == 0) {
            int is, ie;
            clip_interval(s, e, qs, qlen, is, ie);
            if (packed) {
                const int32_t w = st[f];
                int u = w & 0x7fff, cv = (w >> 15) & 0x7fff;
                if (ie > u) { cv += ie - (is > u ? is : u); u = ie; }
                st[f] = u | (cv << 15) | kAnyRow;             // "any row" marker for no-threshold mode
            } else {
                int32_t* cov = st + (f - f_lo);
                int32_t* until = cov + half;
                const int u = *until;
                if (ie > u) { *cov += ie - (is > u ? is : u); *until = ie; }
                *cov |= kAnyRow;
            }
        }
    }
}

// Labels are sparse (a bin overlaps a handful of the features), so a query costs one vectorised zero fill of its
// row plus work proportional to the intervals it overlaps: no per-feature loop anywhere.
template <typename OutT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, kBlocksPerSm) label_bins_kernel(
    const LabelArgs a, const int32_t* __restrict__ q_chrom, const int32_t* __restrict__ q_start,
    const int32_t* __restrict__ q_end, int n_queries, const int32_t* __restrict__ thr_i, OutT* __restrict__ out) {
    extern __shared__ int64_t meta[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_features = a.n_features;
    const int words = (n_features + 1) & ~1;             // per warp
    int32_t* st = a.scratch + ((size_t)blockIdx.x * kWarpsPerBlock + warp) * words;
    const int64_t* coff = a.chrom_off;
    const int64_t* toff = a.tab_off;
    if (a.meta_in_smem) {
        for (int i = threadIdx.x; i <= a.n_chrom; i += blockDim.x) {
            meta[i] = a.chrom_off[i];
            meta[a.n_chrom + 1 + i] = a.tab_off[i];
        }
        coff = meta;
        toff = meta + a.n_chrom + 1;
        __syncthreads();
    }
    const int warps_total = gridDim.x * kWarpsPerBlock;
    int q = blockIdx.x * kWarpsPerBlock + warp;
    int c_n = 0, qs_n = 0, qe_n = 0;
    if (q < n_queries) { c_n = q_chrom[q]; qs_n = q_start[q]; qe_n = q_end[q]; }
    for (; q < n_queries; q += warps_total) {
        const int c = c_n, qs = qs_n, qe = qe_n;
        if (q + warps_total < n_queries) {   // next query's coordinates: off the dependent chain
            c_n = q_chrom[q + warps_total]; qs_n = q_start[q + warps_total]; qe_n = q_end[q + warps_total];
        }
        const int qlen = qe - qs;
        OutT* o = out + (long long)q * n_features;
        zero_row(o, n_features, lane);
        __syncwarp();   // orders the zero fill before the label stores of other lanes
        if (!(c >= 0 && c < a.n_chrom && qlen > 0)) continue;
        const long long lo0 = coff[c], hi0 = coff[c + 1];
        long long lo;
        {
            const long long t0 = toff[c];
            const long long nb = toff[c + 1] - t0;
            const long long b = qs < 0 ? 0 : (long long)(qs >> a.tab_shift);
            if (b >= nb - 1) continue;                   // past the last interval end of the chromosome: no overlap
            lo = qs < 0 ? lo0 : lo0 + __ldg(a.tab + t0 + b);
        }
        const bool packed = qlen <= kPackedMaxLen;
        // ---- the 32 candidates from the table position
        int fs = 0x7fffffff, fe = 0, ff = 0;
        if (lo + lane < hi0) {
            fs = __ldg(a.iv_start + lo + lane); fe = __ldg(a.iv_end + lo + lane); ff = __ldg(a.iv_feat + lo + lane);
        }
        const bool more = __ballot_sync(0xffffffffu, fs < qe) == 0xffffffffu;   // sorted by start: a prefix
        const bool is_hit = fs < qe && fe > qs;                                 // tabix: start < q_end and end > q_start
        const unsigned hitmask = __ballot_sync(0xffffffffu, is_hit);
        if (!more && packed) {
            if (hitmask == 0) continue;
            unsigned group = 0;
            if (is_hit) group = __match_any_sync(hitmask, ff);
            const bool unique = is_hit && group == (1u << lane);
            const bool leader = is_hit && (int)(__ffs(group) - 1) == lane;
            int cov = 0;
            if (unique) {
                int is, ie;
                clip_interval(fs, fe, qs, qlen, is, ie);
                cov = ie - is;
            }
            // features with several intervals: the group's lowest lane merges them in lane (= start) order
            unsigned rem = __ballot_sync(0xffffffffu, is_hit && !unique);
            int until = 0;
            while (rem) {
                const int l = __ffs(rem) - 1;
                rem &= rem - 1;
                const int s = __shfl_sync(0xffffffffu, fs, l);
                const int e = __shfl_sync(0xffffffffu, fe, l);
                if (leader && ((group >> l) & 1)) {
                    int is, ie;
                    clip_interval(s, e, qs, qlen, is, ie);
                    if (ie > until) { cov += ie - (is > until ? is : until); until = ie; }
                }
            }
            if (leader) {
                const int t = __ldg(thr_i + ff);
                // :39-40 with thresholds; genomic_features.py:406-415 (any overlapping row) when thr < 0
                if (t < 0 || cov > t) o[ff] = (OutT)1;
            }
            continue;
        }
        // ---- general path: more than 32 candidates and/or a bin of 2^15 bases or more (features in two passes)
        const int half = words >> 1;
        const int n_pass = packed ? 1 : 2;
        for (int pass = 0; pass < n_pass; ++pass) {
            const int f_lo = packed ? 0 : pass * half;
            const int f_hi = packed ? n_features : min(n_features, f_lo + half);
            for (int sweep = 0; sweep < 3; ++sweep) {
                for (long long base = lo; base < hi0; base += 32) {
                    const long long i = base + lane;
                    int gs = 0x7fffffff, ge = 0, gf = 0;
                    if (i < hi0) { gs = __ldg(a.iv_start + i); ge = __ldg(a.iv_end + i); gf = __ldg(a.iv_feat + i); }
                    const unsigned live = __ballot_sync(0xffffffffu, gs < qe);
                    const bool h = gs < qe && ge > qs && gf >= f_lo && gf < f_hi;
                    if (sweep == 0) {          // union state, in start order
                        replay(__ballot_sync(0xffffffffu, h), gs, ge, gf, qs, qlen, packed, f_lo, half, st, lane);
                    } else if (sweep == 1) {   // every lane labels its interval's feature (duplicates agree)
                        if (h) {
                            const int32_t w = st[gf - f_lo];
                            const int cv = packed ? ((w >> 15) & 0x7fff) : (w & 0x3fffffff);
                            const int t = __ldg(thr_i + gf);
                            if (t < 0 || cv > t) o[gf] = (OutT)1;
                        }
                    } else if (h) {            // clear the touched state
                        st[gf - f_lo] = 0;
                        if (!packed) st[half + gf - f_lo] = 0;
                    }
                    if (live != 0xffffffffu) break;
                }
                __syncwarp();
                __threadfence_block();
            }
        }
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
    // bucket table over positions (see sb_intervals): shift grows until the table is at most 16 M entries
    int64_t span = 1;
    for (int c = 0; c < n_chrom; ++c)
        if (off[(size_t)c + 1] > off[(size_t)c]) span += std::max<int64_t>(m[(size_t)off[(size_t)c + 1] - 1], 0);
    int shift = 10;
    while ((span >> shift) + 2 * n_chrom > (int64_t)(16 << 20)) ++shift;
    std::vector<int64_t> toff((size_t)n_chrom + 1, 0);
    std::vector<int32_t> tab;
    for (int c = 0; c < n_chrom; ++c) {
        const int64_t b0 = off[(size_t)c], b1 = off[(size_t)c + 1];
        SB_REQUIRE(b1 - b0 < INT32_MAX, "sb_intervals_create: chromosome %d has too many intervals", c);
        const int64_t max_end = b1 > b0 ? std::max<int64_t>(m[(size_t)b1 - 1], 0) : 0;
        const int64_t nb = (max_end >> shift) + 2;
        int64_t i = b0;
        for (int64_t b = 0; b < nb; ++b) {
            const int64_t p = b << shift;
            while (i < b1 && (int64_t)m[(size_t)i] <= p) ++i;   // running max is non-decreasing
            tab.push_back((int32_t)(i - b0));
        }
        toff[(size_t)c + 1] = (int64_t)tab.size();
    }
    SB_CHECK_CUDA(cudaSetDevice(device));
    sb_intervals* iv = new sb_intervals();
    memset(iv, 0, sizeof(*iv));
    iv->tab_shift = shift;
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
    SB_CHECK_CUDA(cudaMalloc(&iv->tab, sizeof(int32_t) * tab.size()));
    SB_CHECK_CUDA(cudaMalloc(&iv->tab_off, sizeof(int64_t) * (size_t)(n_chrom + 1)));
    SB_CHECK_CUDA(cudaMemcpy(iv->tab, tab.data(), sizeof(int32_t) * tab.size(), cudaMemcpyHostToDevice));
    SB_CHECK_CUDA(cudaMemcpy(iv->tab_off, toff.data(), sizeof(int64_t) * (size_t)(n_chrom + 1), cudaMemcpyHostToDevice));
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
    cudaFree(iv->tab);
    cudaFree(iv->tab_off);
    if (iv->scratch) cudaFree(iv->scratch);
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
    const int meta_in_smem = iv->n_chrom <= 512;
    const size_t smem = meta_in_smem ? 2 * (size_t)(iv->n_chrom + 1) * sizeof(int64_t) : 0;
    int64_t blocks = ((int64_t)n + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const int64_t cap = (int64_t)sb_sm_count(iv->device) * kBlocksPerSm;   // persistent grid: 48 warps per SM
    if (blocks > cap) blocks = cap;
    // general-path state: one slice of `words` per warp of the largest grid, zeroed once (the kernel leaves it zero)
    const size_t words = (size_t)((iv->n_features + 1) & ~1);
    const size_t need = (size_t)cap * kWarpsPerBlock * words;
    if (iv->scratch == nullptr || iv->scratch_words < need) {
        sb_intervals* mut = const_cast<sb_intervals*>(iv);
        if (mut->scratch) SB_CHECK_CUDA(cudaFree(mut->scratch));
        mut->scratch = nullptr;
        SB_CHECK_CUDA(cudaMalloc(&mut->scratch, need * sizeof(int32_t)));
        SB_CHECK_CUDA(cudaMemset(mut->scratch, 0, need * sizeof(int32_t)));
        mut->scratch_words = need;
    }
    LabelArgs a;
    a.iv_start = iv->start; a.iv_end = iv->end; a.iv_feat = iv->feat; a.iv_cmax = iv->cmax;
    a.chrom_off = iv->chrom_off; a.tab = iv->tab; a.tab_off = iv->tab_off; a.tab_shift = iv->tab_shift;
    a.n_features = iv->n_features; a.n_chrom = iv->n_chrom; a.meta_in_smem = meta_in_smem; a.scratch = iv->scratch;
    cudaStream_t s = (cudaStream_t)stream;
#define SB_LAUNCH_LABEL(T)                                                                                  \
    do {                                                                                                    \
        label_bins_kernel<T><<<(int)blocks, kWarpsPerBlock * 32, smem, s>>>(a, chrom, bin_start, bin_end, n, thr_i, \
                                                                            (T*)out);                       \
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
