// selene_b200 — kernel (a): genome image -> codes / one-hot windows, ISM expansion, SNV ref/alt pairs.
//
// All of these are HBM-bound byte movers: a 1000-bp fp32 window reads 250 B of 2-bit codes +
// 2 x 125 B of masks and writes 16 000 B, so the design goal is simply "every store is a full
// 128-bit coalesced store and nothing is read twice from DRAM".  One CTA owns one output row
// (window / mutant / variant): the packed bytes it needs are served by L1 after the first touch,
// thread t writes bases t, t+blockDim, ... so a warp writes 512 contiguous bytes per instruction.
//
// Reference semantics restated here (each cited at the function that implements it):
//   selene_sdk/sequences/_sequence.pyx:11-25, selene_sdk/sequences/genome.py:96-166,542,
//   selene_sdk/predict/_in_silico_mutagenesis.py:91-143,
//   selene_sdk/predict/_variant_effect_prediction.py:137-245, selene_sdk/predict/_common.py:37-62.
#include "sb_common.cuh"

struct sb_genome {
    int device;
    int n_chrom;
    int64_t n_image;
    uint8_t* packed;   // dev, 2 bits / base
    uint8_t* unk;      // dev, 1 bit / base
    uint8_t* upn;      // dev, 1 bit / base
    int64_t* chrom_off;  // dev
    int64_t* chrom_len;  // dev
};

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
struct GenomeView {
    const uint8_t* __restrict__ packed;
    const uint8_t* __restrict__ unk;
    const uint8_t* __restrict__ upn;
    const int64_t* __restrict__ chrom_off;
    const int64_t* __restrict__ chrom_len;
};

// code (0..3, 4 = unknown) of image base `gi`; sets *upper_n when the character was 'N'.
__device__ __forceinline__ int image_code(const GenomeView& g, int64_t gi, bool* upper_n) {
    const int byte = __ldg(g.packed + (gi >> 2));
    const int code = (byte >> (2 * (int)(gi & 3))) & 3;
    const int m = __ldg(g.unk + (gi >> 3));
    const int un = (m >> (int)(gi & 7)) & 1;
    if (un) {
        const int u = __ldg(g.upn + (gi >> 3));
        *upper_n = ((u >> (int)(gi & 7)) & 1) != 0;
        return 4;
    }
    *upper_n = false;
    return code;
}

// Code of output position i of window [start, start+L) on chromosome of length clen at image
// offset coff.  Out-of-range parts are 'N' (genome.py:156-166); with rc the in-bounds part alone is
// reverse-complemented and the padding stays where it was (the reference concatenates
// N*start_pad + revcomp(inner) + N*end_pad, genome.py:164-166).
__device__ __forceinline__ int window_code(const GenomeView& g, int64_t coff, int64_t clen,
                                           int64_t start, int L, bool rc, int i, bool* upper_n) {
    const int64_t end = start + L;
    const int64_t s0 = start < 0 ? 0 : start;
    const int64_t e0 = end > clen ? clen : end;
    const int64_t start_pad = s0 - start;
    const int64_t inner = e0 - s0;
    const int64_t j = (int64_t)i - start_pad;
    if (j < 0 || j >= inner) {
        *upper_n = true;  // Genome.UNK_BASE == "N"
        return 4;
    }
    if (!rc) return image_code(g, coff + s0 + j, upper_n);
    const int c = image_code(g, coff + e0 - 1 - j, upper_n);
    return c < 4 ? 3 - c : 4;  // A<->T, C<->G; unknown stays unknown (pyfaidx complement keeps N/n)
}

struct Perm {
    uint8_t p[4];
};

// one-hot row of a code as 4 floats in channel order (unknown: 0.25 x 4, _sequence.pyx:17,21-24)
__device__ __forceinline__ float4 onehot_f32(int code, const Perm& perm) {
    if (code > 3) return make_float4(0.25f, 0.25f, 0.25f, 0.25f);
    const int ch = perm.p[code];
    return make_float4(ch == 0 ? 1.f : 0.f, ch == 1 ? 1.f : 0.f, ch == 2 ? 1.f : 0.f,
                       ch == 3 ? 1.f : 0.f);
}
// the same as 4 packed bf16 (0x3F80 = 1.0, 0x3E80 = 0.25)
__device__ __forceinline__ uint2 onehot_bf16(int code, const Perm& perm) {
    if (code > 3) return make_uint2(0x3E803E80u, 0x3E803E80u);
    const int ch = perm.p[code];
    const uint32_t lo = ch == 0 ? 0x00003F80u : (ch == 1 ? 0x3F800000u : 0u);
    const uint32_t hi = ch == 2 ? 0x00003F80u : (ch == 3 ? 0x3F800000u : 0u);
    return make_uint2(lo, hi);
}
template <bool BF16>
__device__ __forceinline__ void store_onehot(void* out, int64_t base_index, int code,
                                             const Perm& perm) {
    if (BF16) {
        reinterpret_cast<uint2*>(out)[base_index] = onehot_bf16(code, perm);
    } else {
        reinterpret_cast<float4*>(out)[base_index] = onehot_f32(code, perm);
    }
}
// complement permutation applied by get_reverse_complement_encoding (_common.py:57-62): the value
// in channel perm[b] moves to channel perm[3-b]; with codes that is simply code -> 3 - code.

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
// One WARP per window, grid-stride over windows.  MODE 0: codes only; 1: one-hot fp32; 2: one-hot bf16.
// Lane l of iteration k owns BPL consecutive bases (1 for fp32, 2 for bf16, 16 for codes) so that every
// store is one 16-byte store and a warp writes 512 contiguous bytes per instruction; the per-window
// flags are a warp reduction (no block barrier, no shared memory).
template <int MODE>
__global__ void __launch_bounds__(256) window_kernel(GenomeView g, const int32_t* __restrict__ chrom,
                                                     const int64_t* __restrict__ start,
                                                     const uint8_t* __restrict__ strand_rc, int n, int L,
                                                     Perm perm, void* __restrict__ out,
                                                     uint8_t* __restrict__ flags_out) {
    constexpr int BPL = MODE == 1 ? 1 : (MODE == 2 ? 2 : 16);
    const int lane = threadIdx.x & 31;
    const int warps_total = gridDim.x * (blockDim.x >> 5);
    for (int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n; w += warps_total) {
        const int c = chrom[w];
        const int64_t coff = g.chrom_off[c];
        const int64_t clen = g.chrom_len[c];
        const int64_t st = start[w];
        const bool rc = strand_rc != nullptr && strand_rc[w] != 0;
        int flags = 0;
        for (int i0 = lane * BPL; i0 < L; i0 += 32 * BPL) {
            int codes[BPL];
#pragma unroll
            for (int b = 0; b < BPL; ++b) {
                bool upn = false;
                int code = 4;
                if (i0 + b < L) {
                    code = window_code(g, coff, clen, st, L, rc, i0 + b, &upn);
                    if (code > 3) flags |= SB_FLAG_HAS_UNK;
                    if (upn) flags |= SB_FLAG_HAS_UPPER_N;
                }
                codes[b] = code;
            }
            const int64_t o = (int64_t)w * L + i0;
            if (MODE == 1) {
                reinterpret_cast<float4*>(out)[o] = onehot_f32(codes[0], perm);
            } else if (MODE == 2) {
                // two bases = 16 bytes when both exist and the address is 16-byte aligned
                uint2* p = reinterpret_cast<uint2*>(out) + o;
                if (i0 + 1 < L && ((o & 1) == 0)) {
                    const uint2 a = onehot_bf16(codes[0], perm), b2 = onehot_bf16(codes[1], perm);
                    *reinterpret_cast<uint4*>(p) = make_uint4(a.x, a.y, b2.x, b2.y);
                } else {
                    p[0] = onehot_bf16(codes[0], perm);
                    if (i0 + 1 < L) p[1] = onehot_bf16(codes[1], perm);
                }
            } else {
                uint8_t* p = reinterpret_cast<uint8_t*>(out) + o;
                if (i0 + 16 <= L && ((o & 15) == 0)) {
                    uint32_t wd[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        wd[k] = (uint32_t)codes[4 * k] | ((uint32_t)codes[4 * k + 1] << 8) |
                                ((uint32_t)codes[4 * k + 2] << 16) | ((uint32_t)codes[4 * k + 3] << 24);
                    *reinterpret_cast<uint4*>(p) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
                } else {
#pragma unroll
                    for (int b = 0; b < BPL; ++b)
                        if (i0 + b < L) p[b] = (uint8_t)codes[b];
                }
            }
        }
        if (flags_out != nullptr) {
            flags = __reduce_or_sync(0xffffffffu, flags);
            if (lane == 0) flags_out[w] = (uint8_t)flags;
        }
    }
}

// One-hot of code rows, flat over n*L bases.
template <bool BF16>
__global__ void __launch_bounds__(256) codes_onehot_kernel(const uint8_t* __restrict__ codes,
                                                           int64_t total, Perm perm,
                                                           void* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
        store_onehot<BF16>(out, i, __ldg(codes + i), perm);
}

// ISM enumeration (single CTA): ascending positions, alternatives in channel order minus the
// reference base, unknown reference -> all four (_in_silico_mutagenesis.py:91-107).
__global__ void __launch_bounds__(1024) ism_enumerate_kernel(const uint8_t* __restrict__ codes,
                                                             int L, int pos0, int pos1, Perm perm,
                                                             int32_t* __restrict__ mut_pos,
                                                             uint8_t* __restrict__ mut_alt,
                                                             int32_t* __restrict__ n_mut, int cap) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    // inverse permutation: channel -> canonical code
    int inv[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) inv[perm.p[b]] = b;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = pos0; base < pos1; base += blockDim.x) {
        const int pos = base + threadIdx.x;
        int code = 4, cnt = 0;
        if (pos < pos1 && pos < L) {
            code = codes[pos];
            cnt = code > 3 ? 4 : 3;
        }
        // block-wide exclusive scan of cnt
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int ws = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
            int wincl = ws;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, wincl, d);
                if (lane >= d) wincl += v;
            }
            warp_sums[lane] = wincl - ws;  // exclusive
        }
        __syncthreads();
        const int block_total = warp_sums[31] + 0;  // exclusive offset of last warp (see below)
        int off = carry + warp_sums[wid] + incl - cnt;
        for (int ch = 0; ch < 4 && cnt > 0; ++ch) {
            const int b = inv[ch];
            if (b == code) continue;
            if (off < cap) {
                mut_pos[off] = pos;
                mut_alt[off] = (uint8_t)b;
            }
            ++off;
        }
        (void)block_total;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = carry + warp_sums[wid] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_mut = carry;
}

// One CTA per mutant row (mutate_sequence, _in_silico_mutagenesis.py:138-143: copy, zero the row,
// set the alt channel — so an unknown reference row becomes a clean one-hot).
template <bool BF16>
__global__ void __launch_bounds__(256) ism_expand_kernel(const uint8_t* __restrict__ codes, int L,
                                                         const int32_t* __restrict__ src,
                                                         const int32_t* __restrict__ mut_pos,
                                                         const uint8_t* __restrict__ mut_alt,
                                                         Perm perm, void* __restrict__ out) {
    const int64_t m = blockIdx.x;
    const uint8_t* row = codes + (int64_t)(src ? src[m] : 0) * L;
    const int pos = mut_pos[m];
    const int alt = mut_alt[m];
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        const int code = (i == pos) ? alt : __ldg(row + i);
        store_onehot<BF16>(out, m * L + i, code, perm);
    }
}

// One CTA per variant: '+' window codes, forced ref row / alt row, optional reverse complement of
// the finished encodings (which mirrors the variant row to L-1-var_row and complements the codes).
template <int MODE>  // 0: no one-hot outputs, 1: fp32, 2: bf16
__global__ void __launch_bounds__(256) snv_pairs_kernel(
    GenomeView g, const int32_t* __restrict__ chrom, const int64_t* __restrict__ start,
    const uint8_t* __restrict__ ref_code, const uint8_t* __restrict__ alt_code,
    const uint8_t* __restrict__ strand_rc, int L, int var_row, Perm perm, void* __restrict__ ref_out,
    void* __restrict__ alt_out, uint8_t* __restrict__ codes_out, uint8_t* __restrict__ ref_match,
    uint8_t* __restrict__ flags_out) {
    const int v = blockIdx.x;
    const int c = chrom[v];
    const int64_t coff = g.chrom_off[c];
    const int64_t clen = g.chrom_len[c];
    const int64_t st = start[v];
    const bool rc = strand_rc != nullptr && strand_rc[v] != 0;
    const int rcode = ref_code[v];
    const int acode = alt_code[v];
    int flags = 0;
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        bool upn;
        const int code = window_code(g, coff, clen, st, L, false, i, &upn);
        if (code > 3) flags |= SB_FLAG_HAS_UNK;
        if (upn) flags |= SB_FLAG_HAS_UPPER_N;
        if (codes_out) {  // strand-applied, unsubstituted window codes for sb_net_forward_*
            const int oc = rc ? L - 1 - i : i;
            codes_out[(int64_t)v * L + oc] = (uint8_t)(rc ? (code < 4 ? 3 - code : 4) : code);
        }
        if (i == var_row && ref_match) ref_match[v] = (uint8_t)(code == rcode);
        if (MODE != 0) {
            int rc_ref = (i == var_row) ? rcode : code;
            int rc_alt = (i == var_row) ? acode : code;
            int o = i;
            if (rc) {
                o = L - 1 - i;
                rc_ref = rc_ref < 4 ? 3 - rc_ref : 4;
                rc_alt = rc_alt < 4 ? 3 - rc_alt : 4;
            }
            if (ref_out) store_onehot<MODE == 2>(ref_out, (int64_t)v * L + o, rc_ref, perm);
            if (alt_out) store_onehot<MODE == 2>(alt_out, (int64_t)v * L + o, rc_alt, perm);
        }
    }
    if (flags_out != nullptr) {
        const int any_unk = __syncthreads_or(flags & SB_FLAG_HAS_UNK);
        const int any_upn = __syncthreads_or(flags & SB_FLAG_HAS_UPPER_N);
        if (threadIdx.x == 0)
            flags_out[v] = (uint8_t)((any_unk ? SB_FLAG_HAS_UNK : 0) |
                                     (any_upn ? SB_FLAG_HAS_UPPER_N : 0));
    }
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
static Perm make_perm(const uint8_t perm[4]) {
    Perm p;
    for (int i = 0; i < 4; ++i) p.p[i] = perm[i];
    return p;
}
static bool perm_ok(const uint8_t perm[4]) {
    if (!perm) return false;
    int seen = 0;
    for (int i = 0; i < 4; ++i) {
        if (perm[i] > 3) return false;
        seen |= 1 << perm[i];
    }
    return seen == 15;
}
static int window_grid(int n, int device) {
    const long long blocks = ((long long)n + 7) / 8;           // 8 warps (windows) per CTA
    const long long cap = (long long)sb_sm_count(device) * 8;  // persistent: 8 CTAs of 256 threads per SM
    return (int)(blocks < cap ? blocks : cap);
}

static GenomeView view_of(const sb_genome* g) {
    GenomeView v;
    v.packed = g->packed;
    v.unk = g->unk;
    v.upn = g->upn;
    v.chrom_off = g->chrom_off;
    v.chrom_len = g->chrom_len;
    return v;
}

extern "C" int sb_genome_create(const uint8_t* packed2bit, const uint8_t* unk_mask,
                                const uint8_t* upn_mask, int64_t n_image_bases,
                                const int64_t* chrom_off, const int64_t* chrom_len, int n_chrom,
                                int device, sb_genome** out) {
    SB_REQUIRE(packed2bit && unk_mask && upn_mask && chrom_off && chrom_len && out,
               "sb_genome_create: null argument");
    SB_REQUIRE(n_chrom > 0 && n_image_bases > 0, "sb_genome_create: empty genome");
    for (int c = 0; c < n_chrom; ++c)
        SB_REQUIRE(chrom_off[c] >= 0 && chrom_len[c] >= 0 &&
                       chrom_off[c] + chrom_len[c] <= n_image_bases,
                   "sb_genome_create: chromosome %d outside the image", c);
    SB_CHECK_CUDA(cudaSetDevice(device));
    sb_genome* g = new sb_genome();
    memset(g, 0, sizeof(*g));
    g->device = device;
    g->n_chrom = n_chrom;
    g->n_image = n_image_bases;
    const size_t nb2 = (size_t)((n_image_bases + 3) / 4);
    const size_t nb1 = (size_t)((n_image_bases + 7) / 8);
    SB_CHECK_CUDA(cudaMalloc(&g->packed, nb2 + 16));
    SB_CHECK_CUDA(cudaMalloc(&g->unk, nb1 + 16));
    SB_CHECK_CUDA(cudaMalloc(&g->upn, nb1 + 16));
    SB_CHECK_CUDA(cudaMalloc(&g->chrom_off, sizeof(int64_t) * n_chrom));
    SB_CHECK_CUDA(cudaMalloc(&g->chrom_len, sizeof(int64_t) * n_chrom));
    SB_CHECK_CUDA(cudaMemcpy(g->packed, packed2bit, nb2, cudaMemcpyHostToDevice));
    SB_CHECK_CUDA(cudaMemcpy(g->unk, unk_mask, nb1, cudaMemcpyHostToDevice));
    SB_CHECK_CUDA(cudaMemcpy(g->upn, upn_mask, nb1, cudaMemcpyHostToDevice));
    SB_CHECK_CUDA(
        cudaMemcpy(g->chrom_off, chrom_off, sizeof(int64_t) * n_chrom, cudaMemcpyHostToDevice));
    SB_CHECK_CUDA(
        cudaMemcpy(g->chrom_len, chrom_len, sizeof(int64_t) * n_chrom, cudaMemcpyHostToDevice));
    *out = g;
    return SB_OK;
}

extern "C" int sb_genome_destroy(sb_genome* g) {
    if (!g) return SB_OK;
    cudaSetDevice(g->device);
    cudaFree(g->packed);
    cudaFree(g->unk);
    cudaFree(g->upn);
    cudaFree(g->chrom_off);
    cudaFree(g->chrom_len);
    delete g;
    return SB_OK;
}

extern "C" int sb_genome_codes(const sb_genome* g, const int32_t* chrom, const int64_t* start,
                               const uint8_t* strand_rc, int n, int L, uint8_t* codes_out,
                               uint8_t* flags_out, void* stream) {
    SB_REQUIRE(g && chrom && start && codes_out, "sb_genome_codes: null argument");
    SB_REQUIRE(n >= 0 && L > 0, "sb_genome_codes: bad sizes n=%d L=%d", n, L);
    if (n == 0) return SB_OK;
    Perm p = {{0, 1, 2, 3}};
    window_kernel<0><<<window_grid(n, g->device), 256, 0, (cudaStream_t)stream>>>(view_of(g), chrom, start, strand_rc,
                                                                                  n, L, p, codes_out, flags_out);
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    return SB_OK;
}

extern "C" int sb_encode_windows(const sb_genome* g, const int32_t* chrom, const int64_t* start,
                                 const uint8_t* strand_rc, int n, int L, const uint8_t perm[4],
                                 int out_dtype, void* out, uint8_t* flags_out, void* stream) {
    SB_REQUIRE(g && chrom && start && out, "sb_encode_windows: null argument");
    SB_REQUIRE(perm_ok(perm), "sb_encode_windows: perm is not a permutation of 0..3");
    SB_REQUIRE(n >= 0 && L > 0, "sb_encode_windows: bad sizes n=%d L=%d", n, L);
    SB_REQUIRE(out_dtype == SB_F32 || out_dtype == SB_BF16, "sb_encode_windows: bad out_dtype %d",
               out_dtype);
    if (n == 0) return SB_OK;
    const Perm p = make_perm(perm);
    cudaStream_t s = (cudaStream_t)stream;
    if (out_dtype == SB_F32)
        window_kernel<1><<<window_grid(n, g->device), 256, 0, s>>>(view_of(g), chrom, start, strand_rc, n, L, p, out,
                                                                   flags_out);
    else
        window_kernel<2><<<window_grid(n, g->device), 256, 0, s>>>(view_of(g), chrom, start, strand_rc, n, L, p, out,
                                                                   flags_out);
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    return SB_OK;
}

extern "C" int sb_encode_codes(const uint8_t* codes, int64_t n, int L, const uint8_t perm[4],
                               int out_dtype, void* out, void* stream) {
    SB_REQUIRE(perm_ok(perm), "sb_encode_codes: perm is not a permutation of 0..3");
    SB_REQUIRE(n >= 0 && L >= 0, "sb_encode_codes: bad sizes");
    SB_REQUIRE(out_dtype == SB_F32 || out_dtype == SB_BF16, "sb_encode_codes: bad out_dtype %d",
               out_dtype);
    const int64_t total = n * (int64_t)L;
    if (total == 0) return SB_OK;  // empty string -> (0,4) array, sequence.py:14-41
    SB_REQUIRE(codes && out, "sb_encode_codes: null argument");
    const Perm p = make_perm(perm);
    int dev = 0;
    cudaGetDevice(&dev);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)sb_sm_count(dev) * 16;
    if (blocks > cap) blocks = cap;
    cudaStream_t s = (cudaStream_t)stream;
    if (out_dtype == SB_F32)
        codes_onehot_kernel<false><<<(int)blocks, 256, 0, s>>>(codes, total, p, out);
    else
        codes_onehot_kernel<true><<<(int)blocks, 256, 0, s>>>(codes, total, p, out);
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    return SB_OK;
}

extern "C" int sb_ism_enumerate(const uint8_t* codes, int L, int pos0, int pos1,
                                const uint8_t perm[4], int32_t* mut_pos, uint8_t* mut_alt,
                                int32_t* n_mut, int cap, void* stream) {
    SB_REQUIRE(codes && mut_pos && mut_alt && n_mut, "sb_ism_enumerate: null argument");
    SB_REQUIRE(perm_ok(perm), "sb_ism_enumerate: perm is not a permutation of 0..3");
    // the reference's argument checks (_in_silico_mutagenesis.py:64-89) stay on the host; here only
    // memory safety
    SB_REQUIRE(L > 0 && pos0 >= 0 && pos1 <= L && pos0 < pos1, "sb_ism_enumerate: bad range [%d,%d) L=%d",
               pos0, pos1, L);
    ism_enumerate_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(codes, L, pos0, pos1, make_perm(perm),
                                                               mut_pos, mut_alt, n_mut, cap);
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    return SB_OK;
}

extern "C" int sb_ism_expand(const uint8_t* codes, int L, const int32_t* src, const int32_t* mut_pos,
                             const uint8_t* mut_alt, int64_t n_mut, const uint8_t perm[4],
                             int out_dtype, void* out, void* stream) {
    SB_REQUIRE(perm_ok(perm), "sb_ism_expand: perm is not a permutation of 0..3");
    SB_REQUIRE(n_mut >= 0 && L > 0 && n_mut < (1ll << 31), "sb_ism_expand: bad sizes");
    SB_REQUIRE(out_dtype == SB_F32 || out_dtype == SB_BF16, "sb_ism_expand: bad out_dtype %d",
               out_dtype);
    if (n_mut == 0) return SB_OK;
    SB_REQUIRE(codes && mut_pos && mut_alt && out, "sb_ism_expand: null argument");
    const Perm p = make_perm(perm);
    cudaStream_t s = (cudaStream_t)stream;
    if (out_dtype == SB_F32)
        ism_expand_kernel<false><<<(unsigned)n_mut, 256, 0, s>>>(codes, L, src, mut_pos, mut_alt, p, out);
    else
        ism_expand_kernel<true><<<(unsigned)n_mut, 256, 0, s>>>(codes, L, src, mut_pos, mut_alt, p, out);
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    return SB_OK;
}

extern "C" int sb_snv_pairs(const sb_genome* g, const int32_t* chrom, const int64_t* start,
                            const uint8_t* ref_code, const uint8_t* alt_code,
                            const uint8_t* strand_rc, int n, int L, int var_row,
                            const uint8_t perm[4], int out_dtype, void* ref_out, void* alt_out,
                            uint8_t* codes_out, uint8_t* ref_match, uint8_t* flags_out,
                            void* stream) {
    SB_REQUIRE(g && chrom && start && ref_code && alt_code, "sb_snv_pairs: null argument");
    SB_REQUIRE(perm_ok(perm), "sb_snv_pairs: perm is not a permutation of 0..3");
    SB_REQUIRE(n >= 0 && L > 0 && var_row >= 0 && var_row < L, "sb_snv_pairs: bad sizes");
    SB_REQUIRE(out_dtype == SB_F32 || out_dtype == SB_BF16, "sb_snv_pairs: bad out_dtype %d",
               out_dtype);
    if (n == 0) return SB_OK;
    const Perm p = make_perm(perm);
    cudaStream_t s = (cudaStream_t)stream;
    const GenomeView v = view_of(g);
    if (!ref_out && !alt_out)
        snv_pairs_kernel<0><<<n, 256, 0, s>>>(v, chrom, start, ref_code, alt_code, strand_rc, L,
                                              var_row, p, nullptr, nullptr, codes_out, ref_match,
                                              flags_out);
    else if (out_dtype == SB_F32)
        snv_pairs_kernel<1><<<n, 256, 0, s>>>(v, chrom, start, ref_code, alt_code, strand_rc, L,
                                              var_row, p, ref_out, alt_out, codes_out, ref_match,
                                              flags_out);
    else
        snv_pairs_kernel<2><<<n, 256, 0, s>>>(v, chrom, start, ref_code, alt_code, strand_rc, L,
                                              var_row, p, ref_out, alt_out, codes_out, ref_match,
                                              flags_out);
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    return SB_OK;
}
