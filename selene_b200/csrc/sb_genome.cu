// selene_b200 — kernel (a): genome image -> codes / one-hot windows, ISM expansion, SNV ref/alt pairs.
//
// All of these are HBM-bound byte movers: a 1000-bp fp32 window reads 250 B of 2-bit codes +
// 2 x 125 B of masks and writes 16 000 B, so the design goal is simply "every store is a full
// 128-bit coalesced store, every load is a full word, and nothing is read twice from DRAM".
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
// One WARP per window (or SNV), grid-stride.  A window is handled in chunks of 512 output positions:
//
//   decode   lane l owns output positions c0 + 16 l ... + 15.  They are 16 CONSECUTIVE genome bases (descending for
//            the minus strand), so the lane fetches them with two aligned 32-bit loads of the 2-bit image and two of
//            each mask, funnel-shifted to the base offset — 6 load instructions per lane per chunk instead of two
//            dependent byte loads per base — expands them to 16 code bytes with bit spreads (complemented and
//            byte-reversed for the minus strand, 4 = unknown where the mask or the chromosome end says so) and parks
//            them in the warp's 512-byte slice of shared memory as one 16-byte store;
//   emit     lane l then reads the codes of the positions it STORES: one base per lane and instruction for fp32
//            (a warp writes 512 contiguous bytes), two for bf16, sixteen for code rows, always one 16-byte store.
//
// Flags are mask arithmetic on the decode side plus one warp OR-reduction.  MODE 0: codes only; 1: one-hot fp32;
// 2: one-hot bf16.  SNV = false: sb_encode_windows / sb_genome_codes (the minus strand reverse-complements the
// in-bounds part only, genome.py:164-166).  SNV = true: sb_snv_pairs — ref and alt tensors with the VCF alleles forced
// at the variant row, ref_match, the strand applied to the finished encodings (model_predict.py:1102-1110).
constexpr int kWinChunk = 512;
constexpr int kWinWarps = 8;
constexpr int kWinMaxChromSmem = 256;   // chromosome offsets / lengths staged in shared memory up to this many

__device__ __forceinline__ uint32_t spread_codes4(uint32_t bits8) {   // 4 x 2-bit fields -> 4 bytes
    uint32_t x = bits8 & 0xFFu;
    x = (x | (x << 12)) & 0x000F000Fu;
    return (x | (x << 6)) & 0x03030303u;
}
__device__ __forceinline__ uint32_t spread_bits4(uint32_t bits4) {    // 4 bits -> bit 0 of 4 bytes
    uint32_t m = bits4 & 0xFu;
    m = (m | (m << 14)) & 0x00030003u;
    return (m | (m << 7)) & 0x01010101u;
}

struct WinMeta {          // what is read per window from the caller's arrays
    int c;
    int64_t st;
    int rc, sub_ref, sub_alt;
};
struct WinRaw {           // the image words of one lane's 16 positions
    uint32_t pw, uw, nw, vm;
};

template <int MODE, bool SNV>
__global__ void __launch_bounds__(kWinWarps * 32, 6) window_kernel(
    GenomeView g, int64_t n_image, int n_chrom, const int32_t* __restrict__ chrom, const int64_t* __restrict__ start,
    const uint8_t* __restrict__ strand_rc, int n, int L, Perm perm, void* __restrict__ out,
    uint8_t* __restrict__ flags_out,
    // SNV only
    const uint8_t* __restrict__ ref_code, const uint8_t* __restrict__ alt_code, int var_row,
    void* __restrict__ alt_out, uint8_t* __restrict__ codes_out, uint8_t* __restrict__ ref_match) {
    __shared__ __align__(16) uint8_t stage[kWinWarps][kWinChunk];
    __shared__ __align__(16) uint2 lut_bf16[5];      // bf16 one-hot rows by code, in the caller's channel order
    __shared__ int64_t chrom_meta[2 * kWinMaxChromSmem];
    const int lane = threadIdx.x & 31;
    uint8_t* sm = stage[threadIdx.x >> 5];
    if (threadIdx.x < 5) lut_bf16[threadIdx.x] = onehot_bf16(threadIdx.x, perm);
    const bool meta_smem = n_chrom <= kWinMaxChromSmem;
    if (meta_smem)
        for (int i = threadIdx.x; i < n_chrom; i += blockDim.x) {
            chrom_meta[i] = g.chrom_off[i];
            chrom_meta[kWinMaxChromSmem + i] = g.chrom_len[i];
        }
    __syncthreads();
    const uint32_t* __restrict__ packed32 = reinterpret_cast<const uint32_t*>(g.packed);
    const uint32_t* __restrict__ unk32 = reinterpret_cast<const uint32_t*>(g.unk);
    const uint32_t* __restrict__ upn32 = reinterpret_cast<const uint32_t*>(g.upn);
    const int64_t max_w2 = (n_image >> 4) + 1, max_w1 = (n_image >> 5) + 1;   // the allocations are padded by 16 bytes
    const int warps_total = gridDim.x * kWinWarps;

    auto load_meta = [&](int w) {
        WinMeta m;
        m.c = chrom[w];
        m.st = start[w];
        m.rc = strand_rc != nullptr && strand_rc[w] != 0;
        m.sub_ref = SNV ? ref_code[w] : 4;
        m.sub_alt = SNV ? alt_code[w] : 4;
        return m;
    };
    // window geometry: in-bounds genome range [s0, e0), image offset, and the genome coordinate of output position p,
    // base + p (plus strand) or base - p (minus strand)
    auto geometry = [&](const WinMeta& m, int64_t& coff, int64_t& s0, int64_t& e0, int64_t& base) {
        coff = meta_smem ? chrom_meta[m.c] : g.chrom_off[m.c];
        const int64_t clen = meta_smem ? chrom_meta[kWinMaxChromSmem + m.c] : g.chrom_len[m.c];
        const int64_t end = m.st + L;
        s0 = m.st < 0 ? 0 : m.st;
        e0 = end > clen ? clen : end;
        if (e0 < s0) e0 = s0;
        base = !m.rc ? m.st : (SNV ? end - 1 : e0 - 1 + (s0 - m.st));
    };
    auto load_raw = [&](int64_t coff, int64_t s0, int64_t e0, int64_t base, int rc, int c0) {
        WinRaw r;
        const int p0 = c0 + 16 * lane;
        const int64_t lo = rc ? base - (p0 + 15) : base + p0;           // lowest genome coordinate of the 16
        const int64_t a64 = s0 - lo, b64 = e0 - lo;
        const int a = a64 < 0 ? 0 : (a64 > 16 ? 16 : (int)a64);
        const int b = b64 < 0 ? 0 : (b64 > 16 ? 16 : (int)b64);
        r.vm = b > a ? (((1u << b) - 1u) & ~((1u << a) - 1u)) : 0u;     // in-bounds bases (ascending-coordinate space)
        r.pw = r.uw = r.nw = 0;
        if (r.vm != 0u && p0 < L) {
            const int64_t gi = coff + lo;
            int64_t w2 = gi >> 4, w1 = gi >> 5;
            w2 = w2 < 0 ? 0 : (w2 > max_w2 ? max_w2 : w2);
            w1 = w1 < 0 ? 0 : (w1 > max_w1 ? max_w1 : w1);
            r.pw = __funnelshift_r(__ldg(packed32 + w2), __ldg(packed32 + w2 + 1), (uint32_t)(gi & 15) * 2u);
            r.uw = __funnelshift_r(__ldg(unk32 + w1), __ldg(unk32 + w1 + 1), (uint32_t)(gi & 31)) & 0xFFFFu;
            r.nw = __funnelshift_r(__ldg(upn32 + w1), __ldg(upn32 + w1 + 1), (uint32_t)(gi & 31)) & 0xFFFFu;
        }
        return r;
    };

    // No software prefetch across windows or chunks: an earlier version that loaded the next chunk's words and the next
    // window's coordinates ahead needed 48+ registers, and the lost occupancy cost the (store-bound) fp32 path more
    // than the hidden load latency gained (r02: 0.81 -> 0.75 of the measured copy bandwidth).
    for (int w = blockIdx.x * kWinWarps + (threadIdx.x >> 5); w < n; w += warps_total) {
        const WinMeta m = load_meta(w);
        int64_t coff, s0, e0, base;
        geometry(m, coff, s0, e0, base);
        const int rc = m.rc;
        int sub_at = -1, sub_ref = m.sub_ref, sub_alt = m.sub_alt;
        if (SNV) {
            sub_at = rc ? L - 1 - var_row : var_row;
            if (rc) {
                sub_ref = sub_ref < 4 ? 3 - sub_ref : 4;
                sub_alt = sub_alt < 4 ? 3 - sub_alt : 4;
            }
        }
        int flags = 0;
        for (int c0 = 0; c0 < L; c0 += kWinChunk) {
            const WinRaw raw = load_raw(coff, s0, e0, base, rc, c0);
            // ---------------------------------------------- decode this lane's 16 positions
            const int p0 = c0 + 16 * lane;
            uint4 v;
            {
                const uint32_t vm = raw.vm, pw = raw.pw, uw = raw.uw, nw = raw.nw;
                const uint32_t um = (uw & vm) | (~vm & 0xFFFFu);            // unknown: masked bases and padding
                const uint32_t nm = (nw & uw & vm) | (~vm & 0xFFFFu);       // 'N': upper-case N bases and padding
                const int cnt = L - p0 < 0 ? 0 : (L - p0 > 16 ? 16 : L - p0);
                const uint32_t tm = (1u << cnt) - 1u;                       // positions of this lane inside the window
                const uint32_t pm = rc ? (__brev(tm) >> 16) : tm;           // the same in ascending-coordinate space
                if (um & pm) flags |= SB_FLAG_HAS_UNK;
                if (nm & pm) flags |= SB_FLAG_HAS_UPPER_N;
                uint32_t wd[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t x = spread_codes4(pw >> (8 * k));
                    const uint32_t mk = spread_bits4(um >> (4 * k));
                    if (rc) x ^= 0x03030303u;                               // complement: code -> 3 - code
                    wd[k] = (x & ((mk ^ 0x01010101u) * 3u)) | (mk << 2);    // unknown -> 4
                }
                if (!rc) v = make_uint4(wd[0], wd[1], wd[2], wd[3]);
                else v = make_uint4(__byte_perm(wd[3], 0, 0x0123), __byte_perm(wd[2], 0, 0x0123),
                                    __byte_perm(wd[1], 0, 0x0123), __byte_perm(wd[0], 0, 0x0123));
            }
            const int c1 = L - c0 < kWinChunk ? L - c0 : kWinChunk;         // positions in this chunk
            const int64_t row0 = (int64_t)w * L + c0;
            // ---------------------------------------------- code rows and ref_match come straight from registers
            if (SNV && ref_match && sub_at >= p0 && sub_at < p0 + 16) {
                const int k = sub_at - p0;
                const uint32_t word = k < 4 ? v.x : (k < 8 ? v.y : (k < 12 ? v.z : v.w));
                ref_match[w] = (uint8_t)((int)((word >> (8 * (k & 3))) & 0xFFu) == sub_ref);
            }
            if (MODE == 0 || (SNV && codes_out != nullptr)) {
                uint8_t* cp = (MODE == 0 && !SNV) ? reinterpret_cast<uint8_t*>(out) : codes_out;
                const int i = 16 * lane;
                if (i < c1) {
                    uint8_t* p = cp + row0 + i;
                    const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
                    if (i + 16 <= c1 && (addr & 15) == 0) {
                        *reinterpret_cast<uint4*>(p) = v;
                    } else if (i + 16 <= c1 && (addr & 7) == 0) {
                        *reinterpret_cast<uint2*>(p) = make_uint2(v.x, v.y);
                        *reinterpret_cast<uint2*>(p + 8) = make_uint2(v.z, v.w);
                    } else {
                        const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int b = 0; b < 16; ++b)
                            if (i + b < c1) p[b] = (uint8_t)(wds[b >> 2] >> (8 * (b & 3)));
                    }
                }
            }
            // ---------------------------------------------- one-hot rows: stage the codes, store 16 bytes per lane
            if (MODE != 0) {
                reinterpret_cast<uint4*>(sm)[lane] = v;
                __syncwarp();
                if (MODE == 1) {
                    float4* o4 = reinterpret_cast<float4*>(out) + row0;
                    float4* a4 = reinterpret_cast<float4*>(alt_out) + row0;
#pragma unroll 4
                    for (int i = lane; i < c1; i += 32) {
                        const int code = sm[i];
                        if (SNV) {
                            const bool hit = c0 + i == sub_at;
                            if (out) o4[i] = onehot_f32(hit ? sub_ref : code, perm);
                            if (alt_out) a4[i] = onehot_f32(hit ? sub_alt : code, perm);
                        } else {
                            o4[i] = onehot_f32(code, perm);      // selects: the shared-memory pipe is busy with the codes
                        }
                    }
                } else {
                    const bool aligned = (row0 & 1) == 0;                   // 16-byte alignment of position pairs
#pragma unroll 4
                    for (int i = 2 * lane; i < c1; i += 64) {
                        const bool two = i + 1 < c1;
                        const int ca = sm[i], cb = two ? sm[i + 1] : 4;
#pragma unroll
                        for (int which = 0; which < (SNV ? 2 : 1); ++which) {
                            void* dst = which == 0 ? out : alt_out;
                            if (dst == nullptr) continue;
                            int xa = ca, xb = cb;
                            if (SNV) {
                                const int sub = which == 0 ? sub_ref : sub_alt;
                                if (c0 + i == sub_at) xa = sub;
                                if (c0 + i + 1 == sub_at) xb = sub;
                            }
                            uint2* p = reinterpret_cast<uint2*>(dst) + row0 + i;
                            const uint2 ua = lut_bf16[xa], ub = lut_bf16[xb];
                            if (two && aligned) {
                                *reinterpret_cast<uint4*>(p) = make_uint4(ua.x, ua.y, ub.x, ub.y);
                            } else {
                                p[0] = ua;
                                if (two) p[1] = ub;
                            }
                        }
                    }
                }
                __syncwarp();
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
    const long long blocks = ((long long)n + kWinWarps - 1) / kWinWarps;   // one warp per window
    const long long cap = (long long)sb_sm_count(device) * 6;              // persistent: 6 CTAs of 8 warps per SM
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
    window_kernel<0, false><<<window_grid(n, g->device), kWinWarps * 32, 0, (cudaStream_t)stream>>>(
        view_of(g), g->n_image, g->n_chrom, chrom, start, strand_rc, n, L, p, codes_out, flags_out, nullptr, nullptr, 0, nullptr,
        nullptr, nullptr);
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
        window_kernel<1, false><<<window_grid(n, g->device), kWinWarps * 32, 0, s>>>(
            view_of(g), g->n_image, g->n_chrom, chrom, start, strand_rc, n, L, p, out, flags_out, nullptr, nullptr, 0, nullptr, nullptr,
            nullptr);
    else
        window_kernel<2, false><<<window_grid(n, g->device), kWinWarps * 32, 0, s>>>(
            view_of(g), g->n_image, g->n_chrom, chrom, start, strand_rc, n, L, p, out, flags_out, nullptr, nullptr, 0, nullptr, nullptr,
            nullptr);
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
    const int grid = window_grid(n, g->device);
    if (!ref_out && !alt_out)
        window_kernel<0, true><<<grid, kWinWarps * 32, 0, s>>>(v, g->n_image, g->n_chrom, chrom, start, strand_rc, n, L, p, nullptr,
                                                               flags_out, ref_code, alt_code, var_row, nullptr, codes_out,
                                                               ref_match);
    else if (out_dtype == SB_F32)
        window_kernel<1, true><<<grid, kWinWarps * 32, 0, s>>>(v, g->n_image, g->n_chrom, chrom, start, strand_rc, n, L, p, ref_out,
                                                               flags_out, ref_code, alt_code, var_row, alt_out, codes_out,
                                                               ref_match);
    else
        window_kernel<2, true><<<grid, kWinWarps * 32, 0, s>>>(v, g->n_image, g->n_chrom, chrom, start, strand_rc, n, L, p, ref_out,
                                                               flags_out, ref_code, alt_code, var_row, alt_out, codes_out,
                                                               ref_match);
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    return SB_OK;
}
