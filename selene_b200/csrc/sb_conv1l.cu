// selene_b200 — first convolution with MANY taps and a WIDE max-pool (DanQ: 4 -> 320 channels, 26 taps, ReLU,
// pool 13; models/danQ.py:22-28) straight from base codes on tcgen05.
//
// sb_conv1.cu covers <= 8 taps with pool 1/2/4 by folding the pool into the MMA schedule; that needs `pool`
// accumulator blocks and does not stretch to 13.  The generic layer kernel on an (n, L, 8) one-hot is bound by
// TMA row requests (every im2col row is a 128-byte request that brings 16 new bytes) and needs a second kernel
// for the pool.  Here each CTA builds the im2col rows of its tile in shared memory from the uint8 codes
// (K = taps*4 <= 128, un-swizzled K-major core-matrix layout) and the GEMM is TRANSPOSED: the weights are the M
// operand (128 channels per pass) and the im2col rows the N operand (a tile = W whole pooling windows = W*pool
// positions <= 256), so in tensor memory a LANE is a channel and consecutive COLUMNS are consecutive positions.
// A pooling window is then `pool` adjacent columns of one thread: one tcgen05.ld and a chain of max — no
// cross-lane traffic — and a warp stores 32 consecutive channels of a pooled position (64 contiguous bytes).
// The unpooled activations never exist in memory.
//
//   tile            = W pooling windows of one sequence (DanQ: 16 windows = 208 positions, 5 tiles per sequence)
//   accumulators    = 2 stages x (W*pool) columns of TMEM, one 128-channel pass at a time
//   warps 0-7       : build the im2col tile of the NEXT tile, then drain the current tile's passes (two warps per
//                     TMEM lane quadrant, half of the windows each)
//   warp 8 (1 lane) : issues tcgen05.mma (K/16 per pass), commits to mbarriers
#include "sb_common.cuh"

namespace {

constexpr int kPassM = 128;          // channels per pass
constexpr int kMaxTaps = 32;
constexpr int kMaxN = 256;           // positions per tile
constexpr int kMaxCodes = kMaxN + kMaxTaps + 16;

struct Conv1LParams {
    const uint8_t* codes;
    const int32_t* src;
    const int32_t* sub_pos;
    const uint8_t* sub_code;
    long long row0, n_rows;
    int L, taps, pool, t_pool, out_pitch, cout, out_ld, relu, n_pass, tiles_per_seq;
    int kpad;            // K padded to a multiple of 16
    int sbo;             // bytes between 8-row groups = kpad/8 * 128
    int win;             // pooling windows per tile (even)
    int n_pos;           // positions per tile = win * pool, a multiple of 16
    const uint4* wpacked;
    const float* bias;
    __nv_bfloat16* out;
};

// un-swizzled K-major layout: 8-row x 16-byte core matrices; K-adjacent core matrices 128 B apart (LBO),
// M/N-adjacent 8-row groups `sbo` bytes apart
__device__ __forceinline__ uint64_t desc_nosw(uint32_t smem_addr, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(128 >> 4) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// four bf16 of one tap: 1.0 in the slot of the base, or 0.25 x 4 for an unknown base
__device__ __forceinline__ uint64_t tap_pattern(int code) {
    return code < 4 ? (0x3F80ull << (16 * code)) : 0x3E803E803E803E80ull;
}

__global__ void __launch_bounds__(288, 1) conv1_long_kernel(const Conv1LParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (sb_smem_u32(smem_raw) & 1023u)) & 1023u);
    const int x_bytes = (p.n_pos / 8) * p.sbo;             // one im2col tile
    const int w_pass_bytes = (kPassM / 8) * p.sbo;         // weights of one pass
    // layout: [X buf0][X buf1][W: n_pass passes][codes][barriers][tap patterns]
    uint8_t* x_buf = smem;
    uint8_t* w_smem = smem + 2 * x_bytes;
    uint8_t* codes_s = w_smem + p.n_pass * w_pass_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(codes_s + ((kMaxCodes + 15) & ~15));
    uint64_t* x_full = bars;        // [2]
    uint64_t* tfull = bars + 2;     // [2]
    uint64_t* tempty = bars + 4;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
    uint64_t* pat_s = bars + 8;     // [6]: the four bf16 of a tap for codes 0..4, and all-zero (index 5) for K padding

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const long long n_tiles = p.n_rows * p.tiles_per_seq;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            sb_mbar_init(&x_full[i], 256);
            sb_mbar_init(&tfull[i], 1);
            sb_mbar_init(&tempty[i], 256);
        }
        sb_fence_mbar_init();
    }
    if (warp == 8) sb_tmem_alloc(tmem_slot, 512);
    if (threadIdx.x < 6) pat_s[threadIdx.x] = threadIdx.x < 5 ? tap_pattern((int)threadIdx.x) : 0ull;
    {
        // weights -> smem (once per CTA)
        const int n16 = p.n_pass * w_pass_bytes / 16;
        uint4* dst = reinterpret_cast<uint4*>(w_smem);
        for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = __ldg(p.wpacked + i);
    }
    sb_fence_proxy_async();
    sb_tc_fence_before();
    __syncthreads();
    sb_tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int ksteps = p.kpad / 16;

    if (warp == 8) {
        // ===================== MMA issuer: D[channel][position] = sum_k W[channel][k] X[position][k] ==============
        if (lane == 0) {
            const uint32_t idesc = sb_umma_idesc_bf16(kPassM, (uint32_t)p.n_pos);
            const uint32_t x_addr = sb_smem_u32(x_buf);
            const uint32_t w_addr = sb_smem_u32(w_smem);
            uint32_t gpass = 0;
            int it = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                sb_mbar_wait(&x_full[buf], (it >> 1) & 1);
                sb_tc_fence_after();
                for (int pass = 0; pass < p.n_pass; ++pass, ++gpass) {
                    const uint32_t stage = gpass & 1;
                    sb_mbar_wait(&tempty[stage], ((gpass >> 1) & 1) ^ 1);
                    sb_tc_fence_after();
                    const uint32_t d = tmem_base + stage * kMaxN;
                    for (int ks = 0; ks < ksteps; ++ks) {
                        // one K step = 2 core matrices = 256 B along K
                        const uint64_t da = desc_nosw(w_addr + pass * w_pass_bytes + ks * 256, (uint32_t)p.sbo);
                        const uint64_t db = desc_nosw(x_addr + buf * x_bytes + ks * 256, (uint32_t)p.sbo);
                        sb_umma_bf16(d, da, db, idesc, ks > 0 ? 1u : 0u);
                    }
                    sb_umma_commit(&tfull[stage]);
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== im2col builder + epilogue (warps 0..7) =====================
        const int t = threadIdx.x;                // 0..255: im2col row (= position of the tile) this thread builds
        const int quad = warp & 3;                // TMEM lane quadrant: channels quad*32 + lane of the pass
        const int half = warp >> 2;               // which half of the tile's windows this warp drains
        const int n_codes = p.n_pos + p.taps;
        const int kchunks = p.kpad / 8;           // 16-byte chunks (2 taps) per row

        auto build = [&](long long tile, int buf) {
            const long long seq = tile / p.tiles_per_seq;
            const int tin = (int)(tile - seq * p.tiles_per_seq);
            const long long row = p.row0 + seq;
            const long long srow = p.src ? (long long)p.src[row] : row;
            const uint8_t* crow = p.codes + srow * p.L;
            const int sp = p.sub_pos ? p.sub_pos[row] : -1;
            const int sc = (p.sub_pos && p.sub_code) ? p.sub_code[row] : 4;
            const int p0 = tin * p.n_pos;
            for (int k = t; k < n_codes; k += 256) {
                const int pos = p0 + k;
                int c = 4;
                if (pos < p.L) {
                    c = __ldg(crow + pos);
                    if (pos == sp) c = sc;
                    if (c > 4) c = 4;
                }
                codes_s[k] = (uint8_t)c;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (t < p.n_pos) {
                uint8_t* x_row = x_buf + (size_t)buf * x_bytes + (t >> 3) * p.sbo + (t & 7) * 16;
                for (int kc = 0; kc < kchunks; ++kc) {
                    const int c0 = (2 * kc < p.taps) ? codes_s[t + 2 * kc] : 5;
                    const int c1 = (2 * kc + 1 < p.taps) ? codes_s[t + 2 * kc + 1] : 5;
                    const uint2 lo = *reinterpret_cast<const uint2*>(pat_s + c0);
                    const uint2 hi = *reinterpret_cast<const uint2*>(pat_s + c1);
                    *reinterpret_cast<uint4*>(x_row + kc * 128) = make_uint4(lo.x, lo.y, hi.x, hi.y);
                }
            }
            sb_fence_proxy_async();      // generic-proxy smem writes -> visible to the tensor core
            sb_mbar_arrive(&x_full[buf]);
            asm volatile("bar.sync 1, 256;" ::: "memory");  // codes_s may be overwritten by the next build
        };

        const int wins_half = p.win / 2;
        uint32_t gpass = 0;
        int it = 0;
        long long tile = blockIdx.x;
        if (tile < n_tiles) build(tile, 0);
        for (; tile < n_tiles; tile += gridDim.x, ++it) {
            const long long next = tile + gridDim.x;
            if (next < n_tiles) build(next, (it + 1) & 1);
            const long long seq = tile / p.tiles_per_seq;
            const int tin = (int)(tile - seq * p.tiles_per_seq);
            const int tp0 = tin * p.win + half * wins_half;            // first pooled position of this warp
            __nv_bfloat16* obase = p.out + (seq * p.out_pitch + tp0) * (long long)p.out_ld;
            for (int pass = 0; pass < p.n_pass; ++pass, ++gpass) {
                const uint32_t stage = gpass & 1;
                const int ch = pass * kPassM + quad * 32 + lane;
                const float bias = ch < p.cout ? __ldg(p.bias + ch) : 0.f;
                sb_mbar_wait(&tfull[stage], (gpass >> 1) & 1);
                sb_tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + stage * kMaxN + half * wins_half * p.pool;
                // two windows in flight: the second's columns are loading while the first is reduced
                uint32_t ra[16], rb[16];
                sb_tmem_ld16(taddr, ra);
                for (int w = 0; w < wins_half; w += 2) {
                    sb_tmem_wait_ld();
                    if (w + 1 < wins_half) sb_tmem_ld16(taddr + (w + 1) * p.pool, rb);
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if (u == 1) {
                            if (w + 1 >= wins_half) break;
                            sb_tmem_wait_ld();
                            if (w + 2 < wins_half) sb_tmem_ld16(taddr + (w + 2) * p.pool, ra);
                        }
                        float m = __uint_as_float(u == 0 ? ra[0] : rb[0]);
#pragma unroll
                        for (int j = 1; j < 16; ++j)
                            if (j < p.pool) m = fmaxf(m, __uint_as_float(u == 0 ? ra[j] : rb[j]));
                        m += bias;
                        if (p.relu) m = fmaxf(m, 0.f);
                        const int tp = tp0 + w + u;
                        if (tp < p.t_pool && ch < p.out_ld)
                            obase[(long long)(w + u) * p.out_ld + ch] = __float2bfloat16_rn(m);
                    }
                }
                sb_tc_fence_before();
                sb_mbar_arrive(&tempty[stage]);
            }
        }
    }

    sb_tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        sb_tc_fence_after();
        sb_tmem_dealloc(tmem_base, 512);
    }
}

struct Conv1LShape { int kpad, sbo, win, n_pos, n_pass; size_t smem; };

bool conv1l_shape(int taps, int pool, int cout, Conv1LShape* s) {
    if (!(taps >= 1 && taps <= kMaxTaps && pool >= 2 && pool <= 15 && cout >= 1 && cout <= 1024)) return false;
    s->kpad = (taps * 4 + 15) / 16 * 16;
    s->sbo = s->kpad / 8 * 128;
    // a multiple of 16 windows (so win * pool is a multiple of 16) within 240 positions: the 16-column TMEM read of
    // the last window then stays inside the stage's 256 columns
    s->win = 16 * (240 / (16 * pool));
    s->n_pos = s->win * pool;
    s->n_pass = (cout + kPassM - 1) / kPassM;
    s->smem = 1024 + 2 * (size_t)(s->n_pos / 8) * s->sbo + (size_t)s->n_pass * (kPassM / 8) * s->sbo +
              ((kMaxCodes + 15) & ~15) + 192;
    return s->win >= 2 && s->smem <= 220 * 1024;
}

}  // namespace

int sb_conv1_long_supported(int taps, int pool, int cout) {
    Conv1LShape s;
    return conv1l_shape(taps, pool, cout, &s) ? 1 : 0;
}

// Host: pack conv weights (cout, 4, taps) fp32 into the smem core-matrix layout, bf16, K = tap*4 + base padded to a
// multiple of 16, 128-channel passes (missing channels are zero rows).  `out` holds n_pass * 128 * kpad elements
// (pass NULL to get the element count).
size_t sb_conv1_long_pack_weights(const float* w, int cout, int taps, const uint8_t perm[4], uint16_t* out,
                                  uint16_t (*to_bf16)(float)) {
    const int n_pass = (cout + kPassM - 1) / kPassM;
    const int kpad = (taps * 4 + 15) / 16 * 16;
    const size_t elems = (size_t)n_pass * kPassM * kpad;
    if (!out) return elems;
    for (size_t i = 0; i < elems; ++i) out[i] = 0;
    const size_t group = (size_t)kpad * 8;      // elements of one 8-row group
    for (int n = 0; n < cout; ++n)
        for (int tap = 0; tap < taps; ++tap)
            for (int b = 0; b < 4; ++b) {
                const int k = tap * 4 + b;
                const size_t off = (size_t)(n >> 3) * group + (size_t)(k >> 3) * 64 + (size_t)(n & 7) * 8 + (k & 7);
                out[off] = to_bf16(w[((size_t)n * 4 + perm[b]) * taps + tap]);
            }
    return elems;
}

// bias: fp32, at least cout entries
int sb_conv1_long_launch(const uint8_t* codes, const int32_t* src, const int32_t* sub_pos, const uint8_t* sub_code,
                         long long row0, long long n_rows, int L, int taps, int pool, int t_pool, int out_pitch, int cout,
                         int out_ld, int relu, const void* wpacked, const float* bias, void* out, int device,
                         cudaStream_t stream) {
    if (n_rows <= 0) return SB_OK;
    Conv1LShape sh;
    SB_REQUIRE(conv1l_shape(taps, pool, cout, &sh), "sb_conv1_long_launch: taps %d / pool %d / cout %d unsupported", taps,
               pool, cout);
    Conv1LParams p;
    p.codes = codes; p.src = src; p.sub_pos = sub_pos; p.sub_code = sub_code;
    p.row0 = row0; p.n_rows = n_rows;
    p.L = L; p.taps = taps; p.pool = pool; p.t_pool = t_pool; p.out_pitch = out_pitch > 0 ? out_pitch : t_pool;
    p.cout = cout; p.out_ld = out_ld; p.relu = relu;
    p.n_pass = sh.n_pass; p.kpad = sh.kpad; p.sbo = sh.sbo; p.win = sh.win; p.n_pos = sh.n_pos;
    p.tiles_per_seq = (t_pool + p.win - 1) / p.win;
    p.wpacked = reinterpret_cast<const uint4*>(wpacked);
    p.bias = bias;
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    size_t smem = sh.smem;
    if (smem < 120 * 1024) smem = 120 * 1024;  // one CTA per SM: each CTA owns all 512 TMEM columns
    static size_t attr = 0;
    if (smem > attr) {
        SB_CHECK_CUDA(cudaFuncSetAttribute(conv1_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    const long long tiles = n_rows * p.tiles_per_seq;
    const int sms = sb_sm_count(device);
    const int grid = (int)(tiles < sms ? tiles : sms);
    conv1_long_kernel<<<grid, 288, smem, stream>>>(p);
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    return SB_OK;
}
