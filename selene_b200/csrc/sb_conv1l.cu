// selene_b200 — first convolution with MANY taps and a WIDE max-pool (DanQ: 4 -> 320 channels, 26 taps, ReLU,
// pool 13; models/danQ.py:22-28) straight from base codes on tcgen05.
//
// sb_conv1.cu covers <= 8 taps with pool 1/2/4 by folding the pool into the MMA schedule; that needs `pool`
// accumulator blocks and does not stretch to 13.  The generic layer kernel on an (n, L, 8) one-hot is bound by
// TMA row requests (every im2col row is a 128-byte request that brings 16 new bytes) and needs a second kernel
// for the pool.  Here, as in sb_conv1.cu, each CTA builds the im2col rows of its tile in shared memory from the
// uint8 codes (K = taps*4 <= 128, un-swizzled K-major core-matrix layout) — and because the builder decides which
// conv position a tile row holds, every 32-row TMEM lane quadrant is given floor(32/pool) WHOLE pooling windows
// (pool 13: rows 0-25 = two windows, rows 26-31 stay zero).  A pooling window is then a group of adjacent lanes
// of one warp: the epilogue reduces it with a segmented shuffle maximum and the window's first lane stores 16
// channels (32 contiguous bytes).  The unpooled activations never exist in memory.
//
//   tile            = 4 quadrants x floor(32/pool) windows of one sequence (DanQ: 8 windows = 104 positions)
//   accumulators    = 2 stages x 64 columns of TMEM, one 64-channel pass at a time
//   warps 0-7       : build A for the NEXT tile, then drain the current tile's passes (two warps per TMEM lane
//                     quadrant, alternate 16-column chunks)
//   warp 8 (1 lane) : issues tcgen05.mma (K/16 per pass), commits to mbarriers
#include "sb_common.cuh"

namespace {

constexpr int kRows = 128;
constexpr int kPassN = 64;
constexpr int kMaxTaps = 32;
constexpr int kMaxCodes = 128 + kMaxTaps + 16;

struct Conv1LParams {
    const uint8_t* codes;
    const int32_t* src;
    const int32_t* sub_pos;
    const uint8_t* sub_code;
    long long row0, n_rows;
    int L, taps, pool, t_pool, out_pitch, cout, out_ld, relu, n_pass, tiles_per_seq;
    int kpad;            // K padded to a multiple of 16
    int sbo;             // bytes between 8-row groups = kpad/8 * 128
    int win_per_quad;    // floor(32 / pool)
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

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(288, 1) conv1_long_kernel(const Conv1LParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (sb_smem_u32(smem_raw) & 1023u)) & 1023u);
    const int a_bytes = (kRows / 8) * p.sbo;
    const int b_pass_bytes = (kPassN / 8) * p.sbo;
    // layout: [A buf0][A buf1][B: n_pass passes][bias][codes][barriers]
    uint8_t* a_buf = smem;
    uint8_t* b_smem = smem + 2 * a_bytes;
    float* bias_s = reinterpret_cast<float*>(b_smem + p.n_pass * b_pass_bytes);
    uint8_t* codes_s = reinterpret_cast<uint8_t*>(bias_s + p.n_pass * kPassN);
    uint64_t* bars = reinterpret_cast<uint64_t*>(codes_s + ((kMaxCodes + 15) & ~15));
    uint64_t* a_full = bars;        // [2]
    uint64_t* tfull = bars + 2;     // [2]
    uint64_t* tempty = bars + 4;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const long long n_tiles = p.n_rows * p.tiles_per_seq;
    const int rows_per_quad = p.win_per_quad * p.pool;    // active rows of a quadrant
    const int pos_per_tile = 4 * rows_per_quad;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            sb_mbar_init(&a_full[i], 256);
            sb_mbar_init(&tfull[i], 1);
            sb_mbar_init(&tempty[i], 256);
        }
        sb_fence_mbar_init();
    }
    if (warp == 8) sb_tmem_alloc(tmem_slot, 128);
    {
        // weights + bias -> smem (once per CTA); both A buffers zeroed: idle rows and the K padding stay zero
        const int n16 = p.n_pass * b_pass_bytes / 16;
        uint4* dst = reinterpret_cast<uint4*>(b_smem);
        for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = __ldg(p.wpacked + i);
        for (int i = threadIdx.x; i < p.n_pass * kPassN; i += blockDim.x) bias_s[i] = __ldg(p.bias + i);
        uint4* az = reinterpret_cast<uint4*>(a_buf);
        for (int i = threadIdx.x; i < 2 * a_bytes / 16; i += blockDim.x) az[i] = make_uint4(0, 0, 0, 0);
    }
    sb_fence_proxy_async();
    sb_tc_fence_before();
    __syncthreads();
    sb_tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int ksteps = p.kpad / 16;

    if (warp == 8) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = sb_umma_idesc_bf16(kRows, kPassN);
            const uint32_t a_addr = sb_smem_u32(a_buf);
            const uint32_t b_addr = sb_smem_u32(b_smem);
            uint32_t gpass = 0;
            int it = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                sb_mbar_wait(&a_full[buf], (it >> 1) & 1);
                sb_tc_fence_after();
                for (int pass = 0; pass < p.n_pass; ++pass, ++gpass) {
                    const uint32_t stage = gpass & 1;
                    sb_mbar_wait(&tempty[stage], ((gpass >> 1) & 1) ^ 1);
                    sb_tc_fence_after();
                    const uint32_t d = tmem_base + stage * kPassN;
                    for (int ks = 0; ks < ksteps; ++ks) {
                        // one K step = 2 core matrices = 256 B along K
                        const uint64_t da = desc_nosw(a_addr + buf * a_bytes + ks * 256, (uint32_t)p.sbo);
                        const uint64_t db = desc_nosw(b_addr + pass * b_pass_bytes + ks * 256, (uint32_t)p.sbo);
                        sb_umma_bf16(d, da, db, idesc, ks > 0 ? 1u : 0u);
                    }
                    sb_umma_commit(&tfull[stage]);
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== A builder + epilogue (warps 0..7) =====================
        const int i = threadIdx.x & 127;          // tile row
        const int half = threadIdx.x >> 7;        // builder: which K chunks; epilogue: which 16-column chunks
        const int quad = i >> 5, ql = i & 31;     // TMEM lane quadrant and lane
        const bool active = ql < rows_per_quad;
        const int rel = quad * rows_per_quad + ql;            // position of this row relative to the tile start
        const int n_codes = pos_per_tile + p.taps;
        const int kchunks = p.kpad / 8;                       // 16-byte chunks (2 taps) per row

        auto build = [&](long long tile, int buf) {
            const long long seq = tile / p.tiles_per_seq;
            const int tin = (int)(tile - seq * p.tiles_per_seq);
            const long long row = p.row0 + seq;
            const long long srow = p.src ? (long long)p.src[row] : row;
            const uint8_t* crow = p.codes + srow * p.L;
            const int sp = p.sub_pos ? p.sub_pos[row] : -1;
            const int sc = (p.sub_pos && p.sub_code) ? p.sub_code[row] : 4;
            const int p0 = tin * pos_per_tile;
            for (int k = threadIdx.x; k < n_codes; k += 256) {
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
            if (active) {
                uint8_t* a_row = a_buf + (size_t)buf * a_bytes + (i >> 3) * p.sbo + (i & 7) * 16;
                // the two halves of the builder take alternate 16-byte chunks (= tap pairs) of the row
                for (int kc = half; kc < kchunks; kc += 2) {
                    const uint64_t lo = (2 * kc < p.taps) ? tap_pattern(codes_s[rel + 2 * kc]) : 0ull;
                    const uint64_t hi = (2 * kc + 1 < p.taps) ? tap_pattern(codes_s[rel + 2 * kc + 1]) : 0ull;
                    *reinterpret_cast<uint4*>(a_row + kc * 128) =
                        make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
                }
            }
            sb_fence_proxy_async();      // generic-proxy smem writes -> visible to the tensor core
            sb_mbar_arrive(&a_full[buf]);
            asm volatile("bar.sync 1, 256;" ::: "memory");  // codes_s may be overwritten by the next build
        };

        const int g = ql / p.pool;                 // window of this lane inside its quadrant
        const int li = ql - g * p.pool;            // lane inside the window
        uint32_t gpass = 0;
        int it = 0;
        long long tile = blockIdx.x;
        if (tile < n_tiles) build(tile, 0);
        for (; tile < n_tiles; tile += gridDim.x, ++it) {
            const long long next = tile + gridDim.x;
            if (next < n_tiles) build(next, (it + 1) & 1);
            const long long seq = tile / p.tiles_per_seq;
            const int tin = (int)(tile - seq * p.tiles_per_seq);
            const int tp = (tin * 4 + quad) * p.win_per_quad + g;       // pooled position of this lane's window
            const bool writer = active && li == 0 && tp < p.t_pool;
            __nv_bfloat16* orow = p.out + (seq * p.out_pitch + tp) * (long long)p.out_ld;
            for (int pass = 0; pass < p.n_pass; ++pass, ++gpass) {
                const uint32_t stage = gpass & 1;
                sb_mbar_wait(&tfull[stage], (gpass >> 1) & 1);
                sb_tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + stage * kPassN;
                uint32_t ra[16], rb[16];
                sb_tmem_ld16(taddr + half * 16, ra);
                sb_tmem_ld16(taddr + (half + 2) * 16, rb);
                sb_tmem_wait_ld();
                sb_tc_fence_before();
                sb_mbar_arrive(&tempty[stage]);      // accumulators are in registers: release them early
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int col0 = pass * kPassN + (half + 2 * c) * 16;
                    // + bias, ReLU, round to bf16 (rounding is monotonic, so the maximum of the rounded values is
                    // the rounded maximum) and pack two channels per register: half the shuffles
                    __nv_bfloat162 v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float a = __uint_as_float(c == 0 ? ra[2 * j] : rb[2 * j]) + bias_s[col0 + 2 * j];
                        float b = __uint_as_float(c == 0 ? ra[2 * j + 1] : rb[2 * j + 1]) + bias_s[col0 + 2 * j + 1];
                        if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                        v[j] = __floats2bfloat162_rn(a, b);
                    }
                    // segmented maximum by doubling: after the step with distance d, lane i holds the maximum over
                    // [i, i + 2d) clipped to its window; the window's first lane ends up with the window maximum
                    for (int d = 1; d < p.pool; d <<= 1) {
                        const bool take = li + d < p.pool;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint32_t mine = *reinterpret_cast<uint32_t*>(&v[j]);
                            uint32_t o = __shfl_down_sync(0xffffffffu, mine, d);
                            if (!take) o = mine;
                            v[j] = __hmax2(v[j], *reinterpret_cast<__nv_bfloat162*>(&o));
                        }
                    }
                    if (writer) {
                        uint4* o = reinterpret_cast<uint4*>(orow + col0);
                        const uint32_t* w = reinterpret_cast<const uint32_t*>(v);
                        if (col0 + 8 <= p.out_ld) o[0] = make_uint4(w[0], w[1], w[2], w[3]);
                        if (col0 + 16 <= p.out_ld) o[1] = make_uint4(w[4], w[5], w[6], w[7]);
                    }
                }
            }
        }
    }

    sb_tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        sb_tc_fence_after();
        sb_tmem_dealloc(tmem_base, 128);
    }
}

size_t conv1l_smem(int taps, int n_pass) {
    const int kpad = (taps * 4 + 15) / 16 * 16;
    const size_t sbo = (size_t)kpad / 8 * 128;
    return 1024 + 2 * (kRows / 8) * sbo + (size_t)n_pass * (kPassN / 8) * sbo + (size_t)n_pass * kPassN * 4 +
           ((kMaxCodes + 15) & ~15) + 128;
}

}  // namespace

int sb_conv1_long_supported(int taps, int pool, int cout) {
    if (!(taps >= 1 && taps <= kMaxTaps && pool >= 2 && pool <= 32 && cout >= 1 && cout <= 1024)) return 0;
    return conv1l_smem(taps, (cout + kPassN - 1) / kPassN) <= 200 * 1024;
}

// Host: pack conv weights (cout, 4, taps) fp32 into the smem core-matrix layout, bf16, K = tap*4 + base padded to a
// multiple of 16, 64-channel passes.  `out` holds n_pass * 64 * kpad elements.  Returns the element count.
size_t sb_conv1_long_pack_weights(const float* w, int cout, int taps, const uint8_t perm[4], uint16_t* out,
                                  uint16_t (*to_bf16)(float)) {
    const int n_pass = (cout + kPassN - 1) / kPassN;
    const int kpad = (taps * 4 + 15) / 16 * 16;
    const size_t elems = (size_t)n_pass * kPassN * kpad;
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

int sb_conv1_long_launch(const uint8_t* codes, const int32_t* src, const int32_t* sub_pos, const uint8_t* sub_code,
                         long long row0, long long n_rows, int L, int taps, int pool, int t_pool, int out_pitch, int cout,
                         int out_ld, int relu, const void* wpacked, const float* bias_padded, void* out, int device,
                         cudaStream_t stream) {
    if (n_rows <= 0) return SB_OK;
    SB_REQUIRE(sb_conv1_long_supported(taps, pool, cout), "sb_conv1_long_launch: taps %d / pool %d / cout %d unsupported",
               taps, pool, cout);
    SB_REQUIRE(out_ld % 8 == 0, "sb_conv1_long_launch: out_ld %d is not a multiple of 8", out_ld);
    Conv1LParams p;
    p.codes = codes; p.src = src; p.sub_pos = sub_pos; p.sub_code = sub_code;
    p.row0 = row0; p.n_rows = n_rows;
    p.L = L; p.taps = taps; p.pool = pool; p.t_pool = t_pool; p.out_pitch = out_pitch > 0 ? out_pitch : t_pool;
    p.cout = cout; p.out_ld = out_ld; p.relu = relu;
    p.n_pass = (cout + kPassN - 1) / kPassN;
    p.kpad = (taps * 4 + 15) / 16 * 16;
    p.sbo = p.kpad / 8 * 128;
    p.win_per_quad = 32 / pool;
    p.tiles_per_seq = (t_pool + 4 * p.win_per_quad - 1) / (4 * p.win_per_quad);
    p.wpacked = reinterpret_cast<const uint4*>(wpacked);
    p.bias = bias_padded;
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    size_t smem = conv1l_smem(taps, p.n_pass);
    if (smem < 120 * 1024) smem = 120 * 1024;  // one CTA per SM
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
