// selene_b200 — first convolution (4 -> cout channels, <= 8 taps) straight from base codes on tcgen05.
//
// Replaces torch.conv1d + relu + max_pool1d of models/deepsea.py:22-25 for one-hot input.  The input is
// never materialised as floats: each CTA builds the im2col rows of its tile in shared memory directly
// from the uint8 codes (one 64-bit pattern per tap: 1.0 in the slot of the base, or 0.25 x 4 for an
// unknown base), in the un-swizzled K-major core-matrix layout tcgen05.mma reads.  K = taps*4 <= 32, so
// the tensor work is tiny; the kernel is bounded by its epilogue.  To keep the epilogue free of
// cross-lane traffic the max-pool is folded into the MMA schedule: row i of accumulator q holds conv
// position pool*i + q, so the `pool` members of a pooling window sit in the SAME TMEM lane of `pool`
// different column blocks and the pooled value is a per-thread max.
//
//   tile            = 128 pooled positions of one sequence (DeepSEA: 2 tiles per 1000-bp sequence)
//   accumulators    = 2 stages x pool x 64 columns of TMEM (<= 512)
//   warps 0-7       : build A for the NEXT tile (two warps per 32 rows, each half of the q blocks), then
//                     drain the current tile's accumulators, the two warps of a TMEM lane quadrant taking
//                     alternate 16-column chunks (tcgen05.ld -> max over q -> +bias -> ReLU -> bf16 ->
//                     2 x 16-byte stores per chunk); eight warps because one epilogue warp per scheduler
//                     leaves the issue slots 3/4 idle (ncu, r01)
//   warp 8 (1 lane) : issues tcgen05.mma (pool x K/16 per 64-channel pass), commits to mbarriers
#include "sb_common.cuh"

namespace {

constexpr int kRows = 128;          // pooled positions per tile
constexpr int kPassN = 64;          // channels per pass
constexpr int kKPad = 32;           // padded K (8 taps x 4 bases)
constexpr int kRowBytes = kKPad * 2;                 // 64
constexpr int kATileBytes = kRows * kRowBytes;       // 8192 per q
constexpr int kMaxPool = 4;
constexpr int kMaxCodes = kRows * kMaxPool + 16;     // codes staged per tile
constexpr int kStagePitch = kPassN * 2 + 16;         // bytes per staged output row (+16: conflict-free)

struct Conv1Params {
    const uint8_t* codes;
    const int32_t* src;
    const int32_t* sub_pos;
    const uint8_t* sub_code;
    long long row0;
    long long n_rows;
    int L, taps, pool, t_pool, out_pitch, cout, out_ld, relu, n_pass, tiles_per_seq;
    const uint4* wpacked;   // n_pass*64 rows x 32 K in the smem core-matrix layout
    const float* bias;      // n_pass*64
    __nv_bfloat16* out;
};

// un-swizzled K-major layout: 8-row x 16-byte core matrices; K-adjacent core matrices 128 B apart
// (LBO), M/N-adjacent 8-row groups 512 B apart (SBO)
__device__ __forceinline__ uint64_t desc_nosw(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(128 >> 4) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__device__ __forceinline__ uint64_t tap_pattern(int code) {
    return code < 4 ? (0x3F80ull << (16 * code)) : 0x3E803E803E803E80ull;
}

__device__ __forceinline__ float max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));   // FMNMX3
    return r;
}

// 16 channels of one pooled row: pooled = max over the POOL accumulator blocks, then bias + ReLU as
// max(m, -bias) + bias (one FMNMX3 absorbs the last pool member and the ReLU), bf16, two 16-byte stores
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

template <int POOL>
__device__ __forceinline__ void finish_chunk(const uint32_t (&r)[4][16], const float* __restrict__ bs, bool relu,
                                             uint8_t* __restrict__ stage_row) {
    float v[16];
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
        const float4 b4 = *reinterpret_cast<const float4*>(bs + 4 * j4);
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = 4 * j4 + jj;
            const float floor_ = relu ? -bb[jj] : -INFINITY;
            float m;
            if (POOL == 1) {
                m = fmaxf(__uint_as_float(r[0][j]), floor_);
            } else if (POOL == 2) {
                m = max3(__uint_as_float(r[0][j]), __uint_as_float(r[1][j]), floor_);
            } else {
                m = max3(__uint_as_float(r[0][j]), __uint_as_float(r[1][j]), __uint_as_float(r[2][j]));
                m = max3(m, __uint_as_float(r[3][j]), floor_);
            }
            v[j] = m + bb[jj];
        }
    }
#pragma unroll
    for (int g = 0; g < 2; ++g)
        *reinterpret_cast<uint4*>(stage_row + 16 * g) =
            make_uint4(pack2(v[8 * g], v[8 * g + 1]), pack2(v[8 * g + 2], v[8 * g + 3]),
                       pack2(v[8 * g + 4], v[8 * g + 5]), pack2(v[8 * g + 6], v[8 * g + 7]));
}

template <int POOL>
__global__ void __launch_bounds__(288, POOL == 4 ? 1 : 2) conv1_tc_kernel(const Conv1Params p) {
    extern __shared__ uint8_t smem_raw[];
    // pointer arithmetic on the __shared__ array (not an integer round trip) keeps the shared address
    // space visible to the compiler: LDS/STS instead of generic LD/ST
    uint8_t* smem = smem_raw + ((1024u - (sb_smem_u32(smem_raw) & 1023u)) & 1023u);
    // layout: [A buf0: 4 x 8 KB][A buf1: 4 x 8 KB][B: n_pass x 4 KB][bias][codes x2][barriers]
    uint8_t* a_buf = smem;
    uint8_t* b_smem = smem + 2 * POOL * kATileBytes;
    float* bias_s = reinterpret_cast<float*>(b_smem + p.n_pass * kPassN * kRowBytes);
    uint8_t* codes_s = reinterpret_cast<uint8_t*>(bias_s + p.n_pass * kPassN);
    uint8_t* stage_s = codes_s + kMaxCodes;                              // [2][128 rows][kStagePitch]
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_s + 2 * kRows * kStagePitch);
    uint64_t* a_full = bars;        // [2]
    uint64_t* tfull = bars + 2;     // [2]
    uint64_t* tempty = bars + 4;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const long long n_tiles = p.n_rows * p.tiles_per_seq;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            sb_mbar_init(&a_full[i], 256);
            sb_mbar_init(&tfull[i], 1);
            sb_mbar_init(&tempty[i], 256);
        }
        sb_fence_mbar_init();
    }
    // 2 accumulator stages x POOL x 64 columns: pool 1 / 2 leave tensor memory (and shared memory) for a second or
    // third CTA on the SM, whose builder / MMA / epilogue phases then interleave with this one's
    constexpr uint32_t kTmem = 2 * POOL * kPassN;
    if (warp == 8) sb_tmem_alloc(tmem_slot, kTmem);
    // weights + bias -> smem (once per CTA)
    {
        const int n16 = p.n_pass * kPassN * kRowBytes / 16;
        uint4* dst = reinterpret_cast<uint4*>(b_smem);
        for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = __ldg(p.wpacked + i);
        for (int i = threadIdx.x; i < p.n_pass * kPassN; i += blockDim.x) bias_s[i] = __ldg(p.bias + i);
    }
    sb_fence_proxy_async();
    sb_tc_fence_before();
    __syncthreads();
    sb_tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int ksteps = (p.taps * 4 + 15) / 16;

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
                    for (int q = 0; q < p.pool; ++q) {
                        const uint32_t d = tmem_base + stage * (POOL * kPassN) + q * kPassN;
                        for (int ks = 0; ks < ksteps; ++ks) {
                            // one K step = 2 core matrices = 256 B along K
                            const uint64_t da = desc_nosw(a_addr + (buf * POOL + q) * kATileBytes + ks * 256);
                            const uint64_t db = desc_nosw(b_addr + pass * (kPassN * kRowBytes) + ks * 256);
                            sb_umma_bf16(d, da, db, idesc, ks > 0 ? 1u : 0u);
                        }
                    }
                    sb_umma_commit(&tfull[stage]);
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== A builder + epilogue (warps 0..7, thread = pooled row x half) ==============
        const int i = threadIdx.x & 127;   // pooled row of the tile
        const int half = threadIdx.x >> 7; // 0: warps 0-3, 1: warps 4-7
        const int n_codes = kRows * p.pool + 8;

        auto build = [&](long long tile, int buf) {
            const long long seq = tile / p.tiles_per_seq;
            const int tin = (int)(tile - seq * p.tiles_per_seq);
            const long long row = p.row0 + seq;
            const long long srow = p.src ? (long long)p.src[row] : row;
            const uint8_t* crow = p.codes + srow * p.L;
            const int sp = p.sub_pos ? p.sub_pos[row] : -1;
            const int sc = (p.sub_pos && p.sub_code) ? p.sub_code[row] : 4;
            const int p0 = tin * kRows * p.pool;
            uint8_t* cs = codes_s;
            for (int k = threadIdx.x; k < n_codes; k += 256) {
                const int pos = p0 + k;
                int c = 4;
                if (pos < p.L) {
                    c = __ldg(crow + pos);
                    if (pos == sp) c = sc;
                    if (c > 4) c = 4;
                }
                cs[k] = (uint8_t)c;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // this thread's pool + 7 codes
            uint64_t pat[kMaxPool + 7];
#pragma unroll
            for (int k = 0; k < kMaxPool + 7; ++k) {
                const int c = (k < p.pool + 7) ? cs[p.pool * i + k] : 4;
                pat[k] = tap_pattern(c);
            }
            uint8_t* a_row = a_buf + (size_t)buf * POOL * kATileBytes + (i >> 3) * 512 + (i & 7) * 16;
#pragma unroll
            for (int q = 0; q < kMaxPool; ++q) {
                // the two halves split the q blocks between them (pool 4: {0,1} / {2,3}; pool 2: {0} / {1})
                const bool mine = p.pool == 1 ? half == 0 : (q * 2 / p.pool) == half;
                if (q < p.pool && mine) {
#pragma unroll
                    for (int kc = 0; kc < 4; ++kc) {
                        const uint64_t lo = (2 * kc < p.taps) ? pat[q + 2 * kc] : 0ull;
                        const uint64_t hi = (2 * kc + 1 < p.taps) ? pat[q + 2 * kc + 1] : 0ull;
                        *reinterpret_cast<uint4*>(a_row + q * kATileBytes + kc * 128) =
                            make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
                    }
                }
            }
            sb_fence_proxy_async();      // generic-proxy smem writes -> visible to the tensor core
            sb_mbar_arrive(&a_full[buf]);
            asm volatile("bar.sync 1, 256;" ::: "memory");  // codes_s may be overwritten by the next build
        };

        uint32_t gpass = 0;
        int it = 0;
        long long tile = blockIdx.x;
        if (tile < n_tiles) build(tile, 0);
        for (; tile < n_tiles; tile += gridDim.x, ++it) {
            const long long next = tile + gridDim.x;
            if (next < n_tiles) build(next, (it + 1) & 1);
            const long long seq = tile / p.tiles_per_seq;
            const int tin = (int)(tile - seq * p.tiles_per_seq);
            for (int pass = 0; pass < p.n_pass; ++pass, ++gpass) {
                const uint32_t stage = gpass & 1;
                sb_mbar_wait(&tfull[stage], (gpass >> 1) & 1);
                sb_tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + stage * (POOL * kPassN);
                // this warp's two 16-column chunks of the pass; the second chunk's TMEM loads are in
                // flight while the first is finished
                const int c0 = half, c1 = half + 2;
                uint32_t ra[4][16], rb[4][16];
#pragma unroll
                for (int q = 0; q < POOL; ++q) sb_tmem_ld16(taddr + q * kPassN + c0 * 16, ra[q]);
                sb_tmem_wait_ld();
#pragma unroll
                for (int q = 0; q < POOL; ++q) sb_tmem_ld16(taddr + q * kPassN + c1 * 16, rb[q]);
                uint8_t* st = stage_s + (gpass & 1) * (kRows * kStagePitch);
                finish_chunk<POOL>(ra, bias_s + pass * kPassN + c0 * 16, p.relu != 0, st + i * kStagePitch + c0 * 32);
                sb_tmem_wait_ld();
                finish_chunk<POOL>(rb, bias_s + pass * kPassN + c1 * 16, p.relu != 0, st + i * kStagePitch + c1 * 32);
                sb_tc_fence_before();
                sb_mbar_arrive(&tempty[stage]);      // accumulators are in registers/smem: release them early
                // staged tile (128 rows x 64 channels) -> global with coalesced stores: a warp instruction
                // writes 4 rows x 128 contiguous bytes instead of 32 rows x 16 bytes
                asm volatile("bar.sync 2, 256;" ::: "memory");
                {
                    const int nb = pass * kPassN;
#pragma unroll
                    for (int k = 0; k < (kRows * 8) / 256; ++k) {
                        const int id = threadIdx.x + 256 * k;
                        const int row = id >> 3, piece = id & 7;
                        const int tpr = tin * kRows + row;
                        if (tpr < p.t_pool && nb + piece * 8 + 8 <= p.out_ld) {
                            const uint4 val = *reinterpret_cast<const uint4*>(st + row * kStagePitch + piece * 16);
                            *reinterpret_cast<uint4*>(p.out + (seq * p.out_pitch + tpr) * (long long)p.out_ld + nb +
                                                      piece * 8) = val;
                        }
                    }
                }
            }
        }
    }

    sb_tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        sb_tc_fence_after();
        sb_tmem_dealloc(tmem_base, 2 * POOL * kPassN);
    }
}

}  // namespace

// Host: pack conv weights (cout, 4, taps) fp32 into the smem core-matrix layout, bf16, K = tap*4 + base.
// Returns the number of bytes written (n_pass * 64 rows * 64 B).
size_t sb_conv1_pack_weights(const float* w, int cout, int taps, const uint8_t perm[4], uint16_t* out,
                             uint16_t (*to_bf16)(float)) {
    const int n_pass = (cout + kPassN - 1) / kPassN;
    const size_t elems = (size_t)n_pass * kPassN * kKPad;
    for (size_t i = 0; i < elems; ++i) out[i] = 0;
    for (int n = 0; n < cout; ++n)
        for (int tap = 0; tap < taps; ++tap)
            for (int b = 0; b < 4; ++b) {
                const int k = tap * 4 + b;
                const size_t off = (size_t)(n >> 3) * 256 + (size_t)(k >> 3) * 64 + (size_t)(n & 7) * 8 + (k & 7);
                out[off] = to_bf16(w[((size_t)n * 4 + perm[b]) * taps + tap]);
            }
    return elems * 2;
}

int sb_conv1_tc_supported(int taps, int pool, int cout) {
    return taps >= 1 && taps <= 8 && (pool == 1 || pool == 2 || pool == 4) && cout >= 1 && cout <= 1024;
}

int sb_conv1_tc_launch(const uint8_t* codes, const int32_t* src, const int32_t* sub_pos, const uint8_t* sub_code,
                       long long row0, long long n_rows, int L, int taps, int pool, int t_pool, int out_pitch, int cout,
                       int out_ld, int relu, const void* wpacked, const float* bias_padded, void* out, int device,
                       cudaStream_t stream) {
    if (n_rows <= 0) return SB_OK;
    Conv1Params p;
    p.codes = codes; p.src = src; p.sub_pos = sub_pos; p.sub_code = sub_code;
    p.row0 = row0; p.n_rows = n_rows;
    p.L = L; p.taps = taps; p.pool = pool; p.t_pool = t_pool; p.out_pitch = out_pitch > 0 ? out_pitch : t_pool; p.cout = cout; p.out_ld = out_ld; p.relu = relu;
    p.n_pass = (cout + kPassN - 1) / kPassN;
    p.tiles_per_seq = (t_pool + kRows - 1) / kRows;
    p.wpacked = reinterpret_cast<const uint4*>(wpacked);
    p.bias = bias_padded;
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    size_t smem = 1024 + 2 * (size_t)pool * kATileBytes + (size_t)p.n_pass * kPassN * (kRowBytes + 4) + kMaxCodes +
                  2 * kRows * kStagePitch + 64 + 64;
    if (pool == 4 && smem < 120 * 1024) smem = 120 * 1024;  // pool 4: one CTA per SM, it owns all 512 TMEM columns
    const int per_sm = pool == 4 ? 1 : 2;     // pool 1 / 2: two CTAs fit (tensor memory 128-256 columns each, 108 registers x 288 threads)
    SB_REQUIRE(smem <= 220 * 1024, "sb_conv1_tc_launch: %zu bytes of shared memory needed", smem);
    static size_t attr = 0;
    if (smem > attr) {
        SB_CHECK_CUDA(cudaFuncSetAttribute(conv1_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SB_CHECK_CUDA(cudaFuncSetAttribute(conv1_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SB_CHECK_CUDA(cudaFuncSetAttribute(conv1_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // two CTAs per SM for pool 1 / 2 need the large shared-memory carve-out even though one CTA would fit a small one
        SB_CHECK_CUDA(cudaFuncSetAttribute(conv1_tc_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                           cudaSharedmemCarveoutMaxShared));
        SB_CHECK_CUDA(cudaFuncSetAttribute(conv1_tc_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                           cudaSharedmemCarveoutMaxShared));
        attr = smem;
    }
    const long long tiles = n_rows * p.tiles_per_seq;
    const int sms = sb_sm_count(device);
    const long long ctas = (long long)sms * per_sm;
    const int grid = (int)(tiles < ctas ? tiles : ctas);
    if (pool == 4) conv1_tc_kernel<4><<<grid, 288, smem, stream>>>(p);
    else if (pool == 2) conv1_tc_kernel<2><<<grid, 288, smem, stream>>>(p);
    else conv1_tc_kernel<1><<<grid, 288, smem, stream>>>(p);
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    return SB_OK;
}
