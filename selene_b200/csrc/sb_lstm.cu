// selene_b200 — DanQ's BiLSTM recurrence as ONE persistent kernel per chunk (bf16 / tcgen05 path).
//
// Reference: models/danQ.py:29-31,42-52 — a bidirectional LSTM(320 -> 320) that, because of the
// batch_first quirk, recurs over the rows of a batch independently for every position.
// A (chain, position) pair is therefore an independent "unit" with `steps` sequential LSTM cell updates.
//
// One CTA owns 128 units of one direction for ALL steps; nothing but the streamed recurrent weights and
// the per-step inputs/outputs touches global memory between steps:
//   * h_{s-1} of the 128 units lives in shared memory in the K-major 128-byte-swizzled layout tcgen05.mma
//     reads as its A operand (double buffered: step s reads one copy while its epilogue writes h_s into
//     the other);
//   * warp 0 (one lane) streams W_hh through a 3-stage TMA ring in tiles of 128 gate rows (4 gates x 32
//     hidden units) x 64 k, identical every step (L2-resident, 800 KB per step per CTA);
//   * warp 1 (one lane) issues the MMAs: per step 10 N-tiles x 5 k-blocks x 4 x (128 x 128 x 16), fp32
//     accumulators in TMEM (2 x 128 columns so the cell math of tile i overlaps the MMAs of tile i+1);
//   * warps 2-5 (thread = unit) do the cell update straight from TMEM: gates = acc + xW (precomputed input
//     projection, fp32), c' = f*c + i*g, h' = o*tanh(c'); c stays in global fp32 (L2), h' goes to the other
//     smem copy (swizzled by hand) and to the layer output.
#include "sb_common.cuh"

#include <cuda.h>

namespace {

constexpr int kUnits = 128;
constexpr int kNTile = 128;          // 4 gates x 32 hidden units
constexpr int kUnitsPerTile = 32;
constexpr int kBStages = 3;
constexpr int kBStageBytes = kNTile * 64 * 2;   // 16384
constexpr int kLThreads = 192;

struct LstmParams {
    long long m_rows;        // rows of the chunk
    long long n_units;       // chains * T
    int H, T, G, I, steps;
    int kblocks;             // H / 64
    int n_tiles;             // H / 32
    const float* xw;         // (m_rows*T, 8H) fp32
    float* cstate;           // (2, n_units, H) fp32
    __nv_bfloat16* out;      // (m_rows*T, out_ld)
    int out_ld;
};

__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__global__ void __launch_bounds__(kLThreads, 1)
lstm_persistent_kernel(const __grid_constant__ CUtensorMap tmap_w0, const __grid_constant__ CUtensorMap tmap_w1,
                       const LstmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (sb_smem_u32(smem_raw) & 1023u)) & 1023u);
    const int a_bytes = p.kblocks * kUnits * 128;          // one copy of h: kblocks x (128 rows x 128 B)
    uint8_t* hA = smem;                                     // [2][a_bytes]
    uint8_t* bring = smem + 2 * a_bytes;                    // [kBStages][16 KB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(bring + kBStages * kBStageBytes);
    uint64_t* b_full = bars;                 // [3]
    uint64_t* b_empty = bars + 3;            // [3]
    uint64_t* tfull = bars + 6;              // [2]
    uint64_t* tempty = bars + 8;             // [2]
    uint64_t* h_ready = bars + 10;           // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    const CUtensorMap* tmap_w = dir == 0 ? &tmap_w0 : &tmap_w1;

    if (warp == 0 && lane == 0) sb_prefetch_tmap(tmap_w);
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kBStages; ++i) { sb_mbar_init(&b_full[i], 1); sb_mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { sb_mbar_init(&tfull[i], 1); sb_mbar_init(&tempty[i], 128); sb_mbar_init(&h_ready[i], 128); }
        sb_fence_mbar_init();
    }
    if (warp == 2) sb_tmem_alloc(tmem_slot, 256);
    // h_{-1} = 0
    for (int i = threadIdx.x; i < a_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(hA)[i] = make_uint4(0, 0, 0, 0);
    sb_fence_proxy_async();
    sb_tc_fence_before();
    __syncthreads();
    sb_tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== W_hh producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int s = 0; s < p.steps; ++s)
                for (int nt = 0; nt < p.n_tiles; ++nt)
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        sb_mbar_wait(&b_empty[stage], phase ^ 1);
                        uint8_t* dst = bring + stage * kBStageBytes;
                        sb_mbar_expect_tx(&b_full[stage], kBStageBytes);
#pragma unroll
                        for (int g = 0; g < 4; ++g)   // gate g, hidden units [32 nt, 32 nt + 32)
                            sb_tma_load_2d(dst + g * 4096, tmap_w, &b_full[stage], kb * 64, g * p.H + nt * kUnitsPerTile);
                        if (++stage == kBStages) { stage = 0; phase ^= 1; }
                    }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = sb_umma_idesc_bf16(kUnits, kNTile);
            uint32_t stage = 0, phase = 0, gt = 0;
            for (int s = 0; s < p.steps; ++s) {
                if (s > 0) {   // h_{s-1} has been written by the epilogue of step s-1
                    sb_mbar_wait(&h_ready[s & 1], ((s - 1) >> 1) & 1);
                    sb_tc_fence_after();
                }
                const uint32_t a_addr = sb_smem_u32(hA + (s & 1) * a_bytes);
                for (int nt = 0; nt < p.n_tiles; ++nt, ++gt) {
                    const uint32_t as = gt & 1;
                    sb_mbar_wait(&tempty[as], ((gt >> 1) & 1) ^ 1);
                    sb_tc_fence_after();
                    const uint32_t tmem_d = tmem_base + as * kNTile;
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        sb_mbar_wait(&b_full[stage], phase);
                        sb_tc_fence_after();
                        const uint64_t da = sb_umma_desc_sw128(a_addr + kb * (kUnits * 128));
                        const uint64_t db = sb_umma_desc_sw128(sb_smem_u32(bring + stage * kBStageBytes));
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            sb_umma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                        sb_umma_commit(&b_empty[stage]);
                        if (++stage == kBStages) { stage = 0; phase ^= 1; }
                    }
                    sb_umma_commit(&tfull[as]);
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== cell update (warps 2..5, thread = unit) =====================
        const int quad = warp & 3;
        const int r = quad * 32 + lane;                       // row of the tile = unit
        const long long unit = (long long)blockIdx.x * kUnits + r;
        // unit -> (chain, position) -> rows of the chunk
        long long first = 0;
        int len = 0, pos = 0;
        if (unit < p.n_units) {
            pos = (int)(unit % p.T);
            const long long chain = unit / p.T;
            const long long block = chain / p.I;
            const int c = (int)(chain - block * p.I);
            first = block * p.G * p.I + c;
            if (first < p.m_rows) {
                long long l = (p.m_rows - first + p.I - 1) / p.I;
                len = (int)(l > p.G ? p.G : l);
            }
        }
        float* cst = p.cstate + ((long long)dir * p.n_units + (unit < p.n_units ? unit : 0)) * p.H;
        uint32_t gt = 0;
        for (int s = 0; s < p.steps; ++s) {
            const bool live = s < len;
            const long long row = first + (long long)p.I * (dir == 0 ? s : len - 1 - s);
            const float* xw = p.xw + ((live ? row : 0) * p.T + pos) * (long long)(8 * p.H) + (long long)dir * 4 * p.H;
            __nv_bfloat16* orow = p.out + ((live ? row : 0) * p.T + pos) * (long long)p.out_ld + (long long)dir * p.H;
            uint8_t* hnext = hA + ((s + 1) & 1) * a_bytes;
            for (int nt = 0; nt < p.n_tiles; ++nt, ++gt) {
                const uint32_t as = gt & 1;
                sb_mbar_wait(&tfull[as], (gt >> 1) & 1);
                sb_tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * kNTile;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t ri[16], rf[16], rg[16], ro[16];
                    sb_tmem_ld16(taddr + 0 * 32 + half * 16, ri);
                    sb_tmem_ld16(taddr + 1 * 32 + half * 16, rf);
                    sb_tmem_ld16(taddr + 2 * 32 + half * 16, rg);
                    sb_tmem_ld16(taddr + 3 * 32 + half * 16, ro);
                    sb_tmem_wait_ld();
                    const int k0 = nt * kUnitsPerTile + half * 16;   // first hidden unit of this group
                    uint32_t hp[8];
                    if (live) {
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const float4 xi = *reinterpret_cast<const float4*>(xw + 0 * p.H + k0 + 4 * j4);
                            const float4 xf = *reinterpret_cast<const float4*>(xw + 1 * p.H + k0 + 4 * j4);
                            const float4 xg = *reinterpret_cast<const float4*>(xw + 2 * p.H + k0 + 4 * j4);
                            const float4 xo = *reinterpret_cast<const float4*>(xw + 3 * p.H + k0 + 4 * j4);
                            float4 cc = s > 0 ? *reinterpret_cast<const float4*>(cst + k0 + 4 * j4) : make_float4(0.f, 0.f, 0.f, 0.f);
                            const float xiv[4] = {xi.x, xi.y, xi.z, xi.w}, xfv[4] = {xf.x, xf.y, xf.z, xf.w};
                            const float xgv[4] = {xg.x, xg.y, xg.z, xg.w}, xov[4] = {xo.x, xo.y, xo.z, xo.w};
                            float cv[4] = {cc.x, cc.y, cc.z, cc.w};
                            float hv[4];
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const int j = 4 * j4 + jj;
                                const float ig = sigmoid_fast(__uint_as_float(ri[j]) + xiv[jj]);
                                const float fg = sigmoid_fast(__uint_as_float(rf[j]) + xfv[jj]);
                                const float gg = tanh_fast(__uint_as_float(rg[j]) + xgv[jj]);
                                const float og = sigmoid_fast(__uint_as_float(ro[j]) + xov[jj]);
                                cv[jj] = fg * cv[jj] + ig * gg;
                                hv[jj] = og * tanh_fast(cv[jj]);
                            }
                            *reinterpret_cast<float4*>(cst + k0 + 4 * j4) = make_float4(cv[0], cv[1], cv[2], cv[3]);
                            __nv_bfloat162 a = __floats2bfloat162_rn(hv[0], hv[1]), b = __floats2bfloat162_rn(hv[2], hv[3]);
                            hp[2 * j4] = *reinterpret_cast<uint32_t*>(&a);
                            hp[2 * j4 + 1] = *reinterpret_cast<uint32_t*>(&b);
                        }
                        // layer output (bf16, 32 contiguous bytes)
                        *reinterpret_cast<uint4*>(orow + k0) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                        *reinterpret_cast<uint4*>(orow + k0 + 8) = make_uint4(hp[4], hp[5], hp[6], hp[7]);
                        // h_s into the other A copy: k-block k0/64, 128-byte rows, 16-byte chunks XOR-swizzled by row&7
                        uint8_t* rowp = hnext + (k0 >> 6) * (kUnits * 128) + r * 128;
                        const int ch0 = ((k0 & 63) * 2) >> 4;
                        *reinterpret_cast<uint4*>(rowp + (((ch0) ^ (r & 7)) << 4)) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                        *reinterpret_cast<uint4*>(rowp + (((ch0 + 1) ^ (r & 7)) << 4)) = make_uint4(hp[4], hp[5], hp[6], hp[7]);
                    }
                }
                sb_tc_fence_before();
                sb_mbar_arrive(&tempty[as]);
            }
            sb_fence_proxy_async();           // this step's h is visible to the tensor core
            sb_mbar_arrive(&h_ready[(s + 1) & 1]);
        }
    }
    sb_tc_fence_before();
    __syncthreads();
    if (warp == 2) { sb_tc_fence_after(); sb_tmem_dealloc(tmem_base, 256); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

int sb_lstm_persistent_supported(int H) { return H % 64 == 0 && H >= 64 && H <= 512; }

// whh[d]: (4H x h_ld) bf16 K-major; xw: (m*T, 8H) fp32; cstate: (2, n_units, H) fp32 scratch; out: (m*T, out_ld) bf16
int sb_lstm_persistent_launch(const void* whh0, const void* whh1, int h_ld, const float* xw, float* cstate, void* out,
                              int out_ld, long long m_rows, int H, int T, int G, int I, int device, cudaStream_t stream) {
    SB_REQUIRE(sb_lstm_persistent_supported(H), "sb_lstm_persistent_launch: hidden size %d unsupported", H);
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        SB_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q));
        SB_REQUIRE(q == cudaDriverEntryPointSuccess && ptr, "cuTensorMapEncodeTiled is not available from the driver");
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    CUtensorMap tm[2];
    const void* w[2] = {whh0, whh1};
    for (int d = 0; d < 2; ++d) {
        cuuint64_t dims[2] = {(cuuint64_t)H, (cuuint64_t)(4 * H)};
        cuuint64_t strides[1] = {(cuuint64_t)h_ld * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)kUnitsPerTile};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = fn(&tm[d], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w[d]), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(W_hh) failed with CUresult %d", (int)r);
    }
    LstmParams p;
    const long long blk = (long long)G * I;
    const long long chains = ((m_rows + blk - 1) / blk) * I;
    p.m_rows = m_rows; p.n_units = chains * T;
    p.H = H; p.T = T; p.G = G; p.I = I;
    long long steps = (m_rows + I - 1) / I;
    p.steps = (int)(steps > G ? G : steps);
    p.kblocks = H / 64; p.n_tiles = H / kUnitsPerTile;
    p.xw = xw; p.cstate = cstate; p.out = reinterpret_cast<__nv_bfloat16*>(out); p.out_ld = out_ld;
    const size_t smem = 1024 + 2 * (size_t)p.kblocks * kUnits * 128 + kBStages * kBStageBytes + 256;
    SB_REQUIRE(smem <= 227 * 1024, "sb_lstm_persistent_launch: %zu bytes of shared memory needed", smem);
    static size_t attr = 0;
    if (smem > attr) {
        SB_CHECK_CUDA(cudaFuncSetAttribute(lstm_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    dim3 grid((unsigned)((p.n_units + kUnits - 1) / kUnits), 2);
    lstm_persistent_kernel<<<grid, kLThreads, smem, stream>>>(tm[0], tm[1], p);
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    (void)device;
    return SB_OK;
}
