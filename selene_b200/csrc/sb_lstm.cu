// selene_b200 — DanQ's BiLSTM recurrence as ONE persistent kernel per chunk (bf16 / tcgen05 path).
//
// Reference: models/danQ.py:29-31,42-52 — a bidirectional LSTM(320 -> 320) that, because of the
// batch_first quirk, recurs over the rows of a batch independently for every position.
// A (chain, position) pair is therefore an independent "unit" with `steps` sequential LSTM cell updates.
//
// One CTA owns 128 units of one direction for ALL steps; between steps only the streamed recurrent weights,
// the per-step input projection and the layer output touch global memory:
//   * h_{s-1} of the 128 units lives in shared memory in the K-major 128-byte-swizzled layout tcgen05.mma
//     reads as its A operand (double buffered: step s reads one copy while its epilogue writes h_s into
//     the other);
//   * c of the 128 units lives in TENSOR MEMORY next to the accumulators (lane = unit, column = hidden
//     unit, fp32), read with tcgen05.ld and written back with tcgen05.st by the thread that owns the unit; the
//     256 columns left after the accumulators hold hidden units 0..255, the rest (DanQ: 64) stay in registers;
//   * warp 0 (one lane) streams W_hh through a TMA ring in tiles of 128 gate rows x 64 k.  The rows are
//     permuted on the host so that every 64 rows hold 2 halves x 4 gates x 8 hidden units (sb_lstm_perm_row):
//     one contiguous TMA box per tile, identical every step (L2-resident);
//   * warp 1 (one lane) issues the MMAs: per step H/32 N-tiles x H/64 k-blocks x 4 x (128 x 128 x 16), fp32
//     accumulators in TMEM (2 x 128 columns, so the cell math of a tile overlaps the MMAs of the next);
//   * warps 2-9 (thread = unit x half tile) do the cell update: gates = acc + xW, where xW is the input
//     projection (+ bias) written by the projection GEMM as bf16 in the same permuted column order, so the 64
//     bytes one thread needs per tile are contiguous; every thread loads its own one tile ahead into
//     registers (and prefetches the next step's line into L2 a whole step ahead).
//
// Small chunks (<= SMs/4 unit tiles, e.g. a 3000-mutant ISM) launch CLUSTERS of two CTAs per unit tile: each CTA
// computes the gate columns of half of the hidden units (half of the MMA tiles, half of W_hh, half of the cell
// math per step) and writes its half of h_s into BOTH CTAs' shared memory (st.shared::cluster through mapa);
// the per-step barrier counts the arrivals of both CTAs (mbarrier.arrive.release.cluster on the peer's barrier,
// try_wait.acquire.cluster + fence.proxy.async before the MMAs that read the peer-written rows).
#include "sb_common.cuh"

#include <cuda.h>

#include <type_traits>

namespace {

constexpr int kUnits = 128;
constexpr int kSubTile = 64;         // 2 halves x 4 gates x 8 hidden units (the unit of the row permutation)
constexpr int kNTile = 128;          // one MMA tile = two sub-tiles
constexpr int kHidPerSub = 16;
constexpr int kBStages = 4;          // 64 KB of W_hh in flight: the ring must cover the L2 latency x bandwidth product
constexpr int kBStageBytes = kNTile * 64 * 2;   // 16384
constexpr int kEpiThreads = 256;
constexpr int kLThreads = 64 + kEpiThreads;
constexpr int kAccStages = 2;
constexpr int kAccCols = kAccStages * kNTile;   // 256: c state starts at this TMEM column
constexpr int kCTmemHidden = 512 - kAccCols;    // hidden units whose c lives in tensor memory; the rest in registers
constexpr int kCRegSubTiles = 4;                // up to 4 sub-tiles (64 hidden units) of c in registers

struct LstmParams {
    long long m_rows;        // rows of the chunk
    long long n_units;       // chains * T
    int H, T, G, I, steps;
    int kblocks;             // H / 64
    int n_sub;               // H / 16 sub-tiles
    int n_tiles;             // n_sub / 2 MMA tiles
    int csize;               // CTAs per unit tile (cluster size 1 or 2): each takes n_tiles / csize tiles of every step
    const __nv_bfloat16* xw; // (m_rows*T, 8H) bf16, permuted columns, bias included
    __nv_bfloat16* out;      // (m_rows*T, out_ld)
    int out_ld;
};

__device__ __forceinline__ float tanh_fast(float x) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// one MUFU instead of ex2 + rcp
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

__global__ void __launch_bounds__(kLThreads, 1)
lstm_persistent_kernel(const __grid_constant__ CUtensorMap tmap_w0, const __grid_constant__ CUtensorMap tmap_w1,
                       const LstmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (sb_smem_u32(smem_raw) & 1023u)) & 1023u);
    const int a_bytes = p.kblocks * kUnits * 128;          // one copy of h: kblocks x (128 rows x 128 B)
    uint8_t* hA = smem;                                     // [2][a_bytes]
    uint8_t* bring = smem + 2 * a_bytes;                    // [kBStages][8 KB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(bring + kBStages * kBStageBytes);
    uint64_t* b_full = bars;                    // [kBStages]
    uint64_t* b_empty = bars + kBStages;        // [kBStages]
    uint64_t* tfull = bars + 2 * kBStages;      // [kAccStages]
    uint64_t* tempty = tfull + kAccStages;      // [kAccStages]
    uint64_t* h_ready = tempty + kAccStages;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_ready + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    const CUtensorMap* tmap_w = dir == 0 ? &tmap_w0 : &tmap_w1;
    // cluster of `csize` CTAs per unit tile: CTA `crank` computes the gate columns of hidden units
    // [crank, crank + 1) * H / csize and writes its part of h_s into EVERY CTA's shared memory (DSMEM)
    const uint32_t crank = p.csize > 1 ? sb_cluster_ctarank() : 0u;
    const int tiles_local = p.n_tiles / p.csize;
    const int tile_off = (int)crank * tiles_local;
    const int sub_local = 2 * tiles_local, sub_off = 2 * tile_off;

    if (warp == 0 && lane == 0) sb_prefetch_tmap(tmap_w);
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kBStages; ++i) { sb_mbar_init(&b_full[i], 1); sb_mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < kAccStages; ++i) { sb_mbar_init(&tfull[i], 1); sb_mbar_init(&tempty[i], kEpiThreads); }
        for (int i = 0; i < 2; ++i) sb_mbar_init(&h_ready[i], kEpiThreads * p.csize);
        sb_fence_mbar_init();
    }
    if (warp == 2) sb_tmem_alloc(tmem_slot, 512);
    // h_{-1} = 0
    for (int i = threadIdx.x; i < a_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(hA)[i] = make_uint4(0, 0, 0, 0);
    sb_fence_proxy_async();
    sb_tc_fence_before();
    __syncthreads();
    if (p.csize > 1) sb_cluster_sync();   // the peer's barriers and h buffers exist before anything remote touches them
    sb_tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== W_hh producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int s = 0; s < p.steps; ++s)
                for (int nt = 0; nt < tiles_local; ++nt)
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        sb_mbar_wait(&b_empty[stage], phase ^ 1);
                        sb_mbar_expect_tx(&b_full[stage], kBStageBytes);
                        sb_tma_load_2d(bring + stage * kBStageBytes, tmap_w, &b_full[stage], kb * 64, (tile_off + nt) * kNTile);
                        if (++stage == kBStages) { stage = 0; phase ^= 1; }
                    }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = sb_umma_idesc_bf16(kUnits, kNTile);
            uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
            for (int s = 0; s < p.steps; ++s) {
                if (s > 0) {   // h_{s-1} has been written by the epilogue of step s-1 (of every CTA of the cluster)
                    if (p.csize > 1) {
                        sb_mbar_wait_cluster(&h_ready[s & 1], ((s - 1) >> 1) & 1);
                        sb_fence_proxy_async();   // the peer's generic-proxy writes -> this CTA's tensor core reads
                    } else {
                        sb_mbar_wait(&h_ready[s & 1], ((s - 1) >> 1) & 1);
                    }
                    sb_tc_fence_after();
                }
                const uint32_t a_addr = sb_smem_u32(hA + (s & 1) * a_bytes);
                for (int nt = 0; nt < tiles_local; ++nt) {
                    sb_mbar_wait(&tempty[as], aphase ^ 1);
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
                    if (++as == kAccStages) { as = 0; aphase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== cell update (warps 2..9; thread = (unit, half of the tile)) =====================
        const int quad = warp & 3;                            // TMEM lane quadrant this warp may touch
        const int half = (warp - 2) >> 2;
        const int r = quad * 32 + lane;                       // row of the tile = unit
        const long long unit = (long long)(blockIdx.x / p.csize) * kUnits + r;
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
        const uint32_t lane_bits = (uint32_t)(quad * 32) << 16;
        const long long xw_ld = 8LL * p.H;
        // element offset of this thread's 32 projection values of tile nt at step s: row(s) * xw_ld + col0 + 64 nt
        const long long col0 = (long long)dir * 4 * p.H + half * 32;
        auto row_of = [&](int s) { return (first + (long long)p.I * (dir == 0 ? s : len - 1 - s)) * p.T + pos; };
        // every thread loads its OWN 64 bytes of the projection one sub-tile ahead, straight into registers (the
        // line was pulled into L2 during the previous step)
        auto fetch = [&](int s, int sub_tile, uint4 (&x)[4]) {
            if (s < len) {
                const uint4* src = reinterpret_cast<const uint4*>(p.xw + row_of(s) * xw_ld + col0 + sub_tile * kSubTile);
#pragma unroll
                for (int g = 0; g < 4; ++g) x[g] = __ldg(src + g);
            }
        };
        uint4 xn[4] = {};
        int fs = 0, fst = 0;   // next (local) sub-tile to fetch
        fetch(0, sub_off, xn);
        if (++fst == sub_local) { fst = 0; ++fs; }
        float creg[kCRegSubTiles * 8];   // c of the hidden units beyond kCTmemHidden (DanQ: 256..319)
#pragma unroll
        for (int i = 0; i < kCRegSubTiles * 8; ++i) creg[i] = 0.f;
        const int h_local = p.H / p.csize;                                               // hidden units of this CTA
        const int n_tmem_sub = (h_local < kCTmemHidden ? h_local : kCTmemHidden) / kHidPerSub;   // (local) sub-tiles with c in TMEM

        uint32_t as = 0, aphase = 0;
        for (int s = 0; s < p.steps; ++s) {
            const bool live = s < len;
            const long long grow = live ? row_of(s) : 0;
            __nv_bfloat16* orow = p.out + grow * (long long)p.out_ld + (long long)dir * p.H + half * 8;
            // the next step's projection row: pull it towards L2 while this step computes
            const __nv_bfloat16* nxt = (s + 1 < len) ? p.xw + row_of(s + 1) * xw_ld + col0 : nullptr;
            uint8_t* hnext = hA + ((s + 1) & 1) * a_bytes;

            // one 64-column sub-tile (16 hidden units, this thread: 8 of them); RS >= 0: c in creg[RS*8 ...]
            auto sub_tile = [&](int st_local, uint32_t acc_col, bool release, auto rs_tag) {
                constexpr int RS = decltype(rs_tag)::value;
                const int st = sub_off + st_local;       // sub-tile of the whole hidden state
                if (nxt) sb_prefetch_l2(nxt + st * kSubTile);
                uint4 xv[4];   // gates i, f, g, o: 8 bf16 each
#pragma unroll
                for (int g = 0; g < 4; ++g) xv[g] = xn[g];
                if (fs < p.steps) fetch(fs, sub_off + fst, xn);   // the next sub-tile
                if (++fst == sub_local) { fst = 0; ++fs; }
                const uint32_t taddr = tmem_base + lane_bits + acc_col + half * 32;
                const uint32_t caddr = tmem_base + lane_bits + kAccCols + st_local * kHidPerSub + half * 8;
                uint32_t ri[8], rf[8], rg[8], ro[8], rc[8];
                sb_tmem_ld8(taddr + 0, ri);
                sb_tmem_ld8(taddr + 8, rf);
                sb_tmem_ld8(taddr + 16, rg);
                sb_tmem_ld8(taddr + 24, ro);
                if (RS < 0) { if (s > 0) sb_tmem_ld8(caddr, rc); }
                sb_tmem_wait_ld();
                if (release) {     // the whole 128-column stage is in registers now: the MMAs of tile + 2 may overwrite it
                    sb_tc_fence_before();
                    sb_mbar_arrive(&tempty[as]);
                }
                const uint32_t xi[4] = {xv[0].x, xv[0].y, xv[0].z, xv[0].w};
                const uint32_t xf[4] = {xv[1].x, xv[1].y, xv[1].z, xv[1].w};
                const uint32_t xg[4] = {xv[2].x, xv[2].y, xv[2].z, xv[2].w};
                const uint32_t xo[4] = {xv[3].x, xv[3].y, xv[3].z, xv[3].w};
                uint32_t hp[4];
#pragma unroll
                for (int j2 = 0; j2 < 4; ++j2) {
                    float hv[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int j = 2 * j2 + q;
                        const float ai = __uint_as_float(ri[j]) + (q ? bf16_hi(xi[j2]) : bf16_lo(xi[j2]));
                        const float af = __uint_as_float(rf[j]) + (q ? bf16_hi(xf[j2]) : bf16_lo(xf[j2]));
                        const float ag = __uint_as_float(rg[j]) + (q ? bf16_hi(xg[j2]) : bf16_lo(xg[j2]));
                        const float ao = __uint_as_float(ro[j]) + (q ? bf16_hi(xo[j2]) : bf16_lo(xo[j2]));
                        float c_old;
                        if (RS >= 0) c_old = creg[(RS >= 0 ? RS : 0) * 8 + j];
                        else c_old = s > 0 ? __uint_as_float(rc[j]) : 0.f;
                        const float c_new = fmaf(sigmoid_fast(af), c_old, sigmoid_fast(ai) * tanh_fast(ag));
                        if (RS >= 0) creg[(RS >= 0 ? RS : 0) * 8 + j] = c_new;
                        else rc[j] = __float_as_uint(c_new);
                        hv[q] = sigmoid_fast(ao) * tanh_fast(c_new);
                    }
                    __nv_bfloat162 a = __floats2bfloat162_rn(hv[0], hv[1]);
                    hp[j2] = *reinterpret_cast<uint32_t*>(&a);
                }
                if (RS < 0) sb_tmem_st8(caddr, rc);
                const uint4 hq = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                const int k0 = st * kHidPerSub + half * 8;          // first hidden unit of this thread's group
                if (live) *reinterpret_cast<uint4*>(orow + st * kHidPerSub) = hq;   // layer output (bf16, 16 bytes)
                // h_s into the other A copy: k-block k0/64, 128-byte rows, 16-byte chunks XOR-swizzled by row&7
                uint8_t* hdst = hnext + (k0 >> 6) * (kUnits * 128) + r * 128 + ((((k0 & 63) >> 3) ^ (r & 7)) << 4);
                *reinterpret_cast<uint4*>(hdst) = hq;
                if (p.csize > 1) sb_st_cluster_v4(sb_mapa(sb_smem_u32(hdst), crank ^ 1u), hq);   // the peer needs all of h_s
            };
            // one 128-column MMA tile = sub-tiles 2 nt and 2 nt + 1
            auto tile = [&](int nt, auto rs0, auto rs1) {
                sb_mbar_wait(&tfull[as], aphase);
                sb_tc_fence_after();
                sub_tile(2 * nt, as * kNTile, false, rs0);
                sub_tile(2 * nt + 1, as * kNTile + kSubTile, true, rs1);
                sb_tmem_wait_st();
                if (++as == kAccStages) { as = 0; aphase ^= 1; }
            };
            using Tm = std::integral_constant<int, -1>;
            const int n_tmem_tiles = n_tmem_sub / 2 < tiles_local ? n_tmem_sub / 2 : tiles_local;
            for (int nt = 0; nt < n_tmem_tiles; ++nt) tile(nt, Tm{}, Tm{});
            if (tiles_local > n_tmem_tiles)
                tile(n_tmem_tiles, std::integral_constant<int, 0>{}, std::integral_constant<int, 1>{});
            if (tiles_local > n_tmem_tiles + 1)
                tile(n_tmem_tiles + 1, std::integral_constant<int, 2>{}, std::integral_constant<int, 3>{});
            sb_fence_proxy_async();           // this step's h is visible to the tensor core
            sb_mbar_arrive(&h_ready[(s + 1) & 1]);
            if (p.csize > 1) sb_mbar_arrive_cluster(sb_mapa(sb_smem_u32(&h_ready[(s + 1) & 1]), crank ^ 1u));
        }
    }
    sb_tc_fence_before();
    __syncthreads();
    if (p.csize > 1) sb_cluster_sync();   // no CTA exits while its peer may still write into its shared memory
    if (warp == 2) { sb_tc_fence_after(); sb_tmem_dealloc(tmem_base, 512); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

size_t lstm_smem_bytes(int H) {
    return 1024 + 2 * (size_t)(H / 64) * kUnits * 128 + kBStages * kBStageBytes + 256;
}

}  // namespace

int sb_lstm_persistent_supported(int H) {
    return H % 64 == 0 && H >= 64 && lstm_smem_bytes(H) <= 227 * 1024 && H <= kCTmemHidden + kCRegSubTiles * kHidPerSub;
}

// Row order of the permuted recurrent / projection matrices of one direction: tile nt holds, for each half,
// the four gates of 8 consecutive hidden units.  Returns the torch row (gate*H + hidden) stored at row n.
int sb_lstm_perm_row(int n, int H) {
    const int nt = n / kSubTile, rem = n % kSubTile;
    const int half = rem / 32, g = (rem % 32) / 8, j = rem % 8;
    return g * H + nt * kHidPerSub + half * 8 + j;
}

// whh[d]: (4H x h_ld) bf16 K-major, rows permuted by sb_lstm_perm_row; xw: (m*T, 8H) bf16 projection + bias with
// columns in the same permuted order per direction; out: (m*T, out_ld) bf16
int sb_lstm_persistent_launch(const void* whh0, const void* whh1, int h_ld, const void* xw, void* out, int out_ld,
                              long long m_rows, int H, int T, int G, int I, int device, cudaStream_t stream) {
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
        cuuint32_t box[2] = {64, (cuuint32_t)kNTile};
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
    p.kblocks = H / 64; p.n_sub = H / kHidPerSub; p.n_tiles = p.n_sub / 2;
    p.xw = reinterpret_cast<const __nv_bfloat16*>(xw); p.out = reinterpret_cast<__nv_bfloat16*>(out); p.out_ld = out_ld;
    const size_t smem = lstm_smem_bytes(H);
    static size_t attr = 0;
    if (smem > attr) {
        SB_CHECK_CUDA(cudaFuncSetAttribute(lstm_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    // split every unit tile over a cluster of 2 CTAs when the unit tiles leave half of the SMs idle anyway and the
    // tiles divide evenly (SB_LSTM_CLUSTER=0/1 overrides)
    const long long unit_tiles = (p.n_units + kUnits - 1) / kUnits;
    int sms = 0;
    SB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const char* cl = getenv("SB_LSTM_CLUSTER");
    bool cluster = p.n_tiles % 2 == 0 && unit_tiles * 4 <= sms;
    if (cl && cl[0] == '0') cluster = false;
    if (cl && cl[0] == '1') cluster = p.n_tiles % 2 == 0;
    p.csize = cluster ? 2 : 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(unit_tiles * p.csize), 2);
    cfg.blockDim = dim3(kLThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = (unsigned)p.csize;
    attrs[0].val.clusterDim.y = 1;
    attrs[0].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    SB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, lstm_persistent_kernel, tm[0], tm[1], p));
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    (void)device;
    return SB_OK;
}
