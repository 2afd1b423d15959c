// selene_b200 — kernel (c): conv1d / FC layers as a persistent, warp-specialised tcgen05 implicit GEMM.
//
// Replaces what torch dispatches for models/deepsea.py:22-50 (torch.conv1d -> cuDNN, linear ->
// cuBLAS, relu, max_pool1d).  One kernel per layer does conv + bias + ReLU + MaxPool + the layout
// change for the next layer:
//
//   * activations are channel-last bf16, rows = sequence x position.  The im2col row of a valid
//     k-tap convolution is then the contiguous run of k*C elements starting at row r, so the A
//     operand is described to TMA as an overlapping-stride 2-D tensor (row stride C, row length
//     k*C) and never materialised.  Rows whose window straddles two sequences are computed and
//     dropped in the epilogue (7/248 of conv2's rows, 7/60 of conv3's).
//   * weights are bf16, K-major (cout x k*C with K index = tap*C + channel), the B operand.
//   * warp 0 (one lane) is the TMA producer, warp 1 (one lane) issues tcgen05.mma with fp32
//     accumulators in TMEM (2 x 256 columns, double buffered), warps 2-5 are the epilogue:
//     tcgen05.ld -> +bias -> max over the 4 rows of a pool window (two xor-shuffles: a pool window is
//     4 adjacent TMEM lanes because tiles start at multiples of 4 and t_in % 4 == 0) -> ReLU -> bf16
//     -> compacted store (sequence, pooled position, channel).
//   * 4-stage smem ring (A 16 KB + B up to 32 KB per stage, 128-byte swizzle), mbarrier full/empty
//     pairs, tcgen05.commit releases stages and publishes accumulators.
//   * persistent grid = min(tiles, SMs); tiles are numbered n-fastest so CTAs running side by side
//     share the A tile in L2 while the (small) weight matrix stays L2-resident.
#include "sb_tc.cuh"

#include <cuda.h>  // CUtensorMap + enums only; the encoder is fetched through the runtime (no -lcuda)

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;   // bf16 elements = one 128-byte swizzle row
constexpr int kStages = 4;
constexpr int kMaxN = 256;
constexpr int kABytes = kBlockM * kBlockK * 2;       // 16384
constexpr int kBBytesMax = kMaxN * kBlockK * 2;      // 32768
constexpr int kStageBytes = kABytes + kBBytesMax;    // 49152
constexpr int kRingBytes = kStages * kStageBytes;    // 196608: tc_layer_kernel cuts it into stages of 16 KB + its B tile
constexpr int kMaxStages = 8;                        // small-N layers: more, smaller stages (the ring must cover the TMA latency)
constexpr int kAuxBytes = 4096;                      // barriers + tmem slot + 2 bias tiles
constexpr int kSmemBytes = kStages * kStageBytes + kAuxBytes + 1024;  // + alignment slack
constexpr int kThreads = 192;
constexpr int kTmemCols = 512;

struct TcParams {
    long long rows_a;
    int t_in, t_conv, pool, t_pool, out_pitch;
    int out_ld, relu, out_f32;
    int block_n, n_tiles_n, n_tiles_m, num_k_blocks;
    int chunks_per_tap;  // 0: overlapping-stride A view; >0: per-tap boxes
    int im2col;          // per-tap boxes through an im2col tensor map: tiles enumerate valid window starts only
    int v_pix;           // im2col: valid window starts per sequence (= pool * t_pool); the epilogue's rows per sequence
    int k_splits, kb_per_split;   // split-K (out_f32 == 3): tile = (range, m, n)
    int last_k16;        // K = 16 slices multiplied in a tap's last 64-channel chunk (4 unless SB_TC_SKIPK=1, see the MMA loop)
    int stages, stage_bytes;   // smem ring: stage = A tile (16 KB) + B tile (block_n x 128 B)
    // narrow layers (block_n <= 128) run TWO CTAs per SM: half the ring, 256 TMEM columns (2 x 128) each.  One
    // issuing thread needs ~600 cycles per k-block (wait, descriptors, 4 MMAs, commit) but a 128 x 64 x 64 k-block
    // is only ~190 tensor-pipe cycles: the second CTA's issuer fills the gap
    int ring_bytes, tmem_cols, acc_stride;
    const float* bias;
    void* out;
    SbScoreEpilogue sc;
};

// The fused epilogue belongs to the bf16 path (gate 2e-2 on sigmoid outputs): exp / log / divide are the hardware
// approximations (ex2 / lg2 / rcp, relative error ~1e-7), a third of the instructions of the library routines that the
// fp32 path's score_kernel keeps (the epilogue of fc2 is instruction-bound: 4 warps finish 128 x 240 values per tile)
__device__ __forceinline__ float tc_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
// logit_score_handler.py:118-124: clamp, then scipy.special.logit in float32
__device__ __forceinline__ float tc_clamp(float p) { return p == 0.f ? 1e-24f : (p >= 1.f ? 0.999999f : p); }
__device__ __forceinline__ float tc_logit(float p) { return __logf(__fdividef(p, 1.f - p)); }

// one (ref, alt) probability pair -> the requested score arrays at flat index `idx`
__device__ __forceinline__ void tc_store_scores(const SbScoreEpilogue& sc, long long idx, float ref, float alt,
                                                bool write_ref) {
    if (sc.abs_diff) sc.abs_diff[idx] = fabsf(ref - alt);   // np.abs(baseline - batch)
    if (sc.diff) sc.diff[idx] = alt - ref;                  // batch - baseline
    if (sc.logit) {
        ref = tc_clamp(ref);
        alt = tc_clamp(alt);
        sc.logit[idx] = tc_logit(alt) - tc_logit(ref);
    }
    if (write_ref && sc.pred_ref) sc.pred_ref[idx] = ref;   // clamped when logits were requested, as the
    if (sc.pred_alt) sc.pred_alt[idx] = alt;                // reference's in-place clamp makes later handlers see
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// PAIR: two CTAs of a cluster (one TPC) work on one 256-row M tile with cta_group::2 MMAs: each CTA loads its own 128
// rows of A and HALF of the B tile, the leader (cluster rank 0) issues the MMAs for both, every CTA drains its own 128
// accumulator lanes.  Per k-block a CTA then pulls 16 KB + N x 64 B instead of 16 KB + N x 128 B from L2.
template <bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
tc_layer_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (sb_smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared pointer
    uint8_t* aux = smem + p.ring_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tfull_bar = empty_bar + kMaxStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* bias_s = reinterpret_cast<float*>(aux + 1024);  // [2][256]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tiles_mn = p.n_tiles_m * p.n_tiles_n;      // PAIR: n_tiles_m counts 256-row tiles
    const int num_tiles = tiles_mn * p.k_splits;
    const uint32_t rank = PAIR ? sb_cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int m_rows_tile = PAIR ? 2 * kBlockM : kBlockM;
    const int b_rows = PAIR ? p.block_n / 2 : p.block_n;       // B rows this CTA loads

    if (warp == 0 && lane == 0) {
        sb_prefetch_tmap(&tmap_a);
        sb_prefetch_tmap(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.stages; ++i) {
            sb_mbar_init(&full_bar[i], 1);
            sb_mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            sb_mbar_init(&tfull_bar[i], 1);
            sb_mbar_init(&tempty_bar[i], PAIR ? 256 : 128);    // PAIR: the epilogue threads of both CTAs arrive at the leader's
        }
        sb_fence_mbar_init();
    }
    if (warp == 2) {
        if (PAIR) sb_tmem_alloc_pair(tmem_slot, (uint32_t)p.tmem_cols);
        else sb_tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    }
    sb_tc_fence_before();
    if (PAIR) sb_cluster_sync();
    else __syncthreads();
    sb_tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const uint32_t tx_bytes = (kABytes + (uint32_t)b_rows * kBlockK * 2) * (PAIR ? 2u : 1u);   // PAIR: both CTAs' bytes
            for (int tile = tile0; tile < num_tiles; tile += tile_step) {
                const int mn = tile % tiles_mn;
                const int m0 = (mn / p.n_tiles_n) * m_rows_tile + (int)rank * kBlockM;
                const int n0 = (mn % p.n_tiles_n) * p.block_n + (int)rank * b_rows;
                // PAIR: the second half of the last 256-row tile may lie wholly past the rows; it then loads tile 0 (its
                // accumulators are never stored)
                const int m0_real = m0;
                const int m0_ld = (PAIR && (long long)m0_real >= p.rows_a) ? 0 : m0_real;
                const int kb0 = (tile / tiles_mn) * p.kb_per_split;
                const int kb1 = kb0 + p.kb_per_split < p.num_k_blocks ? kb0 + p.kb_per_split : p.num_k_blocks;
                for (int kb = kb0; kb < kb1; ++kb) {
                    sb_mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * p.stage_bytes;
                    uint8_t* sb = sa + kABytes;
                    // PAIR: every load of either CTA completes on the LEADER's barrier, which expects both CTAs' bytes
                    if (leader) sb_mbar_expect_tx(&full_bar[stage], tx_bytes);
                    if (p.im2col == 2) {
                        // im2col over the RUN view: a window start's "channels" are its whole taps x C run (the next
                        // positions' rows follow contiguously), so a k-block is 64 elements of that run: no per-tap padding
                        const int n = m0_ld / p.v_pix;
                        if (PAIR) sb_tma_load_im2col_3d_pair(sa, &tmap_a, &full_bar[stage], kb * kBlockK, m0_ld - n * p.v_pix, n, (uint16_t)0);
                        else sb_tma_load_im2col_3d(sa, &tmap_a, &full_bar[stage], kb * kBlockK, m0_ld - n * p.v_pix, n, (uint16_t)0);
                    } else if (p.im2col) {
                        // 128 window starts from (position w, sequence n) on, read at position + tap
                        const int tap = kb / p.chunks_per_tap;
                        const int n = m0_ld / p.v_pix;
                        if (PAIR) sb_tma_load_im2col_3d_pair(sa, &tmap_a, &full_bar[stage], (kb - tap * p.chunks_per_tap) * kBlockK,
                                                             m0_ld - n * p.v_pix, n, (uint16_t)tap);
                        else sb_tma_load_im2col_3d(sa, &tmap_a, &full_bar[stage], (kb - tap * p.chunks_per_tap) * kBlockK,
                                                   m0_ld - n * p.v_pix, n, (uint16_t)tap);
                    } else if (p.chunks_per_tap > 0) {
                        const int tap = kb / p.chunks_per_tap;
                        if (PAIR) sb_tma_load_2d_pair(sa, &tmap_a, &full_bar[stage], (kb - tap * p.chunks_per_tap) * kBlockK, m0_ld + tap);
                        else sb_tma_load_2d(sa, &tmap_a, &full_bar[stage], (kb - tap * p.chunks_per_tap) * kBlockK, m0_ld + tap);
                    } else {
                        if (PAIR) sb_tma_load_2d_pair(sa, &tmap_a, &full_bar[stage], kb * kBlockK, m0_ld);
                        else sb_tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * kBlockK, m0_ld);
                    }
                    if (PAIR) sb_tma_load_2d_pair(sb, &tmap_b, &full_bar[stage], kb * kBlockK, n0);
                    else sb_tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * kBlockK, n0);
                    if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0 && leader) {
            const uint32_t idesc = sb_umma_idesc_bf16((uint32_t)m_rows_tile, (uint32_t)p.block_n);
            uint32_t stage = 0, phase = 0;
            int it = 0;
            for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
                const uint32_t as = it & 1, aphase = (it >> 1) & 1;
                if (PAIR) sb_mbar_wait_cluster(&tempty_bar[as], aphase ^ 1);     // arrivals come from both CTAs
                else sb_mbar_wait(&tempty_bar[as], aphase ^ 1);
                sb_tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * p.acc_stride;
                const int kb0 = (tile / tiles_mn) * p.kb_per_split;
                const int kb1 = kb0 + p.kb_per_split < p.num_k_blocks ? kb0 + p.kb_per_split : p.num_k_blocks;
                for (int kb = kb0; kb < kb1; ++kb) {
                    sb_mbar_wait(&full_bar[stage], phase);
                    sb_tc_fence_after();
                    const uint32_t sa = sb_smem_u32(smem + stage * p.stage_bytes);
                    const uint64_t da = sb_umma_desc_sw128(sa);
                    const uint64_t db = sb_umma_desc_sw128(sa + kABytes);
                    // im2col: the last chunk of a tap may hold fewer than 64 real channels (480 = 7 x 64 + 32; TMA zero-fills
                    // the rest and the per-tap weights are zero-padded).  Its empty K slices are multiplied all the same
                    // unless SB_TC_SKIPK=1: skipping them halves that k-block's tensor time but not its 46 KB of loads, and
                    // the ring then starves (r02, conv3: 11.0 ms with the slices, 12.5 ms without them, 11.6 ms before im2col)
                    int nk = kBlockK / 16;
                    if (p.last_k16 < nk && (kb + 1) % p.chunks_per_tap == 0) nk = p.last_k16;
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k)
                        if (k < nk) {
                            if (PAIR) sb_umma_bf16_pair(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                                        (kb > kb0 || k > 0) ? 1u : 0u);
                            else sb_umma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                              (kb > kb0 || k > 0) ? 1u : 0u);
                        }
                    // frees the smem stage (of both CTAs) when the MMAs finish
                    if (PAIR) sb_umma_commit_pair(&empty_bar[stage]);
                    else sb_umma_commit(&empty_bar[stage]);
                    if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
                }
                // accumulator complete (in both CTAs)
                if (PAIR) sb_umma_commit_pair(&tfull_bar[as]);
                else sb_umma_commit(&tfull_bar[as]);
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quad = warp & 3;                  // TMEM lane quadrant this warp may read
        const int row_in_tile = quad * 32 + lane;
        const int et = threadIdx.x - 64;            // 0..127
        int it = 0;
        for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
            const uint32_t as = it & 1, aphase = (it >> 1) & 1;
            const int mn = tile % tiles_mn;
            const int m0 = (mn / p.n_tiles_n) * m_rows_tile + (int)rank * kBlockM;
            const int n0 = (mn % p.n_tiles_n) * p.block_n;
            float* bs = bias_s + as * kMaxN;
            for (int i = et; i < p.block_n; i += 128) bs[i] = __ldg(p.bias + n0 + i);
            asm volatile("bar.sync 1, 128;" ::: "memory");

            // output row of this lane's pool window
            // (32-bit arithmetic: rows_a < 2^31 is checked at launch; pool is 1, 2 or 4 and t_in % pool == 0, so the
            // window start is a mask and the only division left is the one by t_in)
            const uint32_t r = (uint32_t)m0 + (uint32_t)row_in_tile;
            const uint32_t rg = r & ~(uint32_t)(p.pool - 1);  // first row of the pool window
            const uint32_t rps = (uint32_t)(p.im2col ? p.v_pix : p.t_in);   // rows of the M axis per sequence
            const uint32_t seq = rg / rps;
            const int t = (int)(rg - seq * rps);
            const int tp = p.pool == 4 ? t >> 2 : (p.pool == 2 ? t >> 1 : t);
            const bool valid = ((long long)rg < p.rows_a) && (tp < p.t_pool);
            const long long out_row = (long long)seq * p.out_pitch + tp;

            sb_mbar_wait(&tfull_bar[as], aphase);
            sb_tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * p.acc_stride;
            const int n_chunks = p.block_n / 16;
            for (int c = 0; c < n_chunks; ++c) {
                const int nb = n0 + c * 16;
                if (nb >= (p.out_f32 == 2 ? p.sc.F : p.out_ld)) break;  // uniform
                uint32_t rr[16];
                sb_tmem_ld16(taddr + c * 16, rr);
                sb_tmem_wait_ld();
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rr[j]) + bs[c * 16 + j];
                if (p.pool >= 2) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
                }
                if (p.pool >= 4) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 2));
                }
                if (p.relu) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
                }
                if (p.out_f32 == 2) {
                    // ---- fused sigmoid + score epilogue (kernel (d)) ----
                    const SbScoreEpilogue& sc = p.sc;
                    const long long grow = sc.row0 + r;   // global row of this lane
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = tc_sigmoid(v[j]);
                    if (sc.mode == 0) {
                        if (valid)
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (nb + j < sc.F) sc.pred[grow * sc.F + nb + j] = v[j];
                    } else if (sc.mode == 1) {
                        // rows 2v / 2v+1 are adjacent lanes: exchange, then the even lane scores columns
                        // 0-7 of the chunk and the odd lane columns 8-15
                        const bool odd = lane & 1;
                        float mine[8], other[8];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float o = __shfl_xor_sync(0xffffffffu, v[j], 1);
                            if (j < 8) { if (!odd) { mine[j] = v[j]; other[j] = o; } }
                            else { if (odd) { mine[j - 8] = v[j]; other[j - 8] = o; } }
                        }
                        if (valid) {
                            const long long pair = grow >> 1;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int col = nb + (odd ? 8 : 0) + j;
                                if (col < sc.F)
#ifdef SB_MUTANT_SWAP_LANES   // test-only build (scripts/mutation_check.sh): ref/alt lanes swapped on purpose
                                    tc_store_scores(sc, pair * sc.F + col, odd ? mine[j] : other[j],
                                                    odd ? other[j] : mine[j], true);
#else
                                    tc_store_scores(sc, pair * sc.F + col, odd ? other[j] : mine[j],
                                                    odd ? mine[j] : other[j], true);
#endif
                            }
                        }
                    } else if (valid) {
                        const long long bi = sc.base_index ? (long long)sc.base_index[grow] : 0;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (nb + j < sc.F)
                                tc_store_scores(sc, grow * sc.F + nb + j, __ldg(sc.base_pred + bi * sc.F + nb + j), v[j],
                                                false);
                    }
                } else if (p.out_f32 == 3) {
                    // split-K partial sums (bias is zero, no ReLU / pool): lanes are rows, so each red covers one float
                    float* o = reinterpret_cast<float*>(p.out) + out_row * p.out_ld + nb;
                    if (valid)
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (nb + j < p.out_ld) atomicAdd(o + j, v[j]);
                } else if (p.out_f32) {
                    float* o = reinterpret_cast<float*>(p.out) + out_row * p.out_ld + nb;
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        if (valid && nb + 4 * g + 4 <= p.out_ld)
                            reinterpret_cast<float4*>(o)[g] =
                                make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                } else if (p.pool == 4) {
                    const int sub = lane & 3;  // each lane of the window stores 4 of the 16 channels
                    const float w0 = sub == 0 ? v[0] : sub == 1 ? v[4] : sub == 2 ? v[8] : v[12];
                    const float w1 = sub == 0 ? v[1] : sub == 1 ? v[5] : sub == 2 ? v[9] : v[13];
                    const float w2 = sub == 0 ? v[2] : sub == 1 ? v[6] : sub == 2 ? v[10] : v[14];
                    const float w3 = sub == 0 ? v[3] : sub == 1 ? v[7] : sub == 2 ? v[11] : v[15];
                    if (valid && nb + 4 * sub + 4 <= p.out_ld) {
                        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) +
                                           out_row * p.out_ld + nb + 4 * sub;
                        *reinterpret_cast<uint2*>(o) =
                            make_uint2(pack_bf16x2(w0, w1), pack_bf16x2(w2, w3));
                    }
                } else if (p.pool == 2) {
                    const int sub = lane & 1;
                    float w[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) w[j] = sub == 0 ? v[j] : v[8 + j];
                    if (valid && nb + 8 * sub + 8 <= p.out_ld) {
                        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) +
                                           out_row * p.out_ld + nb + 8 * sub;
                        *reinterpret_cast<uint4*>(o) =
                            make_uint4(pack_bf16x2(w[0], w[1]), pack_bf16x2(w[2], w[3]),
                                       pack_bf16x2(w[4], w[5]), pack_bf16x2(w[6], w[7]));
                    }
                } else {
                    __nv_bfloat16* o =
                        reinterpret_cast<__nv_bfloat16*>(p.out) + out_row * p.out_ld + nb;
#pragma unroll
                    for (int g = 0; g < 2; ++g)
                        if (valid && nb + 8 * g + 8 <= p.out_ld)
                            reinterpret_cast<uint4*>(o)[g] = make_uint4(
                                pack_bf16x2(v[8 * g], v[8 * g + 1]), pack_bf16x2(v[8 * g + 2], v[8 * g + 3]),
                                pack_bf16x2(v[8 * g + 4], v[8 * g + 5]), pack_bf16x2(v[8 * g + 6], v[8 * g + 7]));
                }
            }
            sb_tc_fence_before();
            // 128 arrivals (PAIR: 256, at the leader's barrier) hand the accumulator back to the MMA warp
            if (PAIR) sb_mbar_arrive_cluster(sb_mapa(sb_smem_u32(&tempty_bar[as]), 0u));
            else sb_mbar_arrive(&tempty_bar[as]);
        }
    }

    sb_tc_fence_before();
    if (PAIR) sb_cluster_sync();
    else __syncthreads();
    if (warp == 2) {
        sb_tc_fence_after();
        if (PAIR) sb_tmem_dealloc_pair(tmem_base, (uint32_t)p.tmem_cols);
        else sb_tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------
// host: tensor maps
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

EncodeIm2colFn get_encode_im2col_fn() {
    static EncodeIm2colFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
    fn = reinterpret_cast<EncodeIm2colFn>(p);
    return fn;
}

// Channel-last bf16 activations (sequence, position, channel) as an im2col tensor: dims {channels, t_in, n_seq}; the
// window starts ("base pixels") are positions [0, v_pix) of every sequence (upper corner = v_pix - t_in), a load
// fetches 64 channels of 128 consecutive window starts, continuing in the next sequence, at position + tap.
int make_tmap_im2col(CUtensorMap* tm, const void* base, uint64_t channels, uint64_t ld, uint64_t t_in, uint64_t n_seq,
                     int v_pix) {
    EncodeIm2colFn fn = get_encode_im2col_fn();
    if (!fn) {
        sb_set_error("cuTensorMapEncodeIm2col is not available from the driver");
        return SB_ERR_CUDA;
    }
    cuuint64_t dims[3] = {channels, t_in, n_seq};
    cuuint64_t strides[2] = {ld * 2, t_in * ld * 2};
    int lower[1] = {0};
    int upper[1] = {v_pix - (int)t_in};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, lower, upper,
                    (cuuint32_t)kBlockK, (cuuint32_t)kBlockM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        sb_set_error("cuTensorMapEncodeIm2col failed with CUresult %d (channels=%llu t_in=%llu n_seq=%llu v_pix=%d)", (int)r,
                     (unsigned long long)channels, (unsigned long long)t_in, (unsigned long long)n_seq, v_pix);
        return SB_ERR_CUDA;
    }
    // Drivers up to CUDA 13.1 encode im2col maps of tensors below 128 KB with a flag that makes the load fault; CUTLASS
    // (cute/atom/copy_traits_sm90_im2col.hpp) clears bit 21 of the second descriptor word for them, and so do we.
    int drv = 0;
    if (cudaDriverGetVersion(&drv) == cudaSuccess && drv <= 13010 && n_seq * t_in * ld * 2 < 131072)
        reinterpret_cast<uint64_t*>(tm)[1] &= ~(1ull << 21);
    return SB_OK;
}

// 2-D bf16 tensor: dim0 (contiguous) = inner elements, dim1 = rows with stride row_stride_elems.
int make_tmap(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_elems,
              uint32_t box_inner, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        sb_set_error("cuTensorMapEncodeTiled is not available from the driver");
        return SB_ERR_CUDA;
    }
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {row_stride_elems * 2};
    cuuint32_t box[2] = {box_inner, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        sb_set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu rows=%llu stride=%llu box=%ux%u)",
                     (int)r, (unsigned long long)inner, (unsigned long long)rows,
                     (unsigned long long)row_stride_elems, box_inner, box_rows);
        return SB_ERR_CUDA;
    }
    return SB_OK;
}

}  // namespace

int sb_tc_launch(const SbTcLayer& L, int device, cudaStream_t stream, const SbScoreEpilogue* score) {
    SB_REQUIRE(L.block_n >= 16 && L.block_n <= kMaxN && L.block_n % 16 == 0, "sb_tc_launch: bad block_n %d",
               L.block_n);
    SB_REQUIRE(L.a_ld % 8 == 0 && L.w_ld % 8 == 0, "sb_tc_launch: leading dimensions must be multiples of 8");
    SB_REQUIRE(L.pool == 1 || L.pool == 2 || L.pool == 4, "sb_tc_launch: pool %d unsupported", L.pool);
    SB_REQUIRE(L.t_in % L.pool == 0, "sb_tc_launch: t_in %d not a multiple of pool %d", L.t_in, L.pool);
    SB_REQUIRE(L.out_f32 == 2 ? score != nullptr : (L.out_f32 ? (L.out_ld % 4 == 0) : (L.out_ld % 8 == 0)),
               "sb_tc_launch: bad out_ld %d / missing score epilogue", L.out_ld);
    SB_REQUIRE(!(L.out_f32 && L.pool != 1), "sb_tc_launch: fp32 output cannot pool");
    SB_REQUIRE(L.out_f32 == 3 || L.k_splits <= 1, "sb_tc_launch: split-K needs the accumulating fp32 output mode");
    SB_REQUIRE(!(L.out_f32 == 3 && (L.relu || L.per_tap)), "sb_tc_launch: split-K tiles cannot apply ReLU / per-tap boxes");
    if (L.rows_a <= 0) return SB_OK;
    SB_REQUIRE(L.rows_a < (1ll << 31) - 256, "sb_tc_launch: %lld rows exceed the 32-bit row arithmetic of the epilogue",
               (long long)L.rows_a);

    // The last rows' windows would run past the buffer: shrink the row count so TMA zero-fills them
    // (they are windows that straddle the end of the last sequence and are dropped anyway).
    CUtensorMap tm_a, tm_b;
    int rc;
    int chunks_per_tap = 0;
    const int v_pix = (L.t_conv / L.pool) * L.pool;      // window starts that feed a pooled output
    long long rows_m = L.rows_a;                         // rows of the M axis
    if (!L.per_tap && L.im2col == 2) {
        // run view: every window start owns k_full "channels" (its taps x C run, overlapping the next positions' rows)
        SB_REQUIRE(L.rows_a % L.t_in == 0 && v_pix > 0 && L.t_in - v_pix < 32768,
                   "sb_tc_launch: im2col needs whole sequences (rows %lld, t_in %d)", (long long)L.rows_a, L.t_in);
        const long long n_seq = L.rows_a / L.t_in;
        rc = make_tmap_im2col(&tm_a, L.a, (uint64_t)L.k_full, (uint64_t)L.a_ld, (uint64_t)L.t_in, (uint64_t)n_seq, v_pix);
        rows_m = n_seq * v_pix;
    } else if (L.per_tap && L.im2col) {
        chunks_per_tap = (L.a_ld + kBlockK - 1) / kBlockK;
        SB_REQUIRE(L.k_full == L.taps * chunks_per_tap * kBlockK, "sb_tc_launch: per-tap k_full mismatch");
        SB_REQUIRE(L.rows_a % L.t_in == 0 && v_pix > 0 && L.t_in - v_pix < 32768 && L.taps <= 65535,
                   "sb_tc_launch: im2col needs whole sequences (rows %lld, t_in %d)", (long long)L.rows_a, L.t_in);
        const long long n_seq = L.rows_a / L.t_in;
        rc = make_tmap_im2col(&tm_a, L.a, (uint64_t)L.a_ld, (uint64_t)L.a_ld, (uint64_t)L.t_in, (uint64_t)n_seq, v_pix);
        rows_m = n_seq * v_pix;
    } else if (L.per_tap) {
        chunks_per_tap = (L.a_ld + kBlockK - 1) / kBlockK;
        SB_REQUIRE(L.k_full == L.taps * chunks_per_tap * kBlockK, "sb_tc_launch: per-tap k_full mismatch");
        rc = make_tmap(&tm_a, L.a, (uint64_t)L.a_ld, (uint64_t)L.rows_a, (uint64_t)L.a_ld, kBlockK, kBlockM);
    } else {
        const int64_t span = ((int64_t)L.k_full + L.a_ld - 1) / L.a_ld;  // rows touched by one A row
        const int64_t rows_tma = L.rows_a - (span - 1);
        SB_REQUIRE(rows_tma > 0, "sb_tc_launch: fewer rows than taps");
        rc = make_tmap(&tm_a, L.a, (uint64_t)L.k_full, (uint64_t)rows_tma, (uint64_t)L.a_ld, kBlockK, kBlockM);
    }
    if (rc != SB_OK) return rc;
    // CTA pairs (SB_TC_PAIR=0 switches them off): 256-row tiles, each CTA loads half of the B tile
    const char* narrow_env = getenv("SB_TC_NARROW");
    const bool narrow = L.block_n <= 128 && !(narrow_env && narrow_env[0] == '0');
    const char* pair_env = getenv("SB_TC_PAIR");
    bool pair = !(pair_env && pair_env[0] == '0') && !narrow && L.block_n % 16 == 0 && L.out_f32 != 3 &&
                sb_sm_count(device) % 2 == 0;
    if (pair && !(pair_env && pair_env[0] == '2')) {
        // a pair does the work of two CTAs in the time of one tile, so what decides is the number of waves: a layer whose
        // 128-row tiles fill the SMs exactly (DeepSEA fc1: 148 tiles) would need two half-empty waves of 256-row tiles
        const long long sms = sb_sm_count(device);
        const long long t1 = ((rows_m + kBlockM - 1) / kBlockM) * L.n_tiles_n;
        const long long t2 = ((rows_m + 2 * kBlockM - 1) / (2 * kBlockM)) * L.n_tiles_n;
        const long long w1 = (t1 + sms - 1) / sms, w2 = (t2 + sms / 2 - 1) / (sms / 2);
        pair = w2 * 100 <= w1 * 103;
    }
    rc = make_tmap(&tm_b, L.w, (uint64_t)L.k_full, (uint64_t)L.n_rows_w, (uint64_t)L.w_ld, kBlockK,
                   (uint32_t)(pair ? L.block_n / 2 : L.block_n));
    if (rc != SB_OK) return rc;

    TcParams p;
    p.rows_a = rows_m;
    p.im2col = (L.per_tap && L.im2col) ? 1 : ((!L.per_tap && L.im2col == 2) ? 2 : 0);
    p.v_pix = v_pix;
    p.last_k16 = kBlockK / 16;
    {
        const char* e = getenv("SB_TC_SKIPK");
        if (e && e[0] == '1' && p.im2col == 1) p.last_k16 = (L.a_ld - (chunks_per_tap - 1) * kBlockK + 15) / 16;
    }
    p.t_in = L.t_in;
    p.t_conv = L.t_conv;
    p.pool = L.pool;
    p.t_pool = L.t_conv / L.pool;
    p.out_pitch = L.out_pitch > 0 ? L.out_pitch : p.t_pool;
    p.out_ld = L.out_ld;
    p.relu = L.relu;
    p.out_f32 = L.out_f32;
    p.block_n = L.block_n;
    p.n_tiles_n = L.n_tiles_n;
    p.n_tiles_m = pair ? (int)((rows_m + 2 * kBlockM - 1) / (2 * kBlockM)) : (int)((rows_m + kBlockM - 1) / kBlockM);
    p.num_k_blocks = (L.k_full + kBlockK - 1) / kBlockK;
    p.chunks_per_tap = chunks_per_tap;
    p.k_splits = (L.out_f32 == 3 && L.k_splits > 1) ? (L.k_splits < p.num_k_blocks ? L.k_splits : p.num_k_blocks) : 1;
    p.kb_per_split = (p.num_k_blocks + p.k_splits - 1) / p.k_splits;
    p.k_splits = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;     // no empty range
    p.ring_bytes = narrow ? kRingBytes / 2 : kRingBytes;
    p.tmem_cols = narrow ? 256 : kTmemCols;
    p.acc_stride = narrow ? 128 : kMaxN;
    p.stage_bytes = kABytes + (pair ? L.block_n / 2 : L.block_n) * kBlockK * 2;
    p.stages = p.ring_bytes / p.stage_bytes;
    if (p.stages > kMaxStages) p.stages = kMaxStages;
    p.bias = L.bias;
    p.out = L.out;
    memset(&p.sc, 0, sizeof(p.sc));
    if (score) p.sc = *score;

    static bool attr_set = false;
    if (!attr_set) {
        SB_CHECK_CUDA(cudaFuncSetAttribute(tc_layer_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        SB_CHECK_CUDA(cudaFuncSetAttribute(tc_layer_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        // narrow layers put two CTAs of ~103 KB on an SM: ask for the largest shared-memory carve-out
        SB_CHECK_CUDA(cudaFuncSetAttribute(tc_layer_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                           cudaSharedmemCarveoutMaxShared));
        attr_set = true;
    }
    const int64_t tiles = (int64_t)p.n_tiles_m * p.n_tiles_n * p.k_splits;
    const int sms = sb_sm_count(device);
    const int smem_bytes = p.ring_bytes + kAuxBytes + 1024;
    if (pair) {
        int grid = (int)(2 * tiles < sms ? 2 * tiles : sms) & ~1;      // whole clusters of two CTAs
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = (size_t)smem_bytes;
        cfg.stream = stream;
        cudaLaunchAttribute attrs[1];
        attrs[0].id = cudaLaunchAttributeClusterDimension;
        attrs[0].val.clusterDim.x = 2;
        attrs[0].val.clusterDim.y = 1;
        attrs[0].val.clusterDim.z = 1;
        cfg.attrs = attrs;
        cfg.numAttrs = 1;
        SB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tc_layer_kernel<true>, tm_a, tm_b, p));
    } else {
        const int ctas = narrow ? 2 * sms : sms;
        const int grid = (int)(tiles < ctas ? tiles : ctas);
        tc_layer_kernel<false><<<grid, kThreads, smem_bytes, stream>>>(tm_a, tm_b, p);
    }
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    return SB_OK;
}

// ---------------------------------------------------------------------------------------------
// weight-gradient GEMM:  G[n][m] += sum_r  dY[r][m] * X[r*x_ld + n]        (m < M = cout, n < N = taps*cin)
//
// Both operands are contiguous along their M/N index and strided along the reduction index r (rows), i.e.
// MN-major for tcgen05: TMA boxes of 64 (M/N) x 64 (rows) land in the 128-byte-swizzled MN-major atom
// layout (8-row groups 1024 B apart = SBO, 64-element atoms 8192 B apart = LBO) and the instruction
// descriptor carries a_major = b_major = MN.  The reduction is split over `splits` row ranges; partial
// accumulators are added to the fp32 gradient with red.global.add (lanes = consecutive m = contiguous
// floats of G's rows, so every red instruction covers 128 contiguous bytes).
// ---------------------------------------------------------------------------------------------
namespace {

struct WgParams {
    int M, N;
    long long rows;
    int m_tiles, n_tiles, splits, kb_total;
    float* g;       // (N x ldg) fp32
    int ldg;
    int overwrite;  // one K pass per tile: plain stores over all ldg columns instead of adds into a zeroed buffer
    int m_fastest;  // tile order (overwrite mode only)
};

constexpr int kWgBlockN = 256;

__device__ __forceinline__ uint64_t sb_umma_desc_mn_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(8192 >> 4) << 16;   // LBO: next 64-element M/N atom
    d |= (uint64_t)(1024 >> 4) << 32;   // SBO: next group of 8 reduction rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(kThreads, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const WgParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (sb_smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* aux = smem + kStages * kStageBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tfull_bar = empty_bar + kStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_items = p.m_tiles * p.n_tiles * p.splits;
    const int kb_per = (p.kb_total + p.splits - 1) / p.splits;

    if (warp == 0 && lane == 0) { sb_prefetch_tmap(&tmap_a); sb_prefetch_tmap(&tmap_b); }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) { sb_mbar_init(&full_bar[i], 1); sb_mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { sb_mbar_init(&tfull_bar[i], 1); sb_mbar_init(&tempty_bar[i], 128); }
        sb_fence_mbar_init();
    }
    if (warp == 2) sb_tmem_alloc(tmem_slot, kTmemCols);
    sb_tc_fence_before();
    __syncthreads();
    sb_tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // item -> (m tile, n tile, split); n fastest so neighbouring CTAs share the dY tile.  In overwrite mode (a long,
    // single-pass G) m is fastest: the CTAs in flight then store whole rows of G and share the X tile.
    auto decode = [&](int item, int& mt, int& nt, int& kb0, int& kb1) {
        if (p.m_fastest) {
            mt = item % p.m_tiles;
            nt = item / p.m_tiles;
            kb0 = 0;
            kb1 = p.kb_total;
            return;
        }
        nt = item % p.n_tiles;
        const int rest = item / p.n_tiles;
        mt = rest % p.m_tiles;
        const int sp = rest / p.m_tiles;
        kb0 = sp * kb_per;
        kb1 = kb0 + kb_per < p.kb_total ? kb0 + kb_per : p.kb_total;
    };

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                int mt, nt, kb0, kb1;
                decode(item, mt, nt, kb0, kb1);
                for (int kb = kb0; kb < kb1; ++kb) {
                    sb_mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * kStageBytes;
                    uint8_t* sb = sa + kABytes;
                    sb_mbar_expect_tx(&full_bar[stage], kABytes + kBBytesMax);
#pragma unroll
                    for (int a = 0; a < 2; ++a)   // M = 128 = 2 atoms of 64
                        sb_tma_load_2d(sa + a * 8192, &tmap_a, &full_bar[stage], mt * kBlockM + a * 64, kb * kBlockK);
#pragma unroll
                    for (int b = 0; b < 4; ++b)   // N = 256 = 4 atoms of 64
                        sb_tma_load_2d(sb + b * 8192, &tmap_b, &full_bar[stage], nt * kWgBlockN + b * 64, kb * kBlockK);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // bf16 x bf16 -> fp32, A and B MN-major (bits 15 / 16), M = 128, N = 256
            const uint32_t idesc = sb_umma_idesc_bf16(kBlockM, kWgBlockN) | (1u << 15) | (1u << 16);
            uint32_t stage = 0, phase = 0;
            int it = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
                int mt, nt, kb0, kb1;
                decode(item, mt, nt, kb0, kb1);
                const uint32_t as = it & 1, aphase = (it >> 1) & 1;
                sb_mbar_wait(&tempty_bar[as], aphase ^ 1);
                sb_tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * kMaxN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    sb_mbar_wait(&full_bar[stage], phase);
                    sb_tc_fence_after();
                    const uint32_t sa = sb_smem_u32(smem + stage * kStageBytes);
                    const uint64_t da = sb_umma_desc_mn_sw128(sa);
                    const uint64_t db = sb_umma_desc_mn_sw128(sa + kABytes);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k)   // 16 reduction rows = 2048 bytes
                        sb_umma_bf16(tmem_d, da + (uint64_t)(128 * k), db + (uint64_t)(128 * k), idesc,
                                     (kb > kb0 || k > 0) ? 1u : 0u);
                    sb_umma_commit(&empty_bar[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                sb_umma_commit(&tfull_bar[as]);
            }
        }
        __syncwarp();
    } else {
        const int quad = warp & 3;
        const int row_in_tile = quad * 32 + lane;
        int it = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
            int mt, nt, kb0, kb1;
            decode(item, mt, nt, kb0, kb1);
            const uint32_t as = it & 1, aphase = (it >> 1) & 1;
            const int m = mt * kBlockM + row_in_tile;
            sb_mbar_wait(&tfull_bar[as], aphase);
            sb_tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * kMaxN;
            if (kb1 > kb0) {
                for (int c = 0; c < kWgBlockN / 16; ++c) {
                    const int nb = nt * kWgBlockN + c * 16;
                    if (nb >= p.N) break;
                    uint32_t rr[16];
                    sb_tmem_ld16(taddr + c * 16, rr);
                    sb_tmem_wait_ld();
                    if (p.overwrite) {   // rows of dY^T past M are TMA zero fill: the padded columns get their zeros here
                        if (m < p.ldg) {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (nb + j < p.N) p.g[(long long)(nb + j) * p.ldg + m] = __uint_as_float(rr[j]);
                        }
                    } else if (m < p.M) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (nb + j < p.N) atomicAdd(p.g + (long long)(nb + j) * p.ldg + m, __uint_as_float(rr[j]));
                    }
                }
            }
            sb_tc_fence_before();
            sb_mbar_arrive(&tempty_bar[as]);
        }
    }
    sb_tc_fence_before();
    __syncthreads();
    if (warp == 2) { sb_tc_fence_after(); sb_tmem_dealloc(tmem_base, kTmemCols); }
}

}  // namespace

// G (N x ldg, fp32) += dY^T . Xview ; dy: rows x M bf16 (row stride dy_ld), x: bf16 buffer whose row r is the
// run of N elements starting at r*x_ld (overlapping when N > x_ld: taps > 1).
static int wgrad_splits(long long rows, int M, int N, int device) {
    const int tiles = ((M + kBlockM - 1) / kBlockM) * ((N + kWgBlockN - 1) / kWgBlockN);
    const int kb_total = (int)((rows + kBlockK - 1) / kBlockK);
    int splits = (2 * sb_sm_count(device)) / tiles;
    if (splits < 1) splits = 1;
    return splits > kb_total ? kb_total : splits;
}

// true when sb_tc_wgrad_launch(..., overwrite = 1) stores G (all ldg columns of its N rows) instead of adding to it,
// i.e. the caller need not zero G first: every tile is reduced in one K pass and the M tiles cover the padded columns
bool sb_tc_wgrad_overwrites(long long rows, int M, int N, int ldg, int device) {
    const char* e = getenv("SB_WGRAD_OVERWRITE");
    if (e && e[0] == '0') return false;
    return rows > 0 && wgrad_splits(rows, M, N, device) == 1 && ldg <= ((M + kBlockM - 1) / kBlockM) * kBlockM;
}

int sb_tc_wgrad_launch(const void* dy, int dy_ld, const void* x, int x_ld, long long x_rows_alloc, long long rows,
                       int M, int N, float* g, int ldg, int device, cudaStream_t stream, int overwrite) {
    SB_REQUIRE(dy_ld % 8 == 0 && x_ld % 8 == 0, "sb_tc_wgrad_launch: leading dimensions must be multiples of 8");
    if (rows <= 0) return SB_OK;
    const int span = (N + x_ld - 1) / x_ld;                        // rows touched by one X row view
    long long x_rows = x_rows_alloc - (span - 1);                  // keep every read inside the allocation
    if (x_rows > rows) x_rows = rows;
    SB_REQUIRE(x_rows > 0, "sb_tc_wgrad_launch: fewer rows than taps");
    EncodeTiledFn fn = get_encode_fn();
    SB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    CUtensorMap tm_a, tm_b;
    {
        cuuint64_t dims[2] = {(cuuint64_t)M, (cuuint64_t)rows};
        cuuint64_t strides[1] = {(cuuint64_t)dy_ld * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)kBlockK};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = fn(&tm_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(dy), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(dY) failed with CUresult %d", (int)r);
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)x_rows};
        cuuint64_t strides[1] = {(cuuint64_t)x_ld * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)kBlockK};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = fn(&tm_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(X) failed with CUresult %d", (int)r);
    }
    WgParams p;
    p.M = M; p.N = N; p.rows = rows;
    p.m_tiles = (M + kBlockM - 1) / kBlockM;
    p.n_tiles = (N + kWgBlockN - 1) / kWgBlockN;
    p.kb_total = (int)((rows + kBlockK - 1) / kBlockK);
    const int sms = sb_sm_count(device);
    p.splits = wgrad_splits(rows, M, N, device);
    p.g = g; p.ldg = ldg;
    p.overwrite = (overwrite && sb_tc_wgrad_overwrites(rows, M, N, ldg, device)) ? 1 : 0;
    const char* ord = getenv("SB_WGRAD_ORDER");
    p.m_fastest = (p.overwrite && !(ord && ord[0] == '0')) ? 1 : 0;
    static bool attr_set = false;
    if (!attr_set) {
        SB_CHECK_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        attr_set = true;
    }
    const long long items = (long long)p.m_tiles * p.n_tiles * p.splits;
    const int grid = (int)(items < sms ? items : sms);
    tc_wgrad_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tm_a, tm_b, p);
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    return SB_OK;
}

extern "C" int sb_tc_wgrad_selftest(const void* dy_bf16, const void* x_bf16, long long rows, int M, int x_ld, int N,
                                    float* g, void* stream) {
    SB_REQUIRE(dy_bf16 && x_bf16 && g, "sb_tc_wgrad_selftest: null argument");
    int dev = 0;
    SB_CHECK_CUDA(cudaGetDevice(&dev));
    return sb_tc_wgrad_launch(dy_bf16, M, x_bf16, x_ld, rows, rows, M, N, g, M, dev, (cudaStream_t)stream, 0);
}

// ---------------------------------------------------------------------------------------------
// self-test entry: plain C = A . B^T through the same kernel
// ---------------------------------------------------------------------------------------------
__global__ void zero_bias_kernel(float* b, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = 0.f;
}

extern "C" int sb_tc_gemm_selftest(const void* a_bf16, const void* b_bf16, int M, int N, int K, float* c,
                                   void* stream) {
    SB_REQUIRE(a_bf16 && b_bf16 && c, "sb_tc_gemm_selftest: null argument");
    SB_REQUIRE(M > 0 && N > 0 && K > 0 && K % 8 == 0 && N % 4 == 0, "sb_tc_gemm_selftest: need K%%8==0, N%%4==0");
    int dev = 0;
    SB_CHECK_CUDA(cudaGetDevice(&dev));
    int block_n = N >= 256 ? 256 : ((N + 15) / 16) * 16;
    const int n_tiles_n = (N + block_n - 1) / block_n;
    float* bias = nullptr;
    SB_CHECK_CUDA(cudaMalloc(&bias, sizeof(float) * (size_t)n_tiles_n * block_n));
    zero_bias_kernel<<<(n_tiles_n * block_n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(bias, n_tiles_n * block_n);
    SB_COUNT_LAUNCH();
    SbTcLayer L;
    memset(&L, 0, sizeof(L));
    L.a = a_bf16; L.rows_a = M; L.a_ld = K; L.k_full = K;
    L.w = b_bf16; L.n_rows_w = N; L.w_ld = K;
    L.bias = bias; L.block_n = block_n; L.n_tiles_n = n_tiles_n;
    L.t_in = 1; L.t_conv = 1; L.pool = 1; L.relu = 0; L.out_f32 = 1; L.out = c; L.out_ld = N;
    int rc = sb_tc_launch(L, dev, (cudaStream_t)stream);
    cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(bias);
    return rc;
}
