// selene_b200 — first convolution (4 -> cout channels, <= 8 taps, ReLU, max-pool 4) straight from base codes on the
// register-accumulator tensor path (mma.sync.m16n8k16, bf16 x bf16 -> fp32).
//
// Replaces torch.conv1d + relu + max_pool1d of models/deepsea.py:22-25 for one-hot input.  Why not tcgen05 here
// (sb_conv1.cu): K = taps * 4 <= 32, so the layer is 1.8 % of DeepSEA's FLOPs, but it produces 4 unpooled fp32
// accumulators per pooled output and on the tcgen05 path every one of them has to come back from tensor memory through
// tcgen05.ld at 64 B/clk/SM — 41.6 GB per 32 768 sequences = 2.9 ms at the power-capped clock, measured 3.4 ms, 10 % of
// the whole step (profiles/r01_final_summary.md).  With mma.sync the accumulators ARE registers, so the pool is a
// per-thread max and nothing but the pooled bf16 output (5.2 GB) moves.
//
//   unit      = 32 conv positions (8 pooled positions) of one sequence x 80 output channels, owned by one warp:
//               two m16 tiles x 10 n8 tiles x 2 k16 steps = 40 MMAs.
//   rows      tile X row g / g+8 = positions 4g / 4g+1, tile Y row g / g+8 = positions 4g+2 / 4g+3 (relative to the
//               unit), so the four members of pool window g are the four accumulator registers a thread holds for a
//               column: pooling needs no shuffle.
//   A         never touches memory: K index = tap*4 + base.  Dense path: a thread's A registers are the one-hot patterns of
//               the TEN codes at positions 4g + t/2 ... + 9 (each register a bf16 pair: 1.0 in the slot of the base, or
//               0.25 / 0.25 for an unknown base), read from the code row staged in shared memory.  Sparse path (no unknown
//               base in the unit, kernel co-resident with the tcgen05 layers): the one-hot operand is 2:4 sparse, one
//               mma.sp.m16n8k32 covers all 8 taps (see mma_sp_bf16_16832).
//   B         in registers for the whole kernel (a warp keeps its 80 channels): 40 registers; the bias (the initial
//             accumulator value) is read from shared memory.
//   columns   column c of n-tile j is channel 10 c + j of the warp's 80, so a thread ends up with 20 CONSECUTIVE
//             channels of one pooled row (40 bytes, five 8-byte stores; the four lanes of a row write 160 contiguous bytes).
#include "sb_common.cuh"

namespace {

constexpr int kNT = 10;                 // n8 tiles per warp
constexpr int kChWarp = 8 * kNT;        // 80 channels per warp
constexpr int kWarps = 8;
constexpr int kUnitPos = 32;            // conv positions per unit
constexpr int kUnitPooled = 8;

struct Conv1mParams {
    const uint8_t* codes;
    const int32_t* src;
    const int32_t* sub_pos;
    const uint8_t* sub_code;
    long long row0, n_rows;
    int L, t_pool, out_pitch, out_ld, relu, n_cg, units_per_seq, rows_async, row_bytes, dense_only;
    const uint32_t* bfrag;   // [n_cg][kNT][2 k-steps][2][32 lanes]
    const float* bias;       // [n_cg * 80]
    __nv_bfloat16* out;
};

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t a0, const uint32_t a1, const uint32_t a2,
                                               const uint32_t a3, const uint32_t b0, const uint32_t b1,
                                               const float c0, const float c1, const float c2, const float c3) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(c0), "f"(c1), "f"(c2), "f"(c3));
}

// 2:4 structured-sparse A (one-hot input: exactly one of the four bases of a tap is set): A holds two values per group of
// four K (here (1, 0) or (0, 1)), the metadata nibble of the group names their positions ((0,1) or (2,3)); K = 32 in ONE
// instruction.  Layouts measured with scripts/probes/probe_mma_sp.cu: a0 / a1 = rows g / g+8 of group t, a2 / a3 of group
// t + 4; metadata (selector 0) comes from lanes 4g and 4g+1: nibbles 0-3 = row g, groups 4s .. 4s+3, nibbles 4-7 = row g+8.
__device__ __forceinline__ void mma_sp_bf16_16832(float (&d)[4], const uint32_t a0, const uint32_t a1, const uint32_t a2,
                                                  const uint32_t a3, const uint32_t b0, const uint32_t b1, const uint32_t b2,
                                                  const uint32_t b3, const uint32_t meta) {
    asm volatile(
        "mma.sp::ordered_metadata.sync.aligned.m16n8k32.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
        "{%8,%9,%10,%11}, {%12,%12,%12,%12}, %13, 0x0;"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(b2), "r"(b3), "f"(0.f), "r"(meta));
}

// bf16 pair (base 2h, base 2h + 1) of the one-hot row of `code`; 4 (and anything above) is the unknown base
__device__ __forceinline__ uint32_t pair_pattern(uint32_t code, uint32_t h) {
    const uint32_t hot = (code & 1u) ? 0x3F800000u : 0x00003F80u;
    return code >= 4u ? 0x3E803E80u : ((code >> 1) == h ? hot : 0u);
}

__device__ __forceinline__ float max3f(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

__device__ __forceinline__ uint32_t pack_pair(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

template <bool RELU>
__global__ void __launch_bounds__(kWarps * 32, 2) conv1_mma_kernel(const Conv1mParams p) {
    __shared__ float bias_s[kWarps * kChWarp];
    for (int i = threadIdx.x; i < p.n_cg * kChWarp; i += blockDim.x) bias_s[i] = __ldg(p.bias + i);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int streams_per_cta = kWarps / p.n_cg;            // position streams of this CTA
    if (warp >= streams_per_cta * p.n_cg) return;
    const int cg = warp % p.n_cg;
    const long long stream = (long long)blockIdx.x * streams_per_cta + warp / p.n_cg;
    const long long n_streams = (long long)gridDim.x * streams_per_cta;

    // weights and bias of this warp's 80 channels: registers for the whole kernel
    uint32_t bw[kNT][2][2];
#pragma unroll
    for (int j = 0; j < kNT; ++j)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int r = 0; r < 2; ++r) bw[j][ks][r] = __ldg(p.bfrag + ((((size_t)cg * kNT + j) * 2 + ks) * 2 + r) * 32 + lane);
    // column 2t of n-tile j -> channel 20 t + j of the group, column 2t + 1 -> channel 20 t + 10 + j
    const float* bias_lo = bias_s + cg * kChWarp + 20 * t;
    const float* bias_hi = bias_lo + 10;
    const uint32_t h = (uint32_t)(t & 1);
    const int th = t >> 1;

    // A warp owns whole sequences.  The code row of a sequence (L bytes) is copied to the warp's slice of shared memory
    // with 8-byte cp.async one sequence AHEAD (double buffered), its per-row metadata (source row, substitution) is read
    // two sequences ahead, so nothing in the unit loop waits on global memory.
    extern __shared__ __align__(16) uint8_t rows_s[];
    uint8_t* rbuf0 = rows_s + (size_t)(2 * warp) * p.row_bytes;      // two buffers of row_bytes, back to back
    for (int i = lane * 4; i < 2 * p.row_bytes; i += 128) *reinterpret_cast<uint32_t*>(rbuf0 + i) = 0x04040404u;
    // sparse-metadata nibble of every position of the current row (0x4: bases 0 / 1 -> slots (0,1); 0xE: bases 2 / 3 ->
    // slots (2,3)), eight per word: a unit's metadata registers are then two shifts of one 32-bit window of this stream
    const int mwords = p.row_bytes / 8 + 2;
    uint32_t* mstream = reinterpret_cast<uint32_t*>(rows_s + (size_t)2 * kWarps * p.row_bytes) + (size_t)warp * mwords;
    __syncwarp();

    struct RowMeta { long long srow; int sp; uint32_t sc; };
    auto load_meta = [&](long long sq) {
        RowMeta m;
        m.srow = 0; m.sp = -1; m.sc = 4u;
        if (sq < p.n_rows) {
            const long long row = p.row0 + sq;
            m.srow = p.src ? (long long)__ldg(p.src + row) : row;
            if (p.sub_pos) {
                m.sp = __ldg(p.sub_pos + row);
                if (p.sub_code) m.sc = (uint32_t)__ldg(p.sub_code + row);
            }
        }
        return m;
    };
    auto copy_row = [&](const RowMeta& m, uint8_t* dst) {
        const uint8_t* crow = p.codes + m.srow * p.L;
        if (p.rows_async) {
            const uint32_t d0 = sb_smem_u32(dst);
            for (int i = lane * 8; i < p.L; i += 256)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d0 + i), "l"(crow + i) : "memory");
        } else {
            for (int i = lane; i < p.L; i += 32) dst[i] = __ldg(crow + i);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    long long seq = stream;
    RowMeta m0 = load_meta(seq), m1 = load_meta(seq + n_streams);
    if (seq < p.n_rows) copy_row(m0, rbuf0);
    const int ch0 = cg * kChWarp + 20 * t;
    for (int it = 0; seq < p.n_rows; seq += n_streams, ++it) {
        uint8_t* cur = rbuf0 + (it & 1) * p.row_bytes;
        const bool has_next = seq + n_streams < p.n_rows;
        if (has_next) copy_row(m1, rbuf0 + ((it & 1) ^ 1) * p.row_bytes);
        const RowMeta m2 = load_meta(seq + 2 * n_streams);
        if (has_next) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (lane == 0 && m0.sp >= 0 && m0.sp < p.L) cur[m0.sp] = (uint8_t)m0.sc;
        __syncwarp();
        for (int wi = lane; wi < mwords - 2; wi += 32) {
            const uint32_t c_lo = reinterpret_cast<const uint32_t*>(cur)[2 * wi], c_hi = reinterpret_cast<const uint32_t*>(cur)[2 * wi + 1];
            const uint32_t z = ((c_lo >> 1) & 0x1u) | ((c_lo >> 5) & 0x10u) | ((c_lo >> 9) & 0x100u) | ((c_lo >> 13) & 0x1000u) |
                               ((c_hi << 15) & 0x10000u) | ((c_hi << 11) & 0x100000u) | ((c_hi << 7) & 0x1000000u) |
                               ((c_hi << 3) & 0x10000000u);
            mstream[wi] = 0x44444444u + 10u * z;
        }
        __syncwarp();
        const uint32_t* cw = reinterpret_cast<const uint32_t*>(cur) + g;
        __nv_bfloat16* orow = p.out + (seq * p.out_pitch + g) * (long long)p.out_ld + ch0;
#pragma unroll 1
        for (int u = 0; u < p.units_per_seq; ++u, cw += 8, orow += (long long)kUnitPooled * p.out_ld) {
        // this thread's ten codes: bytes 4g + th ... + 9 of the unit
        const uint32_t w0 = cw[0], w1 = cw[1], w2 = cw[2];
        uint32_t o_lo[kNT / 2], o_hi[kNT / 2];
        float keep_lo = 0.f, keep_hi = 0.f;
        // max over the pool window (tile X: positions 4g, 4g+1; tile Y: 4g+2, 4g+3), then bias + ReLU as max(m, -bias) + bias
        auto finish = [&](int j, const float (&X)[4], const float (&Y)[4]) {
            const float bl = bias_lo[j], bh = bias_hi[j];
            const float lo = max3f(max3f(X[0], X[2], Y[0]), Y[2], RELU ? -bl : -INFINITY) + bl;     // column 2t
            const float hi = max3f(max3f(X[1], X[3], Y[1]), Y[3], RELU ? -bh : -INFINITY) + bh;     // column 2t + 1
            if (j & 1) {
                o_lo[j >> 1] = pack_pair(keep_lo, lo);
                o_hi[j >> 1] = pack_pair(keep_hi, hi);
            } else {
                keep_lo = lo;
                keep_hi = hi;
            }
        };
        if (p.dense_only || __any_sync(0xffffffffu, ((w0 | w1 | w2) & 0xFCFCFCFCu) != 0u)) {
            // an unknown base somewhere in the unit (rare): dense MMAs, general one-hot patterns (0.25 x 4 for the unknown)
            const uint32_t x0 = th ? __funnelshift_r(w0, w1, 8) : w0;    // bytes th .. th+3
            const uint32_t x1 = th ? __funnelshift_r(w1, w2, 8) : w1;    // bytes th+4 .. th+7
            const uint32_t x2 = th ? (w2 >> 8) : w2;                     // bytes th+8, th+9
            uint32_t P[10];
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                P[d] = pair_pattern((x0 >> (8 * d)) & 0xFFu, h);
                P[4 + d] = pair_pattern((x1 >> (8 * d)) & 0xFFu, h);
            }
            P[8] = pair_pattern(x2 & 0xFFu, h);
            P[9] = pair_pattern((x2 >> 8) & 0xFFu, h);
#pragma unroll
            for (int j = 0; j < kNT; ++j) {
                float X[4], Y[4];
                // k-step 0 = taps th, th+2; 1 = th+4, th+6
                mma_bf16_16816(X, P[0], P[1], P[2], P[3], bw[j][0][0], bw[j][0][1], 0.f, 0.f, 0.f, 0.f);
                mma_bf16_16816(Y, P[2], P[3], P[4], P[5], bw[j][0][0], bw[j][0][1], 0.f, 0.f, 0.f, 0.f);
                mma_bf16_16816(X, P[4], P[5], P[6], P[7], bw[j][1][0], bw[j][1][1], X[0], X[1], X[2], X[3]);
                mma_bf16_16816(Y, P[6], P[7], P[8], P[9], bw[j][1][0], bw[j][1][1], Y[0], Y[1], Y[2], Y[3]);
                finish(j, X, Y);
            }
        } else {
            // codes 0..3: ONE sparse MMA per (tile, n-tile).  Value pairs of this thread's taps t and t + 4 at the tile rows:
            // the codes at bytes 4g + t ... + 7 of the unit; the slot is the low bit of the code
            const uint32_t x0 = __funnelshift_r(w0, w1, 8 * t), x1 = __funnelshift_r(w1, w2, 8 * t);
            uint32_t V[8];
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                V[d] = 0x3F80u << ((d == 0 ? (x0 << 4) : (x0 >> (8 * d - 4))) & 16u);
                V[4 + d] = 0x3F80u << ((d == 0 ? (x1 << 4) : (x1 >> (8 * d - 4))) & 16u);
            }
            // metadata: nibbles of positions 4g + 4s ... of the unit (s = lane & 1; only lanes 4g and 4g + 1 are read)
            const int hw = 8 * u + g + (t & 1);
            const uint32_t win = __funnelshift_r(mstream[hw >> 1], mstream[(hw >> 1) + 1], 16 * (hw & 1));
            const uint32_t eX = (win & 0xFFFFu) | ((win << 12) & 0xFFFF0000u);
            const uint32_t eY = ((win >> 8) & 0xFFFFu) | ((win << 4) & 0xFFFF0000u);
#pragma unroll
            for (int j = 0; j < kNT; ++j) {
                float X[4], Y[4];
                mma_sp_bf16_16832(X, V[0], V[1], V[4], V[5], bw[j][0][0], bw[j][0][1], bw[j][1][0], bw[j][1][1], eX);
                mma_sp_bf16_16832(Y, V[2], V[3], V[6], V[7], bw[j][0][0], bw[j][0][1], bw[j][1][0], bw[j][1][1], eY);
                finish(j, X, Y);
            }
        }
        // pooled row 8u + g, channels cg*80 + 20t ... + 19
        if (u * kUnitPooled + g < p.t_pool) {
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const uint32_t a = k < 2 ? o_lo[2 * k] : (k == 2 ? o_lo[4] : o_hi[2 * k - 5]);
                const uint32_t b = k < 2 ? o_lo[2 * k + 1] : (k == 2 ? o_hi[0] : o_hi[2 * k - 4]);
                if (ch0 + 4 * k + 4 <= p.out_ld) *reinterpret_cast<uint2*>(orow + 4 * k) = make_uint2(a, b);
            }
        }
        }
        __syncwarp();      // every lane is done with this buffer before the copy of the sequence after next lands in it
        m0 = m1;
        m1 = m2;
    }
}

}  // namespace

int sb_conv1_mma_supported(int taps, int pool, int cout, int out_ld) {
    const int n_cg = (cout + kChWarp - 1) / kChWarp;
    return taps >= 1 && taps <= 8 && pool == 4 && cout >= 1 && n_cg <= kWarps && out_ld % 4 == 0;
}

// Host: conv weights (cout, 4, taps) fp32 -> the per-lane B fragments of mma.m16n8k16 (bf16 pairs), K = tap*4 + base in the
// caller's channel order, column c of n-tile j = channel 10 c + j of the group.  Returns the number of uint32 written.
size_t sb_conv1_mma_pack_weights(const float* w, int cout, int taps, const uint8_t perm[4], uint32_t* out,
                                 uint16_t (*to_bf16)(float)) {
    const int n_cg = (cout + kChWarp - 1) / kChWarp;
    const size_t n = (size_t)n_cg * kNT * 2 * 2 * 32;
    if (!out) return n;
    auto wk = [&](int ch, int k) -> uint32_t {
        const int tap = k >> 2, b = k & 3;
        if (ch >= cout || tap >= taps) return 0u;
        return (uint32_t)to_bf16(w[((size_t)ch * 4 + perm[b]) * taps + tap]);
    };
    for (int cg = 0; cg < n_cg; ++cg)
        for (int j = 0; j < kNT; ++j)
            for (int ks = 0; ks < 2; ++ks)
                for (int r = 0; r < 2; ++r)
                    for (int lane = 0; lane < 32; ++lane) {
                        const int g = lane >> 2, t = lane & 3;
                        const int ch = cg * kChWarp + g * kNT + j;
                        const int k0 = ks * 16 + r * 8 + 2 * t;
                        out[((((size_t)cg * kNT + j) * 2 + ks) * 2 + r) * 32 + lane] = wk(ch, k0) | (wk(ch, k0 + 1) << 16);
                    }
    return n;
}

int sb_conv1_mma_launch(const uint8_t* codes, const int32_t* src, const int32_t* sub_pos, const uint8_t* sub_code,
                        long long row0, long long n_rows, int L, int taps, int pool, int t_pool, int out_pitch, int cout,
                        int out_ld, int relu, const void* bfrag, const float* bias_padded, void* out, int device,
                        cudaStream_t stream, int ctas_per_sm) {
    if (n_rows <= 0) return SB_OK;
    SB_REQUIRE(sb_conv1_mma_supported(taps, pool, cout, out_ld), "sb_conv1_mma_launch: unsupported shape");
    Conv1mParams p;
    p.codes = codes; p.src = src; p.sub_pos = sub_pos; p.sub_code = sub_code;
    p.row0 = row0; p.n_rows = n_rows;
    p.L = L; p.t_pool = t_pool; p.out_pitch = out_pitch > 0 ? out_pitch : t_pool; p.out_ld = out_ld; p.relu = relu;
    p.n_cg = (cout + kChWarp - 1) / kChWarp;
    p.units_per_seq = (t_pool + kUnitPooled - 1) / kUnitPooled;
    p.rows_async = (L % 8 == 0) && ((reinterpret_cast<uintptr_t>(codes) & 7) == 0);      // 8-byte cp.async per row chunk
    p.row_bytes = ((kUnitPos * p.units_per_seq + 16 > L ? kUnitPos * p.units_per_seq + 16 : L) + 15) & ~15;
    // Alone on the GPU (two CTAs per SM) the kernel is latency-bound and the dense MMAs' shorter setup wins (3.0 vs 3.5 ms
    // per 32 768 sequences); next to a tc_layer_kernel CTA what counts is the tensor-pipe time it takes from the tcgen05
    // MMAs, which the sparse MMAs halve (r02: step 31.35 vs 31.5 ms).  SB_CONV1_SPARSE=0 / 1 forces one or the other.
    {
        const char* e = getenv("SB_CONV1_SPARSE");
        p.dense_only = e ? (e[0] == '0') : (ctas_per_sm != 1);
    }
    p.bfrag = reinterpret_cast<const uint32_t*>(bfrag);
    p.bias = bias_padded;
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    const int streams_per_cta = kWarps / p.n_cg;
    long long ctas = (n_rows + streams_per_cta - 1) / streams_per_cta;      // a position stream owns whole sequences
    const size_t smem = (size_t)2 * kWarps * p.row_bytes + (size_t)kWarps * (p.row_bytes / 8 + 2) * 4;
    SB_REQUIRE(smem <= 100 * 1024, "sb_conv1_mma_launch: sequences of %d bases need %zu bytes of shared memory", L, smem);
    static size_t attr = 0;
    if (smem > attr) {
        SB_CHECK_CUDA(cudaFuncSetAttribute(conv1_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SB_CHECK_CUDA(cudaFuncSetAttribute(conv1_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    // persistent: two CTAs of 8 warps per SM; one when the launch shares the SMs with a tc_layer_kernel CTA (sb_net.cu)
    const long long cap = (long long)sb_sm_count(device) * (ctas_per_sm == 1 ? 1 : 2);
    if (ctas > cap) ctas = cap;
    if (relu) conv1_mma_kernel<true><<<(int)ctas, kWarps * 32, smem, stream>>>(p);
    else conv1_mma_kernel<false><<<(int)ctas, kWarps * 32, smem, stream>>>(p);
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    return SB_OK;
}
