// selene_b200 — kernels (c)+(d), host side: conv1d-stack networks (DeepSEA family) and fused scoring.
//
// Reference: selene_sdk/predict/_common.py:65-97 (predict) running models/deepsea.py:21-58 —
//   conv(4->320,k8)+ReLU+pool4 -> conv(320->480,k8)+ReLU+pool4 -> conv(480->960,k8)+ReLU ->
//   flatten channel-major (deepsea.py:56) -> FC+ReLU -> FC -> sigmoid; Dropout is identity in eval —
// and the score handlers predict_handlers/{absolute_diff,diff,logit}_score_handler.py.
//
// Data layout in HBM: activations are channel-last, (sequence, position, channel), so that
//   * the first conv is a table lookup on base codes (one-hot input => each tap selects a weight
//     column; unknown bases select the 0.25-weighted mean column),
//   * every later conv/FC is the implicit GEMM of sb_tc.cu (bf16 / tcgen05) or of the fp32 FFMA
//     kernel below (accuracy mode, <= 1e-5 abs on the sigmoid outputs),
//   * the reference's channel-major flatten is absorbed into a one-time permutation of the first FC
//     weight matrix.
#include "sb_common.cuh"
#include "sb_tc.cuh"

#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

namespace {

struct NetLayer {
    int type, cin, cout, k, relu, pool;
    int t_in, t_conv, t_pool;  // per-sequence rows in / valid conv rows / rows out (FC: 1,1,1)
    int pitch_in, pitch_out;   // bf16 path: rows per sequence of the in/out buffers (t_in rounded up to this
                               // layer's pool / the next layer's pitch_in) so pool windows never straddle sequences
    // fp32 path
    float* wf;     // dev (k_f32 x ldw_f32): row = K index (tap*cin + c), col = cout (padded to x4)
    float* biasf;  // dev (ldw_f32)
    int k_f32, ldw_f32;
    // bf16 path
    __nv_bfloat16* wb;      // dev (cout x w_ld): K index = tap*a_ld + c   (overlapping-stride A view)
    __nv_bfloat16* wb_tap;  // dev (cout x w_ld_tap): K index = tap*cpk*64 + c (per-tap boxes); conv only
    float* biasb;           // dev, padded to n_tiles_n*block_n
    int a_ld, k_full, w_ld, k_full_tap, w_ld_tap, block_n, n_tiles_n, out_ld_b;
    // first layer from codes
    float* lut;  // dev (k x 5 x cout)
    uint16_t* w1_packed;  // dev, tcgen05 core-matrix layout (sb_conv1.cu); null if unsupported
    uint16_t* w1l_packed; // dev, the same for sb_conv1l.cu (many taps, wide pool); null if unsupported
    float* bias1_padded;  // dev, n_pass*64
    uint32_t* w1m_packed; // dev, mma.sync B fragments (sb_conv1m.cu: pool-4 first conv on register accumulators); null if unsupported
    float* bias1m;        // dev, groups of 80 channels
    // first layer too long / pooled too wide for sb_conv1.cu (DanQ: 26 taps, pool 13): generic layer kernel
    // on a (n, L, 8) bf16 one-hot (4 zero channels), K index = tap*8 + canonical base
    __nv_bfloat16* w0_b;
    int w0_ld, w0_block_n, w0_n_tiles_n;
    float* w0_bias;
    // SB_LAYER_BILSTM: the generic fields above hold the input projection (cin -> 8*hidden, both
    // directions, bias = b_ih + b_hh); these hold the recurrent matrices (hidden -> 4*hidden) per direction
    int hidden;
    float* whh_f[2];            // dev (hidden x 4*hidden): row = K index
    __nv_bfloat16* whh_b[2];    // dev (4*hidden x hidden_ld) K-major
    float* zero_bias;           // dev, zeros (padded for the tcgen05 tiles)
    // one-launch recurrence (sb_lstm.cu): projection weights / bias and recurrent matrices with their gate rows
    // in the tile order of sb_lstm_perm_row; null when the hidden size is unsupported
    __nv_bfloat16* wb_perm;     // dev (8*hidden x w_ld)
    float* biasb_perm;          // dev, padded like biasb
    __nv_bfloat16* whh_perm[2]; // dev (4*hidden x hidden_ld)
    int hh_block_n, hh_n_tiles_n, hidden_ld;
    int fc_prev_c, fc_prev_t;   // first FC after the conv stack: channels / positions it flattens (else 0)
};

}  // namespace

struct sb_train_state;
struct sb_train_tc_state;

struct LstmGraphKey {
    const void* act_in; void* act_out; float* xw; uint8_t* ws;
    long long m; int group, interleave, precision; const void* layer;
};
struct LstmGraph { LstmGraphKey key; cudaGraphExec_t exec; unsigned long long n_launches; };

struct sb_net {
    sb_train_state* train;          // sb_train.inl; null until sb_net_train_init
    sb_train_tc_state* train_tc;    // sb_train_tc.inl; null until sb_net_train_init_tc
    float* zero_bias_tc;            // dev zeros used as the bias of gradient GEMMs
    bool inference_weights_stale;   // set by sb_net_sgd_step: packed bf16 / lookup copies no longer match wf
    int device;
    int L;
    int n_layers;
    int n_out;
    uint8_t perm[4];
    double flops_per_seq;
    bool force_per_tap;
    bool no_im2col;                  // SB_TC_IM2COL=0, or the driver refused an im2col tensor map
    bool no_im2col_run;              // SB_TC_IM2COL=1 (per-tap maps only), or the driver refused the run-view map
    int rec_group, rec_interleave;   // row grouping of BiLSTM recurrences (sb_net_set_recurrence)
    int strand_mode;                 // 0: plain; 1 / 2: NonStrandSpecific mean / max (sb_net_set_strand_mode)
    bool has_lstm;
    std::vector<NetLayer> layers;
    std::vector<LstmGraph> lstm_graphs;   // captured BiLSTM layers (run_lstm)
    cudaStream_t capture_stream;
    // optional per-layer device timing (bench.py's roofline leg): one event pair per launch
    bool timing;
    std::vector<std::vector<cudaEvent_t>> ev;  // per layer: start0, stop0, start1, stop1, ...
    std::vector<size_t> ev_used;
    // first conv one chunk ahead (forward_impl): side stream + ordering events, created on first use
    cudaStream_t side;
    cudaEvent_t ev_fork, ev_c1[2], ev_done[2];
};

// sb_lstm.cu
int sb_lstm_persistent_supported(int H);
int sb_lstm_persistent_launch(const void* whh0, const void* whh1, int h_ld, const void* xw, void* out, int out_ld,
                              long long m_rows, int H, int T, int G, int I, int device, cudaStream_t stream);
int sb_lstm_perm_row(int n, int H);

// sb_conv1.cu
size_t sb_conv1_pack_weights(const float* w, int cout, int taps, const uint8_t perm[4], uint16_t* out,
                             uint16_t (*to_bf16)(float));
int sb_conv1_tc_supported(int taps, int pool, int cout);
// sb_conv1m.cu: pool-4 first conv on mma.sync (accumulators in registers)
int sb_conv1_mma_supported(int taps, int pool, int cout, int out_ld);
size_t sb_conv1_mma_pack_weights(const float* w, int cout, int taps, const uint8_t perm[4], uint32_t* out,
                                 uint16_t (*to_bf16)(float));
int sb_conv1_mma_launch(const uint8_t* codes, const int32_t* src, const int32_t* sub_pos, const uint8_t* sub_code,
                        long long row0, long long n_rows, int L, int taps, int pool, int t_pool, int out_pitch, int cout,
                        int out_ld, int relu, const void* bfrag, const float* bias_padded, void* out, int device,
                        cudaStream_t stream, int ctas_per_sm);
// sb_conv1l.cu: many taps + wide pool (DanQ)
int sb_conv1_long_supported(int taps, int pool, int cout);
size_t sb_conv1_long_pack_weights(const float* w, int cout, int taps, const uint8_t perm[4], uint16_t* out,
                                  uint16_t (*to_bf16)(float));
int sb_conv1_long_launch(const uint8_t* codes, const int32_t* src, const int32_t* sub_pos, const uint8_t* sub_code,
                         long long row0, long long n_rows, int L, int taps, int pool, int t_pool, int out_pitch, int cout,
                         int out_ld, int relu, const void* wpacked, const float* bias_padded, void* out, int device,
                         cudaStream_t stream);
int sb_conv1_tc_launch(const uint8_t* codes, const int32_t* src, const int32_t* sub_pos, const uint8_t* sub_code,
                       long long row0, long long n_rows, int L, int taps, int pool, int t_pool, int out_pitch, int cout,
                       int out_ld, int relu, const void* wpacked, const float* bias_padded, void* out, int device,
                       cudaStream_t stream);

namespace {

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// ---------------------------------------------------------------------------------------------
// first conv from base codes: out[t][n] = bias[n] + sum_j lut[j][code[t+j]][n], ReLU, max-pool
// ---------------------------------------------------------------------------------------------
template <typename OutT>
__device__ __forceinline__ OutT to_out(float v);
template <>
__device__ __forceinline__ float to_out<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename OutT>
__global__ void __launch_bounds__(512) conv_codes_kernel(
    const uint8_t* __restrict__ codes, const int32_t* __restrict__ src, const int32_t* __restrict__ sub_pos,
    const uint8_t* __restrict__ sub_code, long long row0, int L, int k, int cout, int out_ld, int relu, int pool,
    int t_pool, int out_pitch, const float* __restrict__ lut, const float* __restrict__ bias,
    OutT* __restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    float* tab = reinterpret_cast<float*>(smem_raw);             // k*5*cout
    uint8_t* sc = smem_raw + (size_t)k * 5 * cout * sizeof(float);  // L codes
    const long long row = row0 + blockIdx.x;
    const long long srow = src ? src[row] : row;
    const uint8_t* crow = codes + srow * L;
    const int sp = sub_pos ? sub_pos[row] : -1;
    const int scode = (sub_pos && sub_code) ? sub_code[row] : 4;
    for (int i = threadIdx.x; i < k * 5 * cout; i += blockDim.x) tab[i] = __ldg(lut + i);
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        int c = __ldg(crow + i);
        if (i == sp) c = scode;
        sc[i] = (uint8_t)(c > 4 ? 4 : c);
    }
    __syncthreads();
    OutT* orow = out + (long long)blockIdx.x * out_pitch * out_ld;
    for (int n = threadIdx.x; n < out_ld; n += blockDim.x) {
        if (n >= cout) {
            for (int tp = 0; tp < t_pool; ++tp) orow[(long long)tp * out_ld + n] = to_out<OutT>(0.f);
            continue;
        }
        const float b = __ldg(bias + n);
        const float* tn = tab + n;
        for (int tp = 0; tp < t_pool; ++tp) {
            float m = -INFINITY;
            for (int q = 0; q < pool; ++q) {
                const int t = tp * pool + q;
                float acc = b;
                for (int j = 0; j < k; ++j) acc += tn[(j * 5 + sc[t + j]) * cout];
                m = fmaxf(m, acc);
            }
            if (relu) m = fmaxf(m, 0.f);
            orow[(long long)tp * out_ld + n] = to_out<OutT>(m);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fp32 accuracy path: implicit GEMM with FFMA.  full[r][n] = act(bias[n] + sum_k A[r*lda + k] W[k][n])
// ---------------------------------------------------------------------------------------------
constexpr int FBM = 64, FBN = 64, FBK = 16;

__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, long long a_elems, int lda,
                                                       long long rows, int K, const float* __restrict__ W,
                                                       int ldw, const float* __restrict__ bias, int relu,
                                                       float* __restrict__ out, int ldo, int n_cols) {
    __shared__ float As[FBK][FBM + 4];
    __shared__ float Bs[FBK][FBN];
    const int tid = threadIdx.x;
    const long long r0 = (long long)blockIdx.x * FBM;
    const int n0 = blockIdx.y * FBN;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int ar = tid >> 2;         // 0..63
    const int ak = (tid & 3) * 4;    // 0,4,8,12
    const int bk = tid >> 4;         // 0..15
    const int bn = (tid & 15) * 4;   // 0..60
    for (int k0 = 0; k0 < K; k0 += FBK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int kk = k0 + ak + i;
            const long long idx = (r0 + ar) * lda + kk;
            As[ak + i][ar] = (kk < K && (r0 + ar) < rows && idx < a_elems) ? __ldg(A + idx) : 0.f;
        }
        {
            const int kk = k0 + bk;
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kk < K && n0 + bn < ldw) w = *reinterpret_cast<const float4*>(W + (long long)kk * ldw + n0 + bn);
            *reinterpret_cast<float4*>(&Bs[bk][bn]) = w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < FBK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long r = r0 + ty * 4 + i;
        if (r >= rows) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= n_cols) continue;
            float v = acc[i][j] + __ldg(bias + n);
            if (relu) v = fmaxf(v, 0.f);
            out[r * ldo + n] = v;
        }
    }
}

// max-pool + compaction of the "full" conv output (rows = seq*t_in + t) into (seq, t_pool, c)
template <typename OutT>
__global__ void __launch_bounds__(256) pool_compact_kernel(const float* __restrict__ full, int ldf, int t_in,
                                                           int pool, int t_pool, int out_pitch, int c_valid, int out_ld,
                                                           long long n_seq, OutT* __restrict__ out) {
    const long long total = n_seq * t_pool * out_ld;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int c = (int)(i % out_ld);
        const long long rt = i / out_ld;
        const int tp = (int)(rt % t_pool);
        const long long s = rt / t_pool;
        float m = 0.f;
        if (c < c_valid) {
            m = -INFINITY;
            for (int q = 0; q < pool; ++q) m = fmaxf(m, full[(s * t_in + (long long)tp * pool + q) * ldf + c]);
        }
        out[(s * out_pitch + tp) * out_ld + c] = to_out<OutT>(m);
    }
}

// ---------------------------------------------------------------------------------------------
// kernel (d): sigmoid + scores
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoid_f32(float x) { return 1.f / (1.f + expf(-x)); }
// logit_score_handler.py:118-124 clamp, then scipy.special.logit in float32: log(p / (1 - p))
__device__ __forceinline__ float clamp_pred(float p) { return p == 0.f ? 1e-24f : (p >= 1.f ? 0.999999f : p); }
__device__ __forceinline__ float logit_f32(float p) { return logf(p / (1.f - p)); }

__global__ void __launch_bounds__(256) sigmoid_kernel(const float* __restrict__ logits, int ldl, long long n, int F,
                                                      float* __restrict__ pred) {
    const long long total = n * F;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / F;
        const int f = (int)(i - r * F);
        pred[i] = sigmoid_f32(logits[r * ldl + f]);
    }
}

// logits_mode 1: `x` holds pre-sigmoid values with row stride ldx; 0: `x` holds probabilities.
__global__ void __launch_bounds__(256) score_kernel(const float* __restrict__ x, int ldx, int x_is_logits,
                                                    int pair_mode, const float* __restrict__ base_pred,
                                                    const int32_t* __restrict__ base_index, long long n_base,
                                                    long long n_out, int F, float* __restrict__ abs_diff,
                                                    float* __restrict__ diff, float* __restrict__ logit,
                                                    float* __restrict__ pred_ref, float* __restrict__ pred_alt) {
    const long long total = n_out * F;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long v = i / F;
        const int f = (int)(i - v * F);
        float ref, alt;
        if (pair_mode == SB_PAIR_INTERLEAVED) {
            const float a = x[(2 * v) * ldx + f], b = x[(2 * v + 1) * ldx + f];
            ref = x_is_logits ? sigmoid_f32(a) : a;
            alt = x_is_logits ? sigmoid_f32(b) : b;
        } else {
            const float b = x[v * ldx + f];
            alt = x_is_logits ? sigmoid_f32(b) : b;
            long long bi = base_index ? (long long)base_index[v] : (n_base == n_out ? v : 0);
            ref = base_pred[bi * F + f];
        }
        if (abs_diff) abs_diff[i] = fabsf(ref - alt);  // np.abs(baseline - batch)
        if (diff) diff[i] = alt - ref;                 // batch - baseline
        if (logit) {
            ref = clamp_pred(ref);
            alt = clamp_pred(alt);
            logit[i] = logit_f32(alt) - logit_f32(ref);
        }
        if (pred_ref) pred_ref[i] = ref;
        if (pred_alt) pred_alt[i] = alt;
    }
}

// ---------------------------------------------------------------------------------------------
// NonStrandSpecific (utils/non_strand_specific_module.py:62-85): the wrapped model also sees the input with
// its channel axis and its position axis reversed; the two outputs are averaged or maxed.
// ---------------------------------------------------------------------------------------------
// rows [r0, r0+m) of a codes source, substitution applied, positions reversed, channel c -> 3-c.  `map[b]` is the
// canonical code whose channel is 3 - channel(b); the unknown code (0.25 in every channel) maps to itself.
struct RcMap { uint8_t m[5]; };
__global__ void __launch_bounds__(256) rc_codes_kernel(const uint8_t* __restrict__ codes, const int32_t* __restrict__ src,
                                                       const int32_t* __restrict__ sub_pos,
                                                       const uint8_t* __restrict__ sub_code, long long r0, long long m,
                                                       int L, RcMap map, uint8_t* __restrict__ out) {
    const long long total = m * L;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / L;
        const int p = (int)(i - r * L);
        const long long row = r0 + r;
        const int q = L - 1 - p;
        uint8_t c = codes[(src ? (long long)src[row] : row) * L + q];
        if (sub_pos && sub_pos[row] == q) c = sub_code ? sub_code[row] : 4;
        out[i] = map.m[c > 4 ? 4 : c];
    }
}
__global__ void __launch_bounds__(256) rc_onehot_kernel(const float* __restrict__ x, long long r0, long long m, int L,
                                                        float* __restrict__ out) {
    const long long total = m * L * 4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int ch = (int)(i & 3);
        const long long rp = i >> 2;
        const long long r = rp / L;
        const int p = (int)(rp - r * L);
        out[i] = x[((r0 + r) * L + (L - 1 - p)) * 4 + (3 - ch)];
    }
}
__global__ void __launch_bounds__(256) strand_combine_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                             long long n, int mode, float* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = mode == 1 ? (a[i] + b[i]) / 2.f : fmaxf(a[i], b[i]);
}

// ---------------------------------------------------------------------------------------------
// BiLSTM cell update for one step of every chain (torch gate order i, f, g, o)
//   unit = (chain, position); chain = (block, c); row(chain, s) = block*G*I + c + I*s
// ---------------------------------------------------------------------------------------------
template <typename HT, typename OutT>
__global__ void __launch_bounds__(256) lstm_cell_kernel(
    const float* __restrict__ gates_f, const float* __restrict__ gates_b, int ld_gates, const float* __restrict__ xw,
    int H, int T, int G, int I, long long m_rows, int step, long long n_units, HT* __restrict__ h_f,
    HT* __restrict__ h_b, int h_ld, float* __restrict__ c_f, float* __restrict__ c_b, OutT* __restrict__ out,
    int out_ld) {
    const long long total = n_units * H * 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int j = (int)(idx % H);
        long long r = idx / H;
        const int dir = (int)(r & 1);
        const long long unit = r >> 1;
        const int p = (int)(unit % T);
        const long long chain = unit / T;
        const long long block = chain / I;
        const int c = (int)(chain - block * I);
        const long long first = block * G * I + c;
        if (first >= m_rows) continue;
        long long len = (m_rows - first + I - 1) / I;
        if (len > G) len = G;
        if (step >= len) continue;
        const long long sidx = dir == 0 ? step : len - 1 - step;
        const long long row = first + I * sidx;
        const float* g = (dir == 0 ? gates_f : gates_b) + unit * ld_gates;
        const float* x = xw + (row * T + p) * (long long)(8 * H) + (long long)dir * 4 * H;
        float pi = x[j], pf = x[H + j], pg = x[2 * H + j], po = x[3 * H + j];
        float* cs = (dir == 0 ? c_f : c_b) + unit * H + j;
        float c_old = 0.f;
        if (step > 0) {
            pi += g[j]; pf += g[H + j]; pg += g[2 * H + j]; po += g[3 * H + j];
            c_old = *cs;
        }
        const float ig = 1.f / (1.f + expf(-pi));
        const float fg = 1.f / (1.f + expf(-pf));
        const float gg = tanhf(pg);
        const float og = 1.f / (1.f + expf(-po));
        const float c_new = fg * c_old + ig * gg;
        const float h_new = og * tanhf(c_new);
        *cs = c_new;
        (dir == 0 ? h_f : h_b)[unit * h_ld + j] = to_out<HT>(h_new);
        out[(row * T + p) * (long long)out_ld + (long long)dir * H + j] = to_out<OutT>(h_new);
    }
}

__global__ void __launch_bounds__(256) zero_u32_kernel(uint32_t* p, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = 0u;
}

// (n, L, 8) bf16 one-hot rows of the model rows (code row + optional substitution); channels 4..7 zero
__global__ void __launch_bounds__(256) codes_onehot8_kernel(const uint8_t* __restrict__ codes, const int32_t* __restrict__ src,
                                                            const int32_t* __restrict__ sub_pos,
                                                            const uint8_t* __restrict__ sub_code, long long row0,
                                                            long long n_rows, int L, uint4* __restrict__ out) {
    const long long total = n_rows * L;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / L;
        const int t = (int)(i - r * L);
        const long long row = row0 + r;
        const long long srow = src ? (long long)src[row] : row;
        int c = __ldg(codes + srow * L + t);
        if (sub_pos && sub_pos[row] == t) c = sub_code ? sub_code[row] : 4;
        uint32_t lo = 0, hi = 0;
        if (c > 3) { lo = 0x3E803E80u; hi = 0x3E803E80u; }
        else if (c == 0) lo = 0x00003F80u;
        else if (c == 1) lo = 0x3F800000u;
        else if (c == 2) hi = 0x00003F80u;
        else hi = 0x3F800000u;
        out[i] = make_uint4(lo, hi, 0u, 0u);
    }
}

// bf16 max-pool + compaction: out[s][tp][c] = max_q full[s][tp*pool + q][c]
__global__ void __launch_bounds__(256) pool_bf16_kernel(const __nv_bfloat16* __restrict__ full, int t_full, int pool,
                                                        int t_pool, int out_pitch, int ld, long long n_seq,
                                                        __nv_bfloat16* __restrict__ out) {
    const long long total = n_seq * t_pool * ld;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int c = (int)(i % ld);
        const long long rt = i / ld;
        const int tp = (int)(rt % t_pool);
        const long long s = rt / t_pool;
        float m = -INFINITY;
        for (int q = 0; q < pool; ++q)
            m = fmaxf(m, __bfloat162float(full[(s * t_full + (long long)tp * pool + q) * ld + c]));
        out[(s * out_pitch + tp) * ld + c] = __float2bfloat16_rn(m);
    }
}

int grid_for(long long total, int device) {
    long long b = (total + 255) / 256;
    const long long cap = (long long)sb_sm_count(device) * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// ---------------------------------------------------------------------------------------------
// weight packing (host)
// ---------------------------------------------------------------------------------------------
uint16_t f32_to_bf16_bits(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);  // NaN
    const uint32_t lsb = (u >> 16) & 1;
    u += 0x7fffu + lsb;  // round to nearest even
    return (uint16_t)(u >> 16);
}

template <typename T>
int upload(T** dev, const std::vector<T>& host) {
    SB_CHECK_CUDA(cudaMalloc(dev, sizeof(T) * (host.size() ? host.size() : 1)));
    if (!host.empty()) SB_CHECK_CUDA(cudaMemcpy(*dev, host.data(), sizeof(T) * host.size(), cudaMemcpyHostToDevice));
    return SB_OK;
}

}  // namespace

extern "C" int sb_net_create(const sb_layer* layers, int n_layers, int L, const uint8_t perm[4], int device,
                             sb_net** out) {
    SB_REQUIRE(layers && out && n_layers >= 2 && L > 0 && perm, "sb_net_create: bad arguments");
    {
        int seen = 0;
        for (int i = 0; i < 4; ++i) {
            SB_REQUIRE(perm[i] < 4, "sb_net_create: perm is not a permutation of 0..3");
            seen |= 1 << perm[i];
        }
        SB_REQUIRE(seen == 15, "sb_net_create: perm is not a permutation of 0..3");
    }
    SB_REQUIRE(layers[0].type == SB_LAYER_CONV && layers[0].cin == 4, "sb_net_create: first layer must be a 4-channel conv");
    SB_CHECK_CUDA(cudaSetDevice(device));
    sb_net* net = new sb_net();
    net->train = nullptr;
    net->capture_stream = nullptr;
    net->side = nullptr;
    net->train_tc = nullptr;
    net->zero_bias_tc = nullptr;
    net->inference_weights_stale = false;
    net->device = device;
    net->L = L;
    net->n_layers = n_layers;
    memcpy(net->perm, perm, 4);
    net->flops_per_seq = 0.0;
    {
        const char* e2 = getenv("SB_TC_IM2COL");
        net->no_im2col = e2 && e2[0] == '0';
        net->no_im2col_run = e2 && e2[0] == '1';
    }
    const char* env = getenv("SB_TC_PER_TAP");
    net->force_per_tap = env && env[0] == '1';
    if (net->force_per_tap) net->no_im2col = true;
    net->timing = false;
    net->rec_group = 64;
    net->strand_mode = 0;
    net->rec_interleave = 1;
    net->has_lstm = false;

    int t = L, c = 4;          // current activation: t rows of c channels per sequence
    bool seen_fc = false;
    bool prev_lstm = false;    // the FC right after a BiLSTM flattens position-major (danQ.py:46-50)
    std::vector<float> lstm_w, lstm_b;
    for (int li = 0; li < n_layers; ++li) {
        const sb_layer& S = layers[li];
        SB_REQUIRE(S.weight && S.bias, "sb_net_create: layer %d has null weights", li);
        NetLayer N;
        memset(&N, 0, sizeof(N));
        N.type = S.type; N.cin = S.cin; N.cout = S.cout; N.relu = S.relu;
        N.pool = S.pool > 1 ? S.pool : 1;
        if (S.type == SB_LAYER_CONV) {
            SB_REQUIRE(!seen_fc, "sb_net_create: conv layer %d after an FC layer", li);
            SB_REQUIRE(S.cin == c, "sb_net_create: layer %d expects %d input channels, has %d", li, S.cin, c);
            SB_REQUIRE(S.k >= 1 && t >= S.k, "sb_net_create: layer %d kernel %d longer than input %d", li, S.k, t);
            N.k = S.k; N.t_in = t; N.t_conv = t - S.k + 1; N.t_pool = N.t_conv / N.pool;
            SB_REQUIRE(N.t_pool >= 1, "sb_net_create: layer %d pools away everything", li);
            net->flops_per_seq += 2.0 * N.t_conv * (double)S.cout * S.cin * S.k;
        } else if (S.type == SB_LAYER_FC) {
            SB_REQUIRE(S.cin == t * c, "sb_net_create: FC layer %d expects %d inputs, previous layer gives %d x %d", li,
                       S.cin, c, t);
            SB_REQUIRE(N.pool == 1, "sb_net_create: FC layer %d cannot pool", li);
            N.k = 1; N.t_in = 1; N.t_conv = 1; N.t_pool = 1;
            if (!seen_fc && !prev_lstm) { N.fc_prev_c = c; N.fc_prev_t = t; }
            net->flops_per_seq += 2.0 * (double)S.cout * S.cin;
        } else if (S.type == SB_LAYER_BILSTM) {
            SB_REQUIRE(!seen_fc && li > 0, "sb_net_create: BiLSTM layer %d must follow a conv layer", li);
            SB_REQUIRE(S.cin == c, "sb_net_create: layer %d expects %d input channels, has %d", li, S.cin, c);
            for (int a = 0; a < 6; ++a) SB_REQUIRE(S.aux[a], "sb_net_create: BiLSTM layer %d lacks aux[%d]", li, a);
            const int H = S.cout;
            N.hidden = H; N.k = 1; N.t_in = t; N.t_conv = t; N.t_pool = t; N.pool = 1; N.relu = 0;
            N.cout = 8 * H;  // the input projection for both directions
            lstm_w.assign((size_t)8 * H * S.cin, 0.f);
            lstm_b.assign((size_t)8 * H, 0.f);
            for (int d = 0; d < 2; ++d) {
                const float* wih = d == 0 ? S.weight : S.aux[2];
                const float* bih = d == 0 ? S.bias : S.aux[3];
                const float* bhh = d == 0 ? S.aux[1] : S.aux[5];
                for (int o = 0; o < 4 * H; ++o) {
                    lstm_b[(size_t)d * 4 * H + o] = bih[o] + bhh[o];
                    for (int cc = 0; cc < S.cin; ++cc)
                        lstm_w[((size_t)d * 4 * H + o) * S.cin + cc] = wih[(size_t)o * S.cin + cc];
                }
            }
            net->flops_per_seq += 2.0 * t * (8.0 * H * S.cin + 8.0 * H * H);
            net->has_lstm = true;
        } else {
            sb_set_error("sb_net_create: unknown layer type %d", S.type);
            delete net;
            return SB_ERR_ARG;
        }
        const bool is_lstm = S.type == SB_LAYER_BILSTM;
        const float* Wsrc = is_lstm ? lstm_w.data() : S.weight;
        const float* Bsrc = is_lstm ? lstm_b.data() : S.bias;
        const int cout_eff = N.cout;

        // ---- logical K ordering shared by both paths: K index = tap * C + channel, where for the FC after
        //      the conv stack "tap" is the position t and "channel" the conv channel (torch flattens c*T + t)
        const bool conv_like = S.type == SB_LAYER_CONV || is_lstm;
        const int kk = is_lstm ? 1 : S.k;
        const int taps = conv_like ? kk : (seen_fc ? 1 : t);
        const int ch = conv_like ? S.cin : (seen_fc ? S.cin : c);
        auto w_at = [&](int o, int tap, int cc) -> float {
            if (conv_like) return Wsrc[((size_t)o * S.cin + cc) * kk + tap];
            if (seen_fc) return Wsrc[(size_t)o * S.cin + cc];
            if (prev_lstm) return Wsrc[(size_t)o * S.cin + (size_t)tap * c + cc];  // flatten index t*C + c
            return Wsrc[(size_t)o * S.cin + (size_t)cc * t + tap];  // flatten index c*T + t
        };

        // ---- fp32 path
        N.k_f32 = taps * ch;
        N.ldw_f32 = round_up(cout_eff, 4);
        {
            std::vector<float> wf((size_t)N.k_f32 * N.ldw_f32, 0.f), bf((size_t)N.ldw_f32, 0.f);
            for (int o = 0; o < cout_eff; ++o) {
                bf[o] = Bsrc[o];
                for (int tap = 0; tap < taps; ++tap)
                    for (int cc = 0; cc < ch; ++cc) wf[((size_t)tap * ch + cc) * N.ldw_f32 + o] = w_at(o, tap, cc);
            }
            int rc = upload(&N.wf, wf); if (rc) return rc;
            rc = upload(&N.biasf, bf); if (rc) return rc;
        }

        // ---- first layer: lookup table over canonical codes (and the unknown code 4 = 0.25 * sum)
        if (li == 0) {
            std::vector<float> lut((size_t)S.k * 5 * S.cout, 0.f);
            for (int j = 0; j < S.k; ++j)
                for (int o = 0; o < S.cout; ++o) {
                    float s4 = 0.f;
                    for (int b = 0; b < 4; ++b) {
                        const float w = S.weight[((size_t)o * 4 + perm[b]) * S.k + j];
                        lut[((size_t)j * 5 + b) * S.cout + o] = w;
                    }
                    // 0.25*w0 + 0.25*w1 + 0.25*w2 + 0.25*w3 in channel order, as the dense conv would add them
                    for (int chn = 0; chn < 4; ++chn) s4 += 0.25f * S.weight[((size_t)o * 4 + chn) * S.k + j];
                    lut[((size_t)j * 5 + 4) * S.cout + o] = s4;
                }
            int rc = upload(&N.lut, lut); if (rc) return rc;
            const char* force_lut = getenv("SB_CONV1_LUT");
            if (sb_conv1_tc_supported(S.k, N.pool, S.cout) && !(force_lut && force_lut[0] == '1')) {
                const int n_pass = (S.cout + 63) / 64;
                std::vector<uint16_t> wp((size_t)n_pass * 64 * 32, 0);
                sb_conv1_pack_weights(S.weight, S.cout, S.k, perm, wp.data(), f32_to_bf16_bits);
                rc = upload(&N.w1_packed, wp); if (rc) return rc;
                std::vector<float> bp((size_t)n_pass * 64, 0.f);
                for (int o = 0; o < S.cout; ++o) bp[o] = S.bias[o];
                rc = upload(&N.bias1_padded, bp); if (rc) return rc;
                // pool-4 first conv: register-accumulator kernel (SB_CONV1_MMA=0 keeps the tcgen05 one)
                const char* use_mma = getenv("SB_CONV1_MMA");
                if (sb_conv1_mma_supported(S.k, N.pool, S.cout, round_up(S.cout, 8)) && !(use_mma && use_mma[0] == '0')) {
                    std::vector<uint32_t> wm(sb_conv1_mma_pack_weights(S.weight, S.cout, S.k, perm, nullptr, f32_to_bf16_bits), 0);
                    sb_conv1_mma_pack_weights(S.weight, S.cout, S.k, perm, wm.data(), f32_to_bf16_bits);
                    rc = upload(&N.w1m_packed, wm); if (rc) return rc;
                    std::vector<float> bm((size_t)((S.cout + 79) / 80) * 80, 0.f);
                    for (int o = 0; o < S.cout; ++o) bm[o] = S.bias[o];
                    rc = upload(&N.bias1m, bm); if (rc) return rc;
                }
            } else if (sb_conv1_long_supported(S.k, N.pool, S.cout) && !(force_lut && force_lut[0] == '1') &&
                       !(getenv("SB_CONV1_LONG") && getenv("SB_CONV1_LONG")[0] == '0')) {
                const int n_pass = (S.cout + 63) / 64;
                std::vector<uint16_t> wp(sb_conv1_long_pack_weights(S.weight, S.cout, S.k, perm, nullptr, f32_to_bf16_bits), 0);
                sb_conv1_long_pack_weights(S.weight, S.cout, S.k, perm, wp.data(), f32_to_bf16_bits);
                rc = upload(&N.w1l_packed, wp); if (rc) return rc;
                std::vector<float> bp((size_t)n_pass * 64, 0.f);
                for (int o = 0; o < S.cout; ++o) bp[o] = S.bias[o];
                rc = upload(&N.bias1_padded, bp); if (rc) return rc;
            } else {
                N.w0_ld = S.k * 8;
                std::vector<uint16_t> w0((size_t)S.cout * N.w0_ld, 0);
                for (int o = 0; o < S.cout; ++o)
                    for (int j = 0; j < S.k; ++j)
                        for (int b = 0; b < 4; ++b)
                            w0[(size_t)o * N.w0_ld + (size_t)j * 8 + b] =
                                f32_to_bf16_bits(S.weight[((size_t)o * 4 + perm[b]) * S.k + j]);
                rc = upload(reinterpret_cast<uint16_t**>(&N.w0_b), w0); if (rc) return rc;
                const int nt = (S.cout + 255) / 256;
                N.w0_n_tiles_n = nt;
                N.w0_block_n = round_up((S.cout + nt - 1) / nt, 16);
                std::vector<float> bb((size_t)nt * N.w0_block_n, 0.f);
                for (int o = 0; o < S.cout; ++o) bb[o] = S.bias[o];
                rc = upload(&N.w0_bias, bb); if (rc) return rc;
            }
        }

        // ---- bf16 / tcgen05 path (every layer but the first)
        N.out_ld_b = round_up(cout_eff, 8);
        if (li > 0) {
            const int a_ld_ch = round_up(ch, 8);   // channel stride of the incoming activation buffer
            N.a_ld = (S.type == SB_LAYER_FC) ? taps * a_ld_ch : a_ld_ch;
            N.k_full = taps * a_ld_ch;
            N.w_ld = round_up(N.k_full, 8);
            std::vector<uint16_t> wb((size_t)cout_eff * N.w_ld, 0);
            for (int o = 0; o < cout_eff; ++o)
                for (int tap = 0; tap < taps; ++tap)
                    for (int cc = 0; cc < ch; ++cc)
                        wb[(size_t)o * N.w_ld + (size_t)tap * a_ld_ch + cc] = f32_to_bf16_bits(w_at(o, tap, cc));
            int rc = upload(reinterpret_cast<uint16_t**>(&N.wb), wb); if (rc) return rc;
            if (S.type == SB_LAYER_CONV) {
                const int cpk = (a_ld_ch + 63) / 64;
                N.k_full_tap = taps * cpk * 64;
                N.w_ld_tap = N.k_full_tap;
                std::vector<uint16_t> wt((size_t)cout_eff * N.w_ld_tap, 0);
                for (int o = 0; o < cout_eff; ++o)
                    for (int tap = 0; tap < taps; ++tap)
                        for (int cc = 0; cc < ch; ++cc)
                            wt[(size_t)o * N.w_ld_tap + (size_t)tap * cpk * 64 + cc] = f32_to_bf16_bits(w_at(o, tap, cc));
                rc = upload(reinterpret_cast<uint16_t**>(&N.wb_tap), wt); if (rc) return rc;
            }
            // N tiling: fewest tiles of <= 256 columns, each a multiple of 16
            const int nt = (cout_eff + 255) / 256;
            N.n_tiles_n = nt;
            N.block_n = round_up((cout_eff + nt - 1) / nt, 16);
            std::vector<float> bb((size_t)N.n_tiles_n * N.block_n, 0.f);
            for (int o = 0; o < cout_eff; ++o) bb[o] = Bsrc[o];
            rc = upload(&N.biasb, bb); if (rc) return rc;
        }

        if (is_lstm) {
            const int H = N.hidden;
            N.hidden_ld = round_up(H, 8);
            N.out_ld_b = round_up(2 * H, 8);
            const int nt = (4 * H + 255) / 256;
            N.hh_n_tiles_n = nt;
            N.hh_block_n = round_up((4 * H + nt - 1) / nt, 16);
            std::vector<float> zb((size_t)std::max(nt * N.hh_block_n, round_up(4 * H, 4)), 0.f);
            int rc = upload(&N.zero_bias, zb); if (rc) return rc;
            for (int d = 0; d < 2; ++d) {
                const float* whh = d == 0 ? S.aux[0] : S.aux[4];
                const int ldw = round_up(4 * H, 4);
                std::vector<float> wf((size_t)H * ldw, 0.f);
                std::vector<uint16_t> wb((size_t)4 * H * N.hidden_ld, 0);
                for (int o = 0; o < 4 * H; ++o)
                    for (int k2 = 0; k2 < H; ++k2) {
                        wf[(size_t)k2 * ldw + o] = whh[(size_t)o * H + k2];
                        wb[(size_t)o * N.hidden_ld + k2] = f32_to_bf16_bits(whh[(size_t)o * H + k2]);
                    }
                rc = upload(&N.whh_f[d], wf); if (rc) return rc;
                rc = upload(reinterpret_cast<uint16_t**>(&N.whh_b[d]), wb); if (rc) return rc;
                if (sb_lstm_persistent_supported(H)) {
                    std::vector<uint16_t> wp((size_t)4 * H * N.hidden_ld, 0);
                    for (int n = 0; n < 4 * H; ++n) {
                        const int o = sb_lstm_perm_row(n, H);
                        for (int k2 = 0; k2 < H; ++k2)
                            wp[(size_t)n * N.hidden_ld + k2] = f32_to_bf16_bits(whh[(size_t)o * H + k2]);
                    }
                    rc = upload(reinterpret_cast<uint16_t**>(&N.whh_perm[d]), wp); if (rc) return rc;
                }
            }
            if (sb_lstm_persistent_supported(H) && li > 0) {
                // input projection with the same row order per direction (lstm_w / lstm_b: 8H rows, bias = b_ih + b_hh)
                const int a_ld_ch = round_up(S.cin, 8);
                std::vector<uint16_t> wp((size_t)8 * H * N.w_ld, 0);
                std::vector<float> bp((size_t)N.n_tiles_n * N.block_n, 0.f);
                for (int d = 0; d < 2; ++d)
                    for (int n = 0; n < 4 * H; ++n) {
                        const int o = d * 4 * H + sb_lstm_perm_row(n, H);
                        bp[(size_t)d * 4 * H + n] = Bsrc[o];
                        for (int cc = 0; cc < S.cin; ++cc)
                            wp[((size_t)d * 4 * H + n) * N.w_ld + cc] = f32_to_bf16_bits(w_at(o, 0, cc));
                    }
                (void)a_ld_ch;
                rc = upload(reinterpret_cast<uint16_t**>(&N.wb_perm), wp); if (rc) return rc;
                rc = upload(&N.biasb_perm, bp); if (rc) return rc;
            }
        }
        net->layers.push_back(N);
        prev_lstm = false;
        if (S.type == SB_LAYER_CONV) { t = N.t_pool; c = S.cout; }
        else if (is_lstm) { c = 2 * N.hidden; prev_lstm = true; }
        else { seen_fc = true; t = 1; c = S.cout; }
    }
    SB_REQUIRE(seen_fc, "sb_net_create: network must end in FC layers");
    for (size_t li = 0; li < net->layers.size(); ++li) {
        NetLayer& N = net->layers[li];
        N.pitch_in = N.type == SB_LAYER_FC ? 1 : round_up(N.t_in, N.pool);
    }
    for (size_t li = 0; li < net->layers.size(); ++li) {
        NetLayer& N = net->layers[li];
        const bool next_conv = li + 1 < net->layers.size() && net->layers[li + 1].type != SB_LAYER_FC;
        N.pitch_out = next_conv ? net->layers[li + 1].pitch_in : N.t_pool;
    }
    net->n_out = c;
    *out = net;
    return SB_OK;
}

static void free_train_states(sb_net* net);   // sb_train_tc.inl (the training structs are defined there)

extern "C" int sb_net_destroy(sb_net* net) {
    if (!net) return SB_OK;
    cudaSetDevice(net->device);
    free_train_states(net);
    for (auto& v : net->ev)
        for (cudaEvent_t e : v) cudaEventDestroy(e);
    for (auto& g : net->lstm_graphs) cudaGraphExecDestroy(g.exec);
    if (net->capture_stream) cudaStreamDestroy(net->capture_stream);
    for (auto& N : net->layers) {
        cudaFree(N.wf); cudaFree(N.biasf); cudaFree(N.wb); cudaFree(N.wb_tap); cudaFree(N.biasb); cudaFree(N.lut); cudaFree(N.w1_packed); cudaFree(N.w1l_packed); cudaFree(N.bias1_padded); cudaFree(N.w1m_packed); cudaFree(N.bias1m); cudaFree(N.w0_b); cudaFree(N.w0_bias);
        cudaFree(N.whh_f[0]); cudaFree(N.whh_f[1]); cudaFree(N.whh_b[0]); cudaFree(N.whh_b[1]); cudaFree(N.zero_bias);
        cudaFree(N.wb_perm); cudaFree(N.biasb_perm); cudaFree(N.whh_perm[0]); cudaFree(N.whh_perm[1]);
    }
    if (net->side) {
        cudaStreamDestroy(net->side);
        cudaEventDestroy(net->ev_fork);
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(net->ev_c1[i]); cudaEventDestroy(net->ev_done[i]); }
    }
    delete net;
    return SB_OK;
}

extern "C" int sb_net_set_recurrence(sb_net* net, int group, int interleave) {
    SB_REQUIRE(net, "sb_net_set_recurrence: null handle");
    SB_REQUIRE(group >= 1 && (interleave == 1 || interleave == 2), "sb_net_set_recurrence: bad group %d / interleave %d",
               group, interleave);
    net->rec_group = group;
    net->rec_interleave = interleave;
    return SB_OK;
}

extern "C" int sb_net_set_strand_mode(sb_net* net, int mode) {
    SB_REQUIRE(net, "sb_net_set_strand_mode: null handle");
    SB_REQUIRE(mode >= 0 && mode <= 2, "sb_net_set_strand_mode: mode %d is not 0 (off), 1 (mean) or 2 (max)", mode);
    net->strand_mode = mode;
    return SB_OK;
}

extern "C" int sb_net_set_timing(sb_net* net, int enable) {
    SB_REQUIRE(net, "sb_net_set_timing: null handle");
    net->timing = enable != 0;
    for (auto& u : net->ev_used) u = 0;
    return SB_OK;
}

extern "C" int sb_net_n_layers(const sb_net* net) { return net ? (int)net->layers.size() : -1; }

extern "C" double sb_net_layer_flops_per_seq(const sb_net* net, int layer) {
    if (!net || layer < 0 || layer >= (int)net->layers.size()) return 0.0;
    const NetLayer& N = net->layers[(size_t)layer];
    return 2.0 * N.t_conv * (double)N.cout * N.cin * (N.type == SB_LAYER_CONV ? N.k : 1);
}

extern "C" int sb_net_layer_time(sb_net* net, int layer, double* total_ms, int* launches) {
    SB_REQUIRE(net && total_ms && launches, "sb_net_layer_time: null argument");
    SB_REQUIRE(layer >= 0 && layer < (int)net->layers.size(), "sb_net_layer_time: bad layer %d", layer);
    *total_ms = 0.0;
    *launches = 0;
    if ((size_t)layer >= net->ev.size()) return SB_OK;
    const auto& v = net->ev[(size_t)layer];
    const size_t u = net->ev_used[(size_t)layer];
    for (size_t i = 0; i + 1 < u; i += 2) {
        SB_CHECK_CUDA(cudaEventSynchronize(v[i + 1]));
        float ms = 0.f;
        SB_CHECK_CUDA(cudaEventElapsedTime(&ms, v[i], v[i + 1]));
        *total_ms += ms;
        *launches += 1;
    }
    return SB_OK;
}

extern "C" int sb_net_n_outputs(const sb_net* net) { return net ? net->n_out : -1; }
extern "C" double sb_net_flops_per_seq(const sb_net* net) { return net ? net->flops_per_seq : 0.0; }

namespace {

int chunk_rows(int precision) {
    if (precision != SB_PREC_BF16) return 256;
    const char* e = getenv("SB_CHUNK");       // tuning knob: rows per chunk of the bf16 path
    const int v = e ? atoi(e) : 0;
    return v >= 256 ? (v & ~1) : 8192;
}

// bytes of the two ping-pong activation buffers, the fp32 "full" buffer and the logits buffer
struct WsPlan { size_t act, full, logits, lstm, strand, c1, total; long long n_units; };

// the one-launch recurrence (sb_lstm.cu) runs this layer at this precision
bool lstm_persistent(const NetLayer& N, int precision) {
    const char* stepwise = getenv("SB_LSTM_STEPWISE");
    return precision == SB_PREC_BF16 && N.wb_perm && !(stepwise && stepwise[0] == '1');
}

long long chunk_for(const sb_net* net, int precision, bool generic) {
    long long c = generic ? 256 : chunk_rows(precision);
    // first conv one chunk ahead (conv1_ahead): smaller chunks leave less of it exposed at the start of a call
    if (!generic && precision == SB_PREC_BF16 && net->layers[0].w1m_packed && !net->strand_mode && !net->has_lstm &&
        !getenv("SB_CHUNK")) {
        // ... and sized so that the first FC layer (few, long-K tiles) fills the SMs exactly: 128 rows x (SMs / its N tiles),
        // DeepSEA on 148 SMs: 37 x 128 = 4736 rows = 148 fc1 tiles, 120 conv2 and 53 conv3 tiles per SM
        c = 4096;
        for (const NetLayer& N : net->layers)
            if (N.type == SB_LAYER_FC && N.n_tiles_n > 0) {
                const long long per = sb_sm_count(net->device) / N.n_tiles_n;
                if (per > 0) {
                    long long cc = per * 128;
                    while (cc < 3072) cc *= 2;
                    while (cc > 6144 && cc % 256 == 0) cc /= 2;
                    c = cc;
                }
                break;
            }
    }
    if (net->has_lstm) {  // chunks must hold whole recurrence blocks
        const long long blk = (long long)net->rec_group * net->rec_interleave;
        long long cap = 2048;
        for (const NetLayer& N : net->layers)
            if (N.type == SB_LAYER_BILSTM && lstm_persistent(N, precision) && !generic) {
                // the persistent kernel's time does not depend on its grid size while every CTA (128 units of one
                // direction) has its own SM: take as many rows per launch as one wave holds
                const long long units = 128ll * (sb_sm_count(net->device) / 2);
                const long long chains = units / N.t_in;
                cap = std::max<long long>(cap, chains * net->rec_group);
            }
        if (c > cap) c = cap;
        c = c / blk * blk;
        if (c < blk) c = blk;
    }
    return c;
}

// The register-accumulator first conv (sb_conv1m.cu) needs neither tensor memory nor much shared memory, so its CTAs
// fit on an SM next to a tc_layer_kernel CTA: with several chunks to do, the first conv of chunk k+1 runs on a side
// stream while the tcgen05 layers of chunk k run, into one of two extra buffers (SB_CONV1_AHEAD=0 disables it).
bool conv1_ahead(const sb_net* net, long long n, int precision, bool generic) {
    const char* e = getenv("SB_CONV1_AHEAD");
    return precision == SB_PREC_BF16 && !generic && net->layers[0].w1m_packed && !net->strand_mode && !net->has_lstm &&
           n > chunk_for(net, precision, generic) && !(e && e[0] == '0');
}

WsPlan plan_ws(const sb_net* net, long long n, int precision, bool generic) {
    const long long cr = chunk_for(net, precision, generic);
    const long long ch = n < cr ? n : cr;
    size_t act = 0, full = 0;
    const size_t es = precision == SB_PREC_BF16 ? 2 : 4;
    for (size_t li = 0; li < net->layers.size(); ++li) {
        const NetLayer& N = net->layers[li];
        const int ld = precision == SB_PREC_BF16 ? N.out_ld_b : (N.type == SB_LAYER_BILSTM ? 2 * N.hidden : N.cout);
        const size_t bytes = (size_t)ch * (precision == SB_PREC_BF16 ? N.pitch_out : N.t_pool) * ld * es;
        if (bytes > act) act = bytes;
        if (precision == SB_PREC_FP32 && li > 0) {
            const size_t fb = (size_t)ch * N.t_in * N.cout * 4;
            if (fb > full) full = fb;
        }
    }
    if (precision == SB_PREC_BF16 && !generic && net->layers[0].w0_b) {
        const NetLayer& N0 = net->layers[0];
        const size_t fb = (size_t)ch * net->L * 8 * 2 + 1024 + (size_t)ch * N0.t_conv * N0.out_ld_b * 2;
        if (fb > full) full = fb;
    }
    // generic-input first conv also runs through the fp32 "full" path
    if (generic) {
        const NetLayer& N0 = net->layers[0];
        const size_t fb = (size_t)ch * N0.t_in * N0.cout * 4;
        if (fb > full) full = fb;
    }
    size_t lstm = 0;
    long long n_units = 0;
    for (size_t li = 0; li < net->layers.size(); ++li) {
        const NetLayer& N = net->layers[li];
        if (N.type != SB_LAYER_BILSTM) continue;
        const long long blk = (long long)net->rec_group * net->rec_interleave;
        const long long chains = ((ch + blk - 1) / blk) * net->rec_interleave;
        n_units = chains * N.t_in;
        const size_t xw = (size_t)ch * N.t_in * 8 * N.hidden * (lstm_persistent(N, precision) ? 2 : 4);
        if (xw > full) full = xw;
        // gates (2 x units x 4H fp32) + h (2 x units x H_ld, 4 B each) + c (2 x units x H fp32)
        lstm = (size_t)n_units * (2 * 4 * N.hidden * 4 + 2 * N.hidden_ld * 4 + 2 * N.hidden * 4) + 4096;
    }
    auto al = [](size_t b) { return (b + 1023) / 1024 * 1024 + 1024; };
    WsPlan p;
    p.lstm = al(lstm);
    p.n_units = n_units;
    p.act = al(act);
    p.full = al(full);
    p.logits = al((size_t)ch * round_up(net->n_out, 4) * 4);
    // NonStrandSpecific: probabilities of both strands + the reverse-complemented rows of the chunk
    p.strand = net->strand_mode ? al(2 * (size_t)ch * net->n_out * 4) + al((size_t)ch * net->L * (generic ? 16 : 1)) : 0;
    p.c1 = 0;
    if (conv1_ahead(net, n, precision, generic)) {
        const NetLayer& N0 = net->layers[0];
        p.c1 = al((size_t)ch * N0.pitch_out * N0.out_ld_b * 2);
    }
    p.total = 2 * p.act + p.full + p.logits + p.lstm + p.strand + 2 * p.c1 + 4096;
    return p;
}

struct RowSource {
    const uint8_t* codes; const int32_t* src; const int32_t* sub_pos; const uint8_t* sub_code;  // codes input
    const float* onehot;                                                                        // generic input
};

void time_begin(sb_net* net, size_t li, cudaStream_t s) {
    if (!net->timing) return;
    if (net->ev.size() < net->layers.size()) { net->ev.resize(net->layers.size()); net->ev_used.assign(net->layers.size(), 0); }
    auto& v = net->ev[li];
    size_t& u = net->ev_used[li];
    if (u + 2 > v.size()) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        v.push_back(a);
        v.push_back(b);
    }
    cudaEventRecord(v[u], s);
}
void time_end(sb_net* net, size_t li, cudaStream_t s) {
    if (!net->timing) return;
    cudaEventRecord(net->ev[li][net->ev_used[li] + 1], s);
    net->ev_used[li] += 2;
}

// One BiLSTM layer over m rows (T positions each): input projection as one GEMM, then G sequential
// steps of (recurrent GEMM per direction + cell update) batched over every (chain, position) unit.
int run_lstm_body(sb_net* net, const NetLayer& N, const void* act_in, void* act_out, long long m, int precision,
                  float* xw, uint8_t* lstm_ws, cudaStream_t stream) {
    const bool bf = precision == SB_PREC_BF16;
    const int dev = net->device;
    const int H = N.hidden, T = N.t_in, G = net->rec_group, I = net->rec_interleave;
    const long long rows = m * T;
    const long long blk = (long long)G * I;
    const long long chains = ((m + blk - 1) / blk) * I;
    const long long U = chains * T;
    {
        // One persistent kernel for all steps (sb_lstm.cu); SB_LSTM_STEPWISE=1 selects the per-step launches below.
        if (lstm_persistent(N, precision)) {
            // projection (+ bias) as bf16, gate columns in the recurrent kernel's tile order
            SbTcLayer P;
            memset(&P, 0, sizeof(P));
            P.a = act_in; P.rows_a = rows; P.a_ld = N.a_ld; P.k_full = N.k_full;
            P.w = N.wb_perm; P.n_rows_w = 8 * H; P.w_ld = N.w_ld;
            P.bias = N.biasb_perm; P.block_n = N.block_n; P.n_tiles_n = N.n_tiles_n;
            P.t_in = 1; P.t_conv = 1; P.pool = 1; P.relu = 0; P.out_f32 = 0; P.out = xw; P.out_ld = 8 * H;
            int rc = sb_tc_launch(P, dev, stream);
            if (rc) return rc;
            return sb_lstm_persistent_launch(N.whh_perm[0], N.whh_perm[1], N.hidden_ld, xw, act_out, N.out_ld_b, m, H, T, G,
                                             I, dev, stream);
        }
    }
    // ---- input projection for both directions: xw[row*T + p][dir*4H + gate*H + j]
    if (bf) {
        SbTcLayer P;
        memset(&P, 0, sizeof(P));
        P.a = act_in; P.rows_a = rows; P.a_ld = N.a_ld; P.k_full = N.k_full;
        P.w = N.wb; P.n_rows_w = 8 * H; P.w_ld = N.w_ld;
        P.bias = N.biasb; P.block_n = N.block_n; P.n_tiles_n = N.n_tiles_n;
        P.t_in = 1; P.t_conv = 1; P.pool = 1; P.relu = 0; P.out_f32 = 1; P.out = xw; P.out_ld = 8 * H;
        int rc = sb_tc_launch(P, dev, stream);
        if (rc) return rc;
    } else {
        dim3 grid((unsigned)((rows + FBM - 1) / FBM), (unsigned)((8 * H + FBN - 1) / FBN));
        gemm_f32_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(act_in), rows * N.cin, N.cin, rows,
                                                  N.k_f32, N.wf, N.ldw_f32, N.biasf, 0, xw, 8 * H, 8 * H);
        SB_COUNT_LAUNCH();
        SB_CHECK_LAUNCH();
    }
    // ---- state buffers
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(lstm_ws) + 255) & ~(uintptr_t)255);
    float* gates[2];
    gates[0] = reinterpret_cast<float*>(base);
    gates[1] = gates[0] + U * 4 * H;
    uint8_t* hb = reinterpret_cast<uint8_t*>(gates[1] + U * 4 * H);
    const size_t h_bytes = (size_t)U * N.hidden_ld * 4;   // sized for fp32, bf16 uses half
    void* h[2] = {hb, hb + h_bytes};
    float* cst[2];
    cst[0] = reinterpret_cast<float*>(hb + 2 * h_bytes);
    cst[1] = cst[0] + U * H;
    zero_u32_kernel<<<grid_for((long long)(2 * h_bytes / 4), dev), 256, 0, stream>>>(reinterpret_cast<uint32_t*>(hb),
                                                                                     (long long)(2 * h_bytes / 4));
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    long long steps = (m + I - 1) / I;
    if (steps > G) steps = G;
    const int ld4 = round_up(4 * H, 4);
    for (int step = 0; step < (int)steps; ++step) {
        if (step > 0) {
            for (int d = 0; d < 2; ++d) {
                if (bf) {
                    SbTcLayer R;
                    memset(&R, 0, sizeof(R));
                    R.a = h[d]; R.rows_a = U; R.a_ld = N.hidden_ld; R.k_full = N.hidden_ld;
                    R.w = N.whh_b[d]; R.n_rows_w = 4 * H; R.w_ld = N.hidden_ld;
                    R.bias = N.zero_bias; R.block_n = N.hh_block_n; R.n_tiles_n = N.hh_n_tiles_n;
                    R.t_in = 1; R.t_conv = 1; R.pool = 1; R.relu = 0; R.out_f32 = 1; R.out = gates[d]; R.out_ld = 4 * H;
                    int rc = sb_tc_launch(R, dev, stream);
                    if (rc) return rc;
                } else {
                    dim3 grid((unsigned)((U + FBM - 1) / FBM), (unsigned)((4 * H + FBN - 1) / FBN));
                    gemm_f32_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(h[d]), U * N.hidden_ld,
                                                              N.hidden_ld, U, H, N.whh_f[d], ld4, N.zero_bias, 0,
                                                              gates[d], 4 * H, 4 * H);
                    SB_COUNT_LAUNCH();
                    SB_CHECK_LAUNCH();
                }
            }
        }
        const int g = grid_for(U * H * 2, dev);
        if (bf)
            lstm_cell_kernel<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, stream>>>(
                gates[0], gates[1], 4 * H, xw, H, T, G, I, m, step, U, reinterpret_cast<__nv_bfloat16*>(h[0]),
                reinterpret_cast<__nv_bfloat16*>(h[1]), N.hidden_ld, cst[0], cst[1],
                reinterpret_cast<__nv_bfloat16*>(act_out), N.out_ld_b);
        else
            lstm_cell_kernel<float, float><<<g, 256, 0, stream>>>(
                gates[0], gates[1], 4 * H, xw, H, T, G, I, m, step, U, reinterpret_cast<float*>(h[0]),
                reinterpret_cast<float*>(h[1]), N.hidden_ld, cst[0], cst[1], reinterpret_cast<float*>(act_out), 2 * H);
        SB_COUNT_LAUNCH();
        SB_CHECK_LAUNCH();
    }
    return SB_OK;
}

// The recurrence is ~190 small launches per chunk; issued one by one the host cannot keep the GPU fed, so
// the whole layer is captured once per (buffers, rows, grouping) into a CUDA graph and replayed.
int run_lstm(sb_net* net, const NetLayer& N, const void* act_in, void* act_out, long long m, int precision,
             float* xw, uint8_t* lstm_ws, cudaStream_t stream) {
    const char* no_graph = getenv("SB_NO_GRAPH");
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &cs);
    if ((no_graph && no_graph[0] == '1') || cs != cudaStreamCaptureStatusNone)
        return run_lstm_body(net, N, act_in, act_out, m, precision, xw, lstm_ws, stream);
    LstmGraphKey key = {act_in, act_out, xw, lstm_ws, m, net->rec_group, net->rec_interleave, precision, &N};
    for (auto& e : net->lstm_graphs)
        if (memcmp(&e.key, &key, sizeof(key)) == 0) {
            SB_CHECK_CUDA(cudaGraphLaunch(e.exec, stream));
            (void)__atomic_fetch_add(&g_sb_launches, (unsigned long long)e.n_launches, __ATOMIC_RELAXED);
            return SB_OK;
        }
    // warm the lazily initialised pieces (function attributes, driver entry point) outside the capture
    if (net->lstm_graphs.empty()) {
        int rc = run_lstm_body(net, N, act_in, act_out, m, precision, xw, lstm_ws, stream);
        if (rc) return rc;
    }
    const unsigned long long before = g_sb_launches;
    // the caller's stream may be the legacy default stream, which cannot be captured: record on a private
    // stream, replay on the caller's
    if (!net->capture_stream) SB_CHECK_CUDA(cudaStreamCreateWithFlags(&net->capture_stream, cudaStreamNonBlocking));
    cudaStream_t cap = net->capture_stream;
    SB_CHECK_CUDA(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
    int rc = run_lstm_body(net, N, act_in, act_out, m, precision, xw, lstm_ws, cap);
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamEndCapture(cap, &graph);
    if (rc != SB_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    SB_CHECK_CUDA(ce);
    LstmGraph g;
    memset(&g, 0, sizeof(g));
    g.key = key;
    g.n_launches = g_sb_launches - before;
    SB_CHECK_CUDA(cudaGraphInstantiate(&g.exec, graph, 0));
    cudaGraphDestroy(graph);
    if (net->lstm_graphs.size() >= 16) {   // bounded cache
        cudaGraphExecDestroy(net->lstm_graphs.front().exec);
        net->lstm_graphs.erase(net->lstm_graphs.begin());
    }
    net->lstm_graphs.push_back(g);
    SB_CHECK_CUDA(cudaGraphLaunch(g.exec, stream));
    return SB_OK;
}

// Runs rows [r0, r0+m) through the network; leaves fp32 logits (row stride ldl) in ws logits buffer.
int run_chunk(sb_net* net, const RowSource& in, long long r0, long long m, int precision, uint8_t* ws,
              const WsPlan& plan, float** logits_out, int* ldl_out, cudaStream_t stream,
              const SbScoreEpilogue* fused = nullptr, int phase = 0, void* l0_buf = nullptr) {
    // phase 0: the whole network; 1: the first conv only, into l0_buf; 2: every other layer, reading l0_buf
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~(uintptr_t)1023);
    void* act[2] = {base, base + plan.act};
    float* full = reinterpret_cast<float*>(base + 2 * plan.act);
    float* logits = reinterpret_cast<float*>(base + 2 * plan.act + plan.full);
    const int ldl = round_up(net->n_out, 4);
    const bool bf = precision == SB_PREC_BF16;
    const int dev = net->device;
    int cur = 0;

    // ---- layer 0
    if (phase != 2) {
    time_begin(net, 0, stream);
    {
        const NetLayer& N = net->layers[0];
        const int out_ld = bf ? N.out_ld_b : N.cout;
        if (in.codes) {
            const size_t smem = (size_t)N.k * 5 * N.cout * sizeof(float) + (size_t)round_up(net->L, 16);
            SB_REQUIRE(smem <= 220 * 1024, "first-layer lookup table (%zu B) does not fit shared memory", smem);
            int threads = round_up(out_ld < 512 ? out_ld : 512, 32);
            if (bf && N.w1m_packed) {
                int rc = sb_conv1_mma_launch(in.codes, in.src, in.sub_pos, in.sub_code, r0, m, net->L, N.k, N.pool,
                                             N.t_pool, N.pitch_out, N.cout, out_ld, N.relu, N.w1m_packed, N.bias1m,
                                             phase == 1 ? l0_buf : act[cur], dev, stream, phase == 1 && r0 > 0 ? 1 : 2);
                if (rc) return rc;
            } else if (bf && N.w1_packed) {
                int rc = sb_conv1_tc_launch(in.codes, in.src, in.sub_pos, in.sub_code, r0, m, net->L, N.k, N.pool,
                                            N.t_pool, N.pitch_out, N.cout, out_ld, N.relu, N.w1_packed,
                                            N.bias1_padded, act[cur], dev, stream);
                if (rc) return rc;
            } else if (bf && N.w1l_packed) {
                int rc = sb_conv1_long_launch(in.codes, in.src, in.sub_pos, in.sub_code, r0, m, net->L, N.k, N.pool, N.t_pool,
                                              N.pitch_out, N.cout, out_ld, N.relu, N.w1l_packed, N.bias1_padded, act[cur],
                                              dev, stream);
                if (rc) return rc;
            } else if (bf && N.w0_b) {
                // one-hot (n, L, 8) -> generic tcgen05 layer kernel (unpooled) -> bf16 pool
                uint4* oh = reinterpret_cast<uint4*>(full);
                codes_onehot8_kernel<<<grid_for(m * net->L, dev), 256, 0, stream>>>(in.codes, in.src, in.sub_pos, in.sub_code,
                                                                                  r0, m, net->L, oh);
                SB_COUNT_LAUNCH();
                SB_CHECK_LAUNCH();
                __nv_bfloat16* fullb = reinterpret_cast<__nv_bfloat16*>(
                    reinterpret_cast<uint8_t*>(full) + (((size_t)m * net->L * 16 + 1023) / 1024) * 1024);
                SbTcLayer T0;
                memset(&T0, 0, sizeof(T0));
                T0.a = oh; T0.rows_a = m * (long long)net->L; T0.a_ld = 8; T0.k_full = N.k * 8;
                T0.w = N.w0_b; T0.n_rows_w = N.cout; T0.w_ld = N.w0_ld;
                T0.bias = N.w0_bias; T0.block_n = N.w0_block_n; T0.n_tiles_n = N.w0_n_tiles_n;
                T0.t_in = net->L; T0.t_conv = N.t_conv; T0.pool = 1; T0.relu = N.relu; T0.out_f32 = 0;
                T0.out = fullb; T0.out_ld = out_ld; T0.out_pitch = N.t_conv;
                int rc = sb_tc_launch(T0, dev, stream);
                if (rc) return rc;
                const long long total = m * N.t_pool * out_ld;
                pool_bf16_kernel<<<grid_for(total, dev), 256, 0, stream>>>(fullb, N.t_conv, N.pool, N.t_pool, N.pitch_out,
                                                                           out_ld, m, reinterpret_cast<__nv_bfloat16*>(act[cur]));
                SB_COUNT_LAUNCH();
                SB_CHECK_LAUNCH();
            } else if (bf) {
                SB_CHECK_CUDA(cudaFuncSetAttribute(conv_codes_kernel<__nv_bfloat16>,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                conv_codes_kernel<__nv_bfloat16><<<(unsigned)m, threads, smem, stream>>>(
                    in.codes, in.src, in.sub_pos, in.sub_code, r0, net->L, N.k, N.cout, out_ld, N.relu, N.pool,
                    N.t_pool, N.pitch_out, N.lut, N.biasf, reinterpret_cast<__nv_bfloat16*>(act[cur]));
            } else {
                SB_CHECK_CUDA(cudaFuncSetAttribute(conv_codes_kernel<float>,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                conv_codes_kernel<float><<<(unsigned)m, threads, smem, stream>>>(
                    in.codes, in.src, in.sub_pos, in.sub_code, r0, net->L, N.k, N.cout, out_ld, N.relu, N.pool,
                    N.t_pool, N.t_pool, N.lut, N.biasf, reinterpret_cast<float*>(act[cur]));
            }
            if (!(bf && (N.w1_packed || N.w1l_packed || N.w0_b))) {
                SB_COUNT_LAUNCH();
                SB_CHECK_LAUNCH();
            }
        } else {
            // generic (n, L, 4) fp32 input: dense fp32 conv, then pool/compact (+ convert)
            const float* x = in.onehot + r0 * net->L * 4;
            const long long rows = m * N.t_in;
            dim3 grid((unsigned)((rows + FBM - 1) / FBM), (unsigned)((N.cout + FBN - 1) / FBN));
            gemm_f32_kernel<<<grid, 256, 0, stream>>>(x, rows * 4, 4, rows, N.k_f32, N.wf, N.ldw_f32, N.biasf, N.relu,
                                                      full, N.cout, N.cout);
            SB_COUNT_LAUNCH();
            SB_CHECK_LAUNCH();
            const long long total = m * N.t_pool * out_ld;
            if (bf)
                pool_compact_kernel<__nv_bfloat16><<<grid_for(total, dev), 256, 0, stream>>>(
                    full, N.cout, N.t_in, N.pool, N.t_pool, N.pitch_out, N.cout, out_ld, m,
                    reinterpret_cast<__nv_bfloat16*>(act[cur]));
            else
                pool_compact_kernel<float><<<grid_for(total, dev), 256, 0, stream>>>(
                    full, N.cout, N.t_in, N.pool, N.t_pool, N.t_pool, N.cout, out_ld, m, reinterpret_cast<float*>(act[cur]));
            SB_COUNT_LAUNCH();
            SB_CHECK_LAUNCH();
        }
    }

    time_end(net, 0, stream);
    }
    if (phase == 1) return SB_OK;

    // ---- remaining layers
    for (size_t li = 1; li < net->layers.size(); ++li) {
        const NetLayer& N = net->layers[li];
        const bool last = li + 1 == net->layers.size();
        time_begin(net, li, stream);
        if (N.type == SB_LAYER_BILSTM) {
            uint8_t* lstm_ws = base + 2 * plan.act + plan.full + plan.logits;
            int rc = run_lstm(net, N, act[cur], act[cur ^ 1], m, precision, full, lstm_ws, stream);
            if (rc) return rc;
            time_end(net, li, stream);
            cur ^= 1;
            continue;
        }
        if (bf) {
            SbTcLayer T;
            memset(&T, 0, sizeof(T));
            T.a = (phase == 2 && li == 1) ? l0_buf : act[cur];
            T.rows_a = m * (long long)N.pitch_in;
            T.a_ld = N.a_ld;
            T.bias = N.biasb; T.block_n = N.block_n; T.n_tiles_n = N.n_tiles_n;
            T.n_rows_w = N.cout;
            T.t_in = N.pitch_in; T.t_conv = N.t_conv; T.pool = N.pool; T.relu = N.relu;
            T.out_pitch = N.pitch_out;
            T.out_f32 = last ? (fused ? 2 : 1) : 0;
            T.out = last ? (void*)logits : act[cur ^ 1];
            T.out_ld = last ? ldl : N.out_ld_b;
            // conv layers: im2col tensor map over the run view (valid window starts only, no K padding) -> per-tap im2col
            // map -> overlapping-stride view -> per-tap boxes, each the fallback of the one before if the driver refuses
            // the tensor map.  The run view pays whenever rows are dropped (DeepSEA conv2 240 / 248, conv3 53 / 60); the
            // per-tap map pads the channels to 64 per tap (conv3: 512 / 480), so it is only used where rows x K still
            // shrinks; long narrow layers (HeartENN: 986 / 993 rows) keep the overlapping view (r02 sweep, same GPU)
            bool per_tap = net->force_per_tap && N.type == SB_LAYER_CONV;
            int im2col = 0;
            if (N.type == SB_LAYER_CONV && N.k > 1 && !net->no_im2col) {
                const double rows_ratio = (double)((N.t_conv / N.pool) * N.pool) / (double)N.pitch_in;
                if (!net->no_im2col_run && rows_ratio <= 0.985) im2col = 2;
                else if (N.wb_tap) {
                    const double k_ratio = (double)N.k_full_tap / (double)round_up(N.k_full, 64);
                    if (rows_ratio * k_ratio <= 0.975) im2col = 1;
                }
            }
            int rc = SB_OK;
            for (int attempt = 0; attempt < 4; ++attempt) {
                T.im2col = 0;
                if (im2col == 2) {
                    T.per_tap = 0; T.im2col = 2; T.taps = N.k; T.w = N.wb; T.w_ld = N.w_ld; T.k_full = N.k_full;
                } else if (im2col == 1) {
                    T.per_tap = 1; T.im2col = 1; T.taps = N.k; T.w = N.wb_tap; T.w_ld = N.w_ld_tap; T.k_full = N.k_full_tap;
                } else if (per_tap) {
                    T.per_tap = 1; T.taps = N.k; T.w = N.wb_tap; T.w_ld = N.w_ld_tap; T.k_full = N.k_full_tap;
                } else {
                    T.per_tap = 0; T.taps = N.k; T.w = N.wb; T.w_ld = N.w_ld; T.k_full = N.k_full;
                }
                rc = sb_tc_launch(T, dev, stream, last ? fused : nullptr);
                if (rc == SB_OK || N.type != SB_LAYER_CONV) break;
                if (im2col == 2) { net->no_im2col_run = true; im2col = N.wb_tap ? 1 : 0; continue; }
                if (im2col == 1) { im2col = 0; net->no_im2col = true; continue; }
                if (per_tap) break;
                per_tap = true;            // overlapping-stride tensor map refused: fall back to per-tap boxes
                net->force_per_tap = true;
            }
            if (rc != SB_OK) return rc;
        } else {
            const long long rows = m * N.t_in;
            const float* a = reinterpret_cast<const float*>(act[cur]);
            const long long a_elems = rows * N.cin / (N.type == SB_LAYER_CONV ? 1 : N.t_in);
            const int lda = (N.type == SB_LAYER_CONV) ? N.cin : N.cin;
            const bool direct = (N.type == SB_LAYER_FC);
            float* o = direct ? (last ? logits : reinterpret_cast<float*>(act[cur ^ 1])) : full;
            const int ldo = direct ? (last ? ldl : N.cout) : N.cout;
            dim3 grid((unsigned)((rows + FBM - 1) / FBM), (unsigned)((N.cout + FBN - 1) / FBN));
            gemm_f32_kernel<<<grid, 256, 0, stream>>>(a, a_elems, lda, rows, N.k_f32, N.wf, N.ldw_f32, N.biasf, N.relu, o,
                                                      ldo, N.cout);
            SB_COUNT_LAUNCH();
            SB_CHECK_LAUNCH();
            if (!direct) {
                SB_REQUIRE(!last, "network cannot end in a conv layer");
                const long long total = m * N.t_pool * N.cout;
                pool_compact_kernel<float><<<grid_for(total, dev), 256, 0, stream>>>(
                    full, N.cout, N.t_in, N.pool, N.t_pool, N.t_pool, N.cout, N.cout, m,
                    reinterpret_cast<float*>(act[cur ^ 1]));
                SB_COUNT_LAUNCH();
                SB_CHECK_LAUNCH();
            }
        }
        time_end(net, li, stream);
        cur ^= 1;
    }
    *logits_out = logits;
    *ldl_out = ldl;
    return SB_OK;
}

int check_common(sb_net* net, long long n, int precision, bool generic, void* ws, size_t ws_bytes) {
    SB_REQUIRE(net, "null network handle");
    SB_REQUIRE(n >= 0, "negative row count");
    SB_REQUIRE(precision == SB_PREC_FP32 || precision == SB_PREC_BF16, "bad precision %d", precision);
    if (n == 0) return SB_OK;
    const WsPlan plan = plan_ws(net, n, precision, generic);
    SB_REQUIRE(ws && ws_bytes >= plan.total, "workspace too small: need %zu bytes, have %zu", plan.total, ws_bytes);
    return SB_OK;
}

}  // namespace

extern "C" size_t sb_net_workspace_bytes(const sb_net* net, int64_t n_rows, int precision) {
    if (!net || n_rows <= 0) return 0;
    const size_t a = plan_ws(net, n_rows, precision, false).total;
    const size_t b = plan_ws(net, n_rows, precision, true).total;
    return a > b ? a : b;
}

static int forward_impl(sb_net* net, const RowSource& in, int64_t n, int precision, void* ws, size_t ws_bytes,
                        float* pred, int pair_mode, const float* base_pred, const int32_t* base_index, bool scoring,
                        float* abs_diff, float* diff, float* logit, float* pred_ref, float* pred_alt, void* stream) {
    const bool generic = in.codes == nullptr;
    SB_REQUIRE(!(net && net->inference_weights_stale),
               "the weights were updated by sb_net_sgd_step: export them with sb_net_get_layer_params and create an "
               "inference handle from them");
    int rc = check_common(net, n, precision, generic, ws, ws_bytes);
    if (rc) return rc;
    if (n == 0) return SB_OK;
    SB_CHECK_CUDA(cudaSetDevice(net->device));
    cudaStream_t s = (cudaStream_t)stream;
    const WsPlan plan = plan_ws(net, n, precision, generic);
    const int F = net->n_out;
    long long chunk = chunk_for(net, precision, generic);
    if (scoring && pair_mode == SB_PAIR_INTERLEAVED) {
        SB_REQUIRE(n % 2 == 0, "interleaved ref/alt rows need an even row count, got %lld", (long long)n);
        chunk &= ~1ll;
    }
    const bool ahead = conv1_ahead(net, n, precision, generic) && plan.c1 > 0;
    void* c1_buf[2] = {nullptr, nullptr};
    if (ahead) {
        if (!net->side) {
            SB_CHECK_CUDA(cudaStreamCreateWithFlags(&net->side, cudaStreamNonBlocking));
            SB_CHECK_CUDA(cudaEventCreateWithFlags(&net->ev_fork, cudaEventDisableTiming));
            for (int i = 0; i < 2; ++i) {
                SB_CHECK_CUDA(cudaEventCreateWithFlags(&net->ev_c1[i], cudaEventDisableTiming));
                SB_CHECK_CUDA(cudaEventCreateWithFlags(&net->ev_done[i], cudaEventDisableTiming));
            }
        }
        uint8_t* b0 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~(uintptr_t)1023) + 2 * plan.act +
                      plan.full + plan.logits + plan.lstm + plan.strand;
        c1_buf[0] = b0;
        c1_buf[1] = b0 + plan.c1;
        // the side stream starts after everything already queued on the caller's stream (the inputs)
        SB_CHECK_CUDA(cudaEventRecord(net->ev_fork, s));
        SB_CHECK_CUDA(cudaStreamWaitEvent(net->side, net->ev_fork, 0));
        float* lg0 = nullptr;
        int ldl0 = 0;
        const long long m0 = n < chunk ? n : chunk;
        rc = run_chunk(net, in, 0, m0, precision, reinterpret_cast<uint8_t*>(ws), plan, &lg0, &ldl0, net->side, nullptr, 1, c1_buf[0]);
        if (rc) return rc;
        SB_CHECK_CUDA(cudaEventRecord(net->ev_c1[0], net->side));
    }
    long long k_chunk = 0;
    for (long long r0 = 0; r0 < n; r0 += chunk, ++k_chunk) {
        const long long m = (n - r0) < chunk ? (n - r0) : chunk;
        float* lg = nullptr;
        int ldl = 0;
        if (ahead) {
            // first conv of chunk k+1 (side stream; its buffer was last read by chunk k-1), then layers 1.. of chunk k
            const long long r1 = r0 + chunk;
            if (r1 < n) {
                const long long m1 = (n - r1) < chunk ? (n - r1) : chunk;
                if (k_chunk >= 1) SB_CHECK_CUDA(cudaStreamWaitEvent(net->side, net->ev_done[(k_chunk - 1) & 1], 0));
                rc = run_chunk(net, in, r1, m1, precision, reinterpret_cast<uint8_t*>(ws), plan, &lg, &ldl, net->side, nullptr, 1,
                               c1_buf[(k_chunk + 1) & 1]);
                if (rc) return rc;
                SB_CHECK_CUDA(cudaEventRecord(net->ev_c1[(k_chunk + 1) & 1], net->side));
            }
            SB_CHECK_CUDA(cudaStreamWaitEvent(s, net->ev_c1[k_chunk & 1], 0));
            SbScoreEpilogue E;
            memset(&E, 0, sizeof(E));
            E.F = F;
            E.row0 = r0;
            if (!scoring) { E.mode = 0; E.pred = pred; }
            else {
                E.mode = pair_mode == SB_PAIR_INTERLEAVED ? 1 : 2;
                E.base_pred = base_pred; E.base_index = base_index;
                E.abs_diff = abs_diff; E.diff = diff; E.logit = logit;
                E.pred_ref = pair_mode == SB_PAIR_INTERLEAVED ? pred_ref : nullptr;
                E.pred_alt = pred_alt;
            }
            rc = run_chunk(net, in, r0, m, precision, reinterpret_cast<uint8_t*>(ws), plan, &lg, &ldl, s, &E, 2, c1_buf[k_chunk & 1]);
            if (rc) return rc;
            SB_CHECK_CUDA(cudaEventRecord(net->ev_done[k_chunk & 1], s));
            continue;
        }
        if (net->strand_mode) {
            // both strands through the network, combined probabilities, then the handlers' math on those
            uint8_t* sbase = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~(uintptr_t)1023) +
                             2 * plan.act + plan.full + plan.logits + plan.lstm;
            const long long cr = n < chunk ? n : chunk;
            float* p_fwd = reinterpret_cast<float*>(sbase);
            float* p_rev = p_fwd + cr * F;
            uint8_t* rc_rows = sbase + ((2 * (size_t)cr * F * 4 + 1023) / 1024 * 1024 + 1024);
            RowSource rin = {nullptr, nullptr, nullptr, nullptr, nullptr};
            if (generic) {
                rc_onehot_kernel<<<grid_for(m * net->L * 4, net->device), 256, 0, s>>>(in.onehot, r0, m, net->L,
                                                                                      reinterpret_cast<float*>(rc_rows));
                rin.onehot = reinterpret_cast<const float*>(rc_rows);
            } else {
                RcMap map;
                uint8_t inv[4];
                for (int b = 0; b < 4; ++b) inv[net->perm[b]] = (uint8_t)b;
                for (int b = 0; b < 4; ++b) map.m[b] = inv[3 - net->perm[b]];
                map.m[4] = 4;
                rc_codes_kernel<<<grid_for(m * net->L, net->device), 256, 0, s>>>(in.codes, in.src, in.sub_pos, in.sub_code,
                                                                                r0, m, net->L, map, rc_rows);
                rin.codes = rc_rows;
            }
            SB_COUNT_LAUNCH();
            SB_CHECK_LAUNCH();
            for (int pass = 0; pass < 2; ++pass) {
                const RowSource& src_rows = pass == 0 ? in : rin;
                const long long row0 = pass == 0 ? r0 : 0;
                float* dst = pass == 0 ? p_fwd : p_rev;
                if (precision == SB_PREC_BF16) {
                    SbScoreEpilogue E;
                    memset(&E, 0, sizeof(E));
                    E.F = F; E.row0 = 0; E.mode = 0; E.pred = dst;
                    rc = run_chunk(net, src_rows, row0, m, precision, reinterpret_cast<uint8_t*>(ws), plan, &lg, &ldl, s, &E);
                    if (rc) return rc;
                } else {
                    rc = run_chunk(net, src_rows, row0, m, precision, reinterpret_cast<uint8_t*>(ws), plan, &lg, &ldl, s);
                    if (rc) return rc;
                    sigmoid_kernel<<<grid_for(m * F, net->device), 256, 0, s>>>(lg, ldl, m, F, dst);
                    SB_COUNT_LAUNCH();
                    SB_CHECK_LAUNCH();
                }
            }
            float* comb = scoring ? p_fwd : pred + r0 * F;
            strand_combine_kernel<<<grid_for(m * F, net->device), 256, 0, s>>>(p_fwd, p_rev, m * F, net->strand_mode, comb);
            SB_COUNT_LAUNCH();
            SB_CHECK_LAUNCH();
            if (scoring) {
                if (pair_mode == SB_PAIR_INTERLEAVED) {
                    const long long v0 = r0 / 2, nv = m / 2;
                    score_kernel<<<grid_for(nv * F, net->device), 256, 0, s>>>(
                        comb, F, 0, SB_PAIR_INTERLEAVED, nullptr, nullptr, 0, nv, F, abs_diff ? abs_diff + v0 * F : nullptr,
                        diff ? diff + v0 * F : nullptr, logit ? logit + v0 * F : nullptr,
                        pred_ref ? pred_ref + v0 * F : nullptr, pred_alt ? pred_alt + v0 * F : nullptr);
                } else {
                    score_kernel<<<grid_for(m * F, net->device), 256, 0, s>>>(
                        comb, F, 0, SB_PAIR_BASELINE, base_pred, base_index ? base_index + r0 : nullptr, /*n_base*/ 1, m, F,
                        abs_diff ? abs_diff + r0 * F : nullptr, diff ? diff + r0 * F : nullptr,
                        logit ? logit + r0 * F : nullptr, nullptr, pred_alt ? pred_alt + r0 * F : nullptr);
                }
                SB_COUNT_LAUNCH();
                SB_CHECK_LAUNCH();
            }
            continue;
        }
        if (precision == SB_PREC_BF16) {
            // kernel (d): sigmoid + scores are the epilogue of the last layer's tcgen05 kernel
            SbScoreEpilogue E;
            memset(&E, 0, sizeof(E));
            E.F = F;
            E.row0 = r0;
            if (!scoring) { E.mode = 0; E.pred = pred; }
            else {
                E.mode = pair_mode == SB_PAIR_INTERLEAVED ? 1 : 2;
                E.base_pred = base_pred; E.base_index = base_index;
                E.abs_diff = abs_diff; E.diff = diff; E.logit = logit;
                E.pred_ref = pair_mode == SB_PAIR_INTERLEAVED ? pred_ref : nullptr;
                E.pred_alt = pred_alt;
            }
            rc = run_chunk(net, in, r0, m, precision, reinterpret_cast<uint8_t*>(ws), plan, &lg, &ldl, s, &E);
            if (rc) return rc;
            continue;
        }
        rc = run_chunk(net, in, r0, m, precision, reinterpret_cast<uint8_t*>(ws), plan, &lg, &ldl, s);
        if (rc) return rc;
        if (!scoring) {
            sigmoid_kernel<<<grid_for(m * F, net->device), 256, 0, s>>>(lg, ldl, m, F, pred + r0 * F);
        } else if (pair_mode == SB_PAIR_INTERLEAVED) {
            const long long v0 = r0 / 2, nv = m / 2;
            score_kernel<<<grid_for(nv * F, net->device), 256, 0, s>>>(
                lg, ldl, 1, SB_PAIR_INTERLEAVED, nullptr, nullptr, 0, nv, F, abs_diff ? abs_diff + v0 * F : nullptr,
                diff ? diff + v0 * F : nullptr, logit ? logit + v0 * F : nullptr, pred_ref ? pred_ref + v0 * F : nullptr,
                pred_alt ? pred_alt + v0 * F : nullptr);
        } else {
            score_kernel<<<grid_for(m * F, net->device), 256, 0, s>>>(
                lg, ldl, 1, SB_PAIR_BASELINE, base_pred, base_index ? base_index + r0 : nullptr, /*n_base*/ 1, m, F,
                abs_diff ? abs_diff + r0 * F : nullptr, diff ? diff + r0 * F : nullptr,
                logit ? logit + r0 * F : nullptr, nullptr, pred_alt ? pred_alt + r0 * F : nullptr);
        }
        SB_COUNT_LAUNCH();
        SB_CHECK_LAUNCH();
    }
    return SB_OK;
}

extern "C" int sb_net_forward_codes(sb_net* net, const uint8_t* codes, const int32_t* src, const int32_t* sub_pos,
                                    const uint8_t* sub_code, int64_t n, int precision, void* workspace,
                                    size_t workspace_bytes, float* pred, void* stream) {
    SB_REQUIRE(n == 0 || (codes && pred), "sb_net_forward_codes: null argument");
    RowSource in = {codes, src, sub_pos, sub_code, nullptr};
    return forward_impl(net, in, n, precision, workspace, workspace_bytes, pred, 0, nullptr, nullptr, false, nullptr,
                        nullptr, nullptr, nullptr, nullptr, stream);
}

extern "C" int sb_net_forward_onehot(sb_net* net, const float* x, int64_t n, int precision, void* workspace,
                                     size_t workspace_bytes, float* pred, void* stream) {
    SB_REQUIRE(n == 0 || (x && pred), "sb_net_forward_onehot: null argument");
    RowSource in = {nullptr, nullptr, nullptr, nullptr, x};
    return forward_impl(net, in, n, precision, workspace, workspace_bytes, pred, 0, nullptr, nullptr, false, nullptr,
                        nullptr, nullptr, nullptr, nullptr, stream);
}

extern "C" int sb_net_forward_score(sb_net* net, const uint8_t* codes, const int32_t* src, const int32_t* sub_pos,
                                    const uint8_t* sub_code, int64_t n, int precision, int pair_mode,
                                    const float* base_pred, const int32_t* base_index, void* workspace,
                                    size_t workspace_bytes, float* abs_diff, float* diff, float* logit,
                                    float* pred_ref, float* pred_alt, void* stream) {
    SB_REQUIRE(n == 0 || codes, "sb_net_forward_score: null argument");
    SB_REQUIRE(pair_mode == SB_PAIR_INTERLEAVED || pair_mode == SB_PAIR_BASELINE, "sb_net_forward_score: bad pair_mode");
    SB_REQUIRE(pair_mode != SB_PAIR_BASELINE || base_pred, "sb_net_forward_score: baseline mode needs base_pred");
    RowSource in = {codes, src, sub_pos, sub_code, nullptr};
    return forward_impl(net, in, n, precision, workspace, workspace_bytes, nullptr, pair_mode, base_pred, base_index,
                        true, abs_diff, diff, logit, pred_ref, pred_alt, stream);
}

extern "C" int sb_score_pairs(const float* ref, const float* alt, const int32_t* base_index, int64_t n, int64_t n_ref,
                              int F, float* abs_diff, float* diff, float* logit, int clamp_inplace, void* stream) {
    SB_REQUIRE(n >= 0 && F > 0, "sb_score_pairs: bad sizes");
    if (n == 0) return SB_OK;
    SB_REQUIRE(ref && alt, "sb_score_pairs: null argument");
    (void)clamp_inplace;
    int dev = 0;
    SB_CHECK_CUDA(cudaGetDevice(&dev));
    score_kernel<<<grid_for(n * F, dev), 256, 0, (cudaStream_t)stream>>>(alt, F, 0, SB_PAIR_BASELINE, ref, base_index,
                                                                         n_ref, n, F, abs_diff, diff, logit, nullptr,
                                                                         nullptr);
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    return SB_OK;
}

#include "sb_train.inl"
#include "sb_train_tc.inl"

// ---------------------------------------------------------------------------------------------
// fp32 <-> bf16 casts of a flat vector (data-parallel training: gradients travel as bf16, train_model.py)
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ in, long long n, __nv_bfloat16* __restrict__ out) {
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(in) + i);
        __nv_bfloat162 a = __floats2bfloat162_rn(f.x, f.y), b = __floats2bfloat162_rn(f.z, f.w);
        reinterpret_cast<uint2*>(out)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void __launch_bounds__(256) cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ in, long long n, float* __restrict__ out) {
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(in) + i);
        reinterpret_cast<float4*>(out)[i] = make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xFFFF0000u),
                                                        __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xFFFF0000u));
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = __bfloat162float(in[i]);
}
}  // namespace

extern "C" int sb_cast_f32_to_bf16(const float* in, int64_t n, void* out, void* stream) {
    SB_REQUIRE(n >= 0, "sb_cast_f32_to_bf16: bad n");
    if (n == 0) return SB_OK;
    SB_REQUIRE(in && out, "sb_cast_f32_to_bf16: null argument");
    SB_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0,
               "sb_cast_f32_to_bf16: buffers must be 16-byte (fp32) / 8-byte (bf16) aligned");
    int dev = 0;
    SB_CHECK_CUDA(cudaGetDevice(&dev));
    cast_f32_bf16_kernel<<<grid_for(n / 4 + 1, dev), 256, 0, (cudaStream_t)stream>>>(in, n, reinterpret_cast<__nv_bfloat16*>(out));
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    return SB_OK;
}

extern "C" int sb_cast_bf16_to_f32(const void* in, int64_t n, float* out, void* stream) {
    SB_REQUIRE(n >= 0, "sb_cast_bf16_to_f32: bad n");
    if (n == 0) return SB_OK;
    SB_REQUIRE(in && out, "sb_cast_bf16_to_f32: null argument");
    SB_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(in) & 7) == 0,
               "sb_cast_bf16_to_f32: buffers must be 16-byte (fp32) / 8-byte (bf16) aligned");
    int dev = 0;
    SB_CHECK_CUDA(cudaGetDevice(&dev));
    cast_bf16_f32_kernel<<<grid_for(n / 4 + 1, dev), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(in), n, out);
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    return SB_OK;
}
