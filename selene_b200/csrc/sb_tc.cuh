// selene_b200 — tcgen05 implicit-GEMM core: interface between sb_net.cu and sb_tc.cu.
#pragma once
#include "sb_common.cuh"

// One GEMM-shaped layer launch:  D[r][n] = sum_k A[r][k] * W[n][k]  (+ bias, ReLU, max-pool, store)
//
// A is the activation buffer viewed as an OVERLAPPING-stride matrix: row r starts at element
// r * a_ld and is k_full = taps * a_ld elements long, i.e. the im2col row of a valid 1-D
// convolution over a channel-last (rows = sequence x position, a_ld channels) tensor is just a
// longer read of the same memory.  FC layers are the taps = 1 case.
struct SbTcLayer {
    const void* a;       // dev bf16, rows_a x a_ld (row stride a_ld elements)
    int64_t rows_a;      // rows of the activation buffer (sequences x t_in)
    int a_ld;            // elements per row (multiple of 8)
    int k_full;          // logical K (<= taps * a_ld); TMA zero-fills beyond it
    const void* w;       // dev bf16, n_rows_w x w_ld, K-major
    int n_rows_w;        // logical output channels (cout)
    int w_ld;            // elements per weight row (multiple of 8, >= k_full)
    int per_tap;         // 0: overlapping-stride A view (K index = tap*a_ld + c, k_full = taps*a_ld)
                         // 1: one 2-D box per (tap, 64-channel chunk); the weight K axis is then
                         //    taps x (chunks_per_tap*64) with zero padding, k_full = that product
    int taps;            // only used when per_tap
    int im2col;          // 2 (without per_tap): the same map over the RUN view — a window start's channel axis is its whole
                         // taps x C run, k-blocks are 64 elements of it (no per-tap channel padding).
                         // 1 with per_tap: A is fetched through an IM2COL tensor map (channels, position, sequence) whose
                         // pixel bounding box holds only the pool * t_pool valid window starts of every sequence, so an
                         // M tile is 128 VALID rows (it walks across sequence boundaries) and no MMA row is wasted on
                         // windows that straddle two sequences
    const float* bias;   // dev fp32, padded to n_tiles_n * block_n entries
    int block_n;         // UMMA N (multiple of 16, <= 256)
    int n_tiles_n;
    int t_in;            // rows per sequence (1 for FC)
    int t_conv;          // valid conv outputs per sequence: t_in - taps + 1
    int pool;            // 1, 2 or 4 (requires t_in % pool == 0)
    int relu;
    int out_f32;         // 0: bf16 output, 1: fp32 output, 2: fused sigmoid + score epilogue (`score`),
                         // 3: split-K partial sums ADDED to a zeroed fp32 output (no bias / ReLU / pool: the caller finishes)
    int k_splits;        // out_f32 == 3: the K axis is cut into this many ranges, one tile per (m, n, range), so that a
                         // GEMM with few output tiles and a long K (FC over a small batch) still fills every SM
    void* out;           // dev, (sequences * t_pool) x out_ld
    int out_ld;          // bf16: multiple of 8; fp32: multiple of 4
    int out_pitch;       // rows per sequence in `out` (>= t_pool; the consumer's padded t_in); 0 = t_pool
};

// Fused epilogue of the LAST layer (out_f32 == 2): sigmoid of the accumulator rows, then either plain
// predictions or ref/alt scores written straight from registers — logits and predictions never visit HBM
// unless asked for.  Rows of the launch are rows [row0, row0 + rows_a) of the whole job.
struct SbScoreEpilogue {
    int mode;                 // 0: pred[row] = sigmoid; 1: interleaved (ref, alt) row pairs; 2: every row is an
                              //    alt scored against base_pred[base_index[row]]
    int F;                    // valid output columns (row stride of every output array)
    long long row0;           // first global row of this launch
    const float* base_pred;   // mode 2: (n_base, F)
    const int32_t* base_index;  // mode 2: per global row, or NULL (row 0)
    float* pred;              // mode 0: (rows, F)
    float* abs_diff;          // modes 1/2, may be NULL: (pairs or rows, F)
    float* diff;
    float* logit;
    float* pred_ref;          // mode 1 only
    float* pred_alt;
};

int sb_tc_launch(const SbTcLayer& L, int device, cudaStream_t stream, const SbScoreEpilogue* score = nullptr);

// wgrad: G (N x ldg fp32) += dY^T . Xview (see sb_tc.cu); G must be zeroed by the caller.
int sb_tc_wgrad_launch(const void* dy, int dy_ld, const void* x, int x_ld, long long x_rows_alloc, long long rows,
                       int M, int N, float* g, int ldg, int device, cudaStream_t stream, int overwrite = 0);
bool sb_tc_wgrad_overwrites(long long rows, int M, int N, int ldg, int device);
