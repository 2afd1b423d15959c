// selene_b200 — training step with the GEMMs on tcgen05 (bf16 operands, fp32 accumulation, fp32 master
// weights and gradients).  Included by sb_net.cu after sb_train.inl.
//
// Same mathematics and layouts as sb_train.inl; per layer (index >= 1):
//   forward  tc_layer_kernel (pool unfused so the pre-pool output is kept for the backward pass)
//   dgrad    tc_layer_kernel over the shifted overlapping view of dY with the taps-flipped W^T
//   wgrad    tc_wgrad_kernel (MN-major operands, split-K, red.add into the fp32 gradient)
// The first conv (4 input channels) stays on CUDA cores: its forward is a 32-deep dot product per output
// and its weight gradient a 32 x cout table accumulated in registers.

namespace {

typedef __nv_bfloat16 bf16;

__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float* __restrict__ in, long long rows, int cols,
                                                          int ld_in, bf16* __restrict__ out, int ld_out) {
    const long long total = rows * ld_out;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int c = (int)(i % ld_out);
        const long long r = i / ld_out;
        out[i] = __float2bfloat16_rn(c < cols ? in[r * ld_in + c] : 0.f);
    }
}

__global__ void __launch_bounds__(256) bf16_to_f32_kernel(const bf16* __restrict__ in, long long rows, int cols,
                                                          int ld_in, float* __restrict__ out, int ld_out) {
    const long long total = rows * cols;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int c = (int)(i % cols);
        const long long r = i / cols;
        out[r * ld_out + c] = __bfloat162float(in[r * ld_in + c]);
    }
}

// forward B operand from the fp32 master: wb[o][map(kk)] = W[kk][o], map(kk) = (kk / ch) * ch_ld + kk % ch.
// 32 x 32 tiles through shared memory so that both the fp32 reads (o contiguous) and the bf16 writes
// (kk contiguous) are coalesced.
__global__ void __launch_bounds__(256) pack_fwd_kernel(const float* __restrict__ wf, int ldw, long long K, int ch, int ch_ld,
                                                       int cout, bf16* __restrict__ wb, int w_ld) {
    __shared__ float tile[32][33];
    const long long k0 = (long long)blockIdx.x * 32;
    const int o0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const long long kk = k0 + r;
        const int o = o0 + tx;
        tile[r][tx] = (kk < K && o < cout) ? wf[kk * ldw + o] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int o = o0 + r;
        const long long kk = k0 + tx;
        if (o < cout && kk < K) {
            const long long col = (kk / ch) * ch_ld + kk % ch;
            wb[(long long)o * w_ld + col] = __float2bfloat16_rn(tile[tx][r]);
        }
    }
}

// SGD update (torch.optim.SGD with momentum and weight decay, as sgd_kernel) of one layer's fp32 master weights fused
// with both repackings above: one pass over w / momentum / gradient instead of three passes over w.
template <typename GT>
__global__ void __launch_bounds__(256, 8) sgd_pack_kernel(float* __restrict__ wf, float* __restrict__ v, const GT* __restrict__ g,
                                                          int ldw, int K, int cout, float lr, float mom, float wd, int first,
                                                          int ch, int ch_ld, bf16* __restrict__ wb, int w_ld, int taps, int C,
                                                          int cout_ld, bf16* __restrict__ wt, int wt_ld, int o_blocks) {
    __shared__ float tile[32][33];
    // o_blocks > 0: 1-D grid with the channel blocks fastest, so the CTAs in flight sweep whole rows of w / momentum /
    // gradient (contiguous megabytes) instead of a 128-byte column strip with a row-pitch stride.  All index arithmetic
    // is 32-bit and the tap / channel split of a row is skipped for FC layers (ncu, r02k: the 64-bit divisions of the
    // first version kept the issue slots 64 % busy on a kernel that should only wait for HBM).
    const int k0 = (o_blocks > 0 ? (int)(blockIdx.x / o_blocks) : (int)blockIdx.x) * 32;
    const int o0 = (o_blocks > 0 ? (int)(blockIdx.x % o_blocks) : (int)blockIdx.y) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const int o = o0 + tx;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int kk = k0 + r;
        float wn = 0.f;
        if (kk < K && o < ldw) {
            const long long i = (long long)kk * ldw + o;
            const float w0 = wf[i];
            const float gi = grad_at(g, i) + wd * w0;
            const float vi = first ? gi : mom * v[i] + gi;
            v[i] = vi;
            wn = w0 - lr * vi;
            wf[i] = wn;
            if (o < cout) {   // dgrad operand, same orientation: wt[c][tap' * cout_ld + o], taps flipped
                int c = kk, tp = 0;
                if (taps > 1) {
                    const int t = (int)((unsigned)kk / (unsigned)C);
                    c = kk - t * C;
                    tp = taps - 1 - t;
                }
                wt[(long long)c * wt_ld + tp * cout_ld + o] = __float2bfloat16_rn(wn);
            }
        }
        tile[r][tx] = wn;
    }
    __syncthreads();
    const int kk = k0 + tx;
    int col = kk;                                              // forward operand, transposed: wb[o][col(kk)]
    if (ch != ch_ld) {
        const int q = (int)((unsigned)kk / (unsigned)ch);
        col = q * ch_ld + (kk - q * ch);
    }
    if (kk < K) {
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const int oo = o0 + r;
            if (oo < cout) wb[(long long)oo * w_ld + col] = __float2bfloat16_rn(tile[tx][r]);
        }
    }
}

// dgrad B operand: wt[c][tap'*cout_ld + o] = W[(taps-1-tap')*C + c][o]   (taps = 1: plain W^T rows)
__global__ void __launch_bounds__(256) pack_dgrad_kernel(const float* __restrict__ wf, int ldw, int taps, int C, int cout,
                                                         int cout_ld, bf16* __restrict__ wt, int wt_ld) {
    const long long total = (long long)taps * C * cout;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int o = (int)(i % cout);
        const long long r = i / cout;
        const int c = (int)(r % C);
        const int tp = (int)(r / C);
        wt[(long long)c * wt_ld + (long long)tp * cout_ld + o] =
            __float2bfloat16_rn(wf[((long long)(taps - 1 - tp) * C + c) * ldw + o]);
    }
}

// 8 bf16 channels <-> 8 floats (16-byte vectors: every leading dimension here is a multiple of 8)
__device__ __forceinline__ void bf16x8_to_f32(const uint4 v, float (&f)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        f[2 * k] = __uint_as_float(w[k] << 16);
        f[2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u);
    }
}
__device__ __forceinline__ uint4 f32_to_bf16x8(const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
        w[k] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// pooled[s][tp][c] = max_q full[s][tp*pool+q][c]  (bf16 in / out), optional dropout (mask fp32, per element).
// A thread owns 8 channels of one pooled position: `pool` 16-byte loads, one 16-byte store.
__global__ void __launch_bounds__(256) pool_fwd_bf16_kernel(const bf16* __restrict__ full, int t_full, int pool, int t_pool,
                                                            int ld, long long n_seq, bf16* __restrict__ out,
                                                            const float* __restrict__ mask_in, float* __restrict__ mask_out,
                                                            float p, unsigned long long seed) {
    const int ld8 = ld >> 3;
    const long long total = n_seq * t_pool * ld8;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
    for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += stride) {
        const int c8 = (int)(u % ld8);
        const long long rt = u / ld8;
        const int tp = (int)(rt % t_pool);
        const long long s = rt / t_pool;
        float m[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
        const uint4* src = reinterpret_cast<const uint4*>(full + (s * t_full + (long long)tp * pool) * ld) + c8;
        for (int q = 0; q < pool; ++q) {
            float f[8];
            bf16x8_to_f32(__ldg(src + (long long)q * ld8), f);
#pragma unroll
            for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], f[k]);
        }
        const long long i0 = rt * ld + 8 * c8;       // element index of channel 8*c8 in the pooled tensor
        if (mask_in) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(mask_in + i0)), b = __ldg(reinterpret_cast<const float4*>(mask_in + i0) + 1);
            m[0] *= a.x; m[1] *= a.y; m[2] *= a.z; m[3] *= a.w; m[4] *= b.x; m[5] *= b.y; m[6] *= b.z; m[7] *= b.w;
        } else if (p > 0.f) {
            float kk[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                kk[k] = hash_uniform(seed, (unsigned long long)(i0 + k)) >= p ? scale : 0.f;
                m[k] *= kk[k];
            }
            reinterpret_cast<float4*>(mask_out + i0)[0] = make_float4(kk[0], kk[1], kk[2], kk[3]);
            reinterpret_cast<float4*>(mask_out + i0)[1] = make_float4(kk[4], kk[5], kk[6], kk[7]);
        }
        reinterpret_cast<uint4*>(out + i0)[0] = f32_to_bf16x8(m);
    }
}

// dfull[s][t][c] (t < t_in rows per sequence) from gout[s][t/pool][c]: first maximum of the window and > 0, times the
// dropout mask; rows t >= t_pool*pool are zero.  A thread owns 8 channels of one pooling window (or of one tail row).
__global__ void __launch_bounds__(256) pool_relu_bwd_bf16_kernel(const bf16* __restrict__ full, int t_full, const bf16* __restrict__ gout,
                                                                 const float* __restrict__ mask, int t_in, int pool, int t_pool,
                                                                 int ld, long long n_seq, bf16* __restrict__ dfull) {
    const int ld8 = ld >> 3;
    const int tail = t_in - t_pool * pool;                  // zero rows at the end of every sequence
    const int units_seq = t_pool + tail;                    // windows + tail rows
    const long long total = n_seq * units_seq * ld8;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += stride) {
        const int c8 = (int)(u % ld8);
        const long long ru = u / ld8;
        const int w = (int)(ru % units_seq);
        const long long s = ru / units_seq;
        if (w >= t_pool) {                                  // tail row
            const long long t = (long long)t_pool * pool + (w - t_pool);
            reinterpret_cast<uint4*>(dfull + (s * t_in + t) * ld)[c8] = make_uint4(0u, 0u, 0u, 0u);
            continue;
        }
        const uint4* src = reinterpret_cast<const uint4*>(full + (s * t_full + (long long)w * pool) * ld) + c8;
        float best[8];
        int arg[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { best[k] = -INFINITY; arg[k] = 0; }
        for (int q = 0; q < pool; ++q) {
            float f[8];
            bf16x8_to_f32(__ldg(src + (long long)q * ld8), f);
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (f[k] > best[k]) { best[k] = f[k]; arg[k] = q; }      // strict: the FIRST maximum keeps the gradient
        }
        const long long oi = (s * t_pool + w) * ld + 8 * c8;
        float g[8];
        bf16x8_to_f32(__ldg(reinterpret_cast<const uint4*>(gout + oi)), g);
        if (mask) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(mask + oi)), b = __ldg(reinterpret_cast<const float4*>(mask + oi) + 1);
            g[0] *= a.x; g[1] *= a.y; g[2] *= a.z; g[3] *= a.w; g[4] *= b.x; g[5] *= b.y; g[6] *= b.z; g[7] *= b.w;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (!(best[k] > 0.f)) g[k] = 0.f;                // ReLU
        uint4* dst = reinterpret_cast<uint4*>(dfull + (s * t_in + (long long)w * pool) * ld) + c8;
        for (int q = 0; q < pool; ++q) {
            float o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = arg[k] == q ? g[k] : 0.f;
            dst[(long long)q * ld8] = f32_to_bf16x8(o);
        }
    }
}

// FC: dy = gout * mask * (y > 0)  in place (bf16)
__global__ void __launch_bounds__(256) fc_bwd_bf16_kernel(bf16* __restrict__ g, const bf16* __restrict__ y, const float* __restrict__ mask,
                                                          int relu, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float v = __bfloat162float(g[i]);
        if (mask) v *= mask[i];
        if (relu && !(__bfloat162float(y[i]) > 0.f)) v = 0.f;
        g[i] = __float2bfloat16_rn(v);
    }
}

// y[r][c] = act(acc[r][c] + bias[c]) as bf16 (c < ld; columns >= cout are zero): finishes a split-K FC forward
__global__ void __launch_bounds__(256) bias_act_bf16_kernel(const float* __restrict__ acc, const float* __restrict__ bias,
                                                            long long rows, int cout, int ld, int relu, bf16* __restrict__ y) {
    const long long total = rows * ld;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % ld);
        float v = 0.f;
        if (c < cout) {
            v = acc[i] + bias[c];
            if (relu) v = fmaxf(v, 0.f);
        }
        y[i] = __float2bfloat16(v);
    }
}

__global__ void __launch_bounds__(256) dropout_bf16_kernel(bf16* __restrict__ y, long long n, const float* __restrict__ mask_in,
                                                           float* __restrict__ mask_out, float p, unsigned long long seed) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float scale = 1.f / (1.f - p);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float m;
        if (mask_in) m = mask_in[i];
        else { m = hash_uniform(seed, (unsigned long long)i) >= p ? scale : 0.f; mask_out[i] = m; }
        y[i] = __float2bfloat16_rn(__bfloat162float(y[i]) * m);
    }
}

// BCE on fp32 logits -> loss (fp32 atomic) and dz (bf16, padded columns zero)
__global__ void __launch_bounds__(256) bce_bf16_kernel(const float* __restrict__ z, int ldz, const float* __restrict__ y,
                                                       long long n, int F, bf16* __restrict__ dz, int lddz,
                                                       float* __restrict__ loss) {
    const long long total = n * lddz;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float inv = 1.f / (float)(n * F);
    float local = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / lddz;
        const int f = (int)(i - r * lddz);
        float g = 0.f;
        if (f < F) {
            const float p = sigmoid_f32(z[r * ldz + f]);
            const float t = y[r * F + f];
            const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(logf(1.f - p), -100.f);
            local += -(t * lp + (1.f - t) * lq);
            g = (p - t) * inv;
        }
        dz[i] = __float2bfloat16_rn(g);
    }
    local *= inv;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(loss, local);
}

// db[o] += sum_r dy[r][o] (bias gradient).  A CTA owns a strip of 64 channels (one 128-byte line per row) and a slab of
// rows: 8 lanes x 16-byte loads cover the strip, the other 32 threads-per-strip interleave the rows, 8 loads are in
// flight per thread; one atomic per (CTA, channel) at the end, so a channel sees as many adds as there are slabs.
constexpr int kColsumStrip = 64;
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const bf16* __restrict__ dy, long long rows, int ld, int n_cols,
                                                          long long rows_per_slab, float* __restrict__ db) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cq = threadIdx.x & 7;                                // 8 channels = one uint4
    const int rl = threadIdx.x >> 3;                               // 0..31: row lane
    const int c0 = blockIdx.x * kColsumStrip + cq * 8;
    const long long r0 = (long long)blockIdx.y * rows_per_slab;
    const long long r1 = r0 + rows_per_slab < rows ? r0 + rows_per_slab : rows;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    if (c0 < ld) {                                                 // ld % 8 == 0: a uint4 is inside the row or not at all
        const uint4* src = reinterpret_cast<const uint4*>(dy + c0);
        const long long ld8 = ld / 8;
        long long r = r0 + rl;
        for (; r + 7 * 32 < r1; r += 8 * 32) {
            uint4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(src + (r + 32 * u) * ld8);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                float f[8];
                bf16x8_to_f32(v[u], f);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] += f[k];
            }
        }
        for (; r < r1; r += 32) {
            float f[8];
            bf16x8_to_f32(__ldg(src + r * ld8), f);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += f[k];
        }
    }
    // the 4 row lanes of a warp, then the 8 warps
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 8);
        acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 16);
    }
    __shared__ float red[8][kColsumStrip];
    if (lane < 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) red[warp][lane * 8 + k] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x < kColsumStrip) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        const int o = blockIdx.x * kColsumStrip + threadIdx.x;
        if (o < n_cols && t != 0.f) atomicAdd(db + o, t);
    }
}

static void launch_colsum_bf16(const bf16* dy, long long rows, int ld, int n_cols, float* db, int dev, cudaStream_t s) {
    const int strips = (n_cols + kColsumStrip - 1) / kColsumStrip;
    long long slabs = (2 * sb_sm_count(dev) + strips - 1) / strips;
    const long long max_slabs = (rows + 255) / 256;                 // at least 8 rows per row lane
    if (slabs > max_slabs) slabs = max_slabs;
    if (slabs < 1) slabs = 1;
    const long long rows_per_slab = (rows + slabs - 1) / slabs;
    colsum_bf16_kernel<<<dim3((unsigned)strips, (unsigned)((rows + rows_per_slab - 1) / rows_per_slab)), 256, 0, s>>>(
        dy, rows, ld, n_cols, rows_per_slab, db);
}

// First-conv weight gradient: dW[(j*4 + c)][o] = sum_r x[r + j][c] * dY[r][o] with the 4*taps (<= 32)
// accumulators of output channel o in registers; each CTA covers a row range and adds its partial table.
constexpr int kW1MaxK = 32;
__global__ void __launch_bounds__(128) conv1_wgrad_kernel(const float* __restrict__ x, long long x_rows, const float* __restrict__ dy,
                                                          int ld_dy, long long rows, int taps, int cout, int rows_per_block,
                                                          float* __restrict__ gw, int ldw, float* __restrict__ gb) {
    __shared__ float xs[64 + 8][4];
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const int o = blockIdx.y * 128 + threadIdx.x;
    float acc[kW1MaxK];
    float gsum = 0.f;   // bias gradient of channel o over this CTA's rows
#pragma unroll
    for (int k = 0; k < kW1MaxK; ++k) acc[k] = 0.f;
    for (long long rb = r0; rb < r0 + rows_per_block && rb < rows; rb += 64) {
        __syncthreads();
        for (int i = threadIdx.x; i < (64 + 8) * 4; i += 128) {
            const long long rr = rb + i / 4;
            xs[i / 4][i % 4] = rr < x_rows ? x[rr * 4 + (i % 4)] : 0.f;
        }
        __syncthreads();
        if (o < cout) {
            for (int q = 0; q < 64; ++q) {
                const long long r = rb + q;
                if (r >= rows || r >= r0 + rows_per_block) break;
                const float g = dy[r * ld_dy + o];
                gsum += g;
                if (g != 0.f) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (j < taps) acc[j * 4 + c] = fmaf(xs[q + j][c], g, acc[j * 4 + c]);
                }
            }
        }
    }
    if (o < cout) {
        for (int k = 0; k < taps * 4; ++k)
            if (acc[k] != 0.f) atomicAdd(gw + (long long)k * ldw + o, acc[k]);
        if (gsum != 0.f) atomicAdd(gb + o, gsum);
    }
}

// first conv on tcgen05: the (n, L, 4) fp32 input as bf16 rows of 8 channels (4 real + 4 zero), `pad` zero rows appended
__global__ void __launch_bounds__(256) x_to_bf16x8_kernel(const float4* __restrict__ x, long long rows, long long rows_alloc,
                                                          uint4* __restrict__ out) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows_alloc; r += (long long)gridDim.x * blockDim.x) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (r < rows) {
            const float4 f = __ldg(x + r);
            __nv_bfloat162 a = __floats2bfloat162_rn(f.x, f.y), b = __floats2bfloat162_rn(f.z, f.w);
            v.x = *reinterpret_cast<uint32_t*>(&a);
            v.y = *reinterpret_cast<uint32_t*>(&b);
        }
        out[r] = v;
    }
}

// first conv B operand + padded bias from the fp32 master: w0[o][tap*8 + c] = W[tap*4 + c][o] (c < 4, else 0)
__global__ void __launch_bounds__(256) pack_w0_kernel(const float* __restrict__ wf, int ldw, int taps, int cout, int rows_w,
                                                      int w_ld, bf16* __restrict__ w0, const float* __restrict__ bias,
                                                      float* __restrict__ bias_pad, int n_bias) {
    const int total = rows_w * w_ld;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int o = i / w_ld, k = i - o * w_ld;
        const int tap = k >> 3, c = k & 7;
        float v = 0.f;
        if (o < cout && tap < taps && c < 4) v = wf[(size_t)(tap * 4 + c) * ldw + o];
        w0[i] = __float2bfloat16_rn(v);
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_bias; i += gridDim.x * blockDim.x)
        bias_pad[i] = i < cout ? bias[i] : 0.f;
}

// gw[(tap*4 + c)][o] += g8[(tap*8 + c)][o]  (c < 4): the 8-channel rows of the tcgen05 weight gradient back to the 4-channel layout
__global__ void __launch_bounds__(256) w0_grad_unpad_kernel(const float* __restrict__ g8, int ld8, int taps, int cout,
                                                            float* __restrict__ gw, int ldw) {
    const int total = taps * 4 * cout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int kk = i / cout, o = i - kk * cout;
        gw[(size_t)kk * ldw + o] += g8[(size_t)((kk >> 2) * 8 + (kk & 3)) * ld8 + o];
    }
}

struct TcTrainLayer {
    bf16* wb;      // forward B operand (cout x w_ld)
    bf16* wt;      // dgrad B operand (cin x wt_ld)
    int wt_ld;
    int cin_ld, cout_ld;
    int dg_block_n, dg_n_tiles_n;   // tiling of the dgrad output columns (= layer inputs)
};

}  // namespace

struct sb_train_tc_state {
    std::vector<TcTrainLayer> L;
    bool packed_fresh;   // the bf16 operands match the fp32 master weights (sb_net_sgd_step repacks while it updates)
    // first conv on tcgen05 (SB_TRAIN_L0_TC=0 keeps it on CUDA cores): bf16 weights (rows x taps*8), padded bias, tiling
    bool l0_tc;
    bf16* w0b;
    float* bias0;
    int w0_ld, w0_block_n, w0_n_tiles_n;
};

namespace {

bool tc_trainable(const sb_net* net, const char** why) {
    if (net->has_lstm) { *why = "BiLSTM layers"; return false; }
    if (net->layers[0].k > 8) { *why = "first conv longer than 8 taps"; return false; }
    for (size_t li = 1; li < net->layers.size(); ++li) {
        const NetLayer& N = net->layers[li];
        const int ch = N.type == SB_LAYER_CONV ? N.cin : (N.fc_prev_c > 0 ? N.fc_prev_c : N.cin);
        if (N.type == SB_LAYER_CONV && ch % 8 != 0) { *why = "conv input channels not a multiple of 8"; return false; }
        if (N.type == SB_LAYER_FC && N.fc_prev_c > 0 && ch % 8 != 0) { *why = "flattened channels not a multiple of 8"; return false; }
    }
    return true;
}

struct TcPlan {
    std::vector<size_t> full, pooled, mask;   // per layer byte offsets (layer 0: fp32 full / fp32 pooled / mask, + bf16 pooled)
    size_t x1b, g_a, g_b, g32, logits, fc32, x8, g8, total;
};

TcPlan plan_train_tc(const sb_net* net, long long n) {
    TcPlan P;
    size_t off = 0, max_g = 0;
    for (size_t li = 0; li < net->layers.size(); ++li) {
        const NetLayer& N = net->layers[li];
        const bool conv = N.type == SB_LAYER_CONV;
        const int ld = li == 0 ? N.cout : round_up(N.cout, 8);
        const size_t es = li == 0 ? 4 : 2;
        const size_t full_b = (size_t)n * (conv ? (li == 0 ? N.t_in : N.t_conv) : 1) * ld * es;
        const size_t pooled_e = (size_t)n * N.t_pool * ld;
        P.full.push_back(off); off += align256(full_b);
        P.pooled.push_back(off); off += align256(conv ? pooled_e * es : 0);
        P.mask.push_back(off); off += align256((conv ? pooled_e : (size_t)n * ld) * 4);
        // gradient buffers: dfull with (taps-1) leading rows over t_in rows, or the layer input gradient
        const size_t dfull_b = ((size_t)n * N.t_in + N.k) * round_up(N.cout, 8) * 2;
        const size_t dx_b = (size_t)n * N.t_in * round_up(N.cin, 8) * 2;
        max_g = std::max(max_g, std::max(dfull_b, dx_b));
    }
    P.x1b = off; off += align256((size_t)n * net->layers[0].t_pool * round_up(net->layers[0].cout, 8) * 2);
    P.g_a = off; off += align256(max_g + 65536);
    P.g_b = off; off += align256(max_g + 65536);
    // fp32 scratch for the first conv's backward: dfull over all input rows with (taps-1) leading rows
    P.g32 = off; off += align256(((size_t)n * net->layers[0].t_in + 8) * net->layers[0].cout * 4 * 2);
    P.logits = off; off += align256((size_t)n * round_up(net->n_out, 4) * 4);
    // fp32 accumulator of split-K FC forwards (small batches: see the forward pass)
    size_t fc32 = 0;
    for (size_t li = 1; li < net->layers.size(); ++li)
        if (net->layers[li].type == SB_LAYER_FC) fc32 = std::max(fc32, (size_t)n * round_up(net->layers[li].cout, 8) * 4);
    P.fc32 = off; off += align256(fc32);
    // first conv on tcgen05: input as 8-channel bf16 rows (+ taps zero rows), fp32 weight gradient over 8-channel rows
    P.x8 = off; off += align256(((size_t)n * net->layers[0].t_in + net->layers[0].k + 8) * 16);
    P.g8 = off; off += align256((size_t)net->layers[0].k * 8 * net->layers[0].ldw_f32 * 4);
    P.total = off + 1024;
    return P;
}

int launch_tc_plain(const void* a, long long rows_a, int a_ld, int k_full, const void* w, int n_rows_w, int w_ld,
                    const float* bias, int block_n, int n_tiles_n, int t_in, int t_conv, int relu, int out_f32, void* out,
                    int out_ld, int out_pitch, int dev, cudaStream_t s, int k_splits = 1) {
    SbTcLayer T;
    memset(&T, 0, sizeof(T));
    T.k_splits = k_splits;
    T.a = a; T.rows_a = rows_a; T.a_ld = a_ld; T.k_full = k_full;
    T.w = w; T.n_rows_w = n_rows_w; T.w_ld = w_ld;
    T.bias = bias; T.block_n = block_n; T.n_tiles_n = n_tiles_n;
    T.t_in = t_in; T.t_conv = t_conv; T.pool = 1; T.relu = relu; T.out_f32 = out_f32; T.out = out; T.out_ld = out_ld;
    T.out_pitch = out_pitch;
    return sb_tc_launch(T, dev, s);
}

}  // namespace

extern "C" int sb_net_train_init_tc(sb_net* net) {
    int rc = sb_net_train_init(net);
    if (rc) return rc;
    if (net->train_tc) return SB_OK;
    const char* why = "";
    SB_REQUIRE(tc_trainable(net, &why), "sb_net_train_init_tc: unsupported on the tcgen05 training path (%s); use fp32", why);
    sb_train_tc_state* S = new sb_train_tc_state();
    S->packed_fresh = false;
    {
        const NetLayer& N0 = net->layers[0];
        const char* e = getenv("SB_TRAIN_L0_TC");
        S->l0_tc = !(e && e[0] == '0');
        S->w0_ld = N0.k * 8;
        S->w0_n_tiles_n = (N0.cout + 255) / 256;
        S->w0_block_n = round_up((N0.cout + S->w0_n_tiles_n - 1) / S->w0_n_tiles_n, 16);
        const size_t rows_w = (size_t)S->w0_n_tiles_n * S->w0_block_n;
        S->w0b = nullptr;
        S->bias0 = nullptr;
        SB_CHECK_CUDA(cudaMalloc(&S->w0b, rows_w * S->w0_ld * 2));
        SB_CHECK_CUDA(cudaMalloc(&S->bias0, rows_w * 4));
    }
    S->L.resize(net->layers.size());
    size_t max_cols = 256;
    for (size_t li = 1; li < net->layers.size(); ++li) {
        const NetLayer& N = net->layers[li];
        TcTrainLayer& T = S->L[li];
        T.cout_ld = round_up(N.cout, 8);
        const bool conv = N.type == SB_LAYER_CONV;
        const long long cin_rows = conv ? N.cin : N.cin;          // rows of the dgrad B operand = layer inputs
        T.wt_ld = round_up((conv ? N.k : 1) * T.cout_ld, 8);
        T.wb = N.wb;                                              // repacked in place every step
        SB_CHECK_CUDA(cudaMalloc(&T.wt, (size_t)cin_rows * T.wt_ld * 2));
        SB_CHECK_CUDA(cudaMemset(T.wt, 0, (size_t)cin_rows * T.wt_ld * 2));
        const int nt = (int)((cin_rows + 255) / 256);
        T.dg_n_tiles_n = nt;
        T.dg_block_n = round_up((int)((cin_rows + nt - 1) / nt), 16);
        max_cols = std::max(max_cols, (size_t)nt * T.dg_block_n);
    }
    SB_CHECK_CUDA(cudaMalloc(&net->zero_bias_tc, (max_cols + 256) * 4));
    SB_CHECK_CUDA(cudaMemset(net->zero_bias_tc, 0, (max_cols + 256) * 4));
    net->train_tc = S;
    return SB_OK;
}

extern "C" size_t sb_net_train_workspace_bytes_tc(const sb_net* net, int64_t batch) {
    if (!net || batch <= 0) return 0;
    return plan_train_tc(net, batch).total;
}

extern "C" int sb_net_forward_backward_tc(sb_net* net, const float* x, const float* targets, int64_t n,
                                          const float* const* dropout_masks, unsigned long long dropout_seed,
                                          void* workspace, size_t workspace_bytes, float* grads, float* loss, void* stream) {
    SB_REQUIRE(net && net->train && net->train_tc, "sb_net_forward_backward_tc: call sb_net_train_init_tc first");
    SB_REQUIRE(x && targets && grads && loss && n > 0, "sb_net_forward_backward_tc: bad arguments");
    const TcPlan P = plan_train_tc(net, n);
    SB_REQUIRE(workspace && workspace_bytes >= P.total, "sb_net_forward_backward_tc: workspace too small: need %zu, have %zu",
               P.total, workspace_bytes);
    SB_CHECK_CUDA(cudaSetDevice(net->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int dev = net->device;
    sb_train_state* T = net->train;
    sb_train_tc_state* S = net->train_tc;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    const size_t L = net->layers.size();
    const int F = net->n_out;
    const int ldl = round_up(F, 4);
    float* logits = reinterpret_cast<float*>(base + P.logits);
    auto mask_of = [&](size_t li) { return reinterpret_cast<float*>(base + P.mask[li]); };
    auto ext_mask = [&](size_t li) -> const float* { return dropout_masks ? dropout_masks[li] : nullptr; };
    auto seed_of = [&](size_t li) { return dropout_seed * 1315423911ull + li; };

    // ---- repack the bf16 operands from the fp32 master weights unless the last SGD step already did
    for (size_t li = 1; li < L && !S->packed_fresh; ++li) {
        const NetLayer& N = net->layers[li];
        const bool conv = N.type == SB_LAYER_CONV;
        const int ch = conv ? N.cin : (N.fc_prev_c > 0 ? N.fc_prev_c : N.cin);
        const int ch_ld = round_up(ch, 8);
        pack_fwd_kernel<<<dim3((unsigned)((N.k_f32 + 31) / 32), (unsigned)((N.cout + 31) / 32)), 256, 0, s>>>(
            N.wf, N.ldw_f32, N.k_f32, ch, ch_ld, N.cout, N.wb, N.w_ld);
        SB_COUNT_LAUNCH();
        pack_dgrad_kernel<<<grid_for((long long)N.k_f32 * N.cout, dev), 256, 0, s>>>(
            N.wf, N.ldw_f32, conv ? N.k : 1, conv ? N.cin : N.cin, N.cout, S->L[li].cout_ld, S->L[li].wt, S->L[li].wt_ld);
        SB_COUNT_LAUNCH();
        SB_CHECK_LAUNCH();
    }

    // ------------------------------------------------------------------ forward
    // layer 0: tcgen05 over the input as 8-channel bf16 rows (one 64-wide k-block for 8 taps), or CUDA cores (fp32)
    const NetLayer& N0 = net->layers[0];
    float* full0 = reinterpret_cast<float*>(base + P.full[0]);
    float* pooled0 = reinterpret_cast<float*>(base + P.pooled[0]);
    bf16* x1b = reinterpret_cast<bf16*>(base + P.x1b);
    const int ld0b = round_up(N0.cout, 8);
    bf16* x8 = reinterpret_cast<bf16*>(base + P.x8);
    bf16* full0b = reinterpret_cast<bf16*>(base + P.full[0]);
    if (S->l0_tc) {
        const long long rows = (long long)n * N0.t_in;
        const long long rows_alloc = rows + N0.k;
        x_to_bf16x8_kernel<<<grid_for(rows_alloc, dev), 256, 0, s>>>(reinterpret_cast<const float4*>(x), rows, rows_alloc,
                                                                     reinterpret_cast<uint4*>(x8));
        SB_COUNT_LAUNCH();
        const int rows_w = S->w0_n_tiles_n * S->w0_block_n;
        pack_w0_kernel<<<grid_for((long long)rows_w * S->w0_ld, dev), 256, 0, s>>>(N0.wf, N0.ldw_f32, N0.k, N0.cout, rows_w,
                                                                                  S->w0_ld, S->w0b, N0.biasf, S->bias0, rows_w);
        SB_COUNT_LAUNCH();
        SB_CHECK_LAUNCH();
        int rc = launch_tc_plain(x8, rows, 8, N0.k * 8, S->w0b, N0.cout, S->w0_ld, S->bias0, S->w0_block_n, S->w0_n_tiles_n,
                                 N0.t_in, N0.t_conv, N0.relu, 0, full0b, ld0b, N0.t_conv, dev, s);
        if (rc) return rc;
        const long long pe = (long long)n * N0.t_pool * ld0b;
        pool_fwd_bf16_kernel<<<grid_for(pe, dev), 256, 0, s>>>(full0b, N0.t_conv, N0.pool, N0.t_pool, ld0b, n, x1b, ext_mask(0),
                                                               mask_of(0), T->dropout_p[0], seed_of(0));
        SB_COUNT_LAUNCH();
        SB_CHECK_LAUNCH();
    } else {
        const long long rows = (long long)n * N0.t_in;
        int rc = launch_gemm_gen(x, rows * 4, 4, 1, N0.wf, N0.ldw_f32, 1, rows, N0.cout, N0.k_f32, N0.biasf, N0.relu, full0,
                                 N0.cout, s);
        if (rc) return rc;
        const long long pe = (long long)n * N0.t_pool * N0.cout;
        pool_compact_kernel<float><<<grid_for(pe, dev), 256, 0, s>>>(full0, N0.cout, N0.t_in, N0.pool, N0.t_pool, N0.t_pool,
                                                                     N0.cout, N0.cout, n, pooled0);
        SB_COUNT_LAUNCH();
        if (T->dropout_p[0] > 0.f || ext_mask(0)) {
            dropout_kernel<<<grid_for(pe, dev), 256, 0, s>>>(pooled0, pe, ext_mask(0), mask_of(0), T->dropout_p[0], seed_of(0));
            SB_COUNT_LAUNCH();
        }
        f32_to_bf16_kernel<<<grid_for((long long)n * N0.t_pool * ld0b, dev), 256, 0, s>>>(pooled0, (long long)n * N0.t_pool,
                                                                                         N0.cout, N0.cout, x1b, ld0b);
        SB_COUNT_LAUNCH();
        SB_CHECK_LAUNCH();
    }
    std::vector<const bf16*> xin(L, nullptr);   // input activation of each layer (bf16)
    xin[1] = x1b;
    for (size_t li = 1; li < L; ++li) {
        const NetLayer& N = net->layers[li];
        const bool conv = N.type == SB_LAYER_CONV;
        const bool last = li + 1 == L;
        const int ldo = S->L[li].cout_ld;
        bf16* full = reinterpret_cast<bf16*>(base + P.full[li]);
        int rc;
        if (conv) {
            rc = launch_tc_plain(xin[li], (long long)n * N.t_in, N.a_ld, N.k_full, N.wb, N.cout, N.w_ld, N.biasb, N.block_n,
                                 N.n_tiles_n, N.t_in, N.t_conv, N.relu, 0, full, ldo, N.t_conv, dev, s);
            if (rc) return rc;
            bf16* pooled = reinterpret_cast<bf16*>(base + P.pooled[li]);
            const long long pe = (long long)n * N.t_pool * ldo;
            pool_fwd_bf16_kernel<<<grid_for(pe, dev), 256, 0, s>>>(full, N.t_conv, N.pool, N.t_pool, ldo, n, pooled, ext_mask(li),
                                                                   mask_of(li), T->dropout_p[li], seed_of(li));
            SB_COUNT_LAUNCH();
            SB_CHECK_LAUNCH();
            if (!last) xin[li + 1] = pooled;
        } else {
            // A small batch gives an FC layer a handful of output tiles (batch 64 x 925 outputs: 4) over a very long K
            // (50 880): split K over the SMs, partial sums added into an fp32 buffer, bias + ReLU in a finishing pass
            const long long out_tiles = ((n + 127) / 128) * N.n_tiles_n;
            const int kbs = (N.k_full + 63) / 64;
            const int sms = sb_sm_count(dev);
            int splits = 1;
            if (!last && out_tiles * 2 <= sms && kbs >= 32) splits = (int)std::min<long long>(kbs / 4, sms / out_tiles);
            if (splits > 1) {
                float* acc = reinterpret_cast<float*>(base + P.fc32);
                SB_CHECK_CUDA(cudaMemsetAsync(acc, 0, (size_t)n * ldo * 4, s));
                rc = launch_tc_plain(xin[li], n, N.a_ld, N.k_full, N.wb, N.cout, N.w_ld, net->zero_bias_tc, N.block_n,
                                     N.n_tiles_n, 1, 1, 0, 3, acc, ldo, 1, dev, s, splits);
                if (rc) return rc;
                bias_act_bf16_kernel<<<grid_for((long long)n * ldo, dev), 256, 0, s>>>(acc, N.biasb, n, N.cout, ldo, N.relu, full);
                SB_COUNT_LAUNCH();
                SB_CHECK_LAUNCH();
            } else {
                rc = launch_tc_plain(xin[li], n, N.a_ld, N.k_full, N.wb, N.cout, N.w_ld, N.biasb, N.block_n, N.n_tiles_n, 1, 1,
                                     N.relu, last ? 1 : 0, last ? (void*)logits : (void*)full, last ? ldl : ldo, 1, dev, s);
                if (rc) return rc;
            }
            if (!last && (T->dropout_p[li] > 0.f || ext_mask(li))) {
                dropout_bf16_kernel<<<grid_for((long long)n * ldo, dev), 256, 0, s>>>(full, (long long)n * ldo, ext_mask(li),
                                                                                      mask_of(li), T->dropout_p[li], seed_of(li));
                SB_COUNT_LAUNCH();
            }
            if (!last) xin[li + 1] = full;
        }
    }

    // ------------------------------------------------------------------ loss + dL/dz
    bf16* g_cur = reinterpret_cast<bf16*>(base + P.g_a);
    bf16* g_oth = reinterpret_cast<bf16*>(base + P.g_b);
    SB_CHECK_CUDA(cudaMemsetAsync(loss, 0, 4, s));
    // weight-gradient blocks whose wgrad reduces every tile in ONE K pass (FC layers: rows = batch <= one k-block) are
    // stored, not accumulated: everything else (split-K accumulators, bias sums, padded columns) starts from zero
    std::vector<char> wg_stores(L, 0);
    {
        size_t z0 = 0;
        for (size_t li = 1; li < L; ++li) {
            const NetLayer& N = net->layers[li];
            wg_stores[li] = sb_tc_wgrad_overwrites((long long)n * N.t_in, N.cout, N.k_f32, N.ldw_f32, dev) ? 1 : 0;
            if (!wg_stores[li]) continue;
            if (T->goff_w[li] > z0) SB_CHECK_CUDA(cudaMemsetAsync(grads + z0, 0, (T->goff_w[li] - z0) * 4, s));
            z0 = T->goff_b[li];
        }
        if (T->n_params > z0) SB_CHECK_CUDA(cudaMemsetAsync(grads + z0, 0, (T->n_params - z0) * 4, s));
    }
    {
        const int ldz = S->L[L - 1].cout_ld;
        bce_bf16_kernel<<<grid_for((long long)n * ldz, dev), 256, 0, s>>>(logits, ldl, targets, n, F, g_cur, ldz, loss);
        SB_COUNT_LAUNCH();
        SB_CHECK_LAUNCH();
    }

    // ------------------------------------------------------------------ backward, layers L-1 .. 1 on tcgen05
    for (size_t li = L - 1; li >= 1; --li) {
        const NetLayer& N = net->layers[li];
        const bool conv = N.type == SB_LAYER_CONV;
        const bool last = li + 1 == L;
        const int ldo = S->L[li].cout_ld;
        const int ch_in = conv ? N.cin : N.cin;
        const int ld_in = round_up(ch_in, 8);
        float* gw = grads + T->goff_w[li];
        float* gb = grads + T->goff_b[li];
        const bool dropped = !last && (T->dropout_p[li] > 0.f || ext_mask(li));
        const float* mk = dropped ? (ext_mask(li) ? ext_mask(li) : mask_of(li)) : nullptr;
        const bf16* full = reinterpret_cast<const bf16*>(base + P.full[li]);
        const long long rows = (long long)n * N.t_in;
        const bf16* dy;
        int rc;
        if (conv) {
            SB_CHECK_CUDA(cudaMemsetAsync(g_oth, 0, (size_t)(N.k - 1) * ldo * 2, s));
            bf16* dfull = g_oth + (size_t)(N.k - 1) * ldo;
            pool_relu_bwd_bf16_kernel<<<grid_for(rows * ldo, dev), 256, 0, s>>>(full, N.t_conv, g_cur, mk, N.t_in, N.pool,
                                                                               N.t_pool, ldo, n, dfull);
            SB_COUNT_LAUNCH();
            SB_CHECK_LAUNCH();
            dy = dfull;
        } else {
            if (dropped || (N.relu && !last)) {
                fc_bwd_bf16_kernel<<<grid_for(rows * ldo, dev), 256, 0, s>>>(g_cur, full, mk, (N.relu && !last) ? 1 : 0, rows * ldo);
                SB_COUNT_LAUNCH();
            }
            dy = g_cur;
        }
        // wgrad on tcgen05: G[kk][o] += sum_r Xin[r*ld_in + kk] dY[r][o]
        rc = sb_tc_wgrad_launch(dy, ldo, xin[li], conv ? ld_in : N.a_ld, rows, rows, N.cout, N.k_f32, gw, N.ldw_f32, dev, s,
                                wg_stores[li]);
        if (rc) return rc;
        launch_colsum_bf16(dy, rows, ldo, N.cout, gb, dev, s);
        SB_COUNT_LAUNCH();
        SB_CHECK_LAUNCH();
        // dgrad on tcgen05 -> gradient w.r.t. this layer's input (previous layer's final output)
        if (conv) {
            rc = launch_tc_plain(g_oth, rows + (N.k - 1), ldo, N.k * ldo, S->L[li].wt, N.cin, S->L[li].wt_ld, net->zero_bias_tc,
                                 S->L[li].dg_block_n, S->L[li].dg_n_tiles_n, 1, 1, 0, 0, g_cur, ld_in, 1, dev, s);
            if (rc) return rc;
        } else {
            rc = launch_tc_plain(dy, rows, ldo, ldo, S->L[li].wt, N.cin, S->L[li].wt_ld, net->zero_bias_tc,
                                 S->L[li].dg_block_n, S->L[li].dg_n_tiles_n, 1, 1, 0, 0, g_oth, N.a_ld, 1, dev, s);
            if (rc) return rc;
            bf16* t = g_cur; g_cur = g_oth; g_oth = t;
        }
        // recorded after the layer's dgrad (which reads its weights): gradients and weights of layers >= li are final /
        // unread from here on, so their all-reduce AND optimiser step may overlap the rest of the backward pass
        if (T->grad_event[li]) SB_CHECK_CUDA(cudaEventRecord(T->grad_event[li], s));
        if (li == 1) break;
    }

    // ------------------------------------------------------------------ layer 0 backward
    if (S->l0_tc) {
        // dfull over all t_in rows of every sequence (rows >= t_conv are zero, so the input windows that run into the next
        // sequence contribute nothing), weight gradient on tcgen05 over the 8-channel rows, bias gradient as a column sum
        const long long rows = (long long)n * N0.t_in;
        const bool dropped0 = T->dropout_p[0] > 0.f || ext_mask(0);
        const float* mk0 = dropped0 ? (ext_mask(0) ? ext_mask(0) : mask_of(0)) : nullptr;
        bf16* dfull0 = g_oth;
        pool_relu_bwd_bf16_kernel<<<grid_for(rows * ld0b, dev), 256, 0, s>>>(full0b, N0.t_conv, g_cur, mk0, N0.t_in, N0.pool,
                                                                            N0.t_pool, ld0b, n, dfull0);
        SB_COUNT_LAUNCH();
        float* g8 = reinterpret_cast<float*>(base + P.g8);
        SB_CHECK_CUDA(cudaMemsetAsync(g8, 0, (size_t)N0.k * 8 * N0.ldw_f32 * 4, s));
        int rc = sb_tc_wgrad_launch(dfull0, ld0b, x8, 8, rows + N0.k, rows, N0.cout, N0.k * 8, g8, N0.ldw_f32, dev, s);
        if (rc) return rc;
        w0_grad_unpad_kernel<<<grid_for((long long)N0.k * 4 * N0.cout, dev), 256, 0, s>>>(g8, N0.ldw_f32, N0.k, N0.cout,
                                                                                         grads + T->goff_w[0], N0.ldw_f32);
        SB_COUNT_LAUNCH();
        launch_colsum_bf16(dfull0, rows, ld0b, N0.cout, grads + T->goff_b[0], dev, s);
        SB_COUNT_LAUNCH();
        SB_CHECK_LAUNCH();
    } else {
        const long long pe = (long long)n * N0.t_pool * N0.cout;
        float* g32 = reinterpret_cast<float*>(base + P.g32);
        float* dpool32 = g32;                                     // (n, t_pool, cout) fp32
        float* dfull32 = g32 + (((size_t)pe + 63) / 64) * 64;     // (n, t_in, cout) fp32
        bf16_to_f32_kernel<<<grid_for(pe, dev), 256, 0, s>>>(g_cur, (long long)n * N0.t_pool, N0.cout, ld0b, dpool32, N0.cout);
        SB_COUNT_LAUNCH();
        if (T->dropout_p[0] > 0.f || ext_mask(0)) {
            mul_kernel<<<grid_for(pe, dev), 256, 0, s>>>(dpool32, ext_mask(0) ? ext_mask(0) : mask_of(0), pe);
            SB_COUNT_LAUNCH();
        }
        const long long rows = (long long)n * N0.t_in;
        pool_relu_bwd_kernel<<<grid_for(rows * N0.cout, dev), 256, 0, s>>>(full0, dpool32, N0.t_in, N0.pool, N0.t_pool, N0.cout, n,
                                                                           dfull32);
        SB_COUNT_LAUNCH();
        const int rpb = 256;
        dim3 grid((unsigned)((rows + rpb - 1) / rpb), (unsigned)((N0.cout + 127) / 128));
        conv1_wgrad_kernel<<<grid, 128, 0, s>>>(x, rows, dfull32, N0.cout, rows, N0.k, N0.cout, rpb, grads + T->goff_w[0],
                                                N0.ldw_f32, grads + T->goff_b[0]);
        SB_COUNT_LAUNCH();
        SB_CHECK_LAUNCH();
    }
    return SB_OK;
}


template <typename GT>
static bool sgd_pack_layer_tc(sb_net* net, size_t li, const GT* g, float lr, float momentum, float weight_decay,
                              int first, cudaStream_t s) {
    if (!net->train_tc || li == 0 || li >= net->train_tc->L.size()) return false;
    const char* off = getenv("SB_SGD_FUSED_PACK");   // "0": plain SGD kernel, repack at the start of the next step
    if (off && off[0] == '0') return false;
    const NetLayer& N = net->layers[li];
    const TcTrainLayer& T = net->train_tc->L[li];
    const bool conv = N.type == SB_LAYER_CONV;
    const int ch = conv ? N.cin : (N.fc_prev_c > 0 ? N.fc_prev_c : N.cin);
    const unsigned kbs = (unsigned)((N.k_f32 + 31) / 32), obs = (unsigned)((N.ldw_f32 + 31) / 32);
    const char* ord = getenv("SB_SGD_ORDER");                  // "0": 2-D grid, row blocks fastest (r02 first half)
    const bool rows_first = ord && ord[0] == '0';
    sgd_pack_kernel<<<rows_first ? dim3(kbs, obs) : dim3(kbs * obs), 256, 0, s>>>(
        N.wf, net->train->mom_w[li], g, N.ldw_f32, (int)N.k_f32, N.cout, lr, momentum, weight_decay, first, ch, round_up(ch, 8),
        N.wb, N.w_ld, conv ? N.k : 1, N.cin, T.cout_ld, T.wt, T.wt_ld, rows_first ? 0 : (int)obs);
    return true;
}

static void sgd_pack_done_tc(sb_net* net) {
    const char* off = getenv("SB_SGD_FUSED_PACK");
    if (net->train_tc) net->train_tc->packed_fresh = !(off && off[0] == '0');
}

// momentum buffers, dgrad operands and the zero bias of the training paths (sb_net_destroy)
static void free_train_states(sb_net* net) {
    if (net->train) {
        for (float* p : net->train->mom_w) cudaFree(p);
        for (float* p : net->train->mom_b) cudaFree(p);
        delete net->train;
        net->train = nullptr;
    }
    if (net->train_tc) {
        for (auto& T : net->train_tc->L) cudaFree(T.wt);
        cudaFree(net->train_tc->w0b);
        cudaFree(net->train_tc->bias0);
        delete net->train_tc;
        net->train_tc = nullptr;
    }
    cudaFree(net->zero_bias_tc);
    net->zero_bias_tc = nullptr;
}
