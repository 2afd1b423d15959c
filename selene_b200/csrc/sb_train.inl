// selene_b200 — training step of the conv stacks (fp32, CUDA-core path).  Included by sb_net.cu.
//
// Reference: selene_sdk/train_model.py:483-496 — forward in train mode (dropout active), BCELoss
// (mean over B*F, log clamped at -100 as torch does), zero_grad, backward, optimizer.step — with
// models/deepsea.py:60-74 (SGD, momentum 0.9, weight decay 1e-6 added to the gradient, no Nesterov).
//
// Everything is expressed with the same channel-last layouts as inference, so every gradient is again a
// GEMM over a strided or overlapping view and no im2col / col2im buffer exists:
//   forward   Y[r][o]   = sum_k  X[r*C + k]        W[k][o]          (k = tap*C + c; overlapping rows)
//   wgrad     dW[k][o]  = sum_r  X[r*C + k]        dY[r][o]         (rows whose window is invalid have dY = 0)
//   dgrad     dX[r][c]  = sum_kk dY[(r-(taps-1))*O + kk] Wf[kk][c]  (kk = tap'*O + o, Wf = taps-flipped W^T;
//                                                                    the zero rows of dY between sequences
//                                                                    are exactly the padding a full
//                                                                    correlation needs)
// This fp32 path is the parity path of round 1; moving the three GEMMs to tcgen05 is the next step.

namespace {

// C[m][n] = sum_k A(m,k) * B(k,n) (+ bias[n]) (ReLU), A(m,k) = A[m*sam + k*sak] (0 beyond a_elems),
// B(k,n) = B[k*sbk + n*sbn].  64x64x16 tiles, 4x4 per thread.  Tile loads pick the thread mapping that
// is contiguous in memory for the given strides.
__global__ void __launch_bounds__(256) gemm_gen_f32_kernel(const float* __restrict__ A, long long a_elems, long long sam,
                                                           long long sak, const float* __restrict__ B, long long sbk,
                                                           long long sbn, long long M, int N, long long K,
                                                           const float* __restrict__ bias, int relu,
                                                           float* __restrict__ C, long long ldc) {
    __shared__ float As[FBK][FBM + 4];
    __shared__ float Bs[FBK][FBN + 4];
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * FBM;
    const int n0 = blockIdx.y * FBN;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const bool a_k_fast = sak == 1;
    const bool b_n_fast = sbn == 1;
    for (long long k0 = 0; k0 < K; k0 += FBK) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = tid + 256 * e;  // 0..1023 over the 64 x 16 tile
            const int mm = a_k_fast ? idx >> 4 : idx & 63;
            const int kk = a_k_fast ? idx & 15 : idx >> 6;
            const long long m = m0 + mm, k = k0 + kk;
            float v = 0.f;
            if (m < M && k < K) {
                const long long off = m * sam + k * sak;
                if (off < a_elems) v = __ldg(A + off);
            }
            As[kk][mm] = v;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = tid + 256 * e;
            const int nn = b_n_fast ? idx & 63 : idx >> 4;
            const int kk = b_n_fast ? idx >> 6 : idx & 15;
            const long long k = k0 + kk;
            const int n = n0 + nn;
            Bs[kk][nn] = (k < K && n < N) ? __ldg(B + k * sbk + (long long)n * sbn) : 0.f;
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
        const long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
            if (relu) v = fmaxf(v, 0.f);
            C[m * ldc + n] = v;
        }
    }
}

int launch_gemm_gen(const float* A, long long a_elems, long long sam, long long sak, const float* B, long long sbk,
                    long long sbn, long long M, int N, long long K, const float* bias, int relu, float* C,
                    long long ldc, cudaStream_t s) {
    dim3 grid((unsigned)((M + FBM - 1) / FBM), (unsigned)((N + FBN - 1) / FBN));
    gemm_gen_f32_kernel<<<grid, 256, 0, s>>>(A, a_elems, sam, sak, B, sbk, sbn, M, N, K, bias, relu, C, ldc);
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    return SB_OK;
}

// counter-based uniform in [0,1) (splitmix64 finaliser) for dropout masks
__device__ __forceinline__ float hash_uniform(unsigned long long seed, unsigned long long idx) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

// y *= mask.  mask_in given: use it; else draw keep = u >= p, scale 1/(1-p), and write mask_out.
__global__ void __launch_bounds__(256) dropout_kernel(float* __restrict__ y, long long n, const float* __restrict__ mask_in,
                                                      float* __restrict__ mask_out, float p, unsigned long long seed) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float scale = 1.f / (1.f - p);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float m;
        if (mask_in) m = mask_in[i];
        else { m = hash_uniform(seed, (unsigned long long)i) >= p ? scale : 0.f; mask_out[i] = m; }
        y[i] *= m;
    }
}

__global__ void __launch_bounds__(256) mul_kernel(float* __restrict__ y, const float* __restrict__ m, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] *= m[i];
}

// dY *= (Y > 0)
__global__ void __launch_bounds__(256) relu_bwd_kernel(float* __restrict__ dy, const float* __restrict__ y, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (!(y[i] > 0.f)) dy[i] = 0.f;
}

// BCE on sigmoid(z): loss += sum(-(y log p + (1-y) log(1-p))) / N with logs clamped at -100 (torch BCELoss);
// dz = (p - y) / N
__global__ void __launch_bounds__(256) bce_kernel(const float* __restrict__ z, int ldz, const float* __restrict__ y,
                                                  long long n, int F, float* __restrict__ dz, int lddz,
                                                  float* __restrict__ loss) {
    const long long total = n * F;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float inv = 1.f / (float)total;
    float local = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / F;
        const int f = (int)(i - r * F);
        const float p = sigmoid_f32(z[r * ldz + f]);
        const float t = y[i];
        const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(logf(1.f - p), -100.f);
        local += -(t * lp + (1.f - t) * lq);
        dz[r * lddz + f] = (p - t) * inv;
    }
    local = local * inv;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(loss, local);
}

// dFull[s][t][c] = (t is the first maximum of its pool window and Full > 0) ? dPool[s][t/pool][c] : 0;
// rows t >= t_pool*pool (dropped by the pool, or invalid windows) get 0.
__global__ void __launch_bounds__(256) pool_relu_bwd_kernel(const float* __restrict__ full, const float* __restrict__ dpool,
                                                            int t_in, int pool, int t_pool, int C, long long n_seq,
                                                            float* __restrict__ dfull) {
    const long long total = n_seq * t_in * C;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int c = (int)(i % C);
        const long long rt = i / C;
        const int t = (int)(rt % t_in);
        const long long s = rt / t_in;
        float g = 0.f;
        const int tp = t / pool;
        if (tp < t_pool) {
            const float v = full[i];
            if (v > 0.f) {
                bool first_max = true;
                const long long w0 = (s * t_in + (long long)tp * pool) * C + c;
                for (int q = 0; q < pool; ++q) {
                    const float u = full[w0 + (long long)q * C];
                    const int tq = tp * pool + q;
                    if (u > v || (u == v && tq < t)) first_max = false;
                }
                if (first_max) g = dpool[(s * t_pool + tp) * C + c];
            }
        }
        dfull[i] = g;
    }
}

// Wf[(tap'*O + o)*ldc + c] = W[((taps-1-tap')*C + c)*ldw + o]
__global__ void __launch_bounds__(256) flip_weights_kernel(const float* __restrict__ w, int ldw, int taps, int C, int O,
                                                           float* __restrict__ wf, int ldc) {
    const long long total = (long long)taps * O * C;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int c = (int)(i % C);
        const long long r = i / C;
        const int o = (int)(r % O);
        const int tp = (int)(r / O);
        wf[r * ldc + c] = w[((long long)(taps - 1 - tp) * C + c) * ldw + o];
    }
}

// db[o] += sum_r dY[r][o]  (db zeroed by the caller; grid.y splits the rows, partials are added atomically)
constexpr int kColsumRows = 1024;
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dy, long long rows, int ld, int n_cols,
                                                     float* __restrict__ db) {
    const int o = blockIdx.x * 32 + (threadIdx.x & 31);
    const int part = threadIdx.x >> 5;  // 8 row partitions
    __shared__ float red[8][33];
    const long long r0 = (long long)blockIdx.y * kColsumRows;
    const long long r1 = r0 + kColsumRows < rows ? r0 + kColsumRows : rows;
    float s = 0.f;
    if (o < n_cols)
        for (long long r = r0 + part; r < r1; r += 8) s += dy[r * ld + o];
    red[part][threadIdx.x & 31] = s;
    __syncthreads();
    if (part == 0 && o < n_cols) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x & 31];
        if (t != 0.f) atomicAdd(db + o, t);
    }
}

// torch.optim.SGD: g += wd*w; v = mom*v + g (v = g on the first step); w -= lr*v
// gradient element as fp32: the flat gradient vector is fp32, or bf16 when it comes straight out of a bf16 all-reduce
__device__ __forceinline__ float grad_at(const float* g, long long i) { return g[i]; }
__device__ __forceinline__ float grad_at(const __nv_bfloat16* g, long long i) { return __bfloat162float(g[i]); }

template <typename GT>
__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ w, float* __restrict__ v, const GT* __restrict__ g,
                                                  long long n, float lr, float mom, float wd, int first) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gi = grad_at(g, i) + wd * w[i];
        const float vi = first ? gi : mom * v[i] + gi;
        v[i] = vi;
        w[i] -= lr * vi;
    }
}

struct TrainLayerBufs {
    float* x;      // input activation (rows_in x cin), after the previous layer's dropout
    float* full;   // conv: post-ReLU pre-pool output (rows_in x cout); fc: post-activation output (n x cout)
    float* pooled; // conv: pooled (+dropout) output = next layer's x
    float* mask;   // dropout mask of this layer's output (or null)
};

size_t align256(size_t b) { return (b + 255) / 256 * 256; }

}  // namespace

struct sb_train_state {
    std::vector<float*> mom_w, mom_b;      // momentum buffers per layer
    std::vector<size_t> goff_w, goff_b;    // offsets (floats) into the flat gradient vector
    std::vector<float> dropout_p;
    std::vector<cudaEvent_t> grad_event;   // per layer, recorded when its gradients (and those of later layers) are complete
    size_t n_params;
    long long steps;
};

extern "C" int sb_net_train_init(sb_net* net) {
    SB_REQUIRE(net, "sb_net_train_init: null handle");
    SB_REQUIRE(!net->has_lstm, "sb_net_train_init: BiLSTM networks are not trainable yet");
    if (net->train) return SB_OK;
    SB_CHECK_CUDA(cudaSetDevice(net->device));
    sb_train_state* T = new sb_train_state();
    size_t off = 0;
    for (auto& N : net->layers) {
        const size_t nw = (size_t)N.k_f32 * N.ldw_f32, nb = (size_t)N.ldw_f32;
        float *mw = nullptr, *mb = nullptr;
        SB_CHECK_CUDA(cudaMalloc(&mw, nw * 4));
        SB_CHECK_CUDA(cudaMalloc(&mb, nb * 4));
        SB_CHECK_CUDA(cudaMemset(mw, 0, nw * 4));
        SB_CHECK_CUDA(cudaMemset(mb, 0, nb * 4));
        T->mom_w.push_back(mw);
        T->mom_b.push_back(mb);
        T->goff_w.push_back(off); off += nw;
        T->goff_b.push_back(off); off += nb;
        T->dropout_p.push_back(0.f);
        T->grad_event.push_back(nullptr);
    }
    T->n_params = off;
    T->steps = 0;
    net->train = T;
    return SB_OK;
}

extern "C" size_t sb_net_n_params(const sb_net* net) { return (net && net->train) ? net->train->n_params : 0; }

extern "C" size_t sb_net_param_offset(const sb_net* net, int layer) {
    if (!net || !net->train || layer < 0 || layer >= (int)net->layers.size()) return 0;
    return net->train->goff_w[(size_t)layer];
}

extern "C" int sb_net_set_grad_event(sb_net* net, int layer, void* cuda_event) {
    SB_REQUIRE(net && net->train, "sb_net_set_grad_event: call sb_net_train_init first");
    SB_REQUIRE(layer >= 0 && layer < (int)net->layers.size(), "sb_net_set_grad_event: bad layer %d", layer);
    net->train->grad_event[(size_t)layer] = (cudaEvent_t)cuda_event;
    return SB_OK;
}

extern "C" int sb_net_set_dropout(sb_net* net, int layer, float p) {
    SB_REQUIRE(net && net->train, "sb_net_set_dropout: call sb_net_train_init first");
    SB_REQUIRE(layer >= 0 && layer < (int)net->layers.size() && p >= 0.f && p < 1.f, "sb_net_set_dropout: bad arguments");
    net->train->dropout_p[(size_t)layer] = p;
    return SB_OK;
}

namespace {
struct TrainPlan {
    std::vector<size_t> x, full, pooled, mask;  // byte offsets per layer
    size_t dA, dB, wflip, logits, total;
};

TrainPlan plan_train(const sb_net* net, long long n) {
    TrainPlan P;
    size_t off = 0;
    size_t max_act = 0, max_flip = 0;
    for (size_t li = 0; li < net->layers.size(); ++li) {
        const NetLayer& N = net->layers[li];
        const bool conv = N.type == SB_LAYER_CONV;
        const size_t rows_in = (size_t)n * N.t_in;
        const size_t x_b = rows_in * (conv ? N.cin : N.cin) * 4;          // fc: t_in == 1, cin = flattened size
        const size_t full_b = rows_in * N.cout * 4;
        const size_t pooled_b = (size_t)n * N.t_pool * N.cout * 4;
        P.x.push_back(off);                                               // inputs alias the producer's output
        P.full.push_back(off); off += align256(full_b);
        P.pooled.push_back(off); off += align256(conv ? pooled_b : 0);
        P.mask.push_back(off); off += align256(conv ? pooled_b : full_b);
        max_act = std::max(max_act, std::max(full_b + (size_t)N.k * N.cout * 4 * 2, x_b));
        if (conv && li > 0) max_flip = std::max(max_flip, (size_t)N.k * N.cout * round_up(N.cin, 4) * 4);
    }
    P.dA = off; off += align256(max_act + 4096);
    P.dB = off; off += align256(max_act + 4096);
    P.wflip = off; off += align256(max_flip);
    P.logits = off; off += align256((size_t)n * round_up(net->n_out, 4) * 4);
    P.total = off + 1024;
    return P;
}
}  // namespace

extern "C" size_t sb_net_train_workspace_bytes(const sb_net* net, int64_t batch) {
    if (!net || batch <= 0) return 0;
    return plan_train(net, batch).total;
}

// One forward + backward over a batch of generic (n, L, 4) fp32 inputs.  grads: flat fp32 vector of
// sb_net_n_params floats (layout: per layer W as (K x ldw) then bias (ldw)); loss: one float (mean BCE).
extern "C" int sb_net_forward_backward(sb_net* net, const float* x, const float* targets, int64_t n,
                                       const float* const* dropout_masks, unsigned long long dropout_seed,
                                       void* workspace, size_t workspace_bytes, float* grads, float* loss, void* stream) {
    SB_REQUIRE(net && net->train, "sb_net_forward_backward: call sb_net_train_init first");
    SB_REQUIRE(x && targets && grads && loss && n > 0, "sb_net_forward_backward: bad arguments");
    const TrainPlan P = plan_train(net, n);
    SB_REQUIRE(workspace && workspace_bytes >= P.total, "sb_net_forward_backward: workspace too small: need %zu, have %zu",
               P.total, workspace_bytes);
    SB_CHECK_CUDA(cudaSetDevice(net->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int dev = net->device;
    sb_train_state* T = net->train;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    const size_t L = net->layers.size();
    std::vector<TrainLayerBufs> B(L);
    for (size_t li = 0; li < L; ++li) {
        B[li].x = li == 0 ? const_cast<float*>(x) : reinterpret_cast<float*>(base + P.x[li]);
        B[li].full = reinterpret_cast<float*>(base + P.full[li]);
        B[li].pooled = reinterpret_cast<float*>(base + P.pooled[li]);
        B[li].mask = reinterpret_cast<float*>(base + P.mask[li]);
    }
    float* logits = reinterpret_cast<float*>(base + P.logits);
    const int ldl = round_up(net->n_out, 4);

    // ------------------------------------------------------------------ forward (train mode)
    for (size_t li = 0; li < L; ++li) {
        const NetLayer& N = net->layers[li];
        const bool conv = N.type == SB_LAYER_CONV;
        const bool last = li + 1 == L;
        const long long rows = (long long)n * N.t_in;
        float* out = last ? logits : B[li].full;
        const int ldo = last ? ldl : N.cout;
        int rc = launch_gemm_gen(B[li].x, rows * N.cin, N.cin, 1, N.wf, N.ldw_f32, 1, rows, N.cout, N.k_f32, N.biasf,
                                 N.relu, out, ldo, s);
        if (rc) return rc;
        float* y = out;
        long long y_elems = rows * N.cout;
        if (conv) {
            const long long total = (long long)n * N.t_pool * N.cout;
            pool_compact_kernel<float><<<grid_for(total, dev), 256, 0, s>>>(B[li].full, N.cout, N.t_in, N.pool, N.t_pool,
                                                                            N.t_pool, N.cout, N.cout, n, B[li].pooled);
            SB_COUNT_LAUNCH();
            SB_CHECK_LAUNCH();
            y = B[li].pooled;
            y_elems = total;
        }
        const float p = T->dropout_p[li];
        const float* m_in = dropout_masks ? dropout_masks[li] : nullptr;
        if (!last && (p > 0.f || m_in)) {
            dropout_kernel<<<grid_for(y_elems, dev), 256, 0, s>>>(y, y_elems, m_in, B[li].mask, p,
                                                                  dropout_seed * 1315423911ull + li);
            SB_COUNT_LAUNCH();
            SB_CHECK_LAUNCH();
        }
        if (!last) B[li + 1].x = y;   // conv: (n, t_pool, cout) rows; the first FC flattens the same memory
    }

    // ------------------------------------------------------------------ loss + dL/dz
    float* g_cur = reinterpret_cast<float*>(base + P.dA);   // gradient w.r.t. the current layer's final output
    float* g_oth = reinterpret_cast<float*>(base + P.dB);
    SB_CHECK_CUDA(cudaMemsetAsync(loss, 0, 4, s));
    SB_CHECK_CUDA(cudaMemsetAsync(grads, 0, T->n_params * 4, s));   // padded columns must stay zero
    bce_kernel<<<grid_for((long long)n * net->n_out, dev), 256, 0, s>>>(logits, ldl, targets, n, net->n_out, g_cur,
                                                                        net->layers[L - 1].cout, loss);
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();

    // ------------------------------------------------------------------ backward
    for (size_t li = L; li-- > 0;) {
        const NetLayer& N = net->layers[li];
        const bool conv = N.type == SB_LAYER_CONV;
        const bool last = li + 1 == L;
        const long long rows = (long long)n * N.t_in;
        float* gw = grads + T->goff_w[li];
        float* gb = grads + T->goff_b[li];
        const float p = T->dropout_p[li];
        const float* ext_mask = dropout_masks ? dropout_masks[li] : nullptr;
        const bool dropped = !last && (p > 0.f || ext_mask);
        const long long out_elems = conv ? (long long)n * N.t_pool * N.cout : rows * N.cout;
        if (dropped) {
            mul_kernel<<<grid_for(out_elems, dev), 256, 0, s>>>(g_cur, ext_mask ? ext_mask : B[li].mask, out_elems);
            SB_COUNT_LAUNCH();
        }
        const float* dy;   // gradient w.r.t. the layer's pre-pool output, (rows x cout)
        if (conv) {
            // into the other buffer, behind (taps-1) zero rows so that the dgrad view may start before row 0
            SB_CHECK_CUDA(cudaMemsetAsync(g_oth, 0, (size_t)(N.k - 1) * N.cout * 4, s));
            float* dfull = g_oth + (size_t)(N.k - 1) * N.cout;
            pool_relu_bwd_kernel<<<grid_for(rows * N.cout, dev), 256, 0, s>>>(B[li].full, g_cur, N.t_in, N.pool, N.t_pool,
                                                                              N.cout, n, dfull);
            SB_COUNT_LAUNCH();
            SB_CHECK_LAUNCH();
            dy = dfull;
        } else {
            if (N.relu && !last) {
                relu_bwd_kernel<<<grid_for(out_elems, dev), 256, 0, s>>>(g_cur, B[li].full, out_elems);
                SB_COUNT_LAUNCH();
            }
            dy = g_cur;
        }
        // wgrad: dW[k][o] = sum_r X[r*cin + k] dY[r][o];  db[o] = sum_r dY[r][o]
        int rc = launch_gemm_gen(B[li].x, rows * N.cin, 1, N.cin, dy, N.cout, 1, N.k_f32, N.cout, rows, nullptr, 0, gw,
                                 N.ldw_f32, s);
        if (rc) return rc;
        colsum_kernel<<<dim3((N.cout + 31) / 32, (unsigned)((rows + kColsumRows - 1) / kColsumRows)), 256, 0, s>>>(
            dy, rows, N.cout, N.cout, gb);
        SB_COUNT_LAUNCH();
        SB_CHECK_LAUNCH();
        if (li == 0) {
            if (T->grad_event[li]) SB_CHECK_CUDA(cudaEventRecord(T->grad_event[li], s));
            break;
        }
        if (conv) {
            // dgrad as a forward-style GEMM over the shifted view of dfull with the taps-flipped W^T; the
            // result (gradient w.r.t. the previous layer's output) overwrites g_cur, which is consumed
            float* wflip = reinterpret_cast<float*>(base + P.wflip);
            const int ldc = round_up(N.cin, 4);
            flip_weights_kernel<<<grid_for((long long)N.k * N.cout * N.cin, dev), 256, 0, s>>>(N.wf, N.ldw_f32, N.k, N.cin,
                                                                                             N.cout, wflip, ldc);
            SB_COUNT_LAUNCH();
            rc = launch_gemm_gen(g_oth, ((long long)(N.k - 1) + rows) * N.cout, N.cout, 1, wflip, ldc, 1, rows, N.cin,
                                 (long long)N.k * N.cout, nullptr, 0, g_cur, N.cin, s);
            if (rc) return rc;
        } else {
            // dX[b][i] = sum_o dY[b][o] W[i][o]   (W stored (K x ldw): B(k=o, n=i) = wf[i*ldw + o])
            rc = launch_gemm_gen(dy, rows * N.cout, N.cout, 1, N.wf, 1, N.ldw_f32, rows, N.cin, N.cout, nullptr, 0, g_oth,
                                 N.cin, s);
            if (rc) return rc;
            float* t = g_cur; g_cur = g_oth; g_oth = t;
        }
        // recorded after the layer's dgrad: from here on neither the gradients NOR the weights of layers >= li are touched
        // by this pass, so an optimiser step of those layers may run concurrently (sb_net_sgd_step_layers)
        if (T->grad_event[li]) SB_CHECK_CUDA(cudaEventRecord(T->grad_event[li], s));
    }
    return SB_OK;
}

// sb_train_tc.inl: SGD update of layer li's weights fused with the bf16 repacking the tcgen05 training step needs
// (returns false when the tensor-core training state does not exist)
template <typename GT>
static bool sgd_pack_layer_tc(sb_net* net, size_t li, const GT* g, float lr, float momentum, float weight_decay,
                              int first, cudaStream_t s);
static void sgd_pack_done_tc(sb_net* net);

extern "C" int sb_net_sgd_step_layers(sb_net* net, const void* grads, int grads_bf16, float lr, float momentum,
                                      float weight_decay, int layer_begin, int layer_end, int finish, void* stream) {
    SB_REQUIRE(net && net->train && grads, "sb_net_sgd_step_layers: call sb_net_train_init first");
    SB_REQUIRE(layer_begin >= 0 && layer_begin <= layer_end && layer_end <= (int)net->layers.size(),
               "sb_net_sgd_step_layers: bad layer range [%d, %d)", layer_begin, layer_end);
    SB_CHECK_CUDA(cudaSetDevice(net->device));
    cudaStream_t s = (cudaStream_t)stream;
    sb_train_state* T = net->train;
    const int first = T->steps == 0;
    const float* g32 = reinterpret_cast<const float*>(grads);
    const __nv_bfloat16* g16 = reinterpret_cast<const __nv_bfloat16*>(grads);
    for (size_t li = (size_t)layer_begin; li < (size_t)layer_end; ++li) {
        NetLayer& N = net->layers[li];
        const long long nw = (long long)N.k_f32 * N.ldw_f32, nb = N.ldw_f32;
        const bool packed = li > 0 && (grads_bf16 ? sgd_pack_layer_tc(net, li, g16 + T->goff_w[li], lr, momentum, weight_decay, first, s)
                                                  : sgd_pack_layer_tc(net, li, g32 + T->goff_w[li], lr, momentum, weight_decay, first, s));
        if (!packed) {
            if (grads_bf16)
                sgd_kernel<<<grid_for(nw, net->device), 256, 0, s>>>(N.wf, T->mom_w[li], g16 + T->goff_w[li], nw, lr, momentum,
                                                                     weight_decay, first);
            else
                sgd_kernel<<<grid_for(nw, net->device), 256, 0, s>>>(N.wf, T->mom_w[li], g32 + T->goff_w[li], nw, lr, momentum,
                                                                     weight_decay, first);
        }
        if (grads_bf16)
            sgd_kernel<<<grid_for(nb, net->device), 256, 0, s>>>(N.biasf, T->mom_b[li], g16 + T->goff_b[li], nb, lr, momentum,
                                                                 weight_decay, first);
        else
            sgd_kernel<<<grid_for(nb, net->device), 256, 0, s>>>(N.biasf, T->mom_b[li], g32 + T->goff_b[li], nb, lr, momentum,
                                                                 weight_decay, first);
        SB_COUNT_LAUNCH();
        SB_COUNT_LAUNCH();
    }
    SB_CHECK_LAUNCH();
    if (finish) {
        sgd_pack_done_tc(net);
        T->steps += 1;
        net->inference_weights_stale = true;
    }
    return SB_OK;
}

extern "C" int sb_net_sgd_step(sb_net* net, const float* grads, float lr, float momentum, float weight_decay,
                               void* stream) {
    SB_REQUIRE(net && net->train && grads, "sb_net_sgd_step: call sb_net_train_init first");
    return sb_net_sgd_step_layers(net, grads, 0, lr, momentum, weight_decay, 0, (int)net->layers.size(), 1, stream);
}

// device (k_f32 x ldw_f32, K index = tap*cin + channel; FC after a conv stack position-major) <-> torch layout
// (conv (cout, cin, k), fc (cout, cin)); `to_host` selects the direction
static void layer_layout_copy(const NetLayer& N, std::vector<float>& dev, float* torch_w, bool to_host) {
    auto xfer = [&](size_t di, size_t ti) {
        if (to_host) torch_w[ti] = dev[di];
        else dev[di] = torch_w[ti];
    };
    if (N.type == SB_LAYER_CONV) {
        for (int o = 0; o < N.cout; ++o)
            for (int c = 0; c < N.cin; ++c)
                for (int j = 0; j < N.k; ++j)
                    xfer(((size_t)j * N.cin + c) * N.ldw_f32 + o, ((size_t)o * N.cin + c) * N.k + j);
    } else {
        // the FC after the conv stack was packed position-major: K index = t*C + c  <->  torch index c*T + t
        const int C = N.fc_prev_c, Tn = N.fc_prev_t;
        for (int o = 0; o < N.cout; ++o)
            for (int kk = 0; kk < N.cin; ++kk) {
                size_t torch_idx = (size_t)kk;
                if (C > 0 && Tn > 1) { const int t = kk / C, c = kk % C; torch_idx = (size_t)c * Tn + t; }
                xfer((size_t)kk * N.ldw_f32 + o, (size_t)o * N.cin + torch_idx);
            }
    }
}

static int get_layer_tensor(const sb_net* net, int layer, const float* src_w, const float* src_b, float* w_host,
                            float* b_host) {
    const NetLayer& N = net->layers[(size_t)layer];
    SB_CHECK_CUDA(cudaSetDevice(net->device));
    SB_CHECK_CUDA(cudaDeviceSynchronize());
    const size_t nw = (size_t)N.k_f32 * N.ldw_f32;
    std::vector<float> w(nw), b((size_t)N.ldw_f32);
    SB_CHECK_CUDA(cudaMemcpy(w.data(), src_w, nw * 4, cudaMemcpyDeviceToHost));
    SB_CHECK_CUDA(cudaMemcpy(b.data(), src_b, (size_t)N.ldw_f32 * 4, cudaMemcpyDeviceToHost));
    for (int o = 0; o < N.cout; ++o) b_host[o] = b[(size_t)o];
    layer_layout_copy(N, w, w_host, true);
    return SB_OK;
}

// Current fp32 parameters of one layer in torch layout: conv (cout, cin, k), fc (cout, cin); bias (cout).
// `from_grads` != NULL reads that flat gradient vector instead of the weights (same layout mapping).
extern "C" int sb_net_get_layer_params(const sb_net* net, int layer, const float* from_grads, float* w_host,
                                       float* b_host) {
    SB_REQUIRE(net && layer >= 0 && layer < (int)net->layers.size() && w_host && b_host, "sb_net_get_layer_params: bad arguments");
    const NetLayer& N = net->layers[(size_t)layer];
    SB_REQUIRE(N.type != SB_LAYER_BILSTM, "sb_net_get_layer_params: BiLSTM layers are not exportable yet");
    SB_REQUIRE(!from_grads || net->train, "sb_net_get_layer_params: gradients need sb_net_train_init");
    return get_layer_tensor(net, layer, from_grads ? from_grads + net->train->goff_w[(size_t)layer] : N.wf,
                            from_grads ? from_grads + net->train->goff_b[(size_t)layer] : N.biasf, w_host, b_host);
}

// SGD momentum buffers of one layer in torch layout (the ``momentum_buffer`` entries of torch.optim.SGD's state:
// what a Selene checkpoint stores under "optimizer", train_model.py:430-437).  All zeros before the first step.
extern "C" int sb_net_get_layer_momentum(const sb_net* net, int layer, float* w_host, float* b_host) {
    SB_REQUIRE(net && net->train && layer >= 0 && layer < (int)net->layers.size() && w_host && b_host,
               "sb_net_get_layer_momentum: bad arguments (sb_net_train_init first)");
    SB_REQUIRE(net->layers[(size_t)layer].type != SB_LAYER_BILSTM, "sb_net_get_layer_momentum: BiLSTM layers are not trainable");
    return get_layer_tensor(net, layer, net->train->mom_w[(size_t)layer], net->train->mom_b[(size_t)layer], w_host, b_host);
}

// Restores the momentum buffers of one layer from torch-layout host arrays (checkpoint resume, train_model.py:296-326).
extern "C" int sb_net_set_layer_momentum(sb_net* net, int layer, const float* w_host, const float* b_host) {
    SB_REQUIRE(net && net->train && layer >= 0 && layer < (int)net->layers.size() && w_host && b_host,
               "sb_net_set_layer_momentum: bad arguments (sb_net_train_init first)");
    const NetLayer& N = net->layers[(size_t)layer];
    SB_REQUIRE(N.type != SB_LAYER_BILSTM, "sb_net_set_layer_momentum: BiLSTM layers are not trainable");
    SB_CHECK_CUDA(cudaSetDevice(net->device));
    SB_CHECK_CUDA(cudaDeviceSynchronize());
    const size_t nw = (size_t)N.k_f32 * N.ldw_f32;
    std::vector<float> w(nw, 0.f), b((size_t)N.ldw_f32, 0.f);
    layer_layout_copy(N, w, const_cast<float*>(w_host), false);
    for (int o = 0; o < N.cout; ++o) b[(size_t)o] = b_host[o];
    SB_CHECK_CUDA(cudaMemcpy(net->train->mom_w[(size_t)layer], w.data(), nw * 4, cudaMemcpyHostToDevice));
    SB_CHECK_CUDA(cudaMemcpy(net->train->mom_b[(size_t)layer], b.data(), (size_t)N.ldw_f32 * 4, cudaMemcpyHostToDevice));
    return SB_OK;
}
