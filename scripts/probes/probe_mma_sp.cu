// Empirical layout probe for mma.sp::ordered_metadata.sync.aligned.m16n8k32.row.col.f32.bf16.bf16.f32 on sm_100a:
// which thread / nibble of the metadata operand controls which (row, group of 4 logical K), and which (thread, register,
// half) of the compressed A fragment is which (row, group, slot).  nvcc -arch=sm_100a probe_mma_sp.cu -o probe && ./probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ void mma_sp(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[4], uint32_t e, int sel) {
    if (sel == 0)
        asm volatile("mma.sp::ordered_metadata.sync.aligned.m16n8k32.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9,%10,%11}, {%12,%12,%12,%12}, %13, 0x0;"
                     : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "f"(0.f), "r"(e));
    else
        asm volatile("mma.sp::ordered_metadata.sync.aligned.m16n8k32.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9,%10,%11}, {%12,%12,%12,%12}, %13, 0x1;"
                     : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "f"(0.f), "r"(e));
}

// assumed B fragment: register i, half h of thread (g = lane >> 2, t = lane & 3) = B[k = 2t + h + 8i][n = g]
__device__ void make_b(uint32_t (&b)[4], int lane, int mode, int k0, int n0) {
    const int g = lane >> 2, t = lane & 3;
    for (int i = 0; i < 4; ++i) {
        uint32_t v = 0;
        for (int h = 0; h < 2; ++h) {
            const int k = 2 * t + h + 8 * i, n = g;
            bool one = false;
            if (mode == 0) one = k == k0 && n == n0;                 // single element
            if (mode == 1) one = (k / 4 == n) && (k % 4 >= 2);       // group n, bases 2 / 3
            if (mode == 2) one = (k / 4 == n);                       // group n
            if (mode == 3) one = (n < 4) && (k % 4 == n);            // base n
            if (one) v |= 0x3F80u << (16 * h);
        }
        b[i] = v;
    }
}

__global__ void probe(int* out) {
    const int lane = threadIdx.x;
    const int g = lane >> 2, t = lane & 3;
    float d[4];
    uint32_t a[4], b[4];
    int* o = out;
    // ---- 1. B layout check: A all ones, metadata (0,1) per group: D[r][n0] == 1 for every row iff B[k0][n0] is where assumed
    int bad = 0;
    for (int sel = 0; sel < 2; ++sel)
    for (int pass = 0; pass < 2; ++pass)
        for (int k0 = 0; k0 < 32; ++k0) {
            if ((k0 % 4 >= 2) != (pass == 1)) continue;
            for (int n0 = 0; n0 < 8; ++n0) {
                for (int i = 0; i < 4; ++i) a[i] = 0x3F803F80u;
                make_b(b, lane, 0, k0, n0);
                mma_sp(d, a, b, pass ? 0xEEEEEEEEu : 0x44444444u, sel);
                // d0: (g, 2t) d1: (g, 2t+1) d2: (g+8, 2t) d3: (g+8, 2t+1)
                for (int j = 0; j < 4; ++j) {
                    const int n = 2 * t + (j & 1);
                    const float want = n == n0 ? 1.f : 0.f;
                    if (d[j] != want) bad = 1;
                }
            }
        }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) o[0] = bad;
    o += 1;
    // ---- 2. metadata: flip nibble q of thread T to (2,3); B mode 1 -> D[r][n] = 2 at the flipped (row, group)
    for (int sel = 0; sel < 2; ++sel)
        for (int T = 0; T < 32; ++T)
            for (int q = 0; q < 8; ++q) {
                for (int i = 0; i < 4; ++i) a[i] = 0x3F803F80u;
                make_b(b, lane, 1, 0, 0);
                uint32_t e = 0x44444444u;
                if (lane == T) e = (e & ~(0xFu << (4 * q))) | (0xEu << (4 * q));
                mma_sp(d, a, b, e, sel);
                int found = -1;
                for (int j = 0; j < 4; ++j)
                    if (d[j] != 0.f) found = ((g + 8 * (j >> 1)) << 8) | ((2 * t + (j & 1)) << 4) | (int)d[j];
                // gather: at most a few lanes see it
                for (int l = 0; l < 32; ++l) {
                    const int f = __shfl_sync(0xffffffffu, found, l);
                    if (f >= 0 && lane == 0) o[(sel * 32 + T) * 8 + q] = f + 1;
                }
            }
    o += 2 * 32 * 8;
    // ---- 3. A fragment: one compressed element = 1 (thread T, register i, half h); metadata (0,1);
    //         B mode 2 -> group, B mode 3 -> base (slot)
    for (int T = 0; T < 32; ++T)
        for (int i = 0; i < 4; ++i)
            for (int h = 0; h < 2; ++h) {
                int res = 0;
                for (int mode = 2; mode <= 3; ++mode) {
                    for (int ii = 0; ii < 4; ++ii) a[ii] = 0;
                    if (lane == T) a[i] = 0x3F80u << (16 * h);
                    make_b(b, lane, mode, 0, 0);
                    mma_sp(d, a, b, 0x44444444u, 0);
                    int found = -1;
                    for (int j = 0; j < 4; ++j)
                        if (d[j] != 0.f) found = ((g + 8 * (j >> 1)) << 4) | (2 * t + (j & 1));
                    for (int l = 0; l < 32; ++l) {
                        const int f = __shfl_sync(0xffffffffu, found, l);
                        if (f >= 0) res |= (f + 1) << (mode == 2 ? 0 : 12);
                    }
                }
                if (lane == 0) o[(T * 4 + i) * 2 + h] = res;
            }
}

int main() {
    int n = 1 + 2 * 32 * 8 + 32 * 4 * 2;
    int* d;
    cudaMalloc(&d, n * sizeof(int));
    cudaMemset(d, 0, n * sizeof(int));
    probe<<<1, 32>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    int* h = new int[n];
    cudaMemcpy(h, d, n * sizeof(int), cudaMemcpyDeviceToHost);
    printf("B layout assumption %s\n", h[0] ? "WRONG" : "ok");
    int* m = h + 1;
    for (int sel = 0; sel < 2; ++sel) {
        printf("metadata, selector %d: thread: nibble -> (row, group, value)\n", sel);
        for (int T = 0; T < 32; ++T) {
            bool any = false;
            for (int q = 0; q < 8; ++q) any |= m[(sel * 32 + T) * 8 + q] != 0;
            if (!any) continue;
            printf("  T%-2d:", T);
            for (int q = 0; q < 8; ++q) {
                int f = m[(sel * 32 + T) * 8 + q] - 1;
                if (f < 0) printf("  q%d:-", q);
                else printf("  q%d:(r%d,g%d,%d)", q, f >> 8, (f >> 4) & 15, f & 15);
            }
            printf("\n");
        }
    }
    int* af = m + 2 * 32 * 8;
    printf("A fragment: thread reg.half -> (row, group; base)\n");
    for (int T = 0; T < 32; ++T) {
        printf("  T%-2d:", T);
        for (int i = 0; i < 4; ++i)
            for (int hh = 0; hh < 2; ++hh) {
                int r = af[(T * 4 + i) * 2 + hh];
                int a = (r & 0xFFF) - 1, b = (r >> 12) - 1;
                printf(" a%d.%d:(r%d,g%d;b%d)", i, hh, a >> 4, a & 15, b < 0 ? -1 : (b & 15));
            }
        printf("\n");
    }
    return 0;
}
