// selene_b200 — shared device/host helpers (sm_100a only).
//
// Error handling convention of the C ABI (include/selene_b200.h): every entry point returns an
// int status and records a human-readable message retrievable with sb_last_error(); nothing
// throws across the boundary.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdarg.h>

#include "../../include/selene_b200.h"

// ---------------------------------------------------------------------------------------------
// host: error plumbing
// ---------------------------------------------------------------------------------------------
void sb_set_error(const char* fmt, ...);

#define SB_CHECK_CUDA(expr)                                                                    \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            sb_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                    \
                         cudaGetErrorString(_e));                                              \
            return SB_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

#define SB_REQUIRE(cond, ...)                                                                  \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            sb_set_error(__VA_ARGS__);                                                         \
            return SB_ERR_ARG;                                                                 \
        }                                                                                      \
    } while (0)

#define SB_CHECK_LAUNCH()                                                                      \
    do {                                                                                       \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) {                                                               \
            sb_set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__,                \
                         cudaGetErrorString(_e));                                              \
            return SB_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

static inline int sb_ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Global launch counter (bench.py reports gpu_launches from it).
extern unsigned long long g_sb_launches;
// writer threads launch format kernels concurrently with the caller's thread: the count is atomic
#define SB_COUNT_LAUNCH() ((void)__atomic_fetch_add(&g_sb_launches, 1ull, __ATOMIC_RELAXED))

// Cached device properties.
int sb_sm_count(int device);

// ---------------------------------------------------------------------------------------------
// device: PTX wrappers (mbarrier / TMA / tcgen05). One thread issues unless noted.
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t sb_smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool sb_elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void sb_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sb_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void sb_fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void sb_fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void sb_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb_smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void sb_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sb_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool sb_mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(sb_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (CUDA error at the next sync),
// never as a hung GPU box.
__device__ __forceinline__ void sb_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!sb_mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) {
            printf("selene_b200: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x,
                   (int)threadIdx.x);
            __trap();
        }
    }
}

__device__ __forceinline__ void sb_prefetch_tmap(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

// 2-D tiled TMA load: global (tensor map) -> shared, completion signalled on `bar`.
__device__ __forceinline__ void sb_tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar,
                                               int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(sb_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(sb_smem_u32(bar)),
          "r"(c0), "r"(c1)
        : "memory");
}

// 3-D im2col TMA load (channels, position, sequence): `pixelsPerColumn` window starts from (w, n) on, each read at
// position + w_off; completion signalled on `bar`.
__device__ __forceinline__ void sb_tma_load_im2col_3d(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c,
                                                      int32_t w, int32_t n, uint16_t w_off) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2], {%6};"
        :
        : "r"(sb_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(sb_smem_u32(bar)), "r"(c), "r"(w), "r"(n),
          "h"(w_off)
        : "memory");
}

// ---- CTA pairs (cta_group::2): the variants CUTLASS uses in cute/arch/copy_sm100_tma.hpp, mma_sm100_umma.hpp,
// tmem_allocator_sm100.hpp and cutlass/arch/barrier.h.  A shared::cta address of CTA 1 with bit 24 cleared names the same
// location in CTA 0 (the leader), which owns the "full" barriers of the pair.
constexpr uint32_t kSbPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void sb_tma_load_2d_pair(void* smem_dst, const void* tmap, uint64_t* leader_bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(sb_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(sb_smem_u32(leader_bar) & kSbPeerBitMask),
          "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void sb_tma_load_im2col_3d_pair(void* smem_dst, const void* tmap, uint64_t* leader_bar, int32_t c,
                                                           int32_t w, int32_t n, uint16_t w_off) {
    asm volatile(
        "cp.async.bulk.tensor.3d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2], {%6};"
        :
        : "r"(sb_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(sb_smem_u32(leader_bar) & kSbPeerBitMask),
          "r"(c), "r"(w), "r"(n), "h"(w_off)
        : "memory");
}
__device__ __forceinline__ void sb_tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {  // one warp of EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sb_smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void sb_tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs, 128 rows each] (+)= [A of CTA 0; A of CTA 1] * [B half of CTA 0 | B half of CTA 1]; leader thread only
__device__ __forceinline__ void sb_umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                  uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// arrive on `bar` (same offset) in BOTH CTAs once all previously issued MMAs of the pair have completed
__device__ __forceinline__ void sb_umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     sb_smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

// --- tcgen05 / TMEM -------------------------------------------------------------------------
__device__ __forceinline__ void sb_tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     sb_smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void sb_tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void sb_tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void sb_tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 inputs, fp32 accumulate.
__device__ __forceinline__ void sb_umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on `bar` once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void sb_umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     sb_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void sb_tmem_wait_ld() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 32-bit, 16 consecutive columns -> 16 registers per thread (thread i = lane base+i).
__device__ __forceinline__ void sb_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void sb_tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void sb_tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :
                 : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void sb_tmem_wait_st() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// 1-D bulk copy global -> shared (size multiple of 16 bytes, both addresses 16-byte aligned), completion
// counted in bytes on `bar`
__device__ __forceinline__ void sb_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :
                 : "r"(sb_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(sb_smem_u32(bar))
                 : "memory");
}
// ---- thread-block clusters: distributed shared memory
__device__ __forceinline__ uint32_t sb_cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void sb_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address of this CTA's layout) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t sb_mapa(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void sb_st_cluster_v4(uint32_t cluster_addr, uint4 v) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void sb_mbar_arrive_cluster(uint32_t cluster_bar_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
// wait that synchronises with arrivals from other CTAs of the cluster
__device__ __forceinline__ void sb_mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0, ok = 0;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred P;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P;\n\t"
            "}"
            : "=r"(ok)
            : "r"(sb_smem_u32(bar)), "r"(parity)
            : "memory");
        if (!ok && ++spins > (1u << 26)) {
            printf("selene_b200: cluster mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    } while (!ok);
}

// per-thread 16-byte asynchronous copy global -> shared (Ampere-style cp.async; completion by commit/wait groups)
__device__ __forceinline__ void sb_cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sb_smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(gsrc))
                 : "memory");
}
__device__ __forceinline__ void sb_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void sb_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void sb_prefetch_l2(const void* g) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<uint64_t>(g)));
}

// Shared-memory matrix descriptor for a K-major bf16 tile stored as rows of 128 bytes (64 bf16)
// in the 128-byte-swizzle pattern TMA writes (8-row / 1024-byte atoms): SBO = 1024 B, LBO unused,
// descriptor version 1 (sm_100), layout type 2 (SWIZZLE_128B). The tile base must be 1024-byte
// aligned; advancing along K by 16 bf16 adds 32 bytes to the start address.
__device__ __forceinline__ uint64_t sb_umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, dense, M x N tile.
__host__ __device__ constexpr uint32_t sb_umma_idesc_bf16(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ float sb_ld_nc_f32(const float* p) { return __ldg(p); }

#endif  // __CUDACC__
