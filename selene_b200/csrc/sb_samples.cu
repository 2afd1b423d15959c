// selene_b200 — packed training samples: the bit-packed HDF5 sample format of
// selene_sdk/samplers/dataloader.py:129-147 (unpackbits_sequence / unpackbits_targets) and
// scripts/sampler_write_to_h5.py:63-89 (np.packbits(sequences > 0, axis=1), np.packbits(targets > 0, axis=1)).
//
// Layout: sequences uint8 (n, ceil(L/8), 4) — byte (s, g, c) holds positions 8g ... 8g+7 of channel c, MSB first, so
// the four channel bytes of a group of 8 positions are one aligned 32-bit word; an unknown base ('N' = 0.25 in all four
// channels) was packed as four set bits and unpacks to 1/4 in every channel.  targets uint8 (n, ceil(F/8)), MSB first.
//
// Both directions are HBM-bound byte movers (0.5 B read -> 16 B written per base for fp32): a thread owns ONE base,
// so a warp reads four words (L1 broadcast within each group of 8 lanes) and writes 512 contiguous bytes with one
// 16-byte store per lane; the packing direction reduces the eight bases of a group with a ballot.
#include "sb_common.cuh"

namespace {

template <bool BF16>
__global__ void __launch_bounds__(256) unpack_sequences_kernel(const uint32_t* __restrict__ packed, long long n,
                                                               int groups, int seq_start, int out_len,
                                                               void* __restrict__ out) {
    const long long total = n * out_len;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long s = i / out_len;
        const int p = (int)(i - s * out_len) + seq_start;                 // position in the stored sequence
        const uint32_t w = __ldg(packed + s * groups + (p >> 3));        // bytes = channels 0..3 of the group
        const uint32_t bits = (w >> (7 - (p & 7))) & 0x01010101u;        // bit 0 of byte c = channel c at p
        const bool null_row = bits == 0x01010101u;                        // dataloader.py:131-133
        if (BF16) {
            const uint32_t one = 0x3F80u, quarter = 0x3E80u;
            uint32_t lo, hi;
            if (null_row) { lo = quarter | (quarter << 16); hi = lo; }
            else {
                lo = ((bits & 1u) ? one : 0u) | (((bits >> 8) & 1u) ? one << 16 : 0u);
                hi = (((bits >> 16) & 1u) ? one : 0u) | (((bits >> 24) & 1u) ? one << 16 : 0u);
            }
            reinterpret_cast<uint2*>(out)[i] = make_uint2(lo, hi);
        } else {
            float4 v;
            if (null_row) v = make_float4(0.25f, 0.25f, 0.25f, 0.25f);
            else v = make_float4((float)(bits & 1u), (float)((bits >> 8) & 1u), (float)((bits >> 16) & 1u),
                                 (float)((bits >> 24) & 1u));
            reinterpret_cast<float4*>(out)[i] = v;
        }
    }
}

template <typename OutT>
__global__ void __launch_bounds__(256) unpack_targets_kernel(const uint8_t* __restrict__ packed, long long n, int bytes_per_row,
                                                             int F, OutT* __restrict__ out) {
    const long long total = n * F;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long s = i / F;
        const int f = (int)(i - s * F);
        out[i] = (OutT)((__ldg(packed + s * bytes_per_row + (f >> 3)) >> (7 - (f & 7))) & 1);
    }
}

// one-hot / label rows -> packed bits (value > 0 sets the bit, positions past the end pad with zeros)
__global__ void __launch_bounds__(256) pack_sequences_kernel(const float4* __restrict__ x, long long n, int L, int groups,
                                                             uint32_t* __restrict__ packed) {
    // a warp handles 32 consecutive padded positions of one sample = 4 groups; lanes 8j..8j+7 form group j
    const long long padded = (long long)groups * 8;
    const long long total = n * padded;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    for (long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) & ~31ll; i0 < total; i0 += stride) {
        const long long i = i0 + lane;
        const long long s = i / padded;
        const int p = (int)(i - s * padded);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < total && p < L) v = __ldg(x + s * L + p);
        uint32_t word = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float val = c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w));
            const uint32_t b = __ballot_sync(0xffffffffu, val > 0.f);     // lane l = position l of the 32
            // this lane's group: bits 8j..8j+7, MSB first -> reverse the byte
            const uint32_t byte = (__brev(b >> (lane & 24)) >> 24) & 0xFFu;
            word |= byte << (8 * c);
        }
        if ((lane & 7) == 0 && i < total) packed[s * groups + (p >> 3)] = word;
    }
}

__global__ void __launch_bounds__(256) pack_targets_kernel(const float* __restrict__ y, long long n, int F, int bytes_per_row,
                                                           uint8_t* __restrict__ packed) {
    const long long total = n * bytes_per_row;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long s = i / bytes_per_row;
        const int g = (int)(i - s * bytes_per_row);
        uint32_t byte = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int f = 8 * g + b;
            if (f < F && __ldg(y + s * F + f) > 0.f) byte |= 0x80u >> b;
        }
        packed[i] = (uint8_t)byte;
    }
}

int flat_grid(long long total) {
    int dev = 0;
    cudaGetDevice(&dev);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sb_sm_count(dev) * 8;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace

extern "C" int sb_unpack_sequences(const uint8_t* packed, int64_t n, int L_stored, int seq_start, int out_len,
                                   int out_dtype, void* out, void* stream) {
    SB_REQUIRE(n >= 0 && L_stored > 0 && seq_start >= 0 && out_len >= 0 && seq_start + out_len <= ((L_stored + 7) / 8) * 8,
               "sb_unpack_sequences: window [%d, %d) outside the stored %d positions", seq_start, seq_start + out_len, L_stored);
    SB_REQUIRE(out_dtype == SB_F32 || out_dtype == SB_BF16, "sb_unpack_sequences: bad out_dtype %d", out_dtype);
    if (n == 0 || out_len == 0) return SB_OK;
    SB_REQUIRE(packed && out, "sb_unpack_sequences: null argument");
    SB_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 3) == 0, "sb_unpack_sequences: packed sequences must be 4-byte aligned");
    const int groups = (L_stored + 7) / 8;
    const long long total = n * (long long)out_len;
    cudaStream_t s = (cudaStream_t)stream;
    if (out_dtype == SB_F32)
        unpack_sequences_kernel<false><<<flat_grid(total), 256, 0, s>>>(reinterpret_cast<const uint32_t*>(packed), n, groups,
                                                                        seq_start, out_len, out);
    else
        unpack_sequences_kernel<true><<<flat_grid(total), 256, 0, s>>>(reinterpret_cast<const uint32_t*>(packed), n, groups,
                                                                       seq_start, out_len, out);
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    return SB_OK;
}

extern "C" int sb_unpack_targets(const uint8_t* packed, int64_t n, int F, int out_dtype, void* out, void* stream) {
    SB_REQUIRE(n >= 0 && F > 0, "sb_unpack_targets: bad sizes");
    SB_REQUIRE(out_dtype == SB_F32 || out_dtype == SB_U8, "sb_unpack_targets: bad out_dtype %d", out_dtype);
    if (n == 0) return SB_OK;
    SB_REQUIRE(packed && out, "sb_unpack_targets: null argument");
    const long long total = n * (long long)F;
    cudaStream_t s = (cudaStream_t)stream;
    if (out_dtype == SB_F32)
        unpack_targets_kernel<float><<<flat_grid(total), 256, 0, s>>>(packed, n, (F + 7) / 8, F, (float*)out);
    else
        unpack_targets_kernel<uint8_t><<<flat_grid(total), 256, 0, s>>>(packed, n, (F + 7) / 8, F, (uint8_t*)out);
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    return SB_OK;
}

extern "C" int sb_pack_samples(const float* sequences, const float* targets, int64_t n, int L, int F,
                               uint8_t* packed_sequences, uint8_t* packed_targets, void* stream) {
    SB_REQUIRE(n >= 0 && L > 0 && F > 0, "sb_pack_samples: bad sizes");
    if (n == 0) return SB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (sequences) {
        SB_REQUIRE(packed_sequences, "sb_pack_samples: null output for the sequences");
        SB_REQUIRE((reinterpret_cast<uintptr_t>(packed_sequences) & 3) == 0 && (reinterpret_cast<uintptr_t>(sequences) & 15) == 0,
                   "sb_pack_samples: sequence buffers must be 16-byte (input) / 4-byte (output) aligned");
        const int groups = (L + 7) / 8;
        pack_sequences_kernel<<<flat_grid(n * (long long)groups * 8), 256, 0, s>>>(
            reinterpret_cast<const float4*>(sequences), n, L, groups, reinterpret_cast<uint32_t*>(packed_sequences));
        SB_COUNT_LAUNCH();
    }
    if (targets) {
        SB_REQUIRE(packed_targets, "sb_pack_samples: null output for the targets");
        pack_targets_kernel<<<flat_grid(n * (long long)((F + 7) / 8)), 256, 0, s>>>(targets, n, F, (F + 7) / 8, packed_targets);
        SB_COUNT_LAUNCH();
    }
    SB_CHECK_LAUNCH();
    return SB_OK;
}
