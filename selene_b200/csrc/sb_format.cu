// selene_b200 — SURVEY §8 (f1): the score writers' number formatting on the device.
//
// Reference: predict_handlers/handler.py:15-42 writes every value as "{:.2e}".format(v) (v a numpy float32, i.e.
// the correctly rounded, round-half-even 3-significant-digit form of the EXACT value) joined by tabs, one row per
// mutant / variant; that loop is half of the reference's ISM wall time.  Here a warp formats one
// row: pass 1 sizes the rows (a finite value takes 8 bytes, 9 when negative; "nan" / "inf" / "-inf"), the host
// turns the sizes into offsets, pass 2 writes the text.  Rounding is exact: the value is scaled by a power of ten
// held as a double-double (106 bits, far more than a 24-bit mantissa can need), and the remainder decides the
// rounding, ties to even.
#include "sb_common.cuh"

namespace {

// 10^k as hi + lo doubles, k = -38 .. 48 (index k + 38)
__constant__ double kP10[87][2] = {
    {0x1.b38fb9daa78e4p-127, 0x1.2acb73de9ac65p-181},
    {0x1.1039d428a8b8fp-123, -0x1.4540d794df441p-177},
    {0x1.54484932d2e72p-120, 0x1.696ef285e8eafp-174},
    {0x1.a95a5b7f87a0fp-117, -0x1.e1aa86c4e6d2fp-174},
    {0x1.09d8792fb4c49p-113, 0x1.5a5ead789df78p-167},
    {0x1.4c4e977ba1f5cp-110, -0x1.4f09a7293a8aap-164},
    {0x1.9f623d5a8a733p-107, -0x1.a2cc10f3892d4p-161},
    {0x1.039d665896880p-103, -0x1.85bf8a9835bc4p-157},
    {0x1.4484bfeebc2a0p-100, -0x1.e72f6d3e432b6p-154},
    {0x1.95a5efea6b347p-97, 0x1.9f04b7722c09dp-151},
    {0x1.fb0f6be506019p-94, 0x1.06c5e54eb70c4p-148},
    {0x1.3ce9a36f23c10p-90, -0x1.b788a15d9b30bp-145},
    {0x1.8c240c4aecb14p-87, -0x1.12b564da80fe7p-141},
    {0x1.ef2d0f5da7dd9p-84, -0x1.5762be11213e0p-138},
    {0x1.357c299a88ea7p-80, 0x1.a96249354b394p-134},
    {0x1.82db34012b251p-77, 0x1.13badb829e079p-131},
    {0x1.e392010175ee6p-74, -0x1.a7566d9cba769p-128},
    {0x1.2e3b40a0e9b4fp-70, 0x1.f769fb7e0b75ep-124},
    {0x1.79ca10c924223p-67, 0x1.75447a5d8e536p-121},
    {0x1.d83c94fb6d2acp-64, 0x1.a52b31e9e3d07p-119},
    {0x1.2725dd1d243acp-60, -0x1.7c628066e8ceep-114},
    {0x1.70ef54646d497p-57, -0x1.db7b2080a3029p-111},
    {0x1.cd2b297d889bcp-54, 0x1.5b4c2ebe68799p-109},
    {0x1.203af9ee75616p-50, -0x1.937831647f5a0p-104},
    {0x1.6849b86a12b9bp-47, 0x1.ea70909833de7p-107},
    {0x1.c25c268497682p-44, -0x1.ecd79a5a0df95p-99},
    {0x1.19799812dea11p-40, 0x1.97f27f0f6e886p-96},
    {0x1.5fd7fe1796495p-37, 0x1.7f7bc7b4d28aap-91},
    {0x1.b7cdfd9d7bdbbp-34, -0x1.20a5465df8d2cp-88},
    {0x1.12e0be826d695p-30, -0x1.34674bfabb83bp-84},
    {0x1.5798ee2308c3ap-27, -0x1.03023df2d4c94p-82},
    {0x1.ad7f29abcaf48p-24, 0x1.5e1e99483b023p-78},
    {0x1.0c6f7a0b5ed8dp-20, 0x1.b5a63f9a49c2cp-75},
    {0x1.4f8b588e368f1p-17, -0x1.ee78183f91e64p-71},
    {0x1.a36e2eb1c432dp-14, -0x1.6a161e4f765fep-68},
    {0x1.0624dd2f1a9fcp-10, -0x1.89374bc6a7efap-66},
    {0x1.47ae147ae147bp-7, -0x1.eb851eb851eb8p-63},
    {0x1.999999999999ap-4, -0x1.999999999999ap-58},
    {0x1.0000000000000p+0, 0x0.0p+0},
    {0x1.4000000000000p+3, 0x0.0p+0},
    {0x1.9000000000000p+6, 0x0.0p+0},
    {0x1.f400000000000p+9, 0x0.0p+0},
    {0x1.3880000000000p+13, 0x0.0p+0},
    {0x1.86a0000000000p+16, 0x0.0p+0},
    {0x1.e848000000000p+19, 0x0.0p+0},
    {0x1.312d000000000p+23, 0x0.0p+0},
    {0x1.7d78400000000p+26, 0x0.0p+0},
    {0x1.dcd6500000000p+29, 0x0.0p+0},
    {0x1.2a05f20000000p+33, 0x0.0p+0},
    {0x1.74876e8000000p+36, 0x0.0p+0},
    {0x1.d1a94a2000000p+39, 0x0.0p+0},
    {0x1.2309ce5400000p+43, 0x0.0p+0},
    {0x1.6bcc41e900000p+46, 0x0.0p+0},
    {0x1.c6bf526340000p+49, 0x0.0p+0},
    {0x1.1c37937e08000p+53, 0x0.0p+0},
    {0x1.6345785d8a000p+56, 0x0.0p+0},
    {0x1.bc16d674ec800p+59, 0x0.0p+0},
    {0x1.158e460913d00p+63, 0x0.0p+0},
    {0x1.5af1d78b58c40p+66, 0x0.0p+0},
    {0x1.b1ae4d6e2ef50p+69, 0x0.0p+0},
    {0x1.0f0cf064dd592p+73, 0x0.0p+0},
    {0x1.52d02c7e14af6p+76, 0x1.0000000000000p+23},
    {0x1.a784379d99db4p+79, 0x1.0000000000000p+24},
    {0x1.08b2a2c280291p+83, -0x1.b000000000000p+29},
    {0x1.4adf4b7320335p+86, -0x1.1c00000000000p+32},
    {0x1.9d971e4fe8402p+89, -0x1.8c00000000000p+33},
    {0x1.027e72f1f1281p+93, 0x1.8440000000000p+38},
    {0x1.431e0fae6d721p+96, 0x1.f2a8000000000p+42},
    {0x1.93e5939a08ceap+99, -0x1.215c000000000p+44},
    {0x1.f8def8808b024p+102, 0x1.4b26800000000p+48},
    {0x1.3b8b5b5056e17p+106, -0x1.3107f00000000p+52},
    {0x1.8a6e32246c99cp+109, 0x1.82b6140000000p+55},
    {0x1.ed09bead87c03p+112, 0x1.e363990000000p+58},
    {0x1.3426172c74d82p+116, 0x1.5c3c7f4000000p+61},
    {0x1.812f9cf7920e3p+119, -0x1.265a307800000p+65},
    {0x1.e17b84357691bp+122, 0x1.900f436a00000p+68},
    {0x1.2ced32a16a1b1p+126, 0x1.e826288900000p+70},
    {0x1.78287f49c4a1dp+129, 0x1.988becaad0000p+75},
    {0x1.d6329f1c35ca5p+132, -0x1.0151182a7c000p+78},
    {0x1.25dfa371a19e7p+136, -0x1.069578d46c000p+79},
    {0x1.6f578c4e0a061p+139, -0x1.29075ae130e00p+85},
    {0x1.cb2d6f618c879p+142, -0x1.cd24c665f4600p+86},
    {0x1.1efc659cf7d4cp+146, -0x1.c80dbeffee2f0p+92},
    {0x1.66bb7f0435c9ep+149, 0x1.c5eed14016454p+95},
    {0x1.c06a5ec5433c6p+152, 0x1.bb542c80deb48p+95},
    {0x1.18427b3b4a05cp+156, -0x1.babad90bdd33dp+101},
    {0x1.5e531a0a1c873p+159, -0x1.14b4c7a76a406p+105},
};

struct Dec3 { int digits, exp10, kind; };   // kind 0: finite, 1: nan, 2: inf

__device__ __forceinline__ Dec3 dec3(float xf) {
    Dec3 r = {0, 0, 0};
    const uint32_t bits = __float_as_uint(xf) & 0x7fffffffu;
    if (bits > 0x7f800000u) { r.kind = 1; return r; }
    if (bits == 0x7f800000u) { r.kind = 2; return r; }
    if (bits == 0) return r;                       // "0.00e+00"
    const double v = fabs((double)xf);             // exact
    // floor(log10 v) from the binary exponent, corrected below
    int e2;
    frexp(v, &e2);                                  // v = f * 2^e2, f in [0.5, 1)
    int e = (int)floor((e2 - 1) * 0.30102999566398120);
    int d = 0;
    double rem = 0.0;
#pragma unroll 1
    for (int it = 0; it < 3; ++it) {
        const double ph = kP10[2 - e + 38][0], pl = kP10[2 - e + 38][1];
        const double th = v * ph;
        const double tl = fma(v, ph, -th) + v * pl;   // v * 10^(2-e) = th + tl
        if (th + tl < 100.0) { --e; continue; }
        if (th + tl >= 1000.0) { ++e; continue; }
        double fl = floor(th);
        rem = (th - fl) + tl;                          // in (-1, 2): th - fl is exact
        if (rem < 0.0) { fl -= 1.0; rem = (th - fl) + tl; }
        else if (rem >= 1.0) { fl += 1.0; rem = (th - fl) + tl; }
        d = (int)fl;
        break;
    }
    if (rem > 0.5 || (rem == 0.5 && (d & 1))) ++d;      // round half to even on the exact value
    if (d >= 1000) { d = 100; ++e; }
    r.digits = d;
    r.exp10 = e;
    return r;
}

__device__ __forceinline__ int e2_length(float xf) {
    const uint32_t b = __float_as_uint(xf);
    const uint32_t a = b & 0x7fffffffu;
    if (a > 0x7f800000u) return 3;                      // nan
    if (a == 0x7f800000u) return 3 + (b >> 31);         // inf / -inf
    return 8 + (b >> 31);                               // d.dde+XX
}

// writes the text of xf at p, returns its length
__device__ __forceinline__ int e2_write(float xf, uint8_t* p) {
    const Dec3 r = dec3(xf);
    const bool neg = __float_as_uint(xf) >> 31;
    int n = 0;
    if (r.kind == 1) { p[0] = 'n'; p[1] = 'a'; p[2] = 'n'; return 3; }
    if (neg) p[n++] = '-';
    if (r.kind == 2) { p[n] = 'i'; p[n + 1] = 'n'; p[n + 2] = 'f'; return n + 3; }
    const int d = r.digits;
    p[n++] = (uint8_t)('0' + d / 100);
    p[n++] = '.';
    p[n++] = (uint8_t)('0' + (d / 10) % 10);
    p[n++] = (uint8_t)('0' + d % 10);
    p[n++] = 'e';
    int e = r.exp10;
    p[n++] = e < 0 ? '-' : '+';
    if (e < 0) e = -e;
    p[n++] = (uint8_t)('0' + e / 10);
    p[n++] = (uint8_t)('0' + e % 10);
    return n;
}

constexpr int kFmtWarps = 8;

// bytes of every row: values + (F - 1) tabs + '\n'
__global__ void __launch_bounds__(kFmtWarps * 32) e2_row_bytes_kernel(const float* __restrict__ x, long long n_rows, int F,
                                                                     long long* __restrict__ row_bytes) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * kFmtWarps;
    for (long long r = (long long)blockIdx.x * kFmtWarps + (threadIdx.x >> 5); r < n_rows; r += warps) {
        int s = 0;
        for (int f = lane; f < F; f += 32) s += e2_length(x[r * F + f]) + 1;
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) row_bytes[r] = s;
    }
}

__global__ void __launch_bounds__(kFmtWarps * 32) e2_rows_kernel(const float* __restrict__ x, long long n_rows, int F,
                                                                const long long* __restrict__ row_off,
                                                                const uint8_t* __restrict__ prefix,
                                                                const long long* __restrict__ prefix_off,
                                                                uint8_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * kFmtWarps;
    for (long long r = (long long)blockIdx.x * kFmtWarps + (threadIdx.x >> 5); r < n_rows; r += warps) {
        long long base = row_off[r];
        if (prefix) {   // the row's id columns (bytes supplied by the caller) come first
            const long long p0 = prefix_off[r], plen = prefix_off[r + 1] - p0;
            for (long long i = lane; i < plen; i += 32) out[base + i] = prefix[p0 + i];
            base += plen;
        }
        for (int f0 = 0; f0 < F; f0 += 32) {
            const int f = f0 + lane;
            const float v = f < F ? x[r * F + f] : 0.f;
            const int len = f < F ? e2_length(v) + 1 : 0;
            int incl = len;                                  // inclusive scan of the lengths of this batch
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (f < F) {
                uint8_t* p = out + base + (incl - len);
                const int n = e2_write(v, p);
                p[n] = f == F - 1 ? '\n' : '\t';
            }
            base += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

}  // namespace

extern "C" int sb_format_e2_row_bytes(const float* x, int64_t n_rows, int F, int64_t* row_bytes, void* stream) {
    SB_REQUIRE(n_rows >= 0 && F > 0, "sb_format_e2_row_bytes: bad sizes");
    if (n_rows == 0) return SB_OK;
    SB_REQUIRE(x && row_bytes, "sb_format_e2_row_bytes: null argument");
    int dev = 0;
    SB_CHECK_CUDA(cudaGetDevice(&dev));
    long long blocks = (n_rows + kFmtWarps - 1) / kFmtWarps;
    const long long cap = (long long)sb_sm_count(dev) * 8;
    if (blocks > cap) blocks = cap;
    e2_row_bytes_kernel<<<(int)blocks, kFmtWarps * 32, 0, (cudaStream_t)stream>>>(x, n_rows, F,
                                                                                 reinterpret_cast<long long*>(row_bytes));
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    return SB_OK;
}

static int format_rows(const float* x, int64_t n_rows, int F, const int64_t* row_off, const uint8_t* prefix,
                       const int64_t* prefix_off, uint8_t* out, void* stream) {
    SB_REQUIRE(n_rows >= 0 && F > 0, "sb_format_e2_rows: bad sizes");
    if (n_rows == 0) return SB_OK;
    SB_REQUIRE(x && row_off && out, "sb_format_e2_rows: null argument");
    SB_REQUIRE((prefix == nullptr) == (prefix_off == nullptr), "sb_format_e2_rows_prefixed: prefix and prefix_off go together");
    int dev = 0;
    SB_CHECK_CUDA(cudaGetDevice(&dev));
    long long blocks = (n_rows + kFmtWarps - 1) / kFmtWarps;
    const long long cap = (long long)sb_sm_count(dev) * 8;
    if (blocks > cap) blocks = cap;
    e2_rows_kernel<<<(int)blocks, kFmtWarps * 32, 0, (cudaStream_t)stream>>>(
        x, n_rows, F, reinterpret_cast<const long long*>(row_off), prefix, reinterpret_cast<const long long*>(prefix_off), out);
    SB_COUNT_LAUNCH();
    SB_CHECK_LAUNCH();
    return SB_OK;
}

extern "C" int sb_format_e2_rows(const float* x, int64_t n_rows, int F, const int64_t* row_off, uint8_t* out,
                                 void* stream) {
    return format_rows(x, n_rows, F, row_off, nullptr, nullptr, out, stream);
}

extern "C" int sb_format_e2_rows_prefixed(const float* x, int64_t n_rows, int F, const int64_t* row_off,
                                          const uint8_t* prefix, const int64_t* prefix_off, uint8_t* out, void* stream) {
    SB_REQUIRE(prefix && prefix_off, "sb_format_e2_rows_prefixed: null prefix");
    return format_rows(x, n_rows, F, row_off, prefix, prefix_off, out, stream);
}
