// registry.cuh -- the table of compiled fast kernels: (axis length, variant, mode) -> launch entry.
// The instantiations are spread over several translation units (reg_*.cu) so that they build in parallel;
// plan_impl.cuh only sees the fill functions declared at the end of this header.
#pragma once
#include <map>
#include <tuple>

#include "fft_pass.cuh"
#include "fft_r2x.cuh"

namespace cpc {

// ------------------------------------------------------------------------------------------------
// Fast-kernel registry
// ------------------------------------------------------------------------------------------------
template <typename T> struct FastEntry {
    void (*kern)(const cplx_t<T> *, cplx_t<T> *, const PassGeom, const cplx_t<T> *, const SymbolArgs<T>);
    int threads;
    size_t smem;
    int g;          // tiles per CTA
    int tx;         // lines per tile
    int radix[3];
};

// Kernel variants of one axis length.
enum Variant {
    VAR_WIDE = 0,     // strided lines, TX lanes = one 128-byte row (y, z passes)
    VAR_NARROW = 1,   // strided lines, TX = 4 (wave x pass: the 4 components of a cell; long lines)
    VAR_XMAP = 2,     // contiguous lines (scalar x pass)
    VAR_WIDE2 = 3,    // as VAR_WIDE with a different points-per-thread / radix split (512: two butterflies per thread)
    VAR_SMALL = 4,    // as VAR_WIDE with small CTAs, 4 per SM (256-point y lines)
    VAR_R2X = 5,      // 512 = 2 x (16 x 16): radix-2 level in registers + warp shuffle, one shared-memory exchange
    VAR_XR2X = 6,     // the same for contiguous lines: one warp per line, no block barrier (scalar x pass)
    VAR_SLIM = 7,     // strided lines, TX = 4 and one line group per CTA: 64 KB tiles for 1024-point lines, 2 CTAs per SM
    VAR_COUNT = 8
};

template <typename T> using FastKey = std::tuple<int, int, int>;   // (n, variant, mode + 16 * general-addressing)

constexpr int GEN_BIT = 16;   // key offset of the kernels compiled with chunked-layout / peer-push addressing

template <typename T, int VAR, int N, int R0, int R1, int R2, int E, int TX, int G, int MINB, int MINBF, bool GEN, int NG>
static void register_modes_gen(std::map<FastKey<T>, FastEntry<T>> &m)
{
    constexpr int NST = (R1 > 1) + (R2 > 1) + 1;
    constexpr bool XM = (VAR == VAR_XMAP);
    constexpr int threads = (N / E) * TX * G;
    constexpr size_t smem = NST > 1 ? (size_t)G * NG * SmemTile<N, TX / NG, Log2<R0>::v, XM>::elems * sizeof(cplx_t<T>) : 0;
    constexpr int KB = GEN ? GEN_BIT : 0;
    FastEntry<T> e{ nullptr, threads, smem, G, TX, { R0, R1, R2 } };
    e.kern = fft_pass_kernel<T, N, R0, R1, R2, E, TX, G, MODE_FWD, MINB, XM, GEN, NG>;
    m[FastKey<T>(N, VAR, MODE_FWD + KB)] = e;
    e.kern = fft_pass_kernel<T, N, R0, R1, R2, E, TX, G, MODE_INV, MINB, XM, GEN, NG>;
    m[FastKey<T>(N, VAR, MODE_INV + KB)] = e;
    if constexpr (XM && NST > 1 && !GEN) {
        e.kern = fft_pass_kernel<T, N, R0, R1, R2, E, TX, G, MODE_R2C, MINB, XM, false, NG>;
        m[FastKey<T>(N, VAR, MODE_R2C)] = e;
        e.kern = fft_pass_kernel<T, N, R0, R1, R2, E, TX, G, MODE_C2R, MINB, XM, false, NG>;
        m[FastKey<T>(N, VAR, MODE_C2R)] = e;
    }
    if constexpr (!XM) {
        e.kern = fft_pass_kernel<T, N, R0, R1, R2, E, TX, G, MODE_FUSED_SEP, MINBF, XM, GEN, NG>;
        m[FastKey<T>(N, VAR, MODE_FUSED_SEP + KB)] = e;
        e.kern = fft_pass_kernel<T, N, R0, R1, R2, E, TX, G, MODE_FUSED_TABLE, MINBF, XM, GEN, NG>;
        m[FastKey<T>(N, VAR, MODE_FUSED_TABLE + KB)] = e;
        e.kern = fft_pass_kernel<T, N, R0, R1, R2, E, TX, G, MODE_FUSED_WAVE, MINBF, XM, GEN, NG>;
        m[FastKey<T>(N, VAR, MODE_FUSED_WAVE + KB)] = e;
    }
}

// Every variant is compiled with plain strided addressing; the variants used for the y and z passes of power-of-two
// grids (VAR_WIDE and the tuned 512 / 256 ones) also get the general-addressing build needed by multi-rank plans.
template <typename T, int VAR, int N, int R0, int R1, int R2, int E, int TX, int G, int MINB, int MINBF = MINB, int NG = 1>
static void register_modes(std::map<FastKey<T>, FastEntry<T>> &m)
{
    register_modes_gen<T, VAR, N, R0, R1, R2, E, TX, G, MINB, MINBF, false, NG>(m);
    if constexpr (VAR == VAR_WIDE || VAR == VAR_WIDE2 || VAR == VAR_SMALL)
        register_modes_gen<T, VAR, N, R0, R1, R2, E, TX, G, MINB, MINBF, true, NG>(m);
}

// strided lines of 2 H points as 2 x (R0 x R1) (fft_r2x.cuh)
template <typename T, int H, int R0, int R1> static void register_r2x(std::map<FastKey<T>, FastEntry<T>> &m)
{
    constexpr int TX = 128 / (int)sizeof(cplx_t<T>);      // 8 lanes (complex128) or 16 (complex64) = one 128-byte row
    FastEntry<T> e{ nullptr, H * TX / 8, (size_t)H * 2 * TX * sizeof(cplx_t<T>), 1, TX, { R0, R1, 1 } };
#define CPC_R2X(MODE)                                                                            \
    e.kern = fft_r2x_kernel<T, H, R0, R1, MODE, false>; m[FastKey<T>(2 * H, VAR_R2X, MODE)] = e;           \
    e.kern = fft_r2x_kernel<T, H, R0, R1, MODE, true>;  m[FastKey<T>(2 * H, VAR_R2X, MODE + GEN_BIT)] = e;
    CPC_R2X(MODE_FWD)
    CPC_R2X(MODE_INV)
    CPC_R2X(MODE_FUSED_SEP)
    CPC_R2X(MODE_FUSED_TABLE)
    CPC_R2X(MODE_FUSED_WAVE)
#undef CPC_R2X
}

// contiguous 512-point lines, one warp per line (fft_r2x.cuh); plain transforms only
template <typename T> static void register_r2x512_line(std::map<FastKey<T>, FastEntry<T>> &m)
{
    constexpr int LINES = 8;
    FastEntry<T> e{ nullptr, 32 * LINES, (size_t)LINES * 512 * sizeof(cplx_t<T>), 1, LINES, { 16, 16, 1 } };
    e.kern = fft_r2x512_line_kernel<T, MODE_FWD, LINES>; m[FastKey<T>(512, VAR_XR2X, MODE_FWD)] = e;
    e.kern = fft_r2x512_line_kernel<T, MODE_INV, LINES>; m[FastKey<T>(512, VAR_XR2X, MODE_INV)] = e;
}

// contiguous 256-point lines, half a warp per line (fft_r2x.cuh): plain transforms and the r2c / c2r pair
template <typename T> static void register_line256(std::map<FastKey<T>, FastEntry<T>> &m)
{
    constexpr int LINES = 16;
    FastEntry<T> e{ nullptr, 16 * LINES, (size_t)LINES * 256 * sizeof(cplx_t<T>), 1, LINES, { 16, 16, 1 } };
    e.kern = fft_line256_kernel<T, MODE_FWD, LINES>; m[FastKey<T>(256, VAR_XR2X, MODE_FWD)] = e;
    e.kern = fft_line256_kernel<T, MODE_INV, LINES>; m[FastKey<T>(256, VAR_XR2X, MODE_INV)] = e;
    e.kern = fft_line256_kernel<T, MODE_R2C, LINES>; m[FastKey<T>(256, VAR_XR2X, MODE_R2C)] = e;
    e.kern = fft_line256_kernel<T, MODE_C2R, LINES>; m[FastKey<T>(256, VAR_XR2X, MODE_C2R)] = e;
}

// one function per translation unit (reg_<dtype>_<group>.cu)
void fill_fast_f32_odd(std::map<FastKey<float>, FastEntry<float>> &m);
void fill_fast_f32_pow2a(std::map<FastKey<float>, FastEntry<float>> &m);
void fill_fast_f32_pow2b(std::map<FastKey<float>, FastEntry<float>> &m);
void fill_fast_f32_r2x(std::map<FastKey<float>, FastEntry<float>> &m);
void fill_fast_f64_odd(std::map<FastKey<double>, FastEntry<double>> &m);
void fill_fast_f64_pow2a(std::map<FastKey<double>, FastEntry<double>> &m);
void fill_fast_f64_pow2b(std::map<FastKey<double>, FastEntry<double>> &m);
void fill_fast_f64_r2x(std::map<FastKey<double>, FastEntry<double>> &m);

}  // namespace cpc
