// reg_f64_odd.cu -- line lengths 2^a * 3 and 2^a * 5^b (prime-factor butterflies).
// fp64: a quarter warp (8 lanes x 16 B) covers one 128-byte row of the [N][8] tile.
#include "registry.cuh"

namespace cpc {

void fill_fast_f64_odd(std::map<FastKey<double>, FastEntry<double>> &m)
{
    // line lengths 2^a * 3 (radix 6 / 12 last, prime-factor butterflies): 24 points per thread need 96 data registers
    // in fp64, so these run 3 small CTAs per SM at <= 168 registers; single-rank builds only
    register_modes_gen<double, VAR_WIDE,     48,  4, 12,  1, 12,  8,  8, 2, 2, false, 1>(m);
    register_modes_gen<double, VAR_WIDE,     96,  4,  4,  6, 12,  8,  4, 2, 2, false, 1>(m);
    register_modes_gen<double, VAR_WIDE,    192,  8,  4,  6, 24,  8,  2, 3, 3, false, 1>(m);
    register_modes_gen<double, VAR_WIDE,    384,  8,  8,  6, 24,  8,  1, 3, 3, false, 1>(m);
    register_modes_gen<double, VAR_WIDE,    768,  8,  8, 12, 24,  4,  1, 3, 3, false, 1>(m);
    register_modes_gen<double, VAR_NARROW,   96,  4,  4,  6, 12,  4,  8, 2, 2, false, 1>(m);
    register_modes_gen<double, VAR_NARROW,  192,  8,  4,  6, 24,  4,  4, 3, 3, false, 1>(m);
    register_modes_gen<double, VAR_NARROW,  384,  8,  8,  6, 24,  4,  2, 3, 3, false, 1>(m);
    register_modes_gen<double, VAR_XMAP,     48,  4, 12,  1, 12, 64,  1, 2, 2, false, 1>(m);
    register_modes_gen<double, VAR_XMAP,     96,  4,  4,  6, 12, 32,  1, 2, 2, false, 1>(m);
    register_modes_gen<double, VAR_XMAP,    192,  8,  4,  6, 24, 16,  1, 3, 3, false, 1>(m);
    register_modes_gen<double, VAR_XMAP,    384,  8,  8,  6, 24,  8,  1, 3, 3, false, 1>(m);
    register_modes_gen<double, VAR_XMAP,    768,  8,  8, 12, 24,  4,  1, 3, 3, false, 1>(m);
    // line lengths 2^a * 5^b: radix 5, 10 = 2 x 5 and 20 = 4 x 5 butterflies, 10 or 20 points per thread
    //                         variant        N   R0  R1  R2   E  TX   G MINB MINBF
    register_modes_gen<double, VAR_WIDE,    100, 10, 10,  1, 10,  8,  3, 2, 2, false, 1>(m);
    register_modes_gen<double, VAR_WIDE,    160,  4,  4, 10, 20,  8,  2, 3, 3, false, 1>(m);
    register_modes_gen<double, VAR_WIDE,    200, 10, 20,  1, 20,  8,  2, 3, 3, false, 1>(m);
    register_modes_gen<double, VAR_WIDE,    250,  5,  5, 10, 10,  8,  1, 3, 3, false, 1>(m);
    register_modes_gen<double, VAR_WIDE,    320,  4,  4, 20, 20,  8,  1, 3, 3, false, 1>(m);
    register_modes_gen<double, VAR_WIDE,    400, 20, 20,  1, 20,  8,  1, 3, 3, false, 1>(m);
    register_modes_gen<double, VAR_WIDE,    500,  5, 10, 10, 10,  8,  1, 2, 2, false, 1>(m);
    register_modes_gen<double, VAR_WIDE,    800,  4, 10, 20, 20,  4,  1, 2, 3, false, 1>(m);
    register_modes_gen<double, VAR_WIDE,   1000, 10, 10, 10, 10,  4,  1, 2, 2, false, 1>(m);
    register_modes_gen<double, VAR_XMAP,    100, 10, 10,  1, 10, 24,  1, 2, 2, false, 1>(m);
    register_modes_gen<double, VAR_XMAP,    160,  4,  4, 10, 20, 16,  1, 3, 3, false, 1>(m);
    register_modes_gen<double, VAR_XMAP,    200, 10, 20,  1, 20, 16,  1, 3, 3, false, 1>(m);
    register_modes_gen<double, VAR_XMAP,    250,  5,  5, 10, 10,  8,  1, 3, 3, false, 1>(m);
    register_modes_gen<double, VAR_XMAP,    320,  4,  4, 20, 20,  8,  1, 3, 3, false, 1>(m);
    register_modes_gen<double, VAR_XMAP,    400, 20, 20,  1, 20,  8,  1, 3, 3, false, 1>(m);
    register_modes_gen<double, VAR_XMAP,    500,  5, 10, 10, 10,  8,  1, 2, 2, false, 1>(m);
    register_modes_gen<double, VAR_XMAP,    800,  4, 10, 20, 20,  4,  1, 2, 2, false, 1>(m);
    register_modes_gen<double, VAR_XMAP,   1000, 10, 10, 10, 10,  4,  1, 2, 2, false, 1>(m);
}

}  // namespace cpc
