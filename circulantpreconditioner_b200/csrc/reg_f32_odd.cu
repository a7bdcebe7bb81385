// reg_f32_odd.cu -- line lengths 2^a * 3 and 2^a * 5^b (prime-factor butterflies).
// fp32: 16 lanes x 8 B = one 128-byte row.
#include "registry.cuh"

namespace cpc {

void fill_fast_f32_odd(std::map<FastKey<float>, FastEntry<float>> &m)
{
    register_modes_gen<float, VAR_WIDE,     48,  4, 12,  1, 12, 16,  4, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_WIDE,     96,  4,  4,  6, 12, 16,  2, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_WIDE,    192,  8,  4,  6, 24, 16,  2, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_WIDE,    384,  8,  8,  6, 24, 16,  1, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_WIDE,    768,  8,  8, 12, 24,  8,  1, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_NARROW,   96,  4,  4,  6, 12,  4,  8, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_NARROW,  192,  8,  4,  6, 24,  4,  8, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_NARROW,  384,  8,  8,  6, 24,  4,  4, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_XMAP,     48,  4, 12,  1, 12, 64,  1, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_XMAP,     96,  4,  4,  6, 12, 32,  1, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_XMAP,    192,  8,  4,  6, 24, 32,  1, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_XMAP,    384,  8,  8,  6, 24, 16,  1, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_XMAP,    768,  8,  8, 12, 24,  8,  1, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_WIDE,    100, 10, 10,  1, 10, 16,  2, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_WIDE,    160,  4,  4, 10, 20, 16,  2, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_WIDE,    200, 10, 20,  1, 20, 16,  1, 3, 3, false, 1>(m);
    register_modes_gen<float, VAR_WIDE,    250,  5,  5, 10, 10, 16,  1, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_WIDE,    320,  4,  4, 20, 20, 16,  1, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_WIDE,    400, 20, 20,  1, 20, 16,  1, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_WIDE,    500,  5, 10, 10, 10,  8,  1, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_WIDE,    800,  4, 10, 20, 20,  8,  1, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_WIDE,   1000, 10, 10, 10, 10,  8,  1, 1, 1, false, 1>(m);
    register_modes_gen<float, VAR_XMAP,    100, 10, 10,  1, 10, 32,  1, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_XMAP,    160,  4,  4, 10, 20, 32,  1, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_XMAP,    200, 10, 20,  1, 20, 32,  1, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_XMAP,    250,  5,  5, 10, 10, 16,  1, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_XMAP,    320,  4,  4, 20, 20, 16,  1, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_XMAP,    400, 20, 20,  1, 20, 16,  1, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_XMAP,    500,  5, 10, 10, 10,  8,  1, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_XMAP,    800,  4, 10, 20, 20,  8,  1, 2, 2, false, 1>(m);
    register_modes_gen<float, VAR_XMAP,   1000, 10, 10, 10, 10,  8,  1, 1, 1, false, 1>(m);
}

}  // namespace cpc
