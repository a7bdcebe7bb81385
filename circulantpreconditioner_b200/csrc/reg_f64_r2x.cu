// reg_f64_r2x.cu -- 2 x (R0 x R1) kernels and the warp-per-line kernels (fft_r2x.cuh).
// fp64: a quarter warp (8 lanes x 16 B) covers one 128-byte row of the [N][8] tile.
#include "registry.cuh"

namespace cpc {

void fill_fast_f64_r2x(std::map<FastKey<double>, FastEntry<double>> &m)
{
    register_r2x<double, 256, 16, 16>(m);
    register_r2x<double, 128, 16, 8>(m);
    register_r2x512_line<double>(m);
    register_line256<double>(m);
}

}  // namespace cpc
