// reg_f32_r2x.cu -- 2 x (R0 x R1) kernels and the warp-per-line kernels (fft_r2x.cuh).
// fp32: 16 lanes x 8 B = one 128-byte row.
#include "registry.cuh"

namespace cpc {

void fill_fast_f32_r2x(std::map<FastKey<float>, FastEntry<float>> &m)
{
    register_r2x512_line<float>(m);
    register_r2x<float, 256, 16, 16>(m);
    register_r2x<float, 128, 16, 8>(m);
    register_line256<float>(m);
}

}  // namespace cpc
