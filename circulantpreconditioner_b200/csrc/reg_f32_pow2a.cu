// reg_f32_pow2a.cu -- power-of-two line lengths, strided lines with 128-byte rows (y / z passes), incl. the multi-rank builds.
// fp32: 16 lanes x 8 B = one 128-byte row.
#include "registry.cuh"

namespace cpc {

void fill_fast_f32_pow2a(std::map<FastKey<float>, FastEntry<float>> &m)
{
    register_modes<float, VAR_WIDE,     16, 16,  1,  1, 16, 16,  8, 2>(m);
    register_modes<float, VAR_WIDE,     32,  8,  4,  1,  8, 16,  4, 2>(m);
    register_modes<float, VAR_WIDE,     64,  8,  8,  1,  8, 16,  2, 2>(m);
    register_modes<float, VAR_WIDE,    128, 16,  8,  1, 16, 16,  2, 2>(m);
    register_modes<float, VAR_WIDE,    256, 16, 16,  1, 16, 16,  1, 2>(m);
    register_modes<float, VAR_WIDE,    512, 16,  8,  4, 16, 16,  1, 2>(m);
    register_modes<float, VAR_WIDE,   1024, 16,  8,  8, 16, 16,  1, 1>(m);
}

}  // namespace cpc
