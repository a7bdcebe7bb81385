// reg_f64_pow2a.cu -- power-of-two line lengths, strided lines with 128-byte rows (y / z passes), incl. the multi-rank builds.
// fp64: a quarter warp (8 lanes x 16 B) covers one 128-byte row of the [N][8] tile.
#include "registry.cuh"

namespace cpc {

void fill_fast_f64_pow2a(std::map<FastKey<double>, FastEntry<double>> &m)
{
    //                     variant        N   R0  R1  R2   E  TX   G MINB [MINB fused]
    register_modes<double, VAR_WIDE,     16, 16,  1,  1, 16,  8, 16, 2>(m);
    register_modes<double, VAR_WIDE,     32,  8,  4,  1,  8,  8,  8, 2>(m);
    register_modes<double, VAR_WIDE,     64,  8,  8,  1,  8,  8,  4, 2>(m);
    register_modes<double, VAR_WIDE,    128, 16,  8,  1, 16,  8,  4, 2>(m);
    register_modes<double, VAR_WIDE,    256, 16, 16,  1, 16,  8,  2, 2>(m);
    register_modes<double, VAR_WIDE,    512,  8,  8,  8,  8,  8,  1, 2, 1>(m);
    register_modes<double, VAR_WIDE,   1024, 16,  8,  8, 16,  8,  1, 1>(m);
    register_modes<double, VAR_WIDE2,   512,  8,  8,  8, 16,  8,  1, 2>(m);
    register_modes<double, VAR_WIDE2,   256,  8,  8,  4,  8,  8,  2, 2>(m);      // 512 thr, 64 regs
    register_modes<double, VAR_WIDE2,   128,  8,  4,  4,  8,  8,  4, 2>(m);      // 512 thr, 64 regs
    register_modes<double, VAR_SMALL,    256,  8,  8,  4,  8,  8,  1, 4>(m);      // 256 thr, 64 regs, 4 CTAs/SM
    register_modes<double, VAR_SLIM,   1024, 16,  8,  8, 16,  4,  1, 2>(m);
}

}  // namespace cpc
