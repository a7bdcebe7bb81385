// reg_f64_pow2b.cu -- power-of-two line lengths, contiguous lines (x pass) and 4-lane strided lines (wave x pass).
// fp64: a quarter warp (8 lanes x 16 B) covers one 128-byte row of the [N][8] tile.
#include "registry.cuh"

namespace cpc {

void fill_fast_f64_pow2b(std::map<FastKey<double>, FastEntry<double>> &m)
{
    register_modes<double, VAR_NARROW,   16, 16,  1,  1, 16,  4, 32, 2>(m);
    register_modes<double, VAR_NARROW,   32,  8,  4,  1,  8,  4, 16, 2>(m);
    register_modes<double, VAR_NARROW,   64,  8,  8,  1,  8,  4,  8, 2>(m);
    register_modes<double, VAR_NARROW,  128, 16,  8,  1, 16,  4,  8, 2>(m);
    register_modes<double, VAR_NARROW,  256, 16, 16,  1, 16,  4,  4, 2>(m);
    register_modes<double, VAR_NARROW,  512,  8,  8,  8,  8,  4,  2, 2, 1>(m);
    register_modes<double, VAR_NARROW, 1024, 16,  8,  8, 16,  4,  2, 1>(m);
    register_modes<double, VAR_NARROW, 2048, 16, 16,  8, 16,  4,  1, 1>(m);
    register_modes<double, VAR_XMAP,     16, 16,  1,  1, 16, 64,  1, 2>(m);
    register_modes<double, VAR_XMAP,     32,  8,  4,  1,  8, 64,  1, 2>(m);
    register_modes<double, VAR_XMAP,     64,  8,  8,  1,  8, 32,  1, 2>(m);
    register_modes<double, VAR_XMAP,    128, 16,  8,  1, 16, 32,  1, 2>(m);
    register_modes<double, VAR_XMAP,    256, 16, 16,  1, 16, 16,  1, 2>(m);
    register_modes<double, VAR_XMAP,    512,  8,  8,  8,  8,  8,  1, 2>(m);
    register_modes<double, VAR_XMAP,   1024, 16,  8,  8, 16,  4,  1, 2>(m);
    register_modes<double, VAR_XMAP,   2048, 16, 16,  8, 16,  2,  1, 2>(m);
}

}  // namespace cpc
