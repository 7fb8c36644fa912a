// fp32-activation instantiations of the four-step column kernels for column lengths with an odd factor
// (M1 = 5 * 2^a, 3 * 2^a: transform lengths 5 * 2^k and 3 * 2^k, e.g. L = 160 000 -> M = 40 x 4096 instead of 2^18)
#define HY_CONV_ODD_TU 1
#include "hy_conv_launch_impl.cuh"
namespace hy {
template int launch_col_fwd_odd<DT_F32>(const ConvArgs&, int, int, int, int, void*);
template int launch_col_inv_odd<DT_F32>(const ConvArgs&, int, int, int, int, void*);
}  // namespace hy
