// fp32-activation instantiations of the four-step column kernels for the column lengths M1 = 5 * 2^a
// (transform lengths 5 * 2^k; e.g. L = 160 000 -> M = 40 x 4096 instead of 2^18)
#define HY_CONV_ODD_TU 5
#include "hy_conv_launch_impl.cuh"
namespace hy {
template int launch_col_fwd_odd<DT_F32, 5>(const ConvArgs&, int, int, int, int, void*);
template int launch_col_inv_odd<DT_F32, 5>(const ConvArgs&, int, int, int, int, void*);
}  // namespace hy
