// bf16-activation instantiations of the four-step column kernels for the column lengths M1 = 3 * 2^a
// (transform lengths 3 * 2^k; e.g. L = 160 000 -> M = 40 x 4096 instead of 2^18)
#define HY_CONV_ODD_TU 3
#include "hy_conv_launch_impl.cuh"
namespace hy {
template int launch_col_fwd_odd<DT_BF16, 3>(const ConvArgs&, int, int, int, int, void*);
template int launch_col_inv_odd<DT_BF16, 3>(const ConvArgs&, int, int, int, int, void*);
}  // namespace hy
