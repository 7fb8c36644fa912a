// bf16-activation instantiations of the long-convolution kernels
#include "hy_conv_launch_impl.cuh"
namespace hy {
template int launch_fused_fwd<DT_BF16>(const ConvArgs&, int, int, void*);
template int launch_fused_bwd<DT_BF16>(const ConvArgs&, int, void*);
template int launch_fused_bwdg<DT_BF16>(const ConvArgs&, int, void*);
template int launch_col_fwd<DT_BF16>(const ConvArgs&, int, int, int, int, void*);
template int launch_col_inv<DT_BF16>(const ConvArgs&, int, int, int, int, void*);
}  // namespace hy
