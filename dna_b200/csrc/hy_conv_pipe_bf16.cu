// hyena-b200: persistent four-step pipeline, bf16 activations (hy_conv_pipe.cuh).
#include "hy_conv_pipe_launch.cuh"
namespace hy {
template int launch_conv_pipe<DT_BF16>(const ConvArgs&, int, int, int, float2*, unsigned*, void*);
}  // namespace hy
