// hyena-b200: persistent four-step pipeline, fp32 rows (activations, filter spectrum, dk) (hy_conv_pipe.cuh).
#include "hy_conv_pipe_launch.cuh"
namespace hy {
template int launch_conv_pipe<DT_F32>(const ConvArgs&, int, int, int, float2*, unsigned*, void*);
}  // namespace hy
