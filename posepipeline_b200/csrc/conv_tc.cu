// placeholder until the tcgen05 kernel lands
#include "kernels.h"
struct TcConvPlan { int dummy; };
cudaError_t tc_conv_plan_create(TcConvPlan**, const float*, float*, const float*, const float*, const float*, int, int, int, int, int, int, int) { return cudaErrorNotSupported; }
void tc_conv_plan_destroy(TcConvPlan*) {}
cudaError_t tc_conv_launch(TcConvPlan*, int, cudaStream_t) { return cudaErrorNotSupported; }
