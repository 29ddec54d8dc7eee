#include "conv_tc.cuh"
namespace demon {
bool tc_layer_supported(const ConvProblem&) { return false; }
int tc_layer_prepare(TcLayer&, const ConvProblem&, const float*, int) { return fail(DEMON_E_INVALID, "tcgen05 path not built"); }
void tc_layer_free(TcLayer&) {}
int conv_tc_launch(const TcLayer&, const ConvProblem&, cudaStream_t) { return fail(DEMON_E_INVALID, "tcgen05 path not built"); }
}
