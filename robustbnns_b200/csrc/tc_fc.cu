// placeholder (replaced by the tcgen05 FC engine)
#include "common.cuh"
namespace rbnn {
int tc_supported(const rbnn_net*) { return 0; }
int tc_bank_refresh(rbnn_net*, int, int, cudaStream_t) { set_error("tcgen05 engine not built"); return 1; }
int tc_fc_input_grad_sum(rbnn_net*, int, const float*, const int32_t*, int, int, int, const float*, float*, cudaStream_t) { set_error("tcgen05 engine not built"); return 1; }
int tc_fc_forward_probs_sum(rbnn_net*, const float*, int, int, int, float*, cudaStream_t) { set_error("tcgen05 engine not built"); return 1; }
}
