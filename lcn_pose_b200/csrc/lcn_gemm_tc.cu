// bf16 tensor-core (tcgen05 / TMEM) mid-layer kernels -- placeholder translation unit; see DESIGN.md.
#include "lcn_internal.cuh"

bool lcn_tc_enabled() { return false; }

int lcn_tc_gemm(const lcn_model*, const WsLayout&, int, int, const __nv_bfloat16*, const char*, const float*,
                const __nv_bfloat16*, __nv_bfloat16*, float*, cudaStream_t) {
  lcn_set_error("tcgen05 GEMM not built");
  return LCN_ESTATE;
}
int lcn_tc_wgrad(const lcn_model*, const WsLayout&, const __nv_bfloat16*, const __nv_bfloat16*, float*, cudaStream_t) {
  lcn_set_error("tcgen05 wgrad not built");
  return LCN_ESTATE;
}
