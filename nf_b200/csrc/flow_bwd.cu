#include "common.cuh"
size_t nis_flow_bwd_scratch_floats(const DevFlow& F, int64_t B) { return 0; }
extern "C" int nis_flow_backward(const NisFlowDesc* desc, const float* params, const float* saved,
                                 const float* bn_saved, const void* grad_out, int32_t grad_dtype,
                                 float* grad_params, void* grad_in, int32_t bn_mode,
                                 void* workspace, size_t workspace_bytes, int64_t B, void* stream) {
    return NIS_EUNSUPPORTED;
}
