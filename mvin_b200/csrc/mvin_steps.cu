// mvin_steps.cu -- explicit instantiation of the step templates (steps.cuh) for ONE embedding dimension:
// nvcc -DMVIN_DIM=<8|16|32|64|128> -c mvin_steps.cu -o steps_<d>.o   (mvin_b200/build.py compiles the five in parallel)
#include "steps.cuh"

#ifndef MVIN_DIM
#error "compile with -DMVIN_DIM=<8|16|32|64|128>"
#endif

namespace mvin_host {
template int forward_impl<MVIN_DIM>(mvin_handle_t, const int64_t*, const int32_t*, const int32_t*, const int32_t*, int, float*,
                                    float*, void*, cudaStream_t);
template int backward_init<MVIN_DIM>(mvin_handle_t, int, void*, cudaStream_t, cudaEvent_t, bool);
template int backward_impl<MVIN_DIM>(mvin_handle_t, const float*, int, float*, void*, cudaStream_t);
template int xchg_expand_impl<MVIN_DIM>(mvin_handle_t, const int64_t*, int, int32_t*, void*, cudaStream_t);
template int xchg_owner_impl<MVIN_DIM>(mvin_handle_t, int, bool, int, void*, cudaStream_t);
template int xchg_finish_impl<MVIN_DIM>(mvin_handle_t, int, void*, cudaStream_t);
}  // namespace mvin_host
