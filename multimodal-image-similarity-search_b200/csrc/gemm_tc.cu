// placeholder until the tcgen05 kernels land
#include "kernels.h"
namespace vs {
bool tensor_path_available() { return false; }
size_t tensor_workspace_bytes(int, int, int, int) { return 256; }
cudaError_t launch_tensor_topk(const TensorArgs&, const float*, int, int, void*, float*, int64_t*, int, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_tensor_filter(const TensorArgs&, const float*, int, float, void*, uint32_t*, int64_t, int, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_tensor_dedup(const TensorArgs&, int64_t, int64_t, float, int64_t, int64_t*, int64_t*, float*, unsigned long long*, void*, int, cudaStream_t) { return cudaErrorNotSupported; }
}
