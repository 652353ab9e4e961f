// temporary stub
#include <cuda_runtime.h>
#include <cstdlib>
#include "../../include/glx.h"
extern "C" void* glx_pinned_alloc(size_t bytes) { void* p = nullptr; if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr; return p; }
extern "C" void glx_pinned_free(void* p) { cudaFreeHost(p); }
extern "C" {
int glx_create(glx_ctx**, int) { return GLX_E_CUDA; }
void glx_destroy(glx_ctx*) {}
const char* glx_last_error(const glx_ctx*) { return ""; }
int glx_genotype_batch(glx_ctx*, const glx_config*, const glx_records*, const glx_sites*, glx_out*) { return GLX_E_CUDA; }
int glx_upload(glx_ctx*, const glx_config*, const glx_records*, const glx_sites*, glx_dev_batch**) { return GLX_E_CUDA; }
int glx_launch(glx_ctx*, glx_dev_batch*, void*) { return GLX_E_CUDA; }
int glx_download(glx_ctx*, glx_dev_batch*, glx_out*) { return GLX_E_CUDA; }
int glx_check(glx_ctx*, glx_dev_batch*) { return GLX_E_CUDA; }
uint64_t glx_out_bytes(const glx_dev_batch*) { return 0; }
uint64_t glx_in_bytes(const glx_dev_batch*) { return 0; }
int glx_launch_count(const glx_dev_batch*) { return 0; }
void glx_free_batch(glx_ctx*, glx_dev_batch*) {}
int glx_synchronize(glx_ctx*) { return GLX_E_CUDA; }
}
