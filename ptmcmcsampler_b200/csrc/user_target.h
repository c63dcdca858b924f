// Run-time compiled user targets (user_target.cu).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

namespace ptm {

struct UserModule {
    void *module = nullptr;
    void *mh = nullptr, *init_eval = nullptr, *accept = nullptr;  // CUfunction handles of the kernels of mh_kernels.cuh
};

// the source NVRTC sees: PTMCMC_USER_TARGET, the caller's user_logl / user_logp (stubs for absent ones), mh_kernels.cuh
std::string user_translation_unit(const char *logl_src, const char *logp_src);
// compile only (no device needed): cubin for compute capability major.minor; 0 on success, the NVRTC log in `log`
int user_compile_cubin(const char *logl_src, const char *logp_src, int cc_major, int cc_minor, std::vector<char> &cubin,
                       std::string &log);
// compiled + loaded module for `device`, cached by (device, source); nullptr and `err` on failure
UserModule *user_module(const char *logl_src, const char *logp_src, int device, std::string &err);
cudaError_t user_launch(void *fn, unsigned grid, unsigned block, cudaStream_t stream, void **args);

}  // namespace ptm
