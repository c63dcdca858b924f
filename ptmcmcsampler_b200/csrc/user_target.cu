// User targets on the device: the caller's CUDA source for logl / logp (the reference's arbitrary Python
// callables, ref PTMCMCSampler.py :108-109, :605-612, :1072-1086, moved onto the GPU) is compiled at run time with
// NVRTC together with the thread-per-chain MH kernels (mh_kernels.cuh) into a cubin for the device's architecture,
// loaded through the driver API and cached by content.  NVRTC and the driver library are dlopen()ed on first use, so
// the engine library itself loads on machines without a driver.
#include "user_target.h"

#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

namespace ptm {
namespace {

// ---- the few entry points used, resolved at run time --------------------------------------------------------------
typedef struct _nvrtcProgram *nvrtcProgram;
typedef int nvrtcResult;
typedef int CUresult;
typedef struct CUmod_st *CUmodule;
typedef struct CUstream_st *CUstream;

struct Api {
    void *nvrtc = nullptr, *cuda = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char *) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char *) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram *) = nullptr;
    const char *(*GetErrorString)(nvrtcResult) = nullptr;
    CUresult (*ModuleLoadData)(CUmodule *, const void *) = nullptr;
    CUresult (*ModuleGetFunction)(void **, CUmodule, const char *) = nullptr;
    CUresult (*LaunchKernel)(void *, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **,
                             void **) = nullptr;
    CUresult (*GetErrorStringCu)(CUresult, const char **) = nullptr;
    std::string err;
};

Api &api()
{
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char *n : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"})
            if (!a.nvrtc) a.nvrtc = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (!a.nvrtc) {
            a.err = "libnvrtc.so.12 not found: user targets need the CUDA run-time compiler";
            return;
        }
#define NV(sym) *(void **)(&a.sym) = dlsym(a.nvrtc, "nvrtc" #sym)
        NV(CreateProgram); NV(CompileProgram); NV(GetProgramLogSize); NV(GetProgramLog); NV(GetCUBINSize); NV(GetCUBIN);
        NV(DestroyProgram); NV(GetErrorString);
#undef NV
        if (!a.CreateProgram || !a.CompileProgram || !a.GetCUBIN) a.err = "libnvrtc lacks the expected entry points";
    });
    return a;
}

bool load_driver(Api &a, std::string &err)
{
    static std::once_flag once;
    std::call_once(once, [&a] {
        for (const char *n : {"libcuda.so.1", "libcuda.so"})
            if (!a.cuda) a.cuda = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (!a.cuda) return;
        *(void **)(&a.ModuleLoadData) = dlsym(a.cuda, "cuModuleLoadData");
        *(void **)(&a.ModuleGetFunction) = dlsym(a.cuda, "cuModuleGetFunction");
        *(void **)(&a.LaunchKernel) = dlsym(a.cuda, "cuLaunchKernel");
        *(void **)(&a.GetErrorStringCu) = dlsym(a.cuda, "cuGetErrorString");
    });
    if (!a.cuda || !a.ModuleLoadData || !a.ModuleGetFunction || !a.LaunchKernel) {
        err = "libcuda.so.1 not found: no CUDA driver";
        return false;
    }
    return true;
}

std::string source_dir()
{
    Dl_info info;
    if (dladdr((void *)&source_dir, &info) && info.dli_fname) {
        std::string p(info.dli_fname);
        const size_t k = p.rfind('/');
        return (k == std::string::npos ? std::string(".") : p.substr(0, k)) + "/csrc";
    }
    return "csrc";
}

const char *STUB_LOGL = "__device__ double user_logl(const double *, int, const double *) { return 0.0; }\n";
const char *STUB_LOGP = "__device__ double user_logp(const double *, int, const double *) { return 0.0; }\n";

std::mutex g_mu;
std::map<std::string, UserModule *> g_cache;

}  // namespace

// the translation unit handed to NVRTC
std::string user_translation_unit(const char *logl_src, const char *logp_src)
{
    std::string tu = "#define PTMCMC_USER_TARGET 1\n#include \"mh_common.cuh\"\n";
    // the usual <math.h> constants (NVRTC ships no standard headers; the math functions themselves are built in)
    tu += "#ifndef INFINITY\n#define INFINITY (__longlong_as_double(0x7ff0000000000000LL))\n#endif\n"
          "#ifndef NAN\n#define NAN (__longlong_as_double(0x7ff8000000000000LL))\n#endif\n"
          "#ifndef M_PI\n#define M_PI 3.14159265358979323846\n#endif\n"
          "#ifndef M_LN2\n#define M_LN2 0.693147180559945309417\n#endif\n"
          "#ifndef M_E\n#define M_E 2.7182818284590452354\n#endif\n";
    tu += "// ---- user log-likelihood\n";
    tu += (logl_src && *logl_src) ? std::string(logl_src) + "\n" : STUB_LOGL;
    tu += "// ---- user log-prior\n";
    tu += (logp_src && *logp_src) ? std::string(logp_src) + "\n" : STUB_LOGP;
    tu += "#include \"mh_kernels.cuh\"\n";
    return tu;
}

int user_compile_cubin(const char *logl_src, const char *logp_src, int cc_major, int cc_minor, std::vector<char> &cubin,
                       std::string &log)
{
    Api &a = api();
    if (!a.err.empty()) {
        log = a.err;
        return -1;
    }
    const std::string tu = user_translation_unit(logl_src, logp_src);
    nvrtcProgram prog = nullptr;
    nvrtcResult rc = a.CreateProgram(&prog, tu.c_str(), "ptmcmc_user_target.cu", 0, nullptr, nullptr);
    if (rc != 0) {
        log = std::string("nvrtcCreateProgram: ") + (a.GetErrorString ? a.GetErrorString(rc) : "error");
        return -1;
    }
    char arch[64];
    // architecture-specific target for Blackwell (sm_100a), plain sm_XY elsewhere
    snprintf(arch, sizeof arch, "--gpu-architecture=sm_%d%d%s", cc_major, cc_minor, (cc_major >= 9) ? "a" : "");
    const std::string inc = "-I" + source_dir();
    const char *opts[] = {arch, "--std=c++17", inc.c_str(), "-lineinfo", "-default-device"};
    rc = a.CompileProgram(prog, 5, opts);
    size_t n = 0;
    if (a.GetProgramLogSize && a.GetProgramLogSize(prog, &n) == 0 && n > 1) {
        log.resize(n);
        a.GetProgramLog(prog, &log[0]);
    }
    if (rc != 0) {
        log = std::string("NVRTC compilation of the user target failed (") + (a.GetErrorString ? a.GetErrorString(rc) : "error") +
              "):\n" + log;
        a.DestroyProgram(&prog);
        return -1;
    }
    size_t sz = 0;
    if (a.GetCUBINSize(prog, &sz) != 0 || sz == 0) {
        log = "NVRTC produced no cubin";
        a.DestroyProgram(&prog);
        return -1;
    }
    cubin.resize(sz);
    a.GetCUBIN(prog, cubin.data());
    a.DestroyProgram(&prog);
    return 0;
}

UserModule *user_module(const char *logl_src, const char *logp_src, int device, std::string &err)
{
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
    std::string key = std::to_string(device) + ":" + std::to_string(major) + std::to_string(minor) + "\x01";
    key += (logl_src ? logl_src : "");
    key += "\x02";
    key += (logp_src ? logp_src : "");
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_cache.find(key);
    if (it != g_cache.end()) return it->second;
    Api &a = api();
    std::vector<char> cubin;
    if (user_compile_cubin(logl_src, logp_src, major, minor, cubin, err) != 0) return nullptr;
    if (!load_driver(a, err)) return nullptr;
    cudaFree(0);  // the runtime's primary context is current on this thread
    CUmodule mod = nullptr;
    CUresult rc = a.ModuleLoadData(&mod, cubin.data());
    if (rc != 0) {
        const char *s = nullptr;
        if (a.GetErrorStringCu) a.GetErrorStringCu(rc, &s);
        err = std::string("cuModuleLoadData: ") + (s ? s : "error");
        return nullptr;
    }
    UserModule *m = new UserModule();
    m->module = mod;
    struct { void **fn; const char *name; } want[] = {{&m->mh, "mh_generic_kernel"}, {&m->init_eval, "init_eval_kernel"},
                                                      {&m->accept, "accept_kernel"}};
    for (auto &w : want) {
        if (a.ModuleGetFunction(w.fn, mod, w.name) != 0) {
            err = std::string("the compiled user target lacks ") + w.name;
            delete m;
            return nullptr;
        }
    }
    g_cache[key] = m;
    return m;
}

cudaError_t user_launch(void *fn, unsigned grid, unsigned block, cudaStream_t stream, void **args)
{
    Api &a = api();
    const CUresult rc = a.LaunchKernel(fn, grid, 1, 1, block, 1, 1, 0, (CUstream)stream, args, nullptr);
    return rc == 0 ? cudaSuccess : cudaErrorLaunchFailure;
}

}  // namespace ptm
