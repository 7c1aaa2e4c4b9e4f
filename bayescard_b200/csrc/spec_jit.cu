// K-spec runtime: compile the generated source with NVRTC (dlopen'ed, never linked), cache the
// cubin on disk, load it with cudaLibraryLoadData and launch it.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <fstream>

#include "bc_internal.h"

void bc_spec_geometry(const bc_model& m, int* threads, int* min_blocks);

namespace {

typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
    void* h = nullptr;
    int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
    int (*CompileProgram)(nvrtcProgram, int, const char* const*);
    int (*GetCUBINSize)(nvrtcProgram, size_t*);
    int (*GetCUBIN)(nvrtcProgram, char*);
    int (*GetProgramLogSize)(nvrtcProgram, size_t*);
    int (*GetProgramLog)(nvrtcProgram, char*);
    int (*DestroyProgram)(nvrtcProgram*);
    const char* (*GetErrorString)(int);
};

bool load_nvrtc(Nvrtc& n) {
    static const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                                  "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* nm : names) {
        n.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (n.h) break;
    }
    if (!n.h) return false;
#define SYM(field, sym)                                             \
    *(void**)(&n.field) = dlsym(n.h, sym);                          \
    if (!n.field) return false;
    SYM(CreateProgram, "nvrtcCreateProgram")
    SYM(CompileProgram, "nvrtcCompileProgram")
    SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    SYM(GetCUBIN, "nvrtcGetCUBIN")
    SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    SYM(GetProgramLog, "nvrtcGetProgramLog")
    SYM(DestroyProgram, "nvrtcDestroyProgram")
    SYM(GetErrorString, "nvrtcGetErrorString")
#undef SYM
    return true;
}

std::string cache_path(const bc_model& m, const char* dir) {
    char name[64];
    snprintf(name, sizeof(name), "/spec_%016llx.cubin", (unsigned long long)bc_spec_hash_of(m));
    return std::string(dir) + name;
}

}  // namespace

int bc_spec_attach(bc_model* m, const void* image, size_t bytes) {
    BC_CUDA_CHECK(cudaSetDevice(m->device));
    cudaLibrary_t lib = nullptr;
    // the image must outlive the library only for the duration of the call (it is copied)
    cudaError_t e = cudaLibraryLoadData(&lib, image, nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e != cudaSuccess) {
        bc_set_error("cudaLibraryLoadData failed: %s", cudaGetErrorString(e));
        return BC_ECUDA;
    }
    cudaKernel_t kr = nullptr, kd = nullptr;
    e = cudaLibraryGetKernel(&kd, lib, "bc_spec_dense");
    if (e != cudaSuccess) {
        bc_set_error("specialised image has no bc_spec_dense: %s", cudaGetErrorString(e));
        cudaLibraryUnload(lib);
        return BC_ECUDA;
    }
    if (m->max_card <= 256) {
        e = cudaLibraryGetKernel(&kr, lib, "bc_spec_range8");
        if (e != cudaSuccess) {
            bc_set_error("specialised image has no bc_spec_range8: %s", cudaGetErrorString(e));
            cudaLibraryUnload(lib);
            return BC_ECUDA;
        }
    }
    if (m->spec_lib) cudaLibraryUnload(m->spec_lib);
    m->spec_lib = lib;
    m->spec_range8 = kr;
    m->spec_dense = kd;
    bc_spec_geometry(*m, &m->spec_threads, &m->spec_min_blocks);
    int nb = 0;
    if (kr && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)kr, m->spec_threads, 0) == cudaSuccess &&
        nb > 0)
        m->spec_min_blocks = nb;
    else
        cudaGetLastError();
    return BC_OK;
}

int bc_spec_build(bc_model* m, const char* cache_dir) {
    if (m->arena.empty()) {
        bc_set_error("model too large for a specialised kernel");
        return BC_ELIMIT;
    }
    std::string path;
    if (cache_dir && *cache_dir) {
        path = cache_path(*m, cache_dir);
        std::ifstream f(path, std::ios::binary);
        if (f) {
            std::string img((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
            if (!img.empty() && bc_spec_attach(m, img.data(), img.size()) == BC_OK) return BC_OK;
        }
    }
    Nvrtc n{};
    if (!load_nvrtc(n)) {
        bc_set_error("NVRTC (libnvrtc.so.12) could not be loaded: %s", dlerror() ? dlerror() : "missing symbols");
        return BC_ECOMPILE;
    }
    const std::string src = bc_spec_generate(*m);
    nvrtcProgram prog = nullptr;
    int rc = n.CreateProgram(&prog, src.c_str(), "bc_spec.cu", 0, nullptr, nullptr);
    if (rc) {
        bc_set_error("nvrtcCreateProgram: %s", n.GetErrorString(rc));
        return BC_ECOMPILE;
    }
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--fmad=true"};
    rc = n.CompileProgram(prog, 4, opts);
    if (rc) {
        size_t ls = 0;
        n.GetProgramLogSize(prog, &ls);
        std::string log(ls + 1, 0);
        if (ls) n.GetProgramLog(prog, &log[0]);
        bc_set_error("NVRTC compile failed (%s): %.800s", n.GetErrorString(rc), log.c_str());
        n.DestroyProgram(&prog);
        return BC_ECOMPILE;
    }
    size_t cs = 0;
    n.GetCUBINSize(prog, &cs);
    std::string cubin(cs, 0);
    n.GetCUBIN(prog, &cubin[0]);
    n.DestroyProgram(&prog);
    if (!path.empty()) {
        std::ofstream f(path, std::ios::binary);
        if (f) f.write(cubin.data(), (std::streamsize)cubin.size());
    }
    return bc_spec_attach(m, cubin.data(), cubin.size());
}

int bc_spec_launch(bc_model* m, const void* desc, size_t nq, int fmt, const uint32_t* fan_mask, float* out,
                   cudaStream_t stream) {
    cudaKernel_t k = fmt == BC_DESC_DENSE_F32 ? m->spec_dense : m->spec_range8;
    if (!k) {
        bc_set_error("no specialised kernel for descriptor format %d", fmt);
        return BC_ECOMPILE;
    }
    const unsigned char* d = static_cast<const unsigned char*>(desc);
    unsigned long long stride = (unsigned long long)bc_model_desc_stride(m, fmt);
    unsigned long long n = nq;
    void* args[] = {(void*)&d, (void*)&stride, (void*)&fan_mask, (void*)&out, (void*)&n};
    const int threads = m->spec_threads;
    long long grid = (long long)m->sm_count * m->spec_min_blocks;
    const long long needed = (long long)((nq + threads - 1) / threads);
    if (grid > needed) grid = needed;
    BC_CUDA_CHECK(cudaLaunchKernel((const void*)k, dim3((unsigned)grid), dim3(threads), args, 0, stream));
    bc_count_launch();
    return BC_OK;
}
