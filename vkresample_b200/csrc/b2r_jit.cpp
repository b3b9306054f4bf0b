// b2r_jit.cpp -- plan-time JIT of statically scheduled kernels for sizes that have no ahead-of-time
// instantiation (b2r_static_sizes.h).
//
// The reference builds its FFT plan by generating GLSL per axis and compiling it with glslang at plan
// time (shaderGenVkFFT vkFFT.h:4495-4642, compile :7446-7521), so every size runs code with its
// constants baked in.  This is the B200 counterpart: the kernel templates of b2r_kernels.cuh are
// instantiated for the plan's exact schedule (StaticFft<N, T, radices...>) with NVRTC for sm_100a,
// loaded through the driver API, and launched exactly like the ahead-of-time instantiations.
// libnvrtc and libcuda are dlopen()ed on first use, so the library has no link-time dependency on
// them; if either is missing (or B2R_JIT=0) the plan falls back to the dynamic kernels.
// Compiled cubins are cached under $B2R_CACHE_DIR (default ~/.cache/b2resample).
#include "b2r_jit.h"

#include <algorithm>

#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <sstream>
#include <vector>

namespace b2r {
namespace {

// ---- minimal driver-API / NVRTC surface, resolved at run time ------------------------------------
typedef int CUresult;
typedef struct CUmod_st* CUmodule;
typedef struct CUfunc_st* CUfunction;
typedef struct CUstream_st* CUstream;
typedef struct _nvrtcProgram* nvrtcProgram;
constexpr int CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES = 8;

struct Api {
    bool ok = false;
    std::string why;
    CUresult (*cuModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*cuModuleUnload)(CUmodule) = nullptr;
    CUresult (*cuModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*cuFuncSetAttribute)(CUfunction, int, int) = nullptr;
    CUresult (*cuFuncGetAttribute)(int*, int, CUfunction) = nullptr;   // optional (B2R_JIT_VERBOSE)
    CUresult (*cuLaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                               void**, void**) = nullptr;
    CUresult (*cuOccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
    CUresult (*cuGetErrorString)(CUresult, const char**) = nullptr;
    int (*nvrtcCreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*nvrtcDestroyProgram)(nvrtcProgram*) = nullptr;
    int (*nvrtcCompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    int (*nvrtcAddNameExpression)(nvrtcProgram, const char*) = nullptr;
    int (*nvrtcGetLoweredName)(nvrtcProgram, const char*, const char**) = nullptr;
    int (*nvrtcGetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    int (*nvrtcGetCUBIN)(nvrtcProgram, char*) = nullptr;
    int (*nvrtcGetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    int (*nvrtcGetProgramLog)(nvrtcProgram, char*) = nullptr;
    int (*nvrtcVersion)(int*, int*) = nullptr;
};

template <class F> bool sym(void* lib, const char* name, F* out) {
    *out = reinterpret_cast<F>(dlsym(lib, name));
    return *out != nullptr;
}

Api& api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        void* cu = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!cu) { a.why = "libcuda.so.1 not found"; return; }
        void* rtc = nullptr;
        std::vector<std::string> cands;
        if (const char* e = getenv("B2R_NVRTC")) cands.push_back(e);
        if (const char* e = getenv("CUDA_HOME")) cands.push_back(std::string(e) + "/lib64/libnvrtc.so.12");
        cands.insert(cands.end(), {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so", "libnvrtc.so"});
        for (const auto& c : cands)
            if ((rtc = dlopen(c.c_str(), RTLD_NOW))) break;
        if (!rtc) { a.why = "libnvrtc.so.12 not found (set B2R_NVRTC or CUDA_HOME)"; return; }
        sym(cu, "cuFuncGetAttribute", &a.cuFuncGetAttribute);
        bool ok = sym(cu, "cuModuleLoadData", &a.cuModuleLoadData) && sym(cu, "cuModuleUnload", &a.cuModuleUnload) &&
                  sym(cu, "cuModuleGetFunction", &a.cuModuleGetFunction) && sym(cu, "cuFuncSetAttribute", &a.cuFuncSetAttribute) &&
                  sym(cu, "cuLaunchKernel", &a.cuLaunchKernel) &&
                  sym(cu, "cuOccupancyMaxActiveBlocksPerMultiprocessor", &a.cuOccupancyMaxActiveBlocksPerMultiprocessor) &&
                  sym(cu, "cuGetErrorString", &a.cuGetErrorString) &&
                  sym(rtc, "nvrtcCreateProgram", &a.nvrtcCreateProgram) && sym(rtc, "nvrtcDestroyProgram", &a.nvrtcDestroyProgram) &&
                  sym(rtc, "nvrtcCompileProgram", &a.nvrtcCompileProgram) && sym(rtc, "nvrtcAddNameExpression", &a.nvrtcAddNameExpression) &&
                  sym(rtc, "nvrtcGetLoweredName", &a.nvrtcGetLoweredName) && sym(rtc, "nvrtcGetCUBINSize", &a.nvrtcGetCUBINSize) &&
                  sym(rtc, "nvrtcGetCUBIN", &a.nvrtcGetCUBIN) && sym(rtc, "nvrtcGetProgramLogSize", &a.nvrtcGetProgramLogSize) &&
                  sym(rtc, "nvrtcGetProgramLog", &a.nvrtcGetProgramLog) && sym(rtc, "nvrtcVersion", &a.nvrtcVersion);
        if (!ok) { a.why = "a driver / NVRTC entry point is missing"; return; }
        a.ok = true;
    });
    return a;
}

cudaError_t cu2rt(CUresult r) { return r == 0 ? cudaSuccess : cudaErrorLaunchFailure; }

std::string type_of(const Schedule& s) {
    std::ostringstream o;
    o << "b2r::StaticFft<" << s.n << ", " << s.threads;
    for (int i = 0; i < s.nst; ++i) o << ", " << s.radices[i];
    o << ">";
    return o.str();
}

std::string this_library_dir() {
    Dl_info info;
    if (dladdr((const void*)&this_library_dir, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        size_t k = p.rfind('/');
        return k == std::string::npos ? "." : p.substr(0, k);
    }
    return ".";
}

uint64_t fnv(const std::string& s, uint64_t h = 1469598103934665603ull) {
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    return h;
}
std::string slurp(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}
// The cubin cache holds GPU code that is loaded without further checks, so it must live in a directory only
// this user can write: B2R_CACHE_DIR or $HOME/.cache/b2resample, created 0700 and verified (owner == us, no
// group / other write bit).  No private directory (HOME unset, wrong owner, ...) -> "" = caching disabled;
// there is no fallback to a shared location such as /tmp.
std::string cache_dir() {
    std::string d;
    if (const char* e = getenv("B2R_CACHE_DIR")) d = e;
    else if (const char* home = getenv("HOME")) { if (*home) d = std::string(home) + "/.cache/b2resample"; }
    if (d.empty()) return d;
    for (size_t i = 1; i <= d.size(); ++i)
        if (i == d.size() || d[i] == '/') mkdir(d.substr(0, i).c_str(), 0700);
    struct stat st;
    if (stat(d.c_str(), &st) != 0 || !S_ISDIR(st.st_mode) || st.st_uid != geteuid() || (st.st_mode & (S_IWGRP | S_IWOTH))) return "";
    return d;
}
// minimal structural check of a cubin image (ELF64, little endian): header present, section and program header
// tables inside the image -- the driver walks them without knowing the buffer length
bool cubin_is_sane_elf(const std::string& img) {
    if (img.size() < 64 || memcmp(img.data(), "\x7f" "ELF", 4) != 0 || img[4] != 2 || img[5] != 1) return false;
    auto rd = [&](size_t off, int bytes) { unsigned long long v = 0; memcpy(&v, img.data() + off, (size_t)bytes); return v; };
    const unsigned long long phoff = rd(32, 8), shoff = rd(40, 8);
    const unsigned long long phentsize = rd(54, 2), phnum = rd(56, 2), shentsize = rd(58, 2), shnum = rd(60, 2);
    if (shoff > img.size() || shnum * shentsize > img.size() - shoff) return false;
    if (phoff > img.size() || phnum * phentsize > img.size() - phoff) return false;
    if (shentsize >= 64)
        for (unsigned long long i = 0; i < shnum; ++i) {      // every section with file contents lies inside the image
            const size_t sh = (size_t)(shoff + i * shentsize);
            const unsigned long long type = rd(sh + 4, 4), off = rd(sh + 24, 8), size = rd(sh + 32, 8);
            if (type != 8 /* SHT_NOBITS */ && (off > img.size() || size > img.size() - off)) return false;
        }
    return true;
}
// write `data` to `path` through a private temporary name; false (and nothing published) on any I/O error
bool publish(const std::string& path, const std::string& data, const std::string& tag) {
    const std::string tmp = path + tag;
    {
        std::ofstream f(tmp, std::ios::binary);
        f.write(data.data(), (std::streamsize)data.size());
        f.flush();
        if (!f.good()) { f.close(); remove(tmp.c_str()); return false; }
    }
    if (rename(tmp.c_str(), path.c_str()) != 0) { remove(tmp.c_str()); return false; }
    return true;
}

// per-plan state: one module with the kernels this plan needs
struct RowCtx { JitModule* m; CUfunction fn = nullptr, fn_c2c = nullptr; int threads = 0, ppb = 1, ppb_c2c = 1; bool bulk = false; int sms = 0; bool dbl = false; };
struct ColCtx { JitModule* m; CUfunction fn = nullptr; int threads = 0, cc = 4; bool dbl = false; };

}  // namespace

struct JitModule {
    CUmodule mod = nullptr;
    RowCtx r2c, c2r;
    ColCtx cols;
    CUfunction sharpen = nullptr, to_planar = nullptr, to_u8 = nullptr;
    bool sharpen_rows = false;
};

bool jit_available(std::string* why) {
    const char* e = getenv("B2R_JIT");
    if (e && atoi(e) == 0) { if (why) *why = "disabled by B2R_JIT=0"; return false; }
    Api& a = api();
    if (!a.ok && why) *why = a.why;
    return a.ok;
}

void jit_destroy(JitModule* m) {
    if (!m) return;
    if (m->mod && api().ok) api().cuModuleUnload(m->mod);
    delete m;
}

// ---- launch trampolines (the same shapes as the ahead-of-time launchers) --------------------------
namespace {
cudaError_t prep_row(size_t smem, const void* ctx) {
    const RowCtx* k = static_cast<const RowCtx*>(ctx);
    if (smem > 48 * 1024) return cu2rt(api().cuFuncSetAttribute(k->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem));
    return cudaSuccess;
}
cudaError_t prep_row_c2c(size_t smem, const void* ctx) {
    const RowCtx* k = static_cast<const RowCtx*>(ctx);
    if (smem > 48 * 1024) return cu2rt(api().cuFuncSetAttribute(k->fn_c2c, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem));
    return cudaSuccess;
}
cudaError_t prep_col(size_t smem, const void* ctx) {
    const ColCtx* k = static_cast<const ColCtx*>(ctx);
    if (smem > 48 * 1024) return cu2rt(api().cuFuncSetAttribute(k->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem));
    return cudaSuccess;
}
cudaError_t run_r2c(cudaStream_t s, const R2cArgs& a, int, size_t smem, const void* ctx) {
    const RowCtx* k = static_cast<const RowCtx*>(ctx);
    int pairs = 3 * a.dm.h / 2;
    char plan = 0;
    void* args[] = {(void*)&a.in, (void*)&a.spec, (void*)&a.tw, &plan, (void*)&a.dm, &pairs};
    return cu2rt(api().cuLaunchKernel(k->fn, (pairs + k->ppb - 1) / k->ppb, 1, 1, k->threads, k->ppb, 1, (unsigned)smem,
                                      (CUstream)s, args, nullptr));
}
cudaError_t run_c2r(cudaStream_t s, const C2rArgs& a, int, size_t smem, const void* ctx) {
    const RowCtx* k = static_cast<const RowCtx*>(ctx);
    int pairs = 3 * a.dm.up_h / 2;
    char plan = 0;
    float scale = a.scale;
    double scale_d = 1.0 / (double)a.dm.up_w;
    void* args[] = {(void*)&a.spec, (void*)&a.pre, (void*)&a.tw, &plan, (void*)&a.dm, &pairs, k->dbl ? (void*)&scale_d : (void*)&scale};
    unsigned grid = (unsigned)((pairs + k->ppb - 1) / k->ppb), by = (unsigned)k->ppb;
    if (k->bulk) {   // persistent: one CTA per resident slot
        int per_sm = 0;
        if (api().cuOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k->fn, k->threads, smem) != 0 || per_sm < 1) per_sm = 1;
        grid = (unsigned)std::min(pairs, k->sms * per_sm);
        by = 1;
    }
    return cu2rt(api().cuLaunchKernel(k->fn, grid, 1, 1, k->threads, by, 1, (unsigned)smem, (CUstream)s, args, nullptr));
}
cudaError_t run_c2c(cudaStream_t s, const C2rArgs& a, int, size_t smem, const void* ctx) {
    const RowCtx* k = static_cast<const RowCtx*>(ctx);
    int rows = 3 * a.dm.up_h;
    char plan = 0;
    float scale = a.scale;
    double scale_d = 1.0 / (double)a.dm.up_w;
    void* args[] = {(void*)&a.spec, (void*)&a.nyq, (void*)&a.pre, (void*)&a.tw, &plan, (void*)&a.dm, &rows, k->dbl ? (void*)&scale_d : (void*)&scale};
    return cu2rt(api().cuLaunchKernel(k->fn_c2c, (rows + k->ppb_c2c - 1) / k->ppb_c2c, 1, 1, k->threads, k->ppb_c2c, 1, (unsigned)smem,
                                      (CUstream)s, args, nullptr));
}
cudaError_t run_cols(cudaStream_t s, const ColsArgs& a, int, size_t smem, const void* ctx) {
    const ColCtx* k = static_cast<const ColCtx*>(ctx);
    char pf = 0, pi = 0;
    float scale = a.scale;
    double scale_d = 1.0 / (double)a.dm.up_h;
    void* args[] = {(void*)&a.in, (void*)&a.out, (void*)&a.tw_f, (void*)&a.tw_i, &pf, &pi, (void*)&a.dm,
                    k->dbl ? (void*)&scale_d : (void*)&scale, (void*)&a.nyq};
    return cu2rt(api().cuLaunchKernel(k->fn, (a.dm.nx + k->cc - 1) / k->cc, 3, 1, k->threads * k->cc, 1, 1, (unsigned)smem,
                                      (CUstream)s, args, nullptr));
}
}  // namespace

bool jit_build(const JitRequest& rq, JitModule** out_mod, RowImpl* r2c, ColImpl* cols, RowImpl* c2r, std::string* err) {
    Api& a = api();
    if (!a.ok) { *err = a.why; return false; }
    const std::string csrc = this_library_dir() + "/../csrc";
    const std::string hdr = slurp(csrc + "/b2r_kernels.cuh") + slurp(csrc + "/b2r_fft.cuh") + slurp(csrc + "/b2r_common.cuh") + slurp(csrc + "/b2r_cas.cuh") +
                            slurp(csrc + "/b2r_fused.cuh");
    if (hdr.empty()) { *err = "kernel headers not found next to the library (" + csrc + ")"; return false; }

    // ---- source: explicit instantiations of exactly the kernels this plan launches
    const bool dbl = rq.precision == 1;
    const size_t cb = dbl ? 16 : 8;   // bytes of one complex workspace element
    const char* tin = rq.precision == 2 ? "__half" : (dbl ? "double" : "float");
    std::vector<std::string> names;   // name expressions, in the order r2c, cols, c2r, c2c
    const int ppb_w = rq.ppb_w > 0 ? rq.ppb_w : std::max(1, std::min(8, 256 / rq.w.threads));
    const int ppb_uw = std::max(1, std::min(8, 256 / rq.uw.threads));
    if (rq.want_r2c) {
        std::ostringstream n;
        n << "b2r::k_r2c_rows<" << type_of(rq.w) << ", " << tin << ", " << ppb_w << ">";
        names.push_back(n.str());
    }
    if (rq.want_cols) {
        std::ostringstream n;
        n << "b2r::k_cols<" << type_of(rq.h) << ", " << type_of(rq.uh) << ", " << rq.cc << ">";
        names.push_back(n.str());
    }
    // the bulk-copy C2R kernel adds 4*nx staging elements to the workspace; when that does not fit one CTA
    // (e.g. fp64 3840 -> 7680) the direct-load kernel, whose footprint is the workspace alone, is compiled instead
    const size_t c2r_bulk_bytes = 16 + 4 * (size_t)c2r_stage_row_elems(rq.nx) * cb + (size_t)smem_padded_len(rq.uw.n) * cb;
    const bool c2r_bulk = c2r_bulk_bytes <= 227u * 1024u;
    if (rq.want_c2r) {
        std::ostringstream n;
        if (c2r_bulk) n << "b2r::k_c2r_rows_bulk<" << type_of(rq.uw) << ", " << tin << ", " << (rq.up2 ? "true" : "false") << ">";
        else n << "b2r::k_c2r_rows<" << type_of(rq.uw) << ", " << tin << ", 1, " << (rq.up2 ? "true" : "false") << ">";
        names.push_back(n.str());
        if (rq.c2c) {
            std::ostringstream m;
            m << "b2r::k_c2c_rows<" << type_of(rq.uw) << ", " << tin << ", " << ppb_uw << ">";
            names.push_back(m.str());
        }
    }
    const bool rows_sharpen = (rq.up_w % 4) == 0;
    if (rq.want_pixels) {
        const bool ragged = rows_sharpen && sharpen_rows_ragged(rq.up_w, sharpen_rows_block(rq.up_w));
        names.push_back(rows_sharpen ? std::string("b2r::k_sharpen_rows<") + tin + ", b2r::kSharpenRowsPerThread, " +
                                           (ragged ? "true>" : "false>")
                                     : std::string("b2r::k_sharpen<") + tin + ", 4>");
        names.push_back(std::string("b2r::k_u8_to_planar<") + tin + ">");
        names.push_back(std::string("b2r::k_planar_to_u8<") + tin + ">");
    }
    // name expressions instantiate the templates; the source only has to bring the templates in
    std::string source = dbl ? "#define B2R_REAL_IS_DOUBLE 1\n#include \"b2r_kernels.cuh\"\n" : "#include \"b2r_kernels.cuh\"\n";
    for (const auto& n : names) source += "// " + n + "\n";

    int vmaj = 0, vmin = 0;
    a.nvrtcVersion(&vmaj, &vmin);
    const uint64_t key = fnv(hdr, fnv(source + std::to_string(vmaj * 100 + vmin)));
    char keyhex[32];
    snprintf(keyhex, sizeof keyhex, "%016llx", (unsigned long long)key);
    const std::string cdir = cache_dir();   // "" = no private cache directory: compile every time
    const std::string cpath = cdir + "/" + keyhex + ".cubin", npath = cdir + "/" + keyhex + ".names";

    std::string cubin;
    std::vector<std::string> lowered;
    JitModule* m = nullptr;
    // attempt 0 may use a cached cubin; if the driver rejects it (truncated / foreign file) the entry is deleted
    // and attempt 1 compiles afresh
    for (int attempt = 0; attempt < 2 && !m; ++attempt) {
        bool from_cache = false;
        cubin.clear(); lowered.clear();
        if (attempt == 0 && !cdir.empty()) {
            cubin = slurp(cpath);
            if (!cubin.empty()) {
                // cuModuleLoadData takes no length: a truncated file would make the driver read past the buffer.
                // The .names file therefore starts with the size and checksum of the cubin it belongs to, and an
                // image that does not match (or is not an ELF whose section table lies inside it) is never loaded.
                std::istringstream nf(slurp(npath));
                std::string line;
                bool intact = false;
                if (std::getline(nf, line)) {
                    unsigned long long sz = 0, sum = 0;
                    if (sscanf(line.c_str(), "#cubin %llu %llx", &sz, &sum) == 2)
                        intact = sz == cubin.size() && sum == fnv(cubin) && cubin_is_sane_elf(cubin);
                }
                while (std::getline(nf, line)) if (!line.empty()) lowered.push_back(line);
                if (!intact || lowered.size() != names.size()) {
                    cubin.clear(); lowered.clear();
                    remove(npath.c_str()); remove(cpath.c_str());
                } else from_cache = true;
            }
        }
        if (cubin.empty()) {
            if (rq.cache_only) { *err = "no cached JIT build for this size"; return false; }
            nvrtcProgram prog = nullptr;
            if (a.nvrtcCreateProgram(&prog, source.c_str(), "b2r_jit.cu", 0, nullptr, nullptr) != 0) { *err = "nvrtcCreateProgram failed"; return false; }
            for (const auto& n : names) a.nvrtcAddNameExpression(prog, n.c_str());
            const std::string inc1 = "-I" + csrc;
            std::string cuda_inc = "-I/usr/local/cuda/include";
            if (const char* e = getenv("CUDA_HOME")) cuda_inc = std::string("-I") + e + "/include";
            const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", inc1.c_str(), cuda_inc.c_str(), "-default-device",
                                  "-lineinfo", "-diag-suppress=550"};
            const int rc = a.nvrtcCompileProgram(prog, 7, opts);
            if (rc != 0) {
                size_t ls = 0;
                a.nvrtcGetProgramLogSize(prog, &ls);
                std::string log(ls, 0);
                if (ls) a.nvrtcGetProgramLog(prog, &log[0]);
                *err = "NVRTC compilation failed: " + log.substr(0, 400);
                a.nvrtcDestroyProgram(&prog);
                return false;
            }
            for (const auto& n : names) {
                const char* ln = nullptr;
                a.nvrtcGetLoweredName(prog, n.c_str(), &ln);
                lowered.push_back(ln ? ln : "");
            }
            size_t cs = 0;
            a.nvrtcGetCUBINSize(prog, &cs);
            cubin.resize(cs);
            a.nvrtcGetCUBIN(prog, &cubin[0]);
            a.nvrtcDestroyProgram(&prog);
            // publish atomically (several worker threads / processes may build the same key at once): private
            // names first, stream state checked, then rename; the .names file goes last and gates the cache hit
            if (!cdir.empty()) {
                const std::string tag = "." + std::to_string((long long)getpid()) + "." + std::to_string((unsigned long long)(uintptr_t)&cubin);
                char head[64];
                snprintf(head, sizeof head, "#cubin %llu %016llx\n", (unsigned long long)cubin.size(), (unsigned long long)fnv(cubin));
                std::string nm = head;
                for (const auto& l : lowered) nm += l + "\n";
                if (publish(cpath, cubin, tag)) { if (!publish(npath, nm, tag)) remove(cpath.c_str()); }
            }
        }
        m = new JitModule();
        if (a.cuModuleLoadData(&m->mod, cubin.data()) != 0) {
            delete m;
            m = nullptr;
            if (from_cache) { remove(npath.c_str()); remove(cpath.c_str()); continue; }   // bad cache entry: drop it, recompile once
            *err = "cuModuleLoadData failed";
            return false;
        }
    }
    if (!m) { *err = "cuModuleLoadData failed"; return false; }
    size_t idx = 0;
    auto get = [&](CUfunction* f) { return a.cuModuleGetFunction(f, m->mod, lowered[idx++].c_str()) == 0; };
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    bool ok = true;
    auto report = [&](const char* what, CUfunction f) {   // B2R_JIT_VERBOSE=1: registers per thread of each kernel
        const char* e = getenv("B2R_JIT_VERBOSE");
        int regs = 0;
        if (e && atoi(e) != 0 && f && a.cuFuncGetAttribute && a.cuFuncGetAttribute(&regs, 4 /* NUM_REGS */, f) == 0)
            fprintf(stderr, "[b2r jit] %s: %d registers/thread\n", what, regs);
    };
    if (rq.want_r2c) {
        m->r2c.m = m; m->r2c.threads = rq.w.threads; m->r2c.ppb = ppb_w; m->r2c.dbl = dbl;
        ok = ok && get(&m->r2c.fn);
        report(names[idx - 1].c_str(), m->r2c.fn);
        *r2c = RowImpl{};
        r2c->name = "r2c_rows<jit>"; r2c->is_static = true; r2c->is_jit = true; r2c->sched = rq.w; r2c->ppb = ppb_w;
        r2c->smem = (size_t)ppb_w * smem_padded_len(rq.w.n) * cb;
        r2c->ctx = &m->r2c; r2c->prepare = &prep_row; r2c->r2c = &run_r2c;
    }
    if (rq.want_cols) {
        m->cols.m = m; m->cols.threads = rq.uh.threads; m->cols.cc = rq.cc; m->cols.dbl = dbl;
        ok = ok && get(&m->cols.fn);
        report(names[idx - 1].c_str(), m->cols.fn);
        *cols = ColImpl{};
        cols->name = "cols<jit>"; cols->is_static = true; cols->is_jit = true; cols->fwd = rq.h; cols->inv = rq.uh; cols->cc = rq.cc;
        cols->smem = (size_t)smem_padded_len(rq.uh.n * rq.cc) * cb;
        cols->ctx = &m->cols; cols->prepare = &prep_col; cols->launch = &run_cols;
    }
    if (rq.want_c2r) {
        m->c2r.m = m; m->c2r.threads = rq.uw.threads; m->c2r.ppb = c2r_bulk ? ppb_uw : 1; m->c2r.ppb_c2c = ppb_uw; m->c2r.bulk = c2r_bulk; m->c2r.sms = sms; m->c2r.dbl = dbl;
        ok = ok && get(&m->c2r.fn);
        report(names[idx - 1].c_str(), m->c2r.fn);
        if (rq.c2c) ok = ok && get(&m->c2r.fn_c2c);
        *c2r = RowImpl{};
        c2r->name = c2r_bulk ? "c2r_rows_bulk<jit>" : "c2r_rows<jit>"; c2r->is_static = true; c2r->is_jit = true; c2r->sched = rq.uw; c2r->ppb = 1;
        c2r->smem = c2r_bulk ? c2r_bulk_bytes : (size_t)smem_padded_len(rq.uw.n) * cb;
        c2r->ctx = &m->c2r; c2r->prepare = &prep_row; c2r->c2r = &run_c2r;
        c2r->c2c = &run_c2c; c2r->prepare_c2c = &prep_row_c2c; c2r->ppb_c2c = ppb_uw;
        c2r->smem_c2c = (size_t)ppb_uw * smem_padded_len(rq.uw.n) * cb;
    }
    if (rq.want_pixels) {
        m->sharpen_rows = rows_sharpen;
        ok = ok && get(&m->sharpen) && get(&m->to_planar) && get(&m->to_u8);
    }
    if (!ok) { *err = "a JIT-compiled kernel was not found in the module"; jit_destroy(m); return false; }
    *out_mod = m;
    return true;
}

cudaError_t jit_launch_sharpen(const JitModule* m, cudaStream_t s, const SharpenArgs& a) {
    void* args[] = {(void*)&a.pre, (void*)&a.out, (void*)&a.dm};
    if (m->sharpen_rows) {
        const int bx = sharpen_rows_block(a.dm.up_w), ry = kSharpenRowsPerThread;
        return cu2rt(api().cuLaunchKernel(m->sharpen, (a.dm.up_w / 4 + bx - 1) / bx, (a.dm.up_h + ry - 1) / ry, 3, bx, 1, 1, 0,
                                          (CUstream)s, args, nullptr));
    }
    return cu2rt(api().cuLaunchKernel(m->sharpen, (a.dm.up_w + 4 * 256 - 1) / (4 * 256), a.dm.up_h, 3, 256, 1, 1, 0, (CUstream)s,
                                      args, nullptr));
}
cudaError_t jit_launch_u8_to_planar(const JitModule* m, cudaStream_t s, const unsigned char* src, void* dst, const FrameDims& dm) {
    const size_t n4 = ((size_t)dm.w * dm.h + 3) / 4;
    void* args[] = {(void*)&src, (void*)&dst, (void*)&dm};
    return cu2rt(api().cuLaunchKernel(m->to_planar, (unsigned)((n4 + 255) / 256), 1, 1, 256, 1, 1, 0, (CUstream)s, args, nullptr));
}
cudaError_t jit_launch_planar_to_u8(const JitModule* m, cudaStream_t s, const void* src, unsigned char* dst, const FrameDims& dm) {
    const size_t n4 = ((size_t)dm.up_w * dm.up_h + 3) / 4;
    void* args[] = {(void*)&src, (void*)&dst, (void*)&dm};
    return cu2rt(api().cuLaunchKernel(m->to_u8, (unsigned)((n4 + 255) / 256), 1, 1, 256, 1, 1, 0, (CUstream)s, args, nullptr));
}

}  // namespace b2r

// Test hook (not part of include/b2resample.h): the structural check a cached cubin has to pass before it is
// handed to the driver; tests/test_abi.py feeds it whole, truncated and garbage images.
extern "C" int b2r_debug_cubin_image_ok(const void* image, size_t bytes) {
    if (!image) return 0;
    return b2r::cubin_is_sane_elf(std::string(static_cast<const char*>(image), bytes)) ? 1 : 0;
}
