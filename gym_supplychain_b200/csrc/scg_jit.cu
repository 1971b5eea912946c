// scg_jit.cu -- run-time instantiation of the compile-time layered kernel (sc_layered_rollout, scg_sc_static.cuh)
// for chain shapes that have no pre-built instantiation.
//
// The throughput kernel is a template over the chain shape (S-F-W-R nodes per echelon, P products, lead time): every
// loop unrolls, the env state lives in registers, per-node constants are immediate operands.  libscg.so carries the
// instantiations of the reference's named presets; any OTHER layered full-bipartite chain (SupplyChainNPerStage /
// SupplyChainMultiProduct with custom `nodes_per_echelon`, product count or lead times -- supplychain_Nperstage_env.py:37-66)
// used to fall back to the runtime-topology kernel, ~20x slower.  Here the same hand-written template is instantiated
// for the user's shape with NVRTC when the handle first rolls out (10-20 s, once; the cubin is cached on disk) and
// launched through the driver API.  Nothing is traced or generated: the source is csrc/*.cuh as shipped.
//
// Both libnvrtc and the driver entry points are resolved at run time (dlopen / cudaGetDriverEntryPoint), so libscg.so has
// no link-time dependency on them; if anything is missing the handle keeps using the runtime-topology kernel.
#include <cuda.h>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <sstream>

#include "scg_sc_host.cuh"

namespace scg {

// ---------------------------------------------------------------------------------------
// NVRTC + driver entry points
// ---------------------------------------------------------------------------------------
namespace {

struct Nvrtc {
    void* lib = nullptr;
    int (*CreateProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*DestroyProgram)(void**) = nullptr;
    int (*CompileProgram)(void*, int, const char* const*) = nullptr;
    int (*GetProgramLogSize)(void*, size_t*) = nullptr;
    int (*GetProgramLog)(void*, char*) = nullptr;
    int (*GetCUBINSize)(void*, size_t*) = nullptr;
    int (*GetCUBIN)(void*, char*) = nullptr;
    int (*AddNameExpression)(void*, const char*) = nullptr;
    int (*GetLoweredName)(void*, const char*, const char**) = nullptr;
    int (*Version)(int*, int*) = nullptr;
    bool ok = false;
};

struct Driver {
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
    bool ok = false;
};

std::mutex g_mu;
Nvrtc g_rtc;
Driver g_drv;
bool g_tried = false, g_drv_tried = false;

template <class F> bool sym(void* lib, const char* name, F& out) {
    out = reinterpret_cast<F>(dlsym(lib, name));
    return out != nullptr;
}
template <class F> bool drv_sym(const char* name, F& out) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !p) return false;
    out = reinterpret_cast<F>(p);
    return true;
}

bool load_nvrtc() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_tried) return g_rtc.ok;
    g_tried = true;
    // The CUDA toolkit's NVRTC by PATH first (SCG_NVRTC, $CUDA_HOME, /usr/local/cuda): a bare soname resolves to whatever copy
    // the process has already loaded -- `import torch` preloads its bundled, older libnvrtc.so.12 -- and then the compiler
    // (and with it the cache key below) would depend on the import order: a cache warmed by tools/prejit.py without torch
    // was missed by every process that imported torch first.  The soname is the fallback for machines without a toolkit.
    std::vector<std::string> names;
    if (const char* e = getenv("SCG_NVRTC")) names.push_back(e);
    for (const char* var : {"CUDA_HOME", "CUDA_PATH"})
        if (const char* e = getenv(var)) names.push_back(std::string(e) + "/lib64/libnvrtc.so.12");
    names.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
    names.push_back("libnvrtc.so.12");
    names.push_back("libnvrtc.so");
    for (const std::string& name : names) {
        g_rtc.lib = dlopen(name.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (g_rtc.lib) break;
    }
    if (g_rtc.lib) {
        Nvrtc& r = g_rtc;
        r.ok = sym(r.lib, "nvrtcCreateProgram", r.CreateProgram) && sym(r.lib, "nvrtcDestroyProgram", r.DestroyProgram) &&
               sym(r.lib, "nvrtcCompileProgram", r.CompileProgram) && sym(r.lib, "nvrtcGetProgramLogSize", r.GetProgramLogSize) &&
               sym(r.lib, "nvrtcGetProgramLog", r.GetProgramLog) && sym(r.lib, "nvrtcGetCUBINSize", r.GetCUBINSize) &&
               sym(r.lib, "nvrtcGetCUBIN", r.GetCUBIN) && sym(r.lib, "nvrtcAddNameExpression", r.AddNameExpression) &&
               sym(r.lib, "nvrtcGetLoweredName", r.GetLoweredName) && sym(r.lib, "nvrtcVersion", r.Version);
    }
    return g_rtc.ok;
}

bool load_driver() {            // needs a CUDA device (the runtime hands out the driver entry points)
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_drv_tried) return g_drv.ok;
    g_drv_tried = true;
    Driver& d = g_drv;
    d.ok = drv_sym("cuModuleLoadData", d.ModuleLoadData) && drv_sym("cuModuleUnload", d.ModuleUnload) &&
           drv_sym("cuModuleGetFunction", d.ModuleGetFunction) && drv_sym("cuFuncSetAttribute", d.FuncSetAttribute) &&
           drv_sym("cuLaunchKernel", d.LaunchKernel);
    return d.ok;
}

bool load_backends() { return load_nvrtc() && load_driver(); }

std::string lib_dir() {
    Dl_info info;
    if (dladdr(reinterpret_cast<const void*>(&load_backends), &info) == 0 || !info.dli_fname) return ".";
    std::string p(info.dli_fname);
    const size_t k = p.find_last_of('/');
    return k == std::string::npos ? "." : p.substr(0, k);
}

bool read_file(const std::string& path, std::string& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::ostringstream ss; ss << f.rdbuf();
    out = ss.str();
    return true;
}

unsigned long long fnv1a(unsigned long long h, const std::string& s) {
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    return h;
}

// ---------------------------------------------------------------------------------------
// byte image of LayeredConsts<Real, S, F, W, R, P, L, ML, STOCH> (same member order, natural alignment).  The
// generated translation unit static_asserts the size and a handful of offsets, so a drift between this packer and the
// struct is a JIT compile error (-> fallback), never a silent mismatch.
// ---------------------------------------------------------------------------------------
static void init_text(std::string& out, double v) { char t[64]; snprintf(t, sizeof(t), "%.17g", v); out += t; if (!strpbrk(t, ".enai")) out += ".0"; out += ","; }
static void init_text(std::string& out, float v) { char t[64]; snprintf(t, sizeof(t), "%.9g", (double)v); out += t; if (!strpbrk(t, ".enai")) out += ".0"; out += "f,"; }
static void init_text(std::string& out, int v) { out += std::to_string(v) + ","; }
static void init_text(std::string& out, unsigned long long v) { out += std::to_string(v) + "ull,"; }

struct Packed {
    std::vector<unsigned char> b;
    size_t max_align = 1;
    std::vector<std::pair<std::string, size_t>> marks;
    std::string init;                    // the same values as a flat aggregate initializer, member order (brace elision)
    bool finite = true;
    template <class T> size_t put(const char* name, const T* v, size_t count) {
        for (size_t i = 0; i < count; ++i) {
            init_text(init, v[i]);
            if (!(v[i] == v[i]) || (double)v[i] > 1e300 || (double)v[i] < -1e300) finite = false;
        }
        const size_t al = alignof(T);
        if (al > max_align) max_align = al;
        size_t off = (b.size() + al - 1) / al * al;
        b.resize(off + count * sizeof(T), 0);
        if (count) memcpy(b.data() + off, v, count * sizeof(T));
        if (name) marks.emplace_back(name, off);
        return off;
    }
    void finish() { b.resize((b.size() + max_align - 1) / max_align * max_align, 0); }
};

template <class Real>
void pack_consts(const ScHostSpec& hs, int S, int F, int W, int R, int ML, Packed& pk) {
    const int P = hs.P, N = S + F + W + R, NP = N * P, E = S * F + F * W + W * R;
    (void)N;
    auto reals = [&](const char* name, const std::vector<double>& src, size_t first, size_t count) {
        std::vector<Real> t(count);
        for (size_t i = 0; i < count; ++i) t[i] = (Real)src[first + i];
        pk.put<Real>(name, t.data(), count);
    };
    reals("cap", hs.cap, 0, NP); reals(nullptr, hs.hold, 0, NP); reals(nullptr, hs.stock0, 0, NP); reals("max_ship", hs.max_ship, 0, NP);
    {
        std::vector<Real> init((size_t)NP * ML, (Real)0);
        for (int i = 0; i < NP; ++i)
            for (int k = hs.ibeg[i]; k < hs.ibeg[i + 1]; ++k) init[(size_t)i * ML + hs.itime[k] - 1] += (Real)hs.iamt[k];
        pk.put<Real>("init_slot", init.data(), init.size());
    }
    const std::vector<double> qs = sc_ring_scales(hs);
    const bool fixed_ring = sizeof(Real) == 8 && (!hs.stochL || SCG_FIXED_RING_STOCH != 0);
    {
        std::vector<Real> a(NP), b(NP);
        for (int i = 0; i < NP; ++i) { a[i] = (Real)qs[i]; b[i] = (Real)(1.0 / qs[i]); }
        pk.put<Real>("q_scale", a.data(), NP); pk.put<Real>(nullptr, b.data(), NP);
    }
    {
        const std::vector<unsigned long long> thr = sc_leadtime_thresholds(hs);
        std::vector<unsigned long long> t((size_t)ML);
        for (int j = 0; j < ML; ++j) t[j] = j < (int)thr.size() ? thr[j] : ~0ull;
        pk.put<unsigned long long>("lt_thr", t.data(), t.size());
    }
    {
        std::vector<float> a(NP), b(NP), c(NP);
        for (int i = 0; i < NP; ++i) {
            const double q = fixed_ring ? 1.0 / qs[i] : 1.0;
            a[i] = (float)(1.0 / hs.cap[i]);
            b[i] = (float)(hs.max_ship[i] > 0 ? q / hs.max_ship[i] : 0.0);
            c[i] = (float)(hs.max_ship[i] > 0 ? q / (hs.max_ship[i] * (double)(hs.maxL - (hs.avgL - 1))) : 0.0);
        }
        pk.put<float>("obs_stock", a.data(), NP); pk.put<float>(nullptr, b.data(), NP); pk.put<float>("obs_last", c.data(), NP);
    }
    reals("sup_cap", hs.sup_cap, 0, (size_t)S * P); reals(nullptr, hs.sup_cost, 0, (size_t)S * P);
    reals("proc_cap", hs.proc_cap, S, F);
    reals(nullptr, hs.ratio, (size_t)S * P, (size_t)F * P);
    {
        std::vector<Real> inv((size_t)F * P);
        for (int i = 0; i < F * P; ++i) inv[i] = (Real)(1.0 / hs.ratio[(size_t)S * P + i]);
        pk.put<Real>("inv_ratio", inv.data(), inv.size());
    }
    reals(nullptr, hs.proc_cost, (size_t)S * P, (size_t)F * P);
    reals("ecap", hs.ecap, 0, E); reals("ecost", hs.ecost, 0, (size_t)E * P);
    {
        const Real c4[4] = {(Real)hs.c_unmet, (Real)hs.c_xstock, (Real)hs.c_xproc, (Real)hs.c_xship};
        pk.put<Real>("c_unmet", c4, 1); pk.put<Real>(nullptr, c4 + 1, 3);
    }
    std::vector<Real> dlo(P), dspan(P);
    std::vector<float> odem(P);
    std::vector<int> dloi(P), dcnt(P);
    int all_uniform = 1;
    for (int p = 0; p < P; ++p) {
        dlo[p] = (Real)hs.dlo[p]; dspan[p] = (Real)(hs.dhi[p] - hs.dlo[p]);
        odem[p] = (float)(1.0 / (hs.dhi[p] - hs.dlo[p]));
        dloi[p] = (int)hs.dlo[p]; dcnt[p] = (int)hs.dhi[p] - (int)hs.dlo[p] + 1;
        if (hs.dkind[p] != 0) all_uniform = 0;
    }
    pk.put<Real>("dlo", dlo.data(), P); pk.put<Real>(nullptr, dspan.data(), P);
    pk.put<float>("obs_dem", odem.data(), P);
    const float obs_time = (float)(1.0 / (double)hs.T);
    pk.put<float>("obs_time", &obs_time, 1);
    pk.put<int>("dlo_i", dloi.data(), P); pk.put<int>(nullptr, dcnt.data(), P);
    pk.put<int>("all_uniform", &all_uniform, 1);
    const int T = hs.T;
    pk.put<int>("T", &T, 1);
    pk.finish();
}

// The packer against the real struct, on instantiations that exercise every alignment case (odd counts, float and
// double, stochastic lead times).  Run once before the first run-time instantiation; a mismatch disables it.
template <class Real, int S, int F, int W, int R, int P, int L, int ML, bool STOCH>
bool packer_matches() {
    ScHostSpec hs;
    const int N = S + F + W + R, NP = N * P, E = S * F + F * W + W * R;
    hs.N = N; hs.P = P; hs.E = E; hs.T = 77; hs.avgL = L; hs.maxL = ML; hs.stochL = STOCH ? 1 : 0;
    hs.c_unmet = 3.5; hs.c_xstock = 4.5; hs.c_xproc = 5.5; hs.c_xship = 6.5;
    double v = 1.0;
    auto fill = [&v](std::vector<double>& a, size_t n) { a.resize(n); for (auto& x : a) { x = v; v += 0.37; } };
    fill(hs.cap, NP); fill(hs.hold, NP); fill(hs.stock0, NP); fill(hs.max_ship, NP); fill(hs.sup_cap, NP); fill(hs.sup_cost, NP);
    fill(hs.proc_cap, N); fill(hs.ratio, NP); fill(hs.proc_cost, NP); fill(hs.ecap, E); fill(hs.ecost, (size_t)E * P);
    fill(hs.dlo, P); fill(hs.dhi, P);
    for (int p = 0; p < P; ++p) hs.dhi[p] += 100.0;
    hs.dkind.assign(P, 0);
    {   // layered full-bipartite edges (sc_ring_scales walks them)
        const int cnt[4] = {S, F, W, R}, base[4] = {0, S, S + F, S + F + W};
        hs.ebeg.assign(N + 1, 0);
        for (int e = 0; e < 4; ++e)
            for (int j = 0; j < cnt[e]; ++j) {
                const int n = base[e] + j;
                hs.ebeg[n + 1] = hs.ebeg[n] + (e < 3 ? cnt[e + 1] : 0);
                for (int i = 0; e < 3 && i < cnt[e + 1]; ++i) hs.edst.push_back(base[e + 1] + i);
            }
    }
    hs.ibeg.assign(NP + 1, 0);
    for (int i = 0; i < NP; ++i) {
        hs.ibeg[i + 1] = hs.ibeg[i] + 1;
        hs.itime.push_back(1 + i % ML);
        hs.iamt.push_back(10.0 + i);
    }
    LayeredConsts<Real, S, F, W, R, P, L, ML, STOCH> c;
    sc_fill_consts<Real, S, F, W, R, P, L, ML, STOCH>(hs, c);
    Packed pk;
    pack_consts<Real>(hs, S, F, W, R, ML, pk);
    return pk.b.size() == sizeof(c) && memcmp(pk.b.data(), &c, sizeof(c)) == 0;
}
bool packer_selftest() {
    static const bool ok = packer_matches<float, 3, 3, 3, 3, 2, 2, 2, false>() && packer_matches<double, 2, 2, 2, 2, 1, 2, 4, true>() &&
                           packer_matches<float, 1, 2, 3, 5, 3, 3, 3, false>() && packer_matches<double, 5, 4, 7, 10, 3, 2, 3, true>() &&
                           packer_matches<float, 2, 1, 1, 1, 1, 1, 1, false>();
    return ok;
}

struct JitKernel {
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
};

}  // namespace

struct ScJit {
    bool failed = false;
    int S = 0, F = 0, W = 0, R = 0, ML = 0, threads = 0, minb = 0;
    int split_g = 0;           // > 0: the node-parallel kernel (scg_sc_split.cuh) with that many lanes per env
    int act_tiles = 1;         // action tiles per warp (2 = tile s+1 is copied during period s)
    bool one_tile_define = false;   // layered kernel: compile with -DSCG_DBUF_ACTIONS=0 (its own rule would take two tiles)
    bool baked = false;        // the spec's constants are compile-time constants of the instantiation (SCG_JIT_BAKE=1)
    size_t smem = 0;
    Packed consts;
    JitKernel k[2];            // [random actions ?]
};

// ---------------------------------------------------------------------------------------
// shape matching
// ---------------------------------------------------------------------------------------
static bool jit_shape(const ScHostSpec& hs, int s[4]) {
    int S = 0, F = 0, R = 0;
    for (int n = 0; n < hs.N; ++n) {
        if (hs.is_sup[n]) ++S;
        else if (hs.proc_cap[n] > 0) ++F;
        else if (hs.last[n]) ++R;
    }
    const int W = hs.N - S - F - R;
    if (S < 1 || F < 1 || W < 1 || R < 1 || hs.P < 1 || hs.P > SC_MAXP) return false;
    const int ML = hs.stochL ? hs.maxL : hs.avgL;
    if (ML < 1 || ML > 16 || hs.avgL < 1) return false;
    if (!sc_is_layered(hs, S, F, W, R, hs.P, hs.avgL, ML, hs.stochL != 0)) return false;
    s[0] = S; s[1] = F; s[2] = W; s[3] = R;
    return true;
}

static bool jit_plan(ScJit& j, const ScHostSpec& hs, int mode);

bool sc_jit_match(const ScHostSpec& hs, int mode) {
    const char* e = getenv("SCG_JIT");
    if (e && e[0] == '0') return false;
    ScJit probe;
    return jit_plan(probe, hs, mode) && load_backends();
}

int sc_jit_split_lanes(const ScHostSpec& hs, int mode) {
    ScJit probe;
    return jit_plan(probe, hs, mode) ? probe.split_g : 0;
}

void sc_jit_free(ScJit* j) {
    if (!j) return;
    for (auto& k : j->k) if (k.mod && g_drv.ok) g_drv.ModuleUnload(k.mod);
    delete j;
}

// ---------------------------------------------------------------------------------------
// compile (or load from the disk cache) one instantiation
// ---------------------------------------------------------------------------------------
static int jit_compile(const ScJit& j, const ScHostSpec& hs, int mode, bool random_act, std::string& cubin, std::string& lowered) {
    const char* real = mode == SCG_MODE_FAST_F32 ? "float" : "double";
    const char* io = mode == SCG_MODE_FAST_F64 ? "double" : "float";
    char name[512];
    std::ostringstream src;
    std::string baked_inc;
    if (j.split_g > 0) {
        snprintf(name, sizeof(name), "scg::sc_split_rollout<%s, %s, %d, %d, %d, %d, %d, %d, %d, %s, %d, %d, %d>", real, io, j.S, j.F, j.W, j.R,
                 hs.P, hs.avgL, j.split_g, random_act ? "true" : "false", j.threads, j.minb, j.act_tiles);
        src << "#include \"scg_sc_split.cuh\"\n";
    } else {
        snprintf(name, sizeof(name), "scg::sc_layered_rollout<%s, %s, %d, %d, %d, %d, %d, %d, %d, %s, %s, %d, %d>", real, io, j.S, j.F, j.W,
                 j.R, hs.P, hs.avgL, j.ML, hs.stochL ? "true" : "false", random_act ? "true" : "false", j.threads, j.minb);
        std::ostringstream ctype;
        ctype << "LayeredConsts<" << real << ", " << j.S << ", " << j.F << ", " << j.W << ", " << j.R << ", " << hs.P << ", " << hs.avgL
              << ", " << j.ML << ", " << (hs.stochL ? "true" : "false") << ">";
        if (j.baked) {
            // the spec's constants become compile-time constants of this instantiation (immediate operands, folded
            // products): the kernel reads kBaked instead of its parameter block
            baked_inc = "namespace scg { __device__ constexpr " + ctype.str() + " kBaked = {" + j.consts.init + "}; }\n";
            src << "#define SCG_BAKED_CONSTS_INC \"scg_baked.inc\"\n";
        }
        src << "#include \"scg_sc_static.cuh\"\nnamespace scg {\nusing JitC = " << ctype.str() << ";\n"
            << "static_assert(sizeof(JitC) == " << j.consts.b.size() << ", \"LayeredConsts layout: host packer out of date\");\n";
        src << "}\n";
    }
    const std::string dir = lib_dir(), csrc = dir + "/csrc";
    std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-I" + csrc};
    if (mode == SCG_MODE_FAST_F64) opts.push_back("--fmad=false");      // same contract as scg_sc_f64.cu
    if (j.split_g == 0 && j.one_tile_define) opts.push_back("-DSCG_DBUF_ACTIONS=0");
    if (const char* extra = getenv("SCG_JIT_DEFINES")) {                // kernel A/B experiments: "-DX=1 -DY=0"
        std::istringstream is(extra);
        for (std::string tok; is >> tok;) opts.push_back(tok);
    }

    // cache key: everything that determines the cubin
    unsigned long long h = 1469598103934665603ull;
    h = fnv1a(h, name); h = fnv1a(h, src.str()); h = fnv1a(h, baked_inc);
    for (const auto& o : opts) if (o.compare(0, 2, "-I") != 0) h = fnv1a(h, o);
    for (const char* f : {"/scg_common.cuh", "/scg_sc_device.cuh", "/scg_sc_static.cuh", "/scg_sc_split.cuh"}) {
        std::string text;
        if (!read_file(csrc + f, text)) { set_error("jit: cannot read %s%s", csrc.c_str(), f); return SCG_ESTATE; }
        h = fnv1a(h, text);
    }
    int vmaj = 0, vmin = 0;
    g_rtc.Version(&vmaj, &vmin);
    h = fnv1a(h, std::to_string(vmaj * 100 + vmin));
    const char* cdir = getenv("SCG_JIT_CACHE");
    const std::string cache = cdir && cdir[0] ? std::string(cdir) : dir + "/_jit";
    char key[32];
    snprintf(key, sizeof(key), "%016llx", h);
    const std::string f_cubin = cache + "/" + key + ".cubin", f_name = cache + "/" + key + ".name";

    if (!(read_file(f_cubin, cubin) && read_file(f_name, lowered) && !cubin.empty() && !lowered.empty())) {
        fprintf(stderr, "[scg] compiling %s for this chain shape (NVRTC, once; cached under %s)\n", name, cache.c_str());
        void* prog = nullptr;
        const char* inc_src[1] = {baked_inc.c_str()};
        const char* inc_name[1] = {"scg_baked.inc"};
        if (g_rtc.CreateProgram(&prog, src.str().c_str(), "scg_jit.cu", baked_inc.empty() ? 0 : 1, baked_inc.empty() ? nullptr : inc_src,
                                baked_inc.empty() ? nullptr : inc_name) != 0) { set_error("jit: nvrtcCreateProgram failed"); return SCG_ESTATE; }
        g_rtc.AddNameExpression(prog, name);
        std::vector<const char*> copts;
        for (const auto& o : opts) copts.push_back(o.c_str());
        const int rc = g_rtc.CompileProgram(prog, (int)copts.size(), copts.data());
        if (rc != 0) {
            size_t n = 0;
            g_rtc.GetProgramLogSize(prog, &n);
            std::string log(n, '\0');
            if (n) g_rtc.GetProgramLog(prog, &log[0]);
            set_error("jit: NVRTC failed (%d): %.600s", rc, log.c_str());
            g_rtc.DestroyProgram(&prog);
            return SCG_ESTATE;
        }
        const char* low = nullptr;
        size_t n = 0;
        if (g_rtc.GetLoweredName(prog, name, &low) != 0 || !low || g_rtc.GetCUBINSize(prog, &n) != 0 || n == 0) {
            set_error("jit: no cubin / lowered name");
            g_rtc.DestroyProgram(&prog);
            return SCG_ESTATE;
        }
        lowered = low;
        cubin.resize(n);
        g_rtc.GetCUBIN(prog, &cubin[0]);
        g_rtc.DestroyProgram(&prog);
        mkdir(cache.c_str(), 0755);                          // best effort: a read-only tree just recompiles next time
        // both files appear atomically (several ranks may be compiling the same kernel): write, then rename; the
        // reader needs both, and the .name is renamed last
        const std::string sfx = ".tmp" + std::to_string((long long)getpid());
        std::ofstream fc(f_cubin + sfx, std::ios::binary), fn(f_name + sfx, std::ios::binary);
        if (fc && fn) {
            fc.write(cubin.data(), (std::streamsize)cubin.size());
            fn << lowered;
            fc.close(); fn.close();
            rename((f_cubin + sfx).c_str(), f_cubin.c_str());
            rename((f_name + sfx).c_str(), f_name.c_str());
        }
    }
    return SCG_OK;
}

static int jit_build(ScJit& j, const ScRun& r, bool random_act, JitKernel& out) {
    std::string cubin, lowered;
    const int rc = jit_compile(j, *r.hs, r.mode, random_act, cubin, lowered);
    if (rc != SCG_OK) return rc;
    if (g_drv.ModuleLoadData(&out.mod, cubin.data()) != CUDA_SUCCESS) { set_error("jit: cuModuleLoadData failed"); out.mod = nullptr; return SCG_ECUDA; }
    if (g_drv.ModuleGetFunction(&out.fn, out.mod, lowered.c_str()) != CUDA_SUCCESS) { set_error("jit: kernel %s not found in the cubin", lowered.c_str()); return SCG_ECUDA; }
    if (g_drv.FuncSetAttribute(out.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)j.smem) != CUDA_SUCCESS) {
        set_error("jit: %zu bytes of shared memory refused", j.smem);
        return SCG_ECUDA;
    }
    return SCG_OK;
}

// shape, constants image and launch shape of the instantiation for (spec, mode)
static bool jit_plan(ScJit& j, const ScHostSpec& hs, int mode) {
    int s[4];
    if (mode == SCG_MODE_EXACT_F64 || !jit_shape(hs, s) || !packer_selftest()) return false;
    if ((size_t)hs.N * hs.P * (1 + (hs.stochL ? hs.maxL : hs.avgL)) > 768) return false;   // state far beyond the register file
    j.S = s[0]; j.F = s[1]; j.W = s[2]; j.R = s[3];
    j.ML = hs.stochL ? hs.maxL : hs.avgL;
    const bool f64 = mode != SCG_MODE_FAST_F32;
    const size_t es = mode == SCG_MODE_FAST_F64 ? 8 : 4;
    if (f64) pack_consts<double>(hs, j.S, j.F, j.W, j.R, j.ML, j.consts);
    else pack_consts<float>(hs, j.S, j.F, j.W, j.R, j.ML, j.consts);
    // Chains whose state overflows one thread's registers: G lanes per env (scg_sc_split.cuh).  G = the smallest power of
    // two that leaves a lane at most 50 state values (measured on 5-4-7-10 and 10 per stage, profiles/r2_split_*.log:
    // fewer lanes spill, more lanes idle on the echelons that are not a multiple of G); constant lead times, fp32 I/O.
    // SCG_SPLIT_G / SCG_SPLIT_MINB override (experiments; G = 0 keeps the one-thread-per-env kernel).
    // registers the one-thread-per-env kernel would need for its state: fp32 stocks + ring, or fp64 stocks + the 32-bit
    // fixed-point ring of the mixed mode; measured to work up to ~165 (2-3-2-4, 3 products, lead time 3: 0.51 of the HBM peak)
    const int np_all = hs.N * hs.P;
    const int full_state_regs = f64 ? np_all * 2 + np_all * j.ML : np_all * (1 + j.ML);
    if (!hs.stochL && es == 4 && full_state_regs > (f64 ? 180 : 150)) {
        int g = 0;
        for (int cand : {2, 4, 8}) {
            const int kt = (j.S + cand - 1) / cand + (j.F + cand - 1) / cand + (j.W + cand - 1) / cand + (j.R + cand - 1) / cand;
            g = cand;
            if (kt * hs.P * (1 + j.ML) <= 50) break;
        }
        if (const char* e = getenv("SCG_SPLIT_G")) g = atoi(e);
        if (g == 2 || g == 4 || g == 8) {
            j.split_g = g;
            size_t wtile = (size_t)(32 / g) * (hs.A + hs.O) * es;
            const size_t budget_s = (size_t)227 * 1024 - 128 - 1024;
            // fp64 state: one-warp blocks (bulk-copy addresses in uniform registers, +3 % measured); fp32: no difference, 4 warps
            int warps = f64 ? 1 : 4;
            if (const char* e = getenv("SCG_SPLIT_WARPS")) warps = atoi(e) > 0 ? atoi(e) : warps;
            while (warps > 1 && (size_t)warps * wtile > budget_s / 2) warps >>= 1;
            if (wtile <= budget_s) {
                j.threads = 32 * warps;
                // fp64 state with 4 or fewer lanes per env needs the whole register file (255 per thread); else 168
                const int by_regs = 65536 / (j.threads * ((f64 && g <= 4) ? 255 : 168));
                // a second action tile per warp (the copy of tile s+1 overlaps period s, no L2 prefetch) where the register
                // budget, not shared memory, still decides the occupancy
                const size_t wtile2 = wtile + (size_t)(32 / g) * hs.A * es;
                int tiles = (f64 && (int)(((size_t)227 * 1024) / (128 + (size_t)warps * wtile2 + 1024)) >= by_regs) ? 2 : 1;
                if (const char* e = getenv("SCG_SPLIT_TILES")) tiles = atoi(e) == 2 ? 2 : 1;
                j.act_tiles = tiles;
                if (tiles == 2) wtile = wtile2;
                j.smem = 128 + (size_t)warps * wtile;
                const int by_smem = (int)(((size_t)227 * 1024) / (j.smem + 1024));
                int minb = by_smem < by_regs ? by_smem : by_regs;
                if (const char* e = getenv("SCG_SPLIT_MINB")) minb = atoi(e);
                j.minb = minb < 1 ? 1 : (minb > 32 ? 32 : minb);
                return true;
            }
            j.split_g = 0;
        }
    }
    if (const char* e = getenv("SCG_JIT_BAKE")) j.baked = e[0] == '1' && j.consts.finite;
    // the whole state should stay in registers; the warp tiles must fit shared memory
    size_t tile = (size_t)32 * (hs.A + hs.O) * es;
    const size_t budget = (size_t)227 * 1024 - 128 - 1024;
    // state registers: fp32 stocks + ring; fp64 stocks + 32-bit fixed-point ring (mixed, constant lead time); else all fp64
    const bool fixed_ring = f64 && es == 4 && (!hs.stochL || SCG_FIXED_RING_STOCH != 0);
    const int state_regs = fixed_ring ? np_all * 2 + np_all * j.ML : np_all * (1 + j.ML) * (f64 ? 2 : 1);
    int regs = (state_regs + 88 + 7) / 8 * 8;
    // an fp64-state kernel beyond 168 registers wants the whole file: it lives on instruction-level parallelism
    if (f64 && regs > 168) regs = 255;
    regs = regs < 64 ? 64 : (regs > 255 ? 255 : regs);
    // Block size.  One-warp blocks keep every bulk-copy address in uniform registers and give the finest occupancy granularity;
    // measured (profiles/r2_onewarp_blocks_ab_v2.log): +6..12 % with stochastic lead times, +16 % for an fp32 kernel beyond 168
    // registers (4-4-4-4: 11 warps per SM instead of 10), -2 % where the kernel is already HBM-bound -> only in those two cases.
    int warps = regs > 168 ? 2 : 4;
    if (hs.stochL || (!f64 && regs > 168)) warps = 1;
    if (const char* e = getenv("SCG_JIT_WARPS")) warps = atoi(e) > 0 ? atoi(e) : warps;
    while (warps > 1 && (size_t)warps * tile > budget) warps >>= 1;
    if (tile > budget) return false;                                                       // one warp tile must fit shared memory
    j.threads = 32 * warps;
    const size_t hdr = 128 + (hs.stochL ? ((size_t)j.ML * j.ML * (f64 ? 8 : 4) + 127) / 128 * 128 : 0);     // layered_smem_header
    const int by_regs = 65536 / (j.threads * regs);
    // The kernel's own rule gives the fp64-state / fp32-I/O kernels a second action tile (layered_action_tiles).  Keep it only
    // where registers, not shared memory, still decide the occupancy: 3-3-3-3 with 3 products dropped from 8 to 4 warps per SM
    // and from 0.70 to 0.41 of the HBM peak with it (profiles/r2_midsize_chains_v4.log); otherwise compile with one tile.
    const int rule_tiles = (f64 && es == 4) ? (np_all >= 8 ? layered_action_tiles<double, float, 8>() : 1) : 1;
    const size_t tile2 = tile + (size_t)32 * hs.A * es;
    j.act_tiles = (rule_tiles == 2 && (int)(((size_t)227 * 1024) / (hdr + (size_t)warps * tile2 + 1024)) >= by_regs) ? 2 : 1;
    j.one_tile_define = rule_tiles == 2 && j.act_tiles == 1;
    if (j.act_tiles == 2) tile = tile2;
    j.smem = hdr + (size_t)warps * tile;
    const int by_smem = (int)(((size_t)227 * 1024) / (j.smem + 1024));
    const int minb = by_smem < by_regs ? by_smem : by_regs;
    j.minb = minb < 1 ? 1 : (minb > 32 ? 32 : minb);
    return true;
}

// Compile (into the disk cache) without a GPU: build-time warm-up of the cache and the CPU-side test of the packer.
int sc_jit_precompile(const ScHostSpec& hs, int mode, bool random_act) {
    if (mode == SCG_MODE_EXACT_F64) { set_error("jit: the exact mode has no compile-time kernel"); return SCG_EINVAL; }
    if (!load_nvrtc()) { set_error("jit: libnvrtc not found"); return SCG_ESTATE; }
    ScJit j;
    if (!jit_plan(j, hs, mode)) { set_error("jit: not a layered full-bipartite chain"); return SCG_EINVAL; }
    std::string cubin, lowered;
    return jit_compile(j, hs, mode, random_act, cubin, lowered);
}

// Launch the run-time instantiation for this handle; any failure is reported once and the caller falls back to the
// runtime-topology kernel (SCG_ESTATE / SCG_ECUDA returned, *cache marked failed).
int sc_jit_run(ScJit** cache, const ScRun& r) {
    if (!*cache) {
        ScJit* j = new ScJit();
        *cache = j;
        if (!load_backends() || !jit_plan(*j, *r.hs, r.mode)) { j->failed = true; set_error("jit: not available for this spec"); return SCG_ESTATE; }
    }
    ScJit& j = **cache;
    if (j.failed) return SCG_ESTATE;
    JitKernel& k = j.k[r.actions == nullptr ? 1 : 0];
    if (!k.fn) {
        const int rc = jit_build(j, r, r.actions == nullptr, k);
        if (rc != SCG_OK) {
            j.failed = true;
            fprintf(stderr, "[scg] run-time kernel instantiation failed (%s); using the runtime-topology kernel\n", scg_last_error());
            return rc;
        }
    }
    const bool f64 = r.mode != SCG_MODE_FAST_F32;
    const long long n = f64 ? reinterpret_cast<const ScState<double>*>(r.state)->n : reinterpret_cast<const ScState<float>*>(r.state)->n;
    ScInputs in{r.inj_demand, r.inj_leadtimes, r.key, r.first_env, r.auto_reset, r.err};
    int n_steps = r.n_steps;
    const void* actions = r.actions;
    void* obs = r.obs; void* reward = r.reward; unsigned char* done = r.done;
    void* args_layered[] = {(void*)j.consts.b.data(), const_cast<void*>(r.dev_spec), const_cast<void*>(r.state), &in, &n_steps, &actions, &obs, &reward, &done};
    void* args_split[] = {const_cast<void*>(r.dev_spec), const_cast<void*>(r.state), &in, &n_steps, &actions, &obs, &reward, &done};
    void** args = j.split_g > 0 ? args_split : args_layered;
    const long long per_block = j.split_g > 0 ? (long long)j.threads / j.split_g : (long long)j.threads;      // envs per block
    const unsigned blocks = (unsigned)((n + per_block - 1) / per_block);
    if (g_drv.LaunchKernel(k.fn, blocks, 1, 1, (unsigned)j.threads, 1, 1, (unsigned)j.smem, (CUstream)r.stream, args, nullptr) != CUDA_SUCCESS) {
        set_error("jit: cuLaunchKernel failed");
        return SCG_ECUDA;
    }
    return SCG_OK;
}

}  // namespace scg
