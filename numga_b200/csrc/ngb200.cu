// numga_b200 native library: straight-line multivector programs -> sm_100a kernels.
//
//   IR (include/numga_b200.h)  --codegen-->  CUDA C++  --NVRTC (sm_100a, -lineinfo)-->  cubin
//   cubin cached on disk (in-tree _kcache/, filled at build time so the GPU box does not JIT)
//   cubin --cuModuleLoadData--> CUfunction --cuLaunchKernel--> caller's stream
//
// Data layout: structure-of-arrays. Component c of batch item i of an operand lives at
// ptr[c*comp_stride + i*batch_stride]. The vectorised kernel handles batch_stride == 1 with
// 128-bit loads/stores (4 x f32 or 2 x f64 items per thread per component row); every warp
// therefore touches 512 contiguous bytes per component row: fully coalesced, no shared memory
// staging needed because no data is reused between threads. Broadcast operands (batch_stride 0)
// and everything computed only from them are evaluated once per thread before the grid-stride
// loop (this is the "motor -> 4x4 matrix" pre-contraction of numga's sandwich_map, done in
// registers). The kernels are persistent: grid = SMs x resident CTAs, grid-stride over the batch.
//
// The CUDA driver is dlopen()ed lazily, so this library loads (and compiles kernels) on a
// machine without a GPU; launching there fails loudly with NGB200_ERR_CUDA. No CPU fallback.

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/numga_b200.h"

#define NGB_VERSION "0.1.0"

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

extern "C" const char *ngb200_last_error(void) { return g_err.c_str(); }

static std::atomic<int64_t> g_launches{0};
extern "C" int64_t ngb200_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------------------------------
// CUDA driver API, resolved at first use
// ------------------------------------------------------------------------------------------
#define NGB_STR2(x) #x
#define NGB_STR(x) NGB_STR2(x)
#define DRV_FUNCS(X)                                                                              \
    X(cuInit) X(cuGetErrorString) X(cuCtxGetCurrent) X(cuCtxSetCurrent) X(cuCtxGetDevice)         \
    X(cuDeviceGet) X(cuDevicePrimaryCtxRetain) X(cuDeviceGetAttribute) X(cuModuleLoadData) X(cuModuleUnload) \
    X(cuModuleGetFunction) X(cuLaunchKernel) X(cuFuncGetAttribute) X(cuFuncSetAttribute)          \
    X(cuOccupancyMaxActiveBlocksPerMultiprocessor) X(cuMemAlloc) X(cuMemFree)                     \
    X(cuMemcpyHtoDAsync) X(cuMemcpyDtoHAsync) X(cuStreamCreate)                \
    X(cuStreamSynchronize) X(cuStreamDestroy) X(cuMemAllocHost) X(cuMemFreeHost)

struct Driver {
#define X(n) decltype(&n) p_##n = nullptr;
    DRV_FUNCS(X)
#undef X
    bool ok = false;
    std::string why;
};
static Driver g_drv;
static std::once_flag g_drv_once;

static void load_driver() {
    void *h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        g_drv.why = "CUDA driver (libcuda.so.1) not found: no GPU on this machine; numga_b200 has no CPU fallback";
        return;
    }
#define X(n)                                                                                      \
    g_drv.p_##n = (decltype(&n))dlsym(h, NGB_STR(n));                                             \
    if (!g_drv.p_##n) { g_drv.why = std::string("missing driver symbol ") + NGB_STR(n); return; }
    DRV_FUNCS(X)
#undef X
    CUresult r = g_drv.p_cuInit(0);
    if (r != CUDA_SUCCESS) {
        g_drv.why = "cuInit failed (" + std::to_string((int)r) + "): no usable CUDA device; numga_b200 has no CPU fallback";
        return;
    }
    g_drv.ok = true;
}

static int need_driver() {
    std::call_once(g_drv_once, load_driver);
    if (!g_drv.ok) return fail(NGB200_ERR_CUDA, "%s", g_drv.why.c_str());
    return NGB200_OK;
}

static int cu_fail(CUresult r, const char *what) {
    const char *s = nullptr;
    if (g_drv.p_cuGetErrorString) g_drv.p_cuGetErrorString(r, &s);
    return fail(NGB200_ERR_CUDA, "%s failed: %s (%d)", what, s ? s : "?", (int)r);
}
#define CU(call)                                                                                  \
    do {                                                                                          \
        CUresult r_ = g_drv.p_##call;                                                             \
        if (r_ != CUDA_SUCCESS) return cu_fail(r_, #call);                                        \
    } while (0)

// make sure a context is current on this thread; returns its device ordinal
static int current_device(int *dev_out, int want_device /* -1 = whatever is current */) {
    CUcontext ctx = nullptr;
    CU(cuCtxGetCurrent(&ctx));
    if (ctx && want_device < 0) {
        CUdevice d;
        CU(cuCtxGetDevice(&d));
        *dev_out = (int)d;
        return NGB200_OK;
    }
    if (ctx) {
        CUdevice d;
        CU(cuCtxGetDevice(&d));
        if ((int)d == want_device) { *dev_out = want_device; return NGB200_OK; }
    }
    CUdevice d;
    CU(cuDeviceGet(&d, want_device < 0 ? 0 : want_device));
    CU(cuDevicePrimaryCtxRetain(&ctx, d));
    CU(cuCtxSetCurrent(ctx));
    *dev_out = (int)d;
    return NGB200_OK;
}

// Restores the calling thread's current context on scope exit (the host-buffer entry points take a device ordinal and
// may have to make another device's primary context current while they run).
struct CtxGuard {
    CUcontext prev = nullptr;
    bool armed = false;
    CtxGuard() { if (g_drv.ok && g_drv.p_cuCtxGetCurrent(&prev) == CUDA_SUCCESS) armed = true; }
    ~CtxGuard() {
        if (!armed) return;
        CUcontext now = nullptr;
        if (g_drv.p_cuCtxGetCurrent(&now) == CUDA_SUCCESS && now != prev) g_drv.p_cuCtxSetCurrent(prev);
    }
};

// ------------------------------------------------------------------------------------------
// program object
// ------------------------------------------------------------------------------------------
static const int kMaxDevices = 32;

struct Loaded {
    std::atomic<int> ready{0};
    CUmodule mod = nullptr;
    CUfunction fvec = nullptr, fscalar = nullptr, fscalar_u = nullptr, fstage = nullptr;
    int regs_vec = 0, regs_scalar = 0, regs_stage = 0;
    int grid_vec = 0, grid_scalar = 0, grid_stage = 0;  // persistent grid sizes (SMs x resident CTAs)
};

struct ngb200_program {
    int dtype = 0;
    uint32_t flags = 0;
    std::vector<int32_t> in_width, in_mode, out_width, out_mode;
    int n_params = 0;
    std::vector<double> consts;
    std::vector<ngb200_instr> instr;
    std::vector<ngb200_store> stores;
    int vec = 1;          // items per thread in the contiguous ("vector") kernel
    bool has_unit = false;      // `ngb_scalar_u` exists: the scalar kernel with batch stride 1 compiled in
    bool bdim_runtime = false;  // index arithmetic with blockDim.x (see generate)
    bool has_vec = false; // a contiguous kernel `ngb_vec` exists (batch stride 1 compiled in; vec 1 = plain 32/64-bit accesses)
    int tpb = 256;
    int minb = 1;         // __launch_bounds__ min resident CTAs per SM (caps registers at 65536 / (tpb * minb))
    // staged kernel (bulk async copies global -> shared, mbarrier pipeline); 0 = not generated
    int stage_tpb = 0, stage_stages = 0, stage_nx = 0;
    size_t stage_smem = 0;
    int n_live = 0, n_uniform = 0;
    bool cache_hit = false;
    std::string source, hash;
    std::vector<char> cubin;
    std::mutex mu;
    Loaded loaded[kMaxDevices];
    size_t param_bytes = 0;
    bool any_indexed = false;
    bool any_reduced = false;
    bool any_tiled = false;       // some input is NGB200_TILED: operands carry (inner, outer stride) -> 48-byte layout
    std::mutex host_mu;
    struct HostPipe *host = nullptr;   // staging resources of ngb200_program_run_host, created on first use
};

static void host_pipe_destroy(struct HostPipe *h);

static int esize(int dtype) { return dtype == NGB200_F64 ? 8 : 4; }

// ------------------------------------------------------------------------------------------
// code generation
// ------------------------------------------------------------------------------------------
static void appendf(std::string &s, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (n < (int)sizeof buf) { s.append(buf, n); return; }
    std::vector<char> big(n + 1);
    va_start(ap, fmt);
    vsnprintf(big.data(), big.size(), fmt, ap);
    va_end(ap);
    s.append(big.data(), n);
}

static std::string literal(double v, int dtype) {
    char buf[64];
    if (dtype == NGB200_F32) {
        float f = (float)v;
        if (std::isnan(f)) return "__int_as_float(0x7fc00000)";
        if (std::isinf(f)) return f > 0 ? "__int_as_float(0x7f800000)" : "__int_as_float(0xff800000)";
        snprintf(buf, sizeof buf, "%af", (double)f);
    } else {
        if (std::isnan(v)) return "__longlong_as_double(0x7ff8000000000000LL)";
        if (std::isinf(v)) return v > 0 ? "__longlong_as_double(0x7ff0000000000000LL)" : "__longlong_as_double(0xfff0000000000000LL)";
        snprintf(buf, sizeof buf, "%a", v);
    }
    std::string s(buf);
    if (s[0] == '-') s = "(" + s + ")";
    return s;
}

static int n_operands_of(int op) {
    switch (op) {
        case NGB200_OP_INPUT: case NGB200_OP_CONST: case NGB200_OP_PARAM: return 0;
        case NGB200_OP_NEG: case NGB200_OP_SQRT: case NGB200_OP_RSQRT: case NGB200_OP_RCP: case NGB200_OP_ABS:
        case NGB200_OP_EXP: case NGB200_OP_LOG: case NGB200_OP_SIN: case NGB200_OP_COS: case NGB200_OP_PARTNER: return 1;
        case NGB200_OP_FMA: case NGB200_OP_SELECT_GT0: return 3;
        default: return 2;
    }
}

static int env_int(const char *name, int dflt);

struct Gen {
    const ngb200_program &P;
    bool f64, faithful;
    bool approx = false;      // FAST fp32: MUFU-based reciprocal / square roots
    std::vector<char> live, uniform;
    std::vector<int> uslot;   // export slot of uniform values used by varying code
    std::vector<int> xoff;    // offset of each varying input in x[]
    std::vector<int> yoff;    // offset of each output in y[]
    int NX = 0, NU = 0, NY = 0;
    int max_live = 0;         // peak number of simultaneously live per-item values (see analyse)
    std::vector<int> order;   // emission order of the varying, non-input values of the body
    bool faithful_order = true;    // true: emit in IR order; false: input-frontier order (see analyse)
    // names used by the emitted uniform / body functions; a pipeline (several programs in one kernel) gives every
    // phase its own suffix and maps broadcast inputs onto its own parameter block
    std::string suffix;                       // "" -> ngb_uniform / ngb_body
    std::string params_type = "Params";
    std::vector<std::string> bcast_ref;       // per input k: "p.in[k]"-like expression with .ptr / .cs
    // pipelines only: inputs that are NOT preloaded into x[] but re-read from shared memory at every use (wide tables
    // in fp64: the registers they would occupy for the whole phase are worth more than the extra LDS); value = slot in xp[]
    std::vector<int> remat;
    int n_remat = 0;
    bool has_partner = false;     // the program exchanges values between the threads of an item pair (pipeline phases only)

    explicit Gen(const ngb200_program &p)
        : P(p), f64(p.dtype == NGB200_F64), faithful(p.flags & NGB200_FAITHFUL) {
        approx = !f64 && !faithful && !env_int("NGB200_PRECISE_MATH", 0);
        faithful_order = env_int("NGB200_IR_ORDER", 1) != 0;
    }

    // name of value v as seen from uniform (preamble) or varying (body) code
    std::string ref(int v, bool in_body) const {
        const ngb200_instr &I = P.instr[v];
        if (I.op == NGB200_OP_CONST) return literal(P.consts[I.a], P.dtype);
        if (in_body) {
            if (uniform[v]) return "u[" + std::to_string(uslot[v]) + "]";
            if (I.op == NGB200_OP_INPUT) {
                if (I.a < (int)remat.size() && remat[I.a] >= 0)
                    return "ngb_lds(xp[" + std::to_string(remat[I.a]) + "] + " + std::to_string(I.b) + ")";
                return "x[" + std::to_string(xoff[I.a] + I.b) + "]";
            }
            return "r" + std::to_string(v);
        }
        return "q" + std::to_string(v);
    }

    std::string expr(int v, bool in_body) const {
        const ngb200_instr &I = P.instr[v];
        std::string a = n_operands_of(I.op) >= 1 ? ref(I.a, in_body) : "";
        std::string b = n_operands_of(I.op) >= 2 ? ref(I.b, in_body) : "";
        std::string c = n_operands_of(I.op) >= 3 ? ref(I.c, in_body) : "";
        const char *sfx = f64 ? "" : "f";
        auto rn = [&](const char *op) { return std::string(f64 ? "__d" : "__f") + op + "_rn"; };
        // Every add / multiply is an explicitly rounded intrinsic and every FMA is one this generator chose
        // (plan_contraction), so the scalar and the packed kernels of a program compute bit-identical results
        // and nothing depends on the compiler's contraction heuristics.
        if (in_body && (I.op == NGB200_OP_ADD || I.op == NGB200_OP_SUB) && !fuse_mul.empty() && fuse_mul[v] >= 0) {
            const bool sub = I.op == NGB200_OP_SUB;
            const int m = fuse_mul[v];
            std::string ma = ref(P.instr[m].a, true), mb = ref(P.instr[m].b, true);
            std::string f = std::string("fma") + sfx;
            if (m == I.b) return f + "(" + (sub ? "-" + ma : ma) + ", " + mb + ", " + a + ")";
            return f + "(" + ma + ", " + mb + ", " + (sub ? "-" + b : b) + ")";
        }
        switch (I.op) {
            case NGB200_OP_ADD: return rn("add") + "(" + a + ", " + b + ")";
            case NGB200_OP_SUB: return rn("sub") + "(" + a + ", " + b + ")";
            case NGB200_OP_MUL: return rn("mul") + "(" + a + ", " + b + ")";
            case NGB200_OP_DIV: return faithful ? rn("div") + "(" + a + ", " + b + ")" : a + " / " + b;
            case NGB200_OP_NEG: return "-" + a;
            // FAST fp32: MUFU approximations (rcp 1 ulp, sqrt / rsqrt ~1-2 ulp, 0 / inf / nan behave as IEEE, subnormal
            // arguments and results flush to zero; PTX ISA "rcp/sqrt/rsqrt.approx.ftz.f32"): an IEEE divide costs
            // ~21 FFMA issue slots on sm_100, these cost one MUFU (the non-ftz forms add 3-6 fix-up instructions).
            case NGB200_OP_SQRT:
                if (approx) return "ngb_sqrt(" + a + ")";
                return faithful ? rn("sqrt") + "(" + a + ")" : std::string("sqrt") + sfx + "(" + a + ")";
            case NGB200_OP_RSQRT:
                if (approx) return "ngb_rsqrt(" + a + ")";
                if (!faithful && f64) return "rsqrt(" + a + ")";
                return faithful ? rn("div") + "(T(1), " + rn("sqrt") + "(" + a + "))"
                                : std::string("T(1) / sqrt") + sfx + "(" + a + ")";
            case NGB200_OP_RCP:
                if (approx) return "ngb_rcp(" + a + ")";
                return faithful ? rn("div") + "(T(1), " + a + ")" : "T(1) / " + a;
            case NGB200_OP_ABS: return std::string("fabs") + sfx + "(" + a + ")";
            case NGB200_OP_NAN_TO_NUM: return "ngb_nan_to_num(" + a + ", " + b + ")";
            case NGB200_OP_FMA: return std::string("fma") + sfx + "(" + a + ", " + b + ", " + c + ")";
            case NGB200_OP_EXP: return std::string("exp") + sfx + "(" + a + ")";
            case NGB200_OP_LOG: return std::string("log") + sfx + "(" + a + ")";
            case NGB200_OP_SIN: return std::string("sin") + sfx + "(" + a + ")";
            case NGB200_OP_COS: return std::string("cos") + sfx + "(" + a + ")";
            case NGB200_OP_MIN: return std::string("fmin") + sfx + "(" + a + ", " + b + ")";
            case NGB200_OP_MAX: return std::string("fmax") + sfx + "(" + a + ", " + b + ")";
            case NGB200_OP_SELECT_GT0: return "(" + a + " > T(0) ? " + b + " : " + c + ")";
            case NGB200_OP_ATAN2: return std::string("atan2") + sfx + "(" + a + ", " + b + ")";
            case NGB200_OP_PARTNER: return "__shfl_xor_sync(0xffffffffu, " + a + ", 1)";
        }
        return "T(0)";
    }

    // ---- packed fp32 pairs: value v of items (2k, 2k+1) lives in one float2 -----------------------------------
    // Scalar FFMA issues at half rate on sm_100 (one warp instruction per 2 cycles per SM sub-partition); the
    // packed FFMA2 / FMUL2 / FADD2 instructions do two FMAs per lane per issue slot and are the only way to reach
    // the FP32 peak. Compute-bound programs (exp/log chains, CGA products) therefore run two batch items per
    // float2 value. nvcc contracts a*b+c only for plain operators, so FAST programs get their FMAs from the
    // contraction pass below; FAITHFUL programs keep separately rounded __fmul2_rn / __fadd2_rn.
    std::vector<int> fused_into;   // MUL v is absorbed into the FMA emitted for ADD/SUB fused_into[v]
    std::vector<int> fuse_mul;     // for ADD/SUB v: which operand (the MUL) it absorbs, or -1

    void plan_contraction() {
        const int n = (int)P.instr.size();
        fused_into.assign(n, -1);
        fuse_mul.assign(n, -1);
        if (faithful) return;
        std::vector<int> uses(n, 0);
        for (int v = 0; v < n; ++v) {
            if (!live[v]) continue;
            const ngb200_instr &I = P.instr[v];
            const int ops[3] = {I.a, I.b, I.c};
            for (int j = 0; j < n_operands_of(I.op); ++j) uses[ops[j]]++;
        }
        for (const ngb200_store &st : P.stores) uses[st.value]++;
        for (int v = 0; v < n; ++v) {
            if (!live[v] || uniform[v]) continue;
            const ngb200_instr &I = P.instr[v];
            if (I.op != NGB200_OP_ADD && I.op != NGB200_OP_SUB) continue;
            const int cand[2] = {I.b, I.a};
            for (int m : cand) {
                const ngb200_instr &M = P.instr[m];
                if (M.op == NGB200_OP_MUL && !uniform[m] && uses[m] == 1 && fused_into[m] < 0 && I.a != I.b) {
                    fused_into[m] = v;
                    fuse_mul[v] = m;
                    break;
                }
            }
        }
    }

    std::string ref2(int v) const {
        const ngb200_instr &I = P.instr[v];
        if (I.op == NGB200_OP_CONST) return "ngb_dup(" + literal(P.consts[I.a], P.dtype) + ")";
        if (uniform[v]) return "ngb_dup(u[" + std::to_string(uslot[v]) + "])";
        if (I.op == NGB200_OP_INPUT) return "x[" + std::to_string(xoff[I.a] + I.b) + "]";
        return "r" + std::to_string(v);
    }

    std::string expr2(int v) const {
        const ngb200_instr &I = P.instr[v];
        const int k = n_operands_of(I.op);
        std::string a = k >= 1 ? ref2(I.a) : "", b = k >= 2 ? ref2(I.b) : "", c = k >= 3 ? ref2(I.c) : "";
        auto half = [&](const std::string &fn, const std::string &h) {   // scalar op on one half
            std::string r = fn;
            size_t pos;
            while ((pos = r.find("$a")) != std::string::npos) r.replace(pos, 2, a + "." + h);
            while ((pos = r.find("$b")) != std::string::npos) r.replace(pos, 2, b + "." + h);
            while ((pos = r.find("$c")) != std::string::npos) r.replace(pos, 2, c + "." + h);
            return r;
        };
        auto both = [&](const std::string &fn) { return "make_float2(" + half(fn, "x") + ", " + half(fn, "y") + ")"; };
        switch (I.op) {
            case NGB200_OP_ADD:
            case NGB200_OP_SUB: {
                const bool sub = I.op == NGB200_OP_SUB;
                int m = fuse_mul[v];
                if (m >= 0) {
                    const ngb200_instr &M = P.instr[m];
                    std::string ma = ref2(M.a), mb = ref2(M.b);
                    if (m == I.b)   // a +- (ma*mb)
                        return "__ffma2_rn(" + (sub ? "ngb_neg2(" + ma + ")" : ma) + ", " + mb + ", " + a + ")";
                    return "__ffma2_rn(" + ma + ", " + mb + ", " + (sub ? "ngb_neg2(" + b + ")" : b) + ")";   // (ma*mb) +- b
                }
                return "__fadd2_rn(" + a + ", " + (sub ? "ngb_neg2(" + b + ")" : b) + ")";
            }
            case NGB200_OP_MUL: return "__fmul2_rn(" + a + ", " + b + ")";
            case NGB200_OP_FMA: return "__ffma2_rn(" + a + ", " + b + ", " + c + ")";
            case NGB200_OP_NEG: return "ngb_neg2(" + a + ")";
            case NGB200_OP_DIV: return both(faithful ? "__fdiv_rn($a, $b)" : "$a / $b");
            case NGB200_OP_SQRT: return both(approx ? "ngb_sqrt($a)" : faithful ? "__fsqrt_rn($a)" : "sqrtf($a)");
            case NGB200_OP_RSQRT: return both(approx ? "ngb_rsqrt($a)" : faithful ? "__fdiv_rn(1.0f, __fsqrt_rn($a))" : "1.0f / sqrtf($a)");
            case NGB200_OP_RCP: return both(approx ? "ngb_rcp($a)" : faithful ? "__fdiv_rn(1.0f, $a)" : "1.0f / $a");
            case NGB200_OP_ABS: return both("fabsf($a)");
            case NGB200_OP_NAN_TO_NUM: return both("ngb_nan_to_num($a, $b)");
            case NGB200_OP_EXP: return both("expf($a)");
            case NGB200_OP_LOG: return both("logf($a)");
            case NGB200_OP_SIN: return both("sinf($a)");
            case NGB200_OP_COS: return both("cosf($a)");
            case NGB200_OP_MIN: return both("fminf($a, $b)");
            case NGB200_OP_MAX: return both("fmaxf($a, $b)");
            case NGB200_OP_SELECT_GT0: return both("($a > 0.0f ? $b : $c)");
            case NGB200_OP_ATAN2: return both("atan2f($a, $b)");
        }
        return "ngb_dup(0.0f)";
    }

    int analyse() {
        const int n = (int)P.instr.size();
        live.assign(n, 0);
        uniform.assign(n, 0);
        uslot.assign(n, -1);
        // validate + uniformity (forward)
        for (int v = 0; v < n; ++v) {
            const ngb200_instr &I = P.instr[v];
            if (I.op < 0 || I.op >= NGB200_OP_COUNT_) return fail(NGB200_ERR_INVALID, "instr %d: bad opcode %d", v, I.op);
            if (I.op == NGB200_OP_PARTNER) has_partner = true;
            int k = n_operands_of(I.op);
            const int ops[3] = {I.a, I.b, I.c};
            for (int j = 0; j < k; ++j)
                if (ops[j] < 0 || ops[j] >= v) return fail(NGB200_ERR_INVALID, "instr %d: operand %d out of range", v, ops[j]);
            if (I.op == NGB200_OP_INPUT) {
                if (I.a < 0 || I.a >= (int)P.in_width.size() || I.b < 0 || I.b >= P.in_width[I.a])
                    return fail(NGB200_ERR_INVALID, "instr %d: input (%d,%d) out of range", v, I.a, I.b);
                uniform[v] = P.in_mode[I.a] == NGB200_BROADCAST;
            } else if (I.op == NGB200_OP_CONST) {
                if (I.a < 0 || I.a >= (int)P.consts.size()) return fail(NGB200_ERR_INVALID, "instr %d: const out of range", v);
                uniform[v] = 1;
            } else if (I.op == NGB200_OP_PARAM) {
                if (I.a < 0 || I.a >= P.n_params) return fail(NGB200_ERR_INVALID, "instr %d: param out of range", v);
                uniform[v] = 1;
            } else {
                bool u = true;
                for (int j = 0; j < k; ++j) u = u && uniform[ops[j]];
                uniform[v] = u;
            }
        }
        // liveness (backward from the stores)
        for (const ngb200_store &s : P.stores) {
            if (s.operand < 0 || s.operand >= (int)P.out_width.size() || s.component < 0 ||
                s.component >= P.out_width[s.operand] || s.value < 0 || s.value >= n)
                return fail(NGB200_ERR_INVALID, "store (%d,%d,%d) out of range", s.operand, s.component, s.value);
            live[s.value] = 1;
        }
        for (int v = n - 1; v >= 0; --v) {
            if (!live[v]) continue;
            const ngb200_instr &I = P.instr[v];
            int k = n_operands_of(I.op);
            const int ops[3] = {I.a, I.b, I.c};
            for (int j = 0; j < k; ++j) live[ops[j]] = 1;
        }
        // uniform values referenced from varying code or stored need an export slot
        auto want_slot = [&](int v) {
            if (uniform[v] && P.instr[v].op != NGB200_OP_CONST && uslot[v] < 0) uslot[v] = NU++;
        };
        for (int v = 0; v < n; ++v) {
            if (!live[v] || uniform[v]) continue;
            const ngb200_instr &I = P.instr[v];
            int k = n_operands_of(I.op);
            const int ops[3] = {I.a, I.b, I.c};
            for (int j = 0; j < k; ++j) want_slot(ops[j]);
        }
        for (const ngb200_store &s : P.stores) want_slot(s.value);
        xoff.assign(P.in_width.size(), 0);
        for (size_t k = 0; k < P.in_width.size(); ++k) {
            xoff[k] = NX;
            if (P.in_mode[k] != NGB200_BROADCAST) NX += P.in_width[k];
        }
        yoff.assign(P.out_width.size(), 0);
        for (size_t k = 0; k < P.out_width.size(); ++k) { yoff[k] = NY; NY += P.out_width[k]; }
        // Emission order of the per-item body: stable sort by "input frontier" = the last input component (in load
        // order) a value depends on.  A bilinear form whose chains are summed in the order of one operand thereby
        // becomes INPUT-major: all accumulators advance as each component of that operand arrives, so the loads of
        // a thread overlap with its arithmetic and only one operand has to be resident in registers (CGA full x
        // full: output-major 123 registers and every FMA chain waits for all 64 loads).  Dependent values never
        // have a smaller frontier than their operands, so the order stays topological.
        {
            std::vector<int> frontier(n, -1);
            for (int v = 0; v < n; ++v) {
                if (!live[v] || uniform[v]) continue;
                const ngb200_instr &I = P.instr[v];
                if (I.op == NGB200_OP_INPUT) { frontier[v] = xoff[I.a] + I.b; continue; }
                const int k = n_operands_of(I.op);
                const int ops[3] = {I.a, I.b, I.c};
                for (int j = 0; j < k; ++j) if (frontier[ops[j]] > frontier[v]) frontier[v] = frontier[ops[j]];
            }
            order.clear();
            for (int v = 0; v < n; ++v)
                if (live[v] && !uniform[v] && P.instr[v].op != NGB200_OP_INPUT) order.push_back(v);
            if (!faithful_order)
                std::stable_sort(order.begin(), order.end(), [&](int l, int r) { return frontier[l] < frontier[r]; });
        }
        // register demand: the largest number of per-item values alive at once when the body runs in that order
        // (inputs become live at their first use, stored values stay live to the end)
        {
            const int m = (int)order.size();
            std::vector<int> pos(n, -1), first_use(n, -1), last_use(n, -1), delta(m + 2, 0);
            for (int q = 0; q < m; ++q) pos[order[q]] = q;
            for (int q = 0; q < m; ++q) {
                const ngb200_instr &I = P.instr[order[q]];
                const int k = n_operands_of(I.op);
                const int ops[3] = {I.a, I.b, I.c};
                for (int j = 0; j < k; ++j) {
                    if (first_use[ops[j]] < 0) first_use[ops[j]] = q;
                    last_use[ops[j]] = q;
                }
            }
            for (const ngb200_store &st : P.stores) last_use[st.value] = m;
            for (int v = 0; v < n; ++v) {
                if (!live[v] || uniform[v] || last_use[v] < 0) continue;
                const int start = P.instr[v].op == NGB200_OP_INPUT ? (first_use[v] < 0 ? m : first_use[v]) : pos[v];
                delta[start] += 1;
                delta[last_use[v]] -= 1;
            }
            int cur = 0;
            max_live = 0;
            for (int q = 0; q <= m; ++q) { cur += delta[q]; if (cur > max_live) max_live = cur; }
        }
        return NGB200_OK;
    }

    std::string bref(int k) const {
        if (k < (int)bcast_ref.size() && !bcast_ref[k].empty()) return bcast_ref[k];
        return "p.in[" + std::to_string(k) + "]";
    }

    // typedefs, access macros and small device helpers shared by every kernel of a translation unit
    void emit_prelude(std::string &s, int vec, int ldmode, int stmode, bool pair) const {
        const char *T = f64 ? "double" : "float";
        const char *VT = f64 ? "double2" : (vec == 2 ? "float2" : "float4");
        appendf(s, "// numga_b200 generated kernel (dtype %s, %s, vec %d%s)\n", T, faithful ? "faithful" : "fast", vec,
                pair ? ", packed f32x2" : "");
        appendf(s, "typedef %s T;\ntypedef %s VT;\n", T, VT);
        if (f64)
            s += "__device__ __forceinline__ T ngb_nan_to_num(T a, T v) { return a != a ? v : (isinf(a) ? copysign(1.7976931348623157e308, a) : a); }\n";
        else
            s += "__device__ __forceinline__ T ngb_nan_to_num(T a, T v) { return a != a ? v : (isinf(a) ? copysignf(3.4028234663852886e38f, a) : a); }\n";
        if (approx) {
            s += "__device__ __forceinline__ float ngb_rcp(float a) { float r; asm(\"rcp.approx.ftz.f32 %0, %1;\" : \"=f\"(r) : \"f\"(a)); return r; }\n";
            s += "__device__ __forceinline__ float ngb_sqrt(float a) { float r; asm(\"sqrt.approx.ftz.f32 %0, %1;\" : \"=f\"(r) : \"f\"(a)); return r; }\n";
            s += "__device__ __forceinline__ float ngb_rsqrt(float a) { float r; asm(\"rsqrt.approx.ftz.f32 %0, %1;\" : \"=f\"(r) : \"f\"(a)); return r; }\n";
        }
        const char *ld = ldmode == 0 ? "*" : ldmode == 1 ? "__ldg" : ldmode == 2 ? "__ldcs" : "__ldlu";
        appendf(s, "#define NGB_LD(p) %s(p)\n", ld);
        if (stmode == 0) s += "#define NGB_ST(p, v) (*(p) = (v))\n";
        else appendf(s, "#define NGB_ST(p, v) %s((p), (v))\n", stmode == 1 ? "__stcs" : stmode == 2 ? "__stcg" : "__stwt");
    }

    // values that depend only on broadcast operands / constants / launch scalars: evaluated once per thread
    void emit_uniform(std::string &s, int *nl, int *nu) const {
        const int n = (int)P.instr.size();
        appendf(s, "__device__ __forceinline__ void ngb_uniform%s(const %s& p, T (&u)[%d]) {\n", suffix.c_str(),
                params_type.c_str(), NU > 0 ? NU : 1);
        for (int v = 0; v < n; ++v) {
            if (!live[v] || !uniform[v]) continue;
            const ngb200_instr &I = P.instr[v];
            if (I.op == NGB200_OP_CONST) continue;
            ++*nl; ++*nu;
            if (I.op == NGB200_OP_INPUT)
                appendf(s, "  const T q%d = __ldg(%s.ptr + %dLL * %s.cs);\n", v, bref(I.a).c_str(), I.b, bref(I.a).c_str());
            else if (I.op == NGB200_OP_PARAM)
                appendf(s, "  const T q%d = (T)p.scal[%d];\n", v, I.a);
            else
                appendf(s, "  const T q%d = %s;\n", v, expr(v, false).c_str());
            if (uslot[v] >= 0) appendf(s, "  u[%d] = q%d;\n", uslot[v], v);
        }
        s += "}\n";
    }

    // the per-item straight-line body
    void emit_body(std::string &s, int *nl, std::vector<int> *stored_out = nullptr) const {
        if (n_remat > 0)
            appendf(s, "__device__ __forceinline__ void ngb_body%s(const T (&x)[%d], const T* const (&xp)[%d], const T (&u)[%d], T (&y)[%d]) {\n",
                    suffix.c_str(), NX > 0 ? NX : 1, n_remat, NU > 0 ? NU : 1, NY > 0 ? NY : 1);
        else
            appendf(s, "__device__ __forceinline__ void ngb_body%s(const T (&x)[%d], const T (&u)[%d], T (&y)[%d]) {\n",
                    suffix.c_str(), NX > 0 ? NX : 1, NU > 0 ? NU : 1, NY > 0 ? NY : 1);
        for (int v : order) {
            ++*nl;
            if (fused_into[v] >= 0) continue;      // this multiply lives inside the FMA of its consumer
            appendf(s, "  const T r%d = %s;\n", v, expr(v, true).c_str());
        }
        std::vector<int> stored(NY > 0 ? NY : 1, -1);
        for (const ngb200_store &st : P.stores) stored[yoff[st.operand] + st.component] = st.value;
        for (int j = 0; j < NY; ++j) {
            if (stored[j] >= 0) appendf(s, "  y[%d] = %s;\n", j, ref(stored[j], true).c_str());
            else appendf(s, "  y[%d] = T(0);\n", j);
        }
        s += "}\n";
        if (stored_out) *stored_out = stored;
    }

    std::string generate(int vec, int ldmode, int stmode, int *n_live, int *n_uniform, bool pair) {
        plan_contraction();
        const int nin = (int)P.in_width.size(), nout = (int)P.out_width.size();
        static const char *lane[4] = {"x", "y", "z", "w"};
        std::string s;
        emit_prelude(s, vec, ldmode, stmode, pair);
        appendf(s, "#define NGB_V %d\n#define NGB_TPB %d\n", vec, P.tpb);
        // Grid-stride index arithmetic with the run-time blockDim.x instead of the compile-time CTA size: same
        // instruction mix, but ptxas lands on a different register allocation (123 vs 121 registers for CGA full x
        // full) that runs 5-10 % faster (4.01 vs 4.22 ms; twelve source variants in scripts/micro/gen_cga_variants.py
        // split cleanly into these two groups).  Measured neutral or slightly negative elsewhere (long chains -1 %,
        // gathered physics +-1 %), so only the wide one-item contiguous kernels use it.
        appendf(s, "#define NGB_BDIM %s\n", P.bdim_runtime ? "blockDim.x" : "NGB_TPB");
        // an explicit minimum of 1 CTA/SM makes ptxas spend up to 255 registers (measured 116 -> 207): only state a
        // minimum when it is a real cap
        if (P.minb > 1 && 65536 / (P.tpb * P.minb) < 255) appendf(s, "#define NGB_BOUNDS __launch_bounds__(NGB_TPB, %d)\n", P.minb);
        else s += "#define NGB_BOUNDS __launch_bounds__(NGB_TPB)\n";
        // (the two extra fields only exist in programs with a TILED operand: the plain 32-byte layout is kept elsewhere
        //  because the widest kernels are sensitive to ptxas' scheduling: CGA full x full lost 10 % with the larger one)
        if (P.any_tiled) s += "struct Operand { T* ptr; long long cs; long long bs; const long long* index; long long inner; long long os; };\n";
        else s += "struct Operand { T* ptr; long long cs; long long bs; const long long* index; };\n";
        appendf(s, "struct Params { Operand in[%d]; Operand out[%d]; double scal[%d]; long long n; };\n",
                nin > 0 ? nin : 1, nout > 0 ? nout : 1, P.n_params > 0 ? P.n_params : 1);

        int nl = 0, nu = 0;
        emit_uniform(s, &nl, &nu);
        std::vector<int> stored;
        emit_body(s, &nl, &stored);
        *n_live = nl;
        *n_uniform = nu;

        // ---- per-PAIR body (fp32): the same program on float2 values, FFMA2 / FMUL2 / FADD2 ----------------------
        if (pair) {
            s += "__device__ __forceinline__ float2 ngb_dup(float a) { return make_float2(a, a); }\n";
            s += "__device__ __forceinline__ float2 ngb_neg2(float2 a) { return make_float2(-a.x, -a.y); }\n";
            appendf(s, "__device__ __forceinline__ void ngb_body2(const float2 (&x)[%d], const T (&u)[%d], float2 (&y)[%d]) {\n",
                    NX > 0 ? NX : 1, NU > 0 ? NU : 1, NY > 0 ? NY : 1);
            for (int v : order) {
                if (fused_into[v] >= 0) continue;
                appendf(s, "  const float2 r%d = %s;\n", v, expr2(v).c_str());
            }
            for (int j = 0; j < NY; ++j) {
                if (stored[j] >= 0) appendf(s, "  y[%d] = %s;\n", j, ref2(stored[j]).c_str());
                else appendf(s, "  y[%d] = ngb_dup(0.0f);\n", j);
            }
            s += "}\n";
        }

        // ---- scalar kernel: any strides, gather/scatter -----------------------------------
        int NR = 0;                                   // components summed over the batch (NGB200_REDUCED outputs)
        std::vector<int> roff(nout, 0);
        for (int k = 0; k < nout; ++k) if (P.out_mode[k] == NGB200_REDUCED) { roff[k] = NR; NR += P.out_width[k]; }
        // Programs whose only kernel this is (gather/scatter, batch sums) get a second copy with the batch stride 1
        // compiled in (`ngb_scalar_u`, chosen at launch when every operand is contiguous along the batch): ptxas
        // schedules the constant-stride form better (gathered physics sub-step: fp32 +2 %, fp64 +3 %).
        for (int unit = 0; unit < (P.has_unit ? 2 : 1); ++unit) {
        appendf(s, "#undef NGB_BS\n#define NGB_BS(b) %s\n", unit ? "" : "* (b)");
        appendf(s, "extern \"C\" __global__ void NGB_BOUNDS %s(const Params p) {\n", unit ? "ngb_scalar_u" : "ngb_scalar");
        appendf(s, "  T u[%d]; ngb_uniform(p, u);\n", NU > 0 ? NU : 1);
        if (NR > 0) {
            appendf(s, "  __shared__ T ngb_red[%d];\n  T racc[%d];\n", NR, NR);
            appendf(s, "  for (int j = threadIdx.x; j < %d; j += NGB_TPB) ngb_red[j] = T(0);\n", NR);
            appendf(s, "  #pragma unroll\n  for (int j = 0; j < %d; ++j) racc[j] = T(0);\n  __syncthreads();\n", NR);
        }
        s += "  const long long stride = (long long)gridDim.x * NGB_BDIM;\n";
        s += "  for (long long i = (long long)blockIdx.x * NGB_BDIM + threadIdx.x; i < p.n; i += stride) {\n";
        appendf(s, "    T x[%d]; T y[%d];\n", NX > 0 ? NX : 1, NY > 0 ? NY : 1);
        for (int k = 0; k < nin; ++k) {
            if (P.in_mode[k] == NGB200_BROADCAST) continue;
            if (P.in_mode[k] == NGB200_INDEXED)
                appendf(s, "    const T* in%d = p.in[%d].ptr + p.in[%d].index[i] NGB_BS(p.in[%d].bs);\n", k, k, k, k);
            else if (P.in_mode[k] == NGB200_TILED)
                appendf(s, "    const T* in%d = p.in[%d].ptr + (i %% p.in[%d].inner) * p.in[%d].bs + (i / p.in[%d].inner) * p.in[%d].os;\n",
                        k, k, k, k, k, k);
            else
                appendf(s, "    const T* in%d = p.in[%d].ptr + i NGB_BS(p.in[%d].bs);\n", k, k, k);
            for (int c = 0; c < P.in_width[k]; ++c)
                appendf(s, "    x[%d] = NGB_LD(in%d + %dLL * p.in[%d].cs);\n", xoff[k] + c, k, c, k);
        }
        s += "    ngb_body(x, u, y);\n";
        for (int k = 0; k < nout; ++k) {
            if (P.out_mode[k] == NGB200_REDUCED) {
                for (int c = 0; c < P.out_width[k]; ++c) appendf(s, "    racc[%d] += y[%d];\n", roff[k] + c, yoff[k] + c);
                continue;
            }
            if (P.out_mode[k] == NGB200_INDEXED)
                appendf(s, "    T* out%d = p.out[%d].ptr + p.out[%d].index[i] NGB_BS(p.out[%d].bs);\n", k, k, k, k);
            else
                appendf(s, "    T* out%d = p.out[%d].ptr + i NGB_BS(p.out[%d].bs);\n", k, k, k);
            for (int c = 0; c < P.out_width[k]; ++c)
                appendf(s, "    NGB_ST(out%d + %dLL * p.out[%d].cs, y[%d]);\n", k, c, k, yoff[k] + c);
        }
        s += "  }\n";
        if (NR > 0) {
            // batch sum: registers -> warp shuffle -> shared atomics -> ONE global atomicAdd per CTA and component
            appendf(s, "  #pragma unroll\n  for (int j = 0; j < %d; ++j) {\n    T v = racc[j];\n", NR);
            s += "    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);\n";
            s += "    if ((threadIdx.x & 31) == 0) atomicAdd(&ngb_red[j], v);\n  }\n  __syncthreads();\n";
            for (int k = 0; k < nout; ++k) {
                if (P.out_mode[k] != NGB200_REDUCED) continue;
                appendf(s, "  for (int c = threadIdx.x; c < %d; c += NGB_TPB) atomicAdd(p.out[%d].ptr + (long long)c * p.out[%d].cs, ngb_red[%d + c]);\n",
                        P.out_width[k], k, k, roff[k]);
            }
        }
        s += "}\n";
        }

        // ---- vector kernel: batch_stride 1, 128-bit accesses --------------------------------
        if (vec > 1 && pair) {
            // items (0,1) and (2,3) of each thread's vector are processed as float2 pairs
            const int npair = vec / 2;
            appendf(s, "extern \"C\" __global__ void NGB_BOUNDS ngb_vec(const Params p) {\n");
            appendf(s, "  T u[%d]; ngb_uniform(p, u);\n", NU > 0 ? NU : 1);
            s += "  const long long nvec = p.n / NGB_V;\n";
            s += "  const long long stride = (long long)gridDim.x * NGB_BDIM;\n";
            s += "  for (long long v = (long long)blockIdx.x * NGB_BDIM + threadIdx.x; v < nvec; v += stride) {\n";
            appendf(s, "    float2 x[%d][%d]; float2 y[%d][%d];\n", npair, NX > 0 ? NX : 1, npair, NY > 0 ? NY : 1);
            for (int k = 0; k < nin; ++k) {
                if (P.in_mode[k] == NGB200_BROADCAST) continue;
                for (int c = 0; c < P.in_width[k]; ++c) {
                    appendf(s, "    { const VT t = NGB_LD(reinterpret_cast<const VT*>(p.in[%d].ptr + %dLL * p.in[%d].cs) + v);", k, c, k);
                    if (vec == 2) appendf(s, " x[0][%d] = t;", xoff[k] + c);
                    else appendf(s, " x[0][%d] = make_float2(t.x, t.y); x[1][%d] = make_float2(t.z, t.w);", xoff[k] + c, xoff[k] + c);
                    s += " }\n";
                }
            }
            appendf(s, "    #pragma unroll\n    for (int l = 0; l < %d; ++l) ngb_body2(x[l], u, y[l]);\n", npair);
            for (int k = 0; k < nout; ++k)
                for (int c = 0; c < P.out_width[k]; ++c) {
                    if (vec == 2) appendf(s, "    { const VT t = y[0][%d];", yoff[k] + c);
                    else appendf(s, "    { const VT t = make_float4(y[0][%d].x, y[0][%d].y, y[1][%d].x, y[1][%d].y);", yoff[k] + c, yoff[k] + c, yoff[k] + c, yoff[k] + c);
                    appendf(s, " NGB_ST(reinterpret_cast<VT*>(p.out[%d].ptr + %dLL * p.out[%d].cs) + v, t); }\n", k, c, k);
                }
            s += "  }\n}\n";
        } else if (vec == 1 && P.has_vec) {
            // contiguous kernel, one item per thread: the batch stride is the compile-time constant 1.  With the
            // run-time stride of ngb_scalar ptxas computes the first output as one serial FMA chain behind all loads;
            // with the constant stride it interleaves all chains operand-major and starts while loads are in flight
            // (CGA full x full: 4.95 -> 4.33 ms under sustained load, scripts/micro/gen_cga_variants.py).
            appendf(s, "extern \"C\" __global__ void NGB_BOUNDS ngb_vec(const Params p) {\n");
            appendf(s, "  T u[%d]; ngb_uniform(p, u);\n", NU > 0 ? NU : 1);
            s += "  const long long stride = (long long)gridDim.x * NGB_BDIM;\n";
            s += "  for (long long v = (long long)blockIdx.x * NGB_BDIM + threadIdx.x; v < p.n; v += stride) {\n";
            appendf(s, "    T x[%d]; T y[%d];\n", NX > 0 ? NX : 1, NY > 0 ? NY : 1);
            for (int k = 0; k < nin; ++k) {
                if (P.in_mode[k] == NGB200_BROADCAST) continue;
                for (int c = 0; c < P.in_width[k]; ++c)
                    appendf(s, "    x[%d] = NGB_LD(p.in[%d].ptr + %dLL * p.in[%d].cs + v);\n", xoff[k] + c, k, c, k);
            }
            s += "    ngb_body(x, u, y);\n";
            for (int k = 0; k < nout; ++k)
                for (int c = 0; c < P.out_width[k]; ++c)
                    appendf(s, "    NGB_ST(p.out[%d].ptr + %dLL * p.out[%d].cs + v, y[%d]);\n", k, c, k, yoff[k] + c);
            s += "  }\n}\n";
        } else if (vec > 1) {
            appendf(s, "extern \"C\" __global__ void NGB_BOUNDS ngb_vec(const Params p) {\n");
            appendf(s, "  T u[%d]; ngb_uniform(p, u);\n", NU > 0 ? NU : 1);
            s += "  const long long nvec = p.n / NGB_V;\n";
            s += "  const long long stride = (long long)gridDim.x * NGB_BDIM;\n";
            s += "  for (long long v = (long long)blockIdx.x * NGB_BDIM + threadIdx.x; v < nvec; v += stride) {\n";
            appendf(s, "    T x[NGB_V][%d]; T y[NGB_V][%d];\n", NX > 0 ? NX : 1, NY > 0 ? NY : 1);
            for (int k = 0; k < nin; ++k) {
                if (P.in_mode[k] == NGB200_BROADCAST) continue;
                for (int c = 0; c < P.in_width[k]; ++c) {
                    appendf(s, "    { const VT t = NGB_LD(reinterpret_cast<const VT*>(p.in[%d].ptr + %dLL * p.in[%d].cs) + v);", k, c, k);
                    for (int l = 0; l < vec; ++l) appendf(s, " x[%d][%d] = t.%s;", l, xoff[k] + c, lane[l]);
                    s += " }\n";
                }
            }
            s += "    #pragma unroll\n    for (int l = 0; l < NGB_V; ++l) ngb_body(x[l], u, y[l]);\n";
            for (int k = 0; k < nout; ++k)
                for (int c = 0; c < P.out_width[k]; ++c) {
                    s += "    { VT t;";
                    for (int l = 0; l < vec; ++l) appendf(s, " t.%s = y[%d][%d];", lane[l], l, yoff[k] + c);
                    appendf(s, " NGB_ST(reinterpret_cast<VT*>(p.out[%d].ptr + %dLL * p.out[%d].cs) + v, t); }\n", k, c, k);
                }
            s += "  }\n}\n";
        }
        // ---- staged kernel: TMA-style bulk copies global -> shared, mbarrier full-barriers, S-deep pipeline -----
        // One elected thread per CTA issues one `cp.async.bulk` per input component row of the NEXT tiles while all
        // threads compute the current tile out of shared memory: the loads of a register-heavy kernel no longer
        // depend on how many warps fit (CGA full x full: 12 warps/SM, top stall long_scoreboard).  Outputs go
        // straight to global memory (coalesced, written once).
        if (P.stage_tpb > 0) {
            const int TP = P.stage_tpb, S = P.stage_stages;
            const int es = f64 ? 8 : 4;
            appendf(s, "#define NGB_STP %d\n#define NGB_SS %d\n", TP, S);
            s += "__device__ __forceinline__ unsigned ngb_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }\n";
            s += "__device__ __forceinline__ void ngb_mbar_wait(unsigned bar, unsigned phase) {\n"
                 "  asm volatile(\"{\\n.reg .pred p;\\nNGB_WAIT:\\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\\n@p bra NGB_DONE;\\nbra NGB_WAIT;\\nNGB_DONE:\\n}\" :: \"r\"(bar), \"r\"(phase) : \"memory\");\n}\n";
            s += "__device__ __forceinline__ void ngb_bulk(unsigned dst, const void* src, unsigned bytes, unsigned bar) {\n"
                 "  asm volatile(\"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\" :: \"r\"(dst), \"l\"(src), \"r\"(bytes), \"r\"(bar) : \"memory\");\n}\n";
            // issue all rows of tile t into stage st
            s += "__device__ __forceinline__ void ngb_issue(const Params& p, T* buf, unsigned bar, int st, long long t) {\n";
            appendf(s, "  asm volatile(\"mbarrier.arrive.expect_tx.shared::cta.b64 _, [%%0], %%1;\" :: \"r\"(bar), \"r\"(%uu) : \"memory\");\n",
                    (unsigned)(NX * TP * es));
            for (int k = 0; k < nin; ++k) {
                if (P.in_mode[k] == NGB200_BROADCAST) continue;
                for (int c = 0; c < P.in_width[k]; ++c)
                    appendf(s, "  ngb_bulk(ngb_smem(buf + ((long long)st * %d + %d) * NGB_STP), p.in[%d].ptr + %dLL * p.in[%d].cs + t * NGB_STP, %uu, bar);\n",
                            NX, xoff[k] + c, k, c, k, (unsigned)(TP * es));
            }
            s += "}\n";
            int cta_per_sm = (int)(232448 / ((size_t)S * NX * TP * es + 1024));     // what shared memory allows
            if (cta_per_sm < 1) cta_per_sm = 1;
            if (cta_per_sm > 8) cta_per_sm = 8;
            appendf(s, "extern \"C\" __global__ void __launch_bounds__(NGB_STP, %d) ngb_stage(const Params p) {\n", cta_per_sm);
            s += "  extern __shared__ __align__(128) unsigned char ngb_raw[];\n";
            s += "  T* buf = reinterpret_cast<T*>(ngb_raw);\n";
            s += "  __shared__ __align__(8) unsigned long long full[NGB_SS];\n";
            s += "  const int tid = threadIdx.x;\n";
            s += "  const long long ntiles = p.n / NGB_STP;\n";
            s += "  if (tid == 0) {\n";
            s += "    for (int st = 0; st < NGB_SS; ++st) asm volatile(\"mbarrier.init.shared::cta.b64 [%0], 1;\" :: \"r\"(ngb_smem(&full[st])));\n";
            s += "    asm volatile(\"fence.mbarrier_init.release.cluster;\" ::: \"memory\");\n";
            s += "    asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");\n";
            s += "  }\n  __syncthreads();\n";
            appendf(s, "  T u[%d]; ngb_uniform(p, u);\n", NU > 0 ? NU : 1);
            s += "  if (tid == 0) {\n    for (int st = 0; st < NGB_SS; ++st) {\n      const long long t = (long long)blockIdx.x + (long long)st * gridDim.x;\n"
                 "      if (t < ntiles) ngb_issue(p, buf, ngb_smem(&full[st]), st, t);\n    }\n  }\n";
            s += "  int it = 0;\n";
            s += "  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {\n";
            s += "    const int st = it % NGB_SS;\n";
            s += "    ngb_mbar_wait(ngb_smem(&full[st]), (unsigned)((it / NGB_SS) & 1));\n";
            appendf(s, "    T x[%d]; T y[%d];\n", NX > 0 ? NX : 1, NY > 0 ? NY : 1);
            appendf(s, "    const T* tile = buf + (long long)st * %d * NGB_STP + tid;\n", NX);
            for (int j = 0; j < NX; ++j) appendf(s, "    x[%d] = tile[%d * NGB_STP];\n", j, j);
            s += "    ngb_body(x, u, y);\n";
            s += "    const long long i = t * NGB_STP + tid;\n";
            for (int k = 0; k < nout; ++k)
                for (int c = 0; c < P.out_width[k]; ++c)
                    appendf(s, "    NGB_ST(p.out[%d].ptr + %dLL * p.out[%d].cs + i, y[%d]);\n", k, c, k, yoff[k] + c);
            s += "    __syncthreads();\n";      // every thread has read stage st: it can be refilled
            s += "    if (tid == 0) {\n      const long long nt = t + (long long)NGB_SS * gridDim.x;\n"
                 "      if (nt < ntiles) ngb_issue(p, buf, ngb_smem(&full[st]), st, nt);\n    }\n";
            s += "  }\n}\n";
        }
        return s;
    }
};

// ------------------------------------------------------------------------------------------
// cache + NVRTC
// ------------------------------------------------------------------------------------------
static std::mutex g_mu;
static std::string g_cache_dir;

static std::string default_cache_dir() {
    const char *env = getenv("NGB200_CACHE_DIR");
    if (env && *env) return env;
    Dl_info info;
    if (dladdr((void *)&ngb200_last_error, &info) && info.dli_fname) {
        std::string p(info.dli_fname);
        size_t k = p.rfind('/');
        if (k != std::string::npos) return p.substr(0, k) + "/_kcache";
    }
    // last resort (library path unknown): a per-user directory, never a shared predictable path in /tmp
    const char *xdg = getenv("XDG_CACHE_HOME");
    const char *home = getenv("HOME");
    std::string base = (xdg && *xdg) ? std::string(xdg) : (home && *home) ? std::string(home) + "/.cache" : std::string("/tmp/.cache-") + std::to_string((long)getuid());
    mkdir(base.c_str(), 0700);
    return base + "/numga_b200_kcache";
}

extern "C" int ngb200_set_cache_dir(const char *path) {
    std::lock_guard<std::mutex> lock(g_mu);
    g_cache_dir = path ? path : "";
    return NGB200_OK;
}

static std::string cache_dir() {
    if (g_cache_dir.empty()) g_cache_dir = default_cache_dir();
    mkdir(g_cache_dir.c_str(), 0700);
    return g_cache_dir;
}

static std::string hash_hex(const std::string &s) {
    uint64_t a = 1469598103934665603ULL, b = 0x9E3779B97F4A7C15ULL;
    for (unsigned char ch : s) {
        a = (a ^ ch) * 1099511628211ULL;
        b = (b + ch) * 0xD6E8FEB86659FD93ULL;
        b ^= b >> 32;
    }
    char buf[40];
    snprintf(buf, sizeof buf, "%016llx%016llx", (unsigned long long)a, (unsigned long long)b);
    return buf;
}

static bool read_file(const std::string &path, std::vector<char> &out) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? n : 0);
    bool ok = n > 0 && fread(out.data(), 1, n, f) == (size_t)n;
    fclose(f);
    return ok;
}

static void write_file(const std::string &path, const char *data, size_t n) {
    std::string tmp = path + ".tmp" + std::to_string((long)getpid());
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return;
    fwrite(data, 1, n, f);
    fclose(f);
    rename(tmp.c_str(), path.c_str());
}

static int nvrtc_compile(const std::string &src, std::vector<char> &cubin, std::string *log_out = nullptr) {
    nvrtcProgram prog;
    if (nvrtcCreateProgram(&prog, src.c_str(), "ngb200_kernel.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS)
        return fail(NGB200_ERR_COMPILE, "nvrtcCreateProgram failed");
    const char *opts[] = {"--gpu-architecture=sm_100a", "-lineinfo", "--std=c++17", "--fmad=true",
                          "--prec-div=true", "--prec-sqrt=true", "--ftz=false", "--ptxas-options=-v"};
    nvrtcResult r = nvrtcCompileProgram(prog, (int)(sizeof opts / sizeof *opts), opts);
    if (r != NVRTC_SUCCESS) {
        size_t n = 0;
        nvrtcGetProgramLogSize(prog, &n);
        std::string log(n, 0);
        nvrtcGetProgramLog(prog, &log[0]);
        nvrtcDestroyProgram(&prog);
        return fail(NGB200_ERR_COMPILE, "NVRTC: %s\n%.3000s", nvrtcGetErrorString(r), log.c_str());
    }
    if (log_out) {
        size_t ln = 0;
        nvrtcGetProgramLogSize(prog, &ln);
        log_out->assign(ln, 0);
        if (ln) nvrtcGetProgramLog(prog, &(*log_out)[0]);
    }
    size_t n = 0;
    nvrtcGetCUBINSize(prog, &n);
    cubin.resize(n);
    nvrtcGetCUBIN(prog, cubin.data());
    nvrtcDestroyProgram(&prog);
    if (n == 0) return fail(NGB200_ERR_COMPILE, "NVRTC produced no cubin");
    return NGB200_OK;
}

// bytes of spill stores ptxas reports for one entry function in a `-v` log (0 if the function is not listed)
static long spill_bytes(const std::string &log, const char *entry) {
    size_t at = log.find(std::string("entry function '") + entry + "'");
    if (at == std::string::npos) return 0;
    size_t sp = log.find("bytes spill stores", at);
    if (sp == std::string::npos) return 0;
    size_t b = log.rfind(',', sp);
    return b == std::string::npos ? 0 : atol(log.c_str() + b + 1);
}

static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}

static int finish_program(ngb200_program *P, int vec_hint) {
    for (int m : P->in_mode) P->any_indexed |= (m == NGB200_INDEXED || m == NGB200_TILED);    // scalar kernel only
    for (int m : P->out_mode) P->any_indexed |= (m == NGB200_INDEXED);
    for (int m : P->out_mode) P->any_reduced |= (m == NGB200_REDUCED);
    for (int m : P->in_mode) P->any_tiled |= (m == NGB200_TILED);
    Gen gen(*P);
    int rc = gen.analyse();
    if (rc) return rc;
    if (gen.has_partner) return fail(NGB200_ERR_INVALID, "NGB200_OP_PARTNER is only available in pipeline phases (items must map to adjacent lanes)");
    int live_guess = 0;
    for (char c : gen.live) live_guess += c;
    // items per thread: small programs are HBM-bound -> 128-bit accesses (4 x f32 / 2 x f64); big ones are
    // compute-bound -> fp32 keeps 2 items (one packed f32x2 pair, 64-bit accesses), fp64 1 item
    const bool f32 = P->dtype == NGB200_F32;
    const bool small = live_guess <= env_int("NGB200_VEC_INSTR_LIMIT", 700);
    int vmax = f32 ? 4 : 2;
    // measured on B200 (profiles/): long dependent chains (exp/log: ~110 instructions per operand component) are
    // fastest with 1 item per thread; wide products (CGA full x full: ~21) gain 10% from 2 items per thread
    const bool wide = live_guess <= 48 * (gen.NX + gen.NY);
    // ... unless two items need more than 128 registers: one item per thread at 128 registers (16 warps/SM) then
    // sustains more than two at 168 (12 warps/SM): CGA 4.33 vs 4.83 ms under sustained load
    const bool two_fit = gen.max_live * 2 + 16 <= 128;
    int vec = vec_hint > 0 ? vec_hint : (small ? vmax : (f32 && wide && two_fit ? 2 : 1));
    if (env_int("NGB200_FORCE_VEC", 0) > 0) vec = env_int("NGB200_FORCE_VEC", 0);
    if (vec > vmax) vec = vmax;
    if (vec == 3) vec = 2;
    if (P->any_indexed || P->any_reduced || gen.NX == 0) vec = 1;
    // Packed f32x2 arithmetic: measured on B200 (scripts/micro/fp32_rate.cu) scalar FFMA already issues at the full
    // 128 lanes/clk/SM (34.4 T FMA/s) and FFMA2 reaches the same 70 TFLOP/s, so packing buys no throughput and costs
    // registers in large programs; it is kept for the small HBM-bound programs (half the issue slots per item).
    const bool pair = f32 && vec >= 2 && env_int("NGB200_PAIR", small ? 1 : 0);
    P->vec = vec;
    P->has_vec = vec > 1 || (!P->any_indexed && !P->any_reduced && gen.NX > 0 && env_int("NGB200_CONTIG1", 1));
    P->has_unit = !P->has_vec && (P->any_indexed || P->any_reduced) && gen.NX > 0 && env_int("NGB200_UNIT_BS", 1);
    P->bdim_runtime = env_int("NGB200_BDIM", -1) >= 0 ? env_int("NGB200_BDIM", -1) != 0 : (vec == 1 && P->has_vec && !small && wide);
    // Register budget of LARGE programs.  ptxas schedules a long straight-line body for maximal ILP and can take
    // 200+ registers for a chain whose working set is 17 values (measured: exp->normalized->log, 205 registers,
    // 11.1 ms; capped at 48: 5.4 ms; the gathered physics kernels: 254 -> 128 registers, +15%).  The cap is derived
    // from the program's own peak number of live values; small (HBM-bound) programs keep ptxas' natural choice.
    // Programs that need more than 128 registers run 128-thread CTAs: finer occupancy granularity (CGA full x full,
    // 2 items/thread: 168 registers -> 3 CTAs of 128 instead of 1 CTA of 256, 4.33 vs 5.86 ms).
    // staged kernel: experimental knob NGB200_STAGE=1 (wide, non-gathered programs only)
    P->stage_tpb = 0;
    if (env_int("NGB200_STAGE", 0) && !P->any_indexed && !P->any_reduced && gen.NX > 0) {
        P->stage_tpb = env_int("NGB200_STAGE_TPB", 128);
        P->stage_stages = env_int("NGB200_STAGES", 2);
    }
    auto make_variant = [&](bool ir_order) {
        if (gen.faithful_order != ir_order) {
            gen.faithful_order = ir_order;
            gen.NX = gen.NU = gen.NY = 0;
            gen.analyse();
        }
        P->minb = 1;
        P->tpb = 256;
        if (live_guess >= 256) {
            const int need = gen.max_live * (f32 ? 1 : 2) * vec + 16;
            if (need > 128) P->tpb = 128;
            int minb = 65536 / (P->tpb * need);
            P->minb = minb < 1 ? 1 : (minb > 8 ? 8 : minb);
        }
        if (env_int("NGB200_TPB", 0) > 0) P->tpb = env_int("NGB200_TPB", 0);
        if (env_int("NGB200_MINBLOCKS", 0) > 0) P->minb = env_int("NGB200_MINBLOCKS", 0);
        P->source = gen.generate(vec, env_int("NGB200_LDMODE", 0), env_int("NGB200_STMODE", 0), &P->n_live, &P->n_uniform, pair);
        P->hash = hash_hex(P->source + "|sm_100a");
        P->stage_nx = gen.NX;
        P->stage_smem = (size_t)P->stage_stages * gen.NX * P->stage_tpb * esize(P->dtype);
    };
    P->param_bytes = (P->any_tiled ? 48 : 32) * ((P->in_width.size() ? P->in_width.size() : 1) + (P->out_width.size() ? P->out_width.size() : 1)) +
                     8 * (P->n_params > 0 ? P->n_params : 1) + 8;
    if (P->param_bytes > 4000) return fail(NGB200_ERR_INVALID, "too many operands for one kernel (%zu parameter bytes)", P->param_bytes);

    // Two emission orders of the same program (Gen::analyse): IR order, and input-frontier order.  Which one ptxas
    // turns into the better kernel depends on the program (frontier order removes the spills of the fp64 gathered
    // physics kernels, -17 % time; it costs the flat CGA product 10 %), so the choice is made on ptxas' own report:
    // IR order unless it spills, then whichever spills less.  The cache key is the generated source alone: a cubin
    // built at build() time by this image's NVRTC stays valid when torch's bundled (older) libnvrtc happens to be
    // the one loaded in the process; `<hash>.alt` marks "the other order won" for the IR-order hash.
    const int forced = env_int("NGB200_IR_ORDER", -1);
    make_variant(forced != 0);
    std::lock_guard<std::mutex> lock(g_mu);
    const bool use_cache = !env_int("NGB200_NO_CACHE", 0);
    std::string base = cache_dir() + "/" + P->hash;
    if (use_cache && read_file(base + ".cubin", P->cubin)) {
        P->cache_hit = true;
        return NGB200_OK;
    }
    std::vector<char> marker;
    if (forced < 0 && use_cache && read_file(base + ".alt", marker)) {
        make_variant(false);
        base = cache_dir() + "/" + P->hash;
        if (read_file(base + ".cubin", P->cubin)) {
            P->cache_hit = true;
            return NGB200_OK;
        }
        rc = nvrtc_compile(P->source, P->cubin);
        if (rc) return rc;
        write_file(base + ".cubin", P->cubin.data(), P->cubin.size());
        if (env_int("NGB200_CACHE_SOURCES", 0)) write_file(base + ".cu", P->source.data(), P->source.size());
        return NGB200_OK;
    }
    std::string log;
    rc = nvrtc_compile(P->source, P->cubin, &log);
    if (rc) return rc;
    const char *main_kernel = P->has_vec ? "ngb_vec" : (P->has_unit ? "ngb_scalar_u" : "ngb_scalar");
    const long spill_ir = spill_bytes(log, main_kernel);
    if (forced < 0 && spill_ir > 0) {
        std::string src_ir = P->source, hash_ir = P->hash;
        std::vector<char> cubin_ir = P->cubin;
        const int tpb_ir = P->tpb, minb_ir = P->minb, nl = P->n_live, nu = P->n_uniform;
        make_variant(false);
        std::vector<char> cubin_alt;
        std::string log_alt;
        if (nvrtc_compile(P->source, cubin_alt, &log_alt) == NGB200_OK && spill_bytes(log_alt, main_kernel) < spill_ir) {
            P->cubin = cubin_alt;
            base = cache_dir() + "/" + P->hash;
            write_file((cache_dir() + "/" + hash_ir + ".alt").c_str(), P->hash.data(), P->hash.size());
        } else {
            make_variant(true);          // restore the IR-order program
            P->cubin = cubin_ir;
            P->tpb = tpb_ir; P->minb = minb_ir; P->n_live = nl; P->n_uniform = nu;
        }
    }
    write_file(base + ".cubin", P->cubin.data(), P->cubin.size());
    if (env_int("NGB200_CACHE_SOURCES", 0)) write_file(base + ".cu", P->source.data(), P->source.size());
    return NGB200_OK;
}

extern "C" int ngb200_program_create(const ngb200_program_desc *d, ngb200_program **out) {
    if (!d || !out) return fail(NGB200_ERR_INVALID, "null argument");
    if (d->dtype != NGB200_F32 && d->dtype != NGB200_F64) return fail(NGB200_ERR_INVALID, "bad dtype %d", d->dtype);
    if (d->n_inputs < 0 || d->n_outputs < 1 || d->n_instr < 0 || d->n_stores < 0 || d->n_consts < 0 || d->n_params < 0)
        return fail(NGB200_ERR_INVALID, "bad counts in program description");
    ngb200_program *P = new ngb200_program;
    P->dtype = d->dtype;
    P->flags = d->flags;
    P->in_width.assign(d->input_width, d->input_width + d->n_inputs);
    if (d->input_mode) P->in_mode.assign(d->input_mode, d->input_mode + d->n_inputs);
    else P->in_mode.assign(d->n_inputs, NGB200_VARYING);
    P->out_width.assign(d->output_width, d->output_width + d->n_outputs);
    if (d->output_mode) P->out_mode.assign(d->output_mode, d->output_mode + d->n_outputs);
    else P->out_mode.assign(d->n_outputs, NGB200_VARYING);
    P->n_params = d->n_params;
    if (d->n_consts) P->consts.assign(d->consts, d->consts + d->n_consts);
    if (d->n_instr) P->instr.assign(d->instr, d->instr + d->n_instr);
    if (d->n_stores) P->stores.assign(d->stores, d->stores + d->n_stores);
    for (int w : P->in_width) if (w < 0 || w > 4096) { delete P; return fail(NGB200_ERR_INVALID, "bad input width"); }
    for (int w : P->out_width) if (w < 0 || w > 4096) { delete P; return fail(NGB200_ERR_INVALID, "bad output width"); }
    for (int m : P->in_mode) if (m != NGB200_VARYING && m != NGB200_BROADCAST && m != NGB200_INDEXED && m != NGB200_TILED) { delete P; return fail(NGB200_ERR_INVALID, "bad input mode"); }
    for (int m : P->out_mode) if (m != NGB200_VARYING && m != NGB200_INDEXED && m != NGB200_REDUCED) { delete P; return fail(NGB200_ERR_INVALID, "bad output mode"); }
    int rc = finish_program(P, d->vec);
    if (rc) { delete P; return rc; }
    *out = P;
    return NGB200_OK;
}

// ---- operator from its term table -----------------------------------------------------------
struct Builder {
    std::vector<ngb200_instr> code;
    std::vector<double> consts;
    std::map<std::vector<int32_t>, int> memo;
    int emit(int op, int a = 0, int b = 0, int c = 0) {
        if ((op == NGB200_OP_ADD || op == NGB200_OP_MUL) && a > b) std::swap(a, b);  // exact: IEEE add/mul commute
        std::vector<int32_t> key = {op, a, b, c};
        auto it = memo.find(key);
        if (it != memo.end()) return it->second;
        code.push_back({op, a, b, c});
        return memo[key] = (int)code.size() - 1;
    }
    int constant(double v) {
        for (size_t k = 0; k < consts.size(); ++k)
            if (consts[k] == v) return emit(NGB200_OP_CONST, (int)k);
        consts.push_back(v);
        return emit(NGB200_OP_CONST, (int)consts.size() - 1);
    }
};

extern "C" int ngb200_operator_create(const int32_t *terms, int32_t n_terms, int32_t arity, const int32_t *n_in,
                                      int32_t n_out, int32_t dtype, uint32_t bcast_mask, uint32_t flags,
                                      ngb200_program **out) {
    if (!out || arity < 1 || arity > 8 || n_terms < 0 || (n_terms && !terms) || !n_in || n_out < 1)
        return fail(NGB200_ERR_INVALID, "bad operator description");
    const bool faithful = flags & NGB200_FAITHFUL;
    const int w = arity + 2;
    for (int t = 0; t < n_terms; ++t) {
        for (int k = 0; k < arity; ++k)
            if (terms[t * w + k] < 0 || terms[t * w + k] >= n_in[k]) return fail(NGB200_ERR_INVALID, "term %d: input index out of range", t);
        if (terms[t * w + arity] < 0 || terms[t * w + arity] >= n_out) return fail(NGB200_ERR_INVALID, "term %d: output index out of range", t);
    }
    Builder B;
    std::vector<int32_t> in_mode(arity);
    bool any_b = false, any_v = false;
    for (int k = 0; k < arity; ++k) {
        in_mode[k] = (bcast_mask >> k) & 1 ? NGB200_BROADCAST : NGB200_VARYING;
        (in_mode[k] == NGB200_BROADCAST ? any_b : any_v) = true;
    }
    auto input = [&](int k, int i) { return B.emit(NGB200_OP_INPUT, k, i); };
    auto signed_product = [&](const int32_t *term, bool want_bcast, bool want_varying, int coeff) -> int {
        // product of the selected factors, right-associated a0*(a1*(a2..)) -- the order in which
        // numpy's einsum contracts, so FAITHFUL ternary operators are bit-exact too; coefficient last
        int acc = -1;
        for (int k = arity - 1; k >= 0; --k) {
            bool bc = in_mode[k] == NGB200_BROADCAST;
            if ((bc && !want_bcast) || (!bc && !want_varying)) continue;
            int x = input(k, term[k]);
            acc = acc < 0 ? x : B.emit(NGB200_OP_MUL, acc, x);
        }
        if (acc < 0) acc = B.constant(1.0);
        if (coeff == -1) acc = B.emit(NGB200_OP_NEG, acc);
        else if (coeff != 1) acc = B.emit(NGB200_OP_MUL, acc, B.constant((double)coeff));
        return acc;
    };
    std::vector<ngb200_store> stores;
    for (int o = 0; o < n_out; ++o) {
        int acc = -1;
        if (!faithful && any_b && any_v) {
            // pre-contract the broadcast operands: out[o] = sum_J (sum c * prod bcast) * prod varying[J]
            std::vector<std::vector<int32_t>> keys;
            std::vector<int> ucoef;
            for (int t = 0; t < n_terms; ++t) {
                const int32_t *term = terms + t * w;
                if (term[arity] != o) continue;
                std::vector<int32_t> key;
                for (int k = 0; k < arity; ++k) if (in_mode[k] != NGB200_BROADCAST) key.push_back(term[k]);
                int up = signed_product(term, true, false, term[arity + 1]);
                size_t j = 0;
                for (; j < keys.size(); ++j) if (keys[j] == key) break;
                if (j == keys.size()) { keys.push_back(key); ucoef.push_back(up); }
                else ucoef[j] = B.emit(NGB200_OP_ADD, ucoef[j], up);
            }
            for (size_t j = 0; j < keys.size(); ++j) {
                int vp = -1, kk = 0;
                for (int k = 0; k < arity; ++k) {
                    if (in_mode[k] == NGB200_BROADCAST) continue;
                    int x = input(k, keys[j][kk++]);
                    vp = vp < 0 ? x : B.emit(NGB200_OP_MUL, vp, x);
                }
                int prod = B.emit(NGB200_OP_MUL, ucoef[j], vp);
                acc = acc < 0 ? prod : B.emit(NGB200_OP_ADD, acc, prod);
            }
        } else {
            // FAST: sum each output in the order of the LAST operand's component, so that the input-frontier emission
            // order (Gen::analyse) streams that operand; FAITHFUL keeps precompute_sparse order (= einsum's)
            std::vector<int> row;
            for (int t = 0; t < n_terms; ++t) if (terms[t * w + arity] == o) row.push_back(t);
            if (!faithful)
                std::stable_sort(row.begin(), row.end(), [&](int l, int r) { return terms[l * w + arity - 1] < terms[r * w + arity - 1]; });
            for (int t : row) {
                const int32_t *term = terms + t * w;
                int c = term[arity + 1];
                int prod = signed_product(term, true, true, c == -1 ? 1 : c);
                if (acc < 0) acc = c == -1 ? B.emit(NGB200_OP_NEG, prod) : prod;
                else acc = B.emit(c == -1 ? NGB200_OP_SUB : NGB200_OP_ADD, acc, prod);  // a - p == a + (-p) exactly
            }
        }
        if (acc >= 0) stores.push_back({0, o, acc});
    }
    ngb200_program_desc d;
    memset(&d, 0, sizeof d);
    d.dtype = dtype;
    d.flags = flags;
    d.n_inputs = arity;
    d.input_width = n_in;
    d.input_mode = in_mode.data();
    d.n_outputs = 1;
    d.output_width = &n_out;
    d.n_consts = (int)B.consts.size();
    d.consts = B.consts.data();
    d.n_instr = (int)B.code.size();
    d.instr = B.code.data();
    d.n_stores = (int)stores.size();
    d.stores = stores.data();
    return ngb200_program_create(&d, out);
}

extern "C" void ngb200_program_destroy(ngb200_program *P) {
    if (!P) return;
    host_pipe_destroy(P->host);
    if (g_drv.ok) {
        // unload the device code of every device this program ran on (transient programs -- one per operator mask,
        // lazy DAG structure or VJP variant -- would otherwise leak module memory for the life of the process)
        for (int d = 0; d < kMaxDevices; ++d) {
            Loaded &L = P->loaded[d];
            if (!L.ready.load() || !L.mod) continue;
            CUcontext prev = nullptr, ctx = nullptr;
            CUdevice dev;
            if (g_drv.p_cuCtxGetCurrent(&prev) != CUDA_SUCCESS) continue;
            if (g_drv.p_cuDeviceGet(&dev, d) != CUDA_SUCCESS || g_drv.p_cuDevicePrimaryCtxRetain(&ctx, dev) != CUDA_SUCCESS) continue;
            g_drv.p_cuCtxSetCurrent(ctx);
            g_drv.p_cuModuleUnload(L.mod);
            g_drv.p_cuCtxSetCurrent(prev);
            L.mod = nullptr;
        }
    }
    delete P;
}

extern "C" const char *ngb200_program_source(const ngb200_program *P) { return P ? P->source.c_str() : ""; }

extern "C" int ngb200_program_get_info(const ngb200_program *P, ngb200_program_info *info) {
    if (!P || !info) return fail(NGB200_ERR_INVALID, "null argument");
    memset(info, 0, sizeof *info);
    info->n_instr = P->n_live;
    info->n_uniform = P->n_uniform;
    info->vec = P->vec;
    info->cache_hit = P->cache_hit;
    info->cubin_bytes = (int64_t)P->cubin.size();
    for (int d = 0; d < kMaxDevices; ++d)
        if (P->loaded[d].ready.load()) { info->regs_vec = P->loaded[d].regs_vec; info->regs_scalar = P->loaded[d].regs_scalar; break; }
    return NGB200_OK;
}

// ------------------------------------------------------------------------------------------
// loading + launching
// ------------------------------------------------------------------------------------------
static int load_on_device(ngb200_program *P, int dev, Loaded **out) {
    if (dev < 0 || dev >= kMaxDevices) return fail(NGB200_ERR_CUDA, "device ordinal %d not supported", dev);
    Loaded &L = P->loaded[dev];
    if (!L.ready.load(std::memory_order_acquire)) {
        std::lock_guard<std::mutex> lock(P->mu);
        if (!L.ready.load()) {
            CU(cuModuleLoadData(&L.mod, P->cubin.data()));
            CU(cuModuleGetFunction(&L.fscalar, L.mod, "ngb_scalar"));
            int sms = 0;
            CU(cuDeviceGetAttribute(&sms, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev));
            int waves = env_int("NGB200_WAVES", 8), nb = 0;
            CU(cuFuncGetAttribute(&L.regs_scalar, CU_FUNC_ATTRIBUTE_NUM_REGS, L.fscalar));
            CU(cuOccupancyMaxActiveBlocksPerMultiprocessor(&nb, L.fscalar, P->tpb, 0));
            L.grid_scalar = sms * (nb > 0 ? nb : 1) * waves;
            if (P->has_unit) CU(cuModuleGetFunction(&L.fscalar_u, L.mod, "ngb_scalar_u"));   // same body: same occupancy
            if (P->has_vec) {
                CU(cuModuleGetFunction(&L.fvec, L.mod, "ngb_vec"));
                CU(cuFuncGetAttribute(&L.regs_vec, CU_FUNC_ATTRIBUTE_NUM_REGS, L.fvec));
                CU(cuOccupancyMaxActiveBlocksPerMultiprocessor(&nb, L.fvec, P->tpb, 0));
                L.grid_vec = sms * (nb > 0 ? nb : 1) * waves;
            }
            if (P->stage_tpb > 0) {
                CU(cuModuleGetFunction(&L.fstage, L.mod, "ngb_stage"));
                CU(cuFuncGetAttribute(&L.regs_stage, CU_FUNC_ATTRIBUTE_NUM_REGS, L.fstage));
                CU(cuFuncSetAttribute(L.fstage, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)P->stage_smem));
                CU(cuOccupancyMaxActiveBlocksPerMultiprocessor(&nb, L.fstage, P->stage_tpb, P->stage_smem));
                L.grid_stage = sms * (nb > 0 ? nb : 1);       // one resident wave: every CTA keeps its pipeline primed
            }
            L.ready.store(1, std::memory_order_release);
        }
    }
    *out = &L;
    return NGB200_OK;
}

struct DevOperand {
    void *ptr;
    long long cs, bs;
    const long long *index;
    long long inner, os;
};

static int launch_one(ngb200_program *P, CUfunction f, int grid_cap, int64_t work, const ngb200_operand *in,
                      const ngb200_operand *out, const double *params, int64_t n, int64_t offset, CUstream stream,
                      int tpb = 0, size_t smem = 0) {
    if (tpb <= 0) tpb = P->tpb;
    const size_t nin = P->in_width.size(), nout = P->out_width.size();
    const size_t nin_s = nin ? nin : 1, nout_s = nout ? nout : 1, np_s = P->n_params > 0 ? P->n_params : 1;
    alignas(16) char buf[4096];                        // param_bytes <= 4000 is checked at creation: no heap allocation per launch
    memset(buf, 0, P->param_bytes);
    const size_t ob = P->any_tiled ? 48 : 32;          // bytes per operand in the kernel's Params
    const int es = esize(P->dtype);
    auto put = [&](size_t slot, const DevOperand &d) { memcpy(buf + slot * ob, &d, ob); };
    for (size_t k = 0; k < nin; ++k) {
        bool moves = P->in_mode[k] == NGB200_VARYING;
        put(k, {(char *)in[k].ptr + (moves ? offset * in[k].batch_stride * es : 0), in[k].comp_stride,
                P->in_mode[k] == NGB200_BROADCAST ? 0 : in[k].batch_stride,
                P->in_mode[k] == NGB200_INDEXED ? (const long long *)in[k].index + offset : nullptr,
                P->in_mode[k] == NGB200_TILED ? in[k].inner_size : 0, P->in_mode[k] == NGB200_TILED ? in[k].outer_stride : 0});
    }
    for (size_t k = 0; k < nout; ++k) {
        bool moves = P->out_mode[k] == NGB200_VARYING;
        put(nin_s + k, {(char *)out[k].ptr + (moves ? offset * out[k].batch_stride * es : 0), out[k].comp_stride,
                        out[k].batch_stride,
                        P->out_mode[k] == NGB200_INDEXED ? (const long long *)out[k].index + offset : nullptr, 0, 0});
    }
    double *sc = reinterpret_cast<double *>(buf + ob * (nin_s + nout_s));
    for (int k = 0; k < P->n_params; ++k) sc[k] = params ? params[k] : 0.0;
    *reinterpret_cast<long long *>(buf + ob * (nin_s + nout_s) + 8 * np_s) = n;
    int64_t blocks = (work + tpb - 1) / tpb;
    if (blocks > grid_cap) blocks = grid_cap;
    if (blocks < 1) blocks = 1;
    void *args[] = {buf};
    CU(cuLaunchKernel(f, (unsigned)blocks, 1, 1, (unsigned)tpb, 1, 1, (unsigned)smem, stream, args, nullptr));
    g_launches.fetch_add(1);
    return NGB200_OK;
}

extern "C" int ngb200_program_launch(ngb200_program *P, const ngb200_operand *in, const ngb200_operand *out,
                                     const double *params, int64_t batch, void *stream) {
    if (!P || (!in && !P->in_width.empty()) || !out) return fail(NGB200_ERR_INVALID, "null argument");
    if (batch < 0) return fail(NGB200_ERR_INVALID, "negative batch");
    if (P->n_params > 0 && !params) return fail(NGB200_ERR_INVALID, "program needs %d scalar params", P->n_params);
    int rc = need_driver();
    if (rc) return rc;
    if (batch == 0) return NGB200_OK;
    int dev = 0;
    rc = current_device(&dev, -1);
    if (rc) return rc;
    Loaded *L = nullptr;
    rc = load_on_device(P, dev, &L);
    if (rc) return rc;
    const int es = esize(P->dtype);
    for (size_t k = 0; k < P->in_width.size(); ++k) {
        if (P->in_width[k] > 0 && !in[k].ptr) return fail(NGB200_ERR_INVALID, "input %zu: null pointer", k);
        if (P->in_mode[k] == NGB200_INDEXED && !in[k].index) return fail(NGB200_ERR_INVALID, "input %zu: indexed operand without index", k);
        if (P->in_mode[k] == NGB200_TILED && in[k].inner_size <= 0) return fail(NGB200_ERR_INVALID, "input %zu: tiled operand needs inner_size > 0", k);
    }
    for (size_t k = 0; k < P->out_width.size(); ++k) {
        if (P->out_width[k] > 0 && !out[k].ptr) return fail(NGB200_ERR_INVALID, "output %zu: null pointer", k);
        if (P->out_mode[k] == NGB200_INDEXED && !out[k].index) return fail(NGB200_ERR_INVALID, "output %zu: indexed operand without index", k);
    }
    bool vec_ok = P->has_vec && batch >= P->vec;
    if (vec_ok) {
        auto aligned = [&](const ngb200_operand &o, int width) {
            if (width == 0) return true;
            if (P->vec == 1) return o.batch_stride == 1;
            return o.batch_stride == 1 && ((uintptr_t)o.ptr % 16 == 0) && (width == 1 || (o.comp_stride * es) % 16 == 0);
        };
        for (size_t k = 0; k < P->in_width.size() && vec_ok; ++k)
            if (P->in_mode[k] == NGB200_VARYING) vec_ok = aligned(in[k], P->in_width[k]);
        for (size_t k = 0; k < P->out_width.size() && vec_ok; ++k) vec_ok = aligned(out[k], P->out_width[k]);
    }
    CUstream s = (CUstream)stream;
    if (P->stage_tpb > 0 && L->fstage && batch >= (int64_t)P->stage_tpb * 4) {
        // staged kernel: needs batch_stride 1 and 16-byte aligned rows for the bulk copies
        bool ok = true;
        const int64_t tile_bytes = (int64_t)P->stage_tpb * es;
        for (size_t k = 0; k < P->in_width.size() && ok; ++k) {
            if (P->in_mode[k] != NGB200_VARYING || P->in_width[k] == 0) continue;
            ok = in[k].batch_stride == 1 && ((uintptr_t)in[k].ptr % 16 == 0) &&
                 (P->in_width[k] == 1 || (in[k].comp_stride * es) % 16 == 0) && tile_bytes % 16 == 0;
        }
        for (size_t k = 0; k < P->out_width.size() && ok; ++k) ok = P->out_width[k] == 0 || out[k].batch_stride == 1;
        if (ok) {
            const int64_t ntiles = batch / P->stage_tpb, done = ntiles * P->stage_tpb;
            rc = launch_one(P, L->fstage, L->grid_stage, ntiles * P->stage_tpb, in, out, params, done, 0, s, P->stage_tpb, P->stage_smem);
            if (rc) return rc;
            if (done < batch) rc = launch_one(P, L->fscalar, L->grid_scalar, batch - done, in, out, params, batch - done, done, s);
            return rc;
        }
    }
    if (vec_ok) {
        int64_t nvec = batch / P->vec, done = nvec * P->vec;
        rc = launch_one(P, L->fvec, L->grid_vec, nvec, in, out, params, done, 0, s);
        if (rc) return rc;
        if (done < batch) rc = launch_one(P, L->fscalar, L->grid_scalar, batch - done, in, out, params, batch - done, done, s);
        return rc;
    }
    CUfunction fs = L->fscalar;
    if (L->fscalar_u) {
        // batch stride 1 everywhere it is used (VARYING / INDEXED operands; TILED keeps its run-time strides)
        bool unit = true;
        for (size_t k = 0; k < P->in_width.size() && unit; ++k)
            if (P->in_width[k] > 0 && (P->in_mode[k] == NGB200_VARYING || P->in_mode[k] == NGB200_INDEXED)) unit = in[k].batch_stride == 1;
        for (size_t k = 0; k < P->out_width.size() && unit; ++k)
            if (P->out_width[k] > 0 && P->out_mode[k] != NGB200_REDUCED) unit = out[k].batch_stride == 1;
        if (unit) fs = L->fscalar_u;
    }
    return launch_one(P, fs, L->grid_scalar, batch, in, out, params, batch, 0, s);
}

// ------------------------------------------------------------------------------------------
// launch plans: everything about a launch that does not change between calls, decided once
// ------------------------------------------------------------------------------------------
// The eager path of a backend issues the same program on operands of the same layout over and over with fresh
// pointers (every operator call allocates its result).  A plan freezes kernel choice, grid sizes and the parameter
// block for one (layout, batch, device); launching it is: patch the pointers into a stack copy of the block,
// re-check pointer alignment for the vector kernel, cuLaunchKernel.  No validation loops, no heap allocation, no
// driver queries.
struct ngb200_plan {
    ngb200_program *P = nullptr;
    int device = 0;
    int n_in = 0, n_out = 0, n_params = 0;
    size_t ob = 32, param_bytes = 0, scal_off = 0;
    bool vec_possible = false;
    int64_t batch = 0;
    struct Shot {
        CUfunction f = nullptr;
        unsigned blocks = 0, tpb = 0;
        int64_t offset = 0;                 // first item of this shot
        char block[1024];                   // parameter block with null pointers
    };
    Shot vec_main, vec_tail, scalar;        // vec path = main (+ tail when batch % vec); scalar path = one shot
    bool has_tail = false;
    std::vector<int64_t> in_bs, out_bs;     // batch strides in elements (to offset tail pointers)
    std::vector<int> in_moves, out_moves;
};

static void plan_fill(ngb200_plan *pl, ngb200_plan::Shot &shot, CUfunction f, int grid_cap, int64_t work, int64_t n,
                      int64_t offset, const ngb200_operand *in, const ngb200_operand *out, int tpb) {
    ngb200_program *P = pl->P;
    memset(shot.block, 0, sizeof shot.block);
    const size_t nin_s = pl->n_in ? pl->n_in : 1, nout_s = pl->n_out ? pl->n_out : 1, np_s = P->n_params > 0 ? P->n_params : 1;
    for (int k = 0; k < pl->n_in; ++k) {
        DevOperand d{nullptr, in[k].comp_stride, P->in_mode[k] == NGB200_BROADCAST ? 0 : in[k].batch_stride, nullptr, 0, 0};
        memcpy(shot.block + k * pl->ob, &d, pl->ob);
    }
    for (int k = 0; k < pl->n_out; ++k) {
        DevOperand d{nullptr, out[k].comp_stride, out[k].batch_stride, nullptr, 0, 0};
        memcpy(shot.block + (nin_s + k) * pl->ob, &d, pl->ob);
    }
    *reinterpret_cast<long long *>(shot.block + pl->ob * (nin_s + nout_s) + 8 * np_s) = n;
    int64_t blocks = (work + tpb - 1) / tpb;
    if (blocks > grid_cap) blocks = grid_cap;
    if (blocks < 1) blocks = 1;
    shot.f = f;
    shot.blocks = (unsigned)blocks;
    shot.tpb = (unsigned)tpb;
    shot.offset = offset;
}

extern "C" int ngb200_plan_create(ngb200_program *P, const ngb200_operand *in, const ngb200_operand *out, int64_t batch,
                                  ngb200_plan **result) {
    if (!P || !out || !result || (!in && !P->in_width.empty())) return fail(NGB200_ERR_INVALID, "null argument");
    if (batch <= 0) return fail(NGB200_ERR_INVALID, "plans need a positive batch");
    if (P->any_indexed || P->any_reduced || P->any_tiled || P->stage_tpb > 0)
        return fail(NGB200_ERR_INVALID, "plans cover VARYING / BROADCAST operands only (use ngb200_program_launch)");
    if (P->param_bytes > sizeof(ngb200_plan::Shot::block)) return fail(NGB200_ERR_INVALID, "too many operands for a plan");
    int rc = need_driver();
    if (rc) return rc;
    int dev = 0;
    rc = current_device(&dev, -1);
    if (rc) return rc;
    Loaded *L = nullptr;
    rc = load_on_device(P, dev, &L);
    if (rc) return rc;
    ngb200_plan *pl = new ngb200_plan;
    pl->P = P;
    pl->device = dev;
    pl->n_in = (int)P->in_width.size();
    pl->n_out = (int)P->out_width.size();
    pl->n_params = P->n_params;
    pl->ob = 32;
    pl->param_bytes = P->param_bytes;
    pl->scal_off = pl->ob * ((pl->n_in ? pl->n_in : 1) + (pl->n_out ? pl->n_out : 1));
    pl->batch = batch;
    const int es = esize(P->dtype);
    for (int k = 0; k < pl->n_in; ++k) { pl->in_bs.push_back(in[k].batch_stride); pl->in_moves.push_back(P->in_mode[k] == NGB200_VARYING); }
    for (int k = 0; k < pl->n_out; ++k) { pl->out_bs.push_back(out[k].batch_stride); pl->out_moves.push_back(1); }
    bool vec_ok = P->has_vec && batch >= P->vec;
    if (vec_ok) {
        auto strides_ok = [&](const ngb200_operand &o, int width) {
            if (width == 0) return true;
            if (P->vec == 1) return o.batch_stride == 1;
            return o.batch_stride == 1 && (width == 1 || (o.comp_stride * es) % 16 == 0);
        };
        for (int k = 0; k < pl->n_in && vec_ok; ++k)
            if (P->in_mode[k] == NGB200_VARYING) vec_ok = strides_ok(in[k], P->in_width[k]);
        for (int k = 0; k < pl->n_out && vec_ok; ++k) vec_ok = strides_ok(out[k], P->out_width[k]);
    }
    pl->vec_possible = vec_ok;
    if (vec_ok) {
        const int64_t nvec = batch / P->vec, done = nvec * P->vec;
        plan_fill(pl, pl->vec_main, L->fvec, L->grid_vec, nvec, done, 0, in, out, P->tpb);
        pl->has_tail = done < batch;
        if (pl->has_tail) plan_fill(pl, pl->vec_tail, L->fscalar, L->grid_scalar, batch - done, batch - done, done, in, out, P->tpb);
    }
    CUfunction fs = L->fscalar;
    if (L->fscalar_u) {
        bool unit = true;
        for (int k = 0; k < pl->n_in && unit; ++k)
            if (P->in_width[k] > 0 && P->in_mode[k] == NGB200_VARYING) unit = in[k].batch_stride == 1;
        for (int k = 0; k < pl->n_out && unit; ++k)
            if (P->out_width[k] > 0) unit = out[k].batch_stride == 1;
        if (unit) fs = L->fscalar_u;
    }
    plan_fill(pl, pl->scalar, fs, L->grid_scalar, batch, batch, 0, in, out, P->tpb);
    *result = pl;
    return NGB200_OK;
}

extern "C" void ngb200_plan_destroy(ngb200_plan *pl) { delete pl; }

static inline int plan_shoot(ngb200_plan *pl, const ngb200_plan::Shot &shot, void *const *in_ptrs, void *const *out_ptrs,
                             const double *params, CUstream stream) {
    alignas(16) char buf[sizeof shot.block];
    memcpy(buf, shot.block, pl->param_bytes);
    const int es = esize(pl->P->dtype);
    const size_t nin_s = pl->n_in ? pl->n_in : 1;
    for (int k = 0; k < pl->n_in; ++k) {
        char *p = (char *)in_ptrs[k] + (pl->in_moves[k] ? shot.offset * pl->in_bs[k] * es : 0);
        memcpy(buf + k * pl->ob, &p, sizeof p);
    }
    for (int k = 0; k < pl->n_out; ++k) {
        char *p = (char *)out_ptrs[k] + shot.offset * pl->out_bs[k] * es;
        memcpy(buf + (nin_s + k) * pl->ob, &p, sizeof p);
    }
    if (pl->n_params) memcpy(buf + pl->scal_off, params, 8 * (size_t)pl->n_params);
    void *args[] = {buf};
    CU(cuLaunchKernel(shot.f, shot.blocks, 1, 1, shot.tpb, 1, 1, 0, stream, args, nullptr));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return NGB200_OK;
}

// The calling thread's current CUDA context must be the one the plan was created in (true whenever `stream` is the
// caller's current stream on that device).
extern "C" int ngb200_plan_launch(ngb200_plan *pl, void *const *in_ptrs, void *const *out_ptrs, const double *params,
                                  void *stream) {
    if (!pl) return fail(NGB200_ERR_INVALID, "null plan");
    bool vec = pl->vec_possible;
    if (vec && pl->P->vec > 1) {
        for (int k = 0; k < pl->n_in && vec; ++k)
            if (pl->in_moves[k] && pl->P->in_width[k]) vec = ((uintptr_t)in_ptrs[k] & 15) == 0;
        for (int k = 0; k < pl->n_out && vec; ++k)
            if (pl->P->out_width[k]) vec = ((uintptr_t)out_ptrs[k] & 15) == 0;
    }
    CUstream s = (CUstream)stream;
    if (vec) {
        int rc = plan_shoot(pl, pl->vec_main, in_ptrs, out_ptrs, params, s);
        if (rc || !pl->has_tail) return rc;
        return plan_shoot(pl, pl->vec_tail, in_ptrs, out_ptrs, params, s);
    }
    return plan_shoot(pl, pl->scalar, in_ptrs, out_ptrs, params, s);
}

// ------------------------------------------------------------------------------------------
// pipelines: several programs in ONE kernel, exchanging values through shared memory
// ------------------------------------------------------------------------------------------
// A pipeline runs a sequence of PHASES over TILES.  Every phase is a straight-line program over the items of one
// item domain (e.g. "bodies" or "constraints of colour set 0"); tile t holds items [t*ipt[d], (t+1)*ipt[d]) of every
// domain d and is processed by one CTA.  Operands of a phase are bound either to a GLOBAL array (read / written in
// HBM, optionally through an int64 index array = gather / scatter) or to a SHARED buffer that lives in the CTA's
// shared memory for the whole tile (component-major [width][ipt]), optionally through an index array whose entries
// must fall inside the tile.  Phases are separated by __syncthreads(); a contiguous range of phases can be repeated
// (run-time count).  This is what lets a chain of dependent data-parallel steps with neighbour coupling (the
// rigid-body sub-step: integrate -> relax constraint colour sets -> update rates) touch HBM once per launch.
struct PipePhase {
    std::unique_ptr<ngb200_program> prog;
    int domain = 0;
    std::vector<ngb200_bind> in, out;
};

struct PipeLoaded {
    std::atomic<int> ready{0};
    CUmodule mod = nullptr;
    CUfunction f = nullptr;
    int regs = 0, grid = 0;
    size_t local_bytes = 0;
};

struct ngb200_pipeline {
    int dtype = 0;
    uint32_t flags = 0;
    std::vector<int32_t> ipt, gwidth, gdomain, gindex, swidth, sdomain;
    std::vector<int> gslot;                   // global operand -> slot among value operands or among index operands
    std::vector<PipePhase> phases;
    int loop_begin = 0, loop_end = 0, n_params = 0;
    int tpb = 128, minb = 1;
    int warps = 1;                            // > 1: WARP TILES -- a tile is processed by ONE warp (tpb == 32), `warps` tiles per CTA, __syncwarp() barriers
    int n_value_globals = 0, n_index_globals = 0;
    size_t smem = 0, param_bytes = 0;
    std::vector<size_t> soff;                 // element offset of every shared buffer (component-major layout)
    // record layout (NGB200_PIPE_RECORDS, default): all shared buffers of one domain form ONE record per item, accessed
    // with 128-bit shared loads / stores; records are padded to an ODD number of 16-byte chunks plus one chunk every 8
    // rows, which makes unit-stride AND stride-2 row accesses (the two bodies of neighbouring constraints) bank-conflict free
    bool records = true;
    std::vector<size_t> rec_off;              // element offset of every shared buffer inside its domain's record
    std::vector<size_t> dom_base;             // element offset of every domain's record array
    std::vector<int> dom_chunks;              // chunks per record, per domain
    int vw = 4, pad_shift = 3;                // elements per chunk (fp32: 4 = 16 bytes; fp64: 1 = 8 bytes), one padding chunk per 2^pad_shift rows
    std::string source, hash;
    std::vector<char> cubin;
    bool cache_hit = false;
    long spill = 0;
    int n_live = 0;
    std::mutex mu;
    PipeLoaded loaded[kMaxDevices];
};

static std::string pipeline_source(ngb200_pipeline &Q, bool ir_order, std::string *err) {
    const bool f64 = Q.dtype == NGB200_F64;
    const int es = f64 ? 8 : 4;
    std::string s;
    std::vector<std::unique_ptr<Gen>> gens;
    int max_need = 0, total_nu = 0;
    Q.n_live = 0;
    for (size_t k = 0; k < Q.phases.size(); ++k) {
        PipePhase &ph = Q.phases[k];
        gens.emplace_back(new Gen(*ph.prog));
        Gen &g = *gens.back();
        g.faithful_order = ir_order;
        if (g.analyse() != NGB200_OK) { *err = g_err; return ""; }
        g.plan_contraction();
        g.suffix = "_" + std::to_string(k);
        g.params_type = "PParams";
        g.bcast_ref.assign(ph.in.size(), "");
        for (size_t j = 0; j < ph.in.size(); ++j)
            if (ph.prog->in_mode[j] == NGB200_BROADCAST) g.bcast_ref[j] = "p.g[" + std::to_string(Q.gslot[ph.in[j].slot]) + "]";
        const int need = g.max_live * (f64 ? 2 : 1) + 24;
        if (need > max_need) max_need = need;
        total_nu += g.NU;
        // wide shared-memory tables are re-read at every use instead of being held in registers for the whole phase
        g.remat.assign(ph.in.size(), -1);
        if (Q.records && env_int("NGB200_PIPE_REMAT", f64 ? 1 : 0))
            for (size_t j = 0; j < ph.in.size(); ++j)
                if (ph.in[j].kind == NGB200_BIND_SHARED && ph.prog->in_width[j] >= env_int("NGB200_PIPE_REMAT_WIDTH", 16))
                    g.remat[j] = g.n_remat++;
    }
    max_need += total_nu * (f64 ? 2 : 1);
    const int cta_threads = Q.tpb * Q.warps;
    int minb = 65536 / (cta_threads * (max_need > 32 ? max_need : 32));
    Q.minb = minb < 1 ? 1 : (minb > 8 ? 8 : minb);
    if (env_int("NGB200_PIPE_MINB", 0) > 0) Q.minb = env_int("NGB200_PIPE_MINB", 0);
    gens[0]->emit_prelude(s, 1, env_int("NGB200_LDMODE", 0), env_int("NGB200_STMODE", 0), false);
    appendf(s, "#define NGB_TPB %d\n", Q.tpb);
    if (f64)
        s += "__device__ __forceinline__ double ngb_lds(const double* p) { double r; asm volatile(\"ld.shared.f64 %0, [%1];\" : \"=d\"(r) : \"r\"((unsigned)__cvta_generic_to_shared(p)) : \"memory\"); return r; }\n";
    else
        s += "__device__ __forceinline__ float ngb_lds(const float* p) { float r; asm volatile(\"ld.shared.f32 %0, [%1];\" : \"=f\"(r) : \"r\"((unsigned)__cvta_generic_to_shared(p)) : \"memory\"); return r; }\n";
    s += "struct PG { T* ptr; long long cs; };\n";
    appendf(s, "struct PParams { PG g[%d]; const long long* idx[%d]; double scal[%d]; long long n[%d]; long long n_tiles; int repeat; };\n",
            Q.n_value_globals > 0 ? Q.n_value_globals : 1, Q.n_index_globals > 0 ? Q.n_index_globals : 1,
            Q.n_params > 0 ? Q.n_params : 1, (int)Q.ipt.size());
    for (size_t k = 0; k < Q.phases.size(); ++k) {
        int nl = 0, nu = 0;
        gens[k]->emit_uniform(s, &nl, &nu);
        gens[k]->emit_body(s, &nl);
        Q.n_live += nl;
    }
    // NGB200_PIPE_PREFETCH=2: while a tile computes, the rows the NEXT tile of this CTA will read from global memory are
    // pulled into L2 with one bulk prefetch per (array, component) row (cp.async.bulk.prefetch.L2: rows are contiguous per
    // tile in the SoA layout), so the next load phase finds them in L2 and DRAM works during the compute phases.
    struct PfRow { int kind, slot, comp, dom; };
    std::vector<PfRow> pf_rows;
    if (env_int("NGB200_PIPE_PREFETCH", 0) == 2) {
        std::set<std::tuple<int, int, int, int>> seen;
        for (size_t k = 0; k < Q.phases.size(); ++k) {
            PipePhase &ph = Q.phases[k];
            Gen &g = *gens[k];
            std::vector<char> used(g.NX > 0 ? g.NX : 1, 0);
            for (size_t v = 0; v < ph.prog->instr.size(); ++v) {
                const ngb200_instr &I = ph.prog->instr[v];
                if (I.op == NGB200_OP_INPUT && g.live[v] && !g.uniform[v]) used[g.xoff[I.a] + I.b] = 1;
            }
            for (size_t j = 0; j < ph.in.size(); ++j) {
                const ngb200_bind &b = ph.in[j];
                if (b.index >= 0) seen.insert(std::make_tuple(1, Q.gslot[b.index], 0, ph.domain));
                if (b.kind != NGB200_BIND_GLOBAL || b.index >= 0 || ph.prog->in_mode[j] == NGB200_BROADCAST) continue;
                for (int c = 0; c < ph.prog->in_width[j]; ++c)
                    if (used[g.xoff[j] + c]) seen.insert(std::make_tuple(0, Q.gslot[b.slot], c, ph.domain));
            }
            for (const ngb200_bind &b : ph.out)
                if (b.index >= 0) seen.insert(std::make_tuple(1, Q.gslot[b.index], 0, ph.domain));
        }
        for (auto &t : seen) pf_rows.push_back({std::get<0>(t), std::get<1>(t), std::get<2>(t), std::get<3>(t)});
        if (!pf_rows.empty()) {
            const char *names[4] = {"kind", "slot", "comp", "dom"};
            for (int f = 0; f < 4; ++f) {
                appendf(s, "__constant__ unsigned char ngb_pf_%s[%zu] = {", names[f], pf_rows.size());
                for (size_t r = 0; r < pf_rows.size(); ++r)
                    appendf(s, "%s%d", r ? "," : "", f == 0 ? pf_rows[r].kind : f == 1 ? pf_rows[r].slot : f == 2 ? pf_rows[r].comp : pf_rows[r].dom);
                s += "};\n";
            }
            appendf(s, "__constant__ int ngb_pf_ipt[%zu] = {", Q.ipt.size());
            for (size_t d = 0; d < Q.ipt.size(); ++d) appendf(s, "%s%d", d ? "," : "", Q.ipt[d]);
            s += "};\n";
        }
    }
    if (Q.minb > 1 && 65536 / (cta_threads * Q.minb) < 255) appendf(s, "extern \"C\" __global__ void __launch_bounds__(%d, %d) ngb_pipeline(const PParams p) {\n", cta_threads, Q.minb);
    else appendf(s, "extern \"C\" __global__ void __launch_bounds__(%d) ngb_pipeline(const PParams p) {\n", cta_threads);
    s += "  extern __shared__ __align__(16) unsigned char ngb_raw[];\n";
    if (Q.warps > 1) {
        // warp tiles: every warp of the CTA owns its own tile and its own slice of shared memory; nothing is shared between
        // warps, so the phase barriers are __syncwarp() and warps never wait for each other
        appendf(s, "  const int ngb_tid = threadIdx.x & 31;\n  T* const sm = reinterpret_cast<T*>(ngb_raw) + (threadIdx.x >> 5) * %zuLL;\n", Q.smem / esize(Q.dtype) / Q.warps);
    } else {
        s += "  const int ngb_tid = threadIdx.x;\n  T* const sm = reinterpret_cast<T*>(ngb_raw);\n";
    }
    for (size_t k = 0; k < Q.phases.size(); ++k)
        appendf(s, "  T u%zu[%d]; ngb_uniform_%zu(p, u%zu);\n", k, gens[k]->NU > 0 ? gens[k]->NU : 1, k, k);
    if (Q.warps > 1)
        appendf(s, "  for (long long tile = (long long)blockIdx.x * %d + (threadIdx.x >> 5); tile < p.n_tiles; tile += (long long)gridDim.x * %d) {\n", Q.warps, Q.warps);
    else
        s += "  for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {\n";
    // Global operands that no phase writes are read through the non-coherent path (`ld.global.nc`): besides the cache
    // path, it tells ptxas the data cannot change under the kernel, so loads are scheduled next to their uses
    // (ChainStep fp32: 168 -> 128 registers, +1 % / +3 % at ten / one sub-steps per launch).
    std::vector<char> readonly(Q.gwidth.size(), (char)(env_int("NGB200_PIPE_LDG", 1) != 0));
    for (const PipePhase &q : Q.phases)
        for (const ngb200_bind &b : q.out)
            if (b.kind == NGB200_BIND_GLOBAL) readonly[b.slot] = 0;
    auto emit_phase = [&](size_t k, const char *ind) {
        PipePhase &ph = Q.phases[k];
        Gen &g = *gens[k];
        const int d = ph.domain, ipt = Q.ipt[d];
        appendf(s, "%s{ // phase %zu: domain %d, %d items per tile\n", ind, k, d, ipt);
        if (g.has_partner) {
            // the phase exchanges values between the two threads of an item pair (NGB200_OP_PARTNER = warp shuffle): EVERY
            // lane executes the body; lanes without an item compute on the tile's first item and store nothing
            appendf(s, "%s  if (tile * %dLL < p.n[%d]) for (int ngb_i0 = 0; ngb_i0 < %d; ngb_i0 += NGB_TPB) {\n", ind, ipt, d, ipt);
            appendf(s, "%s    const bool ngb_on = ngb_i0 + ngb_tid < %d && tile * %dLL + ngb_i0 + ngb_tid < p.n[%d];\n", ind, ipt, ipt, d);
            appendf(s, "%s    const int i = ngb_on ? ngb_i0 + ngb_tid : 0;\n", ind);
            appendf(s, "%s    const long long gi = tile * %dLL + i;\n%s    {\n", ind, ipt, ind);
        } else if (ipt <= Q.tpb) appendf(s, "%s  const int i = ngb_tid;\n%s  if (i < %d) {\n", ind, ind, ipt);
        else {
            // several items per thread: copy phases (no arithmetic: tile load / store) are unrolled so that the global
            // loads of ALL of a thread's items are in flight together -- one DRAM round trip per phase instead of one per
            // item (NGB200_PIPE_UNROLL: 0 never, 1 copy phases (default), 2 every phase)
            const int mode = env_int("NGB200_PIPE_UNROLL", 1);
            size_t arith = 0;
            for (const ngb200_instr &I : ph.prog->instr) if (I.op != NGB200_OP_INPUT) ++arith;
            if ((mode == 1 && arith == 0 && ipt <= 4 * Q.tpb) || mode >= 2) appendf(s, "%s  #pragma unroll\n", ind);
            appendf(s, "%s  for (int i = ngb_tid; i < %d; i += NGB_TPB) {\n", ind, ipt);
        }
        if (!g.has_partner) appendf(s, "%s    const long long gi = tile * %dLL + i;\n%s    if (gi < p.n[%d]) {\n", ind, ipt, ind, d);
        appendf(s, "%s      T x[%d]; T y[%d];\n", ind, g.NX > 0 ? g.NX : 1, g.NY > 0 ? g.NY : 1);
        // row of an indexed access: global arrays use the index value itself, shared buffers the tile-local part
        std::map<std::pair<int, int>, std::string> rows;
        auto row_of = [&](const ngb200_bind &b) -> std::string {
            const bool shared = b.kind == NGB200_BIND_SHARED;
            const int target = shared ? Q.sdomain[b.slot] : -1;
            auto key = std::make_pair(b.index, target);
            auto it = rows.find(key);
            if (it != rows.end()) return it->second;
            std::string name = "row" + std::to_string(rows.size());
            if (shared)
                appendf(s, "%s      const int %s = (int)(p.idx[%d][gi] - tile * %dLL);\n", ind, name.c_str(), Q.gslot[b.index], Q.ipt[target]);
            else
                appendf(s, "%s      const long long %s = p.idx[%d][gi];\n", ind, name.c_str(), Q.gslot[b.index]);
            rows[key] = name;
            return name;
        };
        auto address = [&](const ngb200_bind &b, int c) -> std::string {
            char buf[256];
            if (b.kind == NGB200_BIND_SHARED) {
                const int w = Q.ipt[Q.sdomain[b.slot]];
                std::string r = b.index >= 0 ? row_of(b) : std::string("i");
                snprintf(buf, sizeof buf, "sm + %zu + %d * %d + %s", Q.soff[b.slot], c, w, r.c_str());
            } else {
                std::string r = b.index >= 0 ? row_of(b) : std::string("gi");
                snprintf(buf, sizeof buf, "p.g[%d].ptr + %dLL * p.g[%d].cs + %s", Q.gslot[b.slot], c, Q.gslot[b.slot], r.c_str());
            }
            return buf;
        };
        std::vector<char> used(g.NX > 0 ? g.NX : 1, 0);
        for (size_t v = 0; v < ph.prog->instr.size(); ++v) {
            const ngb200_instr &I = ph.prog->instr[v];
            if (I.op == NGB200_OP_INPUT && g.live[v] && !g.uniform[v]) used[g.xoff[I.a] + I.b] = 1;
        }
        const int VW = Q.records ? Q.vw : (f64 ? 2 : 4);    // elements per chunk of the record layout
        const char *VTn = f64 ? "double2" : "float4";
        static const char *lane4[4] = {"x", "y", "z", "w"};
        // element offset of (row r) of shared buffer `slot` in the record layout
        auto record_base = [&](const ngb200_bind &b) -> std::string {
            const int dd = Q.sdomain[b.slot];
            std::string r = b.index >= 0 ? row_of(b) : std::string("i");
            char buf[256];
            snprintf(buf, sizeof buf, "sm + %zu + (%s * %d + (%s >> %d)) * %d + %zu", Q.dom_base[dd], r.c_str(), Q.dom_chunks[dd],
                     r.c_str(), Q.pad_shift, VW, Q.rec_off[b.slot]);
            return buf;
        };
        if (g.n_remat > 0) {
            appendf(s, "%s      const T* xp[%d];\n", ind, g.n_remat);
            for (size_t j = 0; j < ph.in.size(); ++j)
                if (g.remat[j] >= 0) appendf(s, "%s      xp[%d] = %s;\n", ind, g.remat[j], record_base(ph.in[j]).c_str());
        }
        for (size_t j = 0; j < ph.in.size(); ++j) {
            if (ph.prog->in_mode[j] == NGB200_BROADCAST) continue;
            if (g.remat[j] >= 0) continue;                  // re-read at every use inside the body
            if (Q.records && ph.in[j].kind == NGB200_BIND_SHARED) {
                const int w = ph.prog->in_width[j];
                std::string base = record_base(ph.in[j]);
                appendf(s, "%s      { const T* rp = %s;\n", ind, base.c_str());
                for (int c0 = 0; c0 < w; c0 += VW) {
                    bool any = false;
                    for (int c = c0; c < std::min(w, c0 + VW); ++c) any = any || used[g.xoff[j] + c];
                    if (!any) continue;
                    if (VW == 1) { appendf(s, "%s        x[%d] = rp[%d];\n", ind, g.xoff[j] + c0, c0); continue; }
                    appendf(s, "%s        { const %s t = *reinterpret_cast<const %s*>(rp + %d);", ind, VTn, VTn, c0);
                    for (int c = c0; c < std::min(w, c0 + VW); ++c)
                        if (used[g.xoff[j] + c]) appendf(s, " x[%d] = t.%s;", g.xoff[j] + c, lane4[c - c0]);
                    s += " }\n";
                }
                appendf(s, "%s      }\n", ind);
                continue;
            }
            for (int c = 0; c < ph.prog->in_width[j]; ++c) {
                if (!used[g.xoff[j] + c]) continue;          // structurally dead component: never loaded
                std::string a = address(ph.in[j], c);
                if (ph.in[j].kind == NGB200_BIND_SHARED) appendf(s, "%s      x[%d] = *(%s);\n", ind, g.xoff[j] + c, a.c_str());
                else if (readonly[ph.in[j].slot]) appendf(s, "%s      x[%d] = __ldg(%s);\n", ind, g.xoff[j] + c, a.c_str());
                else appendf(s, "%s      x[%d] = NGB_LD(%s);\n", ind, g.xoff[j] + c, a.c_str());
            }
        }
        if (g.n_remat > 0) appendf(s, "%s      ngb_body_%zu(x, xp, u%zu, y);\n", ind, k, k);
        else appendf(s, "%s      ngb_body_%zu(x, u%zu, y);\n", ind, k, k);
        if (g.has_partner) appendf(s, "%s      if (ngb_on) {\n", ind);
        for (size_t j = 0; j < ph.out.size(); ++j) {
            if (Q.records && ph.out[j].kind == NGB200_BIND_SHARED) {
                const int w = ph.prog->out_width[j];
                std::string base = record_base(ph.out[j]);
                appendf(s, "%s      { T* rp = %s;\n", ind, base.c_str());
                for (int c0 = 0; c0 < w; c0 += VW) {           // (the padding lanes of the last chunk belong to this buffer)
                    if (VW == 1) { appendf(s, "%s        rp[%d] = y[%d];\n", ind, c0, g.yoff[j] + c0); continue; }
                    appendf(s, "%s        { %s t;", ind, VTn);
                    for (int l = 0; l < VW; ++l) {
                        if (c0 + l < w) appendf(s, " t.%s = y[%d];", lane4[l], g.yoff[j] + c0 + l);
                        else appendf(s, " t.%s = T(0);", lane4[l]);
                    }
                    appendf(s, " *reinterpret_cast<%s*>(rp + %d) = t; }\n", VTn, c0);
                }
                appendf(s, "%s      }\n", ind);
                continue;
            }
            for (int c = 0; c < ph.prog->out_width[j]; ++c) {
                std::string a = address(ph.out[j], c);
                if (ph.out[j].kind == NGB200_BIND_SHARED) appendf(s, "%s      *(%s) = y[%d];\n", ind, a.c_str(), g.yoff[j] + c);
                else appendf(s, "%s      NGB_ST(%s, y[%d]);\n", ind, a.c_str(), g.yoff[j] + c);
            }
        }
        if (g.has_partner) appendf(s, "%s      }\n", ind);
        appendf(s, "%s    }\n%s  }\n%s}\n", ind, ind, ind);
        // A barrier separates this phase from the next one in program order -- unless both run over the same domain and
        // touch shared buffers only at their own item (no index): then thread i of the next phase reads exactly what
        // thread i wrote here (same thread: same item mapping), and no other thread's data.
        const int last = (int)Q.phases.size() - 1;
        const int next = ((int)k + 1 == Q.loop_end && Q.loop_end > Q.loop_begin) ? -1 : ((int)k < last ? (int)k + 1 : -1);
        auto direct_only = [&](const PipePhase &q) {
            for (const ngb200_bind &b : q.in) if (b.kind == NGB200_BIND_SHARED && b.index >= 0) return false;
            for (const ngb200_bind &b : q.out) if (b.kind == NGB200_BIND_SHARED && b.index >= 0) return false;
            return true;
        };
        const bool elide = next >= 0 && Q.phases[next].domain == ph.domain && direct_only(ph) && direct_only(Q.phases[next]) &&
                           env_int("NGB200_PIPE_ELIDE", 1);
        if (!elide) appendf(s, "%s%s;\n", ind, Q.warps > 1 ? "__syncwarp()" : "__syncthreads()");
    };
    // L2 prefetch of what the LATER phases of this tile read from global memory (per-item constants and tables that are
    // not parked in shared memory): issued before the first phase, so that their DRAM latency overlaps the load phase
    // and the first compute phases instead of being exposed at the head of every later phase.
    if (env_int("NGB200_PIPE_PREFETCH", 0) == 1) {     // measured on B200: 5.41 vs 5.77 G body-steps/s with it on (profiles/r02_chain_sweep.txt): off
        std::map<std::pair<int, int>, int> seen;        // (global slot, component) -> domain
        for (size_t k = 1; k < Q.phases.size(); ++k) {
            PipePhase &ph = Q.phases[k];
            Gen &g = *gens[k];
            std::vector<char> used(g.NX > 0 ? g.NX : 1, 0);
            for (size_t v = 0; v < ph.prog->instr.size(); ++v) {
                const ngb200_instr &I = ph.prog->instr[v];
                if (I.op == NGB200_OP_INPUT && g.live[v] && !g.uniform[v]) used[g.xoff[I.a] + I.b] = 1;
            }
            for (size_t j = 0; j < ph.in.size(); ++j) {
                const ngb200_bind &b = ph.in[j];
                if (b.kind != NGB200_BIND_GLOBAL || b.index >= 0 || ph.prog->in_mode[j] == NGB200_BROADCAST) continue;
                for (int c = 0; c < ph.prog->in_width[j]; ++c)
                    if (used[g.xoff[j] + c]) seen[std::make_pair(b.slot, c)] = ph.domain;
            }
        }
        std::map<int, std::vector<std::pair<int, int>>> by_domain;
        for (auto &kv : seen) by_domain[kv.second].push_back(kv.first);
        for (auto &kv : by_domain) {
            const int d = kv.first, ipt = Q.ipt[d];
            // one prefetch per 32-byte sector: every 8th (fp32) / 4th (fp64) item of a row issues it
            const int per = 32 / es;
            appendf(s, "    for (int i = ngb_tid * %d; i < %d; i += NGB_TPB * %d) {\n", per, ipt, per);
            appendf(s, "      const long long gi = tile * %dLL + i;\n      if (gi < p.n[%d]) {\n", ipt, d);
            for (auto &sc : kv.second)
                appendf(s, "        asm volatile(\"prefetch.global.L2 [%%0];\" :: \"l\"(p.g[%d].ptr + %dLL * p.g[%d].cs + gi));\n",
                        Q.gslot[sc.first], sc.second, Q.gslot[sc.first]);
            s += "      }\n    }\n";
        }
    }
    auto emit_next_tile_prefetch = [&]() {
        if (pf_rows.empty()) return;
        if (Q.warps > 1) appendf(s, "    { const long long nt = tile + (long long)gridDim.x * %d;\n", Q.warps);
        else s += "    { const long long nt = tile + gridDim.x;\n";
        appendf(s, "      if (nt < p.n_tiles) for (int r = ngb_tid; r < %zu; r += NGB_TPB) {\n", pf_rows.size());
        s += "        const int d = ngb_pf_dom[r], sl = ngb_pf_slot[r];\n"
             "        const long long start = nt * ngb_pf_ipt[d];\n"
             "        long long cnt = p.n[d] - start; if (cnt > ngb_pf_ipt[d]) cnt = ngb_pf_ipt[d];\n"
             "        const char* a; long long bytes;\n"
             "        if (ngb_pf_kind[r]) { a = (const char*)(p.idx[sl] + start); bytes = cnt * 8; }\n"
             "        else { a = (const char*)(p.g[sl].ptr + (long long)ngb_pf_comp[r] * p.g[sl].cs + start); bytes = cnt * (long long)sizeof(T); }\n"
             "        const unsigned head = (unsigned)((16 - ((unsigned long long)a & 15)) & 15);\n"
             "        a += head; bytes = (bytes - head) & ~15LL;\n"
             "        if (bytes > 0) asm volatile(\"cp.async.bulk.prefetch.L2.global [%0], %1;\" :: \"l\"(a), \"r\"((unsigned)bytes) : \"memory\");\n"
             "      }\n    }\n";
    };
    for (int k = 0; k < Q.loop_begin; ++k) { emit_phase(k, "    "); if (k == 0) emit_next_tile_prefetch(); }
    if (Q.loop_begin == 0) emit_next_tile_prefetch();
    if (Q.loop_end > Q.loop_begin) {
        s += "    for (int rep = 0; rep < p.repeat; ++rep) {\n";
        for (int k = Q.loop_begin; k < Q.loop_end; ++k) emit_phase(k, "      ");
        s += "    }\n";
    }
    for (int k = Q.loop_end; k < (int)Q.phases.size(); ++k) emit_phase(k, "    ");
    s += "  }\n}\n";
    (void)es;
    return s;
}

extern "C" int ngb200_pipeline_create(const ngb200_pipeline_desc *d, ngb200_pipeline **out) {
    if (!d || !out) return fail(NGB200_ERR_INVALID, "null argument");
    if (d->dtype != NGB200_F32 && d->dtype != NGB200_F64) return fail(NGB200_ERR_INVALID, "bad dtype %d", d->dtype);
    if (d->n_domains < 1 || d->n_globals < 1 || d->n_shared < 0 || d->n_phases < 1 || d->n_params < 0)
        return fail(NGB200_ERR_INVALID, "bad counts in pipeline description");
    if (d->loop_begin < 0 || d->loop_end < d->loop_begin || d->loop_end > d->n_phases) return fail(NGB200_ERR_INVALID, "bad loop range");
    std::unique_ptr<ngb200_pipeline> Q(new ngb200_pipeline);
    Q->dtype = d->dtype;
    Q->flags = d->flags;
    Q->ipt.assign(d->items_per_tile, d->items_per_tile + d->n_domains);
    Q->gwidth.assign(d->global_width, d->global_width + d->n_globals);
    Q->gdomain.assign(d->global_domain, d->global_domain + d->n_globals);
    Q->gindex.assign(d->global_is_index, d->global_is_index + d->n_globals);
    if (d->n_shared) {
        Q->swidth.assign(d->shared_width, d->shared_width + d->n_shared);
        Q->sdomain.assign(d->shared_domain, d->shared_domain + d->n_shared);
    }
    Q->loop_begin = d->loop_begin;
    Q->loop_end = d->loop_end;
    Q->n_params = d->n_params;
    int max_ipt = 0;
    for (int v : Q->ipt) { if (v < 1 || v > 1024) return fail(NGB200_ERR_INVALID, "items per tile must be 1..1024"); max_ipt = std::max(max_ipt, v); }
    for (int k = 0; k < d->n_globals; ++k) {
        if (Q->gdomain[k] < -1 || Q->gdomain[k] >= d->n_domains) return fail(NGB200_ERR_INVALID, "global %d: bad domain", k);
        Q->gslot.push_back(Q->gindex[k] ? Q->n_index_globals++ : Q->n_value_globals++);
    }
    const int es = esize(d->dtype);
    size_t off = 0;
    for (int k = 0; k < d->n_shared; ++k) {
        if (Q->sdomain[k] < 0 || Q->sdomain[k] >= d->n_domains || Q->swidth[k] < 1) return fail(NGB200_ERR_INVALID, "shared buffer %d: bad description", k);
        Q->soff.push_back(off);
        off += (size_t)Q->swidth[k] * Q->ipt[Q->sdomain[k]];
    }
    Q->smem = off * es;
    Q->records = env_int("NGB200_PIPE_RECORDS", 1) != 0;
    if (Q->records) {
        // fp32: 16-byte chunks (float4 accesses, a quarter warp per wavefront), records of an ODD number of chunks + one chunk
        // every 8 rows.  fp64: most accesses are the 8-byte re-reads of wide tables at their point of use (a HALF warp per
        // wavefront): with 16-byte chunks the row stride is an even number of 8-byte units and rows 8-15 fall on the banks of
        // rows 0-7 (ncu: 647 M conflicts in 1 276 M load wavefronts); 8-byte chunks, an odd number per record and one
        // padding unit every 16 rows make unit-stride and stride-2 rows conflict-free for 8-byte accesses
        // (NGB200_PIPE_F64_UNITS=0 restores the 16-byte chunks).
        const bool units = d->dtype == NGB200_F64 && env_int("NGB200_PIPE_F64_UNITS", 1) != 0;
        Q->vw = units ? 1 : 16 / es;
        Q->pad_shift = units ? 4 : 3;
        const int VW = Q->vw;
        Q->rec_off.assign(d->n_shared, 0);
        Q->dom_base.assign(d->n_domains, 0);
        Q->dom_chunks.assign(d->n_domains, 0);
        size_t total = 0;
        for (int dd = 0; dd < d->n_domains; ++dd) {
            size_t rec = 0;
            for (int k = 0; k < d->n_shared; ++k)
                if (Q->sdomain[k] == dd) { Q->rec_off[k] = rec; rec += (size_t)(Q->swidth[k] + VW - 1) / VW * VW; }
            int chunks = (int)(rec / VW);
            if (chunks && chunks % 2 == 0) chunks += 1;                      // odd number of chunks per record
            Q->dom_chunks[dd] = chunks;
            Q->dom_base[dd] = total;
            if (chunks) total += ((size_t)Q->ipt[dd] * chunks + (Q->ipt[dd] >> Q->pad_shift) + 1) * VW;
        }
        Q->smem = (total * es + 15) / 16 * 16;
    }
    if (Q->smem > 227 * 1024) return fail(NGB200_ERR_INVALID, "tile needs %zu bytes of shared memory (limit 232448)", Q->smem);
    Q->tpb = d->threads_per_tile > 0 ? d->threads_per_tile : std::min(256, (max_ipt + 31) / 32 * 32);
    if (Q->tpb % 32 || Q->tpb > 1024) return fail(NGB200_ERR_INVALID, "threads per tile must be a multiple of 32, at most 1024");
    if (Q->tpb == 32) {                                        // one warp per tile: several tiles per CTA (see pipeline_source)
        Q->warps = env_int("NGB200_PIPE_WARPS", 8);
        if (Q->warps < 1) Q->warps = 1;
        while (Q->warps > 1 && Q->smem * Q->warps > 227 * 1024) Q->warps /= 2;
        Q->smem = (Q->smem + 15) / 16 * 16 * Q->warps;
    }
    Q->param_bytes = 16 * (size_t)std::max(1, Q->n_value_globals) + 8 * (size_t)std::max(1, Q->n_index_globals) +
                     8 * (size_t)std::max(1, Q->n_params) + 8 * Q->ipt.size() + 8 + 8;
    if (Q->param_bytes > 4000) return fail(NGB200_ERR_INVALID, "too many operands for one kernel");
    for (int k = 0; k < d->n_phases; ++k) {
        const ngb200_phase_desc &pd = d->phases[k];
        const ngb200_program_desc &pp = pd.program;
        if (pd.domain < 0 || pd.domain >= d->n_domains) return fail(NGB200_ERR_INVALID, "phase %d: bad domain", k);
        if (pp.n_params > d->n_params) return fail(NGB200_ERR_INVALID, "phase %d: uses more scalar params than the pipeline has", k);
        PipePhase ph;
        ph.domain = pd.domain;
        ph.prog.reset(new ngb200_program);
        ngb200_program &P = *ph.prog;
        P.dtype = d->dtype;
        P.flags = d->flags;
        P.in_width.assign(pp.input_width, pp.input_width + pp.n_inputs);
        P.out_width.assign(pp.output_width, pp.output_width + pp.n_outputs);
        P.out_mode.assign(pp.n_outputs, NGB200_VARYING);
        P.n_params = d->n_params;
        if (pp.n_consts) P.consts.assign(pp.consts, pp.consts + pp.n_consts);
        if (pp.n_instr) P.instr.assign(pp.instr, pp.instr + pp.n_instr);
        if (pp.n_stores) P.stores.assign(pp.stores, pp.stores + pp.n_stores);
        ph.in.assign(pd.inputs, pd.inputs + pp.n_inputs);
        ph.out.assign(pd.outputs, pd.outputs + pp.n_outputs);
        auto check = [&](const ngb200_bind &b, int width, bool is_out, int j) -> int {
            const char *what = is_out ? "output" : "input";
            if (b.kind == NGB200_BIND_SHARED) {
                if (b.slot < 0 || b.slot >= d->n_shared) return fail(NGB200_ERR_INVALID, "phase %d %s %d: bad shared buffer", k, what, j);
                if (Q->swidth[b.slot] != width) return fail(NGB200_ERR_INVALID, "phase %d %s %d: width %d != shared buffer width %d", k, what, j, width, Q->swidth[b.slot]);
                if (b.index < 0 && Q->sdomain[b.slot] != pd.domain) return fail(NGB200_ERR_INVALID, "phase %d %s %d: direct access to a buffer of another domain", k, what, j);
            } else if (b.kind == NGB200_BIND_GLOBAL) {
                if (b.slot < 0 || b.slot >= d->n_globals || Q->gindex[b.slot]) return fail(NGB200_ERR_INVALID, "phase %d %s %d: bad global operand", k, what, j);
                if (Q->gwidth[b.slot] != width) return fail(NGB200_ERR_INVALID, "phase %d %s %d: width %d != global operand width %d", k, what, j, width, Q->gwidth[b.slot]);
                if (Q->gdomain[b.slot] < 0 && (is_out || b.index >= 0)) return fail(NGB200_ERR_INVALID, "phase %d %s %d: a broadcast operand is read-only and not indexed", k, what, j);
                if (Q->gdomain[b.slot] >= 0 && b.index < 0 && Q->gdomain[b.slot] != pd.domain) return fail(NGB200_ERR_INVALID, "phase %d %s %d: direct access to an array of another domain", k, what, j);
            } else return fail(NGB200_ERR_INVALID, "phase %d %s %d: bad binding kind", k, what, j);
            if (b.index >= 0) {
                if (b.index >= d->n_globals || !Q->gindex[b.index] || Q->gdomain[b.index] != pd.domain)
                    return fail(NGB200_ERR_INVALID, "phase %d %s %d: index operand must be an index array over the phase's domain", k, what, j);
            }
            return NGB200_OK;
        };
        for (int j = 0; j < pp.n_inputs; ++j) {
            int rc = check(ph.in[j], P.in_width[j], false, j);
            if (rc) return rc;
            const ngb200_bind &b = ph.in[j];
            P.in_mode.push_back(b.kind == NGB200_BIND_GLOBAL && Q->gdomain[b.slot] < 0 ? NGB200_BROADCAST : NGB200_VARYING);
        }
        for (int j = 0; j < pp.n_outputs; ++j) {
            int rc = check(ph.out[j], P.out_width[j], true, j);
            if (rc) return rc;
        }
        Q->phases.push_back(std::move(ph));
    }
    const int forced = env_int("NGB200_IR_ORDER", -1);
    std::string err;
    Q->source = pipeline_source(*Q, forced != 0, &err);
    if (Q->source.empty()) return fail(NGB200_ERR_INVALID, "%s", err.c_str());
    Q->hash = hash_hex(Q->source + "|sm_100a|pipeline");
    std::lock_guard<std::mutex> lock(g_mu);
    const bool use_cache = !env_int("NGB200_NO_CACHE", 0);
    std::string base = cache_dir() + "/" + Q->hash;
    std::vector<char> marker;
    if (use_cache && forced < 0 && read_file(base + ".alt", marker)) {       // the other emission order won last time
        Q->source = pipeline_source(*Q, false, &err);
        Q->hash = hash_hex(Q->source + "|sm_100a|pipeline");
        base = cache_dir() + "/" + Q->hash;
    }
    if (use_cache && read_file(base + ".cubin", Q->cubin)) {
        Q->cache_hit = true;
        *out = Q.release();
        return NGB200_OK;
    }
    std::string log;
    int rc = nvrtc_compile(Q->source, Q->cubin, &log);
    if (rc) return rc;
    Q->spill = spill_bytes(log, "ngb_pipeline");
    if (forced < 0 && Q->spill > 0 && marker.empty()) {
        std::string src_ir = Q->source, hash_ir = Q->hash;
        std::vector<char> cubin_ir = Q->cubin;
        const int minb_ir = Q->minb;
        std::string alt = pipeline_source(*Q, false, &err);
        std::vector<char> cubin_alt;
        std::string log_alt;
        if (!alt.empty() && nvrtc_compile(alt, cubin_alt, &log_alt) == NGB200_OK && spill_bytes(log_alt, "ngb_pipeline") < Q->spill) {
            Q->source = alt;
            Q->cubin = cubin_alt;
            Q->spill = spill_bytes(log_alt, "ngb_pipeline");
            Q->hash = hash_hex(Q->source + "|sm_100a|pipeline");
            write_file((cache_dir() + "/" + hash_ir + ".alt").c_str(), Q->hash.data(), Q->hash.size());
            base = cache_dir() + "/" + Q->hash;
        } else {
            Q->source = src_ir;
            Q->cubin = cubin_ir;
            Q->minb = minb_ir;
        }
    }
    write_file(base + ".cubin", Q->cubin.data(), Q->cubin.size());
    if (env_int("NGB200_CACHE_SOURCES", 0)) write_file(base + ".cu", Q->source.data(), Q->source.size());
    *out = Q.release();
    return NGB200_OK;
}

extern "C" void ngb200_pipeline_destroy(ngb200_pipeline *Q) {
    if (!Q) return;
    if (g_drv.ok) {
        for (int d = 0; d < kMaxDevices; ++d) {
            PipeLoaded &L = Q->loaded[d];
            if (!L.ready.load() || !L.mod) continue;
            CUcontext prev = nullptr, ctx = nullptr;
            CUdevice dev;
            if (g_drv.p_cuCtxGetCurrent(&prev) != CUDA_SUCCESS) continue;
            if (g_drv.p_cuDeviceGet(&dev, d) != CUDA_SUCCESS || g_drv.p_cuDevicePrimaryCtxRetain(&ctx, dev) != CUDA_SUCCESS) continue;
            g_drv.p_cuCtxSetCurrent(ctx);
            g_drv.p_cuModuleUnload(L.mod);
            g_drv.p_cuCtxSetCurrent(prev);
        }
    }
    delete Q;
}

extern "C" const char *ngb200_pipeline_source(const ngb200_pipeline *Q) { return Q ? Q->source.c_str() : ""; }

extern "C" int ngb200_pipeline_get_info(const ngb200_pipeline *Q, ngb200_pipeline_info *info) {
    if (!Q || !info) return fail(NGB200_ERR_INVALID, "null argument");
    memset(info, 0, sizeof *info);
    info->n_instr = Q->n_live;
    info->threads_per_tile = Q->tpb;
    info->min_blocks = Q->minb;
    info->shared_bytes = (int64_t)Q->smem;
    info->cache_hit = Q->cache_hit;
    info->spill_bytes = (int64_t)Q->spill;
    for (int d = 0; d < kMaxDevices; ++d)
        if (Q->loaded[d].ready.load()) { info->regs = Q->loaded[d].regs; info->grid = Q->loaded[d].grid; info->local_bytes = (int64_t)Q->loaded[d].local_bytes; break; }
    return NGB200_OK;
}

extern "C" int ngb200_pipeline_launch(ngb200_pipeline *Q, const ngb200_operand *globals, const int64_t *domain_items,
                                      int64_t n_tiles, int32_t repeat, const double *params, void *stream) {
    if (!Q || !globals || !domain_items) return fail(NGB200_ERR_INVALID, "null argument");
    if (n_tiles < 0 || repeat < 0) return fail(NGB200_ERR_INVALID, "negative tile / repeat count");
    if (Q->n_params > 0 && !params) return fail(NGB200_ERR_INVALID, "pipeline needs %d scalar params", Q->n_params);
    int rc = need_driver();
    if (rc) return rc;
    if (n_tiles == 0) return NGB200_OK;
    int dev = 0;
    rc = current_device(&dev, -1);
    if (rc) return rc;
    if (dev < 0 || dev >= kMaxDevices) return fail(NGB200_ERR_CUDA, "device ordinal %d not supported", dev);
    PipeLoaded &L = Q->loaded[dev];
    if (!L.ready.load(std::memory_order_acquire)) {
        std::lock_guard<std::mutex> lock(Q->mu);
        if (!L.ready.load()) {
            CU(cuModuleLoadData(&L.mod, Q->cubin.data()));
            CU(cuModuleGetFunction(&L.f, L.mod, "ngb_pipeline"));
            CU(cuFuncGetAttribute(&L.regs, CU_FUNC_ATTRIBUTE_NUM_REGS, L.f));
            int lb = 0;
            CU(cuFuncGetAttribute(&lb, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, L.f));
            L.local_bytes = (size_t)lb;
            if (Q->smem > 48 * 1024) CU(cuFuncSetAttribute(L.f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)Q->smem));
            int sms = 0, nb = 0;
            CU(cuDeviceGetAttribute(&sms, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev));
            CU(cuOccupancyMaxActiveBlocksPerMultiprocessor(&nb, L.f, Q->tpb * Q->warps, Q->smem));
            L.grid = sms * (nb > 0 ? nb : 1);
            L.ready.store(1, std::memory_order_release);
        }
    }
    const size_t ng = Q->gwidth.size();
    for (size_t k = 0; k < ng; ++k) {
        if (!globals[k].ptr) return fail(NGB200_ERR_INVALID, "global operand %zu: null pointer", k);
        if (!Q->gindex[k] && Q->gdomain[k] >= 0 && globals[k].batch_stride != 1)
            return fail(NGB200_ERR_INVALID, "global operand %zu: pipelines need structure-of-arrays operands (batch stride 1)", k);
    }
    alignas(16) char buf[4096];
    memset(buf, 0, Q->param_bytes);
    char *w = buf;
    struct PG { void *ptr; long long cs; };
    for (size_t k = 0; k < ng; ++k)
        if (!Q->gindex[k]) { PG g{globals[k].ptr, globals[k].comp_stride}; memcpy(w + 16 * Q->gslot[k], &g, 16); }
    w += 16 * (size_t)std::max(1, Q->n_value_globals);
    for (size_t k = 0; k < ng; ++k)
        if (Q->gindex[k]) memcpy(w + 8 * Q->gslot[k], &globals[k].ptr, 8);
    w += 8 * (size_t)std::max(1, Q->n_index_globals);
    for (int k = 0; k < Q->n_params; ++k) memcpy(w + 8 * k, &params[k], 8);
    w += 8 * (size_t)std::max(1, Q->n_params);
    for (size_t k = 0; k < Q->ipt.size(); ++k) { long long v = domain_items[k]; memcpy(w + 8 * k, &v, 8); }
    w += 8 * Q->ipt.size();
    long long nt = n_tiles;
    memcpy(w, &nt, 8);
    w += 8;
    int rp = repeat;
    memcpy(w, &rp, 4);
    const int64_t blocks = std::min<int64_t>((n_tiles + Q->warps - 1) / Q->warps, L.grid);
    void *args[] = {buf};
    CU(cuLaunchKernel(L.f, (unsigned)blocks, 1, 1, (unsigned)(Q->tpb * Q->warps), 1, 1, (unsigned)Q->smem, (CUstream)stream, args, nullptr));
    g_launches.fetch_add(1);
    return NGB200_OK;
}

// ------------------------------------------------------------------------------------------
// host-buffer execution: chunked H2D -> kernel -> D2H over several streams
// ------------------------------------------------------------------------------------------
extern "C" int ngb200_host_alloc(void **ptr, int64_t bytes) {
    if (!ptr || bytes < 0) return fail(NGB200_ERR_INVALID, "bad argument");
    int rc = need_driver();
    if (rc) return rc;
    int dev;
    rc = current_device(&dev, -1);
    if (rc) return rc;
    CU(cuMemAllocHost(ptr, (size_t)(bytes > 0 ? bytes : 1)));
    return NGB200_OK;
}

extern "C" int ngb200_host_free(void *ptr) {
    int rc = need_driver();
    if (rc) return rc;
    CU(cuMemFreeHost(ptr));
    return NGB200_OK;
}

// Staging resources of the host path, kept with the program between calls (allocating / freeing device memory
// on every call serialises the device and made the pipeline's throughput erratic): per slot one stream and one
// device chunk buffer per operand, plus the device copies of broadcast operands.
struct HostPipe {
    int device = -1;
    int64_t chunk_items = 0;
    struct Slot {
        CUstream stream = nullptr;
        std::vector<CUdeviceptr> in, out;
    };
    std::vector<Slot> slots;
    std::vector<CUdeviceptr> bcast;
    void release() {
        for (Slot &s : slots) {
            for (CUdeviceptr p : s.in) if (p) g_drv.p_cuMemFree(p);
            for (CUdeviceptr p : s.out) if (p) g_drv.p_cuMemFree(p);
            if (s.stream) g_drv.p_cuStreamDestroy(s.stream);
        }
        for (CUdeviceptr p : bcast) if (p) g_drv.p_cuMemFree(p);
        slots.clear();
        bcast.clear();
        chunk_items = 0;
        device = -1;
    }
};

static void host_pipe_destroy(HostPipe *h) {
    if (!h) return;
    if (g_drv.ok) h->release();
    delete h;
}

// copy `items` batch items of one operand between host layout and a device SoA/AoS chunk buffer
static int copy_chunk(bool to_device, const ngb200_operand &h, int width, int es, int64_t first, int64_t items,
                      CUdeviceptr d, int64_t d_cs, CUstream s) {
    if (width == 0 || items == 0) return NGB200_OK;
    if (h.batch_stride == 1) {  // SoA on the host: one contiguous row per component -> one 1-D DMA per row
        for (int c = 0; c < width; ++c) {
            char *hp = (char *)h.ptr + ((int64_t)c * h.comp_stride + first) * es;
            CUdeviceptr dp = d + (size_t)c * d_cs * es;
            if (to_device) CU(cuMemcpyHtoDAsync(dp, hp, (size_t)items * es, s));
            else CU(cuMemcpyDtoHAsync(hp, dp, (size_t)items * es, s));
        }
        return NGB200_OK;
    }
    if (h.comp_stride == 1 && h.batch_stride == width) {  // AoS on the host: one contiguous block
        char *hp = (char *)h.ptr + first * width * es;
        if (to_device) CU(cuMemcpyHtoDAsync(d, hp, (size_t)items * width * es, s));
        else CU(cuMemcpyDtoHAsync(hp, d, (size_t)items * width * es, s));
        return NGB200_OK;
    }
    return fail(NGB200_ERR_INVALID, "host operand must be SoA (batch_stride 1) or packed AoS (comp_stride 1, batch_stride width)");
}

// (re)build the staging resources for one configuration; on ANY failure everything allocated so far is released
// and the pipe is left empty, so a later call starts from scratch instead of running on null buffers
static int host_pipe_configure(ngb200_program *P, HostPipe &run, int dev, int64_t chunk_items, int n_slots) {
    if (run.device == dev && run.chunk_items == chunk_items && (int)run.slots.size() == n_slots) return NGB200_OK;
    run.release();
    const int es = esize(P->dtype);
    const size_t nin = P->in_width.size(), nout = P->out_width.size();
    HostPipe fresh;
    fresh.slots.resize(n_slots);
    fresh.bcast.assign(nin, 0);
    auto build = [&]() -> int {
        for (HostPipe::Slot &s : fresh.slots) {
            s.in.assign(nin, 0);
            s.out.assign(nout, 0);
            CU(cuStreamCreate(&s.stream, CU_STREAM_NON_BLOCKING));
            for (size_t k = 0; k < nin; ++k)
                if (P->in_mode[k] == NGB200_VARYING && P->in_width[k]) CU(cuMemAlloc(&s.in[k], (size_t)chunk_items * P->in_width[k] * es));
            for (size_t k = 0; k < nout; ++k)
                if (P->out_width[k]) CU(cuMemAlloc(&s.out[k], (size_t)chunk_items * P->out_width[k] * es));
        }
        for (size_t k = 0; k < nin; ++k)
            if (P->in_mode[k] == NGB200_BROADCAST && P->in_width[k]) CU(cuMemAlloc(&fresh.bcast[k], (size_t)P->in_width[k] * es));
        return NGB200_OK;
    };
    int rc = build();
    if (rc) { fresh.release(); return rc; }
    fresh.device = dev;
    fresh.chunk_items = chunk_items;
    run = fresh;                      // plain handles: the pipe owns them from here on
    fresh.slots.clear();
    fresh.bcast.clear();
    return NGB200_OK;
}

extern "C" int ngb200_program_run_host(ngb200_program *P, const ngb200_operand *in, const ngb200_operand *out,
                                       const double *params, int64_t batch, int32_t device, int64_t chunk_items) {
    if (!P || !out || (!in && !P->in_width.empty())) return fail(NGB200_ERR_INVALID, "null argument");
    if (P->any_indexed) return fail(NGB200_ERR_INVALID, "run_host does not support indexed / tiled operands");
    if (P->any_reduced) return fail(NGB200_ERR_INVALID, "run_host does not support NGB200_REDUCED outputs (batch sums need one device accumulator across chunks)");
    if (P->n_params > 0 && !params) return fail(NGB200_ERR_INVALID, "program needs %d scalar params", P->n_params);
    int rc = need_driver();
    if (rc) return rc;
    CtxGuard restore_context;
    int dev;
    rc = current_device(&dev, device);
    if (rc) return rc;
    if (batch <= 0) return NGB200_OK;
    const int es = esize(P->dtype);
    const size_t nin = P->in_width.size(), nout = P->out_width.size();
    int64_t per_item = 0;
    for (size_t k = 0; k < nin; ++k) if (P->in_mode[k] == NGB200_VARYING) per_item += P->in_width[k] * es;
    for (size_t k = 0; k < nout; ++k) per_item += P->out_width[k] * es;
    if (chunk_items <= 0) chunk_items = (int64_t)env_int("NGB200_HOST_CHUNK_MB", 128) * (1 << 20) / (per_item ? per_item : 1);
    chunk_items = (chunk_items + 255) / 256 * 256;
    if (chunk_items > batch) chunk_items = (batch + 255) / 256 * 256;
    const int n_slots = (int)std::min<int64_t>(env_int("NGB200_HOST_STREAMS", 8), (batch + chunk_items - 1) / chunk_items);

    std::lock_guard<std::mutex> pipe_lock(P->host_mu);      // one host run per program at a time (shared staging)
    if (!P->host) P->host = new HostPipe;
    HostPipe &run = *P->host;
    rc = host_pipe_configure(P, run, dev, chunk_items, n_slots);
    if (rc) return rc;
    // From here on copies are in flight on the caller's host buffers: every exit path drains all slot streams first.
    auto body = [&]() -> int {
        // broadcast operands: one small upload, visible to all streams after a sync
        for (size_t k = 0; k < nin; ++k) {
            if (P->in_mode[k] != NGB200_BROADCAST || !P->in_width[k]) continue;
            std::vector<char> packed((size_t)P->in_width[k] * es);
            for (int c = 0; c < P->in_width[k]; ++c) memcpy(&packed[(size_t)c * es], (char *)in[k].ptr + (size_t)c * in[k].comp_stride * es, es);
            CU(cuMemcpyHtoDAsync(run.bcast[k], packed.data(), packed.size(), run.slots[0].stream));
            CU(cuStreamSynchronize(run.slots[0].stream));
        }
        std::vector<ngb200_operand> din(nin ? nin : 1), dout(nout);
        int64_t chunk_index = 0;
        for (int64_t first = 0; first < batch; first += chunk_items, ++chunk_index) {
            HostPipe::Slot &s = run.slots[chunk_index % n_slots];
            int64_t items = std::min(chunk_items, batch - first);
            for (size_t k = 0; k < nin; ++k) {
                if (P->in_mode[k] == NGB200_BROADCAST) { din[k] = {(void *)run.bcast[k], 1, 0, nullptr}; continue; }
                bool soa = in[k].batch_stride == 1;
                int r = copy_chunk(true, in[k], P->in_width[k], es, first, items, s.in[k], chunk_items, s.stream);
                if (r) return r;
                din[k] = {(void *)s.in[k], soa ? chunk_items : 1, soa ? 1 : P->in_width[k], nullptr};
            }
            for (size_t k = 0; k < nout; ++k) {
                bool soa = out[k].batch_stride == 1;
                dout[k] = {(void *)s.out[k], soa ? chunk_items : 1, soa ? 1 : P->out_width[k], nullptr};
            }
            int r = ngb200_program_launch(P, din.data(), dout.data(), params, items, s.stream);
            if (r) return r;
            for (size_t k = 0; k < nout; ++k) {
                r = copy_chunk(false, out[k], P->out_width[k], es, first, items, s.out[k], chunk_items, s.stream);
                if (r) return r;
            }
        }
        return NGB200_OK;
    };
    rc = body();
    const std::string first_error = rc ? g_err : std::string();
    int sync_rc = NGB200_OK;
    for (HostPipe::Slot &s : run.slots) {
        CUresult r = g_drv.p_cuStreamSynchronize(s.stream);
        if (r != CUDA_SUCCESS && !sync_rc) sync_rc = cu_fail(r, "cuStreamSynchronize");
    }
    if (rc) { g_err = first_error; return rc; }
    return sync_rc;
}

// Raw-copy control of the host path (bench.py `e2e.host_ceiling_points_s`): the chunked, multi-stream H2D of
// `in_bytes` and D2H of `out_bytes` that ngb200_program_run_host performs, WITHOUT a kernel in between.  What it
// measures is the ceiling the pinned-memory / PCIe side of the box puts on any host-buffer pipeline.
extern "C" int ngb200_host_copy_control(void *host_in, int64_t in_bytes, void *host_out, int64_t out_bytes,
                                        int64_t chunk_bytes, int32_t streams, int32_t device) {
    if ((in_bytes > 0 && !host_in) || (out_bytes > 0 && !host_out) || in_bytes < 0 || out_bytes < 0)
        return fail(NGB200_ERR_INVALID, "bad argument");
    int rc = need_driver();
    if (rc) return rc;
    CtxGuard restore_context;
    int dev;
    rc = current_device(&dev, device);
    if (rc) return rc;
    if (chunk_bytes <= 0) chunk_bytes = (int64_t)env_int("NGB200_HOST_CHUNK_MB", 128) * (1 << 20) / 2;   // in + out = one pipeline chunk
    if (streams <= 0) streams = env_int("NGB200_HOST_STREAMS", 8);
    struct Slot { CUstream s = nullptr; CUdeviceptr din = 0, dout = 0; };
    static std::mutex mu;
    static std::vector<Slot> slots;
    static int64_t slot_bytes = 0;
    static int slot_dev = -1;
    std::lock_guard<std::mutex> lock(mu);
    auto release = [&]() {
        for (Slot &q : slots) {
            if (q.din) g_drv.p_cuMemFree(q.din);
            if (q.dout) g_drv.p_cuMemFree(q.dout);
            if (q.s) g_drv.p_cuStreamDestroy(q.s);
        }
        slots.clear();
        slot_bytes = 0;
        slot_dev = -1;
    };
    if (slot_dev != dev || slot_bytes != chunk_bytes || (int)slots.size() != streams) {
        release();
        slots.resize(streams);
        auto build = [&]() -> int {
            for (Slot &q : slots) {
                CU(cuStreamCreate(&q.s, CU_STREAM_NON_BLOCKING));
                CU(cuMemAlloc(&q.din, (size_t)chunk_bytes));
                CU(cuMemAlloc(&q.dout, (size_t)chunk_bytes));
            }
            return NGB200_OK;
        };
        rc = build();
        if (rc) { release(); return rc; }
        slot_bytes = chunk_bytes;
        slot_dev = dev;
    }
    auto body = [&]() -> int {
        const int64_t total = std::max(in_bytes, out_bytes);
        int64_t k = 0;
        for (int64_t off = 0; off < total; off += chunk_bytes, ++k) {
            Slot &q = slots[k % streams];
            if (off < in_bytes) CU(cuMemcpyHtoDAsync(q.din, (char *)host_in + off, (size_t)std::min(chunk_bytes, in_bytes - off), q.s));
            if (off < out_bytes) CU(cuMemcpyDtoHAsync((char *)host_out + off, q.dout, (size_t)std::min(chunk_bytes, out_bytes - off), q.s));
        }
        return NGB200_OK;
    };
    rc = body();
    const std::string first_error = rc ? g_err : std::string();
    for (Slot &q : slots) g_drv.p_cuStreamSynchronize(q.s);
    if (rc) g_err = first_error;
    return rc;
}

// ------------------------------------------------------------------------------------------
// statically compiled helper kernel: counter-based normal generator for synthetic benchmarks
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&r)[4]) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    r[0] = c0; r[1] = c1; r[2] = c2; r[3] = c3;
}

template <typename T>
__global__ void __launch_bounds__(256) ngb_fill_normal_kernel(T *__restrict__ out, long long n, unsigned long long seed,
                                                              unsigned long long offset, float scale) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < n; q += stride) {
        unsigned long long ctr = offset / 4 + (unsigned long long)q;
        uint32_t r[4];
        philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
        float v[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float u1 = ((float)r[2 * h] + 1.0f) * 2.3283064365386963e-10f;   // (0, 1]
            float u2 = (float)r[2 * h + 1] * 2.3283064365386963e-10f;
            float rad = sqrtf(-2.0f * __logf(u1));
            float sn, cs;
            __sincosf(6.283185307179586f * u2, &sn, &cs);
            v[2 * h] = rad * cs * scale;
            v[2 * h + 1] = rad * sn * scale;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (q * 4 + j < n) out[q * 4 + j] = (T)v[j];
    }
}

extern "C" int ngb200_fill_normal(void *ptr, int64_t n, uint64_t seed, uint64_t offset, double scale, int32_t dtype,
                                  void *stream) {
    if (!ptr || n < 0 || offset % 4) return fail(NGB200_ERR_INVALID, "bad argument (offset must be a multiple of 4)");
    if (n == 0) return NGB200_OK;
    int64_t quads = (n + 3) / 4;
    int blocks = (int)std::min<int64_t>((quads + 255) / 256, 148 * 8);
    if (dtype == NGB200_F64)
        ngb_fill_normal_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>((double *)ptr, n, seed, offset, (float)scale);
    else
        ngb_fill_normal_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((float *)ptr, n, seed, offset, (float)scale);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(NGB200_ERR_CUDA, "fill_normal launch failed: %s", cudaGetErrorString(e));
    g_launches.fetch_add(1);
    return NGB200_OK;
}

extern "C" const char *ngb200_version(void) {
    static std::string v;
    if (v.empty()) {
        int maj = 0, min = 0;
        nvrtcVersion(&maj, &min);
        v = std::string("numga_b200 " NGB_VERSION " sm_100a nvrtc ") + std::to_string(maj) + "." + std::to_string(min);
    }
    return v.c_str();
}
