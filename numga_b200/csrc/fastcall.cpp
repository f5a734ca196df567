// numga_b200 fast eager path: one C++ call per operator application.
//
// The steady state of eager code is "same compiled program, same operand layout, fresh tensors": XOperator.__call__
// allocates its result on every call (numga/backend/numpy/operator.py:40-49,117-132).  In Python that costs a layout
// signature, a dict lookup, torch.empty_strided per result, ctypes marshalling and the C launch: 15-22 us per call on
// the GPU box, more than the 16 us kernel of a 2^20-item product.  Here the whole sequence -- layout check against the
// plan, result allocation from torch's caching allocator, current stream, ngb200_plan_launch -- is one function.
// torch is used for what it is in this project: device memory and streams.  The kernels are launched through the
// C ABI of libnumga_b200.so (function pointers handed over at start-up; no link-time dependency).
#include <ATen/ATen.h>
#include <ATen/core/grad_mode.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/extension.h>

#include <cstdint>
#include <vector>

namespace py = pybind11;

typedef int (*plan_launch_fn)(void *plan, void *const *in_ptrs, void *const *out_ptrs, const double *params, void *stream);
typedef const char *(*last_error_fn)();

static plan_launch_fn g_plan_launch = nullptr;
static last_error_fn g_last_error = nullptr;

struct FastPlan {
    void *plan = nullptr;                         // ngb200_plan* (owned by the python side)
    int device = 0;
    int n_params = 0;
    at::ScalarType dtype = at::kFloat;
    std::vector<std::vector<int64_t>> in_sizes, in_strides, out_sizes, out_strides;
};

static void bind(uint64_t plan_launch, uint64_t last_error) {
    g_plan_launch = reinterpret_cast<plan_launch_fn>(plan_launch);
    g_last_error = reinterpret_cast<last_error_fn>(last_error);
}

static FastPlan *make_plan(uint64_t plan, int device, int n_params, bool f64, std::vector<std::vector<int64_t>> in_sizes,
                           std::vector<std::vector<int64_t>> in_strides, std::vector<std::vector<int64_t>> out_sizes,
                           std::vector<std::vector<int64_t>> out_strides) {
    FastPlan *p = new FastPlan;
    p->plan = reinterpret_cast<void *>(plan);
    p->device = device;
    p->n_params = n_params;
    p->dtype = f64 ? at::kDouble : at::kFloat;
    p->in_sizes = std::move(in_sizes);
    p->in_strides = std::move(in_strides);
    p->out_sizes = std::move(out_sizes);
    p->out_strides = std::move(out_strides);
    return p;
}

// Returns the result tensors, or None when an input does not have the layout the plan was made for (the caller then
// takes the general path).  `params` holds the scalar launch parameters (may be empty).
static py::object launch(FastPlan &p, const std::vector<at::Tensor> &inputs, const std::vector<double> &params) {
    const size_t nin = p.in_sizes.size(), nout = p.out_sizes.size();
    if (inputs.size() != nin || (int)params.size() != p.n_params) return py::none();
    void *in_ptrs[32];
    void *out_ptrs[32];
    if (nin > 32 || nout > 32) return py::none();
    const bool grad_mode = at::GradMode::is_enabled();
    for (size_t k = 0; k < nin; ++k) {
        const at::Tensor &t = inputs[k];
        if (t.scalar_type() != p.dtype || !t.is_cuda() || t.get_device() != p.device) return py::none();
        if (grad_mode && t.requires_grad()) return py::none();        // differentiable call: the python path builds the autograd node
        const auto sz = t.sizes(), st = t.strides();
        const auto &es = p.in_sizes[k], &et = p.in_strides[k];
        if (sz.size() != es.size()) return py::none();
        for (size_t d = 0; d < es.size(); ++d)
            if (sz[d] != es[d] || (es[d] > 1 && st[d] != et[d])) return py::none();
        in_ptrs[k] = t.data_ptr();
    }
    const c10::cuda::CUDAGuard device_guard((c10::DeviceIndex)p.device);   // the plan's device is current while we allocate and launch
    std::vector<at::Tensor> outs;
    outs.reserve(nout);
    const auto opts = at::TensorOptions().dtype(p.dtype).device(at::kCUDA, p.device);
    for (size_t k = 0; k < nout; ++k) {
        outs.push_back(at::empty_strided(p.out_sizes[k], p.out_strides[k], opts));
        out_ptrs[k] = outs.back().data_ptr();
    }
    void *stream = c10::cuda::getCurrentCUDAStream(p.device).stream();
    const int rc = g_plan_launch(p.plan, in_ptrs, out_ptrs, params.empty() ? nullptr : params.data(), stream);
    if (rc != 0) throw std::runtime_error(std::string("numga_b200 error ") + std::to_string(rc) + ": " + (g_last_error ? g_last_error() : "?"));
    return py::cast(outs);
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "numga_b200 fast eager launch path";
    py::class_<FastPlan>(m, "FastPlan");
    m.def("bind", &bind);
    m.def("make_plan", &make_plan, py::return_value_policy::take_ownership);
    m.def("launch", &launch);
}
