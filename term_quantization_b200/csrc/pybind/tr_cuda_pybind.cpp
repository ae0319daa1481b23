// tr_cuda_pybind.cpp -- the reference's extension module, re-bound to libtq_b200.so.
//
// Replaces kernels/tr_cuda.cpp:1-28 (pybind11 module `tr_cuda`, one function `tr`) for callers that want a
// compiled torch extension instead of the ctypes binding (term_quantization_b200/tr_cuda.py): same signature
//     tr(Tensor input, float sf, int bitwidth, int group_size, int num_keep_terms) -> Tensor
// same precondition messages (kernels/tr_cuda.cpp:12-18), a NEW output tensor of the input's shape / dtype /
// device (kernels/tr_cuda_kernel.cu:145), no autograd.  Differences, all additive: the launch goes to the
// CURRENT stream of the input's device (the reference uses the legacy default stream and no device guard,
// kernels/tr_cuda_kernel.cu:147), bf16 / fp16 inputs are accepted, and a failed launch raises.
// Built ahead of time by __graft_entry__.build() (csrc/pybind/build.py); nothing is compiled at import time.
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/extension.h>

#include "tq_b200.h"

#define CHECK_CUDA(x) TORCH_CHECK(x.is_cuda(), #x " must be a CUDA tensor")
#define CHECK_CONTIGUOUS(x) TORCH_CHECK(x.is_contiguous(), #x " must be contiguous")
#define CHECK_INPUT(x) \
    CHECK_CUDA(x);     \
    CHECK_CONTIGUOUS(x)

static int tq_dtype(const at::Tensor &t)
{
    switch (t.scalar_type()) {
    case at::kFloat: return TQ_F32;
    case at::kDouble: return TQ_F64;
    case at::kBFloat16: return TQ_BF16;
    case at::kHalf: return TQ_F16;
    default: TORCH_CHECK(false, "\"tr_cuda\" not implemented for '", t.scalar_type(), "'");
    }
}

static at::Tensor tr_impl(const at::Tensor &input, float sf, int32_t bitwidth, int32_t group_size, int32_t num_keep_terms,
                          int encoding)
{
    CHECK_INPUT(input);
    TORCH_CHECK(input.dim() >= 2, "tr expects a tensor with at least 2 dimensions (B, C, ...)");
    const int64_t B = input.size(0), C = input.size(1);          // kernels/tr_cuda_kernel.cu:133-141
    int64_t WH = 1;
    for (int64_t d = 2; d < input.dim(); ++d) WH *= input.size(d);
    const c10::cuda::CUDAGuard guard(input.device());
    at::Tensor output = at::empty_like(input);
    const int rc = tq_tr_encode(input.data_ptr(), output.data_ptr(), tq_dtype(input), B, C, WH, sf, bitwidth, group_size,
                                num_keep_terms, encoding, 0u, c10::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == TQ_OK, "tq_b200: ", tq_last_error());
    return output;
}

at::Tensor tr(const at::Tensor input, const float sf, const int32_t bitwidth, const int32_t group_size,
              const int32_t num_keep_terms)
{
    return tr_impl(input, sf, bitwidth, group_size, num_keep_terms, TQ_ENC_HESE);
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m)
{
    m.def("tr", &tr, "Term Revealing (TR) (CUDA)");
    m.def("tr_encoding", &tr_impl, "TR with an explicit term encoding (0 hese, 1 binary, 2 booth)");
    m.def("version", &tq_version, "C ABI version of libtq_b200.so");
}
