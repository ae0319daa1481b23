"""ctypes front-end of the CPU checkers under oracle/.

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of bench.py.  The product package
(``term_quantization_b200``) never imports this module.

Two libraries:

* ``libtq_oracle.so`` -- our restatement (oracle/tq_oracle.c), always available.
* ``_ref/libtq_ref.so`` -- the reference kernel body (kernels/tr_cuda_kernel.cu:13-126)
  compiled for the host by oracle/Makefile where /root/reference exists; ships to
  the GPU box prebuilt.  ``have_ref()`` says whether it is loadable.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ENC_HESE, ENC_BINARY, ENC_BOOTH = 0, 1, 2
_lib = None
_ref = None


def build():
    """(Re)build the checkers; never needs a GPU."""
    subprocess.run(["make", "-C", _HERE, "--no-print-directory"], check=True,
                   stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libtq_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.tqo_quantize_f32.restype = C.c_int32
        L.tqo_quantize_f32.argtypes = [C.c_float, C.c_float, C.c_int]
        L.tqo_quantize_f64.restype = C.c_int32
        L.tqo_quantize_f64.argtypes = [C.c_double, C.c_float, C.c_int]
        L.tqo_terms.restype = None
        L.tqo_terms.argtypes = [C.c_uint32, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.tqo_tr.restype = C.c_int
        L.tqo_tr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64,
                             C.c_int64, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.tqo_hist_f32.restype = None
        L.tqo_hist_f32.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_float, C.c_float]
        L.tqo_mse_profile.restype = C.c_int
        L.tqo_mse_profile.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                      C.c_int, C.c_int, C.c_void_p]
        L.tqo_hese_term_count_f32.restype = C.c_int64
        L.tqo_hese_term_count_f32.argtypes = [C.c_void_p, C.c_int64, C.c_float]
        L.tqo_gemm_i32.restype = None
        L.tqo_gemm_i32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64]
        L.tqo_fma_channels_f32.restype = None
        L.tqo_fma_channels_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        _lib = L
    return _lib


def have_ref():
    return os.path.exists(os.path.join(_HERE, "_ref", "libtq_ref.so"))


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(os.path.join(_HERE, "_ref", "libtq_ref.so"))
        for name, ct in (("tq_ref_tr_f32", C.c_float), ("tq_ref_tr_f64", C.c_double)):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_float] + [C.c_int] * 7
        L.tq_ref_hese_terms_f32.restype = C.c_int
        L.tq_ref_hese_terms_f32.argtypes = [C.c_float, C.c_float, C.c_int, C.c_void_p]
        L.tq_ref_set_threads.argtypes = [C.c_int]
        L.tq_ref_get_threads.restype = C.c_int
        _ref = L
    return _ref


def _dims(shape):
    """(B, C, WH) exactly as the reference launcher reads them
    (kernels/tr_cuda_kernel.cu:133-141): size(0), size(1), size(2)*size(3) if 4-D."""
    if len(shape) == 4:
        return shape[0], shape[1], shape[2] * shape[3]
    if len(shape) == 2:
        return shape[0], shape[1], 1
    raise ValueError("oracle expects 2-D or 4-D arrays")


def tr(x, sf, bits, g, alpha, encoding=ENC_HESE, relu=False, return_codes=False):
    """Our restatement of tr_cuda.tr on a numpy array (float32/float64, 2-D or 4-D)."""
    x = np.ascontiguousarray(x)
    assert x.dtype in (np.float32, np.float64)
    B, Cc, WH = _dims(x.shape)
    out = np.empty_like(x)
    codes = np.empty(x.shape, dtype=np.int32)
    rc = lib().tqo_tr(x.ctypes.data, out.ctypes.data, codes.ctypes.data,
                      0 if x.dtype == np.float32 else 1, B, Cc, WH,
                      float(np.float32(sf)), bits, g, alpha, encoding, int(relu))
    if rc != 0:
        raise ValueError("tqo_tr: invalid arguments")
    return (out, codes) if return_codes else out


def ref_tr(x, sf, bits, g, alpha, threads=0):
    """The REFERENCE kernel body on the host (needs oracle/_ref)."""
    x = np.ascontiguousarray(x)
    assert x.dtype in (np.float32, np.float64)
    if x.ndim == 4:
        B, Cc, W, H = x.shape
    else:
        (B, Cc), W, H = x.shape, 1, 1
    out = np.empty_like(x)
    R = ref()
    R.tq_ref_set_threads(threads)
    fn = R.tq_ref_tr_f32 if x.dtype == np.float32 else R.tq_ref_tr_f64
    rc = fn(x.ctypes.data, out.ctypes.data, float(np.float32(sf)), bits, g, alpha, B, Cc, W, H)
    if rc != 0:
        raise ValueError("reference kernel is undefined for these arguments (C % g != 0, g > 32)")
    return out


_ref_gpu = None


def have_ref_gpu():
    """oracle/_ref/libtq_ref_gpu.so: the reference kernel body compiled for sm_100a (ref_gpu_shim.cu)."""
    return os.path.exists(os.path.join(_HERE, "_ref", "libtq_ref_gpu.so"))


def ref_gpu_tr(x, sf, bits, g, alpha, out=None):
    """The REFERENCE kernel (kernels/tr_cuda_kernel.cu:58-125, byte-identical body) launched on the GPU exactly as
    the reference's launcher does (zeros_like + <<<ceil(n/128), 128>>>), on a torch CUDA tensor, on the current
    stream.  The GPU-side oracle and "the kernel to beat"."""
    import torch
    global _ref_gpu
    if _ref_gpu is None:
        L = C.CDLL(os.path.join(_HERE, "_ref", "libtq_ref_gpu.so"))
        for name in ("tq_refgpu_tr_f32", "tq_refgpu_tr_f64"):
            fn = getattr(L, name)
            fn.restype = C.c_int
            fn.argtypes = [C.c_void_p, C.c_void_p, C.c_float] + [C.c_int] * 7 + [C.c_void_p]
        _ref_gpu = L
    assert x.is_cuda and x.is_contiguous() and x.dtype in (torch.float32, torch.float64)
    if x.dim() == 4:
        B, Cc, W, H = x.shape
    else:
        (B, Cc), W, H = x.shape, 1, 1
    if out is None:
        out = torch.empty_like(x)
    fn = _ref_gpu.tq_refgpu_tr_f32 if x.dtype == torch.float32 else _ref_gpu.tq_refgpu_tr_f64
    with torch.cuda.device(x.device):
        rc = fn(x.data_ptr(), out.data_ptr(), float(np.float32(sf)), bits, g, alpha, B, Cc, W, H,
                torch.cuda.current_stream(x.device).cuda_stream)
    if rc != 0:
        raise ValueError(f"reference GPU kernel: launch refused / failed (rc={rc})")
    return out


def ref_hese_terms(x, sf, bits):
    buf = (C.c_int32 * 64)()
    n = ref().tq_ref_hese_terms_f32(float(x), float(np.float32(sf)), bits, buf)
    return [buf[i] for i in range(n)]


def terms(q, encoding=ENC_HESE):
    p, n = C.c_uint32(), C.c_uint32()
    lib().tqo_terms(int(q), encoding, C.byref(p), C.byref(n))
    return p.value, n.value


def quantize(x, sf, bits):
    return lib().tqo_quantize_f32(float(np.float32(x)), float(np.float32(sf)), bits)


def hist(x, hist_bins, lo, hi):
    x = np.ascontiguousarray(x, dtype=np.float32).ravel()
    assert hist_bins.dtype == np.float32 and hist_bins.flags.c_contiguous
    lib().tqo_hist_f32(x.ctypes.data, x.size, hist_bins.ctypes.data, hist_bins.size,
                       float(lo), float(hi))
    return hist_bins


def mse_profile(hist_bins, x, sfs, bits, terms_):
    hist_bins = np.ascontiguousarray(hist_bins, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    sfs = np.ascontiguousarray(sfs, dtype=np.float32)
    errs = np.empty(sfs.size, dtype=np.float64)
    idx = lib().tqo_mse_profile(hist_bins.ctypes.data, x.ctypes.data, x.size, sfs.ctypes.data,
                                sfs.size, bits, terms_, errs.ctypes.data)
    return idx, errs


def hese_term_count(w, sf):
    w = np.ascontiguousarray(w, dtype=np.float32).ravel()
    return lib().tqo_hese_term_count_f32(w.ctypes.data, w.size, float(np.float32(sf)))


def gemm_i32(a, w):
    a = np.ascontiguousarray(a, dtype=np.int16)
    w = np.ascontiguousarray(w, dtype=np.int16)
    M, K = a.shape
    N, K2 = w.shape
    assert K == K2
    acc = np.empty((M, N), dtype=np.int32)
    lib().tqo_gemm_i32(a.ctypes.data, w.ctypes.data, acc.ctypes.data, M, N, K)
    return acc


def fma_channels(x, a, b):
    """fmaf(x, a[c], b[c]) over the last (channel) axis of a float32 array, one rounding per element."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    assert x.shape[-1] == a.size == b.size
    y = np.empty_like(x)
    lib().tqo_fma_channels_f32(x.ctypes.data, a.ctypes.data, b.ctypes.data, y.ctypes.data, x.size, a.size)
    return y
