"""term_quantization_b200 -- B200 (sm_100a) implementation of the term-quantization hot path.

Layout
------
csrc/            hand-written CUDA kernels + the C ABI (include/tq_b200.h) -> libtq_b200.so
_lib.py          ctypes binding of that C ABI (fails loudly if the library is missing)
tr_cuda.py       drop-in for the reference extension module `tr_cuda` (kernels/tr_cuda.cpp)
tr_layer.py      drop-in for the reference module `tr_layer` (LinearQuantize, TR*Layer, ...)
profile_model.py / thop/ / cnn_models/ / lstm_models/ / train_mlp.py / util.py /
evaluate_*.py    host-side mirrors of the reference drivers on synthetic data

`dropin/` at the repository root re-exports these under the reference's top-level module
names, so `import tr_layer` keeps working for code written against the reference.
"""
__version__ = "0.1.0"
