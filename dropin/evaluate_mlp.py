"""Drop-in alias: `import evaluate_mlp` resolves to term_quantization_b200.evaluate_mlp (put dropin/ on sys.path)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import term_quantization_b200.evaluate_mlp as _impl  # noqa: E402

_sys.modules[__name__] = _impl
