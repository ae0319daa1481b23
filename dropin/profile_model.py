"""Drop-in alias: `import profile_model` resolves to term_quantization_b200.profile_model (put dropin/ on sys.path)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import term_quantization_b200.profile_model as _impl  # noqa: E402

_sys.modules[__name__] = _impl
