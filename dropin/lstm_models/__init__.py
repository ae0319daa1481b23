"""Drop-in alias: `import lstm_models` resolves to term_quantization_b200.lstm_models (put dropin/ on sys.path)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
import term_quantization_b200.lstm_models as _impl  # noqa: E402

_sys.modules[__name__] = _impl
