"""Forward-hook op/parameter counter with the API of the reference's thop fork
(thop/__init__.py:1-2): ``profile``, ``clever_format`` and the ``count_hooks`` functions.
Imports cleanly on Python 3.12 (the fork's ``collections.Iterable`` / ``distutils`` do not)."""
from . import count_hooks  # noqa: F401
from .profile import profile, register_hooks  # noqa: F401
from .utils import clever_format  # noqa: F401
