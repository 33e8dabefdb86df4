"""xformers stand-in: only `import xformers` must succeed (xformer_attention.py:3); ops are never called
because diffusers.utils.import_utils.is_xformers_available() is False in the shim (math path, attention.py:461)."""
from . import ops  # noqa: F401
