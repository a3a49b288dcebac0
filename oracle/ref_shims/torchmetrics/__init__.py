"""Stub of `torchmetrics` (not installed): esme/variant.py:5 imports `torchmetrics.text.Perplexity` at module
level; only predict_pseudoperplexity uses it.  TEST INFRASTRUCTURE."""
from . import text  # noqa: F401
