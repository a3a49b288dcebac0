"""Stub of the `accelerate` package for running the UNMODIFIED reference (oracle/_ref/esme_ref.zip).  TEST INFRASTRUCTURE.

`accelerate` is not installed in this image.  The reference imports it at module level (esme/esm.py:3) and uses
two names, both in its checkpoint loader only (esme/esm.py:363-372) -- no arithmetic goes through it:
  * init_empty_weights()           -> here: a null context (parameters are allocated normally)
  * load_checkpoint_and_dispatch() -> here: load_state_dict from the safetensors file, then .to(device)
"""
import contextlib

init_empty_weights = contextlib.nullcontext


def load_checkpoint_and_dispatch(model, checkpoint, device_map=None, **kw):
    from safetensors.torch import load_file
    missing, unexpected = model.load_state_dict(load_file(checkpoint), strict=True)
    assert not missing and not unexpected
    device = (device_map or {'': 'cpu'})['']
    return model.to(device)
