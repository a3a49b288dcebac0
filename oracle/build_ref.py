"""Recipe for oracle/_ref: the UNMODIFIED reference package, built into a form the GPU box can run.
TEST INFRASTRUCTURE ONLY (see oracle/esm_oracle.py for who may use oracle/).

The reference (`/root/reference/esme`) is pure Python, so its "build" is byte-compilation: every module of the
package is compiled where it lies (`py_compile`, no source is copied or edited) and the resulting `.pyc` files are
packed into ONE archive, `oracle/_ref/esme_ref.zip`, which Python imports directly (zipimport).  `oracle/_ref/` is
git-ignored (nothing derived from the reference enters the history) but NOT gpurun-ignored, so the archive travels
to the GPU box next to the built libesmk.so, where `/root/reference` does not exist.  `__graft_entry__.build()`
calls build_ref() whenever /root/reference is present.

The archive is only ever used out of process, by oracle/ref_runner.py, with oracle/ref_shims/ (stubs for the two
absent third-party imports `accelerate` and `torchmetrics`) on its path:
  * on the GPU: as is -- real flash_attn_varlen_func (flash-attn 2.8.3 wheel of the image): the parity target;
  * on the CPU: with the one symbol esme.attention.flash_attn_varlen_func replaced by a per-sequence torch SDPA
    (the reference has no CPU attention path): the `--impl reference` / cpu_baseline arm of bench.py.
"""
import hashlib
import os
import py_compile
import sys
import tempfile
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference/esme'
REF_ZIP = os.path.join(HERE, '_ref', 'esme_ref.zip')


def ref_available() -> bool:
    return os.path.isfile(REF_ZIP)


def _source_digest() -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(REF_SRC)):
        if name.endswith('.py'):
            h.update(name.encode())
            h.update(open(os.path.join(REF_SRC, name), 'rb').read())
    h.update(sys.version.encode())
    return h.hexdigest()


def build_ref(force: bool = False) -> bool:
    """Byte-compile the reference package into oracle/_ref/esme_ref.zip (True if the archive is usable afterwards)."""
    if not os.path.isdir(REF_SRC):
        return ref_available()
    os.makedirs(os.path.dirname(REF_ZIP), exist_ok=True)
    stamp = os.path.join(os.path.dirname(REF_ZIP), 'PROVENANCE')
    digest = _source_digest()
    if not force and ref_available() and os.path.isfile(stamp) and digest in open(stamp).read():
        return True
    with tempfile.TemporaryDirectory() as tmp, zipfile.ZipFile(REF_ZIP + '.tmp', 'w', zipfile.ZIP_DEFLATED) as z:
        for name in sorted(os.listdir(REF_SRC)):
            if not name.endswith('.py'):
                continue
            cfile = os.path.join(tmp, name + 'c')
            py_compile.compile(os.path.join(REF_SRC, name), cfile=cfile, dfile=f'<reference>/esme/{name}', doraise=True)
            z.write(cfile, f'esme/{name}c')
    os.replace(REF_ZIP + '.tmp', REF_ZIP)
    with open(stamp, 'w') as f:
        f.write('esme_ref.zip: /root/reference/esme/*.py byte-compiled (py_compile) by oracle/build_ref.py with '
                f'{sys.version.split()[0]}; not tracked by git\nsha256(sources + interpreter) = {digest}\n')
    return ref_available()


if __name__ == '__main__':
    print('oracle/_ref available:', build_ref(force='--force' in sys.argv))
