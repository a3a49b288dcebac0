"""Recipe for oracle/_ref: the UNMODIFIED reference package, placed where the GPU box can run it.
TEST INFRASTRUCTURE ONLY (see oracle/esm_oracle.py for who may use oracle/).

The reference (`/root/reference/esme`) is pure Python: there is nothing to compile, "building" it means copying
the package directory byte for byte into `oracle/_ref/esme/`.  `oracle/_ref/` is git-ignored (no reference source
enters the history) but NOT gpurun-ignored, so it travels to the GPU box next to the built libesmk.so, where
`/root/reference` does not exist.  `__graft_entry__.build()` calls build_ref() whenever /root/reference is present.

The copied package is only ever run out of process, by oracle/ref_runner.py, with oracle/ref_shims/ (stubs for
the two absent third-party imports `accelerate` and `torchmetrics`) on its PYTHONPATH:
  * on the GPU: as is -- real flash_attn_varlen_func (flash-attn 2.8.3 wheel of the image): the parity target;
  * on the CPU: with the one symbol esme.attention.flash_attn_varlen_func replaced by a per-sequence torch SDPA
    (the reference has no CPU attention path): the `--impl reference` / cpu_baseline arm of bench.py.
"""
import filecmp
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference/esme'
REF_DST = os.path.join(HERE, '_ref', 'esme')


def ref_available() -> bool:
    return os.path.isfile(os.path.join(REF_DST, 'esm.py'))


def build_ref(force: bool = False) -> bool:
    """Copy the reference package into oracle/_ref/esme (True if oracle/_ref is usable afterwards)."""
    if not os.path.isdir(REF_SRC):
        return ref_available()
    os.makedirs(REF_DST, exist_ok=True)
    for name in sorted(os.listdir(REF_SRC)):
        if not name.endswith('.py'):
            continue
        src, dst = os.path.join(REF_SRC, name), os.path.join(REF_DST, name)
        if force or not os.path.isfile(dst) or not filecmp.cmp(src, dst, shallow=False):
            shutil.copyfile(src, dst)
    with open(os.path.join(HERE, '_ref', 'PROVENANCE'), 'w') as f:
        f.write('byte-for-byte copy of /root/reference/esme/*.py made by oracle/build_ref.py; not tracked by git\n')
    return ref_available()


if __name__ == '__main__':
    print('oracle/_ref available:', build_ref())
