"""TEST / BENCHMARK INFRASTRUCTURE -- not part of the product path.

Recipe that puts an UNMODIFIED copy of the reference (cbg-ethz/BnpC v0.2.1) where the GPU box can
see it.  `/root/reference` exists only in the build container; `baseline/_ref/` is git-ignored
(the copy never enters the history) but travels with `gpurun`, like the built `.so`.

    python oracle/fetch_ref.py            # copies /root/reference -> baseline/_ref/BnpC

The reference has no packaging metadata (no setup.py / pyproject.toml), so `pip install --target
baseline/_ref /root/reference` has nothing to install: this script is that install step.  It
copies files byte for byte and records their SHA-256 in baseline/_ref/MANIFEST.json; nothing is
edited.  Consumers: `oracle/ref_shim.py` (imports the reference's modules with a numpy stand-in
for the missing `bottleneck`), `bench.py --impl reference` (times the reference's own classes on
the box's host cores) and `tests/test_oracle_vs_reference.py`.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get('BNPC_REFERENCE_SRC', '/root/reference')
DST = os.path.join(ROOT, 'baseline', '_ref', 'BnpC')
FILES = ['run_BnpC.py', 'requirements.txt', 'LICENSE', 'libs/__init__.py', 'libs/CRP.py',
         'libs/CRP_learning_errors.py', 'libs/MCMC.py', 'libs/utils.py', 'libs/dpmmIO.py', 'libs/plotting.py',
         'example_data/data.csv', 'example_data/data_params.txt']


def fetch(verbose=True):
    """Copy the reference tree; returns the destination or None when the source is absent (the GPU
    box: the copy made in the build container is used as it is)."""
    if not os.path.isfile(os.path.join(SRC, 'libs', 'CRP.py')):
        return DST if os.path.isfile(os.path.join(DST, 'libs', 'CRP.py')) else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, 'rb') as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(os.path.dirname(DST), 'MANIFEST.json'), 'w') as f:
        json.dump(dict(source='cbg-ethz/BnpC v0.2.1 (unmodified copy of /root/reference)', sha256=manifest), f,
                  indent=1)
    if verbose:
        print(f'reference copied to {DST} ({len(FILES)} files)')
    return DST


if __name__ == '__main__':
    sys.exit(0 if fetch() else 1)
