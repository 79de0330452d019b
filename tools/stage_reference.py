"""Stage the reference's hot-path sources for the GPU box -- TEST / BASELINE INFRASTRUCTURE.

/root/reference exists only in the build container.  The GPU box receives a snapshot of
this repository, including git-ignored directories, so the files the lifted reference
needs (oracle/reference_lift.py: `render` out of run.py by AST, the reference Generator,
its synthesis network, nerf_utils, ops, pose_utils) are copied UNMODIFIED into
`baseline/_ref/` -- git-ignored (never part of this repository's history), not
gpurun-ignored (travels with the snapshot).  SURVEY.md section 8c; the reference has no
setup.py / pyproject.toml, so `pip install --target baseline/_ref` is not possible and a
file-level staging of the Apache-2.0 sources is what "install" means here.

Used by: `__graft_entry__.build()` (when /root/reference is present), then on the box by
`oracle/reference_lift.py` -> tests marked gpu that compare `render()` against the
reference running eagerly on the same GPU, and `bench.py --impl reference`
(`cpu_baseline.kind == "reference"`).
"""
import hashlib
import json
import os
import shutil
import sys

FILES = ('run.py', 'LICENSE', 'lib/nerf_utils.py', 'lib/ops.py', 'lib/pose_utils.py',
         'models/generator.py', 'models/stylegan.py')

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, 'baseline', '_ref')


def stage(src='/root/reference', dest=DEST, quiet=False):
    """Copies FILES from ``src`` to ``dest``; returns the manifest (path -> sha256)."""
    if not os.path.isfile(os.path.join(src, 'run.py')):
        raise FileNotFoundError('no reference checkout at %s' % src)
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(dest, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        with open(d, 'rb') as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(dest, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': src, 'files': manifest}, f, indent=1, sort_keys=True)
    if not quiet:
        print('staged %d reference files into %s' % (len(manifest), dest))
    return manifest


if __name__ == '__main__':
    stage(*(sys.argv[1:2] or ['/root/reference']))
