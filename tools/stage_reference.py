#!/usr/bin/env python
"""Stage the files of the reference that the reference arm drives (baseline/refarm.py) into the git-ignored
``baseline/_ref/``: model.py, model_manager.py, swap_batch_transform.py, utils.py and the demo fixtures.  Nothing is
modified; nothing under baseline/_ref is committed (it travels to the GPU box with the gpurun snapshot).

    python tools/stage_reference.py [--ref /root/reference]
"""
import argparse
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import refarm      # noqa: E402


def stage(ref='/root/reference', quiet=False):
    if not os.path.exists(os.path.join(ref, 'model.py')):
        return None
    dst = refarm.STAGED
    os.makedirs(os.path.join(dst, 'demo_files'), exist_ok=True)
    for f in refarm.FILES:
        shutil.copy2(os.path.join(ref, f), os.path.join(dst, f))
    for f in refarm.DEMO:
        shutil.copy2(os.path.join(ref, 'demo_files', f), os.path.join(dst, 'demo_files', f))
    mdst = os.path.join(dst, 'demo_files', 'meshes')
    if os.path.isdir(mdst):
        shutil.rmtree(mdst)
    shutil.copytree(os.path.join(ref, 'demo_files', 'meshes'), mdst)
    if not quiet:
        n = sum(len(fs) for _, _, fs in os.walk(dst))
        print('staged %d reference files into %s' % (n, dst))
    return dst


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--ref', default='/root/reference')
    a = ap.parse_args()
    if stage(a.ref) is None:
        raise SystemExit('no reference at ' + a.ref)
