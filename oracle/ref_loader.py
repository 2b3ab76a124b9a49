"""TEST / BASELINE INFRASTRUCTURE - loads the UNMODIFIED reference classes (lbusellato/KMP_demos) for the CPU arm of
bench.py and for tests that compare against the live reference.  Never imported by the product (kmp_demos_b200/, kmp/).

The reference is not an installable distribution (no setup.py / pyproject.toml: `pip install --target baseline/_ref
/root/reference` answers "Directory is not installable"), so `__graft_entry__.build()` places a verbatim, git-ignored
copy of its `kmp/` package under baseline/_ref/ when /root/reference is mounted (the authoring container); that copy
travels to the GPU box with the snapshot.  Nothing of it is tracked in git.

`import kmp.<anything>` fails for the reference as shipped (kmp/__init__.py imports matplotlib, kmp/types/point.py
breaks on Python >= 3.11) and would collide with this repo's own drop-in `kmp` package.  The four
hot-path files are self-contained modules (numpy / scipy / sklearn imports only), so they are executed from their
file paths under private module names:

    kmp/model/KMP.py                         -> KMP
    kmp/mixture/GaussianMixtureModel.py      -> GaussianMixtureModel
    kmp/cluster/Cluster.py                   -> KMeans, Uniform
    kmp/types/quaternion.py                  -> quaternion
"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_FILES = {
    'KMP': ('model', 'KMP.py'),
    'GaussianMixtureModel': ('mixture', 'GaussianMixtureModel.py'),
    'Cluster': ('cluster', 'Cluster.py'),
    'quaternion': ('types', 'quaternion.py'),
}


def find_reference():
    """Directory that holds the reference's `kmp/` package, or None."""
    for cand in (os.environ.get('KMP_REFERENCE'), os.path.join(ROOT, 'baseline', '_ref'), '/root/reference'):
        if cand and os.path.isfile(os.path.join(cand, 'kmp', 'model', 'KMP.py')):
            return cand
    return None


def available():
    return find_reference() is not None


def load_module(name):
    """Executes one of the reference's hot-path files and returns the module object."""
    ref = find_reference()
    if ref is None:
        raise ImportError('the reference is neither under baseline/_ref nor /root/reference')
    sub, fn = _FILES[name]
    key = f'_kmp_reference_{name}'
    if key in sys.modules:
        return sys.modules[key]
    spec = importlib.util.spec_from_file_location(key, os.path.join(ref, 'kmp', sub, fn))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[key] = mod
    spec.loader.exec_module(mod)
    return mod


def classes():
    """dict(KMP=..., GaussianMixtureModel=..., KMeans=..., Uniform=..., quaternion=...) of the unmodified reference."""
    cl = load_module('Cluster')
    return dict(KMP=load_module('KMP').KMP, GaussianMixtureModel=load_module('GaussianMixtureModel').GaussianMixtureModel,
                KMeans=cl.KMeans, Uniform=cl.Uniform, quaternion=load_module('quaternion').quaternion)


def install(src='/root/reference'):
    """Copies the reference's kmp/ package (python files only) to baseline/_ref/kmp.  Returns the destination or None
    when the source is not mounted."""
    import shutil
    if not os.path.isfile(os.path.join(src, 'kmp', 'model', 'KMP.py')):
        return None
    dst = os.path.join(ROOT, 'baseline', '_ref')
    shutil.copytree(os.path.join(src, 'kmp'), os.path.join(dst, 'kmp'), dirs_exist_ok=True,
                    ignore=shutil.ignore_patterns('*.pyc', '__pycache__'))
    return dst
