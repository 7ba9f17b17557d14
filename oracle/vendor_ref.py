"""
ORACLE (test infrastructure) -- vendors the UNMODIFIED reference package into
``oracle/_ref/`` so that the CPU arm of ``bench.py`` can time the reference itself on
the GPU box, where ``/root/reference`` does not exist.

    python oracle/vendor_ref.py [/root/reference]

``oracle/_ref/`` is git-ignored (no reference sources enter the history) but not
gpurun-ignored, so it travels with the snapshot.  ``__graft_entry__.build()`` runs this
whenever the reference tree is present.  Nothing under ``qspectra_b200/`` imports it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, '_ref')


def vendor(source='/root/reference'):
    pkg = os.path.join(source, 'qspectra')
    if not os.path.isdir(pkg):
        return False
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    shutil.copytree(pkg, os.path.join(DEST, 'qspectra'),
                    ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    for name in ('LICENSE', 'README.md'):
        if os.path.exists(os.path.join(source, name)):
            shutil.copy(os.path.join(source, name), os.path.join(DEST, name))
    return True


def load():
    """Import the vendored reference (``import qspectra``) with the one shim Python 3.11+
    needs (``inspect.getargspec``, used at simulate/decorators.py:30-36).  Returns the
    module, or None if ``oracle/_ref`` has not been vendored."""
    if not os.path.isdir(os.path.join(DEST, 'qspectra')):
        return None
    import collections
    import inspect
    if not hasattr(inspect, 'getargspec'):
        ArgSpec = collections.namedtuple('ArgSpec', 'args varargs keywords defaults')

        def getargspec(func):
            full = inspect.getfullargspec(func)
            return ArgSpec(full.args, full.varargs, full.varkw, full.defaults)
        inspect.getargspec = getargspec
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')      # SyntaxWarnings of '\s' in the reference docstrings
        import qspectra
    return qspectra


if __name__ == '__main__':
    ok = vendor(sys.argv[1] if len(sys.argv) > 1 else '/root/reference')
    print('vendored' if ok else 'reference tree not found', '->', DEST)
