"""Import the UNMODIFIED reference (numga) in this container, for fixture generation only.

The reference at /root/reference is pure Python but does not import as-is here
(SURVEY.md §8c): it needs the third-party `numpy_indexed` (absent, no network) and
one line trips numpy-2 integer-overflow rules.  This loader
  * copies /root/reference/numga to a scratch dir under /tmp (never into the repo),
  * rewrites the dead `blades * -1` expression at subspace/factory.py:50
    (its value is overwritten two lines later, so semantics are unchanged),
  * installs a tiny `numpy_indexed` stand-in (integer set/sort helpers only; no
    floating-point arithmetic passes through it),
and returns the imported `numga` package.

Nothing here runs on the GPU box: /root/reference does not exist there.  Only
`make_golden.py` (which writes tests/golden/*.npz) uses this module.
"""
import os
import shutil
import sys
import tempfile
import types

import numpy as np

REFERENCE = '/root/reference'


def _numpy_indexed_standin():
    npi = types.ModuleType('numpy_indexed')

    def all_unique(a):
        a = np.asarray(a)
        return len(np.unique(a)) == len(a)

    def indices(this, that, missing='raise'):
        this = np.asarray(this)
        that = np.asarray(that)
        order = np.argsort(this, kind='stable')
        pos = np.searchsorted(this[order], that)
        pos_c = np.clip(pos, 0, max(len(this) - 1, 0))
        found = (this[order][pos_c] == that) if len(this) else np.zeros(that.shape, bool)
        if missing == 'raise':
            if not np.all(found):
                raise KeyError('Not all keys in `that` are present in `this`')
            return order[pos_c]
        out = np.where(found, order[pos_c] if len(this) else 0, missing)
        return out

    class GroupBy:
        def __init__(self, keys):
            self.keys = np.asarray(keys)
            self.unique, self.inverse = np.unique(self.keys, return_inverse=True)

        def split(self, values):
            values = np.asarray(values)
            return [values[self.inverse == i] for i in range(len(self.unique))]

        def sum(self, values):
            values = np.asarray(values)
            out = np.zeros((len(self.unique),) + values.shape[1:], values.dtype)
            np.add.at(out, self.inverse, values)
            return self.unique, out

    def group_by(keys, values=None):
        g = GroupBy(keys)
        if values is None:
            return g
        return g.unique, g.split(values)

    npi.all_unique = all_unique
    npi.indices = indices
    npi.group_by = group_by
    return npi


_cached = None


def load_reference():
    """Return the imported reference `numga` package (patched copy under /tmp)."""
    global _cached
    if _cached is not None:
        return _cached
    if not os.path.isdir(REFERENCE):
        raise RuntimeError('reference not present (expected on the GPU box); golden fixtures are committed')
    root = tempfile.mkdtemp(prefix='numga_ref_')
    shutil.copytree(os.path.join(REFERENCE, 'numga'), os.path.join(root, 'numga'))
    fac = os.path.join(root, 'numga', 'subspace', 'factory.py')
    src = open(fac).read()
    assert 'blades * -1' in src
    src = src.replace('blades * -1', 'blades.astype(np.int64) * -1')
    open(fac, 'w').write(src)
    sys.modules['numpy_indexed'] = _numpy_indexed_standin()
    sys.path.insert(0, root)
    import numga  # noqa
    _cached = numga
    return numga
