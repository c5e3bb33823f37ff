"""Make the UNMODIFIED reference importable in this container (no GPU box use).

Usage (only from oracle/make_golden.py and the optional reference
cross-check tests, which skip when /root/reference is absent):

    from oracle import refboot; refboot.boot()
    import pb_chime5.core, pb_bss.distribution

Recipe = SURVEY.md appendix B: numpy alias shims, two ``sys.modules`` aliases,
dummy modules for two Cython helpers that are not on the numeric path, and the
stub packages in ``oracle/refshim`` for the un-vendored third-party deps.
"""
import os
import sys
import types
from pathlib import Path

REFERENCE = Path(os.environ.get('GSS_REFERENCE_ROOT', '/root/reference'))


def available():
    return (REFERENCE / 'pb_chime5' / 'core.py').exists()


def boot():
    import numpy as np
    if not available():
        raise RuntimeError(f'reference tree not found at {REFERENCE}')
    sys.dont_write_bytecode = True           # the tree is read-only
    for name, val in (('int', int), ('object', object), ('complex', complex),
                      ('float', float), ('bool', bool)):
        if name not in np.__dict__:
            setattr(np, name, val)
    if not hasattr(np, 'asfarray'):
        np.asfarray = lambda a, dtype=np.float64: np.asarray(a, dtype=dtype)
    if not hasattr(np.linalg, 'linalg'):
        np.linalg.linalg = np.linalg
    import sklearn.mixture._gaussian_mixture as _gm
    sys.modules.setdefault('sklearn.mixture.gaussian_mixture', _gm)
    import numpy.testing._private.utils as _tu
    sys.modules.setdefault('numpy.testing.utils', _tu)
    for mod, names in (
        ('pb_chime5.utils.intervall_array_util',
         ('cy_non_intersection', 'cy_intersection', 'cy_parse_item',
          'cy_str_to_intervalls')),
        ('pb_chime5.utils.alignment_util', ('cy_alignment_id2phone',)),
    ):
        m = types.ModuleType(mod)
        for n in names:
            setattr(m, n, None)
        sys.modules.setdefault(mod, m)
    here = Path(__file__).resolve().parent
    for p in (str(REFERENCE / 'pb_bss'), str(REFERENCE), str(here / 'refshim'),
              str(here.parent)):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
