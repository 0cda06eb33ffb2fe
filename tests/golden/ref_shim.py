"""Import shim for running the *live* reference (/root/reference) in the build container.

Only used by tests/golden/make_golden.py (golden-vector generation); never imported at test/bench time
on the GPU box, where /root/reference does not exist.

Two shims, both outside the reference tree:
  * numpy.float = float            (mem_dataset.py:150 uses the alias removed in numpy 1.24)
  * permissive stub modules for tensorflow / matplotlib (DRecPy/__init__ eagerly imports Recommender -> tf,
    Evaluation -> loss_tracker -> matplotlib). None of the golden paths execute TF code.
"""
import sys
import types
import importlib.machinery

REFERENCE_ROOT = '/root/reference'


class _Stub(types.ModuleType):
    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []
        self.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)

    def __getattr__(self, item):
        if item.startswith('__'):
            raise AttributeError(item)
        full = f'{self.__name__}.{item}'
        mod = sys.modules.get(full)
        if mod is None:
            mod = _Stub(full)
            sys.modules[full] = mod
        return mod

    def __call__(self, *a, **k):
        return self


def install():
    import os
    import tempfile
    import warnings
    import numpy as np
    warnings.filterwarnings('ignore', category=SyntaxWarning)
    # keep the reference's data folder out of $HOME (file_utils.py:6 reads DATA_FOLDER)
    os.environ.setdefault('DATA_FOLDER', os.path.join(tempfile.gettempdir(), 'drecpy_ref_data'))
    if not hasattr(np, 'float'):
        np.float = float
    for name in ('tensorflow', 'matplotlib', 'matplotlib.pyplot'):
        if name not in sys.modules:
            sys.modules[name] = _Stub(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
