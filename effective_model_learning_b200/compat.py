"""
Module-name compatibility with the reference's flat script directory (scope row N3).

The reference's analysis scripts do `import learning_model`, `import configs`, ... and unpickle `*_best.pickle` /
`*_proposals.pickle` files whose classes were pickled as `learning_model.LearningModel`,
`two_level_system.TwoLevelSystem` (reference find_clusters.py:71-90, cluster_accepted.py:42-58,
models_params_corr_mat.py:38-56, collate.py:122-128).  Two helpers make this package a drop-in there:

  install_reference_aliases()   registers this package's modules under the reference's flat names, so that an
                                unmodified analysis script imports / unpickles against this package
                                (`model.build_Liouvillian()`, `.vectorise_under_library()`, `.final_loss`, ... are
                                the same calls);
  reference_module_names()      context manager: objects pickled inside it carry the reference's module names
                                (`learning_model.LearningModel`), i.e. the files look like the reference's own and load
                                wherever those names resolve -- to this package through the aliases, or to any other
                                implementation of the same classes.  `output.Output(reference_module_names=True)` uses it.
"""
from __future__ import annotations

import contextlib
import importlib
import sys

FLAT_MODULES = ('definitions', 'two_level_system', 'basic_model', 'learning_model', 'learning_chain',
                'process_handling', 'params_handling', 'configs', 'output')
_PICKLED_CLASSES = (('two_level_system', 'TwoLevelSystem'), ('basic_model', 'BasicModel'), ('learning_model', 'LearningModel'))


def install_reference_aliases(overwrite: bool = False) -> list[str]:
    """sys.modules[<flat name>] = effective_model_learning_b200.<flat name>; returns the names that were installed."""
    done = []
    for name in FLAT_MODULES:
        if name in sys.modules and not overwrite:
            if not getattr(sys.modules[name], '__name__', '').startswith(__package__ + '.'):
                continue        # somebody else's module of that name (e.g. the real reference): leave it alone
        sys.modules[name] = importlib.import_module(f'{__package__}.{name}')
        done.append(name)
    return done


@contextlib.contextmanager
def reference_module_names():
    """Pickle this package's model classes under the reference's module names for the duration of the block."""
    saved_modules = {name: sys.modules.get(name) for name, _ in _PICKLED_CLASSES}
    classes = []
    try:
        for mod_name, cls_name in _PICKLED_CLASSES:
            mod = importlib.import_module(f'{__package__}.{mod_name}')
            cls = getattr(mod, cls_name)
            classes.append((cls, cls.__module__))
            sys.modules[mod_name] = mod          # pickle verifies that <module>.<name> is the object being pickled
            cls.__module__ = mod_name
        yield
    finally:
        for cls, original in classes:
            cls.__module__ = original
        for name, mod in saved_modules.items():
            if mod is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = mod
